#!/usr/bin/env python3
"""End-to-end rate of BASELINE.json config #2 with T host threads calling sg_search_batch_candidates concurrently (the
reference's callers are one goroutine per request): every thread has its own page-locked query batches and result rows.
usage: e2e_callers.py [calls per thread]      (SG_DIRECT_CHUNKS=0 selects the sliced path)"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_dictionary, synthetic_queries  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nq, k = 65536, 10
d, off, rng = synthetic_dictionary(1_000_000)
ix = S.NewRAMBuilder((d, off), IndexDescription(Name="p", NGramSize=3)).Build()
m = S.JaccardMetric()
for T in (1, 2, 3, 4):
    batches = []
    for t in range(T):
        ring = []
        for _ in range(4):
            q, qo, _ = synthetic_queries(d, off, nq, rng)
            ring.append((torch.from_numpy(q).pin_memory(), torch.from_numpy(qo.astype(np.int32)).pin_memory()))
        batches.append((ring, S.PinnedCandidateRows(nq, k)))

    def worker(t, n):
        ring, rows = batches[t]
        for b in range(n):
            hq, ho = ring[b % len(ring)]
            ix.SuggestBatchCandidates(None, 0.5, m, k, packed=(hq.numpy(), ho.numpy().view(np.uint32)), out=rows.out)

    for t in range(T):
        worker(t, 8)
    th = [threading.Thread(target=worker, args=(t, calls)) for t in range(T)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    print(f"chunks={os.environ.get('SG_DIRECT_CHUNKS', 'default')} callers {T}: {T * calls * nq / dt / 1e6:.1f}M q/s, {dt / calls * 1e6:.0f} us per call per thread", flush=True)
    for _, rows in batches:
        rows.close()
