#!/usr/bin/env python3
"""where to cut a 65,536-query batch (SG_DIRECT_SPLIT, percent) for sg_search_batch_candidates with page-locked rows: one
caller and three concurrent callers, config #2.  usage (GPU box): python tools/e2e_split_sweep.py"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_dictionary, synthetic_queries  # noqa: E402

NQ, K, T = 65536, 10, 3
d, off, rng = synthetic_dictionary(1_000_000)
ring = []
for _ in range(6):
    q, qo, _ = synthetic_queries(d, off, NQ, rng)
    ring.append((torch.from_numpy(q).pin_memory(), torch.from_numpy(qo.astype(np.int32)).pin_memory()))
rows = [S.PinnedCandidateRows(NQ, K) for _ in range(T)]
m = S.JaccardMetric()
for env in (dict(), dict(SG_DIRECT_SPLIT="12"), dict(SG_DIRECT_SPLIT="40"), dict(SG_DIRECT_SPLIT="50"), dict(SG_DIRECT_SPLIT="15,50"),
            dict(SG_DIRECT_SPLIT="33,66"), dict(SG_DIRECT_SPLIT="10,35,65"), dict(SG_DIRECT_SLICE_QUERIES=65536)):
    for k_, v in env.items():
        os.environ[k_] = str(v)
    ix = S.NewRAMBuilder((d, off), IndexDescription(Name="p", NGramSize=3)).Build()
    for k_ in env:
        os.environ.pop(k_)

    def call(t, b):
        hq, ho = ring[b % len(ring)]
        ix.SuggestBatchCandidates(None, 0.5, m, K, packed=(hq.numpy(), ho.numpy().view(np.uint32)), out=rows[t].out)

    def worker(t, n):
        for b in range(t, n * T, T):
            call(t, b)

    for b in range(6):
        call(0, b)
    t0 = time.perf_counter()
    for b in range(100):
        call(0, b)
    one = 100 * NQ / (time.perf_counter() - t0)
    for n in (20, 100):  # the first round creates the call contexts
        th = [threading.Thread(target=worker, args=(t, n)) for t in range(T)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        three = T * n * NQ / (time.perf_counter() - t0)
    print(f"{env or 'default (25)'}: one caller {one / 1e6:.1f} M q/s, three callers {three / 1e6:.1f} M q/s", flush=True)
    ix.close()
