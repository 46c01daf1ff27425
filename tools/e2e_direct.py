#!/usr/bin/env python3
"""e2e throughput of sg_search_batch with page-locked result buffers: rows staged in HBM and copied back per slice
(SG_DIRECT_OUT=0) against rows stored by the kernel straight into host memory, for several slice sizes.
usage (GPU box): python tools/e2e_direct.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_dictionary, synthetic_queries  # noqa: E402

NQ, K = 65536, 10
d_bytes, d_off, rng = synthetic_dictionary(1_000_000)
q_bytes, q_off, _ = synthetic_queries(d_bytes, d_off, NQ, rng)
desc = IndexDescription(Name="b", NGramSize=3, Alphabet=("english", "russian", "numbers", "$"), Pad="$", Wrap=("$", "$"))
torch.cuda.set_device(0)
hq = torch.from_numpy(q_bytes).pin_memory()
hoff = torch.from_numpy(q_off.astype(np.int32)).pin_memory()
packed = (hq.numpy(), hoff.numpy().view(np.uint32))
buf = S.PinnedBuffers(NQ, K)
ref = None
for metric in (S.JaccardMetric(), S.CosineMetric()):
    ref = None
    for env in (dict(SG_DIRECT_CHUNKS=4), dict(SG_DIRECT_CHUNKS=8), dict(SG_DIRECT_OUT=0), dict(SG_DIRECT_SLICE_QUERIES=65536), dict(SG_DIRECT_SLICE_QUERIES=32768),
                dict(SG_DIRECT_SLICE_QUERIES=16384), dict(), dict(SG_DIRECT_SPLIT="12,38"), dict(SG_DIRECT_SPLIT="8,30"),
                dict(SG_DIRECT_SPLIT="25"), dict(SG_DIRECT_SPLIT="12"), dict(SG_DIRECT_SPLIT="6,20,50"), dict(SG_DIRECT_SPLIT="15,50"),
                dict(SG_DIRECT_SPLIT="50"), dict(SG_DIRECT_SPLIT="33,66"), dict(SG_DIRECT_SPLIT="25,50,75"), dict(SG_DIRECT_SPLIT="40")):
        for k_, v in env.items():
            os.environ[k_] = str(v)
        index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
        for k_ in env:
            os.environ.pop(k_)
        buf.ids[...] = 0
        buf.scores[...] = 0
        for _ in range(5):
            index.SuggestBatch(None, 0.5, metric, K, packed=packed, out=buf.out)
        best = 1e9
        t0 = time.perf_counter()
        for _ in range(30):
            t1 = time.perf_counter()
            index.SuggestBatch(None, 0.5, metric, K, packed=packed, out=buf.out)
            best = min(best, time.perf_counter() - t1)
        dt = (time.perf_counter() - t0) / 30
        mask = np.arange(K)[None, :] < buf.counts[:, None]
        sig = (buf.counts.copy(), buf.ids[mask].copy(), buf.scores[mask].copy())
        same = True if ref is None else all(np.array_equal(a, b) for a, b in zip(ref, sig))
        ref = ref or sig
        print(f"{type(metric).__name__} {env}: {dt * 1e3:.3f} ms/batch (best {best * 1e3:.3f})  {NQ / dt / 1e6:.1f} M q/s  "
              f"results {'same' if same else 'DIFFER'}  rows written {int(mask.sum())}", flush=True)
        index.close()
