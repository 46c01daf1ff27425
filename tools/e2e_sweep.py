#!/usr/bin/env python3
"""e2e (host buffers) throughput of sg_search_batch for several slice sizes, next to raw pinned-copy bandwidth.
usage (GPU box): python tools/e2e_sweep.py"""
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

# raw copies
dev = torch.device("cuda", 0)
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for n in (1 << 20, 8 << 20, 64 << 20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dst[:n].copy_(src[:n], non_blocking=True)
        torch.cuda.synchronize()
        print(f"{name} {n >> 20} MiB: {10 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)

hq = torch.from_numpy(q_bytes).pin_memory()
hoff = torch.from_numpy(q_off.astype(np.int32)).pin_memory()
h_ids = torch.zeros(NQ * K, dtype=torch.int32).pin_memory()
h_sc = torch.zeros(NQ * K, dtype=torch.float64).pin_memory()
h_cnt = torch.zeros(NQ, dtype=torch.int32).pin_memory()
out = (h_ids.numpy().view(np.uint32).reshape(NQ, K), h_sc.numpy().reshape(NQ, K), h_cnt.numpy().view(np.uint32))
packed = (hq.numpy(), hoff.numpy().view(np.uint32))
for sq in (4096, 8192, 16384, 32768, 65536):
    os.environ["SG_SLICE_QUERIES"] = str(sq)
    index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
    for _ in range(3):
        index.SuggestBatch(None, 0.5, S.JaccardMetric(), K, packed=packed, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        index.SuggestBatch(None, 0.5, S.JaccardMetric(), K, packed=packed, out=out)
    dt = (time.perf_counter() - t0) / 20
    print(f"slice {sq}: {dt * 1e3:.3f} ms/batch  {NQ / dt / 1e6:.1f} M q/s", flush=True)
    index.close()
