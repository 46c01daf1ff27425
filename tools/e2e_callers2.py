#!/usr/bin/env python3
"""what makes three concurrent callers slow inside bench.py but not in e2e_callers.py: toggles one condition at a time"""
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

nq, k, T, calls = 65536, 10, 3, 150
d, off, rng = synthetic_dictionary(1_000_000)
ix = S.NewRAMBuilder((d, off), IndexDescription(Name="p", NGramSize=3)).Build()
m = S.JaccardMetric()
shared = []
for _ in range(8):
    q, qo, _ = synthetic_queries(d, off, nq, rng)
    shared.append((torch.from_numpy(q).pin_memory(), torch.from_numpy(qo.astype(np.int32)).pin_memory()))
rows = [S.PinnedCandidateRows(nq, k) for _ in range(T)]


def run(label, stride_shared=True):
    def worker(t):
        for b in range(t, calls * T, T):
            hq, ho = shared[b % 8] if stride_shared else shared[(2 * t + (b // T) % 2) % 8]
            ix.SuggestBatchCandidates(None, 0.5, m, k, packed=(hq.numpy(), ho.numpy().view(np.uint32)), out=rows[t].out)
    for t in range(T):
        for b in range(4):
            ix.SuggestBatchCandidates(None, 0.5, m, k, packed=(shared[b][0].numpy(), shared[b][1].numpy().view(np.uint32)), out=rows[t].out)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    print(f"{label}: {T * calls * nq / dt / 1e6:.1f}M q/s", flush=True)


run("own batches per caller", stride_shared=False)
run("shared ring of 8")
# device-side activity of the bench before the e2e leg: device-resident calls on a side stream, a 256 MB flush buffer
side = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(side):
    flush.fill_(1)
torch.cuda.synchronize()
run("after a torch side stream was used")
stop = threading.Event()


def poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        time.sleep(0.002)


p = threading.Thread(target=poll, daemon=True)
p.start()
run("with the NVML polling thread")
stop.set()
p.join()
run("polling thread stopped again")
