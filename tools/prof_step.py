#!/usr/bin/env python3
"""A short device-resident run of BASELINE.json config #2 (or a variant) for ncu: index build, N calls of the path.
usage: prof_step.py [--calls N] [--metric Jaccard|Cosine|Dice] [--ngram n] [--data uniform|zipf] [--docs D]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_dictionary, synthetic_queries  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--calls", type=int, default=6)
ap.add_argument("--metric", default="Jaccard")
ap.add_argument("--ngram", type=int, default=3)
ap.add_argument("--data", default="uniform")
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--stages", action="store_true", help="print the per-kernel times of one call (SG_TRACE=1 adds the pipeline's counters)")
a = ap.parse_args()
d, off, rng = synthetic_dictionary(a.docs, skew=None if a.data == "uniform" else a.data)
q, qo, _ = synthetic_queries(d, off, 65536, rng)
ix = S.NewRAMBuilder((d, off), IndexDescription(Name="p", NGramSize=a.ngram)).Build()
m = {"Jaccard": S.JaccardMetric(), "Cosine": S.CosineMetric(), "Dice": S.DiceMetric()}[a.metric]
dev = torch.device("cuda:0")
dq, doff = torch.from_numpy(q).to(dev), torch.from_numpy(qo.astype(np.int32)).to(dev)
k, nq = 10, 65536
ids = torch.zeros(nq * k, dtype=torch.int32, device=dev)
sc = torch.zeros(nq * k, dtype=torch.float64, device=dev)
cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(a.calls):
    ix.SuggestBatchDevice(dq.data_ptr(), doff.data_ptr(), nq, 0.5, m, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), 0, st)
torch.cuda.synchronize()
if a.stages:
    acc = {}
    for _ in range(10):
        for name, ms in ix.StageTimes(dq.data_ptr(), doff.data_ptr(), nq, 0.5, m, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), st).items():
            acc[name] = acc.get(name, 0.0) + ms / 10
    print({k_: round(v, 4) for k_, v in acc.items()}, ix.layout(), flush=True)
print("queries with a match", float((cnt > 0).float().mean()), flush=True)
