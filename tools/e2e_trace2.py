#!/usr/bin/env python3
"""Device timeline of sg_search_batch per slice (SG_TRACE=2) for the staged and the direct result path.
usage (GPU box): SG_TRACE=2 python tools/e2e_trace2.py 2> trace.txt"""
import os
import sys

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
for env in (dict(SG_DIRECT_OUT=0), dict(), dict(SG_DIRECT_SPLIT="25")):
    for k_, v in env.items():
        os.environ[k_] = str(v)
    index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
    for k_ in env:
        os.environ.pop(k_)
    for i in range(6):
        if i == 4:
            os.write(2, f"==== {env}\n".encode())
        index.SuggestBatch(None, 0.5, S.JaccardMetric(), K, packed=packed, out=buf.out)
    index.close()
