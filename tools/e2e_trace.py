import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import suggest_b200 as S
from suggest_b200.suggest import IndexDescription
from suggest_b200.workload import synthetic_dictionary, synthetic_queries
NQ, K = 65536, 10
d_bytes, d_off, rng = synthetic_dictionary(1_000_000)
q_bytes, q_off, _ = synthetic_queries(d_bytes, d_off, NQ, rng)
desc = IndexDescription(Name="b", NGramSize=3, Alphabet=("english", "russian", "numbers", "$"), Pad="$", Wrap=("$", "$"))
index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
hq = torch.from_numpy(q_bytes).pin_memory(); hoff = torch.from_numpy(q_off.astype(np.int32)).pin_memory()
h_ids = torch.zeros(NQ * K, dtype=torch.int32).pin_memory(); h_sc = torch.zeros(NQ * K, dtype=torch.float64).pin_memory(); h_cnt = torch.zeros(NQ, dtype=torch.int32).pin_memory()
out = (h_ids.numpy().view(np.uint32).reshape(NQ, K), h_sc.numpy().reshape(NQ, K), h_cnt.numpy().view(np.uint32))
packed = (hq.numpy(), hoff.numpy().view(np.uint32))
for i in range(8):
    t0 = time.perf_counter()
    index.SuggestBatch(None, 0.5, S.JaccardMetric(), K, packed=packed, out=out)
    print("python call %.1f us" % ((time.perf_counter() - t0) * 1e6), flush=True)
