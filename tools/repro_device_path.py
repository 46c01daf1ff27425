#!/usr/bin/env python3
"""Host-buffer path vs device-buffer path on a cars shard, repeated; prints any difference (debug aid, GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200.suggest import IndexDescription, pack_strings  # noqa: E402

lines = open(os.path.join(ROOT, "tests", "golden", "cars.dict"), "rb").read().split(b"\n")[:-1]
desc = IndexDescription(Name="t", NGramSize=3, Alphabet=("russian", "english", "numbers", "$"), Pad="$", Wrap=("$", "$"))
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 3700
sh = S.NewRAMBuilder(lines[lo:], desc, id_base=lo).Build()
q = lines[::4]
nq, k = len(q), 10
ids, sc, cnt = sh.SuggestBatch(q, 0.5, S.JaccardMetric(), k)
dev = torch.device("cuda:0")
data, off = pack_strings(q)
d_q = torch.from_numpy(data).to(dev)
d_off = torch.from_numpy(off.astype(np.int32)).to(dev)
for it in range(6):
    d_ids = torch.zeros(nq * k, dtype=torch.int32, device=dev)
    d_sc = torch.zeros(nq * k, dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    sh.SuggestBatchDevice(d_q.data_ptr(), d_off.data_ptr(), nq, 0.5, S.JaccardMetric(), k, d_ids.data_ptr(), d_sc.data_ptr(),
                          d_cnt.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g_sc = d_sc.cpu().numpy().reshape(nq, k)
    g_ids = d_ids.cpu().numpy().astype(np.uint32).reshape(nq, k)
    bad = np.nonzero((g_sc != sc) | (g_ids != ids))
    print("iter", it, "differences", len(bad[0]), [(int(a), int(b), float(g_sc[a, b]), float(sc[a, b]), q[a]) for a, b in zip(*bad)][:5], flush=True)
ids2, sc2, cnt2 = sh.SuggestBatch(q, 0.5, S.JaccardMetric(), k)
print("host again equal:", np.array_equal(sc2, sc), flush=True)
