#!/usr/bin/env python3
"""words.dict, count -> resolve pipeline forced (SG_PIPELINE=lean), 65,536 queries: device-buffer path against the host-buffer
paths (pageable rows, page-locked rows, page-locked candidate rows) and an oracle sample.  Debug aid (GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SG_PIPELINE", "lean")
import torch  # noqa: E402

import suggest_b200 as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_queries  # noqa: E402

lines = open(os.path.join(ROOT, "tests", "golden", "words.dict"), "rb").read().split(b"\n")[:-1]
off = np.zeros(len(lines) + 1, dtype=np.uint64)
off[1:] = np.cumsum([len(x) for x in lines])
data = np.frombuffer(b"".join(lines), dtype=np.uint8).copy()
nq, k = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 10
rng = np.random.default_rng(5)
q, qo, _ = synthetic_queries(data, off, nq, rng, 1)
ix = S.NewRAMBuilder((data, off), IndexDescription(Name="w", NGramSize=3, Wrap=("^", "$"), Pad="$", Alphabet=("english", "numbers", "$^"))).Build()
print(ix.layout(), flush=True)
m = S.CosineMetric()
dev = torch.device("cuda:0")
dq, doff = torch.from_numpy(q).to(dev), torch.from_numpy(qo.astype(np.int32)).to(dev)
res = {}
for rep in range(2):
    ids = torch.zeros(nq * k, dtype=torch.int32, device=dev)
    sc = torch.zeros(nq * k, dtype=torch.float64, device=dev)
    cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    ix.SuggestBatchDevice(dq.data_ptr(), doff.data_ptr(), nq, 0.5, m, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res[f"device{rep}"] = (ids.cpu().numpy().view(np.uint32).reshape(nq, k), sc.cpu().numpy().reshape(nq, k), cnt.cpu().numpy().view(np.uint32))
for _ in range(6):
    print(ix.StageTimes(dq.data_ptr(), doff.data_ptr(), nq, 0.5, m, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream), flush=True)
res["pageable"] = ix.SuggestBatch(None, 0.5, m, k, packed=(q, qo))
buf = S.PinnedBuffers(nq, k)
hq, hoff = torch.from_numpy(q).pin_memory(), torch.from_numpy(qo.astype(np.int32)).pin_memory()
for rep in range(2):
    ix.SuggestBatch(None, 0.5, m, k, packed=(hq.numpy(), hoff.numpy().view(np.uint32)), out=buf.out)
    res[f"pinned{rep}"] = (buf.ids.copy(), buf.scores.copy(), buf.counts.copy())
rows = S.PinnedCandidateRows(nq, k)
ix.SuggestBatchCandidates(None, 0.5, m, k, packed=(hq.numpy(), hoff.numpy().view(np.uint32)), out=rows.out)
res["rows"] = (rows.rows["key"].copy(), rows.rows["score"].copy(), rows.counts.copy())
ox = O.OracleIndex(3, ("^", "$"), "$", ("english", "numbers", "$^")).add_packed(data, off)
ns = nq
o = ox.suggest_batch(None, O.COSINE, 0.5, k, O.CANONICAL, threads=8, packed=(q[:int(qo[ns])], qo[:ns + 1].astype(np.uint64)))
ref = res["device0"]
for name, (i_, s_, c_) in res.items():
    mask = np.arange(k)[None, :] < ref[2][:, None]
    same_c = np.array_equal(c_, ref[2])
    bad = np.nonzero(c_ != ref[2])[0]
    same_i = same_c and np.array_equal(i_[mask], ref[0][mask]) and np.array_equal(s_[mask], ref[1][mask])
    mo = np.arange(k)[None, :] < o[2][:, None]
    vs_o = np.array_equal(c_[:ns], o[2]) and np.array_equal(i_[:ns][mo], o[0][mo]) and np.array_equal(s_[:ns][mo], o[1][mo])
    bad_o = np.nonzero(c_[:ns] != o[2])[0]
    print(f"{name}: equals device0: counts {same_c} rows {same_i}; equals oracle on {ns}: {vs_o}; count mismatches {len(bad)} first {bad[:8]} "
          f"{[(int(c_[b]), int(ref[2][b])) for b in bad[:8]]}; against the oracle {len(bad_o)}: "
          f"{[(int(b), int(c_[b]), int(o[2][b]), bytes(q[int(qo[b]):int(qo[b + 1])])) for b in bad_o[:8]]}", flush=True)
