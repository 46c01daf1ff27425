"""Small run of the collect kernel (sg_candidates_batch, built-in and tabulated thresholds) and of the one-process shards
(sg_sharded_search_batch: peer reads and peer copies, page-locked and pageable rows) for compute-sanitizer.
usage: compute-sanitizer --tool memcheck python tools/sanitize_new.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import suggest_b200 as S  # noqa: E402
from suggest_b200.metric import Metric  # noqa: E402
from suggest_b200.sharding import ShardedNGramIndex  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_workload  # noqa: E402


class Containment(Metric):
    name = "Containment"

    def MinY(self, alpha, size):
        return int(np.ceil(alpha * size))

    def MaxY(self, alpha, size):
        return 1 << 15

    def Threshold(self, alpha, sizeA, sizeB):
        return int(np.ceil(alpha * sizeA))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / float(sizeA)


docs, (qb, qo), _ = synthetic_workload(20000, 20000)
desc = IndexDescription(Name="san", NGramSize=3)
for shift in (None, 0, 3):
    if shift is not None:
        os.environ["SG_BUCKET_SHIFT"] = str(shift)
    gx = S.NewRAMBuilder(docs, desc).Build()
    os.environ.pop("SG_BUCKET_SHIFT", None)
    for m, alpha in ((S.JaccardMetric(), 0.4), (S.CosineMetric(), 0.5), (Containment(), 0.8)):
        cq, cid, cov, cseg, size_a = gx.CandidatesBatch(None, alpha, m, packed=(qb, qo), cap=64)  # overflows once, then fits
        print("candidates", shift, m, len(cq))
    want = gx.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo))
    gx.close()
pinned = S.PinnedBuffers(len(qo) - 1, 10)
for copy in ("0", "1"):
    os.environ["SG_SHARD_GATHER_COPY"] = copy
    sx = ShardedNGramIndex(docs, desc, [0, 0, 0])
    for out in (None, pinned.out):
        ids, sc, cnt = sx.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo), out=out)
        m = np.arange(10)[None, :] < want[2][:, None]
        assert np.array_equal(cnt, want[2]) and np.array_equal(ids[m], want[0][m]) and np.array_equal(sc[m], want[1][m])
    print("sharded", "copies" if copy == "1" else "peer reads", int((cnt > 0).sum()))
    sx.close()
print("sanitize_new ok")
