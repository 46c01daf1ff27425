"""SG_TRACE=1 timeline of sg_sharded_search_batch on every visible GPU (stderr).  usage: SG_TRACE=1 python tools/shard_capi_trace.py [n_docs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import suggest_b200 as S  # noqa: E402
from suggest_b200.sharding import ShardedNGramIndex  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_workload  # noqa: E402

import torch  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
docs, (qb, qo), _ = synthetic_workload(n_docs, 65536)
pinned = S.PinnedBuffers(65536, 10)
for n in sorted({1, 2, torch.cuda.device_count()}):
    if n > torch.cuda.device_count():
        continue
    sx = ShardedNGramIndex(docs, IndexDescription(Name="t", NGramSize=3), list(range(n)))
    print(f"==== {n} shard(s)", file=sys.stderr, flush=True)
    for _ in range(6):
        sx.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo), out=pinned.out)
    sx.close()
