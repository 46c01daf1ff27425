"""Config #4 through the single-process C ABI (sg_sharded_*): one host process, one shard per visible GPU, host buffers in
and out.  Prints one JSON line per gather mode.  usage: python tools/bench_sharded_capi.py [n_docs] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import suggest_b200 as S  # noqa: E402
from suggest_b200.sharding import ShardedNGramIndex  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_workload  # noqa: E402


def main():
    import torch
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n_gpus = torch.cuda.device_count()
    docs, (qb, qo), _ = synthetic_workload(n_docs, 65536)
    desc = IndexDescription(Name="bench", NGramSize=3)
    pinned = S.PinnedBuffers(65536, 10)
    pageable = (np.zeros((65536, 10), np.uint32), np.zeros((65536, 10), np.float64), np.zeros(65536, np.uint32))  # touched once, reused
    for mode, rows in (("peer", "page-locked"), ("peer", "pageable"), ("copy", "page-locked")):
        os.environ["SG_SHARD_GATHER_COPY"] = "1" if mode == "copy" else "0"
        t0 = time.perf_counter()
        sx = ShardedNGramIndex(docs, desc, list(range(n_gpus)))
        build_s = time.perf_counter() - t0
        out = pinned.out if rows == "page-locked" else pageable
        for _ in range(3):
            ids, sc, cnt = sx.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo), out=out)
        t0 = time.perf_counter()
        for _ in range(steps):
            ids, sc, cnt = sx.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo), out=out)
        dt = (time.perf_counter() - t0) / steps
        print(json.dumps({"workload": f"{n_docs}-entry dictionary, {n_gpus} record-id-range shards, one process (sg_sharded_search_batch, "
                                      "host buffers in and out)", "gather": mode, "result_rows": rows, "peer_reads": sx.info()["peer_reads"],
                          "queries_per_s": 65536 / dt, "ms_per_batch": dt * 1e3, "n_gpus": n_gpus, "build_s": round(build_s, 2),
                          "queries_with_a_match": float((cnt > 0).mean()), "timing": "host wall clock around the call"}))
        sx.close()


if __name__ == "__main__":
    main()
