#!/bin/bash
# the micro-batcher with one and with two batches in flight
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batcher.py -x -q 2>&1 | tail -3
for w in 1 2 3; do
SG_BATCHER_WORKERS=$w timeout 600 python - <<'PY' | tee -a $O/batcher_workers.txt
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from suggest_b200.workload import synthetic_dictionary, synthetic_queries
d, off, rng = synthetic_dictionary(1_000_000)
q, qo, _ = synthetic_queries(d, off, 65536, rng)
r = bench.measure_single_query(d, off, q, qo)
for k in ("one_caller", "callers_64", "callers_512"):
    v = r[k]
    print("workers", os.environ["SG_BATCHER_WORKERS"], k, {x: v.get(x) for x in ("qps", "p50_us", "p99_us", "mean_batch", "largest_batch")}, flush=True)
PY
done
