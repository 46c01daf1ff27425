#!/bin/bash
# round 2, first GPU pass: new tests, L2 microbench, new bench.py
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
tools/l2bench gpurun_out/r2a/l2.json > gpurun_out/r2a/l2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_disk.py tests/test_gpu_sharded_ranks.py "tests/test_gpu_parity.py::test_synthetic_stats_match_oracle" -x -q > gpurun_out/r2a/pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a/pytest_new.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench1.json 2> gpurun_out/r2a/bench1.err
echo "bench rc=$?" >> gpurun_out/r2a/bench1.err
tail -c 1500 gpurun_out/r2a/pytest_new.log
tail -c 600 gpurun_out/r2a/bench1.err
head -c 1500 gpurun_out/r2a/bench1.json
