#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 300 python tools/prof_step.py --calls 6 --stages 2>&1 | grep sg_tokens_count | cut -c1-260
