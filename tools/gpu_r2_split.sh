#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python tools/e2e_split_sweep.py 2>&1 | grep "caller" | tee $O/split_sweep.txt
