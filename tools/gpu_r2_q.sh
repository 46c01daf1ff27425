#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 300 python tools/e2e_callers2.py 2>&1 | grep "q/s" | tee $O/callers2.txt
