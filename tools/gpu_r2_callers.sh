#!/bin/bash
# end-to-end rate against the number of concurrent callers, with 32 and with the default 8 CUDA hardware queues
O=gpurun_out/r2q; mkdir -p $O
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/e2e_callers.py 200 2>&1 | grep callers | tee $O/callers_conn32.txt
timeout 300 python tools/e2e_callers.py 200 2>&1 | grep callers | tee $O/callers_conn8.txt
