#!/bin/bash
# compute-sanitizer memcheck over the tests added last: submit / wait, the device autocomplete entry, every scratch-overflow point
O=gpurun_out/r2s; mkdir -p $O
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "submit or autocomplete_device or scratch_overflow" > $O/memcheck_new.log 2>&1
grep -A8 "Invalid\|ERROR SUMMARY" $O/memcheck_new.log | head -30; tail -3 $O/memcheck_new.log
