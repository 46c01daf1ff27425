#!/bin/bash
O=gpurun_out/r2san; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_parity.py::test_candidate_rows_equal_separate_arrays" -x -q > $O/memcheck.log 2>&1
grep -A22 "Invalid\|ERROR SUMMARY" $O/memcheck.log | head -70
