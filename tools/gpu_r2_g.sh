#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
for d in 0 1 2 3 4; do echo "== SG_RESOLVE_DEBUG=$d"; SG_RESOLVE_DEBUG=$d python tools/prof_step.py --calls 3 --stages 2>&1 | tail -2; done > $O/debug_bits.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sg_resolve_kernel' -s 3 -c 1 -f -o $O/resolve python tools/prof_step.py --calls 4 > $O/ncu.log 2>&1; echo "ncu rc=$?"
cat $O/debug_bits.txt
timeout 600 python -m pytest tests/test_gpu_candidates.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
