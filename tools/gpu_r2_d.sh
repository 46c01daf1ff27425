#!/bin/bash
# ncu captures of the pipeline kernels on config #2 + counters
O=gpurun_out/r2d; mkdir -p $O
SG_TRACE=1 python tools/prof_step.py --calls 3 --stages > $O/stages.txt 2>&1
SG_TRACE=1 python tools/prof_step.py --calls 3 --stages --data zipf --metric Cosine > $O/stages_zipf.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sg_count_kernel|sg_resolve_kernel' -s 6 -c 2 -f -o $O/lean python tools/prof_step.py --calls 4 > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -4 $O/stages.txt; tail -4 $O/stages_zipf.txt; ls -la $O
