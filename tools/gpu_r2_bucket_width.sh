#!/bin/bash
# bucket width against the count -> resolve pipeline: 10M entries on one GPU and config #2
O=gpurun_out/r2w; mkdir -p $O
for bs in 7 8 9 10; do
  echo "== 10M docs, SG_BUCKET_SHIFT=$bs"
  SG_BUCKET_SHIFT=$bs timeout 300 python tools/prof_step.py --calls 4 --docs 10000000 --stages 2>&1 | grep "sg_tokens_count" | cut -c1-330 | tee -a $O/shift_10m.txt
done
for bs in 6 7 8; do
  echo "== 1M docs, SG_BUCKET_SHIFT=$bs"
  SG_BUCKET_SHIFT=$bs timeout 300 python tools/prof_step.py --calls 4 --stages 2>&1 | grep "sg_tokens_count" | cut -c1-330 | tee -a $O/shift_1m.txt
done
