#!/bin/bash
# count-kernel variants: stage times of config #2 (10 runs each) + parity of the headline test
O=gpurun_out/r2v; mkdir -p $O
for v in "" align32 noalloc align32_noalloc; do
  if [ -n "$v" ]; then export SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_$v.so; else unset SUGGEST_B200_LIB; fi
  echo "== ${v:-default}"
  timeout 300 python tools/prof_step.py --calls 6 --stages 2>&1 | grep sg_tokens_count | cut -c1-260
  timeout 300 python tools/prof_step.py --calls 6 --stages --metric Cosine 2>&1 | grep sg_tokens_count | cut -c1-200
done
export SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_align32.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
