#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
for v in "" resolve5 resolve6 resolve8 count3; do
  if [ -n "$v" ]; then export SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_$v.so; else unset SUGGEST_B200_LIB; fi
  echo "== ${v:-default}"
  timeout 300 python tools/prof_step.py --calls 6 --stages 2>&1 | grep sg_tokens_count | sed "s/.*'sg_tokens_count_kernel': \([0-9.]*\), 'sg_resolve_kernel': \([0-9.]*\),.*/config2: count \1 resolve \2/"
  timeout 300 python tools/prof_step.py --calls 4 --stages --data zipf --metric Cosine 2>&1 | grep sg_tokens_count | sed "s/.*'sg_tokens_count_kernel': \([0-9.]*\), 'sg_resolve_kernel': \([0-9.]*\),.*/zipf Cosine: count \1 resolve \2/"
done
export SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_count3.so
timeout 300 python bench.py --no-config3 --no-config4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('count3 bench: value %.1fM one stream %.1fM e2e %.1fM'%(d['value']/1e6, d['run']['value_with_one_stream']/1e6, d['e2e']['value']/1e6))"
