#!/bin/bash
# bench.py under a list of "VAR=VALUE[,VAR=VALUE]" tuning settings; one JSON line each -> gpurun_out/<tag>/sweep.jsonl
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
: > $OUT/sweep.jsonl
for setting in "$@"; do
  envs=$(echo "$setting" | tr ',' ' ')
  line=$(env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>>$OUT/sweep.err)
  echo "{\"setting\": \"$setting\", \"result\": ${line:-null}}" >> $OUT/sweep.jsonl
  echo "$setting => $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,2),"Mq/s kernel_ms",round(d["roofline"]["kernel_ms"],3),"e2e",round(d["e2e"]["value"]/1e6,2),"frac",round(d["roofline"]["frac"],3))' 2>/dev/null)"
done
