#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
for w in 3 4; do
SG_SUBMIT_WORKERS=$w timeout 600 python bench.py --no-config3 --no-config4 --no-cpu-baseline > $O/bench_e2e.json 2> $O/bench_e2e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p/bench_e2e.json'))
e=d['e2e']; print('value %.1fM e2e %.1fM'%(d['value']/1e6, e['value']/1e6), e.get('how'), 'threads %.1fM'%((e['concurrent_host_threads']['value'] or 0)/1e6), 'submit/wait %.1fM'%((e['submit_wait_one_thread']['value'] or 0)/1e6), 'one caller %.1fM'%(e['one_caller']['value']/1e6), e['host'])
PY
done
