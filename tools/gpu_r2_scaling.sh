#!/bin/bash
# the driver's scaling command at N ranks: default bench line (replicated headline + config #4 sharded over the N ranks)
N=${1:-8}
O=gpurun_out/r2n$N; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?"
tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open('$O/bench.json'))
print('N=$N value %.1fM (one stream %.1fM) e2e %.1fM'%(d['value']/1e6,(d['run']['value_with_one_stream'] or 0)/1e6,d['e2e']['value']/1e6), d.get('gpu_results_identical'))
e=d['e2e']; print(e.get('how'), 'threads %.1fM'%((e['concurrent_host_threads']['value'] or 0)/1e6), 'submit/wait %.1fM'%((e['submit_wait_one_thread']['value'] or 0)/1e6), 'one caller %.1fM'%(e['one_caller']['value']/1e6)); print(e['host']); print(e['one_caller']['host'])
c=d['config4']; print({k:(round(v['value']/1e6,1), round(v['e2e']/1e6,1), v['stage_ms'], v.get('gpu_results_identical'), v.get('e2e_host')) for k,v in c['exchange'].items()})
PY
nproc; cat /proc/cpuinfo | grep "model name" | sort | uniq -c
