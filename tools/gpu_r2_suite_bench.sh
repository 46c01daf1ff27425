#!/bin/bash
# full GPU suite, then the default bench line
O=gpurun_out/r2p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p/bench.json'))
print('value %.1fM e2e %.1fM ms/step %.2f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['roofline'].get('stage_ms'), d.get('gpu_results_identical'), d.get('config3_min_qps'))
print({k:v for k,v in d['roofline'].items() if k!='stage_ms'})
print(d['e2e']); print(d['cpu_baseline']); print(d.get('clocks')); print(d.get('single_query'))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p.get('metric'),p.get('ngram'),p.get('letters'),p.get('bucket_shift'),'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p.get('host_equals_device'))
print(d.get('config4'))
PY
tail -n 5 $O/bench.err
