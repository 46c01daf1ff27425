#!/bin/bash
# full GPU suite, then the default bench line
O=gpurun_out/r2p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p/bench.json'))
print('value %.1fM (one stream %.1fM) e2e %.1fM ms/step %.2f'%(d['value']/1e6,(d['run']['value_with_one_stream'] or 0)/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['roofline'].get('stage_ms'), d.get('gpu_results_identical'), d.get('config3_min_qps'))
e=d['e2e']; print(e.get('how'), e['host'], 'threads %.1fM'%((e['concurrent_host_threads']['value'] or 0)/1e6), 'submit/wait %.1fM'%((e['submit_wait_one_thread']['value'] or 0)/1e6), 'one caller %.1fM'%(e['one_caller']['value']/1e6), 'arrays %.1fM'%(e['separate_id_and_score_arrays']['value']/1e6))
print(d['cpu_baseline']); print(d.get('clocks'))
sq=d.get('single_query') or {}
for k in ('one_caller','callers_64','callers_512'): print(k, {x:(sq.get(k) or {}).get(x) for x in ('qps','p50_us','p99_us','mean_batch')})
for p in (d.get('config3') or {}).get('points',[]): print('  ',p.get('metric'),p.get('ngram'),p.get('letters'),p.get('bucket_shift'),'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p.get('host_equals_device'))
c=d.get('config4') or {}
print({k:(round(v['value']/1e6,1), round(v['e2e']/1e6,1), v.get('gpu_results_identical')) for k,v in (c.get('exchange') or {}).items()})
PY
tail -n 5 $O/bench.err
