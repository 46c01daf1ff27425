#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 900 python bench.py --no-config4 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r/bench.json'))
print('value %.1fM (one stream %.1fM) e2e %.1fM ms/step %.2f'%(d['value']/1e6,(d['run']['value_with_one_stream'] or 0)/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['roofline'].get('stage_ms'), d.get('gpu_results_identical'), d.get('config3_min_qps'))
e=d['e2e']; print('callers', e.get('callers'), 'one caller %.1fM'%(e['one_caller']['value']/1e6), e['host'], 'arrays %.1fM'%(e['separate_id_and_score_arrays']['value']/1e6))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p.get('metric'),p.get('ngram'),p.get('letters'),p.get('bucket_shift'),'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p.get('host_equals_device'))
PY
tail -n 5 $O/bench.err
