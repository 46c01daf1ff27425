#!/bin/bash
O=gpurun_out/r2a; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "autocomplete" 2>&1 | tail -3
timeout 600 python bench.py --workload autocomplete > $O/bench_autocomplete.json 2> $O/bench_autocomplete.err; echo "rc=$?"; tail -3 $O/bench_autocomplete.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a/bench_autocomplete.json'))
print('value %.1fM e2e %.1fM'%(d['value']/1e6, d['e2e']['value']/1e6), d['roofline'], d['cpu_baseline'], d['results'])
PY
