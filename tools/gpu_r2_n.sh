#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -B2 -A14 "Error" $O/pytest_gpu.log | head -50; tail -4 $O/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-config4 > $O/bench.json 2> $O/bench.err
SG_PIPELINE=classic timeout 400 python bench.py --steps 5 --warmup 3 --no-config4 --no-cpu-baseline > $O/bench_classic.json 2> $O/bench_classic.err
for f in $O/bench.json $O/bench_classic.json; do echo $f; python -c "
import json
d=json.load(open('$f'))
print('value %.1fM e2e %.1fM arrays %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6, d['e2e']['separate_id_and_score_arrays']['value']/1e6), d['roofline']['stage_ms'], d.get('gpu_results_identical'), d.get('config3_min_qps'))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p['metric'],p['ngram'],p['letters'],p['bucket_shift'],'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p['host_equals_device'], round(p['queries_with_a_match'],3))
"; done
tail -3 $O/bench.err $O/bench_classic.err
