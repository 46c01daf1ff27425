#!/bin/bash
# round 2, second GPU pass: full GPU suite on the count -> resolve pipeline, bench with the count kernel at 4/5/6 CTAs per SM
O=gpurun_out/r2c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-config4 > $O/bench_lean4.json 2> $O/bench_lean4.err
for v in count3 count5; do
  SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
tail -5 $O/pytest_gpu.log
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value %.1fM e2e %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6), d['roofline']['stage_ms'], d.get('gpu_results_identical'), d.get('config3_min_qps'))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p['metric'],p['ngram'],p['letters'],p['bucket_shift'],'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p['host_equals_device'])
"; done
for f in $O/*.err; do tail -n 3 $f; done
