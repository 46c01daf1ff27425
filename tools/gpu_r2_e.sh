#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-config4 > $O/bench.json 2> $O/bench.err
SG_FUSED_TOKENS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-config4 --no-config3 --no-cpu-baseline > $O/bench_unfused.json 2> $O/bench_unfused.err
SG_TRACE=1 python tools/prof_step.py --calls 3 --stages --data zipf --metric Cosine > $O/stages_zipf.txt 2>&1
SG_TRACE=2 python tools/e2e_trace.py > $O/e2e_trace.txt 2>&1
grep -B2 -A12 "Error" $O/pytest_gpu.log | head -60
tail -4 $O/pytest_gpu.log
for f in $O/bench.json $O/bench_unfused.json; do python -c "
import json
d=json.load(open('$f'))
print('value %.1fM e2e %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6), d['roofline']['stage_ms'], d.get('gpu_results_identical'), d.get('config3_min_qps'))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p['metric'],p['ngram'],p['letters'],p['bucket_shift'],'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p['host_equals_device'])
"; done
tail -3 $O/bench.err; tail -3 $O/stages_zipf.txt; tail -8 $O/e2e_trace.txt
