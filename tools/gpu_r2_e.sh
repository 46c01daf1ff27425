#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
# the chunked-arrival path first, on a short leash: a device-side wait must never cost more than this
timeout 180 python -m pytest "tests/test_gpu_parity.py::test_chunked_arrival_equals_sliced_path" "tests/test_gpu_parity.py::test_pinned_result_buffers" -x -q > $O/pytest_chunked.log 2>&1
rc=$?; echo "chunked rc=$rc"; tail -5 $O/pytest_chunked.log
if [ $rc -ne 0 ]; then grep -B2 -A25 "Error\|error" $O/pytest_chunked.log | head -80; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-config4 > $O/bench.json 2> $O/bench.err
SUGGEST_B200_LIB=$PWD/suggest_b200/variants/libsuggest_b200_count3.so timeout 200 python bench.py --steps 5 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/bench_count3.json 2> $O/bench_count3.err
SG_DIRECT_CHUNKS=8 timeout 200 python bench.py --steps 5 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/bench_chunks8.json 2> $O/bench_chunks8.err
SG_DIRECT_CHUNKS=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/bench_chunks0.json 2> $O/bench_chunks0.err
SG_TRACE=1 timeout 120 python tools/prof_step.py --calls 3 --stages --data zipf --metric Cosine > $O/stages_zipf.txt 2>&1
SG_TRACE=2 timeout 120 python tools/e2e_trace.py > $O/e2e_trace.txt 2>&1
grep -B2 -A12 "Error" $O/pytest_gpu.log | head -60
tail -4 $O/pytest_gpu.log
for f in $O/bench.json $O/bench_count3.json $O/bench_chunks8.json $O/bench_chunks0.json; do echo $f; python -c "
import json
d=json.load(open('$f'))
print('value %.1fM e2e %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6), d['roofline']['stage_ms'], d.get('gpu_results_identical'), d.get('config3_min_qps'))
for p in (d.get('config3') or {}).get('points',[]): print('  ',p['metric'],p['ngram'],p['letters'],p['bucket_shift'],'%.1fM e2e %.1fM'%(p['value']/1e6,p['e2e']/1e6), p['host_equals_device'])
print(json.dumps(d.get('single_query'))[:1500])
"; done
tail -3 $O/bench.err; tail -3 $O/stages_zipf.txt; tail -8 $O/e2e_trace.txt
