#!/bin/bash
# round 2 profiles: launch list of the bench command, one full capture per pipeline kernel (config #2 and the 10M single-GPU
# shard), compute-sanitizer memcheck over the pipeline's tests
O=gpurun_out/r2prof; mkdir -p $O
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-config3 --no-config4 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
for K in sg_tokens_count_kernel sg_resolve_kernel; do
  echo "== full capture $K (config #2)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $O/$K python tools/prof_step.py --calls 6 > $O/ncu_$K.log 2>&1; echo "rc=$?"
done
echo "== full capture sg_tokens_count_kernel (10M-entry dictionary on one GPU)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sg_tokens_count_kernel -s 3 -c 1 -f -o $O/sg_tokens_count_kernel_10m python tools/prof_step.py --calls 6 --docs 10000000 > $O/ncu_10m.log 2>&1; echo "rc=$?"
timeout 300 python tools/prof_step.py --calls 6 --docs 10000000 --stages > $O/stages_10m.txt 2>&1; tail -2 $O/stages_10m.txt | cut -c1-400
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "words_dict or large_k or long or chunked or candidate_rows" > $O/memcheck.log 2>&1
grep -A12 "Invalid\|ERROR SUMMARY" $O/memcheck.log | head -40; tail -3 $O/memcheck.log
ls -la $O
