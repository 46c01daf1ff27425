#!/bin/bash
# Round artefacts on the GPU box: default bench line, reference arm, ncu launch list + full captures, sanitizer runs.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench.err; echo rc=$?; cut -c1-300 $OUT/bench_1gpu.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo rc=$?; cut -c1-300 $OUT/bench_reference.json
echo "== e2e: staged against direct result rows"; timeout 600 python tools/e2e_direct.py > $OUT/e2e_direct.txt 2>> $OUT/bench.err; echo rc=$?; cat $OUT/e2e_direct.txt
echo "== e2e device timeline per slice"; SG_TRACE=2 timeout 300 python tools/e2e_trace2.py 2> $OUT/e2e_trace.txt; grep -A4 "====" $OUT/e2e_trace.txt | head -40
echo "== autocomplete / spellchecker"; timeout 600 python bench.py --workload autocomplete > $OUT/bench_autocomplete.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_autocomplete.json
timeout 600 python bench.py --workload spellchecker > $OUT/bench_spellchecker.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_spellchecker.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "rc=$?"
for K in sg_bitmap_search_kernel sg_tokens_kernel; do
echo "== ncu full $K"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/$K \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$K.log 2>&1 ; echo "rc=$?"
done
echo "== compute-sanitizer memcheck (smoke)"; timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > $OUT/memcheck.log 2>&1; tail -3 $OUT/memcheck.log
echo "== compute-sanitizer racecheck (smoke)"; timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
ls -la $OUT
