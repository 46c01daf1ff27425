#!/bin/bash
# Round artefacts on the GPU box: default bench line, reference arm, ncu launch list + full captures, sanitizer runs.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench.err; echo rc=$?; cut -c1-300 $OUT/bench_1gpu.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo rc=$?; cut -c1-300 $OUT/bench_reference.json
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
