#!/bin/bash
# One GPU-box session: smoke, parity tests, bench, ncu launch list + full capture of the search kernel.
# usage (under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $OUT/smoke.log
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $OUT/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -15 $OUT/pytest.log
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; cat $OUT/bench.json ; tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu1 rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-sg_bitmap_search_kernel} -s 2 -c 1 -f -o $OUT/prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1 ; echo "ncu2 rc=$?"
ls -la $OUT
fi
