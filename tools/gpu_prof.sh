#!/bin/bash
# ncu session on the GPU box: launch list of a short bench run + one full capture per kernel named in $KERNELS.
# usage (under gpurun): KERNELS="sg_bitmap_search_kernel sg_tokens_kernel" bash tools/gpu_prof.sh [tag]
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KERNELS=${KERNELS:-sg_bitmap_search_kernel}
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; cat $OUT/bench.json ; tail -3 $OUT/bench.err
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > $OUT/ncu_bench.log 2>&1 ; echo "ncu1 rc=$?"
for K in $KERNELS; do
echo "== ncu full $K"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/$K \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > $OUT/ncu_full_$K.log 2>&1 ; echo "ncu2 rc=$?"
done
ls -la $OUT
