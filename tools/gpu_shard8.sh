#!/bin/bash
# config #4 at N GPUs: the sharded bench line and the per-piece timing of the step
N=${1:-8}
OUT=gpurun_out/shard$N
mkdir -p $OUT
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
echo "== sharded 10M x$N"; run bench.py --gpus $N --steps 10 --warmup 3 --workload sharded --no-cpu-baseline > $OUT/sharded10m_$N.json 2> $OUT/sharded10m_$N.err; echo rc=$?; cut -c1-260 $OUT/sharded10m_$N.json; tail -n 3 $OUT/sharded10m_$N.err
echo "== shard trace x$N"; run tools/shard_trace.py > $OUT/shard_trace_$N.txt 2> $OUT/shard_trace_$N.err; echo rc=$?; cat $OUT/shard_trace_$N.txt; tail -n 3 $OUT/shard_trace_$N.err
