#!/bin/bash
# N-GPU runs of bench.py: replicated 1M index (weak scaling) and the record-id-range sharded 10M dictionary (config #4)
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
echo "== replicated x$N"; run --steps 20 --warmup 3 --no-cpu-baseline > $OUT/replicated_$N.json 2> $OUT/replicated_$N.err; cut -c1-260 $OUT/replicated_$N.json
echo "== sharded 10M x$N"; run --steps 10 --warmup 3 --workload sharded --no-cpu-baseline > $OUT/sharded10m_$N.json 2> $OUT/sharded10m_$N.err; cut -c1-260 $OUT/sharded10m_$N.json
for f in $OUT/*.err; do tail -n 3 "$f"; done
