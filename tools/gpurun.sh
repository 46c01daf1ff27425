#!/bin/bash
# tools/gpurun.sh <timeout-seconds> <script-or-command> [--gpus N]: rebuild the library (and the oracle) here, then run on the GPU box.
# The .so travels with the snapshot, so a stale one would silently be what gets measured.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" # (forced rebuild)
[ -f tools/l2bench.cu ] && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bench tools/l2bench.cu
T=$1; shift
CMD=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
