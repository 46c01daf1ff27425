#!/bin/bash
# N ranks, result rows staged in HBM and copied (SG_DIRECT_OUT=0) instead of stored by the kernels over PCIe: does the
# end-to-end rate per rank still fall with N?
N=${1:-8}
O=gpurun_out/r2n$N; mkdir -p $O
SG_DIRECT_OUT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-config4 > $O/bench_staged.json 2> $O/bench_staged.err; echo "rc=$?"
tail -2 $O/bench_staged.err
python - <<PY
import json
d=json.load(open('$O/bench_staged.json'))
print('staged rows N=$N value %.1fM e2e %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6), d.get('gpu_results_identical'), d['e2e']['result_path'])
e=d['e2e']; print('callers', e.get('callers'), 'one caller %.1fM'%(e['one_caller']['value']/1e6)); print(e['host']); print(e['one_caller']['host'])
PY
