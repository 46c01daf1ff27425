#!/bin/bash
# multi-GPU pass: rank tests (NCCL + fused exchange on distinct GPUs), sharded bench at N ranks, replicated bench at N ranks
N=${1:-2}
O=gpurun_out/r2m$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded_ranks.py tests/test_gpu_sharded_capi.py -x -q > $O/pytest_ranks.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ranks.log
tail -5 $O/pytest_ranks.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload sharded > $O/bench_sharded.json 2> $O/bench_sharded.err; echo "sharded rc=$?"
tail -3 $O/bench_sharded.err
python -c "
import json
d=json.load(open('$O/bench_sharded.json'))
c=d['config4']; print('config4 N=$N', {k:(round(v['value']/1e6,1), round(v['e2e']/1e6,1), v['stage_ms'], v.get('gpu_results_identical')) for k,v in c['exchange'].items()})
"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-config4 > $O/bench_repl.json 2> $O/bench_repl.err; echo "repl rc=$?"
tail -3 $O/bench_repl.err
python -c "
import json
d=json.load(open('$O/bench_repl.json'))
print('replicated N=$N value %.1fM e2e %.1fM'%(d['value']/1e6,d['e2e']['value']/1e6), d.get('gpu_results_identical'))
"
