#!/usr/bin/env python3
"""Time the pieces of the sharded step (search, all-gather, merge) with CUDA events.  torchrun, N GPUs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import suggest_b200 as S  # noqa: E402
from suggest_b200 import _capi  # noqa: E402
from suggest_b200.sharding import ShardedIndex  # noqa: E402
from suggest_b200.suggest import IndexDescription  # noqa: E402
from suggest_b200.workload import synthetic_dictionary, synthetic_queries  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d_bytes, d_off, rng = synthetic_dictionary(N)
q_bytes, q_off, _ = synthetic_queries(d_bytes, d_off, 65536, rng)
desc = IndexDescription(Name="b", NGramSize=3, Alphabet=("english", "russian", "numbers", "$"), Pad="$", Wrap=("$", "$"), Device=local)
sh = ShardedIndex((d_bytes, d_off), desc, rank, world, S.NewRAMBuilder)
nq, K = 65536, 10
dq = torch.from_numpy(q_bytes).to(dev)
doff = torch.from_numpy(q_off.astype(np.int32)).to(dev)
o_ids = torch.zeros(nq * K, dtype=torch.int32, device=dev)
o_sc = torch.zeros(nq * K, dtype=torch.float64, device=dev)
o_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
b = sh._alloc(nq, K, dev)
L = _capi.lib()
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3)
for it in range(13):
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    _capi.check(L.sg_search_batch_packed_device(sh.index.handle, dq.data_ptr(), doff.data_ptr(), nq, S.JaccardMetric().code, 0.5, K, b["mine"].data_ptr(), st or None))
    ev[1].record()
    dist.all_gather_into_tensor(b["all"], b["mine"])
    ev[2].record()
    _capi.check(L.sg_merge_topk_packed_device(local, world, nq, K, b["all"].data_ptr(), o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), st or None))
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
if rank == 0:
    print("search %.3f ms  all-gather %.3f ms  merge %.3f ms  (bytes per rank %d)" % (*(acc / 10), b["mine"].numel()), flush=True)
dist.barrier()
dist.destroy_process_group()
