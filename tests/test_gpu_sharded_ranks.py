"""Record-id-range shards with ONE PROCESS PER GPU (suggest_b200/sharding.py: ShardedIndex) against the oracle.
Needs a B200: `pytest -m gpu`.

world ranks are spawned, each builds its shard, all run ShardedIndex.SuggestBatchDevice on the same batch (collective),
and EVERY rank's merged rows must equal the oracle over the whole dictionary (ids in order, scores bit-equal).
  * distinct GPUs, nccl backend: world = 2 .. min(GPUs, 8); both exchanges (fused peer-memory kernel, NCCL all-gather + merge)
  * one GPU shared by two ranks, gloo for the 64-byte handle exchange: the fused kernel over CUDA IPC on a 1-GPU box
    (NCCL refuses two ranks on one device, so only the fused exchange can be exercised there)
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_DOCS, N_Q, K = 120000, 4096, 10


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def n_gpus():
    import torch
    return torch.cuda.device_count()


def worker(rank, world, port, backend, exchange, share_gpu, out_dir):
    import torch
    import torch.distributed as dist
    import suggest_b200 as S
    from suggest_b200.sharding import ShardedIndex
    from suggest_b200.suggest import IndexDescription
    from suggest_b200.workload import synthetic_workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = 0 if share_gpu else rank
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        docs, (qb, qo), _ = synthetic_workload(N_DOCS, N_Q)
        desc = IndexDescription(Name="s", NGramSize=3, Device=device)
        sx = ShardedIndex(docs, desc, rank, world, S.NewRAMBuilder, max_queries=N_Q, max_k=32, exchange=exchange)
        assert sx.exchange == exchange, sx.exchange_note
        dq = torch.from_numpy(qb).to(dev)
        doff = torch.from_numpy(qo.astype(np.int32)).to(dev)
        results = {}
        for name, metric, alpha, k in (("jaccard", S.JaccardMetric(), 0.5, K), ("cosine", S.CosineMetric(), 0.45, 3),
                                       ("dice", S.DiceMetric(), 0.4, 25)):
            ids = torch.zeros(N_Q * k, dtype=torch.int32, device=dev)
            sc = torch.zeros(N_Q * k, dtype=torch.float64, device=dev)
            cnt = torch.zeros(N_Q, dtype=torch.int32, device=dev)
            for _ in range(3):  # repeated steps reuse the regions and the flags
                sx.SuggestBatchDevice(dq, doff, N_Q, alpha, metric, k, ids, sc, cnt)
            sx.check_exchange()
            torch.cuda.synchronize()
            results[name + "_ids"] = ids.cpu().numpy().view(np.uint32).reshape(N_Q, k)
            results[name + "_scores"] = sc.cpu().numpy().reshape(N_Q, k)
            results[name + "_counts"] = cnt.cpu().numpy().view(np.uint32)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **results)
        if world > 1:
            dist.barrier()
        sx.close()
    finally:
        dist.destroy_process_group()


def run_and_check(world, backend, exchange, share_gpu, tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from suggest_b200.workload import synthetic_workload
    mp.spawn(worker, args=(world, free_port(), backend, exchange, share_gpu, str(tmp_path)), nprocs=world, join=True)
    docs, (qb, qo), _ = synthetic_workload(N_DOCS, N_Q)
    ox = O.OracleIndex(3, ("$", "$"), "$", ("english", "russian", "numbers", "$")).add_packed(*docs)
    for name, code, alpha, k in (("jaccard", O.JACCARD, 0.5, K), ("cosine", O.COSINE, 0.45, 3), ("dice", O.DICE, 0.4, 25)):
        o_ids, o_sc, o_cnt = ox.suggest_batch(None, code, alpha, k, O.CANONICAL, threads=8, packed=(qb, qo.astype(np.uint64)))
        m = np.arange(k)[None, :] < o_cnt[:, None]
        assert o_cnt.sum() > 0
        for rank in range(world):
            got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
            assert np.array_equal(got[name + "_counts"], o_cnt), (name, rank)
            assert np.array_equal(got[name + "_ids"][m], o_ids[m]), (name, rank)
            assert np.array_equal(got[name + "_scores"][m], o_sc[m]), (name, rank)


@pytest.mark.parametrize("exchange", ["fused", "nccl"])
def test_ranks_on_distinct_gpus_equal_oracle(exchange, tmp_path):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for world in sorted({2, min(n, 4), min(n, 8)}):
        d = tmp_path / f"w{world}"
        d.mkdir()
        run_and_check(world, "nccl", exchange, False, d)


def test_two_ranks_sharing_one_gpu_fused_exchange(tmp_path):
    """CUDA IPC between two processes on one device: the fused kernel's peer loads / stores and both flag barriers
    (the contexts time-slice the GPU, so each barrier costs a time slice; results must still be exact)"""
    run_and_check(2, "gloo", "fused", True, tmp_path)
