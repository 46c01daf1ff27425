"""Host logic of the record-id-range sharding on CPU: world_size 2, gloo.

Every rank searches its shard with the oracle (there is no GPU here), the rows are all-gathered in the layout the
product uses ([part][query][k]) and merged under (score desc, id asc); the result must equal the oracle over the
whole dictionary.  The GPU version of the same flow is tests/test_gpu_parity.py::test_shards_and_merge.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CARS_DESCRIPTION, GOLDEN
from oracle import oracle as O
from suggest_b200.sharding import merge_rows_reference, shard_bounds, slice_packed
from suggest_b200.suggest import pack_strings


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, lines, queries, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data, off = pack_strings(lines, np.uint64)
        lo, hi = shard_bounds(len(lines), world)[rank]
        sub, sub_off = slice_packed(data, off, lo, hi)
        ox = O.OracleIndex(**CARS_DESCRIPTION).add_packed(sub, sub_off)
        ids, sc, cnt = ox.suggest_batch(queries, O.JACCARD, 0.5, k, O.CANONICAL)
        ids = ids + np.uint32(lo)  # id_base
        nq = len(queries)
        t_ids = torch.from_numpy(ids.astype(np.int64).reshape(-1))  # flat rows, as ShardedIndex gathers them
        t_sc = torch.from_numpy(sc.reshape(-1))
        t_cnt = torch.from_numpy(cnt.astype(np.int64))
        g_ids = torch.zeros(world * nq * k, dtype=torch.int64)
        g_sc = torch.zeros(world * nq * k, dtype=torch.float64)
        g_cnt = torch.zeros(world * nq, dtype=torch.int64)
        dist.all_gather_into_tensor(g_ids, t_ids)
        dist.all_gather_into_tensor(g_sc, t_sc)
        dist.all_gather_into_tensor(g_cnt, t_cnt)
        m = merge_rows_reference(g_ids.numpy().reshape(world, nq, k), g_sc.numpy().reshape(world, nq, k),
                                 g_cnt.numpy().reshape(world, nq), k)
        if rank == 0:
            np.savez(out, ids=m[0], scores=m[1], counts=m[2])
    finally:
        dist.destroy_process_group()


def test_two_shards_equal_one_index(tmp_path):
    with open(os.path.join(GOLDEN, "cars.dict"), "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    queries = lines[::9]
    k = 10
    out = str(tmp_path / "merged.npz")
    mp.spawn(worker, args=(2, free_port(), lines, queries, k, out), nprocs=2, join=True)
    got = np.load(out)
    ox = O.OracleIndex(**CARS_DESCRIPTION).add_docs(lines)
    ids, sc, cnt = ox.suggest_batch(queries, O.JACCARD, 0.5, k, O.CANONICAL, threads=4)
    assert np.array_equal(got["counts"], cnt)
    mask = np.arange(k)[None, :] < cnt[:, None]
    assert np.array_equal(got["ids"][mask], ids[mask])
    assert np.array_equal(got["scores"][mask], sc[mask])


def test_shard_bounds_cover_everything():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (0, 2)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
