"""Record-id-range shards of one dictionary over the GPUs of a box (BASELINE.json config #4).

The reference is a single process and has nothing like this.  A query's answer is the top-k of the union of
the per-shard top-k lists, so: rank r indexes documents [lo_r, hi_r) with id_base = lo_r (returned ids stay
global, which keeps the (score desc, id asc) order meaningful across shards), every rank searches the same
query batch and writes its rows as one packed block (scores | ids | counts); the exchange and the merge are one
kernel per rank over NVLink peer memory (sg_exchange.cu: rank r merges 1/world of the queries from all shards' rows
and stores the winners into every rank's result), or - SG_SHARD_EXCHANGE=nccl - ONE all-gather of the blocks (NCCL)
and sg_merge_topk_packed_device re-selecting k per query on every rank.
"""
import numpy as np

from . import _capi


def shard_bounds(n_docs, world):
    """contiguous, balanced [lo, hi) per rank"""
    return [(r * n_docs // world, (r + 1) * n_docs // world) for r in range(world)]


def slice_packed(data, off, lo, hi):
    """documents [lo, hi) of a packed (bytes, offsets) dictionary, offsets rebased to 0"""
    sub_off = (off[lo:hi + 1] - off[lo]).astype(off.dtype)
    return data[int(off[lo]):int(off[hi])], sub_off


class ShardedIndex:
    """One rank's shard plus the cross-shard reduce.  Needs torch.distributed initialised (any backend: it only carries
    the 64-byte IPC handles of the exchange regions at construction; the nccl backend is needed for exchange="nccl").

    exchange="fused" (default; SG_SHARD_EXCHANGE overrides): sg_exchange_* - the per-shard rows stay in the HBM of the GPU
    that wrote them, ONE kernel per rank merges that rank's 1/world of the queries reading the other shards' rows over
    NVLink peer memory and stores the winners into every rank's result (sg_exchange.cu).
    exchange="nccl": one NCCL all-gather of the full fixed-stride blocks + sg_merge_topk_packed_device of every query on
    every rank (round 1; also what is used when CUDA IPC is not available)."""

    def __init__(self, dictionary, description, rank, world, builder_factory, max_queries=65536, max_k=16, exchange=None):
        import os
        data, off = dictionary
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_bounds(len(off) - 1, world)[rank]
        self.index = builder_factory(slice_packed(data, off, self.lo, self.hi), description, self.lo).Build()
        self._buffers = None
        self._ex = None
        self.exchange = exchange or os.environ.get("SG_SHARD_EXCHANGE", "fused")
        self.exchange_note = ""
        if world > 1 and self.exchange == "fused":
            self._connect(description.Device, max_queries, max_k)

    def _connect(self, device, max_queries, max_k):
        import ctypes as C
        import torch.distributed as dist
        L = _capi.lib()
        h = C.c_void_p()
        _capi.check(L.sg_exchange_create(int(device), self.rank, self.world, int(max_queries), int(max_k), C.byref(h)))
        mine = C.create_string_buffer(_capi.SG_EXCHANGE_HANDLE_BYTES)
        _capi.check(L.sg_exchange_handle(h, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.raw)
        rc = L.sg_exchange_connect(h, b"".join(handles))
        ok = [None] * self.world
        dist.all_gather_object(ok, int(rc))
        if any(r != 0 for r in ok):  # collective decision: every rank takes the same path
            self.exchange_note = "CUDA IPC unavailable (%s): NCCL all-gather path" % (L.sg_last_error() or b"").decode()
            L.sg_exchange_free(h)
            self.exchange = "nccl"
            return
        self._ex = h
        self._max = (int(max_queries), int(max_k))

    def _alloc(self, n_q, k, device):
        import torch
        key = (n_q, k, str(device))
        if self._buffers is None or self._buffers[0] != key:
            nbytes = int(_capi.lib().sg_packed_rows_bytes(n_q, k))
            self._buffers = (key, dict(mine=torch.zeros(nbytes, dtype=torch.uint8, device=device),
                                       all=torch.zeros(self.world * nbytes, dtype=torch.uint8, device=device)))
        return self._buffers[1]

    def SuggestBatchDevice(self, d_q, d_off, n_q, similarity, metric, k, out_ids, out_scores, out_counts):
        """d_q / d_off / out_*: torch tensors on this rank's device.  Enqueued on torch's current stream.  Collective."""
        import torch
        import torch.distributed as dist
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            self.index.SuggestBatchDevice(d_q.data_ptr(), d_off.data_ptr(), n_q, similarity, metric, k, out_ids.data_ptr(),
                                          out_scores.data_ptr(), out_counts.data_ptr(), 0, stream)
            return
        L = _capi.lib()
        if self._ex is not None and n_q <= self._max[0] and k <= self._max[1]:
            _capi.check(L.sg_exchange_search(self._ex, self.index.handle, d_q.data_ptr(), d_off.data_ptr(), n_q, metric.code,
                                             float(similarity), int(k), out_ids.data_ptr(), out_scores.data_ptr(),
                                             out_counts.data_ptr(), stream or None))
            return
        b = self._alloc(n_q, k, d_q.device)
        _capi.check(L.sg_search_batch_packed_device(self.index.handle, d_q.data_ptr(), d_off.data_ptr(), n_q, metric.code,
                                                    float(similarity), int(k), b["mine"].data_ptr(), stream or None))
        dist.all_gather_into_tensor(b["all"], b["mine"])
        _capi.check(L.sg_merge_topk_packed_device(d_q.device.index, self.world, n_q, k, b["all"].data_ptr(), out_ids.data_ptr(),
                                                  out_scores.data_ptr(), out_counts.data_ptr(), stream or None))

    def check_exchange(self):
        """SG_OK, or raises if a barrier of the fused exchange timed out (synchronises the current stream)"""
        import torch
        if self._ex is not None:
            _capi.check(_capi.lib().sg_exchange_status(self._ex, torch.cuda.current_stream().cuda_stream or None))

    def close(self):
        if self._ex is not None:
            _capi.lib().sg_exchange_free(self._ex)
            self._ex = None
        self.index.close()


class ShardedNGramIndex:
    """The same record-id-range shards driven by ONE process (sg_sharded_*): what a Go service, which cannot be one
    process per GPU, calls.  `devices`: CUDA ordinal per shard (a GPU may hold several).  Suggest / SuggestBatch as
    NGramIndex; the exchange is the merge kernel reading the other GPUs' rows over NVLink peer access."""

    def __init__(self, dictionary, description, devices):
        import ctypes as C
        from .suggest import pack_strings
        if isinstance(dictionary, tuple) and len(dictionary) == 2 and isinstance(dictionary[0], np.ndarray):
            data, off = dictionary
        else:
            data, off = pack_strings(dictionary, np.uint64)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        cfg, keep = description.c_config()
        h = C.c_void_p()
        rc = _capi.lib().sg_sharded_build(C.byref(cfg), data.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(off) - 1,
                                          dev.ctypes.data_as(C.c_void_p), len(dev), C.byref(h))
        del keep
        _capi.check(rc)
        self._h = h

    def info(self):
        import ctypes as C
        n, d, p = C.c_uint32(), C.c_uint32(), C.c_int32()
        _capi.check(_capi.lib().sg_sharded_get_info(self._h, C.byref(n), C.byref(d), C.byref(p)))
        return dict(n_shards=n.value, n_docs=d.value, peer_reads=p.value)

    def shard_info(self, s):
        import ctypes as C
        info = _capi.SgIndexInfo()
        _capi.check(_capi.lib().sg_index_get_info(C.c_void_p(_capi.lib().sg_sharded_shard(self._h, s)), C.byref(info)))
        return {name: getattr(info, name) for name, _ in info._fields_}

    def SuggestBatch(self, queries, similarity, metric, topK, packed=None, out=None):
        """`out`: PinnedBuffers(n_q, k).out - the merge kernel then stores the valid entries straight into them."""
        import ctypes as C
        from .suggest import pack_strings
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_q, k = len(off) - 1, max(int(topK), 0)
        if out is None:
            ids = np.zeros((n_q, k), dtype=np.uint32)
            scores = np.zeros((n_q, k), dtype=np.float64)
            counts = np.zeros(n_q, dtype=np.uint32)
        else:
            ids, scores, counts = out
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _capi.check(_capi.lib().sg_sharded_search_batch(self._h, p(data), p(off), n_q, metric.code, float(similarity), k, p(ids),
                                                        p(scores), p(counts)))
        return ids, scores, counts

    def Suggest(self, query, similarity, metric, topK):
        from .collector import Candidate
        ids, scores, counts = self.SuggestBatch([query], similarity, metric, topK)
        return [Candidate(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def close(self):
        if self._h is not None and self._h.value:
            _capi.lib().sg_sharded_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_rows_reference(part_ids, part_scores, part_counts, k):
    """numpy statement of the merge order ([part][query][k] -> [query][k]); what sg_merge_topk_kernel computes.
    Used by the CPU tests of the sharding plan; never on the product path."""
    n_parts, n_q = part_counts.shape
    ids = np.zeros((n_q, k), dtype=np.uint32)
    scores = np.zeros((n_q, k), dtype=np.float64)
    counts = np.zeros(n_q, dtype=np.uint32)
    for q in range(n_q):
        rows = [(-float(part_scores[p, q, i]), int(part_ids[p, q, i])) for p in range(n_parts) for i in range(int(part_counts[p, q]))]
        rows.sort()
        rows = rows[:k]
        counts[q] = len(rows)
        for i, (neg, idx) in enumerate(rows):
            ids[q, i], scores[q, i] = idx, -neg
    return ids, scores, counts
