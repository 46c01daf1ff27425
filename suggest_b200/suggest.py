"""pkg/suggest mirror over the C ABI: IndexDescription, Builder, NGramIndex (Suggester), Service, SearchConfig.

Same names, argument meaning and error behaviour as the reference for the Suggest path
(pkg/suggest/{config,ngram_index_builder,ngram_index,suggester,service,search}.go); the work itself
happens in libsuggest_b200.so on the GPU.  `SuggestBatch` is the one addition: the reference has no
batched call, a single Suggest is a batch of one.
"""
import ctypes as C
import json
import os
import threading
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _capi
from . import collector as _col
from ._capi import SuggestError
from .collector import Candidate
from .metric import Metric

RAMDriver = "RAM"
DiscDriver = "DISC"


def _b(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def pack_strings(strings, offset_dtype=np.uint32):
    """list of str/bytes -> (uint8 array, offsets[n+1])"""
    bs = [_b(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=offset_dtype)
    if bs:
        off[1:] = np.cumsum([len(x) for x in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return data, off


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


CANDIDATE_DTYPE = np.dtype([("key", "<u4"), ("reserved", "<u4"), ("score", "<f8")])  # sg_candidate = suggest.Candidate in memory


class Ticket:
    """a submitted call (NGramIndex.SubmitBatchCandidates); keeps the call's buffers alive until it is waited for"""

    def __init__(self, handle, buffers):
        self._handle, self._buffers = handle, buffers

    def wait(self):
        handle, self._handle = self._handle, None
        if handle is None:
            raise RuntimeError("the ticket was already waited for")
        _capi.check(_capi.lib().sg_ticket_wait(handle))
        _, _, rows, counts = self._buffers
        return rows, counts

    def __del__(self):
        try:
            if self._handle is not None:  # never waited for: the library still owns the buffers; wait before they go
                self.wait()
        except Exception:
            pass


class PinnedCandidateRows:
    """Page-locked rows of sg_candidate entries + counts for NGramIndex.SuggestBatchCandidates (pass `.out`)."""

    def __init__(self, n_q, k):
        self._ptrs = []
        self.rows = self._alloc((n_q, k), CANDIDATE_DTYPE)
        self.counts = self._alloc((n_q,), np.uint32)
        self.out = (self.rows, self.counts)

    def _alloc(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _capi.check(_capi.lib().sg_pinned_alloc(max(n, 1), C.byref(p)))
        self._ptrs.append(p.value)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        ptrs, self._ptrs = self._ptrs, []
        self.rows = self.counts = self.out = None
        for p in ptrs:
            _capi.lib().sg_pinned_free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffers:
    """Result buffers for SuggestBatch / AutocompleteBatch in page-locked memory (sg_pinned_alloc): the search kernel
    stores the candidates straight into them, entries at and behind counts[q] are left as they were.  Pass `.out` as
    the `out=` argument; the arrays are views of the allocation and die with this object."""

    def __init__(self, n_q, k):
        self._ptrs = []
        self.ids = self._alloc((n_q, k), np.uint32)
        self.scores = self._alloc((n_q, k), np.float64)
        self.counts = self._alloc((n_q,), np.uint32)
        self.out = (self.ids, self.scores, self.counts)

    def _alloc(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _capi.check(_capi.lib().sg_pinned_alloc(max(n, 1), C.byref(p)))
        self._ptrs.append(p.value)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        a[...] = 0
        return a

    def close(self):
        ptrs, self._ptrs = self._ptrs, []
        self.ids = self.scores = self.counts = self.out = None
        for p in ptrs:
            _capi.lib().sg_pinned_free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class IndexDescription:
    """suggest.IndexDescription, pkg/suggest/config.go:25-35"""
    Name: str = ""
    NGramSize: int = 3
    Alphabet: Sequence[str] = ("english", "russian", "numbers", "$")
    Pad: str = "$"
    Wrap: Sequence[str] = ("$", "$")
    Driver: str = RAMDriver
    SourcePath: str = ""
    OutputPath: str = ""
    basePath: str = ""
    Device: int = 0  # CUDA ordinal holding the index (not in the reference)

    def GetSourcePath(self):
        return self.SourcePath if os.path.isabs(self.SourcePath) else f"{self.basePath}/{self.SourcePath}"

    def GetIndexPath(self):
        return self.OutputPath if os.path.isabs(self.OutputPath) else f"{self.basePath}/{self.OutputPath}"

    def GetDictionaryFile(self):
        return f"{self.GetIndexPath()}/{self.Name}.cdb"

    def getHeaderFile(self):
        return f"{self.Name}.hd"

    def getDocumentListFile(self):
        return f"{self.Name}.dl"

    def c_config(self):
        return _capi.make_config(self.NGramSize, tuple(self.Wrap), self.Pad, tuple(self.Alphabet), self.Device)


def ReadConfigs(config_path):
    """suggest.ReadConfigs, pkg/suggest/config.go:84-112"""
    with open(config_path, "rb") as f:
        raw = json.load(f)
    base = os.path.dirname(config_path)
    return [IndexDescription(Name=c.get("name", ""), NGramSize=c.get("nGramSize", 0), Alphabet=tuple(c.get("alphabet", ())),
                             Pad=c.get("pad", ""), Wrap=tuple(c.get("wrap", ("", ""))), Driver=c.get("driver", ""),
                             SourcePath=c.get("source", ""), OutputPath=c.get("output", ""), basePath=base) for c in raw]


@dataclass
class ResultItem:
    """suggest.ResultItem, pkg/suggest/service.go:11-16"""
    Score: float
    Value: str


@dataclass
class SearchConfig:
    query: str
    topK: int
    metric: Metric
    similarity: float


def NewSearchConfig(query, topK, metric, similarity):
    """suggest.NewSearchConfig, pkg/suggest/search.go:18-33"""
    if topK <= 0:
        raise ValueError("topK should be greater or equal to 1")
    if similarity <= 0 or similarity > 1:
        raise ValueError("similarity shouble be in (0.0, 1.0]")
    return SearchConfig(query, topK, metric, similarity)


def open_ram_dictionary(path):
    """dictionary.OpenRAMDictionary, pkg/dictionary/helpers.go:25-48: one entry per line, id = line number"""
    with open(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]  # bufio.ScanLines drops a trailing \r


class NGramIndex:
    """suggest.NGramIndex (pkg/suggest/ngram_index.go:7-10) backed by an sg_index handle in HBM."""

    def __init__(self, handle, description):
        self._h = C.c_void_p(handle)
        self.description = description
        self._lock = threading.Lock()

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        with self._lock:
            if self._h is not None and self._h.value:
                _capi.lib().sg_index_free(self._h)
                self._h = None

    def __del__(self):  # the reference relies on GC finalizers as well (pkg/index/index_reader.go:49-51)
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        if self._h is None:
            raise SuggestError(_capi.SG_ERR_INVALID, "index is closed")
        return self._h

    def info(self):
        info = _capi.SgIndexInfo()
        _capi.check(_capi.lib().sg_index_get_info(self.handle, C.byref(info)))
        return {name: getattr(info, name) for name, _ in info._fields_}

    def layout(self):
        """Slots, bucket width and engine of the handle (sg_index_get_layout)."""
        lay = _capi.SgIndexLayout()
        _capi.check(_capi.lib().sg_index_get_layout(self.handle, C.byref(lay)))
        return {name: getattr(lay, name) for name, _ in lay._fields_}

    # -- Suggester ------------------------------------------------------------------------------
    def Suggest(self, query, similarity, metric, topK) -> List[Candidate]:
        """nGramSuggester.Suggest (pkg/suggest/suggester.go:46-131).  `topK` is a number - FuzzyCollectorManager(topK), the
        reference's newFuzzyCollectorManager - or a CollectorManagerFactory as in the Go signature."""
        return self.SuggestMany([query], similarity, metric, topK)[0]

    def SuggestMany(self, queries, similarity, metric, topK) -> List[List[Candidate]]:
        """Suggest for a list of queries, as lists of Candidate.  A built-in metric with a FuzzyCollectorManager runs
        sg_search_batch (top-k on the device).  Anything else - a metric.Metric of the caller's own, or another
        CollectorManager - gets every candidate of the T-occurrence count from sg_candidates_batch and is driven on the
        host the way suggester.go:66-108 drives it (the two interface wrinkles)."""
        factory = topK if callable(topK) else None
        k = None if factory else int(topK)
        if factory is not None:
            probe = factory()
            if type(probe) is _col.FuzzyCollectorManager:
                k, factory = probe.topK, None
        if factory is None and metric.code is not None:
            ids, scores, counts = self.SuggestBatch(queries, similarity, metric, k)
            return [[Candidate(int(ids[q, i]), float(scores[q, i])) for i in range(int(counts[q]))] for q in range(len(queries))]
        if factory is None:
            factory = _col.NewFuzzyCollectorManager(k)
        cq, cid, cov, cseg, size_a = self.CandidatesBatch(queries, similarity, metric)
        order = np.lexsort((cid, cseg, cq))  # per query, per segment, ids ascending: the order the mergers emit
        cq, cid, cov, cseg = cq[order], cid[order], cov[order], cseg[order]
        starts = np.searchsorted(cq, np.arange(len(queries) + 1))
        S = self.info()["n_segments"]
        out = []
        for q in range(len(queries)):
            lo, hi = int(starts[q]), int(starts[q + 1])
            by_segment = {}
            for i in range(lo, hi):
                by_segment.setdefault(int(cseg[i]), []).append(_col.MergeCandidate(int(cid[i]), int(cov[i])))
            out.append(_replay(factory(), metric, float(similarity), int(size_a[q]), S, by_segment))
        return out

    def CandidatesBatch(self, queries, similarity, metric, packed=None, cap=None):
        """sg_candidates_batch: every (query, document, overlap, segment) with overlap >= Threshold over the admissible
        segments, in no particular order, plus len(tokens) per query.  A metric without a device code is tabulated here
        (Threshold over its [MinY, MaxY] window for every len(tokens) up to 128) and passed as a byte table."""
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_q = len(off) - 1
        table = None
        if metric.code is None:
            S = self.info()["n_segments"]
            table = np.zeros((_capi.SG_MAX_QUERY_TOKENS + 1, S), dtype=np.uint8)
            for a in range(1, _capi.SG_MAX_QUERY_TOKENS + 1):
                for B in range(max(int(metric.MinY(similarity, a)), 0), min(int(metric.MaxY(similarity, a)), S - 1) + 1):
                    table[a, B] = min(max(int(metric.Threshold(similarity, a, B)), 0), 255)
        cap = int(cap) if cap is not None else max(4 * n_q, 1024)
        size_a = np.zeros(n_q, dtype=np.uint32)
        while True:
            bufs = [np.zeros(cap, dtype=np.uint32) for _ in range(4)]
            total = C.c_uint64(0)
            rc = _capi.lib().sg_candidates_batch(self.handle, _ptr(data), _ptr(off), n_q, metric.code if table is None else 0,
                                                 float(similarity), _ptr(table), cap, *[_ptr(b) for b in bufs], C.byref(total),
                                                 _ptr(size_a))
            _capi.check(rc)
            if total.value <= cap:
                n = int(total.value)
                return bufs[0][:n], bufs[1][:n], bufs[2][:n], bufs[3][:n], size_a
            cap = int(total.value)

    def SuggestBatch(self, queries, similarity, metric, topK, packed=None, out=None):
        """Batched Suggest through sg_search_batch (host buffers).

        Returns (ids[n_q, k] uint32, scores[n_q, k] float64, counts[n_q] uint32); row q holds counts[q]
        candidates ordered (score desc, id asc)."""
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_q = len(off) - 1
        k = int(topK)
        if out is None:
            ids = np.zeros((n_q, max(k, 0)), dtype=np.uint32)
            scores = np.zeros((n_q, max(k, 0)), dtype=np.float64)
            counts = np.zeros(n_q, dtype=np.uint32)
        else:
            ids, scores, counts = out
        rc = _capi.lib().sg_search_batch(self.handle, _ptr(data), _ptr(off), n_q, metric.code, float(similarity),
                                         max(k, 0), _ptr(ids), _ptr(scores), _ptr(counts))
        _capi.check(rc)
        return ids, scores, counts

    def SuggestBatchCandidates(self, queries, similarity, metric, topK, packed=None, out=None):
        """sg_search_batch_candidates: the rows as suggest.Candidate lays them out (CANDIDATE_DTYPE: key, reserved, score;
        16 bytes).  Returns (rows[n_q, k], counts[n_q]); `out` = PinnedCandidateRows(n_q, k).out for the direct path."""
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_q, k = len(off) - 1, max(int(topK), 0)
        if out is None:
            rows = np.zeros((n_q, k), dtype=CANDIDATE_DTYPE)
            counts = np.zeros(n_q, dtype=np.uint32)
        else:
            rows, counts = out
        _capi.check(_capi.lib().sg_search_batch_candidates(self.handle, _ptr(data), _ptr(off), n_q, metric.code, float(similarity), k,
                                                           _ptr(rows), _ptr(counts)))
        return rows, counts

    def SubmitBatchCandidates(self, similarity, metric, topK, packed, out):
        """sg_search_batch_candidates_submit: SuggestBatchCandidates without waiting for it.  -> Ticket; `.wait()` returns
        (rows, counts).  `packed` = (bytes uint8, offsets uint32) and `out` = PinnedCandidateRows(...).out must stay
        untouched until then (they are passed as they are: no copies are made here)."""
        data, off = packed
        if data.dtype != np.uint8 or off.dtype != np.uint32 or not data.flags.c_contiguous or not off.flags.c_contiguous:
            raise ValueError("SubmitBatchCandidates takes contiguous uint8 bytes and uint32 offsets")
        rows, counts = out
        ticket = C.c_void_p()
        _capi.check(_capi.lib().sg_search_batch_candidates_submit(self.handle, _ptr(data), _ptr(off), len(off) - 1, metric.code, float(similarity),
                                                                  max(int(topK), 0), _ptr(rows), _ptr(counts), C.byref(ticket)))
        return Ticket(ticket, (data, off, rows, counts))

    # -- Autocomplete ---------------------------------------------------------------------------
    def Autocomplete(self, query, limit) -> List[Candidate]:
        """nGramAutocomplete.Autocomplete (pkg/suggest/autocomplete.go:40-77) with a FirstKCollectorManager(limit)."""
        ids, scores, counts = self.AutocompleteBatch([query], limit)
        return [Candidate(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def AutocompleteBatch(self, queries, limit, packed=None, out=None):
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_q = len(off) - 1
        k = max(int(limit), 0)
        if out is None:
            ids = np.zeros((n_q, k), dtype=np.uint32)
            scores = np.zeros((n_q, k), dtype=np.float64)
            counts = np.zeros(n_q, dtype=np.uint32)
        else:
            ids, scores, counts = out
        _capi.check(_capi.lib().sg_autocomplete_batch(self.handle, _ptr(data), _ptr(off), n_q, k, _ptr(ids), _ptr(scores),
                                                      _ptr(counts)))
        return ids, scores, counts

    def SuggestBatchDevice(self, d_q_bytes, d_q_off, n_q, similarity, metric, topK, d_ids, d_scores, d_counts,
                           d_stats=0, stream=0):
        """sg_search_batch_device: every argument is a device pointer (int); asynchronous on `stream`."""
        rc = _capi.lib().sg_search_batch_device(self.handle, d_q_bytes, d_q_off, n_q, metric.code, float(similarity),
                                                int(topK), d_ids, d_scores, d_counts, d_stats or None, stream or None)
        _capi.check(rc)


    def StageTimes(self, d_q_bytes, d_q_off, n_q, similarity, metric, topK, d_ids, d_scores, d_counts, stream=0):
        """sg_search_stage_times: {kernel name: ms} of one device-resident launch (CUDA events between its kernels)."""
        ms = (C.c_float * 8)()
        names = C.create_string_buffer(256)
        n = _capi.check(_capi.lib().sg_search_stage_times(self.handle, d_q_bytes, d_q_off, n_q, metric.code, float(similarity),
                                                          int(topK), d_ids, d_scores, d_counts, stream or None, ms, names, 256))
        return dict(zip(names.value.decode().split(","), [float(ms[i]) for i in range(n)]))


    def AutocompleteBatchDevice(self, d_q_bytes, d_q_off, n_q, limit, d_ids, d_scores, d_counts, d_stats=0, stream=0):
        """sg_autocomplete_batch_device: every argument is a device pointer (int); asynchronous on `stream`."""
        _capi.check(_capi.lib().sg_autocomplete_batch_device(self.handle, d_q_bytes, d_q_off, n_q, int(limit), d_ids, d_scores, d_counts,
                                                             d_stats or None, stream or None))

    def AutocompleteStageTimes(self, d_q_bytes, d_q_off, n_q, limit, d_ids, d_scores, d_counts, stream=0):
        """sg_autocomplete_stage_times: {kernel name: ms} of one device-resident Autocomplete launch."""
        ms = (C.c_float * 8)()
        names = C.create_string_buffer(256)
        n = _capi.check(_capi.lib().sg_autocomplete_stage_times(self.handle, d_q_bytes, d_q_off, n_q, int(limit), d_ids, d_scores, d_counts,
                                                                stream or None, ms, names, 256))
        return dict(zip(names.value.decode().split(","), [float(ms[i]) for i in range(n)]))


class Batcher:
    """sg_batcher_*: the micro-batcher in front of sg_search_batch for callers that issue ONE query per thread, as the
    reference's do (internal/suggest/api/suggest_handler.go:42-76: one goroutine per HTTP request).  Suggest blocks the
    calling thread; any number of threads may call it concurrently (ctypes releases the GIL for the duration)."""

    def __init__(self, index, max_batch=16384, max_wait_us=100, max_k=256):
        self._index = index  # keeps the index alive: it must outlive the batcher
        self._h = C.c_void_p()
        _capi.check(_capi.lib().sg_batcher_create(index.handle, int(max_batch), int(max_wait_us), int(max_k), C.byref(self._h)))
        self.max_k = int(max_k)

    def Suggest(self, query, similarity, metric, topK) -> List[Candidate]:
        q = _b(query)
        k = int(topK)
        ids = (C.c_uint32 * max(k, 1))()
        scores = (C.c_double * max(k, 1))()
        count = C.c_uint32(0)
        _capi.check(_capi.lib().sg_suggest_one(self._h, q, len(q), metric.code, float(similarity), k, ids, scores, C.byref(count)))
        return [Candidate(int(ids[i]), float(scores[i])) for i in range(count.value)]

    def stats(self):
        st = _capi.SgBatcherStats()
        _capi.check(_capi.lib().sg_batcher_get_stats(self._h, C.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_ if name != "reserved"}

    def close(self):
        if self._h is not None and self._h.value:
            _capi.lib().sg_batcher_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def NewBatcher(index, max_batch=16384, max_wait_us=100, max_k=256):
    return Batcher(index, max_batch, max_wait_us, max_k)


def _replay(manager, metric, similarity, sizeA, n_segments, by_segment):
    """The host half of nGramSuggester.Suggest for a CollectorManager of the caller's own (suggester.go:46-131): segments
    in the reference's feed order sizeA, sizeA+1, sizeA-1, ... (:110-118; its five workers make the order of the
    manager.Collect calls nondeterministic, here it is the feed order), per admissible segment Create / SetScorer /
    Collect* / manager.Collect.  ErrCollectionTerminated ends a segment as it does inside the mergers
    (pkg/merger/cp_merge.go:109-111).  Thresholds below 1 skip the segment."""
    if sizeA == 0:
        return []  # suggester.go:49-51
    bMin, bMax = int(metric.MinY(similarity, sizeA)), min(int(metric.MaxY(similarity, sizeA)), n_segments - 1)
    i, j = sizeA, sizeA + 1
    feed = []
    while i >= bMin or j <= bMax:
        if i >= bMin:
            feed.append(i)
        if j <= bMax:
            feed.append(j)
        i, j = i - 1, j + 1
    for sizeB in feed:
        if sizeB < 0 or sizeB >= n_segments:
            continue  # indices.Get(sizeB) == nil
        T = int(metric.Threshold(similarity, sizeA, sizeB))
        if T <= 0 or T > sizeB or T > sizeA:
            continue
        c = manager.Create()
        c.SetScorer(_col.NewMetricScorer(metric, sizeA, sizeB))
        try:
            for cand in by_segment.get(sizeB, ()):
                c.Collect(cand)
        except _col.ErrCollectionTerminated:
            pass
        manager.Collect(c)
    return manager.GetCandidates()


class Builder:
    """suggest.Builder, pkg/suggest/ngram_index_builder.go:14-17"""

    def Build(self) -> NGramIndex:
        raise NotImplementedError


class _RAMBuilder(Builder):
    def __init__(self, dictionary, description, id_base=0):
        self.dictionary = dictionary
        self.description = description
        self.id_base = id_base

    def Build(self):
        d = self.dictionary
        if isinstance(d, tuple) and len(d) == 2 and isinstance(d[0], np.ndarray):
            data, off = d
        else:
            data, off = pack_strings(d, np.uint64)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        cfg, keep = self.description.c_config()
        h = C.c_void_p()
        rc = _capi.lib().sg_index_build(C.byref(cfg), _ptr(data), _ptr(off), len(off) - 1, self.id_base, C.byref(h))
        del keep
        if rc < 0:
            raise SuggestError(rc, "failed to build NGramIndex: " + (_capi.lib().sg_last_error() or b"").decode())
        return NGramIndex(h.value, self.description)


class _FSBuilder(Builder):
    def __init__(self, description):
        self.description = description

    def Build(self):
        d = self.description
        cfg, keep = d.c_config()
        h = C.c_void_p()
        base = d.GetIndexPath()
        rc = _capi.lib().sg_index_open_disk(C.byref(cfg), _b(os.path.join(base, d.getHeaderFile())),
                                            _b(os.path.join(base, d.getDocumentListFile())), C.byref(h))
        del keep
        if rc < 0:
            raise SuggestError(rc, "failed to open FS inverted index: " + (_capi.lib().sg_last_error() or b"").decode())
        return NGramIndex(h.value, d)


def NewRAMBuilder(dictionary, description, id_base=0) -> Builder:
    """suggest.NewRAMBuilder (pkg/suggest/ngram_index_builder.go:27-35).  `dictionary` is the list of values in id
    order (dictionary.Dictionary.Iterate) or an already packed (uint8 bytes, uint64 offsets) pair."""
    return _RAMBuilder(dictionary, description, id_base)


def NewFSBuilder(description) -> Builder:
    """suggest.NewFSBuilder (pkg/suggest/ngram_index_builder.go:38-57): `<output>/<name>.hd` + `.dl`"""
    return _FSBuilder(description)


class Service:
    """suggest.Service, pkg/suggest/service.go:18-139"""

    def __init__(self):
        self._lock = threading.RLock()
        self.indexes = {}
        self.dictionaries = {}

    def AddIndexByDescription(self, description):
        if description.Driver == RAMDriver:
            return self.AddRunTimeIndex(description)
        return self.AddOnDiscIndex(description)

    def AddRunTimeIndex(self, description):
        try:
            dictionary = open_ram_dictionary(description.GetSourcePath())
        except OSError as e:
            raise SuggestError(_capi.SG_ERR_IO, f"failed to create RAMDriver builder: {e}")
        return self.AddIndex(description.Name, dictionary, NewRAMBuilder(dictionary, description))

    def AddOnDiscIndex(self, description, dictionary=None):
        """The reference opens `<name>.cdb` for the values; the CDB reader stays on the Go side of the shim
        (SURVEY.md section 2 row 9), so here the values come from `dictionary` or the description's source file."""
        if dictionary is None:
            try:
                dictionary = open_ram_dictionary(description.GetSourcePath())
            except OSError as e:
                raise SuggestError(_capi.SG_ERR_IO, f"failed to create CDB dictionary: {e}")
        return self.AddIndex(description.Name, dictionary, NewFSBuilder(description))

    def AddIndex(self, name, dictionary, builder):
        try:
            index = builder.Build()
        except SuggestError as e:
            raise SuggestError(e.code, f"failed to build NGramIndex: {e}")
        with self._lock:
            self.indexes[name] = index  # an index being replaced stays alive until its searches drain (refcount)
            self.dictionaries[name] = dictionary

    def GetDictionaries(self):
        with self._lock:
            return list(self.dictionaries)

    def Suggest(self, dictName, config) -> List[ResultItem]:
        with self._lock:
            index = self.indexes.get(dictName)
            dictionary = self.dictionaries.get(dictName)
        if index is None or dictionary is None:
            raise KeyError(f"given dictionary {dictName} is not exists")
        candidates = index.Suggest(config.query, config.similarity, config.metric, config.topK)
        out = []
        for c in candidates:
            v = dictionary[c.Key]
            out.append(ResultItem(c.Score, v.decode("utf-8", "replace") if isinstance(v, bytes) else v))
        return out


    def _lookup(self, dictName):
        with self._lock:
            index = self.indexes.get(dictName)
            dictionary = self.dictionaries.get(dictName)
        if index is None or dictionary is None:
            raise KeyError(f"given dictionary {dictName} is not exists")
        return index, dictionary

    def Autocomplete(self, dictName, query, limit) -> List[ResultItem]:
        """Service.Autocomplete, pkg/suggest/service.go:141-172 (the reference reports score 0 for every item)"""
        index, dictionary = self._lookup(dictName)
        out = []
        for c in index.Autocomplete(query, limit):
            v = dictionary[c.Key]
            out.append(ResultItem(0.0, v.decode("utf-8", "replace") if isinstance(v, bytes) else v))
        return out


def NewService():
    return Service()
