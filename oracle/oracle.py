"""ctypes front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs.  Nothing under suggest_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

JACCARD, COSINE, DICE, OVERLAP, EXACT = range(5)
SCAN_COUNT, CP_MERGE, MERGE_SKIP, DIVIDE_SKIP = range(4)
CANONICAL, FAITHFUL = 0, 1
CODEC_VB, CODEC_SKIPPING, CODEC_BINARY = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def build(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.so_index_new.restype = C.c_void_p
        L.so_index_new.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
        L.so_index_add_docs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.so_index_commit.argtypes = [C.c_void_p]
        L.so_index_free.argtypes = [C.c_void_p]
        L.so_index_segments.argtypes = [C.c_void_p]
        L.so_index_segments.restype = C.c_uint32
        L.so_index_postings.argtypes = [C.c_void_p]
        L.so_index_postings.restype = C.c_uint64
        L.so_index_lists.argtypes = [C.c_void_p]
        L.so_index_lists.restype = C.c_uint64
        L.so_index_get_list.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64]
        L.so_index_get_list.restype = C.c_int64
        L.so_index_list_at.argtypes = [C.c_void_p, C.c_uint64, _u32p, C.POINTER(C.c_void_p), _u32p,
                                       C.POINTER(C.c_void_p), _u32p]
        L.so_ngram_tokenize.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_char_p, C.c_uint32, _u32p, C.c_uint32]
        L.so_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, _u32p, C.c_uint32]
        L.so_alphabet_has.argtypes = [C.c_void_p, C.c_uint32]
        L.so_to_lower.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]
        for f in (L.so_metric_min_y, L.so_metric_max_y):
            f.argtypes = [C.c_int, C.c_double, C.c_int]
        L.so_metric_threshold.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int]
        for f in (L.so_metric_distance, L.so_score):
            f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
            f.restype = C.c_double
        L.so_merge.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]
        L.so_merge.restype = C.c_int64
        L.so_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
        L.so_intersect.restype = C.c_int64
        L.so_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
        L.so_encode.restype = C.c_int64
        L.so_decode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        L.so_decode.restype = C.c_int64
        L.so_posting_lower_bound_tail.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, _u32p,
                                                  C.POINTER(C.c_int), C.c_void_p, C.c_uint32]
        L.so_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, _f64p]
        L.so_suggest.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]
        L.so_suggest_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double,
                                       C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.so_query_stats.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_double, _u64p, _u64p, _u32p,
                                     _u32p]
        L.so_has_duplicate_tokens.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.so_autocomplete.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _b(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def pack_strings(strings):
    """list of str/bytes -> (uint8 array, uint64 offsets[n+1])"""
    bs = [_b(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(x) for x in bs], dtype=np.uint64)
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return data, off


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleIndex:
    """pkg/suggest IndexDescription + NewRAMBuilder(...).Build() on the CPU."""

    def __init__(self, ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("english", "russian", "numbers", "$")):
        L = lib()
        arr = (C.c_char_p * len(alphabet))(*[_b(a) for a in alphabet])
        self._h = L.so_index_new(ngram_size, _b(wrap[0]), _b(wrap[1]), _b(pad), arr, len(alphabet))
        if not self._h:
            raise ValueError("bad index description")
        self.n_docs = 0

    def close(self):
        if self._h:
            lib().so_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_docs(self, docs):
        data, off = pack_strings(docs)
        return self.add_packed(data, off)

    def add_packed(self, data, off):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        rc = lib().so_index_add_docs(self._h, _ptr(data), _ptr(off), len(off) - 1)
        assert rc == 0
        self.n_docs += len(off) - 1
        return self

    def commit(self):
        assert lib().so_index_commit(self._h) == 0
        return self

    @property
    def segments(self):
        return lib().so_index_segments(self._h)

    @property
    def postings(self):
        return lib().so_index_postings(self._h)

    @property
    def lists(self):
        return lib().so_index_lists(self._h)

    def get_list(self, segment, term):
        t = _b(term)
        n = lib().so_index_get_list(self._h, segment, t, len(t), None, 0)
        if n < 0:
            return None
        out = np.zeros(n, dtype=np.uint32)
        lib().so_index_get_list(self._h, segment, t, len(t), _ptr(out), n)
        return out

    def iter_lists(self):
        L = lib()
        seg, tl, n = C.c_uint32(), C.c_uint32(), C.c_uint32()
        tp, ip = C.c_void_p(), C.c_void_p()
        for i in range(self.lists):
            L.so_index_list_at(self._h, i, C.byref(seg), C.byref(tp), C.byref(tl), C.byref(ip), C.byref(n))
            term = C.string_at(tp.value, tl.value)
            ids = np.ctypeslib.as_array(C.cast(ip.value, _u32p), shape=(n.value,)).copy()
            yield seg.value, term, ids

    def tokenize(self, text):
        t = _b(text)
        cap = 4 * (len(t) + 64) * 8 + 64
        out = C.create_string_buffer(cap)
        off = (C.c_uint32 * (len(t) + 66))()
        n = lib().so_tokenize(self._h, t, len(t), out, cap, off, len(t) + 65)
        assert n >= 0
        return [out.raw[off[i]:off[i + 1]] for i in range(n)]

    def has(self, rune):
        return bool(lib().so_alphabet_has(self._h, ord(rune) if isinstance(rune, str) else rune))

    def has_duplicate_tokens(self, text):
        t = _b(text)
        return bool(lib().so_has_duplicate_tokens(self._h, t, len(t)))

    def suggest(self, query, metric, alpha, k, mode=CANONICAL, merger=CP_MERGE):
        q = _b(query)
        ids = np.zeros(k, dtype=np.uint32)
        scores = np.zeros(k, dtype=np.float64)
        n = lib().so_suggest(self._h, q, len(q), metric, alpha, k, mode, merger, _ptr(ids), _ptr(scores))
        if n < 0:
            raise RuntimeError("oracle suggest failed")
        return ids[:n].copy(), scores[:n].copy()

    def autocomplete(self, query, limit):
        q = _b(query)
        ids = np.zeros(limit, dtype=np.uint32)
        scores = np.zeros(limit, dtype=np.float64)
        n = lib().so_autocomplete(self._h, q, len(q), limit, _ptr(ids), _ptr(scores))
        if n < 0:
            raise RuntimeError("oracle autocomplete failed")
        return ids[:n].copy(), scores[:n].copy()

    def suggest_batch(self, queries, metric, alpha, k, mode=CANONICAL, merger=CP_MERGE, threads=1, packed=None):
        data, off = packed if packed is not None else pack_strings(queries)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        nq = len(off) - 1
        ids = np.zeros((nq, k), dtype=np.uint32)
        scores = np.zeros((nq, k), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.uint32)
        rc = lib().so_suggest_batch(self._h, _ptr(data), _ptr(off), nq, metric, alpha, k, mode, merger, threads,
                                    _ptr(ids), _ptr(scores), _ptr(counts))
        if rc != 0:
            raise RuntimeError("oracle batch failed")
        return ids, scores, counts

    def query_stats(self, query, metric, alpha):
        q = _b(query)
        p, l = C.c_uint64(), C.c_uint64()
        s, a = C.c_uint32(), C.c_uint32()
        lib().so_query_stats(self._h, q, len(q), metric, alpha, C.byref(p), C.byref(l), C.byref(s), C.byref(a))
        return dict(postings=p.value, lists=l.value, segments=s.value, size_a=a.value)


# ---- free functions used by the known-answer tests ----
def ngram_tokenize(text, n):
    t = _b(text)
    cap = len(t) * (n + 1) * 4 + 64
    out = C.create_string_buffer(cap)
    off = (C.c_uint32 * (len(t) + 2))()
    cnt = lib().so_ngram_tokenize(t, len(t), n, out, cap, off, len(t) + 1)
    assert cnt >= 0
    return [out.raw[off[i]:off[i + 1]] for i in range(cnt)]


def to_lower(text):
    t = _b(text)
    out = C.create_string_buffer(3 * len(t) + 8)
    n = lib().so_to_lower(t, len(t), out, 3 * len(t) + 8)
    return out.raw[:n]


def _flat(lists):
    off = np.zeros(len(lists) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(x) for x in lists])
    ids = np.array([v for l in lists for v in l], dtype=np.uint32)
    return ids, off


def merge(algo, lists, threshold):
    ids, off = _flat(lists)
    out = np.zeros(int(off[-1]) + 1, dtype=np.uint64)
    n = lib().so_merge(algo, _ptr(ids), _ptr(off), len(lists), threshold, _ptr(out), len(out))
    assert n >= 0
    return [(int(c >> 32), int(c & 0xFFFFFFFF)) for c in out[:n]]


def intersect(lists):
    ids, off = _flat(lists)
    out = np.zeros(int(off[-1]) + 1, dtype=np.uint64)
    n = lib().so_intersect(_ptr(ids), _ptr(off), len(lists), _ptr(out), len(out))
    assert n >= 0
    return [int(c >> 32) for c in out[:n]]


def encode(codec, lst, gap=64):
    a = np.asarray(lst, dtype=np.uint32)
    out = np.zeros(len(a) * 7 + 16, dtype=np.uint8)
    n = lib().so_encode(codec, gap, _ptr(a), len(a), _ptr(out), len(out))
    assert n >= 0
    return out[:n].tobytes()


def decode(codec, data, n, gap=64):
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(n, dtype=np.uint32)
    got = lib().so_decode(codec, gap, _ptr(buf), len(buf), _ptr(out), n)
    assert got == n, (got, n)
    return out


def posting_lower_bound_tail(kind, lst, to, gap=64):
    a = np.asarray(lst, dtype=np.uint32)
    tail = np.zeros(len(a) + 1, dtype=np.uint32)
    lb, err = C.c_uint32(), C.c_int()
    n = lib().so_posting_lower_bound_tail(kind, gap, _ptr(a), len(a), to, C.byref(lb), C.byref(err), _ptr(tail),
                                          len(tail))
    assert n >= 0
    return lb.value, bool(err.value), tail[:n].tolist()


def topk(pairs, k):
    ids = np.array([p[0] for p in pairs], dtype=np.uint32)
    sc = np.array([p[1] for p in pairs], dtype=np.float64)
    oi = np.zeros(k, dtype=np.uint32)
    os_ = np.zeros(k, dtype=np.float64)
    low = C.c_double()
    n = lib().so_topk(_ptr(ids), _ptr(sc), len(ids), k, _ptr(oi), _ptr(os_), C.byref(low))
    return [(int(oi[i]), float(os_[i])) for i in range(n)], low.value


def metric_min_y(m, alpha, size):
    return lib().so_metric_min_y(m, alpha, size)


def metric_max_y(m, alpha, size):
    return lib().so_metric_max_y(m, alpha, size)


def metric_threshold(m, alpha, a, b):
    return lib().so_metric_threshold(m, alpha, a, b)


def metric_distance(m, inter, a, b):
    return lib().so_metric_distance(m, inter, a, b)


def score(m, inter, a, b):
    return lib().so_score(m, inter, a, b)
