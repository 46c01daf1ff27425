"""ctypes binding of libsuggest_b200.so (include/suggest_b200.h).

The library is the product: if it is missing this module raises, there is no fallback of any kind.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SUGGEST_B200_LIB") or os.path.join(_HERE, "libsuggest_b200.so")  # override: tuning builds

SG_OK = 0
SG_ERR_INVALID, SG_ERR_UNSUPPORTED, SG_ERR_CUDA, SG_ERR_NOMEM, SG_ERR_QUERY_TOO_LONG, SG_ERR_IO, SG_ERR_FORMAT = \
    -1, -2, -3, -4, -5, -6, -7
SG_JACCARD, SG_COSINE, SG_DICE, SG_OVERLAP, SG_EXACT = range(5)
SG_MAX_QUERY_TOKENS = 128
SG_MAX_TOPK = 16384
SG_MAX_TOPK_SHARED = 1024
SG_EXCHANGE_HANDLE_BYTES = 64
SG_COUNT_UNSUPPORTED = 0xFFFFFFFF


class SgConfig(C.Structure):
    _fields_ = [("ngram_size", C.c_int32), ("wrap_start", C.c_char_p), ("wrap_end", C.c_char_p), ("pad", C.c_char_p),
                ("alphabet", C.POINTER(C.c_char_p)), ("n_alphabet", C.c_int32), ("device", C.c_int32)]


class SgIndexInfo(C.Structure):
    _fields_ = [("n_docs", C.c_uint32), ("n_segments", C.c_uint32), ("n_terms", C.c_uint32), ("n_lists", C.c_uint64),
                ("n_postings", C.c_uint64), ("device_bytes", C.c_uint64), ("id_base", C.c_uint32), ("device", C.c_int32)]


class SgBatcherStats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("queries", C.c_uint64), ("largest_batch", C.c_uint32), ("max_batch", C.c_uint32),
                ("max_wait_us", C.c_uint32), ("reserved", C.c_uint32)]


class SgIndexLayout(C.Structure):
    _fields_ = [("n_slots", C.c_uint32), ("bucket_shift", C.c_uint32), ("row_words", C.c_uint32), ("engine", C.c_uint32),
                ("built_on_device", C.c_uint32), ("pipeline", C.c_uint32), ("bitmap_bytes", C.c_uint64)]


# every symbol include/suggest_b200.h declares, with its signature
SIGNATURES = {
    "sg_index_build": (C.c_int, [C.POINTER(SgConfig), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sg_index_from_lists": (C.c_int, [C.POINTER(SgConfig), C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "sg_index_open_disk": (C.c_int, [C.POINTER(SgConfig), C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "sg_index_free": (None, [C.c_void_p]),
    "sg_index_get_info": (C.c_int, [C.c_void_p, C.POINTER(SgIndexInfo)]),
    "sg_index_get_layout": (C.c_int, [C.c_void_p, C.POINTER(SgIndexLayout)]),
    "sg_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_search_batch_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                             C.c_void_p, C.c_void_p]),
    "sg_search_batch_candidates_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "sg_ticket_wait": (C.c_int, [C.c_void_p]),
    "sg_autocomplete_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "sg_candidates_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "sg_search_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_search_stage_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_char_p, C.c_uint32]),
    "sg_autocomplete_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_autocomplete_stage_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_char_p, C.c_uint32]),
    "sg_merge_topk_device": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_packed_rows_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "sg_search_batch_packed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                                C.c_void_p, C.c_void_p]),
    "sg_merge_topk_packed_device": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "sg_sharded_build": (C.c_int, [C.POINTER(SgConfig), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sg_sharded_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_sharded_get_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "sg_sharded_shard": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "sg_sharded_free": (None, [C.c_void_p]),
    "sg_exchange_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sg_exchange_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sg_exchange_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sg_exchange_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_exchange_result": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "sg_exchange_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sg_exchange_free": (None, [C.c_void_p]),
    "sg_batcher_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sg_suggest_one": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_uint32)]),
    "sg_batcher_get_stats": (C.c_int, [C.c_void_p, C.POINTER(SgBatcherStats)]),
    "sg_batcher_free": (None, [C.c_void_p]),
    "sg_lm_create": (C.c_int, [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "sg_lm_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "sg_lm_free": (None, [C.c_void_p]),
    "sg_lm_score_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "sg_lm_score_next_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sg_predict_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                   C.c_uint32, C.c_void_p, C.c_void_p]),
    "sg_pinned_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "sg_pinned_free": (None, [C.c_void_p]),
    "sg_is_pinned": (C.c_int, [C.c_void_p, C.c_uint64]),
    "sg_kernel_launches": (C.c_uint64, []),
    "sg_last_error": (C.c_char_p, []),
    "sg_version": (C.c_char_p, []),
    "sg_host_tokenize": (C.c_int, [C.POINTER(SgConfig), C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]),
    "sg_host_to_lower": (C.c_int, [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]),
    "sg_host_index_build": (C.c_int, [C.POINTER(SgConfig), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sg_host_index_open_disk": (C.c_int, [C.POINTER(SgConfig), C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "sg_host_index_free": (None, [C.c_void_p]),
    "sg_host_index_get_info": (C.c_int, [C.c_void_p, C.POINTER(SgIndexInfo)]),
    "sg_host_index_get_list": (C.c_int64, [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "sg_host_index_get_layout": (C.c_int, [C.c_void_p, C.POINTER(SgIndexLayout)]),
    "sg_host_index_get_segments": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "sg_host_index_get_list_slots": (C.c_int64, [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "sg_host_index_get_bitmap": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "sg_host_last_error": (C.c_char_p, []),
}

_lib = None


def lib():
    """Load libsuggest_b200.so.  Raises if it has not been built (python -m suggest_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libsuggest_b200.so is missing: build it with `python -m suggest_b200.build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the Suggest path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class SuggestError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def check(rc, host=False):
    if rc < 0:
        msg = (lib().sg_host_last_error() if host else lib().sg_last_error()) or b""
        raise SuggestError(rc, msg.decode("utf-8", "replace"))
    return rc


def make_config(ngram_size, wrap, pad, alphabet, device=0):
    """-> (SgConfig, keepalive)"""
    enc = [a.encode("utf-8") if isinstance(a, str) else bytes(a) for a in alphabet]
    arr = (C.c_char_p * max(len(enc), 1))(*enc)
    w0, w1 = (w.encode("utf-8") if isinstance(w, str) else bytes(w) for w in wrap)
    p = pad.encode("utf-8") if isinstance(pad, str) else bytes(pad)
    cfg = SgConfig(ngram_size, w0, w1, p, arr, len(enc), device)
    return cfg, (arr, enc, w0, w1, p)
