"""The C-ABI library loads and exports every symbol include/suggest_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

from conftest import ROOT
from suggest_b200 import _capi, build


def header_functions():
    with open(os.path.join(ROOT, "include", "suggest_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    L = C.CDLL(_capi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 19
    for name in names:
        assert hasattr(L, name), name
    assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)


def test_version_and_error_strings_are_callable_without_gpu():
    L = _capi.lib()
    assert b"sm_100a" in L.sg_version()
    assert L.sg_last_error() is not None
    assert L.sg_kernel_launches() == 0


def test_argument_validation_happens_before_any_cuda_call():
    L = _capi.lib()
    # k = 0 and similarity out of range are rejected like suggest.NewSearchConfig (pkg/suggest/search.go:18-25)
    assert L.sg_search_batch(None, None, None, 0, 0, 0.5, 10, None, None, None) == _capi.SG_ERR_INVALID
    assert L.sg_merge_topk_device(0, 0, 1, 1, None, None, None, None, None, None, None) == _capi.SG_ERR_INVALID
    assert L.sg_merge_topk_device(0, 2, 1, 0, None, None, None, None, None, None, None) == _capi.SG_ERR_INVALID
