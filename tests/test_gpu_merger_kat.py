"""The reference's own T-occurrence known-answer tests (pkg/merger/list_merger_test.go:48-140: seven cases, exact
overlap -> ids maps) run on the GPU: the posting lists go in through sg_index_from_lists, the threshold through a
tabulated metric, and sg_candidates_batch must return exactly the (id, overlap) pairs the Go mergers are tested for.
Needs a B200: `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest

import suggest_b200 as S
from suggest_b200 import _capi
from suggest_b200.metric import Metric
from suggest_b200.suggest import IndexDescription
from test_oracle_kat import MERGE_CASES

pytestmark = pytest.mark.gpu

SEGMENT = 5  # sizeB of every list: at least the largest threshold of the cases
TERMS = [b"aaa", b"bbb", b"ccc", b"ddd", b"eee"]


class Fixed(Metric):
    """Threshold T for every (sizeA, sizeB): what the merger tests pass as `threshold`"""
    name = "Fixed"

    def __init__(self, t):
        self.t = t

    def MinY(self, alpha, size):
        return 1

    def MaxY(self, alpha, size):
        return 1 << 15

    def Threshold(self, alpha, sizeA, sizeB):
        return self.t

    def Distance(self, inter, sizeA, sizeB):
        return 0.0


def index_of(rid):
    desc = IndexDescription(Name="kat", NGramSize=3, Alphabet=("english", "$"))
    segs = np.full(len(rid), SEGMENT, dtype=np.uint32)
    ids = np.concatenate([np.asarray(lst, dtype=np.uint32) for lst in rid])
    loff = np.zeros(len(rid) + 1, dtype=np.uint64)
    loff[1:] = np.cumsum([len(lst) for lst in rid])
    toff = np.arange(len(rid) + 1, dtype=np.uint64) * 3
    tb = np.frombuffer(b"".join(TERMS[:len(rid)]), dtype=np.uint8)
    cfg, keep = desc.c_config()
    h = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _capi.check(_capi.lib().sg_index_from_lists(C.byref(cfg), SEGMENT + 1, len(rid), p(segs), p(tb), p(toff), p(ids), p(loff), C.byref(h)))
    del keep
    return S.NGramIndex(h.value, desc)


@pytest.mark.parametrize("case", range(len(MERGE_CASES)))
def test_list_merger_test_go_on_the_gpu(case):
    rid, t, expected = MERGE_CASES[case]
    gx = index_of(rid)
    # the query's n-grams include every term of the index once ("aaabbb..." -> $aa aaa aab abb bbb ...); the others are
    # in no list and only count towards len(tokens)
    query = b"".join(TERMS[:len(rid)]).decode()
    cq, cid, cov, cseg, size_a = gx.CandidatesBatch([query], 0.5, Fixed(t))
    assert int(size_a[0]) == 3 * len(rid)
    assert set(cseg.tolist()) <= {SEGMENT} and set(cq.tolist()) <= {0}
    actual = {}
    for pos, overlap in sorted(zip(cid.tolist(), cov.tolist())):
        actual.setdefault(overlap, []).append(pos)
    assert actual == expected
    gx.close()
