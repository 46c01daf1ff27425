"""sg_candidates_batch and the host-driven half of Suggest for a CollectorManager / metric.Metric of the caller's own
(the two interface wrinkles).  Needs a B200: `pytest -m gpu`."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION, COLLECTION, TEST_DESCRIPTION
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200 import _capi
from suggest_b200 import collector as col
from suggest_b200.metric import Metric
from suggest_b200.suggest import IndexDescription, pack_strings
from suggest_b200.workload import synthetic_workload, unpack
from test_gpu_parity import METRICS, build_pair, description

pytestmark = pytest.mark.gpu


def perturbed(lines, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in rng.integers(0, len(lines), n):
        s = bytearray(lines[i])
        for _ in range(int(rng.integers(0, 3))):
            if s:
                s[int(rng.integers(0, len(s)))] = int(rng.integers(97, 123))
        out.append(bytes(s).decode("utf-8", "replace"))
    return out


def brute_force(ox, docs_tokens, query, threshold, window):
    """SURVEY 8(c) rule 5 from the tokens alone: overlap = sum over query tokens (with multiplicity) of [token in doc]"""
    qt = ox.tokenize(query)
    a = len(qt)
    out = {}
    if a == 0:
        return a, out
    lo, hi = window(a)
    for d, toks in enumerate(docs_tokens):
        b = len(toks)
        if b < lo or b > hi:
            continue
        T = threshold(a, b)
        if T <= 0 or T > a or T > b:
            continue
        st = set(toks)
        ov = sum(1 for t in qt if t in st)
        if ov >= T:
            out[d] = (ov, b)
    return a, out


def group(cq, cid, cov, cseg, n_q):
    rows = [dict() for _ in range(n_q)]
    for q, i, o, s in zip(cq.tolist(), cid.tolist(), cov.tolist(), cseg.tolist()):
        assert i not in rows[q], "a document is reported once per query"
        rows[q][i] = (o, s)
    return rows


@pytest.mark.parametrize("bshift", [None, 0, 2, 5])
def test_candidates_equal_brute_force_on_cars(cars_lines, bshift):
    env = {} if bshift is None else {"SG_BUCKET_SHIFT": bshift}
    gx, ox = build_pair(CARS_DESCRIPTION, cars_lines, env)
    docs_tokens = [ox.tokenize(d.decode("utf-8", "replace")) for d in cars_lines]
    queries = perturbed(cars_lines, 60, 11) + ["", "zz", "RAM RAM", "Nissan March"]
    for code, alpha in ((O.JACCARD, 0.5), (O.COSINE, 0.7), (O.DICE, 0.4), (O.OVERLAP, 0.9), (O.EXACT, 1.0)):
        m = METRICS[code]
        cq, cid, cov, cseg, size_a = gx.CandidatesBatch(queries, alpha, m)
        rows = group(cq, cid, cov, cseg, len(queries))
        S_ = gx.info()["n_segments"]
        for q, text in enumerate(queries):
            a, want = brute_force(ox, docs_tokens, text, lambda a_, b_: m.Threshold(alpha, a_, b_),
                                  lambda a_: (m.MinY(alpha, a_), min(m.MaxY(alpha, a_), S_ - 1)))
            assert int(size_a[q]) == a
            assert rows[q] == want, (m, text, sorted(rows[q].items())[:5], sorted(want.items())[:5])
    gx.close()


def test_candidates_against_the_oracle_top_everything():
    # every candidate = the oracle's top-k with k above the candidate count; score from (overlap, sizeA, sizeB) on the host
    docs, (qb, qo), _ = synthetic_workload(30000, 400)
    desc = IndexDescription(Name="c", NGramSize=3)
    gx = S.NewRAMBuilder(docs, desc).Build()
    ox = O.OracleIndex(3, ("$", "$"), "$", ("english", "russian", "numbers", "$")).add_packed(*docs)
    k = 512
    for code, alpha in ((O.JACCARD, 0.35), (O.COSINE, 0.5), (O.DICE, 0.5)):
        m = METRICS[code]
        cq, cid, cov, cseg, size_a = gx.CandidatesBatch(None, alpha, m, packed=(qb, qo))
        rows = group(cq, cid, cov, cseg, len(qo) - 1)
        o_ids, o_sc, o_cnt = ox.suggest_batch(None, code, alpha, k, O.CANONICAL, threads=8, packed=(qb, qo.astype(np.uint64)))
        assert int(o_cnt.max()) < k
        for q in range(len(qo) - 1):
            want = {int(o_ids[q, i]): float(o_sc[q, i]) for i in range(int(o_cnt[q]))}
            got = {i: 1 - m.Distance(ov, int(size_a[q]), sb) for i, (ov, sb) in rows[q].items()}
            assert got == want
    gx.close()


class MyJaccard(Metric):
    """a metric.Metric the library does not know by type: same arithmetic as pkg/metric/jaccard.go"""
    name = "MyJaccard"

    def MinY(self, alpha, size):
        return int(math.ceil(alpha * float(size)))

    def MaxY(self, alpha, size):
        return int(math.floor(float(size) / alpha))

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(alpha * float(sizeA + sizeB) / (1 + alpha)))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / float(sizeA + sizeB - inter)


class Containment(Metric):
    """how much of the query a document holds: not one of the five built-ins"""
    name = "Containment"

    def MinY(self, alpha, size):
        return int(math.ceil(alpha * size))

    def MaxY(self, alpha, size):
        return 1 << 15

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(alpha * sizeA))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / float(sizeA)


def test_custom_metric_equals_builtin(cars_lines):
    gx, ox = build_pair(CARS_DESCRIPTION, cars_lines)
    queries = perturbed(cars_lines, 150, 5) + ["", "Nissan March", "RAM RAM", "Nissan МАРЧ", "ЖИГУЛИ 2107"]  # non-ASCII: host ToLower
    for k in (1, 5, 40):
        want = gx.SuggestMany(queries, 0.5, S.JaccardMetric(), k)
        got = gx.SuggestMany(queries, 0.5, MyJaccard(), k)
        assert got == want
    # the Go signature: a CollectorManagerFactory; FuzzyCollectorManager is recognised and stays on the device
    assert gx.SuggestMany(queries, 0.5, S.JaccardMetric(), col.NewFuzzyCollectorManager(5)) == gx.SuggestMany(queries, 0.5, S.JaccardMetric(), 5)
    gx.close()


def test_custom_metric_against_brute_force(cars_lines):
    gx, ox = build_pair(CARS_DESCRIPTION, cars_lines)
    docs_tokens = [ox.tokenize(d.decode("utf-8", "replace")) for d in cars_lines]
    m, alpha, k = Containment(), 0.8, 7
    queries = perturbed(cars_lines, 80, 9)
    got = gx.SuggestMany(queries, alpha, m, k)
    S_ = gx.info()["n_segments"]
    for text, g in zip(queries, got):
        a, cands = brute_force(ox, docs_tokens, text, lambda a_, b_: m.Threshold(alpha, a_, b_),
                               lambda a_: (m.MinY(alpha, a_), min(m.MaxY(alpha, a_), S_ - 1)))
        want = sorted(((1 - m.Distance(ov, a, b), d) for d, (ov, b) in cands.items()), key=lambda t: (-t[0], t[1]))[:k]
        assert [(c.Score, c.Key) for c in g] == want, text
    gx.close()


def test_custom_collector_manager(cars_lines):
    """A CollectorManager that is neither fuzzy nor first-k: counts candidates per segment and keeps the largest overlap."""

    class Census(col.CollectorManager):
        def __init__(self):
            self.best, self.n = {}, 0

        def Create(self):
            return col._FirstKCollector(1 << 30)

        def Collect(self, *cs):
            for c in cs:
                for it in c.items:
                    self.n += 1
                    self.best[it.Position()] = it.Overlap()

        def GetCandidates(self):
            return [col.Candidate(p, float(o)) for p, o in sorted(self.best.items())]

    gx, ox = build_pair(CARS_DESCRIPTION, cars_lines)
    docs_tokens = [ox.tokenize(d.decode("utf-8", "replace")) for d in cars_lines]
    queries = perturbed(cars_lines, 40, 2)
    m, alpha = S.DiceMetric(), 0.6
    got = gx.SuggestMany(queries, alpha, m, Census)
    S_ = gx.info()["n_segments"]
    for text, g in zip(queries, got):
        _, cands = brute_force(ox, docs_tokens, text, lambda a_, b_: m.Threshold(alpha, a_, b_),
                               lambda a_: (m.MinY(alpha, a_), min(m.MaxY(alpha, a_), S_ - 1)))
        assert [(c.Key, c.Score) for c in g] == [(d, float(ov)) for d, (ov, _) in sorted(cands.items())]
    # FirstKCollectorManager through Suggest: the first `limit` candidates of every segment, lowest ids kept
    first = gx.SuggestMany(queries[:10], alpha, m, col.NewFirstKCollectorManager(3))
    for text, g in zip(queries[:10], first):
        assert len(g) <= 3 and [c.Score for c in g] == [-float(c.Key) for c in g]
    gx.close()


def test_capacity_protocol_and_errors():
    gx = S.NewRAMBuilder(COLLECTION, description(TEST_DESCRIPTION)).Build()
    data, off = pack_strings(["Nissan ma", "Toyota"])
    bufs = [np.zeros(2, dtype=np.uint32) for _ in range(4)]
    size_a = np.zeros(2, dtype=np.uint32)
    total = C.c_uint64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = _capi.lib().sg_candidates_batch(gx.handle, p(data), p(off), 2, _capi.SG_JACCARD, 0.3, None, 2, *[p(b) for b in bufs],
                                         C.byref(total), p(size_a))
    assert rc == 0 and total.value > 2            # found more than the buffers hold: only the first two were written
    assert size_a.tolist() == [9, 6]
    full = gx.CandidatesBatch(["Nissan ma", "Toyota"], 0.3, S.JaccardMetric(), cap=1)  # the mirror retries with the reported size
    assert len(full[0]) == total.value
    rc = _capi.lib().sg_candidates_batch(gx.handle, p(data), p(off), 2, 9, 0.3, None, 2, *[p(b) for b in bufs], C.byref(total), p(size_a))
    assert rc == _capi.SG_ERR_INVALID
    long_q = "".join(chr(c) for c in np.random.default_rng(1).integers(97, 123, 300))  # ~300 distinct 3-grams
    with pytest.raises(S.SuggestError) as e:
        gx.CandidatesBatch([long_q], 0.5, S.JaccardMetric())
    assert e.value.code == _capi.SG_ERR_QUERY_TOO_LONG
    gx.close()
