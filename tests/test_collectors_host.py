"""Host mirror of the collector interfaces and of the built-in metrics' arithmetic (no GPU): the pieces NGramIndex.Suggest
uses when a caller brings its own CollectorManager / metric.Metric."""
import numpy as np

from oracle import oracle as O
from suggest_b200 import collector as col
from suggest_b200 import metric as M
from suggest_b200.suggest import _replay

PAIRS = [(O.JACCARD, M.JaccardMetric()), (O.COSINE, M.CosineMetric()), (O.DICE, M.DiceMetric()), (O.OVERLAP, M.OverlapMetric()),
         (O.EXACT, M.ExactMetric())]


def test_host_metrics_equal_the_oracle():
    # pkg/metric/*.go restated twice (C oracle, Python host mirror): same integers, bit-equal doubles
    for code, m in PAIRS:
        for alpha in (0.3, 0.5, 0.55, 0.7, 0.9, 1.0):
            for a in range(1, 60, 3):
                assert m.MinY(alpha, a) == O.metric_min_y(code, alpha, a)
                assert m.MaxY(alpha, a) == O.metric_max_y(code, alpha, a)
                for b in range(1, 60, 2):
                    assert m.Threshold(alpha, a, b) == O.metric_threshold(code, alpha, a, b), (m, alpha, a, b)
                    for c in range(1, min(a, b) + 1, 2):
                        assert m.Distance(c, a, b) == O.metric_distance(code, c, a, b)
                        assert col.NewMetricScorer(m, a, b).Score(col.MergeCandidate(7, c)) == O.score(code, c, a, b)


def test_topk_queue_kat():
    # pkg/suggest/topk_test.go:10-39 through the oracle's restatement
    rng = np.random.default_rng(3)
    for k in (1, 3, 10):
        pairs = [(int(i), float(s)) for i, s in zip(rng.permutation(200), rng.integers(0, 20, 200) / 20.0)]
        q = col.TopKQueue(k)
        for pos, score in pairs:
            q.Add(pos, score)
        got = [(c.Key, c.Score) for c in q.GetCandidates()]
        want = sorted(pairs, key=lambda p: (-p[1], p[0]))[:k]
        assert got == want
        got_o, low = O.topk(pairs, k)
        assert got_o == want and low == want[-1][1]
        assert q.IsFull() and q.GetLowestScore() == want[-1][1]
    assert col.TopKQueue(4).GetLowestScore() == float("-inf")


def test_replay_feed_order_and_termination():
    seen = []

    class Recording(col.CollectorManager):
        def __init__(self):
            self.all = []

        def Create(self):
            return col._FirstKCollector(2)

        def Collect(self, *cs):
            for c in cs:
                seen.append([x.Position() for x in c.items])
                self.all += c.items

        def GetCandidates(self):
            return [col.Candidate(x.Position(), float(x.Overlap())) for x in self.all]

    by_segment = {4: [col.MergeCandidate(i, 3) for i in (5, 6, 7)], 5: [col.MergeCandidate(9, 4)], 7: [col.MergeCandidate(1, 4)]}
    out = _replay(Recording(), M.JaccardMetric(), 0.5, 5, 8, by_segment)
    # Jaccard 0.5, sizeA 5: window [3, 10] clipped to 7 segments; feed order 5, 6, 4, 7, 3 (suggester.go:110-118)
    assert seen == [[9], [], [5, 6], [1], []]
    assert [c.Key for c in out] == [9, 5, 6, 1]
    assert _replay(Recording(), M.JaccardMetric(), 0.5, 0, 8, {}) == []


def test_fuzzy_and_first_k_managers():
    f = col.NewFuzzyCollectorManager(2)()
    for seg, items in ((4, [(1, 2), (2, 4)]), (5, [(3, 4), (0, 5)])):
        c = f.Create()
        c.SetScorer(col.NewMetricScorer(M.JaccardMetric(), 5, seg))
        for pos, ov in items:
            c.Collect(col.MergeCandidate(pos, ov))
        f.Collect(c)
    got = f.GetCandidates()
    assert [c.Key for c in got] == [0, 2] and got[0].Score == 1.0 and got[1].Score == O.score(O.JACCARD, 4, 5, 4)
    assert f.GetLowestScore() == got[1].Score
    k = col.NewFirstKCollectorManager(3)()
    c = k.Create()
    for pos in (4, 8, 9):
        c.Collect(col.MergeCandidate(pos, 1))
    try:
        c.Collect(col.MergeCandidate(10, 1))
        raise AssertionError("expected ErrCollectionTerminated")
    except col.ErrCollectionTerminated:
        pass
    k.Collect(c)
    assert [(x.Key, x.Score) for x in k.GetCandidates()] == [(4, -4.0), (8, -8.0), (9, -9.0)]
