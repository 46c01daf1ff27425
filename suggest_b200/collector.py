"""Host mirror of the reference's collector interfaces (pkg/merger/collector.go, pkg/suggest/{collector,scorer,topk}.go).

The device serves FuzzyCollectorManager(topK) and FirstKCollectorManager(limit) itself (sg_search_batch,
sg_autocomplete_batch).  A CollectorManager of the caller's own gets every candidate of the T-occurrence count from
sg_candidates_batch and is driven from here exactly as nGramSuggester.Suggest drives it (suggester.go:66-108): per
admissible segment Create() -> SetScorer(NewMetricScorer) -> Collect(candidate)* -> manager.Collect(collector).
"""
import heapq
from dataclasses import dataclass


class ErrCollectionTerminated(Exception):
    """merger.ErrCollectionTerminated (pkg/merger/collector.go:5-7): swallowed by whoever feeds the collector."""


@dataclass(frozen=True)
class MergeCandidate:
    """merger.MergeCandidate (pkg/merger/list_merger.go:33-48): position + overlap"""
    position: int
    overlap: int

    def Position(self):
        return self.position

    def Overlap(self):
        return self.overlap


@dataclass
class Candidate:
    """suggest.Candidate, pkg/suggest/collector.go:12-17"""
    Key: int
    Score: float

    def Less(self, o):  # collector.go:20-26
        if self.Score == o.Score:
            return self.Key > o.Key
        return self.Score < o.Score


class Scorer:
    """suggest.Scorer, pkg/suggest/scorer.go:9-12"""

    def Score(self, candidate):
        raise NotImplementedError


class MetricScorer(Scorer):
    """metricScorer, pkg/suggest/scorer.go:14-31"""

    def __init__(self, metric, sizeA, sizeB):
        self.metric, self.sizeA, self.sizeB = metric, sizeA, sizeB

    def Score(self, candidate):
        return 1 - self.metric.Distance(candidate.Overlap(), self.sizeA, self.sizeB)


def NewMetricScorer(metric, sizeA, sizeB):
    return MetricScorer(metric, sizeA, sizeB)


class Collector:
    """suggest.Collector, pkg/suggest/collector.go:28-33"""

    def Collect(self, candidate):
        raise NotImplementedError

    def SetScorer(self, scorer):
        pass


class CollectorManager:
    """suggest.CollectorManager, pkg/suggest/collector.go:35-43"""

    def Create(self):
        raise NotImplementedError

    def Collect(self, *collectors):
        raise NotImplementedError

    def GetCandidates(self):
        raise NotImplementedError


class TopKQueue:
    """topKQueue (pkg/suggest/topk.go:66-175): the topK greatest under Candidate.Less, returned best first."""

    def __init__(self, topK):
        self.topK = topK
        self._h = []  # min-heap under Less: (score, -key)

    def Add(self, position, score):
        if self.topK <= 0:
            return
        item = (score, -position)
        if len(self._h) < self.topK:
            heapq.heappush(self._h, item)
        elif self._h[0] < item:
            heapq.heapreplace(self._h, item)

    def Merge(self, other):
        for score, neg in other._h:
            self.Add(-neg, score)

    def IsFull(self):
        return len(self._h) == self.topK

    def GetLowestScore(self):
        return self._h[0][0] if self._h else float("-inf")

    def GetCandidates(self):
        return [Candidate(-neg, score) for score, neg in sorted(self._h, reverse=True)]


class _FuzzyCollector(Collector):
    def __init__(self, queue):
        self.topKQueue, self.scorer = queue, None

    def Collect(self, item):
        self.topKQueue.Add(item.Position(), self.scorer.Score(item))

    def SetScorer(self, scorer):
        self.scorer = scorer


class FuzzyCollectorManager(CollectorManager):
    """pkg/suggest/collector.go:139-191.  NGramIndex.Suggest recognises this type and runs sg_search_batch instead of
    replaying (the reference type-switches on it as well, suggester.go:101)."""

    def __init__(self, topK):
        self.topK = topK
        self.globalQueue = TopKQueue(topK)

    def Create(self):
        return _FuzzyCollector(TopKQueue(self.topK))

    def Collect(self, *collectors):
        for c in collectors:
            if not isinstance(c, _FuzzyCollector):
                raise TypeError("expected Collector created by FirstKCollectorManager")  # sic, collector.go:170
            self.globalQueue.Merge(c.topKQueue)

    def GetCandidates(self):
        return self.globalQueue.GetCandidates()

    def GetLowestScore(self):
        return self.globalQueue.GetLowestScore() if self.globalQueue.IsFull() else float("-inf")


def NewFuzzyCollectorManager(topK):
    """newFuzzyCollectorManager(topK), pkg/suggest/collector.go:143-149: a CollectorManagerFactory"""
    return lambda: FuzzyCollectorManager(topK)


class _FirstKCollector(Collector):
    def __init__(self, limit):
        self.limit, self.items = limit, []

    def Collect(self, item):
        if self.limit == len(self.items):
            raise ErrCollectionTerminated()
        self.items.append(item)


class FirstKCollectorManager(CollectorManager):
    """pkg/suggest/collector.go:48-115"""

    def __init__(self, limit):
        self.limit = limit
        self.queue = TopKQueue(limit)

    def Create(self):
        return _FirstKCollector(self.limit)

    def Collect(self, *collectors):
        for c in collectors:
            if not isinstance(c, _FirstKCollector):
                raise TypeError("expected Collector created by FirstKCollectorManager")
            for cand in c.items:
                self.queue.Add(cand.Position(), -float(cand.Position()))

    def GetCandidates(self):
        return self.queue.GetCandidates()


def NewFirstKCollectorManager(limit):
    return lambda: FirstKCollectorManager(limit)
