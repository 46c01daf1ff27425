"""suggest_b200 — B200 (sm_100a) implementation of suggest-go/suggest's Service.Suggest / NGramIndex.Suggest path.

The package is a thin host-side mirror of the reference's Go interfaces over the C ABI of
libsuggest_b200.so (include/suggest_b200.h).  Names follow the reference: pkg/suggest
(IndexDescription, Builder, NGramIndex, Service, SearchConfig, Candidate, ResultItem) and pkg/metric.
"""
from . import collector, metric
from .metric import CosineMetric, DiceMetric, ExactMetric, JaccardMetric, OverlapMetric
from .suggest import (CANDIDATE_DTYPE, Batcher, Candidate, PinnedCandidateRows, Ticket, IndexDescription, NewBatcher, PinnedBuffers, NGramIndex, NewRAMBuilder, NewFSBuilder, NewSearchConfig, NewService,
                      ResultItem, SearchConfig, Service, SuggestError, pack_strings)

__all__ = ["collector", "metric", "CosineMetric", "DiceMetric", "ExactMetric", "JaccardMetric", "OverlapMetric", "CANDIDATE_DTYPE", "PinnedCandidateRows", "Ticket", "Batcher", "NewBatcher", "Candidate", "PinnedBuffers",
           "IndexDescription", "NGramIndex", "NewRAMBuilder", "NewFSBuilder", "NewSearchConfig", "NewService",
           "ResultItem", "SearchConfig", "Service", "SuggestError", "pack_strings"]
