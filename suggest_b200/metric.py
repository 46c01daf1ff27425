"""pkg/metric mirror.  The five built-in metrics are evaluated on the device (float64, the reference's
operation order, pkg/metric/{jaccard,cosine,dice,overlap,exact}.go); these objects only name them."""
from . import _capi


class Metric:
    """metric.Metric (pkg/metric/metric.go:7-16).  `code` is the sg_metric enum passed through the C ABI."""
    code = None
    name = None

    def __repr__(self):
        return f"{self.name}Metric()"


class _Jaccard(Metric):
    code, name = _capi.SG_JACCARD, "Jaccard"


class _Cosine(Metric):
    code, name = _capi.SG_COSINE, "Cosine"


class _Dice(Metric):
    code, name = _capi.SG_DICE, "Dice"


class _Overlap(Metric):
    code, name = _capi.SG_OVERLAP, "Overlap"


class _Exact(Metric):
    code, name = _capi.SG_EXACT, "Exact"


def JaccardMetric():
    return _Jaccard()


def CosineMetric():
    return _Cosine()


def DiceMetric():
    return _Dice()


def OverlapMetric():
    return _Overlap()


def ExactMetric():
    return _Exact()


BY_NAME = {"Jaccard": JaccardMetric, "Cosine": CosineMetric, "Dice": DiceMetric, "Exact": ExactMetric,
           "Overlap": OverlapMetric}  # internal/suggest/api/suggest_handler.go:26-34
