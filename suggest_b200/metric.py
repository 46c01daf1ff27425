"""pkg/metric mirror.  The five built-in metrics are evaluated on the device (float64, the reference's
operation order, pkg/metric/{jaccard,cosine,dice,overlap,exact}.go); these objects only name them."""
import math

from . import _capi


class Metric:
    """metric.Metric (pkg/metric/metric.go:7-16).  `code` is the sg_metric enum passed through the C ABI.

    A metric of the caller's own subclasses this with `code = None` and implements the interface's four methods; the
    index then tabulates Threshold over the window on the host and the device returns every candidate with its overlap
    (sg_candidates_batch), which is scored with Distance here (pkg/suggest/scorer.go:29-31)."""
    code = None
    name = None

    def MinY(self, alpha, size):
        raise NotImplementedError

    def MaxY(self, alpha, size):
        raise NotImplementedError

    def Threshold(self, alpha, sizeA, sizeB):
        raise NotImplementedError

    def Distance(self, inter, sizeA, sizeB):
        raise NotImplementedError

    def __repr__(self):
        return f"{self.name}Metric()"


# Host-side arithmetic of the built-ins (same operation order as pkg/metric/*.go, IEEE doubles): used only when a
# caller brings its own CollectorManager and the candidates are replayed on the host (NGramIndex.Suggest with a
# factory) - the Go shim calls the reference's own metric objects there.  Searches with a top-k never come here.
class _Jaccard(Metric):  # pkg/metric/jaccard.go:12-27
    code, name = _capi.SG_JACCARD, "Jaccard"

    def MinY(self, alpha, size):
        return int(math.ceil(alpha * float(size)))

    def MaxY(self, alpha, size):
        return int(math.floor(float(size) / alpha))

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(alpha * float(sizeA + sizeB) / (1 + alpha)))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / float(sizeA + sizeB - inter)


class _Cosine(Metric):  # pkg/metric/cosine.go:12-26
    code, name = _capi.SG_COSINE, "Cosine"

    def MinY(self, alpha, size):
        return int(math.ceil(alpha * alpha * float(size)))

    def MaxY(self, alpha, size):
        return int(math.floor(float(size) / (alpha * alpha)))

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(alpha * math.sqrt(float(sizeA * sizeB))))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / math.sqrt(float(sizeA * sizeB))


class _Dice(Metric):  # pkg/metric/dice.go:12-26
    code, name = _capi.SG_DICE, "Dice"

    def MinY(self, alpha, size):
        return int(math.ceil(alpha / (2 - alpha) * float(size)))

    def MaxY(self, alpha, size):
        return int(math.floor((2 - alpha) / alpha * float(size)))

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(0.5 * alpha * float(sizeA + sizeB)))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(2 * inter) / float(sizeA + sizeB)


class _Overlap(Metric):  # pkg/metric/overlap.go:12-26
    code, name = _capi.SG_OVERLAP, "Overlap"

    def MinY(self, alpha, size):
        return 1

    def MaxY(self, alpha, size):
        return 32767  # math.MaxInt16

    def Threshold(self, alpha, sizeA, sizeB):
        return int(math.ceil(alpha * min(float(sizeA), float(sizeB))))

    def Distance(self, inter, sizeA, sizeB):
        return 1 - float(inter) / min(float(sizeA), float(sizeB))


class _Exact(Metric):  # pkg/metric/exact.go:10-24
    code, name = _capi.SG_EXACT, "Exact"

    def MinY(self, alpha, size):
        return size

    def MaxY(self, alpha, size):
        return size

    def Threshold(self, alpha, sizeA, sizeB):
        return sizeA

    def Distance(self, inter, sizeA, sizeB):
        return 0.0


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
