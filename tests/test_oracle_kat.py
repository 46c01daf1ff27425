"""The CPU oracle against every known-answer test the reference holds for the Suggest path."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import TEST_DESCRIPTION


# pkg/analysis/ngram_tokenizer_test.go:16-45
@pytest.mark.parametrize("word,k,ngrams", [
    ("tet", 2, ["te", "et"]),
    ("te", 2, ["te"]),
    ("testing", 3, ["tes", "est", "sti", "tin", "ing"]),
    ("жигули", 2, ["жи", "иг", "гу", "ул", "ли"]),
    ("", 2, []),
    ("lalala", 2, ["la", "al"]),
])
def test_tokenize_ngrams(word, k, ngrams):
    assert O.ngram_tokenize(word, k) == [g.encode() for g in ngrams]


def test_ngram_fewer_runes_than_n_but_enough_bytes():
    # ngram_tokenizer.go:18 compares BYTES with n: two Cyrillic runes are 4 bytes >= 3 -> one token, the whole text
    assert O.ngram_tokenize("жи", 3) == ["жи".encode()]
    assert O.ngram_tokenize("ab", 3) == []


# pkg/alphabet/alphabet_test.go:10-61
def test_alphabets():
    rus = O.OracleIndex(alphabet=("russian",))
    for ch, exp in [("а", True), ("е", True), ("ё", True), ("я", True), ("j", False), ("7", False)]:
        assert rus.has(ch) is exp
    comp = O.OracleIndex(alphabet=("russian", "english", "numbers"))
    for ch, exp in [("a", True), ("b", True), ("z", True), ("а", True), ("ё", True), ("е", True), ("ж", True),
                    ("я", True), ("7", True), ("-", False)]:
        assert comp.has(ch) is exp


def test_full_chain_dedupe_before_normalise():
    # SURVEY §8c rule 3: "RAM RAM" -> [$ra ram am$ m$r $ra am$], cardinality 6
    ix = O.OracleIndex(**TEST_DESCRIPTION)
    assert ix.tokenize("RAM RAM") == [b"$ra", b"ram", b"am$", b"m$r", b"$ra", b"am$"]
    assert ix.tokenize("Nissan ma") == [b"$ni", b"nis", b"iss", b"ssa", b"san", b"an$", b"n$m", b"$ma", b"ma$"]
    assert ix.tokenize("") == []
    assert ix.tokenize("Ёж") == ["$ёж".encode(), "ёж$".encode()]


def test_to_lower_go_semantics():
    assert O.to_lower("ABC xyZ") == b"abc xyz"
    assert O.to_lower("ЖИГУЛИ Ё") == "жигули ё".encode()
    assert O.to_lower("İ") == b"i"                       # simple mapping of U+0130
    assert O.to_lower(b"A\xffB\xc3") == b"a\xef\xbf\xbdb\xef\xbf\xbd"  # strings.Map: invalid byte -> U+FFFD
    assert O.to_lower(b"abc\x80") == b"abc\xef\xbf\xbd"


RID_A = [[1, 2, 3], [1, 2], [2, 3], [2]]
RID_B = [[1, 2, 3, 5, 7, 10, 30, 50], [10, 11, 13, 16, 50, 60, 131], [40, 50, 60], [50, 100], [100, 200]]
# pkg/merger/list_merger_test.go:48-140
MERGE_CASES = [
    (RID_A, 2, {2: [1, 3], 4: [2]}),
    (RID_A, 3, {4: [2]}),
    (RID_A, 4, {4: [2]}),
    (RID_B, 4, {4: [50]}),
    (RID_B, 3, {4: [50]}),
    (RID_B, 2, {2: [10, 60, 100], 4: [50]}),
    (RID_B, 1, {1: [1, 2, 3, 5, 7, 11, 13, 16, 30, 40, 131, 200], 2: [10, 60, 100], 4: [50]}),
]


@pytest.mark.parametrize("algo", [O.SCAN_COUNT, O.CP_MERGE, O.MERGE_SKIP, O.DIVIDE_SKIP])
@pytest.mark.parametrize("case", range(len(MERGE_CASES)))
def test_merge(algo, case):
    rid, t, expected = MERGE_CASES[case]
    actual = {}
    for pos, overlap in O.merge(algo, rid, t):
        actual.setdefault(overlap, []).append(pos)
    assert actual == expected


# pkg/merger/list_intersector_test.go:14-46
def test_intersect():
    assert O.intersect(RID_A) == [2]
    assert O.intersect(RID_B) == []
    assert O.intersect(RID_B[:4] + [[50, 100, 200]]) == [50]


# pkg/compression/compression_test.go:28-56
@pytest.mark.parametrize("codec,gap", [(O.CODEC_BINARY, 0), (O.CODEC_VB, 0), (O.CODEC_SKIPPING, 3)])
@pytest.mark.parametrize("lst", [[824, 829, 215406], [1, 9, 13, 180, 999, 12345],
                                 [1, 13, 29, 101, 506, 10003, 10004, 12000, 12901]])
def test_encode_decode(codec, gap, lst):
    data = O.encode(codec, lst, gap)
    assert O.decode(codec, data, len(lst), gap).tolist() == lst


def test_varint_wire_format():
    # pkg/store/byte_output.go:26-38: 7-bit groups, least significant first, 0x80 = continuation
    assert O.encode(O.CODEC_VB, [1, 129, 129 + 16384]) == bytes([0x01, 0x80, 0x01, 0x80, 0x80, 0x01])
    # skipping.go comment block, gap 3: header = uint16 LE (bytes + 2), bit 15 on the last block
    enc = O.encode(O.CODEC_SKIPPING, [1, 13, 29, 101, 506, 10003, 10004, 12000, 12901], 3)
    assert enc[:5] == bytes([5, 0, 1, 12, 16])
    assert O.decode(O.CODEC_SKIPPING, enc, 9, 3).tolist() == [1, 13, 29, 101, 506, 10003, 10004, 12000, 12901]


PL = [1, 13, 29, 101, 506, 10003, 10004, 12000, 12001]
# pkg/index/posting_list_test.go:39-132
@pytest.mark.parametrize("kind,gap", [(O.CODEC_SKIPPING, 3), (O.CODEC_VB, 0)])
@pytest.mark.parametrize("to,lb,tail,err", [
    (1, 1, PL, False), (2, 13, PL[1:], False), (12000, 12000, [12000, 12001], False), (12001, 12001, [12001], False),
    (0, 1, PL, False), (12002, 0, [], True),
])
def test_posting_lower_bound(kind, gap, to, lb, tail, err):
    got_lb, got_err, got_tail = O.posting_lower_bound_tail(kind, PL, to, gap)
    assert (got_lb, got_err, got_tail) == (lb, err, tail)


def test_skipping_iterator_random_walks():
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(66, 257))
        lst = np.sort(rng.choice(5_000_000, size=n, replace=False)).astype(np.uint32)
        to = int(rng.integers(0, 5_000_100))
        lb, err, tail = O.posting_lower_bound_tail(O.CODEC_SKIPPING, lst, to, 64)
        j = int(np.searchsorted(lst, to))
        if j == n:
            assert err and tail == []
        else:
            assert not err and lb == lst[j] and tail == lst[j:].tolist()


# pkg/suggest/topk_test.go:10-39
def test_topk_queue():
    cands = [(1, 0.1), (2, 0.01), (3, 0.91), (4, 0.24), (5, 0.13), (6, 0.07), (7, 0.9), (8, 0.12345), (9, 0.65),
             (10, 0.6565)]
    top, lowest = O.topk(cands, 3)
    assert top == [(3, 0.91), (7, 0.9), (10, 0.6565)]
    assert lowest == 0.6565


def test_topk_tie_break_lowest_id_first():
    top, _ = O.topk([(9, 0.5), (3, 0.5), (7, 0.5), (1, 0.4)], 2)
    assert top == [(3, 0.5), (7, 0.5)]  # collector.go:20-26


def test_metric_formulas():
    # pkg/metric/*.go restated in Python floats (IEEE double, no FMA)
    import math
    for a in (1, 2, 7, 20, 33):
        for alpha in (0.3, 0.5, 0.7, 1.0):
            assert O.metric_min_y(O.JACCARD, alpha, a) == math.ceil(alpha * a)
            assert O.metric_max_y(O.JACCARD, alpha, a) == math.floor(a / alpha)
            assert O.metric_min_y(O.COSINE, alpha, a) == math.ceil(alpha * alpha * a)
            assert O.metric_max_y(O.COSINE, alpha, a) == math.floor(a / (alpha * alpha))
            assert O.metric_min_y(O.DICE, alpha, a) == math.ceil(alpha / (2 - alpha) * a)
            assert O.metric_max_y(O.DICE, alpha, a) == math.floor((2 - alpha) / alpha * a)
            for b in (1, 5, 19, 20, 21, 40):
                assert O.metric_threshold(O.JACCARD, alpha, a, b) == math.ceil(alpha * (a + b) / (1 + alpha))
                assert O.metric_threshold(O.COSINE, alpha, a, b) == math.ceil(alpha * math.sqrt(a * b))
                assert O.metric_threshold(O.DICE, alpha, a, b) == math.ceil(0.5 * alpha * (a + b))
                assert O.metric_threshold(O.OVERLAP, alpha, a, b) == math.ceil(alpha * min(a, b))
                assert O.metric_threshold(O.EXACT, alpha, a, b) == a
                for c in range(0, min(a, b) + 1):
                    assert O.score(O.JACCARD, c, a, b) == 1 - (1 - c / (a + b - c))
                    assert O.score(O.COSINE, c, a, b) == 1 - (1 - c / math.sqrt(a * b))
                    assert O.score(O.DICE, c, a, b) == 1 - (1 - (2 * c) / (a + b))
                    assert O.score(O.OVERLAP, c, a, b) == 1 - (1 - c / min(a, b))
                    assert O.score(O.EXACT, c, a, b) == 1.0
