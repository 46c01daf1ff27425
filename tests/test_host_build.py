"""Host side of libsuggest_b200 (tokenizer chain, CSR build) against the oracle.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION, COLLECTION, TEST_DESCRIPTION
from oracle import oracle
from suggest_b200 import _capi
from suggest_b200.suggest import pack_strings


def host_tokenize(desc, text):
    t = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    cfg, keep = _capi.make_config(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"])
    cap = 4 * (len(t) + 64) * 8 + 64
    out = C.create_string_buffer(cap)
    off = (C.c_uint32 * (len(t) + 66))()
    n = _capi.check(_capi.lib().sg_host_tokenize(C.byref(cfg), t, len(t), out, cap, off, len(t) + 65), host=True)
    return [out.raw[off[i]:off[i + 1]] for i in range(n)]


TEXTS = ["Nissan March", "niss ma", "", "a", "ab", "  padded  ", "RAM RAM", "жигули", "ЖИГУЛИ Ёлка ёж", "lalala",
         "MiXeD Case 42", "tab\tand-dash", "日本語テキスト", "İstanbul", "ǅ Ǆ ǆ", b"bad\xff\xfebytes", b"\xc3", b"\xe2\x82",
         "x" * 130, "ab$cd", "$", "$$", " ", "ÀÉÎÕÜ", "ß ẞ", "Ω ω", "á", "🙂 emoji", "q"]


@pytest.mark.parametrize("desc", [TEST_DESCRIPTION, CARS_DESCRIPTION,
                                  dict(ngram_size=2, wrap=("", ""), pad="_", alphabet=("english",)),
                                  dict(ngram_size=4, wrap=("^", "$"), pad="$", alphabet=("english", "numbers", "$^")),
                                  dict(ngram_size=3, wrap=("  ", " "), pad="#", alphabet=("russian", "abcё")),
                                  dict(ngram_size=1, wrap=("", ""), pad="?", alphabet=("numbers",)),
                                  dict(ngram_size=8, wrap=("$", "$"), pad="$", alphabet=("english", "russian", "numbers", "$"))])
def test_tokenizer_chain_matches_oracle(desc):
    ox = oracle.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"])
    for text in TEXTS:
        assert host_tokenize(desc, text) == ox.tokenize(text), (desc, text)


def test_tokenizer_on_every_cars_line(cars_lines):
    ox = oracle.OracleIndex(**{"ngram_size": 3, "wrap": ("$", "$"), "pad": "$", "alphabet": CARS_DESCRIPTION["alphabet"]})
    for line in cars_lines[::7]:
        assert host_tokenize(CARS_DESCRIPTION, line) == ox.tokenize(line)


def test_to_lower_matches_oracle():
    for text in TEXTS:
        t = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        out = C.create_string_buffer(3 * len(t) + 8)
        n = _capi.lib().sg_host_to_lower(t, len(t), out, 3 * len(t) + 8)
        assert out.raw[:n] == oracle.to_lower(t)


def test_unsupported_descriptions_are_rejected():
    for desc in (dict(ngram_size=3, wrap=("$", "$"), pad="", alphabet=("english",)),
                 dict(ngram_size=3, wrap=("$", "$"), pad="ab", alphabet=("english",)),
                 dict(ngram_size=9, wrap=("$", "$"), pad="$", alphabet=("english",)),
                 dict(ngram_size=0, wrap=("$", "$"), pad="$", alphabet=("english",))):
        with pytest.raises(_capi.SuggestError):
            host_tokenize(desc, "abc")


def host_index(desc, docs):
    data, off = pack_strings(docs, np.uint64)
    cfg, keep = _capi.make_config(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"])
    h = C.c_void_p()
    _capi.check(_capi.lib().sg_host_index_build(C.byref(cfg), data.ctypes.data_as(C.c_void_p),
                                                off.ctypes.data_as(C.c_void_p), len(docs), C.byref(h)), host=True)
    return h


def host_list(h, seg, term):
    n = _capi.lib().sg_host_index_get_list(h, seg, term, len(term), None, 0)
    if n < 0:
        return None
    out = np.zeros(n, dtype=np.uint32)
    assert _capi.lib().sg_host_index_get_list(h, seg, term, len(term), out.ctypes.data_as(C.c_void_p), n) == n
    return out


def check_index_equals_oracle(desc, docs):
    ox = oracle.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    h = host_index(desc, docs)
    try:
        info = _capi.SgIndexInfo()
        _capi.lib().sg_host_index_get_info(h, C.byref(info))
        assert info.n_docs == len(docs)
        assert info.n_segments == ox.segments
        assert info.n_lists == ox.lists
        total = 0
        for seg, term, ids in ox.iter_lists():
            want = np.unique(ids)  # an id repeated inside one list is stored once (SURVEY.md 8c rule 5)
            got = host_list(h, seg, term)
            assert got is not None, (seg, term)
            assert np.array_equal(got, want), (seg, term)
            total += len(want)
        assert info.n_postings == total
    finally:
        _capi.lib().sg_host_index_free(h)


def test_index_of_test_collection_matches_oracle():
    check_index_equals_oracle(TEST_DESCRIPTION, COLLECTION)


def test_index_of_cars_matches_oracle(cars_lines):
    check_index_equals_oracle(CARS_DESCRIPTION, cars_lines)


def test_index_with_empty_and_short_documents():
    check_index_equals_oracle(TEST_DESCRIPTION, ["", "a", " ", "ab", "abc", "abc", "ёжик", "RAM RAM", ""])
    check_index_equals_oracle(dict(ngram_size=3, wrap=("", ""), pad="$", alphabet=("english",)), ["", "a", "ab", "abc", "é", "éa"])


def test_index_synthetic_matches_oracle():
    rng = np.random.default_rng(7)
    lens = rng.integers(8, 33, size=3000)
    docs = ["".join(chr(97 + c) for c in rng.integers(0, 26, size=l)) for l in lens]
    check_index_equals_oracle(TEST_DESCRIPTION, docs)


def check_bitmaps(desc, docs, monkeypatch, shift):
    """Segments start on bucket boundaries; bit b of a term's row is set iff one of its lists holds a slot of bucket b."""
    if shift is None:
        monkeypatch.delenv("SG_BUCKET_SHIFT", raising=False)
    else:
        monkeypatch.setenv("SG_BUCKET_SHIFT", str(shift))
    ox = oracle.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    h = host_index(desc, docs)
    L = _capi.lib()
    try:
        lay = _capi.SgIndexLayout()
        assert L.sg_host_index_get_layout(h, C.byref(lay)) == 0
        s = lay.bucket_shift
        if shift is not None:
            assert s == min(shift, 8)
        assert lay.row_words % 64 == 0 and lay.row_words * 32 >= -(-lay.n_slots // (1 << s))
        seg = np.zeros(ox.segments + 1, dtype=np.uint32)
        assert L.sg_host_index_get_segments(h, seg.ctypes.data_as(C.c_void_p), len(seg)) == len(seg)
        assert seg[-1] == lay.n_slots and np.all(np.diff(seg.astype(np.int64)) >= 0)
        assert np.all(seg[:-1] % (1 << s) == 0)
        want_bits = {}
        n_real = 0
        for b, term, ids in ox.iter_lists():
            n = L.sg_host_index_get_list_slots(h, b, term, len(term), None, 0)
            slots = np.zeros(n, dtype=np.uint32)
            assert L.sg_host_index_get_list_slots(h, b, term, len(term), slots.ctypes.data_as(C.c_void_p), n) == n
            assert np.all(np.diff(slots.astype(np.int64)) > 0)
            assert np.all((slots >= seg[b]) & (slots < seg[b + 1]))
            want_bits.setdefault(term, set()).update((slots >> s).tolist())
            n_real += n
        for term, buckets in want_bits.items():
            row = np.zeros(lay.row_words, dtype=np.uint32)
            assert L.sg_host_index_get_bitmap(h, term, len(term), row.ctypes.data_as(C.c_void_p), len(row)) == len(row)
            got = set(np.flatnonzero(np.unpackbits(row.view(np.uint8), bitorder="little")).tolist())
            assert got == buckets, term
    finally:
        L.sg_host_index_free(h)


@pytest.mark.parametrize("shift", [None, 0, 3, 7, 12])
def test_bitmap_rows_match_lists_on_cars(cars_lines, monkeypatch, shift):
    check_bitmaps(CARS_DESCRIPTION, cars_lines[::3], monkeypatch, shift)


def test_bitmap_rows_synthetic_auto_shift(monkeypatch):
    rng = np.random.default_rng(11)
    lens = rng.integers(8, 33, size=20000)
    docs = ["".join(chr(97 + c) for c in rng.integers(0, 26, size=l)) for l in lens]
    check_bitmaps(TEST_DESCRIPTION, docs, monkeypatch, None)


def test_index_without_bitmaps_when_over_budget(monkeypatch):
    monkeypatch.setenv("SG_BITMAP_MAX_MB", "0")
    h = host_index(TEST_DESCRIPTION, COLLECTION)
    try:
        lay = _capi.SgIndexLayout()
        assert _capi.lib().sg_host_index_get_layout(h, C.byref(lay)) == 0
        assert lay.row_words == 0 and lay.engine == 0 and lay.bitmap_bytes == 0
    finally:
        _capi.lib().sg_host_index_free(h)
