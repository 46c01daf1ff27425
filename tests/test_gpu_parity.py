"""Parity of the CUDA Suggest path (through the C ABI) with the CPU oracle.  Needs a B200: `pytest -m gpu`.

ids must be identical, order included; scores are compared with == (both sides evaluate the same
float64 expression) and, per the contract, within 1e-6.
"""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION, COLLECTION, TEST_DESCRIPTION
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200 import _capi
from suggest_b200.suggest import IndexDescription, pack_strings
from suggest_b200.workload import synthetic_workload, unpack

pytestmark = pytest.mark.gpu

METRICS = {O.JACCARD: S.JaccardMetric(), O.COSINE: S.CosineMetric(), O.DICE: S.DiceMetric(), O.OVERLAP: S.OverlapMetric(),
           O.EXACT: S.ExactMetric()}
SCORE_TOL = 1e-6


def description(d, name="t"):
    return IndexDescription(Name=name, NGramSize=d["ngram_size"], Alphabet=tuple(d["alphabet"]), Pad=d["pad"],
                            Wrap=tuple(d["wrap"]))


def build_gpu(desc, docs, env=None):
    """GPU index; env = tuning knobs read at index creation"""
    old = {}
    for k_, v in (env or {}).items():
        old[k_] = os.environ.get(k_)
        os.environ[k_] = str(v)
    try:
        gx = S.NewRAMBuilder(docs, description(desc)).Build()
    finally:
        for k_, v in old.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
    return gx


def build_pair(desc, docs, env=None):
    """(GPU index, oracle index) over the same documents"""
    ox = O.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    return build_gpu(desc, docs, env), ox


def assert_same(gx, ox, queries, metric, alpha, k, what=""):
    ids_g, sc_g, n_g = gx.SuggestBatch(queries, alpha, METRICS[metric], k)
    ids_o, sc_o, n_o = ox.suggest_batch(queries, metric, alpha, k, O.CANONICAL, threads=8)
    bad = np.nonzero(n_g != n_o)[0]
    assert len(bad) == 0, (what, "count differs", bad[:5], [queries[i] for i in bad[:5]], n_g[bad[:5]], n_o[bad[:5]])
    mask = np.arange(k)[None, :] < n_o[:, None]
    diff = (ids_g != ids_o) & mask
    if diff.any():
        q = int(np.nonzero(diff.any(axis=1))[0][0])
        raise AssertionError((what, "ids differ", q, queries[q], ids_g[q, :n_g[q]], ids_o[q, :n_o[q]], sc_g[q, :n_g[q]], sc_o[q, :n_o[q]]))
    assert np.all(np.abs(sc_g - sc_o)[mask] <= SCORE_TOL), what
    assert np.array_equal(sc_g[mask], sc_o[mask]), (what, "scores not bit-equal")
    return n_o


# ---------------------------------------------------------------------------------------------------
# the reference's own end-to-end expectations
# ---------------------------------------------------------------------------------------------------
def test_ngram_index_test_go():
    # pkg/suggest/ngram_index_test.go:15-40
    gx = S.NewRAMBuilder(COLLECTION, description(TEST_DESCRIPTION)).Build()
    got = gx.Suggest("Nissan ma", 0.5, S.JaccardMetric(), 2)
    assert [c.Key for c in got] == [2, 0]
    gx.close()


def test_example_test_go():
    # pkg/suggest/example_test.go:14-72
    d = IndexDescription(Name="cars", NGramSize=3, Wrap=("$", "$"), Pad="$", Alphabet=("english", "$"))
    gx = S.NewRAMBuilder(COLLECTION, d).Build()
    got = gx.Suggest("niss ma", 0.4, S.CosineMetric(), 5)
    assert [COLLECTION[c.Key] for c in got] == ["Nissan Maxima", "Nissan March"]


def test_service_test_go(cars_lines, tmp_path):
    # pkg/suggest/service_test.go:11-80 (RAM driver): Cosine 0.7, k=5, concurrent re-adds of the index
    src = tmp_path / "cars.dict"
    src.write_bytes(b"\n".join(cars_lines) + b"\n")
    d = description(CARS_DESCRIPTION, "cars")
    d.SourcePath = str(src)
    service = S.NewService()
    service.AddRunTimeIndex(d)
    words = ["Nissan March", "Honda Fitt", "Wolfsvagen", "Tayota Corolla", "Micra Nissan"]
    expected = [["NISSAN MARCH"], ["HONDA FIT"], [], ["TOYOTA COROLLA"], ["NISSAN MICRA"]]
    errors = []

    def search():
        try:
            for _ in range(3):
                for w, exp in zip(words, expected):
                    res = service.Suggest("cars", S.NewSearchConfig(w, 5, S.CosineMetric(), 0.7))
                    if [r.Value for r in res] != exp:
                        errors.append((w, res))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def reindex():
        try:
            service.AddRunTimeIndex(d)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=search) for _ in range(5)] + [threading.Thread(target=reindex) for _ in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    with pytest.raises(KeyError):
        service.Suggest("nope", S.NewSearchConfig("x", 5, S.CosineMetric(), 0.7))
    with pytest.raises(ValueError):
        S.NewSearchConfig("x", 0, S.CosineMetric(), 0.7)
    with pytest.raises(ValueError):
        S.NewSearchConfig("x", 5, S.CosineMetric(), 1.5)


# ---------------------------------------------------------------------------------------------------
# cars: every entry as a query, all metrics
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cars_pair(cars_lines):
    return build_pair(CARS_DESCRIPTION, cars_lines)


@pytest.mark.parametrize("metric,alpha,k", [(O.COSINE, 0.7, 5), (O.JACCARD, 0.5, 10), (O.DICE, 0.5, 10), (O.COSINE, 0.5, 5),
                                            (O.OVERLAP, 0.9, 7), (O.EXACT, 1.0, 3), (O.JACCARD, 1.0, 4), (O.JACCARD, 0.2, 40),
                                            (O.DICE, 0.35, 100)])
def test_cars_every_entry(cars_pair, cars_lines, metric, alpha, k):
    gx, ox = cars_pair
    n = assert_same(gx, ox, cars_lines, metric, alpha, k, f"cars m={metric} a={alpha} k={k}")
    assert n.sum() > 0


def test_cars_misspelled_queries(cars_pair):
    # pkg/suggest/ngram_index_test.go:196-206 (benchmark query set, no expected output in the reference)
    gx, ox = cars_pair
    queries = ["Nissan Mar", "Hnda Fi", "Mersdes Benz", "Tayota Corolla", "Nssan Skylike", "Nissan Juke", "Dodje iper",
               "Hummer", "tayota"]
    for metric, alpha in ((O.COSINE, 0.5), (O.JACCARD, 0.3), (O.DICE, 0.4)):
        assert_same(gx, ox, queries, metric, alpha, 5)


SCAN = dict(SG_ENGINE="scancount")


@pytest.mark.parametrize("env", [dict(SCAN, SG_FORCE_SHIFT=0, SG_TBL_BYTES=2048), dict(SCAN, SG_FORCE_SHIFT=2, SG_TBL_BYTES=2048),
                                 dict(SCAN, SG_FORCE_SHIFT=5), dict(SCAN, SG_FORCE_SHIFT=13), dict(SCAN, SG_TBL_BYTES=4096, SG_WARPS=3),
                                 dict(SCAN, SG_TBL_BYTES=65536), dict(SCAN), dict(SG_BITMAP_MAX_MB=0),
                                 dict(SG_BUCKET_SHIFT=0), dict(SG_BUCKET_SHIFT=1), dict(SG_BUCKET_SHIFT=3), dict(SG_BUCKET_SHIFT=5),
                                 dict(SG_BUCKET_SHIFT=8), dict(SG_ENGINE="bitmap", SG_BUCKET_SHIFT=2)])
def test_cars_bucket_widths_and_table_sizes(cars_lines, cars_pair, env):
    """Both engines, every bucket width / chunking give the same answer (scan-count: exact counters, bucket filter + merge,
    multi-chunk; bitmap: one bit per document up to 256 documents per bit)."""
    gx = build_gpu(CARS_DESCRIPTION, cars_lines, env)
    assert gx.layout()["engine"] == (0 if env.get("SG_ENGINE") == "scancount" or "SG_BITMAP_MAX_MB" in env else 1)
    _, ox = cars_pair
    q = cars_lines[::3]
    assert_same(gx, ox, q, O.JACCARD, 0.5, 10, str(env))
    assert_same(gx, ox, q, O.COSINE, 0.6, 3, str(env))
    gx.close()


# ---------------------------------------------------------------------------------------------------
# synthetic a-z dictionaries (BASELINE.json configs #2/#3 at a size the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------
def synthetic(n_docs, n_queries, seed=12345):
    (d, off), (q, q_off), _ = synthetic_workload(n_docs, n_queries, seed)
    return unpack(d, off), unpack(q, q_off)


@pytest.fixture(scope="module")
def synth_pairs():
    docs, queries = synthetic(60000, 3000)
    out = {}
    for n in (2, 3, 4):
        desc = dict(TEST_DESCRIPTION, ngram_size=n)
        out[n] = build_pair(desc, docs) + (queries,)
    return out


@pytest.mark.parametrize("n", [2, 3, 4])
@pytest.mark.parametrize("metric,alpha", [(O.JACCARD, 0.5), (O.COSINE, 0.5), (O.DICE, 0.5), (O.JACCARD, 0.25), (O.OVERLAP, 0.7)])
def test_synthetic_sweep(synth_pairs, n, metric, alpha):
    gx, ox, queries = synth_pairs[n]
    cnt = assert_same(gx, ox, queries, metric, alpha, 10, f"synthetic n={n} m={metric} a={alpha}")
    assert (cnt > 0).mean() > 0.3


@pytest.mark.parametrize("env", [SCAN, dict(SG_BUCKET_SHIFT=5), dict(SG_BUCKET_SHIFT=8), dict(SG_BUCKET_SHIFT=0)])
def test_synthetic_other_engines_and_widths(synth_pairs, env):
    """The default index of the sweep above uses the bitmap engine at the width the build picks; here the same data
    through the scan-count engine and through other bucket widths."""
    _, ox, queries = synth_pairs[3]
    docs, _ = synthetic(60000, 3000)
    gx = build_gpu(TEST_DESCRIPTION, docs, env)
    assert gx.layout()["engine"] == (0 if env is SCAN else 1)
    for metric, alpha in ((O.JACCARD, 0.5), (O.COSINE, 0.4), (O.OVERLAP, 0.8)):
        assert_same(gx, ox, queries[:1500], metric, alpha, 10, f"{env} m={metric} a={alpha}")
    gx.close()


def test_default_engine_is_bitmap(synth_pairs, cars_pair):
    assert cars_pair[0].layout()["engine"] == 1 and cars_pair[0].layout()["bucket_shift"] == 0
    for n in (2, 3, 4):
        lay = synth_pairs[n][0].layout()
        assert lay["engine"] == 1 and lay["row_words"] > 0, lay


def test_synthetic_stats_match_oracle(synth_pairs):
    """{admissible postings, lists} of sg_search_batch_device = SURVEY.md 8(d) ingredients from the oracle"""
    import torch
    gx, ox, queries = synth_pairs[3]
    q = queries[:200]
    data, off = pack_strings(q)
    dev = torch.device("cuda:0")
    d_q = torch.from_numpy(data).to(dev)
    d_off = torch.from_numpy(off.astype(np.int32)).to(dev)
    k = 10
    d_ids = torch.zeros(len(q) * k, dtype=torch.int32, device=dev)
    d_sc = torch.zeros(len(q) * k, dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(len(q), dtype=torch.int32, device=dev)
    d_st = torch.zeros(len(q) * 4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    gx.SuggestBatchDevice(d_q.data_ptr(), d_off.data_ptr(), len(q), 0.5, S.JaccardMetric(), k, d_ids.data_ptr(),
                          d_sc.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = d_st.cpu().numpy().reshape(-1, 4)
    ids_o, sc_o, n_o = ox.suggest_batch(q, O.JACCARD, 0.5, k, O.CANONICAL, threads=4)
    assert np.array_equal(d_cnt.cpu().numpy().astype(np.uint32), n_o)
    assert np.array_equal(d_ids.cpu().numpy().astype(np.uint32).reshape(-1, k), ids_o)
    for i, query in enumerate(q):
        want = ox.query_stats(query, O.JACCARD, 0.5)
        assert (int(st[i, 0]), int(st[i, 1])) == (want["postings"], want["lists"]), (i, query)
        assert int(st[i, 3]) == 0


# ---------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------
def test_empty_short_and_degenerate_queries(cars_pair):
    gx, ox = cars_pair
    queries = ["", " ", "a", "zz", "$", "    ", "x" * 120, "qqqqqqqqqqqqqqqqqqqqqqqq", "RAM RAM", "ram  ram", "A4", "Z"]
    for metric, alpha in ((O.JACCARD, 0.5), (O.JACCARD, 0.9), (O.COSINE, 0.3), (O.OVERLAP, 0.5), (O.EXACT, 0.5)):
        assert_same(gx, ox, queries, metric, alpha, 5, f"edge m={metric} a={alpha}")


def test_empty_batch_and_empty_dictionary():
    gx = S.NewRAMBuilder([], description(TEST_DESCRIPTION)).Build()
    assert gx.Suggest("anything", 0.5, S.JaccardMetric(), 3) == []
    ids, sc, n = gx.SuggestBatch([], 0.5, S.JaccardMetric(), 3)
    assert ids.shape == (0, 3) and n.shape == (0,)
    gx2 = S.NewRAMBuilder(["", "", "a"], description(TEST_DESCRIPTION)).Build()
    assert [c.Key for c in gx2.Suggest("a", 0.5, S.JaccardMetric(), 3)] == [2]


def test_unicode_dictionary_and_queries():
    docs = ["Жигули", "жигулёвское", "Ёлка", "ёлочка", "Москвич 412", "МОСКВА", "Nissan ёж", "日本語", "İstanbul", "straße", "STRASSE"]
    queries = ["жигули", "ЖИГУЛИ", "елка", "Ёлка", "москвич", "Москва 41", "nissan еж", "日本", "istanbul", "ıstanbul", "Straße",
               b"\xd0\xb6\xd0\xb8\xd0\xb3\xff\xd1\x83", "ж", "ЖИГУЛЁВСКОЕ"]
    gx, ox = build_pair(TEST_DESCRIPTION, docs)
    for metric, alpha in ((O.JACCARD, 0.3), (O.COSINE, 0.4), (O.DICE, 0.2)):
        assert_same(gx, ox, queries, metric, alpha, 4, f"unicode m={metric}")


def test_alternative_descriptions():
    docs = [f"item {i:04d} " + "abc"[i % 3] * (i % 5) for i in range(500)] + ["^caret$", "tail ", " lead"]
    queries = ["item 0042", "item 42", "ITEM 0499 c", "caret", "tail", "lead", "0007"]
    for desc in (dict(ngram_size=2, wrap=("", ""), pad="_", alphabet=("english", "numbers")),
                 dict(ngram_size=4, wrap=("^", "$"), pad="$", alphabet=("english", "numbers", "$^")),
                 dict(ngram_size=3, wrap=("  ", " "), pad="#", alphabet=("english", "numbers", " ")),
                 dict(ngram_size=1, wrap=("", ""), pad="?", alphabet=("numbers",))):
        gx, ox = build_pair(desc, docs)
        for metric, alpha in ((O.JACCARD, 0.4), (O.COSINE, 0.5)):
            assert_same(gx, ox, queries, metric, alpha, 6, str(desc))


def test_large_k(cars_pair, cars_lines):
    gx, ox = cars_pair
    q = cars_lines[:200]
    for k in (1, 32, 33, 250, 1024):
        assert_same(gx, ox, q, O.JACCARD, 0.15, k, f"k={k}")
    for k in (1025, 3000, 6000):  # above 1024 the per-warp top-k lives in HBM (SearchParams::tk_global)
        n = assert_same(gx, ox, q[:60], O.JACCARD, 0.01, k, f"k={k}")
        assert n.max() > 1024
    assert_same_autocomplete(gx, ox, [b"NISSAN", b"TOYOTA C", b"A"], 2000, "autocomplete limit 2000")


def test_invalid_arguments_and_too_long_query(cars_pair):
    gx, ox = cars_pair
    with pytest.raises(S.SuggestError) as e:
        gx.SuggestBatch(["a"], 0.5, S.JaccardMetric(), 0)
    assert e.value.code == _capi.SG_ERR_INVALID
    with pytest.raises(S.SuggestError):
        gx.SuggestBatch(["a"], 0.0, S.JaccardMetric(), 5)
    with pytest.raises(S.SuggestError):
        gx.SuggestBatch(["a"], 1.01, S.JaccardMetric(), 5)
    with pytest.raises(S.SuggestError):
        gx.SuggestBatch(["a"], 0.5, S.JaccardMetric(), _capi.SG_MAX_TOPK + 1)
    # a query of more than 128 n-grams is answered (host tokenization + sg_long_query_kernel), not refused
    long_q = ["ok", "y" * 300, "Nissan March", "nissan " * 30, "NISSAN MARCH " + "z" * 200]
    assert_same(gx, ox, long_q, O.JACCARD, 0.5, 5, "too long for the batched kernels")
    assert_same(gx, ox, long_q, O.OVERLAP, 0.6, 5, "too long for the batched kernels, overlap")
    ids, sc, cnt = gx.SuggestBatch(long_q, 0.5, S.JaccardMetric(), 5)
    assert cnt[2] > 0
    buf = S.PinnedBuffers(len(long_q), 5)  # the direct result path learns of such a query from a flag the kernel sets
    gx.SuggestBatch(long_q, 0.5, S.JaccardMetric(), 5, out=buf.out)
    assert np.array_equal(buf.counts, cnt) and np.array_equal(buf.ids[2, :cnt[2]], ids[2, :cnt[2]])
    gx.SuggestBatch(["ok", "Nissan March"], 0.5, S.JaccardMetric(), 5, out=(buf.ids[:2], buf.scores[:2], buf.counts[:2]))  # and the flag is per call
    with pytest.raises(S.SuggestError) as e:  # the device-buffer call still reports it per query
        import torch
        data, off = pack_strings(["y" * 300])
        d_q, d_off = torch.from_numpy(data).cuda(), torch.from_numpy(off.astype(np.int32)).cuda()
        d_ids, d_sc, d_cnt = torch.zeros(5, dtype=torch.int32).cuda(), torch.zeros(5, dtype=torch.float64).cuda(), torch.zeros(1, dtype=torch.int32).cuda()
        gx.SuggestBatchDevice(d_q.data_ptr(), d_off.data_ptr(), 1, 0.5, S.JaccardMetric(), 5, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr())
        torch.cuda.synchronize()
        if int(d_cnt.cpu().numpy().view(np.uint32)[0]) == _capi.SG_COUNT_UNSUPPORTED:
            raise S.SuggestError(_capi.SG_ERR_QUERY_TOO_LONG, "reported per query")
    assert e.value.code == _capi.SG_ERR_QUERY_TOO_LONG
    buf.close()
    with pytest.raises(S.SuggestError):
        S.NewRAMBuilder(["a"], IndexDescription(NGramSize=3, Pad="")).Build()


def test_from_lists_matches_build(cars_lines, cars_pair):
    """sg_index_from_lists (what a Go shim that decoded .hd/.dl itself would call) == sg_index_build"""
    _, ox = cars_pair
    segs, terms, ids, loff, toff = [], [], [], [0], [0]
    for seg, term, lst in ox.iter_lists():
        segs.append(seg)
        terms.append(term)
        ids.append(lst)
        loff.append(loff[-1] + len(lst))
        toff.append(toff[-1] + len(term))
    segs = np.array(segs, dtype=np.uint32)
    ids = np.concatenate(ids).astype(np.uint32)
    loff = np.array(loff, dtype=np.uint64)
    toff = np.array(toff, dtype=np.uint64)
    tb = np.frombuffer(b"".join(terms), dtype=np.uint8)
    cfg, keep = description(CARS_DESCRIPTION).c_config()
    h = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _capi.check(_capi.lib().sg_index_from_lists(C.byref(cfg), ox.segments, len(segs), p(segs), p(tb), p(toff), p(ids), p(loff), C.byref(h)))
    gx = S.NGramIndex(h.value, description(CARS_DESCRIPTION))
    assert_same(gx, ox, cars_lines[::5], O.COSINE, 0.7, 5, "from_lists")
    assert gx.info()["n_postings"] == cars_pair[0].info()["n_postings"]


def test_shards_and_merge(cars_lines, cars_pair):
    """record-id-range shards + sg_merge_topk_device == one index"""
    import torch
    _, ox = cars_pair
    n_parts, k = 3, 10
    bounds = [0, 1500, 3700, len(cars_lines)]
    shards = [S.NewRAMBuilder(cars_lines[bounds[i]:bounds[i + 1]], description(CARS_DESCRIPTION), id_base=bounds[i]).Build()
              for i in range(n_parts)]
    q = cars_lines[::4]
    nq = len(q)
    ids = np.zeros((n_parts, nq, k), dtype=np.uint32)
    sc = np.zeros((n_parts, nq, k), dtype=np.float64)
    cnt = np.zeros((n_parts, nq), dtype=np.uint32)
    for i, sh in enumerate(shards):
        ids[i], sc[i], cnt[i] = sh.SuggestBatch(q, 0.5, S.JaccardMetric(), k)
    # the reference's segment count is a per-index property: a shard sees fewer segments than the whole
    # dictionary, which only removes empty segments from the window, never candidates
    dev = torch.device("cuda:0")
    d_ids = torch.from_numpy(ids.view(np.int32)).to(dev)
    d_sc = torch.from_numpy(sc).to(dev)
    d_cnt = torch.from_numpy(cnt.view(np.int32)).to(dev)
    o_ids = torch.zeros(nq * k, dtype=torch.int32, device=dev)
    o_sc = torch.zeros(nq * k, dtype=torch.float64, device=dev)
    o_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    _capi.check(_capi.lib().sg_merge_topk_device(0, n_parts, nq, k, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(),
                                                 o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ids_o, sc_o, n_o = ox.suggest_batch(q, O.JACCARD, 0.5, k, O.CANONICAL, threads=8)
    got_n = o_cnt.cpu().numpy().astype(np.uint32)
    assert np.array_equal(got_n, n_o)
    mask = np.arange(k)[None, :] < n_o[:, None]
    got_ids = o_ids.cpu().numpy().astype(np.uint32).reshape(nq, k)
    assert np.array_equal(got_ids[mask], ids_o[mask])
    assert np.array_equal(o_sc.cpu().numpy().reshape(nq, k)[mask], sc_o[mask])
    # the packed form used by suggest_b200/sharding.py: one block per shard, as one all-gather delivers them
    L = _capi.lib()
    data, off = pack_strings([O.to_lower(x) for x in q])  # device buffers: non-ASCII queries arrive lower-cased (suggest_b200.h)
    d_q = torch.from_numpy(data).to(dev)
    d_off = torch.from_numpy(off.astype(np.int32)).to(dev)
    nbytes = int(L.sg_packed_rows_bytes(nq, k))
    blocks = torch.zeros(n_parts * nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for i, sh in enumerate(shards):
        _capi.check(L.sg_search_batch_packed_device(sh.handle, d_q.data_ptr(), d_off.data_ptr(), nq, S.JaccardMetric().code, 0.5, k,
                                                    blocks.data_ptr() + i * nbytes, st))
    torch.cuda.synchronize()
    raw = blocks.cpu().numpy()
    for i in range(n_parts):
        blk = raw[i * nbytes:(i + 1) * nbytes]
        b_sc = blk[:nq * k * 8].view(np.float64).reshape(nq, k)
        b_ids = blk[nq * k * 8:nq * k * 12].view(np.uint32).reshape(nq, k)
        b_cnt = blk[nq * k * 12:nq * k * 12 + nq * 4].view(np.uint32)
        assert np.array_equal(b_cnt, cnt[i]) and np.array_equal(b_ids, ids[i]), i
        bad = np.nonzero(b_sc != sc[i])
        assert len(bad[0]) == 0, (i, bad[0][:5], bad[1][:5], b_sc[bad][:5], sc[i][bad][:5])
    o_ids.zero_(); o_sc.zero_(); o_cnt.zero_()
    _capi.check(L.sg_merge_topk_packed_device(0, n_parts, nq, k, blocks.data_ptr(), o_ids.data_ptr(), o_sc.data_ptr(),
                                              o_cnt.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(o_cnt.cpu().numpy().astype(np.uint32), n_o)
    assert np.array_equal(o_ids.cpu().numpy().astype(np.uint32).reshape(nq, k)[mask], ids_o[mask])
    assert np.array_equal(o_sc.cpu().numpy().reshape(nq, k)[mask], sc_o[mask])


def test_concurrent_searches_on_one_handle(cars_pair, cars_lines):
    gx, ox = cars_pair
    q = cars_lines[100:400]
    want = ox.suggest_batch(q, O.COSINE, 0.6, 5, O.CANONICAL, threads=8)
    errors = []

    def worker():
        try:
            for _ in range(4):
                got = gx.SuggestBatch(q, 0.6, S.CosineMetric(), 5)
                if not (np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0])):
                    errors.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker) for _ in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:2]


def test_concurrent_callers_with_page_locked_rows():
    """three host threads, each with its own page-locked query batch and candidate rows, calling
    sg_search_batch_candidates on one index at the same time (large batches: the kernels store the rows straight into the
    callers' buffers).  Every call must equal the oracle."""
    docs, _, _ = synthetic_workload(60000, 16)
    gx = build_gpu(TEST_DESCRIPTION, (docs[0], docs[1]))
    ox = O.OracleIndex(**TEST_DESCRIPTION).add_packed(docs[0], docs[1])
    nq, k, errors = 20000, 10, []
    from suggest_b200.workload import synthetic_queries

    def worker(seed):
        try:
            q, qo, _ = synthetic_queries(docs[0], docs[1], nq, np.random.default_rng(seed))
            want = ox.suggest_batch(None, O.JACCARD, 0.5, k, O.CANONICAL, threads=2, packed=(q, qo.astype(np.uint64)))
            rows = S.PinnedCandidateRows(nq, k)
            for _ in range(6):
                rows.counts[:] = 0xFFFFFFFF
                gx.SuggestBatchCandidates(None, 0.5, S.JaccardMetric(), k, packed=(q, qo.astype(np.uint32)), out=rows.out)
                m = np.arange(k)[None, :] < want[2][:, None]
                if not (np.array_equal(rows.counts, want[2]) and np.array_equal(rows.rows["key"][m], want[0][m])
                        and np.array_equal(rows.rows["score"][m], want[1][m])):
                    errors.append(f"caller {seed}: rows differ from the oracle")
            rows.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(s_,)) for s_ in (1, 2, 3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    gx.close()
    ox.close()
    assert not errors, errors[:2]


# ---------------------------------------------------------------------------------------------------
# BASELINE.json config #2 at full size: properties + a sample against the oracle
# ---------------------------------------------------------------------------------------------------
def test_full_size_1m_jaccard():
    (d, off), (qb, q_off), pick = synthetic_workload(1_000_000, 65536)
    gx = S.NewRAMBuilder((d, off), description(TEST_DESCRIPTION)).Build()
    ox = O.OracleIndex(**TEST_DESCRIPTION).add_packed(d, off)
    info = gx.info()
    assert info["n_docs"] == 1_000_000 and info["n_postings"] == ox.postings and info["n_lists"] == ox.lists
    k = 10
    ids, sc, n = gx.SuggestBatch(None, 0.5, S.JaccardMetric(), k, packed=(qb, q_off))
    # properties that hold for every row: scores descending, ties by ascending id, scores in [alpha, 1], ids valid
    mask = np.arange(k)[None, :] < n[:, None]
    assert n.max() <= k and ids[mask].max() < 1_000_000
    assert np.all(sc[mask] >= 0.5) and np.all(sc[mask] <= 1.0)
    pair_ok = (sc[:, :-1] > sc[:, 1:]) | ((sc[:, :-1] == sc[:, 1:]) & (ids[:, :-1] < ids[:, 1:]))
    assert np.all(pair_ok | ~mask[:, 1:])
    found = n > 0
    assert found.mean() > 0.6  # two substitutions keep most queries above Jaccard 0.5 of their source entry
    assert (ids[found, 0] == pick[found]).mean() > 0.99  # and the source entry is the best match
    # idempotence: a second call returns the same bytes
    ids2, sc2, n2 = gx.SuggestBatch(None, 0.5, S.JaccardMetric(), k, packed=(qb, q_off))
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2) and np.array_equal(n, n2)
    # a sample against the oracle
    sample = np.arange(0, 65536, 16)
    queries = unpack(qb, q_off)
    ids_o, sc_o, n_o = ox.suggest_batch([queries[i] for i in sample], O.JACCARD, 0.5, k, O.CANONICAL, threads=8)
    assert np.array_equal(n[sample], n_o)
    m2 = np.arange(k)[None, :] < n_o[:, None]
    assert np.array_equal(ids[sample][m2], ids_o[m2])
    assert np.array_equal(sc[sample][m2], sc_o[m2])


# ---------------------------------------------------------------------------------------------------
# result rows in page-locked caller buffers: the kernel stores the valid entries straight into host memory
# ---------------------------------------------------------------------------------------------------
def test_pinned_result_buffers(cars_pair, cars_lines, synth_pairs):
    L = _capi.lib()
    SENT_ID, SENT_SC = 0xDEADBEEF, -7.25
    cases = [(cars_pair[0], cars_pair[1], cars_lines, O.COSINE, 0.5, 5), (cars_pair[0], cars_pair[1], cars_lines, O.JACCARD, 0.2, 40)]
    gx3, ox3, queries3 = synth_pairs[3]
    cases.append((gx3, ox3, queries3, O.JACCARD, 0.5, 10))
    for gx, ox, queries, metric, alpha, k in cases:
        buf = S.PinnedBuffers(len(queries), k)
        assert L.sg_is_pinned(buf.ids.ctypes.data, buf.ids.nbytes) == 1 and L.sg_is_pinned(buf.scores.ctypes.data, buf.scores.nbytes) == 1
        plain = np.zeros(16, dtype=np.uint32)
        assert L.sg_is_pinned(plain.ctypes.data, plain.nbytes) == 0
        buf.ids[...] = SENT_ID
        buf.scores[...] = SENT_SC
        buf.counts[...] = 0x55555555
        ids, sc, n = gx.SuggestBatch(queries, alpha, METRICS[metric], k, out=buf.out)
        ids_o, sc_o, n_o = ox.suggest_batch(queries, metric, alpha, k, O.CANONICAL, threads=8)
        assert np.array_equal(n, n_o)
        mask = np.arange(k)[None, :] < n_o[:, None]
        assert np.array_equal(ids[mask], ids_o[mask]) and np.array_equal(sc[mask], sc_o[mask])
        # the direct path writes nothing else: what lies behind a row's count is untouched
        assert np.all(ids[~mask] == SENT_ID) and np.all(sc[~mask] == SENT_SC)
        # the staged path (pageable buffers) returns the same candidates
        ids_p, sc_p, n_p = gx.SuggestBatch(queries, alpha, METRICS[metric], k)
        assert np.array_equal(n_p, n) and np.array_equal(ids_p[mask], ids[mask]) and np.array_equal(sc_p[mask], sc[mask])
        # autocomplete through the same buffers
        lim = min(k, 10)
        buf2 = S.PinnedBuffers(len(queries), lim)
        prefixes = [q[: max(3, len(q) // 2)] for q in queries]
        a_ids, a_sc, a_n = gx.AutocompleteBatch(prefixes, lim, out=buf2.out)
        b_ids, b_sc, b_n = gx.AutocompleteBatch(prefixes, lim)
        m2 = np.arange(lim)[None, :] < b_n[:, None]
        assert np.array_equal(a_n, b_n) and np.array_equal(a_ids[m2], b_ids[m2]) and np.array_equal(a_sc[m2], b_sc[m2])
        buf.close()
        buf2.close()
    # a batch large enough for the default cut into unequal slices (SG_DIRECT_SPLIT): the same rows as query by query
    big = (queries3 * 7)[:20000]
    buf = S.PinnedBuffers(len(big), 10)
    ids, sc, n = gx3.SuggestBatch(big, 0.5, METRICS[O.JACCARD], 10, out=buf.out)
    ids_1, sc_1, n_1 = gx3.SuggestBatch(queries3, 0.5, METRICS[O.JACCARD], 10)
    rep = np.arange(len(big)) % len(queries3)
    mask = np.arange(10)[None, :] < n[:, None]
    assert np.array_equal(n, n_1[rep]) and np.array_equal(ids[mask], ids_1[rep][mask]) and np.array_equal(sc[mask], sc_1[rep][mask])
    buf.close()
    # the batch cut into slices that alternate between two streams; and the scan-count engine, which stages its rows
    docs, _ = synthetic(60000, 3000)
    for env in (dict(SG_DIRECT_SLICE_QUERIES=300), dict(SG_DIRECT_SPLIT="3,5,50"), SCAN, dict(SG_DIRECT_OUT=0)):
        gx = build_gpu(TEST_DESCRIPTION, docs, env)
        buf = S.PinnedBuffers(len(queries3), 10)
        ids, sc, n = gx.SuggestBatch(queries3, 0.5, METRICS[O.JACCARD], 10, out=buf.out)
        ids_o, sc_o, n_o = ox3.suggest_batch(queries3, O.JACCARD, 0.5, 10, O.CANONICAL, threads=8)
        mask = np.arange(10)[None, :] < n_o[:, None]
        assert np.array_equal(n, n_o) and np.array_equal(ids[mask], ids_o[mask]) and np.array_equal(sc[mask], sc_o[mask]), env
        buf.close()
        gx.close()


# ---------------------------------------------------------------------------------------------------
# Autocomplete: pkg/suggest/autocomplete.go:40-77 with a FirstKCollectorManager
# ---------------------------------------------------------------------------------------------------
def assert_same_autocomplete(gx, ox, queries, limit, what=""):
    ids_g, sc_g, n_g = gx.AutocompleteBatch(queries, limit)
    for q, query in enumerate(queries):
        ids_o, sc_o = ox.autocomplete(query, limit)
        assert n_g[q] == len(ids_o), (what, query, ids_g[q, :n_g[q]], ids_o)
        assert np.array_equal(ids_g[q, :n_g[q]], ids_o), (what, query, ids_g[q, :n_g[q]], ids_o)
        assert np.array_equal(sc_g[q, :n_g[q]], sc_o), (what, query)


def test_autocomplete_ngram_index_test_go():
    # pkg/suggest/ngram_index_test.go:42-67
    gx = S.NewRAMBuilder(COLLECTION, description(TEST_DESCRIPTION)).Build()
    assert [c.Key for c in gx.Autocomplete("Niss", 5)] == [0, 1, 2, 3, 4]
    assert [c.Score for c in gx.Autocomplete("Niss", 3)] == [0.0, -1.0, -2.0]


def test_autocomplete_cars(cars_pair, cars_lines):
    gx, ox = cars_pair
    rng = np.random.default_rng(11)
    queries = ["", "N", "Ni", "Nis", "niss", "NISSAN ", "toyota c", "zzzz", "RAM RAM", "ram", "a", "4", "BMW 3", "x" * 40]
    for line in cars_lines[::97]:
        cut = int(rng.integers(1, len(line) + 1))
        queries.append(line[:cut])
    for limit in (1, 5, 50):
        assert_same_autocomplete(gx, ox, queries, limit, f"cars limit={limit}")


def test_autocomplete_synthetic_and_service(synth_pairs, cars_lines, tmp_path):
    gx, ox, queries = synth_pairs[3]
    prefixes = [q[:n] for q, n in zip(queries[:300], [3, 4, 5, 6, 8, 12] * 50)]
    assert_same_autocomplete(gx, ox, prefixes, 10, "synthetic prefixes")
    src = tmp_path / "cars.dict"
    src.write_bytes(b"\n".join(cars_lines) + b"\n")
    d = description(CARS_DESCRIPTION, "cars")
    d.SourcePath = str(src)
    service = S.NewService()
    service.AddRunTimeIndex(d)
    items = service.Autocomplete("cars", "Nissan M", 3)
    want = O.OracleIndex(**CARS_DESCRIPTION).add_docs(cars_lines).autocomplete("Nissan M", 3)[0]
    assert [i.Value for i in items] == [cars_lines[k].decode() for k in want] and len(items) == 3
    assert all(i.Value.startswith("NISSAN M") and i.Score == 0.0 for i in items)


def test_autocomplete_forced_bucket_widths(cars_lines, cars_pair):
    _, ox = cars_pair
    q = [l[:max(2, len(l) // 2)] for l in cars_lines[::41]]
    for env in (dict(SCAN, SG_FORCE_SHIFT=0, SG_TBL_BYTES=2048), dict(SCAN, SG_FORCE_SHIFT=4), dict(SCAN, SG_FORCE_SHIFT=9),
                dict(SG_BUCKET_SHIFT=0), dict(SG_BUCKET_SHIFT=4), dict(SG_BUCKET_SHIFT=8)):
        gx = build_gpu(CARS_DESCRIPTION, cars_lines, env)
        assert_same_autocomplete(gx, ox, q, 7, str(env))
        gx.close()


# ---------------------------------------------------------------------------------------------------
# randomised sweep: small alphabets give dense overlaps, ties and duplicate n-grams
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_random_dictionaries(seed):
    rng = np.random.default_rng(1000 + seed)
    letters = ["ab", "abc", "abcde", "abcdefgh", "abcdefghijkl", "ab ", "aB1", "абвг"][seed % 8]
    n_docs = int(rng.integers(50, 3000))
    docs = ["".join(rng.choice(list(letters), size=int(rng.integers(0, 24)))) for _ in range(n_docs)]
    queries = ["".join(rng.choice(list(letters), size=int(rng.integers(0, 20)))) for _ in range(150)] + docs[::max(1, n_docs // 60)]
    desc = dict(ngram_size=int(rng.integers(1, 5)), wrap=[("$", "$"), ("", ""), ("^", ""), ("<<", ">")][int(rng.integers(0, 4))],
                pad=["$", "_", "a"][int(rng.integers(0, 3))],
                alphabet=[("english", "russian", "numbers", "$"), ("english",), ("abв",), ("numbers", " ")][int(rng.integers(0, 4))])
    env = [None, dict(SCAN, SG_FORCE_SHIFT=1, SG_TBL_BYTES=2048), dict(SG_BUCKET_SHIFT=3), dict(SCAN, SG_TBL_BYTES=2048),
           dict(SG_BUCKET_SHIFT=6), dict(SCAN, SG_FORCE_SHIFT=3), dict(SG_BUCKET_SHIFT=1), dict(SG_BUCKET_SHIFT=8)][seed % 8]
    gx = build_gpu(desc, docs, env)
    ox = O.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    for metric in (O.JACCARD, O.COSINE, O.DICE, O.OVERLAP, O.EXACT):
        alpha = float(rng.choice([0.2, 0.34, 0.5, 0.75, 1.0]))
        k = int(rng.choice([1, 3, 10, 37]))
        assert_same(gx, ox, queries, metric, alpha, k, f"seed={seed} desc={desc} m={metric} a={alpha} k={k} env={env}")
    assert_same_autocomplete(gx, ox, queries[:60], int(rng.choice([1, 4, 20])), f"seed={seed} autocomplete")
    gx.close()


# ---------------------------------------------------------------------------------------------------
# device index build (sg_gpubuild.cu) against the host build (sg_index.cpp)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["cars", "synthetic3", "synthetic2", "synthetic4", "unicode"])
def test_device_build_equals_host_build(cars_lines, case):
    if case == "cars":
        desc, docs, queries = CARS_DESCRIPTION, cars_lines, cars_lines[::5]
    elif case == "unicode":
        desc = TEST_DESCRIPTION
        docs = ["Жигули", "жигули 2106", "Ёлка", "ёж", "İstanbul", "naïve café", "日本語", "abc", "", "a", "RAM RAM", b"bad\xff\xfe bytes"] * 3
        queries = ["жигули", "ЁЛКА", "istanbul", "cafe", "ram ram", "日本", ""]
    else:
        n = int(case[-1])
        desc = dict(TEST_DESCRIPTION, ngram_size=n)
        docs, queries = synthetic(30000, 1500, seed=77)
    host = build_gpu(desc, docs, dict(SG_BUILD="host"))
    dev = build_gpu(desc, docs, dict(SG_BUILD="gpu"))
    assert host.layout()["built_on_device"] == 0 and dev.layout()["built_on_device"] == 1
    hi, di = host.info(), dev.info()
    for key in ("n_docs", "n_segments", "n_terms", "n_lists", "n_postings"):
        assert hi[key] == di[key], (key, hi[key], di[key])
    hl, dl = host.layout(), dev.layout()
    for key in ("n_slots", "bucket_shift", "row_words", "engine", "bitmap_bytes"):
        assert hl[key] == dl[key], (key, hl[key], dl[key])
    for metric, alpha, k in ((S.JaccardMetric(), 0.5, 10), (S.CosineMetric(), 0.3, 7), (S.OverlapMetric(), 0.8, 3)):
        a = host.SuggestBatch(queries, alpha, metric, k)
        b = dev.SuggestBatch(queries, alpha, metric, k)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    a = host.AutocompleteBatch([q[:3] for q in queries], 6)
    b = dev.AutocompleteBatch([q[:3] for q in queries], 6)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    host.close()
    dev.close()


def test_device_build_falls_back_for_long_documents(cars_pair):
    docs = ["short one", "x" * 300, "another short entry", "y" * 140 + " tail"]
    gx = build_gpu(TEST_DESCRIPTION, docs)
    assert gx.layout()["built_on_device"] == 0          # a document of more than 128 n-grams: host build
    ox = O.OracleIndex(**{"ngram_size": 3, "wrap": ("$", "$"), "pad": "$", "alphabet": TEST_DESCRIPTION["alphabet"]}).add_docs(docs)
    assert_same(gx, ox, ["short one", "another short", "x" * 40, "tail"], O.JACCARD, 0.3, 3)
    gx.close()
    with pytest.raises(_capi.SuggestError):
        build_gpu(TEST_DESCRIPTION, docs, dict(SG_BUILD="gpu"))
    assert cars_pair[0].layout()["built_on_device"] == 1  # the default build of the suite's dictionaries is the device build


def test_long_documents_and_queries():
    """40-126 n-grams per entry: the 8-plane adders (32 or more lists), many segments, k-th score ties; and entries of
    more than 128 n-grams in the dictionary (host build, more than 256 segments) next to queries that are too long"""
    rng = np.random.default_rng(404)
    docs = ["".join(chr(97 + c) for c in rng.integers(0, 8, size=int(rng.integers(38, 125)))) for _ in range(3000)]
    queries = []
    for d in docs[::10]:
        q = list(d)
        for _ in range(int(rng.integers(0, 6))):
            q[int(rng.integers(len(q)))] = chr(97 + int(rng.integers(0, 8)))
        queries.append("".join(q)[:int(rng.integers(30, len(q) + 1))])
    gx, ox = build_pair(TEST_DESCRIPTION, docs)
    assert gx.layout()["built_on_device"] == 1
    for metric, alpha, k in ((O.JACCARD, 0.5, 10), (O.COSINE, 0.35, 5), (O.DICE, 0.6, 3), (O.OVERLAP, 0.9, 4)):
        n = assert_same(gx, ox, queries, metric, alpha, k, f"long m={metric} a={alpha}")
        assert n.sum() > 0
    assert_same_autocomplete(gx, ox, [q[:int(rng.integers(3, 60))] for q in queries[:80]], 5, "long autocomplete")
    gx.close()
    docs2 = docs[:500] + ["".join(chr(97 + c) for c in rng.integers(0, 26, size=int(rng.integers(150, 400)))) for _ in range(40)]
    gx, ox = build_pair(TEST_DESCRIPTION, docs2)
    assert gx.layout()["built_on_device"] == 0 and gx.info()["n_segments"] > 256
    q2 = queries[:40] + [d[:100] for d in docs2[500:520]]
    assert_same(gx, ox, q2, O.JACCARD, 0.3, 10, "long docs host build")
    assert_same(gx, ox, q2, O.OVERLAP, 0.8, 10, "long docs host build overlap")
    # queries of 129 .. 399 n-grams: answered like the reference answers them (pkg/merger/list_merger.go:9 saturates at 0xFFFF)
    q3 = [d[:int(rng.integers(131, len(d) + 1))] for d in docs2[500:]] + [docs2[-1], docs2[-2][5:] + "qq", "short"] + q2[:5]
    for metric, alpha, k in ((O.JACCARD, 0.3, 10), (O.COSINE, 0.5, 4), (O.DICE, 0.4, 7), (O.OVERLAP, 0.7, 5), (O.EXACT, 1.0, 2)):
        n = assert_same(gx, ox, q3, metric, alpha, k, f"long queries m={metric}")
        assert n[:40].sum() > 0
    assert_same_autocomplete(gx, ox, [d[:int(rng.integers(131, 150))] for d in docs2[500:520]] + ["ab"], 5, "long autocomplete queries")
    buf = S.PinnedBuffers(len(q3), 10)  # and through the direct result path
    ids, sc, cnt = gx.SuggestBatch(q3, 0.3, S.JaccardMetric(), 10)
    gx.SuggestBatch(q3, 0.3, S.JaccardMetric(), 10, out=buf.out)
    assert np.array_equal(buf.counts, cnt)
    assert all(np.array_equal(buf.ids[i, :cnt[i]], ids[i, :cnt[i]]) for i in range(len(q3)))
    gx.close()


@pytest.mark.parametrize("chunks", ["4", "0"])
def test_chunked_arrival_equals_sliced_path(chunks, monkeypatch):
    """sg_search_batch with page-locked rows and >= 16,384 queries.  SG_DIRECT_CHUNKS=4: one launch, queries arriving in
    chunks while the kernels run; 0 (the default): slices on two streams, rows stored into the caller's buffers.  Both must
    equal the staged path (pageable rows) query by query, including chunks that hold non-ASCII bytes (lower-cased on the host
    into their own area), a query of more than 128 n-grams, empty queries and a ragged last chunk."""
    monkeypatch.setenv("SG_DIRECT_CHUNKS", chunks)  # read when the index is created
    docs, (qb, qo), _ = synthetic_workload(50000, 20011)
    queries = unpack(qb, qo)
    queries[5] = b""
    queries[9000] = "ЖИГУЛИ nissan".encode("utf-8")          # chunk 1: non-ASCII
    queries[9001] = b"\xff\xfe" + queries[9001]
    queries[17000] = queries[17001] * 9                          # more than 128 n-grams
    queries[20010] = "Ünïcode tail".encode("utf-8")             # last, ragged chunk
    gx = build_gpu(TEST_DESCRIPTION, (docs[0], docs[1]))
    data, off = pack_strings(queries)
    for metric, alpha, k in ((S.JaccardMetric(), 0.5, 10), (S.CosineMetric(), 0.6, 3)):
        ids, sc, cnt = gx.SuggestBatch(None, alpha, metric, k, packed=(data, off))
        buf = S.PinnedBuffers(len(queries), k)
        for _ in range(3):  # the arrival counter keeps growing across calls
            gx.SuggestBatch(None, alpha, metric, k, packed=(data, off), out=buf.out)
            assert np.array_equal(buf.counts, cnt)
            m = np.arange(k)[None, :] < cnt[:, None]
            assert np.array_equal(buf.ids[m], ids[m]) and np.array_equal(buf.scores[m], sc[m])
        assert cnt.sum() > 10000
        buf.close()
    ids, sc, cnt = gx.AutocompleteBatch(None, 5, packed=(data, off))
    buf = S.PinnedBuffers(len(queries), 5)
    gx.AutocompleteBatch(None, 5, packed=(data, off), out=buf.out)
    m = np.arange(5)[None, :] < cnt[:, None]
    assert np.array_equal(buf.counts, cnt) and np.array_equal(buf.ids[m], ids[m])
    gx.close()


def test_candidate_rows_equal_separate_arrays(cars_pair, cars_lines):
    """sg_search_batch_candidates (rows of 16-byte {key, 0, score} entries, suggest.Candidate's layout) returns what
    sg_search_batch returns: pageable rows (staged), page-locked rows (stored by the kernels), a batch large enough for the
    chunked-arrival path, a query of more than 128 n-grams, k above 1024."""
    gx, ox = cars_pair
    base = list(cars_lines[:700]) + [b"", "ЖИГУЛИ".encode("utf-8")]
    for metric, alpha, k in ((S.JaccardMetric(), 0.5, 10), (S.CosineMetric(), 0.3, 50), (S.JaccardMetric(), 0.05, 1500)):
        queries = base + ([b"nissan " * 30] if k <= 1024 else [])  # (a query of more than 128 n-grams is served up to topK 1024)
        ids, sc, cnt = gx.SuggestBatch(queries, alpha, metric, k)
        m = np.arange(k)[None, :] < cnt[:, None]
        rows, n = gx.SuggestBatchCandidates(queries, alpha, metric, k)
        assert np.array_equal(n, cnt) and np.array_equal(rows["key"][m], ids[m]) and np.array_equal(rows["score"][m], sc[m])
        assert not rows["reserved"][m].any()
        buf = S.PinnedCandidateRows(len(queries), k)
        buf.rows["key"][...] = 0xDEADBEEF
        rows, n = gx.SuggestBatchCandidates(queries, alpha, metric, k, out=buf.out)
        assert np.array_equal(n, cnt) and np.array_equal(rows["key"][m], ids[m]) and np.array_equal(rows["score"][m], sc[m])
        assert np.all(rows["key"][~m] == 0xDEADBEEF)  # the direct path writes the valid entries only
        buf.close()
    docs, (qb, qo), _ = synthetic_workload(40000, 18000)
    big = build_gpu(TEST_DESCRIPTION, (docs[0], docs[1]))
    ids, sc, cnt = big.SuggestBatch(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo))
    buf = S.PinnedCandidateRows(18000, 10)
    rows, n = big.SuggestBatchCandidates(None, 0.5, S.JaccardMetric(), 10, packed=(qb, qo), out=buf.out)
    m = np.arange(10)[None, :] < cnt[:, None]
    assert np.array_equal(n, cnt) and np.array_equal(rows["key"][m], ids[m]) and np.array_equal(rows["score"][m], sc[m]) and cnt.sum() > 9000
    buf.close()
    big.close()


# ---------------------------------------------------------------------------------------------------
# the reference's real-language dictionary (pkg/suggest/testdata/words.dict), both pipelines
# ---------------------------------------------------------------------------------------------------
WORDS_DESCRIPTION = dict(ngram_size=3, wrap=("^", "$"), pad="$", alphabet=("english", "numbers", "$^"))


@pytest.mark.parametrize("pipeline", ["classic", "lean"])
def test_words_dict_against_oracle(pipeline):
    """235,886 English words, one bit per document: frequent n-grams, low thresholds, hundreds of candidates per query -
    the launch-wide scratch of the count -> resolve pipeline overflows and the fallback kernel answers what is left"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "words.dict")
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    rng = np.random.default_rng(31)
    queries = []
    for i in rng.integers(0, len(lines), size=6000):
        w = bytearray(lines[int(i)])
        if w:
            w[int(rng.integers(0, len(w)))] = int(rng.integers(97, 123))
        queries.append(bytes(w))
    gx = build_gpu(WORDS_DESCRIPTION, lines, dict(SG_PIPELINE=pipeline))
    assert gx.layout()["pipeline"] == (1 if pipeline == "lean" else 0) and gx.layout()["bucket_shift"] == 0
    ox = O.OracleIndex(**WORDS_DESCRIPTION).add_docs(lines)
    for metric, alpha, k in ((O.JACCARD, 0.5, 10), (O.COSINE, 0.5, 10), (O.DICE, 0.5, 5)):
        n = assert_same(gx, ox, queries, metric, alpha, k, f"words.dict {pipeline} m={metric}")
        assert n.sum() > 3000
    # the same through page-locked rows, and once more: the scratch is per call
    data, off = pack_strings(queries)
    ids, sc, cnt = gx.SuggestBatch(None, 0.5, S.CosineMetric(), 10, packed=(data, off))
    buf = S.PinnedBuffers(len(queries), 10)
    for _ in range(2):
        gx.SuggestBatch(None, 0.5, S.CosineMetric(), 10, packed=(data, off), out=buf.out)
        m = np.arange(10)[None, :] < cnt[:, None]
        assert np.array_equal(buf.counts, cnt) and np.array_equal(buf.ids[m], ids[m]) and np.array_equal(buf.scores[m], sc[m])
    buf.close()
    gx.close()


@pytest.mark.parametrize("flags,nodes", [(1, 8), (64, 0), (2, 1), (0, 8)])
def test_pipeline_scratch_overflow_against_oracle(flags, nodes):
    """The count -> resolve pipeline with its launch-wide scratch shrunk (SG_LEAN_FLAGS_PER_QUERY / SG_LEAN_NODES_PER_QUERY
    entries per query, pooled): the list of flagged words overflows, the survivor nodes overflow, both, or nothing fits at
    all - the queries concerned are marked dirty at different points of the pipeline and answered by the fallback kernel.
    Every query must equal the oracle, whichever kernel answered it."""
    docs, (qb, qo), _ = synthetic_workload(40000, 3000, seed=77)
    queries = unpack(qb, qo)
    gx = build_gpu(TEST_DESCRIPTION, (docs[0], docs[1]), dict(SG_LEAN_FLAGS_PER_QUERY=flags, SG_LEAN_NODES_PER_QUERY=nodes, SG_BUCKET_SHIFT=4))
    assert gx.layout()["pipeline"] == 1
    ox = O.OracleIndex(**TEST_DESCRIPTION).add_packed(docs[0], docs[1])
    for metric, alpha, k in ((O.JACCARD, 0.3, 10), (O.COSINE, 0.4, 40), (O.DICE, 0.5, 5)):
        n = assert_same(gx, ox, queries, metric, alpha, k, f"scratch {flags}/{nodes} m={metric}")
        assert n.sum() > 2000
    gx.close()
    ox.close()


def test_submit_and_wait_equal_the_synchronous_call():
    """sg_search_batch_candidates_submit / sg_ticket_wait: five batches in flight from this one thread (three worker threads
    of the library), page-locked and pageable rows, an invalid call (its status and message arrive at wait), and tickets
    still open when the index is freed."""
    docs, _, _ = synthetic_workload(60000, 16)
    gx = build_gpu(TEST_DESCRIPTION, (docs[0], docs[1]))
    from suggest_b200.workload import synthetic_queries
    nq, k = 20000, 10
    batches, want, tickets, bufs = [], [], [], []
    for seed in range(5):
        q, qo, _ = synthetic_queries(docs[0], docs[1], nq, np.random.default_rng(100 + seed))
        batches.append((q, qo.astype(np.uint32)))
        rows, counts = gx.SuggestBatchCandidates(None, 0.5, S.JaccardMetric(), k, packed=batches[-1])
        want.append((rows.copy(), counts.copy()))
    for i, b in enumerate(batches):
        if i % 2 == 0:
            buf = S.PinnedCandidateRows(nq, k)
            out = buf.out
        else:  # pageable rows: staged in HBM and copied
            buf = None
            out = (np.zeros((nq, k), dtype=S.CANDIDATE_DTYPE), np.zeros(nq, dtype=np.uint32))
        bufs.append(buf)
        tickets.append(gx.SubmitBatchCandidates(0.5, S.JaccardMetric(), k, b, out))
    for t, (w_rows, w_counts) in zip(tickets, want):
        rows, counts = t.wait()
        m = np.arange(k)[None, :] < w_counts[:, None]
        assert np.array_equal(counts, w_counts)
        assert np.array_equal(rows["key"][m], w_rows["key"][m]) and np.array_equal(rows["score"][m], w_rows["score"][m])
    with pytest.raises(RuntimeError):
        tickets[0].wait()  # a ticket is waited for once
    bad = gx.SubmitBatchCandidates(1.5, S.JaccardMetric(), k, batches[0], bufs[0].out)  # similarity out of range
    with pytest.raises(_capi.SuggestError, match="similarity"):
        bad.wait()
    # tickets still open when the index goes: sg_index_free serves them first
    open_tickets = [gx.SubmitBatchCandidates(0.5, S.JaccardMetric(), k, batches[i], bufs[i].out) for i in (0, 2, 4)]
    gx.close()
    for t, i in zip(open_tickets, (0, 2, 4)):
        rows, counts = t.wait()
        assert np.array_equal(counts, want[i][1])
    for b in bufs:
        if b is not None:
            b.close()


def test_autocomplete_device_entry_equals_host_entry(cars_pair, cars_lines):
    """sg_autocomplete_batch_device (buffers resident on the device) against sg_autocomplete_batch (itself checked against
    the oracle above), the stage-time aid, and the sanity of the stats pass in autocomplete mode"""
    import torch
    gx, ox = cars_pair
    prefixes = [bytes(line[:n]).lower() for line in cars_lines[:400] for n in (3, 5, 8) if len(line) >= n]
    limit = 7
    ids, sc, cnt = gx.AutocompleteBatch(prefixes, limit)
    data, off = pack_strings(prefixes)
    dev = torch.device("cuda", 0)
    dq, doff = torch.from_numpy(np.ascontiguousarray(data)).to(dev), torch.from_numpy(off.astype(np.int32)).to(dev)
    n = len(prefixes)
    d_ids = torch.zeros(n * limit, dtype=torch.int32, device=dev)
    d_sc = torch.zeros(n * limit, dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    d_stats = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    gx.AutocompleteBatchDevice(dq.data_ptr(), doff.data_ptr(), n, limit, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), d_stats.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g_cnt = d_cnt.cpu().numpy().view(np.uint32)
    m = np.arange(limit)[None, :] < cnt[:, None]
    assert np.array_equal(g_cnt, cnt)
    assert np.array_equal(d_ids.cpu().numpy().view(np.uint32).reshape(n, limit)[m], ids[m])
    assert np.array_equal(d_sc.cpu().numpy().reshape(n, limit)[m], sc[m])
    times = gx.AutocompleteStageTimes(dq.data_ptr(), doff.data_ptr(), n, limit, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    assert "sg_bitmap_search_kernel" in times and all(v >= 0 for v in times.values())
    # the stats pass (SURVEY.md 8(d) figures for Autocomplete): a prefix with completions has admissible lists, every list a posting
    stats = d_stats.cpu().numpy().view(np.uint32).reshape(n, 4)
    assert np.all(stats[cnt > 0, 1] > 0) and np.all(stats[:, 0] >= stats[:, 1]) and np.all(stats[cnt > 0, 0] >= cnt[cnt > 0])
