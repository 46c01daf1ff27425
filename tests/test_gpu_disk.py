"""DISC driver on the GPU: sg_index_open_disk / NewFSBuilder / Service.AddOnDiscIndex against the reference's shipped
index files and against RAM-built indexes of the same dictionaries.  Needs a B200: `pytest -m gpu`.

pkg/suggest/service_test.go:25-33 runs its expectations through AddOnDiscIndex as well as AddRunTimeIndex, and the
shipped pkg/suggest/testdata/config.json uses `driver: DISC`; this is that half of the test.
"""
import json
import os
import shutil

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION, GOLDEN, REFERENCE_TESTDATA
from diskfmt import write_index
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200.suggest import IndexDescription, ReadConfigs
from suggest_b200.workload import synthetic_dictionary, synthetic_queries, unpack

pytestmark = pytest.mark.gpu

WORDS_DESCRIPTION = dict(ngram_size=3, wrap=("^", "$"), pad="$", alphabet=("english", "numbers", "$^"))


def disc_description(d, name, output, source=""):
    return IndexDescription(Name=name, NGramSize=d["ngram_size"], Alphabet=tuple(d["alphabet"]), Pad=d["pad"], Wrap=tuple(d["wrap"]),
                            Driver="DISC", OutputPath=output, SourcePath=source)


def same_results(a, b, queries, metric, alpha, k):
    ia, sa, na = a.SuggestBatch(queries, alpha, metric, k)
    ib, sb, nb = b.SuggestBatch(queries, alpha, metric, k)
    assert np.array_equal(na, nb)
    m = np.arange(k)[None, :] < na[:, None]
    assert np.array_equal(ia[m], ib[m]) and np.array_equal(sa[m], sb[m])
    return na


def perturbed(lines, n, seed=7):
    rng = np.random.default_rng(seed)
    out = []
    for i in rng.integers(0, len(lines), size=n):
        w = bytearray(lines[int(i)])
        for _ in range(2):
            if w:
                w[int(rng.integers(0, len(w)))] = int(rng.integers(97, 123))
        out.append(bytes(w))
    return out


def test_service_test_go_disc_driver(cars_lines, tmp_path):
    # pkg/suggest/service_test.go:11-59 with the DISC branch of :25-33, on the reference's own cars.{hd,dl}
    db = tmp_path / "db"
    db.mkdir()
    for name in ("cars.hd", "cars.dl"):
        shutil.copyfile(os.path.join(GOLDEN, name), db / name)
    (tmp_path / "cars.dict").write_bytes(b"\n".join(cars_lines) + b"\n")
    cfg = [dict(driver="DISC", name="cars", nGramSize=3, alphabet=["russian", "english", "numbers", "$"], source="cars.dict",
                output="db", pad="$", wrap=["$", "$"])]  # pkg/suggest/testdata/config.json, first entry
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    descriptions = ReadConfigs(str(tmp_path / "config.json"))
    assert descriptions[0].Driver == "DISC"
    service = S.NewService()
    for d in descriptions:
        service.AddIndexByDescription(d)
    words = ["Nissan March", "Honda Fitt", "Wolfsvagen", "Tayota Corolla", "Micra Nissan"]
    expected = [["NISSAN MARCH"], ["HONDA FIT"], [], ["TOYOTA COROLLA"], ["NISSAN MICRA"]]
    for w, exp in zip(words, expected):
        res = service.Suggest("cars", S.NewSearchConfig(w, 5, S.CosineMetric(), 0.7))
        assert [r.Value for r in res] == exp, (w, res)


def test_cars_disk_index_equals_ram_index(cars_lines):
    disk = S.NewFSBuilder(disc_description(CARS_DESCRIPTION, "cars", GOLDEN)).Build()
    ram = S.NewRAMBuilder(cars_lines, disc_description(CARS_DESCRIPTION, "cars", GOLDEN)).Build()
    di, ri = disk.info(), ram.info()
    for key in ("n_docs", "n_segments", "n_terms", "n_lists"):
        assert di[key] == ri[key], key
    # the shipped lists keep the reference's repeated ids (SURVEY 8c rule 5b); they collapse to the same documents
    assert di["n_postings"] >= ri["n_postings"]
    queries = perturbed(cars_lines, 1000) + list(cars_lines[:200])
    for metric, alpha, k in ((S.CosineMetric(), 0.7, 5), (S.JaccardMetric(), 0.5, 10), (S.DiceMetric(), 0.4, 20)):
        n = same_results(disk, ram, queries, metric, alpha, k)
        assert n.sum() > 0
    got = disk.Autocomplete("niss", 5)
    assert [c.Key for c in got] == [c.Key for c in ram.Autocomplete("niss", 5)]
    disk.close()
    ram.close()


def test_written_index_with_roaring_lists(tmp_path):
    """An index with VB, skipping(64) and roaring lists (array, bitmap and run containers), written the way
    Writer.Commit does, searched from disk: equal to the RAM build and to the oracle."""
    desc = dict(ngram_size=2, wrap=("$", "$"), pad="$", alphabet=("english", "$"))
    data, off, rng = synthetic_dictionary(70000, seed=99, lo=4, hi=14, skew="zipf")
    docs = unpack(data, off)
    ox = O.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    n_lists, n_roaring = write_index(ox, str(tmp_path), "synth")
    assert n_roaring > 50 and n_lists > n_roaring
    disk = S.NewFSBuilder(disc_description(desc, "synth", str(tmp_path))).Build()
    ram = S.NewRAMBuilder(docs, disc_description(desc, "synth", str(tmp_path))).Build()
    assert disk.info()["n_lists"] == n_lists == ram.info()["n_lists"]
    assert disk.info()["n_postings"] == ram.info()["n_postings"]
    q, q_off, _ = synthetic_queries(data, off, 1000, rng, subs=1)
    queries = unpack(q, q_off)
    for metric, om, alpha, k in ((S.JaccardMetric(), O.JACCARD, 0.6, 10), (S.CosineMetric(), O.COSINE, 0.7, 5)):
        n = same_results(disk, ram, queries, metric, alpha, k)
        ids, sc, cnt = disk.SuggestBatch(queries, alpha, metric, k)
        o_ids, o_sc, o_cnt = ox.suggest_batch(queries, om, alpha, k, O.CANONICAL, threads=8)
        assert np.array_equal(cnt, o_cnt)
        m = np.arange(k)[None, :] < o_cnt[:, None]
        assert np.array_equal(ids[m], o_ids[m]) and np.array_equal(sc[m], o_sc[m])
        assert n.sum() > 0
    disk.close()
    ram.close()


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTDATA), reason="reference checkout not present")
def test_words_disk_index_equals_ram_index():
    # the reference's words.{hd,dl}: 71,217 VB + 6,857 skipping + 1,378 roaring lists written by the Go indexer
    with open(os.path.join(REFERENCE_TESTDATA, "words.dict"), "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    d = disc_description(WORDS_DESCRIPTION, "words", os.path.join(REFERENCE_TESTDATA, "db"))
    disk = S.NewFSBuilder(d).Build()
    ram = S.NewRAMBuilder(lines, d).Build()
    assert disk.info()["n_lists"] == ram.info()["n_lists"]
    same_results(disk, ram, perturbed(lines, 1000), S.JaccardMetric(), 0.5, 10)
    disk.close()
    ram.close()


def test_invalid_bytes_triple_when_lowered(cars_lines):
    """strings.ToLower turns every invalid UTF-8 byte into U+FFFD (1 -> 3 bytes): a batch of 0xFF / CP1251 bytes lowers
    to three times its size.  First call on a fresh handle, so the device query buffer is sized by this batch."""
    gx = S.NewRAMBuilder(cars_lines, disc_description(CARS_DESCRIPTION, "cars", GOLDEN)).Build()
    ox = O.OracleIndex(CARS_DESCRIPTION["ngram_size"], CARS_DESCRIPTION["wrap"], CARS_DESCRIPTION["pad"],
                       CARS_DESCRIPTION["alphabet"]).add_docs(cars_lines)
    cp1251 = "ниссан марч".encode("cp1251")
    queries = [b"\xff" * 60, cp1251 * 3, b"\xff\xfe NISSAN \xff MARCH \xfd" * 2] * 40 + [b"\xff" * 100] * 20
    assert sum(len(q) for q in queries) > 1024
    ids, sc, cnt = gx.SuggestBatch(queries, 0.3, S.JaccardMetric(), 5)
    o_ids, o_sc, o_cnt = ox.suggest_batch(queries, O.JACCARD, 0.3, 5, O.CANONICAL, threads=4)
    assert np.array_equal(cnt, o_cnt)
    m = np.arange(5)[None, :] < o_cnt[:, None]
    assert np.array_equal(ids[m], o_ids[m]) and np.array_equal(sc[m], o_sc[m])
    # and the handle still answers ordinary queries afterwards
    got = gx.Suggest("Nissan March", 0.7, S.CosineMetric(), 5)
    assert [cars_lines[c.Key] for c in got] == [b"NISSAN MARCH"]
    gx.close()
