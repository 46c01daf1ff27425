"""sg_disk.cpp (index.Reader.Read replacement) against the reference's shipped on-disk indexes.  No GPU."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION, GOLDEN, REFERENCE_TESTDATA
from diskfmt import gob_header, roaring_blob
from oracle import oracle
from suggest_b200 import _capi

WORDS_DESCRIPTION = dict(ngram_size=3, wrap=("^", "$"), pad="$", alphabet=("english", "numbers", "$^"))


def open_disk(desc, hd, dl):
    cfg, keep = _capi.make_config(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"])
    h = C.c_void_p()
    rc = _capi.lib().sg_host_index_open_disk(C.byref(cfg), hd.encode(), dl.encode(), C.byref(h))
    _capi.check(rc, host=True)
    return h


def get_list(h, seg, term):
    n = _capi.lib().sg_host_index_get_list(h, seg, term, len(term), None, 0)
    if n < 0:
        return None
    out = np.zeros(n, dtype=np.uint32)
    _capi.lib().sg_host_index_get_list(h, seg, term, len(term), out.ctypes.data_as(C.c_void_p), n)
    return out


def check_against_rebuild(desc, lines, hd, dl):
    ox = oracle.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(lines)
    h = open_disk(desc, hd, dl)
    try:
        info = _capi.SgIndexInfo()
        _capi.lib().sg_host_index_get_info(h, C.byref(info))
        assert info.n_segments == ox.segments and info.n_lists == ox.lists
        n = 0
        for seg, term, ids in ox.iter_lists():
            got = get_list(h, seg, term)
            assert got is not None and np.array_equal(got, np.unique(ids)), (seg, term)
            n += 1
        assert n == info.n_lists
    finally:
        _capi.lib().sg_host_index_free(h)


def test_cars_db_equals_rebuild_from_dict(cars_lines):
    # pkg/suggest/testdata/db/cars.{hd,dl}: 36,276 VB lists + 9 skipping lists
    check_against_rebuild(CARS_DESCRIPTION, cars_lines, os.path.join(GOLDEN, "cars.hd"), os.path.join(GOLDEN, "cars.dl"))


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTDATA), reason="reference checkout not present")
def test_words_db_equals_rebuild_from_dict():
    # 71,217 VB + 6,857 skipping + 1,378 roaring lists
    with open(os.path.join(REFERENCE_TESTDATA, "words.dict"), "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    check_against_rebuild(WORDS_DESCRIPTION, lines, os.path.join(REFERENCE_TESTDATA, "db", "words.hd"),
                          os.path.join(REFERENCE_TESTDATA, "db", "words.dl"))


def test_roaring_samples_from_words_dl():
    """A few roaring blobs cut out of the reference's words.dl (tests/golden/make_golden.py) decode to the committed ids."""
    path = os.path.join(GOLDEN, "roaring_samples.npz")
    z = np.load(path)
    n = int(z["n"])
    assert n >= 8
    for i in range(n):
        blob, want = z[f"blob{i}"].tobytes(), z[f"ids{i}"]
        got = decode_single_list(blob, len(want))
        assert np.array_equal(got, want), i


def decode_single_list(blob, length, tmp=[0]):
    import tempfile
    d = tempfile.mkdtemp()
    hd, dl = os.path.join(d, "x.hd"), os.path.join(d, "x.dl")
    with open(dl, "wb") as f:
        f.write(blob)
    with open(hd, "wb") as f:
        f.write(gob_header([(b"abc", 3, len(blob), 0, length)], 4))
    h = open_disk(dict(ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("english", "$")), hd, dl)
    try:
        return get_list(h, 3, b"abc")
    finally:
        _capi.lib().sg_host_index_free(h)


@pytest.mark.parametrize("runs", [False, True])
def test_roaring_container_kinds(runs):
    rng = np.random.default_rng(3)
    sparse = rng.choice(1 << 20, size=700, replace=False)                       # array containers, many keys
    dense = np.concatenate([np.arange(70000, 76000), rng.choice(65536, 6000, replace=False) + (5 << 16)])  # bitmap containers
    runny = np.concatenate([np.arange(100, 900), np.arange(2000, 2400), np.arange(1 << 16, (1 << 16) + 300)])
    for values in (sparse, dense, runny, np.concatenate([sparse, dense, runny])):
        want = np.unique(values.astype(np.uint32))
        got = decode_single_list(roaring_blob(values, runs), len(want))
        assert np.array_equal(got, want)


def test_codecs_through_oracle_encoder():
    """VB and skipping(64) bytes produced by the oracle's encoder (itself pinned to cars.dl) decode identically"""
    rng = np.random.default_rng(5)
    for n in (1, 2, 65, 66, 127, 128, 129, 200, 256):
        lst = np.sort(rng.choice(5_000_000, size=n, replace=False)).astype(np.uint32)
        codec = oracle.CODEC_VB if n <= 65 else oracle.CODEC_SKIPPING
        assert np.array_equal(decode_single_list(oracle.encode(codec, lst, 64), n), lst)


def test_bad_files_are_format_errors(tmp_path):
    hd, dl = tmp_path / "a.hd", tmp_path / "a.dl"
    dl.write_bytes(b"\x01\x02")
    for content in (b"", b"\x05abc", gob_header([(b"abc", 3, 9, 0, 3)], 4), gob_header([(b"abc", 3, 2, 0, 3)], 4).replace(b"v5.1", b"v4.0")):
        hd.write_bytes(content)
        cfg, keep = _capi.make_config(3, ("$", "$"), "$", ("english", "$"))
        h = C.c_void_p()
        rc = _capi.lib().sg_host_index_open_disk(C.byref(cfg), str(hd).encode(), str(dl).encode(), C.byref(h))
        assert rc == _capi.SG_ERR_FORMAT, content
    rc = _capi.lib().sg_host_index_open_disk(C.byref(cfg), b"/nonexistent.hd", str(dl).encode(), C.byref(h))
    assert rc == _capi.SG_ERR_IO


def test_written_index_with_all_three_codecs(tmp_path):
    """diskfmt.write_index (what tests/test_gpu_disk.py feeds sg_index_open_disk) reads back list by list"""
    from diskfmt import write_index
    from suggest_b200.workload import synthetic_dictionary, unpack
    desc = dict(ngram_size=2, wrap=("$", "$"), pad="$", alphabet=("english", "$"))
    data, off, _ = synthetic_dictionary(70000, seed=99, lo=4, hi=14, skew="zipf")
    docs = unpack(data, off)
    ox = oracle.OracleIndex(desc["ngram_size"], desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(docs)
    n_lists, n_roaring = write_index(ox, str(tmp_path), "synth")
    assert n_roaring > 50
    check_against_rebuild(desc, docs, str(tmp_path / "synth.hd"), str(tmp_path / "synth.dl"))
