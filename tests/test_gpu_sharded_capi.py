"""sg_sharded_*: record-id-range shards driven by one process; results must equal the unsharded index
and the oracle.  Runs on one GPU (shards share it) and uses every GPU the box has.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200 import _capi
from suggest_b200.sharding import ShardedNGramIndex
from suggest_b200.suggest import IndexDescription
from suggest_b200.workload import synthetic_workload
from test_gpu_parity import METRICS, description

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def device_plans():
    n = n_gpus()
    plans = [[0], [0, 0, 0]]
    if n >= 2:
        plans += [list(range(n)), [i % n for i in range(2 * n + 1)]]
    return plans


def with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


@pytest.mark.parametrize("gather", ["peer", "copy"])
def test_sharded_equals_unsharded_and_oracle(gather):
    docs, (qb, qo), _ = synthetic_workload(60000, 3000)
    desc = IndexDescription(Name="s", NGramSize=3)
    single = S.NewRAMBuilder(docs, desc).Build()
    ox = O.OracleIndex(3, ("$", "$"), "$", ("english", "russian", "numbers", "$")).add_packed(*docs)
    for devices in device_plans():
        sx = with_env("SG_SHARD_GATHER_COPY", "1" if gather == "copy" else "0", lambda: ShardedNGramIndex(docs, desc, devices))
        info = sx.info()
        assert info["n_shards"] == len(devices) and info["n_docs"] == 60000
        # peer reads need peer access from the first GPU to every other one (any NVSwitch box); shards of one GPU always have it
        assert info["peer_reads"] == 0 if gather == "copy" else (info["peer_reads"] == 1 or len(set(devices)) > 1)
        assert sum(sx.shard_info(s)["n_docs"] for s in range(len(devices))) == 60000
        assert [sx.shard_info(s)["id_base"] for s in range(len(devices))] == [60000 * s // len(devices) for s in range(len(devices))]
        for code, alpha, k in ((O.JACCARD, 0.5, 10), (O.COSINE, 0.45, 3), (O.DICE, 0.4, 25)):
            m = METRICS[code]
            ids, sc, cnt = sx.SuggestBatch(None, alpha, m, k, packed=(qb, qo))
            ids1, sc1, cnt1 = single.SuggestBatch(None, alpha, m, k, packed=(qb, qo))
            o_ids, o_sc, o_cnt = ox.suggest_batch(None, code, alpha, k, O.CANONICAL, threads=8, packed=(qb, qo.astype(np.uint64)))
            mask = np.arange(k)[None, :] < o_cnt[:, None]
            for a_ids, a_sc, a_cnt in ((ids, sc, cnt), (ids1, sc1, cnt1)):
                assert np.array_equal(a_cnt, o_cnt), (devices, m)
                assert np.array_equal(a_ids[mask], o_ids[mask]), (devices, m)
                assert np.array_equal(a_sc[mask], o_sc[mask]), (devices, m)
            assert (cnt > 0).mean() > 0.5
        # page-locked result rows: the merge kernel stores the valid entries (and the counts) straight into them
        pinned = S.PinnedBuffers(len(qo) - 1, 10)
        pinned.ids[...] = 0xABABABAB
        ids, sc, cnt = sx.SuggestBatch(None, 0.5, METRICS[O.JACCARD], 10, packed=(qb, qo), out=pinned.out)
        ids1, sc1, cnt1 = single.SuggestBatch(None, 0.5, METRICS[O.JACCARD], 10, packed=(qb, qo))
        mask = np.arange(10)[None, :] < cnt1[:, None]
        assert np.array_equal(cnt, cnt1) and np.array_equal(ids[mask], ids1[mask]) and np.array_equal(sc[mask], sc1[mask])
        assert np.all(ids[~mask] == 0xABABABAB)
        del ids, sc, cnt
        pinned.close()
        sx.close()
    single.close()


def test_sharded_on_cars_and_errors(cars_lines):
    desc = description(CARS_DESCRIPTION, "cars")
    n = n_gpus()
    sx = ShardedNGramIndex(cars_lines, desc, [i % n for i in range(4)])
    # pkg/suggest/service_test.go:35-59 through four shards
    words = ["Nissan March", "Honda Fitt", "Wolfsvagen", "Tayota Corolla", "Micra Nissan"]
    expected = [[b"NISSAN MARCH"], [b"HONDA FIT"], [], [b"TOYOTA COROLLA"], [b"NISSAN MICRA"]]
    for w, exp in zip(words, expected):
        assert [cars_lines[c.Key] for c in sx.Suggest(w, 0.7, S.CosineMetric(), 5)] == exp
    ids, sc, cnt = sx.SuggestBatch(["", "Nissan Марч", "ТОЙОТА"], 0.5, S.JaccardMetric(), 4)  # empty, non-ASCII (host ToLower)
    single = S.NewRAMBuilder(cars_lines, desc).Build()
    ids1, sc1, cnt1 = single.SuggestBatch(["", "Nissan Марч", "ТОЙОТА"], 0.5, S.JaccardMetric(), 4)
    assert np.array_equal(cnt, cnt1) and cnt[0] == 0
    m = np.arange(4)[None, :] < cnt1[:, None]
    assert np.array_equal(ids[m], ids1[m]) and np.array_equal(sc[m], sc1[m])
    with pytest.raises(S.SuggestError) as e:
        sx.SuggestBatch(["x"], 0.0, S.JaccardMetric(), 4)
    assert e.value.code == _capi.SG_ERR_INVALID
    long_q = "".join(chr(c) for c in np.random.default_rng(1).integers(97, 123, 300))
    with pytest.raises(S.SuggestError) as e:
        sx.SuggestBatch(["nissan", long_q], 0.5, S.JaccardMetric(), 4)
    assert e.value.code == _capi.SG_ERR_QUERY_TOO_LONG
    with pytest.raises(S.SuggestError):
        ShardedNGramIndex(cars_lines, desc, [n + 7])
    sx.close()
    single.close()
