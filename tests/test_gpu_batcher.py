"""sg_batcher_*: concurrent single-query Suggest calls coalesced into batches (the reference's calling pattern:
internal/suggest/api/suggest_handler.go:42-76, one goroutine per request).  Needs a B200: `pytest -m gpu`."""
import threading

import numpy as np
import pytest

from conftest import CARS_DESCRIPTION
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200 import _capi
from test_gpu_parity import build_pair

pytestmark = pytest.mark.gpu


def test_64_threads_get_the_oracle_answers(cars_lines):
    gx, ox = build_pair(CARS_DESCRIPTION, cars_lines)
    batcher = S.NewBatcher(gx, max_batch=256, max_wait_us=200, max_k=16)
    rng = np.random.default_rng(5)
    queries = []
    for i in rng.integers(0, len(cars_lines), size=64 * 40):
        w = bytearray(cars_lines[int(i)])
        if w:
            w[int(rng.integers(0, len(w)))] = int(rng.integers(65, 91))
        queries.append(bytes(w))
    metrics = [(S.CosineMetric(), O.COSINE, 0.7, 5), (S.JaccardMetric(), O.JACCARD, 0.5, 10), (S.CosineMetric(), O.COSINE, 0.7, 3)]
    want = {}
    for m, code, alpha, k in metrics:
        ids, sc, cnt = ox.suggest_batch(queries, code, alpha, k, O.CANONICAL, threads=8)
        want[(code, alpha, k)] = (ids, sc, cnt)
    errors = []

    def caller(t):
        try:
            for j in range(40):
                q = t * 40 + j
                m, code, alpha, k = metrics[(t + j) % len(metrics)]  # mixed (metric, similarity, k) in flight at once
                got = batcher.Suggest(queries[q], alpha, m, k)
                ids, sc, cnt = want[(code, alpha, k)]
                exp = [(int(ids[q, i]), float(sc[q, i])) for i in range(int(cnt[q]))]
                if [(c.Key, c.Score) for c in got] != exp:
                    errors.append((q, queries[q], got, exp))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(64)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    st = batcher.stats()
    assert st["queries"] == 64 * 40
    assert st["batches"] < st["queries"], st   # calls were coalesced
    assert 1 < st["largest_batch"] <= 256, st
    batcher.close()
    gx.close()


def test_errors_are_per_query_and_validation_matches_search_config(cars_lines):
    gx, _ = build_pair(CARS_DESCRIPTION, cars_lines)
    batcher = S.NewBatcher(gx, max_batch=8, max_wait_us=50, max_k=4)
    assert [cars_lines[c.Key] for c in batcher.Suggest("Nissan March", 0.7, S.CosineMetric(), 4)][:1] == [b"NISSAN MARCH"]
    assert batcher.Suggest("", 0.7, S.CosineMetric(), 4) == []
    with pytest.raises(S.SuggestError):
        batcher.Suggest("x", 0.7, S.CosineMetric(), 5)      # above max_k
    with pytest.raises(S.SuggestError):
        batcher.Suggest("x", 1.5, S.CosineMetric(), 2)      # pkg/suggest/search.go:23-25
    with pytest.raises(S.SuggestError):
        batcher.Suggest("x", 0.5, S.CosineMetric(), 0)      # pkg/suggest/search.go:18-21
    assert batcher.stats()["max_batch"] == 8
    batcher.close()
    gx.close()
