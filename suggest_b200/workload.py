"""Synthetic dictionaries and query batches of BASELINE.json's configs.

Dictionary: n_docs strings, length uniform in [lo, hi], characters i.i.d. uniform a-z; lengths are
drawn first, then all characters (numpy default_rng(seed)).  Queries: a uniformly chosen entry with
`subs` single-character substitutions (uniform position, uniform a-z), the same stream continued.
Everything is returned packed: (uint8 bytes, offsets[n+1]).
"""
import numpy as np


def synthetic_dictionary(n_docs, seed=12345, lo=8, hi=32, rng=None, skew=None):
    """skew=None: uniform letters (BASELINE.json config #2).  skew="zipf": letter i drawn with probability ~ 1/(i+1), the
    labelled real-language-like variant of SURVEY.md 8(d) (posting lists about 10x longer for the frequent n-grams)."""
    rng = rng if rng is not None else np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n_docs)
    off = np.zeros(n_docs + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    if skew == "zipf":
        p = 1.0 / np.arange(1, 27)
        data = rng.choice(26, size=int(off[-1]), p=p / p.sum()).astype(np.uint8) + np.uint8(97)
    else:
        data = (rng.integers(0, 26, size=int(off[-1]), dtype=np.uint8) + np.uint8(97))
    return data, off, rng


def synthetic_queries(data, off, n_queries, rng, subs=2):
    n_docs = len(off) - 1
    pick = rng.integers(0, n_docs, size=n_queries)
    start = off[pick].astype(np.int64)
    lens = (off[pick + 1] - off[pick]).astype(np.int64)
    q_off = np.zeros(n_queries + 1, dtype=np.uint32)
    q_off[1:] = np.cumsum(lens)
    idx = np.repeat(start - q_off[:-1].astype(np.int64), lens) + np.arange(int(q_off[-1]), dtype=np.int64)
    q = data[idx].copy()
    for _ in range(subs):
        pos = (rng.random(n_queries) * lens).astype(np.int64)
        q[q_off[:-1].astype(np.int64) + pos] = rng.integers(0, 26, size=n_queries, dtype=np.uint8) + np.uint8(97)
    return q, q_off, pick


def synthetic_workload(n_docs, n_queries, seed=12345, lo=8, hi=32, subs=2):
    """-> ((doc bytes, doc offsets uint64), (query bytes, query offsets uint32), source entry of every query)"""
    data, off, rng = synthetic_dictionary(n_docs, seed, lo, hi)
    q, q_off, pick = synthetic_queries(data, off, n_queries, rng, subs)
    return (data, off), (q, q_off), pick


def unpack(data, off):
    return [data[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
