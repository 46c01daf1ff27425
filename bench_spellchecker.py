#!/usr/bin/env python3
"""BASELINE.json config #5: SpellChecker.Predict (completions ranked by the n-gram LM + fuzzy Cosine candidates rescored by
it) on 1xB200.  Prints ONE JSON line: predictions/s through sg_predict_batch (host buffers in and out: the call a Go shim
would make), the share of each kernel, and the CPU oracle on a bounded sample with a parity check of the ids.

Workload: vocabulary of 1,000,000 synthetic words (4-14 letters a-z, seed 12345), trigram model counted from 400,000
synthetic sentences (Zipf word choice), 65,536 queries = two context words of a sentence + the next word cut to a prefix
(half) or with one substituted letter (half); topK 5, similarity 0.5.
usage (GPU box): python bench.py --workload spellchecker [--steps K] [--queries N]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--words", type=int, default=1_000_000)
    ap.add_argument("--sentences", type=int, default=400_000)
    ap.add_argument("--queries", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=300)
    args = ap.parse_args(argv)
    import suggest_b200 as S
    from suggest_b200 import _capi
    from suggest_b200 import lm as P
    from suggest_b200.suggest import IndexDescription
    from suggest_b200.workload import synthetic_dictionary, unpack

    K, SIM = 5, 0.5
    d_bytes, d_off, rng = synthetic_dictionary(args.words, lo=4, hi=14)
    desc = IndexDescription(Name="vocab", NGramSize=3, Alphabet=("english", "$"), Pad="$", Wrap=("$", "$"))
    t0 = time.perf_counter()
    index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
    build_index_s = time.perf_counter() - t0
    n_words = args.words
    start, end = n_words, n_words + 1  # <S>, </S> get ids behind the vocabulary: they are never completion candidates
    lens = rng.integers(3, 12, size=args.sentences)
    ids = np.minimum(rng.zipf(1.15, size=int(lens.sum())) - 1, n_words - 1).astype(np.int64)
    ids = (ids * 2654435761 % n_words).astype(np.uint32)   # spread the frequent words over the id space
    bounds = np.zeros(args.sentences + 1, dtype=np.int64)
    bounds[1:] = np.cumsum(lens)
    sents = [ids[bounds[i]:bounds[i + 1]] for i in range(args.sentences)]
    t0 = time.perf_counter()
    levels = P.levels_from_sentences(sents, 3, start, end)
    model = P.NGramModel.from_levels(levels)
    build_lm_s = time.perf_counter() - t0

    # queries
    nq = args.queries
    pick = rng.integers(0, args.sentences, size=nq)
    ctxs, last = [], []
    for qi, si in enumerate(pick):
        s = sents[int(si)]
        j = int(rng.integers(2, len(s)))
        ctxs.append([int(s[j - 2]), int(s[j - 1])])
        w = bytes(d_bytes[int(d_off[s[j]]):int(d_off[s[j] + 1])])
        if qi & 1:
            w = w[:max(2, len(w) // 2)]
        else:
            p_ = int(rng.integers(len(w)))
            w = w[:p_] + bytes([97 + int(rng.integers(26))]) + w[p_ + 1:]
        last.append(w)
    from suggest_b200.suggest import pack_strings
    data, off = pack_strings(last)
    off = off.astype(np.uint32)
    ctx = np.array(ctxs, dtype=np.uint32).reshape(-1)
    ctx_off = (np.arange(nq + 1, dtype=np.uint32) * 2)
    out_ids = np.zeros((nq, K + 1), dtype=np.uint32)
    out_cnt = np.zeros(nq, dtype=np.uint32)
    L = _capi.lib()

    def step():
        _capi.check(L.sg_predict_batch(index.handle, model.handle, data.ctypes.data, off.ctypes.data, ctx.ctypes.data, ctx_off.ctypes.data,
                                       nq, SIM, K, out_ids.ctypes.data, out_cnt.ctypes.data))

    for _ in range(args.warmup):
        step()
    launches0 = L.sg_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    launches = L.sg_kernel_launches() - launches0
    line = {
        "metric": "predictions/sec (SpellChecker.Predict, topK=5, similarity 0.5, trigram LM) on a 1M-word vocabulary",
        "value": nq / dt, "unit": "predictions/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "data": "synthetic", "dtype": "u32 bitmap words / u64 packed n-grams",
        "config": {"workload": "BASELINE.json config #5: completions ranked by the LM + fuzzy Cosine candidates, merged and re-sorted by the LM",
                   "vocabulary": n_words, "sentences": args.sentences, "lm_entries": [int(len(v)) for _, v, _ in levels],
                   "queries_per_step": nq, "k": K, "similarity": SIM, "index_build_s": round(build_index_s, 2), "lm_build_s": round(build_lm_s, 2),
                   "timing": "host wall clock around sg_predict_batch (host buffers in and out, H2D + 6 kernels + D2H inside)"},
        "gpu_launches": int(launches),
        "results": {"mean_candidates": float(out_cnt.mean()), "queries_with_candidates": float((out_cnt > 0).mean())},
    }
    if args.cpu_sample:
        from oracle import lm_oracle as LM
        from oracle import oracle as O
        words = unpack(d_bytes, d_off) + [b"<S>", b"</S>"]
        ox = O.OracleIndex(3, ("$", "$"), "$", ("english", "$")).add_packed(d_bytes, d_off)
        om = LM.NGramModel([LM.PackedArray([int(x) for x in c], [int(x) for x in v], t) for c, v, t in levels])

        class _LM:  # the pieces LM.predict uses, over already mapped word ids
            model = om

            @staticmethod
            def word_id(t):
                return t

            @staticmethod
            def next(seq):
                return om.next(seq)

        n = min(args.cpu_sample, nq)
        t0 = time.perf_counter()
        same = 0
        for i in range(n):
            want = LM.predict(ox, _LM, ctxs[i] + [last[i]], K, SIM, O.COSINE, O.CANONICAL)
            same += want == [int(d) for d in out_ids[i, :int(out_cnt[i])]]
        cdt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / cdt, "unit": "predictions/s", "cores": 1, "kind": "port",
                                "sample": f"first {n} queries, oracle/lm_oracle.py predict (Python over the C suggest oracle)",
                                "gpu_results_identical": same == n}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
