#!/usr/bin/env python3
"""NGramIndex.Autocomplete on the config #2 dictionary: 65,536 prefixes (the first 3-8 letters of random
entries), limit 10, through sg_autocomplete_batch with host buffers.  Prints ONE JSON line with the oracle on a CPU sample.
usage (GPU box): python bench.py --workload autocomplete"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import suggest_b200 as S
    from suggest_b200 import _capi
    from suggest_b200.suggest import IndexDescription, pack_strings
    from suggest_b200.workload import synthetic_dictionary
    from oracle import oracle as O
    N, NQ, K, STEPS = 1_000_000, 65536, 10, 20
    d_bytes, d_off, rng = synthetic_dictionary(N)
    desc = IndexDescription(Name="ac", NGramSize=3, Alphabet=("english", "russian", "numbers", "$"), Pad="$", Wrap=("$", "$"))
    index = S.NewRAMBuilder((d_bytes, d_off), desc).Build()
    pick = rng.integers(0, N, size=NQ)
    plen = rng.integers(3, 9, size=NQ)
    queries = [bytes(d_bytes[int(d_off[p]):int(d_off[p]) + int(l)]) for p, l in zip(pick, plen)]
    data, off = pack_strings(queries)
    buf = S.PinnedBuffers(NQ, K)  # page-locked rows: the kernel stores the completions straight into them
    for _ in range(3):
        index.AutocompleteBatch(None, K, packed=(data, off), out=buf.out)
    L = _capi.lib()
    l0 = L.sg_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        ids, scores, counts = index.AutocompleteBatch(None, K, packed=(data, off), out=buf.out)
    dt = (time.perf_counter() - t0) / STEPS
    line = {"metric": "completions/sec (Autocomplete, limit 10) on 1M-entry 3-gram index", "value": NQ / dt, "unit": "queries/s",
            "n_gpus": 1, "steps": STEPS, "warmup": 3, "ms_per_step": dt * 1e3, "higher_is_better": True, "data": "synthetic",
            "config": {"workload": "prefixes of 3-8 letters of dictionary entries, 64K-query batch, host buffers (queries pageable, result rows page-locked; H2D + kernels + rows to the host timed)",
                       "n_docs": N, "k": K},
            "gpu_launches": int(L.sg_kernel_launches() - l0), "results": {"mean_completions": float(counts.mean())}}
    ox = O.OracleIndex(3, ("$", "$"), "$", ("english", "russian", "numbers", "$")).add_packed(d_bytes, d_off)
    n = 2000
    t0 = time.perf_counter()
    same = 0
    for i in range(n):
        o_ids, _ = ox.autocomplete(queries[i], K)
        same += int(len(o_ids) == counts[i] and np.array_equal(o_ids, ids[i, :counts[i]]))
    cdt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": n / cdt, "unit": "queries/s", "cores": 1, "kind": "port", "sample": f"first {n} prefixes, oracle so_autocomplete",
                            "gpu_results_identical": same == n}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
