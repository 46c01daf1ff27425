#!/usr/bin/env python3
"""bench.py — queries/sec of the Suggest hot path on BASELINE.json config #2.

    python bench.py --gpus N --steps K --warmup W            # this repo, CUDA path through the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload: 1,000,000 synthetic 8-32 character a-z strings, 3-gram index, Jaccard >= 0.5,
k = 10, 65,536 queries per step (dictionary entries with two substituted characters).  A step is one
pass of the whole path (tokenise -> posting fetch -> T-occurrence count -> score -> top-k) over one batch.
With N > 1 every rank holds a replica of the 1M index and searches its own batch (independent queries,
no data-path collective; weak scaling).  `--workload sharded` runs config #4 instead: the dictionary is
split by record-id range, every rank searches the same batch, per-shard top-k is all-gathered (NCCL) and
merged on the device.

`--workload spellchecker` / `--workload autocomplete` run the SURVEY.md 8(f) rows (bench_spellchecker.py, bench_autocomplete.py).

Prints ONE JSON line (rank 0).  `value` = queries/s with the batch resident in HBM, CUDA-event timed;
`e2e` = the same through sg_search_batch with pinned host buffers (H2D + kernel + D2H per step);
`roofline` = algorithmic bytes (SURVEY.md 8(d) formula, counted by the kernel's own stats pass) over
the kernel's event-timed duration against MEASURED_PEAKS.json; `cpu_baseline` = the oracle's
line-faithful CPMerge path on the host cores over a bounded sample, which also checks the GPU results.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DOCS = 1_000_000
N_QUERIES = 65536
K = 10
ALPHA = 0.5
DESCRIPTION = dict(ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("english", "russian", "numbers", "$"))
METRIC_NAME = "queries/sec (k=10, Jaccard>=0.5) on 1M-entry 3-gram index"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"{kernel}_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """SM clock and throttle reasons during the timed regions: an NVML polling thread (a few ms per sample; nvidia-smi's
    own loop is too coarse for millisecond steps), nvidia-smi as the fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag = threading.Event()
        self.thread = self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.thread = None
            self._start_smi()

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def _start_smi(self):
        self.path = f"/tmp/sg_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                out = {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                       "samples": len(self.samples), "how": "nvml polling thread"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                   "how": "nvidia-smi -lms 20"}
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def oracle_index(docs):
    from oracle import oracle as O
    ox = O.OracleIndex(DESCRIPTION["ngram_size"], DESCRIPTION["wrap"], DESCRIPTION["pad"], DESCRIPTION["alphabet"])
    ox.add_packed(docs[0], docs[1])
    ox.commit()  # VB / skipping(64) bytes, decoded lazily per Next() like the reference
    return ox


def oracle_run(ox, q_bytes, q_off, lo, hi, threads):
    """line-faithful reference path (CPMerge, lazy codecs, per-segment queues, dynamic alpha) on queries [lo, hi)"""
    from oracle import oracle as O
    off = q_off[lo:hi + 1].astype(np.uint64)
    data = q_bytes[int(off[0]):int(off[-1])]
    off = off - off[0]
    t0 = time.perf_counter()
    res = ox.suggest_batch(None, O.JACCARD, ALPHA, K, O.FAITHFUL, O.CP_MERGE, threads=threads, packed=(data, off))
    return time.perf_counter() - t0, res


def run_reference(args, out_stream):
    """--impl reference: the reference's own CPU implementation of the path.  No Go toolchain exists in this image,
    so this is the oracle's line-faithful port (oracle/so_suggest.c FAITHFUL mode), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from suggest_b200.workload import synthetic_workload
    docs, (q_bytes, q_off), _ = synthetic_workload(N_DOCS, N_QUERIES)
    ox = oracle_index(docs)
    threads = host_threads()
    sample = 8192
    for w in range(args.warmup):
        oracle_run(ox, q_bytes, q_off, 0, min(sample, 2048), threads)
    total_t, total_q = 0.0, 0
    for s in range(args.steps):
        lo = (s * sample) % N_QUERIES
        hi = min(lo + sample, N_QUERIES)
        dt, _ = oracle_run(ox, q_bytes, q_off, lo, hi, threads)
        total_t += dt
        total_q += hi - lo
    qps = total_q / total_t
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "1M synthetic 8-32-char a-z strings, 3-gram, Jaccard 0.5, k=10", "n_docs": N_DOCS,
                   "queries_per_step": sample, "note": "each step is a bounded 8192-query sample of the 65536-query batch"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {sample} queries, oracle FAITHFUL mode (CPMerge, lazy VB/skipping decode, "
                                   "bounded heap), one query per thread"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out_stream, flush=True)
    return 0


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to stdout) must not add to
    it: keep the real stdout for the result and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out_stream = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="replicated", choices=["replicated", "sharded", "spellchecker", "autocomplete"])
    ap.add_argument("--docs", type=int, default=None, help="dictionary size (default 1M; sharded: 10M)")
    ap.add_argument("--metric", default="Jaccard", choices=["Jaccard", "Cosine", "Dice"])
    ap.add_argument("--ngram", type=int, default=3)
    ap.add_argument("--data", default="uniform", choices=["uniform", "zipf"], help="letter distribution of the dictionary")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload in ("spellchecker", "autocomplete"):
        # the "next" rows of SURVEY.md 8(f): BASELINE.json config #5 (bench_spellchecker.py) and Autocomplete
        # (bench_autocomplete.py); single GPU, their own JSON line with a cpu_baseline from the oracle
        sys.stdout = out_stream
        if args.workload == "spellchecker":
            import bench_spellchecker
            return bench_spellchecker.main(["--steps", str(min(args.steps, 20)), "--warmup", str(args.warmup)]) or 0
        import bench_autocomplete
        return bench_autocomplete.main() or 0
    if args.impl == "reference":
        return run_reference(args, out_stream)

    import torch
    import torch.distributed as dist
    import suggest_b200 as S
    from suggest_b200 import _capi
    from suggest_b200.suggest import IndexDescription
    from suggest_b200.workload import synthetic_dictionary, synthetic_queries

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Suggest path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sharded = args.workload == "sharded"
    n_docs = args.docs or (10_000_000 if sharded else N_DOCS)
    metric = {"Jaccard": S.JaccardMetric(), "Cosine": S.CosineMetric(), "Dice": S.DiceMetric()}[args.metric]
    desc = dict(DESCRIPTION, ngram_size=args.ngram)

    # ---- data: identical dictionary on every rank; replicated mode gives every rank its own query batch ----
    d_bytes, d_off, rng = synthetic_dictionary(n_docs, skew=None if args.data == "uniform" else args.data)
    if not sharded and rank > 0:
        rng = np.random.default_rng(12345 + 7919 * rank)
    q_bytes, q_off, pick = synthetic_queries(d_bytes, d_off, N_QUERIES, rng)
    description = IndexDescription(Name="bench", NGramSize=desc["ngram_size"], Alphabet=desc["alphabet"], Pad=desc["pad"],
                                   Wrap=desc["wrap"], Device=local)
    t0 = time.perf_counter()
    sharded_index = None
    if sharded:
        from suggest_b200.sharding import ShardedIndex
        sharded_index = ShardedIndex((d_bytes, d_off), description, rank, world, S.NewRAMBuilder)
        index = sharded_index.index
    else:
        index = S.NewRAMBuilder((d_bytes, d_off), description).Build()
    build_s = time.perf_counter() - t0
    info = index.info()
    L = _capi.lib()

    # ---- device-resident buffers for `value`, pinned host buffers for `e2e` ----
    nq = N_QUERIES
    dq = torch.from_numpy(q_bytes).to(dev)
    doff = torch.from_numpy(q_off.astype(np.int32)).to(dev)
    d_ids = torch.zeros(nq * K, dtype=torch.int32, device=dev)
    d_sc = torch.zeros(nq * K, dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_stats = torch.zeros(nq * 2, dtype=torch.int32, device=dev)
    g_ids = None
    m_ids = m_sc = m_cnt = None
    if sharded and world > 1:
        g_ids = True  # per-shard rows are all-gathered and merged (suggest_b200/sharding.py)
        m_ids, m_sc, m_cnt = torch.zeros_like(d_ids), torch.zeros_like(d_sc), torch.zeros_like(d_cnt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device(stats_ptr=0):
        if g_ids is not None and not stats_ptr:  # config #4: local search, all-gather of per-shard top-k, merge on the device
            sharded_index.SuggestBatchDevice(dq, doff, nq, ALPHA, metric, K, m_ids, m_sc, m_cnt)
            return
        index.SuggestBatchDevice(dq.data_ptr(), doff.data_ptr(), nq, ALPHA, metric, K, d_ids.data_ptr(), d_sc.data_ptr(),
                                 d_cnt.data_ptr(), stats_ptr, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # algorithmic bytes of one launch, counted by the kernel's stats pass (untimed)
    step_device(d_stats.data_ptr())
    torch.cuda.synchronize()
    st = d_stats.cpu().numpy().astype(np.int64).reshape(-1, 2)
    alg_bytes = int(4 * st[:, 0].sum() + 8 * st[:, 1].sum() + int(q_off[-1]) + 12 * K * nq)
    first_counts = d_cnt.cpu().numpy().copy()
    first_ids = d_ids.cpu().numpy().copy()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None  # runs through both timed regions (device-resident and e2e)
    launches0 = L.sg_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)  # L2 flush between timed steps (untimed)
        ev[s][0].record()
        step_device()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = L.sg_kernel_launches() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    queries_per_step = nq if sharded else nq * world
    value = queries_per_step / (ms_per_step * 1e-3)

    # the dominant kernel alone (roofline): CUDA events between the kernels of a launch, on the launching stream,
    # L2 flushed before every launch like the timed steps above
    stage_ms = {}
    n_stage_runs = min(args.steps, 20)
    for s in range(n_stage_runs):
        flush.fill_(s & 0xFF)
        for name, ms in index.StageTimes(dq.data_ptr(), doff.data_ptr(), nq, ALPHA, metric, K, d_ids.data_ptr(), d_sc.data_ptr(),
                                         d_cnt.data_ptr(), stream.cuda_stream).items():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms / n_stage_runs
    top_kernel = max(stage_ms, key=stage_ms.get)
    kernel_ms = stage_ms[top_kernel]
    layout = index.layout()

    # ---- e2e: sg_search_batch, pinned host buffers, H2D + kernel + D2H inside the timed region ----
    hq = torch.from_numpy(q_bytes).pin_memory()
    hoff = torch.from_numpy(q_off.astype(np.int32)).pin_memory()
    h_ids = torch.zeros(nq * K, dtype=torch.int32).pin_memory()
    h_sc = torch.zeros(nq * K, dtype=torch.float64).pin_memory()
    h_cnt = torch.zeros(nq, dtype=torch.int32).pin_memory()
    out = (h_ids.numpy().view(np.uint32).reshape(nq, K), h_sc.numpy().reshape(nq, K), h_cnt.numpy().view(np.uint32))
    packed = (hq.numpy(), hoff.numpy().view(np.uint32))

    def step_e2e():
        if g_ids is not None:  # queries from pinned host memory, merged rows back to the host
            dq.copy_(hq, non_blocking=True)
            doff.copy_(hoff, non_blocking=True)
            sharded_index.SuggestBatchDevice(dq, doff, nq, ALPHA, metric, K, m_ids, m_sc, m_cnt)
            h_ids.copy_(m_ids, non_blocking=True)
            h_sc.copy_(m_sc, non_blocking=True)
            h_cnt.copy_(m_cnt, non_blocking=True)
            torch.cuda.synchronize()
            return
        index.SuggestBatch(None, ALPHA, metric, K, packed=packed, out=out)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = queries_per_step * args.steps / e2e_s
    clocks = sampler.stop() if sampler else None
    if g_ids is None:
        assert np.array_equal(out[2].astype(np.int32), first_counts) and np.array_equal(h_ids.numpy(), first_ids), \
            "host-buffer and device-buffer paths disagree"
    h2d = int(q_bytes.nbytes + 4 * (nq + 1))
    d2h = int(nq * K * 4 + nq * K * 8 + nq * 4)
    # page-locked result buffers: sg_search_batch lets the kernel store the valid entries of every row (12 B each) and
    # the counts straight into host memory; nothing is staged in HBM and no copy follows the kernel
    direct = (g_ids is None and layout["engine"] == 1 and os.environ.get("SG_DIRECT_OUT", "1") != "0"
              and all(L.sg_is_pinned(t.data_ptr(), t.numel() * t.element_size()) == 1 for t in (h_ids, h_sc, h_cnt)))
    if direct:
        d2h = int(nq * 4 + 12 * int(first_counts.astype(np.int64).clip(0, K).sum()))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_kind = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "u32 bitmap words / f64 scores" if layout["engine"] == 1 else "u32 postings / u8 counters / f64 scores",
        "data": "synthetic",
        "config": {"workload": (f"{n_docs}-entry dictionary sharded by record-id range over {world} GPU(s), per-shard top-k + NCCL "
                                "all-gather + merge" if sharded else "1M synthetic 8-32-char a-z strings, 3-gram, Jaccard 0.5, k=10, 64K-query batch"),
                   "n_docs": n_docs, "queries_per_step_per_gpu": nq, "k": K, "similarity": ALPHA, "metric": args.metric,
                   "ngram": args.ngram, "letters": args.data, "parallelism": ("record-id-range shards" if sharded else "replicated index, queries split"),
                   "postings": int(info["n_postings"]), "index_bytes": int(info["device_bytes"]), "index_build_s": round(build_s, 2),
                   "engine": "bitmap" if layout["engine"] == 1 else "scancount", "bucket_shift": int(layout["bucket_shift"]),
                   "bitmap_bytes": int(layout["bitmap_bytes"]),
                   "l2": "flushed between timed steps (256 MiB write, untimed); inside a step the index is re-read ~58x "
                         "and stays L2-resident, which is the steady state of this workload"},
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "result_path": ("kernel stores into the caller's page-locked rows (valid entries + counts only)" if direct
                                else "rows staged in HBM, cudaMemcpyAsync per slice")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(top_kernel), "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg_bytes,
                     "algorithmic_bytes_per_query": alg_bytes / nq, "kernel": top_kernel, "kernel_ms": kernel_ms,
                     "stage_ms": {k_: round(v, 5) for k_, v in stage_ms.items()},
                     "note": ("algorithmic bytes = posting-list bytes of SURVEY.md 8(d) (implementation independent); the bitmap "
                              "engine answers the same queries from per-term bucket bitmaps (~4-5x fewer bytes, L2 resident), "
                              "so frac can exceed 1: it measures the path against a posting-list scan at HBM speed"
                              if layout["engine"] == 1 else "posting lists are read once per query by sg_search_kernel")},
        "clocks": clocks, "wall_s_timed_region": wall,
        "results": {"queries_with_a_match": float((first_counts > 0).mean())},
    }

    if world == 1 and not args.no_cpu_baseline and not sharded and args.metric == "Jaccard" and args.ngram == 3 and args.data == "uniform":
        ox = oracle_index((d_bytes, d_off))
        threads = host_threads()
        sample = N_QUERIES  # the whole batch: about a second on 16 cores, and every GPU result of the step is checked
        dt, (o_ids, o_sc, o_cnt) = oracle_run(ox, q_bytes, q_off, 0, sample, threads)
        ok = bool(np.array_equal(o_cnt, first_counts[:sample].astype(np.uint32)))
        m = np.arange(K)[None, :] < o_cnt[:, None]
        ok = ok and bool(np.array_equal(o_ids[m], first_ids.view(np.uint32).reshape(nq, K)[:sample][m]))
        ok = ok and bool(np.array_equal(o_sc[m], out[1][:sample][m]))
        line["cpu_baseline"] = {"value": sample / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"all {sample} queries of the batch, oracle FAITHFUL mode (CPMerge, lazy VB/skipping "
                                          "decode, bounded heap), one query per thread", "gpu_results_identical": ok}
        if not ok:
            line["parity_error"] = "GPU results differ from the oracle on the cpu_baseline sample"
    print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
