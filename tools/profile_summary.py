#!/usr/bin/env python3
"""ncu report -> the text summary committed under profiles/ (run here, no GPU needed).

usage: profile_summary.py <prof.ncu-rep> <out.md> [title] [kernel name] [workload of the launch]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_lines  # noqa: E402

KEYS = [
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
    kernel = sys.argv[4] if len(sys.argv) > 4 else "sg_bitmap_search_kernel"
    workload = sys.argv[5] if len(sys.argv) > 5 else "65,536 queries of BASELINE.json config #2"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on; one launch of "
             f"{kernel} = {workload})", "", "| metric | unit | value |", "|---|---|---|"]
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals[k] = rows[2][i]
            lines.append(f"| {k} | {units[i]} | {rows[2][i]} |")
    src_csv = rep.replace(".ncu-rep", "_source.csv")
    with open(src_csv, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                               capture_output=True, text=True).stdout)
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        import ncu_top
        ncu_top.main(src_csv, 65536, 32)
    finally:
        sys.stdout = old
    lines += ["", "## instructions and stall samples per phase / per source line", "", "```", buf.getvalue().rstrip(), "```"]
    try:
        dram = float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])
        unit = units[hdr.index("dram__bytes_read.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        lines += ["", f"DRAM traffic per launch: {dram * scale / 1e6:.1f} MB (read + write)"]
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:  # noqa: BLE001
            traffic = {}
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(dict(traffic, **{f"{kernel}_dram_bytes_per_launch": dram * scale, f"{kernel}_source": os.path.basename(out)}), f, indent=1)
    except Exception as e:  # noqa: BLE001
        lines.append(f"(no dram figures: {e})")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
