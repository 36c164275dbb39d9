#!/usr/bin/env python
"""Summarise an ncu report (read here, on the CPU box) into the small text files kept under profiles/.

  python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x          -> .md + .json
  python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r01_launches.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ,
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kidx = hdr.index("Kernel Name")
    summary = OrderedDict()
    md = [f"# ncu --set full summary of `{rep}`", "",
          "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (see the command in the",
          "round's profile notes); values are per launch.  Times under the profiler are not bench values.", ""]
    for n, r in enumerate(rows[2:]):
        name = r[kidx].split("(")[0]
        md.append(f"## launch {n}: `{name}`")
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        rec = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append(f"| {k} | {r[i]} | {units[i]} |")
                rec[k] = (r[i], units[i])
        md.append("")
        rd = to_bytes(*rec["dram__bytes_read.sum"]) if "dram__bytes_read.sum" in rec else 0
        wr = to_bytes(*rec["dram__bytes_write.sum"]) if "dram__bytes_write.sum" in rec else 0
        summary.setdefault(name, []).append({"dram_bytes": rd + wr, "time": rec.get("gpu__time_duration.sum")})
    traffic = {f"{k}_bytes_per_launch": sum(x["dram_bytes"] for x in v) / len(v) for k, v in summary.items()}
    open(out + ".md", "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(out + ".json", "w"), indent=1)
    print("\n".join(md))


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    k, v, g, b = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[k].split("(")[0]
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "grid": r[g], "block": r[b]})
        a["n"] += 1
        a["ns"] += float(r[v].replace(",", ""))
    total = sum(a["ns"] for a in agg.values())
    md = [f"# launch list summary of `{csv_path}`", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised launches -- compare SHARES.", "",
          "| kernel | launches | grid | block | total us | avg us | share |", "|---|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        md.append(f"| `{name}` | {a['n']} | {a['grid']} | {a['block']} | {a['ns'] / 1e3:.1f} | {a['ns'] / a['n'] / 1e3:.1f} | {100 * a['ns'] / total:.1f}% |")
    open(out, "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2])
