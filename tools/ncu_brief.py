#!/usr/bin/env python
"""One-screen digest of an ncu report (read on the CPU box): duration, instruction count, issue rate, pipes, stall mix.
    python tools/ncu_brief.py gpurun_out/x.ncu-rep [launch_index]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2 + idx]))
    u = dict(zip(hdr, units))
    print(d.get("Kernel Name"))
    for k in WANT:
        if k in d:
            print(f"  {k:78s} {d[k]} {u[k]}")
    stalls = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
    for v, k in stalls[:9]:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.2f}")


if __name__ == "__main__":
    main()
