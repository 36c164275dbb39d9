#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report
(needs -lineinfo and --import-source on):  python tools/ncu_source_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = ""
lines = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        lines.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"]), int(d["# Samples"]),
                      int(d["stall_long_sb"]), int(d["stall_sleep"]) if "stall_sleep" in d else 0))
tot_i = sum(l[3] for l in lines) or 1
tot_s = sum(l[4] for l in lines) or 1
print(f"total instructions {tot_i}, samples {tot_s}")
print("by instructions:")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*l[3]/tot_i:5.1f}% inst {100*l[4]/tot_s:5.1f}% smp  {l[0]}:{l[1]:<4d} {l[2][:100]}")
print("by stall samples:")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{100*l[4]/tot_s:5.1f}% smp {100*l[3]/tot_i:5.1f}% inst (long_sb {l[5]}, sleep {l[6]})  {l[0]}:{l[1]:<4d} {l[2][:100]}")
