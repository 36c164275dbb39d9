#!/usr/bin/env python
"""Top stall locations of an ncu report's source page (SASS view):  python tools/ncu_hot.py x.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
texec = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions executed {texec}")
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))
for i in order[:top]:
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100.0 * s / tot:5.1f}%  exec {int(r[ix['Instructions Executed']] or 0):9d}  {r[ix['Source']].strip():60s} {st}")
