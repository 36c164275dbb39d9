#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode of one kernel from an ncu report:
   python tools/ncu_opcodes.py rep.ncu-rep"""
import csv, io, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
ops = collections.Counter()
for r in rows:
    if r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        d = dict(zip(hdr, r))
        toks = d["Source"].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "SHFL", "ISETP", "VABS")) and "." in op else "")] += int(d["Instructions Executed"])
tot = sum(ops.values())
print("total", tot)
for k, v in ops.most_common(40):
    print(f"{100*v/tot:5.1f}%  {v:>12d}  {k}")
