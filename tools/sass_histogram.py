#!/usr/bin/env python
"""Opcode histogram of every kernel in libx264dsp_b200.so from `cuobjdump -sass` (static counts: what the compiler
emitted, not what a launch executes -- ncu's smsp__inst_executed is the dynamic side).

    python tools/sass_histogram.py [out.md]        default profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "x264-dsp_b200", "libx264dsp_b200.so")
# the instructions the design leans on and the memory operations by width
FOCUS = ["VABSDIFF4", "VABSDIFF", "IDP.4A", "IDP.2A", "VIADD.16x2", "VIMNMX.U16x2", "VIMNMX.S16x2", "VIMNMX", "VIADDMNMX", "I2IP",
         "PRMT", "SHF", "LOP3", "IADD3", "IMAD", "LEA", "ISETP", "SEL", "SHFL", "VOTE", "REDUX", "ATOMG", "RED",
         "LDG.E.128", "LDG.E.64", "LDG.E", "LDG.E.U8", "LDG.E.U16", "LDGSTS", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS",
         "UTMALDG", "UBLKCP", "HMMA", "IMMA", "UTCIMMA", "BAR", "WARPSYNC", "MEMBAR", "ERRBAR", "NANOSLEEP"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.md")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    with open(out, "w") as f:
        f.write("# SASS opcode histogram per kernel (`cuobjdump -sass x264-dsp_b200/libx264dsp_b200.so`, sm_100a)\n\n")
        f.write("Static instruction counts per kernel; columns are prefix matches on the full mnemonic (so `LDG.E.128` counts\n"
                "`LDG.E.128.CONSTANT`, `LDG.E.128.STRONG.GPU`, ...; plain `LDG.E` counts only the 32-bit loads).  `UTMALDG`\n"
                "(TMA) appears in one kernel only, `xd_hpel_tma_kernel`, a measured variant that is off by default; no `*MMA` (tensor\n"
                "cores) anywhere: the path is byte-integer work on the ALU / FMA pipes, staged through registers and shared memory\n"
                "(DESIGN.md section 3 says why, with the measurements).\n\n")
        for name, c in kernels.items():
            total = sum(c.values())
            f.write(f"## `{demangle(name)}`\n\n{total} instructions.  ")
            parts = []
            for key in FOCUS:
                if key == "LDG.E":
                    n = sum(v for k, v in c.items() if re.match(r"LDG\.E(\.(CONSTANT|STRONG|SYS|GPU|CTA|EL|EF|LTC\w+))*$", k))
                elif key == "STG.E":
                    n = sum(v for k, v in c.items() if re.match(r"STG\.E(\.(STRONG|SYS|GPU|CTA|EF|EL))*$", k))
                elif key in ("VABSDIFF", "VIMNMX"):
                    n = sum(v for k, v in c.items() if k == key or (k.startswith(key + ".") and not k.startswith(key + "4")
                                                                      and "16x2" not in k))
                else:
                    n = sum(v for k, v in c.items() if k == key or k.startswith(key + "."))
                if n:
                    parts.append(f"{key} {n}")
            f.write(", ".join(parts) + "\n\n")
            top = ", ".join(f"{k} {v}" for k, v in c.most_common(12))
            f.write(f"Most frequent: {top}\n\n")
    print(f"{len(kernels)} kernels -> {out}")


if __name__ == "__main__":
    main()
