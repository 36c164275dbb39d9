#!/usr/bin/env python
"""Timing of x264dsp_p_frames_host (pictures in pinned host memory in, coded P frames out, every copy inside) against the
number of stream groups the call splits its frames into (X264DSP_PF_HOST_GROUPS).
   python tools/bench_pframe_host.py [--frames 384] [--groups 2,4,8,16] [--me 0 --subme 1]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=384)
    ap.add_argument("--groups", default="1,2,4,8,16")
    ap.add_argument("--me", type=int, default=0)
    ap.add_argument("--subme", type=int, default=1)
    ap.add_argument("--qp", type=int, default=26)
    ap.add_argument("--part", type=int, default=0)
    ap.add_argument("--packed", type=int, default=0, help="1: x264dsp_p_frames_host_packed (compact levels, reconstruction stays on the device)")
    args = ap.parse_args()
    import __graft_entry__ as ge
    pkg = ge.load_package()
    ctx = pkg.Context(0)
    w, h, n = 1920, 1080, args.frames
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    nv = 4 if args.part else 1
    pics = ctx.pinned_empty((n + 1, w * h * 3 // 2), np.uint8)
    for i in range(n + 1):
        pics[i] = pkg.synth_frame(w, h, i % 25)
    o = {"mb_type": ctx.pinned_empty((n, nmb), np.int8), "partition": ctx.pinned_empty((n, nmb), np.uint8),
         "mv": ctx.pinned_empty((n, nmb, nv, 2), np.int16), "mvr": ctx.pinned_empty((n, nmb, 2), np.int16),
         "mvd": ctx.pinned_empty((n, nmb, nv, 2), np.int16), "levels": ctx.pinned_empty((n, nmb, pkg.RES_LEVELS_PER_MB), np.int16),
         "nnz": ctx.pinned_empty((n, nmb, pkg.RES_NNZ_PER_MB), np.uint8), "cbp": ctx.pinned_empty((n, nmb), np.int16)}
    recon = ctx.pinned_empty((n, w * h * 3 // 2), np.uint8)
    prm = pkg.PFrameParams(args.me, args.subme, 16, args.qp, 512, 1, 0, args.part)

    packed = ctx.pinned_empty((n * nmb * 392 // 4,), np.int16) if args.packed else None
    f_off = np.zeros(n + 1, np.int64)
    mb_off = ctx.pinned_empty((n, nmb), np.int32)

    def run():
        if args.packed:
            ctx.p_frames_host_packed(w, h, n, pics, prm, o["mb_type"], o["partition"] if args.part else None, o["mv"], o["mvr"], o["mvd"],
                                     packed, f_off, mb_off, o["nnz"], o["cbp"])
            return
        if args.part:
            ctx.p_frames_part_host(w, h, n, pics, prm, o["mb_type"], o["partition"], o["mv"], o["mvr"], o["mvd"], o["levels"],
                                   o["nnz"], o["cbp"], recon)
        else:
            ctx.p_frames_host(w, h, n, pics, prm, o["mb_type"], o["mv"], o["mvr"], o["mvd"], o["levels"], o["nnz"], o["cbp"], recon)
    out = []
    for gr in [int(x) for x in args.groups.split(",")]:
        os.environ["X264DSP_PF_HOST_GROUPS"] = str(gr)
        run()
        t0 = time.perf_counter()
        for _ in range(3):
            run()
        dt = (time.perf_counter() - t0) / 3
        out.append({"groups": gr, "ms_per_call": 1e3 * dt, "frames_per_s": n / dt,
                    "d2h_GBps": ((sum(a.nbytes for k, a in o.items() if k != "levels") + 2 * int(f_off[-1]) + mb_off.nbytes) if args.packed
                                 else (sum(a.nbytes for a in o.values()) + recon.nbytes)) / dt / 1e9,
                    "packed_bytes_per_frame": 2 * int(f_off[-1]) / n if args.packed else None, "h2d_GBps": pics.nbytes / dt / 1e9})
    print(json.dumps({"frames": n, "me": args.me, "subme": args.subme, "part": args.part, "packed": args.packed, "runs": out}))


if __name__ == "__main__":
    main()
