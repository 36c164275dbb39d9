#!/usr/bin/env python
"""Timing of x264dsp_p_frames_dev (the P-slice macroblock loop on the device): N independent 1080p P frames per launch,
CUDA events on the context's stream.   python tools/bench_pframe.py [--frames 8,32,96] [--me 1 --subme 5 --qp 26]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", default="1,8,32,96")
    ap.add_argument("--me", type=int, default=1)
    ap.add_argument("--subme", type=int, default=5)
    ap.add_argument("--qp", type=int, default=26)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--part", type=int, default=-1, help="-1: x264dsp_p_frames_dev; 0 / 1: x264dsp_p_frames_part_dev with analyse_inter = 0 / PSUB16x16")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    ctx = pkg.Context(0)
    stream = ctx.torch_stream()
    w, h = args.width, args.height
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    counts = [int(x) for x in args.frames.split(",")]
    nmax = max(counts)
    distinct = min(nmax, 24) + 1
    frames = np.stack([pkg.synth_frame(w, h, i) for i in range(distinct)])
    src = torch.zeros((nmax + 1) * g.slot_bytes, dtype=torch.uint8, device="cuda")
    one = torch.zeros(distinct * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(frames).cuda(), one, distinct)
    ctx.frame_expand_border(g, one, distinct)
    ctx.frame_filter(g, one, distinct)
    ctx.frame_init_lowres(g, one, distinct)
    for k in range(nmax + 1):           # pair k = (slot k -> slot k+1); slots cycle through the distinct frames
        j = k % distinct
        src[k * g.slot_bytes:(k + 1) * g.slot_bytes] = one[j * g.slot_bytes:(j + 1) * g.slot_bytes]
    torch.cuda.synchronize()            # torch's copies run on its own stream; the context's streams do not wait for it
    b = np.arange(1, nmax + 1, dtype=np.int32)
    d_lmv = torch.zeros((nmax, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((nmax, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((nmax, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, src, b, b - 1, np.ones(nmax, np.uint8), d_lmv, d_lc, d_ls)
    ctx.sync()
    out = {"config": {"width": w, "height": h, "me": args.me, "subme": args.subme, "qp": args.qp, "part": args.part}, "runs": []}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for n in counts:
        o = dict(mb_type=torch.zeros((n, nmb), dtype=torch.int8, device="cuda"),
                 mv=torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda"),
                 mvr=torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda"),
                 levels=torch.zeros((n, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
                 nnz=torch.zeros((n, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
                 cbp=torch.zeros((n, nmb), dtype=torch.int16, device="cuda"))
        recon = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
        prm = pkg.PFrameParams(args.me, args.subme, 16, args.qp, 512, 1, 0, max(args.part, 0))
        part = torch.zeros((n, nmb), dtype=torch.uint8, device="cuda")
        mv8 = torch.zeros((n, nmb, 4, 2), dtype=torch.int16, device="cuda")

        def run():
            if args.part >= 0:
                ctx.p_frames_part(g, src[g.slot_bytes:], src, recon, n, prm, d_lmv[:n], None, o["mb_type"], part, mv8, o["mvr"],
                                  o["levels"], o["nnz"], o["cbp"])
                return
            ctx.p_frames(g, src[g.slot_bytes:], src, recon, n, prm, d_lmv[:n], None, o["mb_type"], o["mv"], o["mvr"],
                         o["levels"], o["nnz"], o["cbp"])
        run()
        torch.cuda.synchronize()
        ev[0].record(stream)
        for _ in range(args.reps):
            run()
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / args.reps
        t = o["mb_type"].cpu().numpy()
        out["runs"].append({"frames_per_launch": n, "ms_per_launch": ms, "ms_per_frame": ms / n, "frames_per_s": 1e3 * n / ms,
                            "skipped_mb_share": float((t == pkg.MB_P_SKIP).mean()),
                            "partition_share": {str(k): float(((part.cpu().numpy() == k) & (t != pkg.MB_P_SKIP)).mean())
                                                for k in (13, 14, 15, 16)} if args.part >= 0 else None})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
