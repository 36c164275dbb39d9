#!/usr/bin/env python
"""Closed GOPs on the device: x264dsp_gops_encode_dev (device-resident, CUDA events) and x264dsp_gops_encode_host (pictures in
pinned host memory in, the entropy coder's input out, wall clock).  1080p, I + P frames, in-loop filter on.
   python tools/bench_gops.py [--gops 96 --len 8 --me 0 --subme 1 --psub 0]"""
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
    ap.add_argument("--gops", type=int, default=96)
    ap.add_argument("--len", type=int, default=8)
    ap.add_argument("--me", type=int, default=0)
    ap.add_argument("--subme", type=int, default=1)
    ap.add_argument("--psub", type=int, default=0)
    ap.add_argument("--qp", type=int, default=26)
    ap.add_argument("--host", type=int, default=1)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    ctx = pkg.Context(0)
    stream = ctx.torch_stream()
    w, h, G, L = 1920, 1080, args.gops, args.len
    g = pkg.geometry(w, h)
    nmb, n = g.mb_count, args.gops * args.len
    prm = pkg.GopEncodeParams(args.me, args.subme, 16, args.qp - 3, args.qp, 512, 1, args.psub, 1, 0, 0)
    distinct = [pkg.synth_frame(w, h, i) for i in range(25)]
    res = {"config": {"gops": G, "gop_len": L, "me": args.me, "subme": args.subme, "psub": args.psub, "qp": args.qp}}
    # ---- device-resident
    fenc = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    one = torch.zeros(25 * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(np.stack(distinct)).cuda(), one, 25)
    ctx.frame_expand_border(g, one, 25)
    ctx.frame_init_lowres(g, one, 25)
    for t in range(L):
        for gop in range(G):
            k = t * G + gop
            j = (3 * gop + t) % 25
            fenc[k * g.slot_bytes:(k + 1) * g.slot_bytes] = one[j * g.slot_bytes:(j + 1) * g.slot_bytes]
    torch.cuda.synchronize()            # torch's copies run on its own stream; the context's streams do not wait for it
    del one
    b = np.arange(G, n, dtype=np.int32)
    d_lmv = torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    if L > 1:
        ctx.lookahead_frame_cost(g, fenc, b, b - G, np.zeros(b.size, np.uint8), d_lmv[G:], d_lc[G:], d_ls[G:])
    dev = {"mb_type": torch.zeros((n, nmb), dtype=torch.int8, device="cuda"), "partition": torch.zeros((n, nmb), dtype=torch.uint8, device="cuda"),
           "mv8": torch.zeros((n, nmb, 4, 2), dtype=torch.int16, device="cuda"), "mvr": torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda"),
           "mvd8": torch.zeros((n, nmb, 4, 2), dtype=torch.int16, device="cuda"),
           "levels": torch.zeros((n, nmb, 392), dtype=torch.int16, device="cuda"), "nnz": torch.zeros((n, nmb, 27), dtype=torch.uint8, device="cuda"),
           "cbp": torch.zeros((n, nmb), dtype=torch.int16, device="cuda"), "mode16": torch.zeros((G, nmb), dtype=torch.uint8, device="cuda"),
           "chroma_mode": torch.zeros((G, nmb), dtype=torch.uint8, device="cuda"), "modes4": torch.zeros((G, nmb, 16), dtype=torch.uint8, device="cuda"),
           "luma_dc": torch.zeros((G, nmb, 16), dtype=torch.int16, device="cuda")}
    recon = torch.zeros_like(fenc)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ctx.gops_encode(g, fenc, recon, G, L, prm, d_lmv, dev)
    torch.cuda.synchronize()
    ev[0].record(stream)
    for _ in range(2):
        ctx.gops_encode(g, fenc, recon, G, L, prm, d_lmv, dev)
    ev[1].record(stream)
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 2
    t = dev["mb_type"].cpu().numpy()
    res["device"] = {"ms_per_call": ms, "frames_per_s": 1e3 * n / ms, "us_per_frame": 1e3 * ms / n,
                     "skipped_mb_share_of_p_frames": float((t[G:] == 6).mean()) if L > 1 else None}
    del fenc, recon, dev
    torch.cuda.empty_cache()
    # ---- through host memory
    if args.host:
        pics = ctx.pinned_empty((n, w * h * 3 // 2), np.uint8)
        for gop in range(G):
            for t in range(L):
                pics[gop * L + t] = distinct[(3 * gop + t) % 25]
        shapes = {"mb_type": ((n, nmb), np.int8), "partition": ((n, nmb), np.uint8), "mv8": ((n, nmb, 4, 2), np.int16),
                  "mvr": ((n, nmb, 2), np.int16), "mvd8": ((n, nmb, 4, 2), np.int16), "nnz": ((n, nmb, 27), np.uint8), "cbp": ((n, nmb), np.int16),
                  "mode16": ((G, nmb), np.uint8), "chroma_mode": ((G, nmb), np.uint8), "modes4": ((G, nmb, 16), np.uint8),
                  "luma_dc": ((G, nmb, 16), np.int16)}
        out = {k: ctx.pinned_empty(s, t) for k, (s, t) in shapes.items()}
        packed = ctx.pinned_empty((n * nmb * 392 // 3,), np.int16)
        f_off, f_size = np.zeros(n, np.int64), np.zeros(n, np.int32)
        mb_off = ctx.pinned_empty((n, nmb), np.int32)
        run = lambda: ctx.gops_encode_host(w, h, G, L, pics, prm, out, packed, f_off, f_size, mb_off)
        run()
        t0 = time.perf_counter()
        for _ in range(2):
            run()
        dt = (time.perf_counter() - t0) / 2
        d2h = sum(a.nbytes for a in out.values()) + mb_off.nbytes + 2 * int(f_size.sum())
        res["host"] = {"ms_per_call": 1e3 * dt, "frames_per_s": n / dt, "h2d_bytes": int(pics.nbytes), "d2h_bytes": int(d2h),
                       "packed_level_bytes_per_frame": 2 * float(f_size.mean())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
