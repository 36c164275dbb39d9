#!/usr/bin/env python
"""Phase-cycle breakdown of xd_la_inter_kernel (debug aid): runs the device-resident lookahead step of
bench.py with `--clips` clips and prints the average clock64() cycles per block spent in each phase."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=16)
    ap.add_argument("--clip-len", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-timing", action="store_true", help="only time the launch with CUDA events (production kernel)")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    ctx = pkg.Context(0)
    w, h = args.width, args.height
    g = pkg.geometry(w, h)
    n = args.clips * args.clip_len
    luma = np.stack([pkg.synth_frame(w, h, i, luma_only=True) for i in range(n)])
    luma_dev = torch.from_numpy(luma).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    d_mvs = torch.zeros((n, g.mb_count, 2), dtype=torch.int16, device="cuda")
    d_costs = torch.zeros((n, g.mb_count), dtype=torch.int32, device="cuda")
    d_sums = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    b = np.arange(n, dtype=np.int32)
    p0 = np.where(b % args.clip_len == 0, -1, b - 1).astype(np.int32)
    wi = np.ones(n, np.uint8)
    ctx.frame_load_luma(g, luma_dev, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    lib = pkg.lib()
    out = (C.c_uint64 * 10)()
    for _ in range(2):
        ctx.lookahead_frame_cost(g, slots, b, p0, wi, d_mvs, d_costs, d_sums)
    if args.no_timing:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = ctx.torch_stream()
        ctx.profile_enable(True)
        for _ in range(5):
            ctx.lookahead_frame_cost(g, slots, b, p0, wi, d_mvs, d_costs, d_sums)
        ms, cnt = ctx.profile_read(pkg.PROF_LA_INTER)
        print(f"clips {args.clips} ctas/sm {os.environ.get('X264DSP_LA_CTAS_PER_SM', 'max')}: "
              f"la_inter {ms / cnt:.3f} ms/launch -> {n / (ms / cnt) * 1e3:.0f} frames/s (kernel only)")
        ctx.close()
        return
    lib.x264dsp_debug_lookahead_timing(ctx._h, 1, 1, out)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = ctx.torch_stream()
    ev0.record(stream)
    ctx.lookahead_frame_cost(g, slots, b, p0, wi, d_mvs, d_costs, d_sums)
    ev1.record(stream)
    torch.cuda.synchronize()
    lib.x264dsp_debug_lookahead_timing(ctx._h, 0, 1, out)
    names = ["wait", "setup", "zero_satd", "candidates", "diamond", "subpel", "tail"]
    blocks = out[7]
    tot = sum(out[k] for k in range(7))
    print(f"clips {args.clips}: launch {ev0.elapsed_time(ev1):.3f} ms, blocks {blocks}, cycles/block {tot / blocks:.0f}")
    for k, nm in enumerate(names):
        print(f"  {nm:10s} {out[k] / blocks:9.0f} cycles/block  {100 * out[k] / tot:5.1f}%")
    print(f"  first wait of a row: {out[8] / max(out[9], 1):.0f} cycles/row over {out[9]} rows "
          f"(row work {tot / max(out[9], 1):.0f} cycles)")
    ctx.close()


if __name__ == "__main__":
    main()
