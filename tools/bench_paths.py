#!/usr/bin/env python
"""Per-path measurements of SURVEY.md 8(d) configs 3 and 4 (and the frame kernels they need) on one
GPU: every kernel timed with CUDA events on the context's stream over `--pairs` distinct frame pairs
(so that the working set exceeds the 126 MB L2), algorithmic bytes per frame from DESIGN.md, and the
pixel comparisons counted by the instrumented CPU oracle on the same input.

    python tools/bench_paths.py [--width 1920 --height 1080 --pairs 12 --reps 3] > gpurun_out/paths.json

Prints ONE JSON object.  Not the headline bench (that is bench.py); this is the evidence behind the
per-kernel roofline table in DESIGN.md.  `--only me16` etc. restricts the run (for ncu captures).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZES = [(0, 16, 16, "16x16"), (1, 16, 8, "16x8"), (2, 8, 16, "8x16"), (3, 8, 8, "8x8"), (4, 8, 4, "8x4"),
         (5, 4, 8, "4x8"), (6, 4, 4, "4x4")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--pairs", type=int, default=12)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--qp", type=int, default=26)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--batched-only", action="store_true", help="skip the one-frame-per-launch variants (ncu captures)")
    args = ap.parse_args()

    import torch
    import __graft_entry__ as ge
    import cpu_checkers as cc
    pkg = ge.load_package()
    ctx = pkg.Context(0)
    stream = ctx.torch_stream()
    w, h, P = args.width, args.height, args.pairs
    g = pkg.geometry(w, h)
    nf = P + 1
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    only = set(x for x in args.only.split(",") if x)

    def want(name):
        return not only or name in only

    frames = np.stack([pkg.synth_frame(w, h, i) for i in range(nf)])
    i420 = torch.from_numpy(frames).cuda()
    slots = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn, reps=args.reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev[0].record(stream)
        for _ in range(reps):
            fn()
        ev[1].record(stream)
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    out = {"config": {"width": w, "height": h, "pairs": P, "frames": nf, "qp": args.qp, "mb_count": g.mb_count,
                      "slot_MB": g.slot_bytes / 1e6, "working_set_MB": nf * g.slot_bytes / 1e6,
                      "hbm_peak_gbs": peak, "timing": "CUDA events on the context stream, mean of reps"},
           "kernels": {}}

    def report(name, ms_total, frames_n, bytes_per_frame, note=""):
        ms = ms_total / frames_n
        gbs = bytes_per_frame / (ms / 1e3) / 1e9 if bytes_per_frame else None
        out["kernels"][name] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                                "algorithmic_MB_per_frame": bytes_per_frame / 1e6 if bytes_per_frame else None,
                                "achieved_GBs": gbs, "frac_of_hbm_peak": gbs / peak if gbs else None, "note": note}

    luma_px, pad_luma = g.luma_w * g.luma_h, g.luma_stride * (g.luma_h + 64)
    # ---- frame kernels
    t = timed(lambda: ctx.frame_load_i420(g, i420, slots, nf))
    report("frame_load_i420", t, nf, w * h * 1.5 + luma_px * 1.5, "read I420, write padded luma + NV12")
    t = timed(lambda: ctx.frame_expand_border(g, slots, nf))
    report("frame_expand_border", t, nf, 2 * (pad_luma - luma_px) * 1.5, "write (and read the edge of) the padding only")
    t = timed(lambda: ctx.frame_filter(g, slots, nf))
    report("frame_filter_hpel", t, nf, luma_px * 4, "SURVEY 8(d): 2.09 MB read + 6.27 MB write")
    t = timed(lambda: ctx.frame_init_lowres(g, slots, nf))
    report("frame_init_lowres", t, nf, luma_px * 2, "SURVEY 8(d): 2.09 MB read + 2.09 MB write")

    # ---- lookahead on the pairs: gives the lowres MVs that seed config 3's mvp
    n = nf
    b = np.arange(n, dtype=np.int32)
    p0 = (b - 1).astype(np.int32)
    wi = np.ones(n, np.uint8)
    d_mvs = torch.zeros((n, g.mb_count, 2), dtype=torch.int16, device="cuda")
    d_costs = torch.zeros((n, g.mb_count), dtype=torch.int32, device="cuda")
    d_sums = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    t = timed(lambda: ctx.lookahead_frame_cost(g, slots, b, p0, wi, d_mvs, d_costs, d_sums))
    report("lookahead_frame_cost", t, P, 5 * g.lowres_w * g.lowres_h + 8 * g.mb_count,
           f"{P} pairs in one launch (chain of frames): latency bound at this batch size")
    la_mv = d_mvs.cpu().numpy()

    # ---- config 3: HEX + subme 5 + qpel refine, every partition size
    o = None if args.no_oracle else cc.oracle()
    go = None if args.no_oracle else cc.oracle_geom(w, h)
    host_slots = None
    if o is not None:
        host_slots = slots[: 2 * g.slot_bytes].cpu().numpy()
    prm = pkg.MeParams(pkg.ME_HEX, 5, 16, args.qp, 1)
    me_out = {}
    mv16 = None
    for size, bw, bh, name in SIZES:
        if not want("me" + name) and not (size == 0 and (want("residual") or want("deblock"))):
            continue
        blocks = [pkg.tiling_blocks(g, size, la_mv[p + 1]) for p in range(P)]
        nb = len(blocks[0])
        d_blocks = [torch.from_numpy(bk.view(np.uint8)).cuda() for bk in blocks]
        d_res = [torch.zeros(nb * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda") for _ in range(P)]

        def run_me():
            for p in range(P):
                ctx.me_search_batch(g, slots[(p + 1) * g.slot_bytes:(p + 2) * g.slot_bytes],
                                    slots[p * g.slot_bytes:(p + 1) * g.slot_bytes], prm, nb, d_blocks[p], d_res[p])
        t_generic = timed(run_me)

        def run_me_sized():
            for p in range(P):
                ctx.me_search_sized(g, slots[(p + 1) * g.slot_bytes:(p + 2) * g.slot_bytes],
                                    slots[p * g.slot_bytes:(p + 1) * g.slot_bytes], prm, size, nb, d_blocks[p], d_res[p])
        t_single = timed(run_me_sized)
        d_blocks_all = torch.cat(d_blocks)
        d_res_all = torch.zeros(P * nb * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        t = timed(lambda: ctx.me_search_sized_frames(g, slots[g.slot_bytes:], slots, P, prm, size, nb, d_blocks_all, d_res_all))
        assert torch.equal(d_res_all, torch.cat(d_res)), "batched and per-frame launches disagree"
        rec = {"blocks_per_frame": nb, "ms_per_frame": t / P, "frames_per_s": 1e3 * P / t,
               "one_frame_per_launch_ms_per_frame": t_single / P,
               "generic_warp_per_block_ms_per_frame": t_generic / P}
        if size == 0:
            mv16 = [r.cpu().numpy().view(cc.ME_RESULT_DTYPE)["mv"].copy() for r in d_res]
        if o is not None:
            wantr = np.zeros(nb, cc.ME_RESULT_DTYPE)
            cnt0, cnt1 = (C.c_int64 * 4)(), (C.c_int64 * 4)()
            cprm = cc.MeParams(1, 5, 16, args.qp, 1)
            o.xo_work_counters(cnt0, 1)
            t0 = time.perf_counter()
            o.xo_me_search_batch(C.byref(go), cc.ptr(host_slots[g.slot_bytes:]), cc.ptr(host_slots[: g.slot_bytes]),
                                 C.byref(cprm), nb, blocks[0].ctypes.data_as(C.c_void_p), wantr.ctypes.data_as(C.c_void_p))
            cpu_s = time.perf_counter() - t0
            o.xo_work_counters(cnt1, 0)
            got = d_res[0].cpu().numpy().view(cc.ME_RESULT_DTYPE)
            rec["parity_vs_oracle_frame0"] = bool(np.array_equal(got, wantr))
            sad_px, satd_px = int(cnt1[0]), int(cnt1[1])
            rec.update({"oracle_sad_pix_per_frame": sad_px, "oracle_satd_pix_per_frame": satd_px,
                        "sad_gpix_per_s": sad_px / (t / P / 1e3) / 1e9, "satd_gpix_per_s": satd_px / (t / P / 1e3) / 1e9,
                        "cpu_oracle_1core_ms_per_frame": cpu_s * 1e3})
        rec["compulsory_MB_per_frame"] = 5 * luma_px / 1e6
        rec["compulsory_GBs"] = 5 * luma_px / (t / P / 1e3) / 1e9
        me_out[name] = rec
    out["me_hex_subme5"] = me_out
    if me_out:
        tot = sum(r["ms_per_frame"] for r in me_out.values())
        out["me_hex_subme5_all_sizes"] = {"ms_per_frame": tot, "frames_per_s": 1e3 / tot}
        if o is not None:
            out["me_hex_subme5_all_sizes"].update({
                "sad_gpix_per_s": sum(r["oracle_sad_pix_per_frame"] for r in me_out.values()) / (tot / 1e3) / 1e9,
                "satd_gpix_per_s": sum(r["oracle_satd_pix_per_frame"] for r in me_out.values()) / (tot / 1e3) / 1e9,
                "cpu_oracle_1core_ms_per_frame": sum(r["cpu_oracle_1core_ms_per_frame"] for r in me_out.values())})

    # ---- config 4: MC from the 16x16 MVs, residual coding, deblocking
    if mv16 is not None and (want("residual") or want("deblock")):
        nmb = g.mb_count
        d_mv = [torch.from_numpy(m).cuda() for m in mv16]
        pred = torch.zeros(P * g.slot_bytes, dtype=torch.uint8, device="cuda")
        lv = torch.zeros((nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
        nz = torch.zeros((nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
        cbp = [torch.zeros(nmb, dtype=torch.int16, device="cuda") for _ in range(P)]
        sl = lambda buf, p: buf[p * g.slot_bytes:(p + 1) * g.slot_bytes]

        def run_mc():
            for p in range(P):
                ctx.mc_frame(g, sl(slots, p), d_mv[p], sl(pred, p))
        t = timed(run_mc)
        report("mc_frame_16x16", t, P, nmb * 384 * 2, "one frame per launch: 384 B in (+halo) + 384 B out per MB")
        d_mv_all = torch.stack(d_mv)
        t = timed(lambda: ctx.mc_frames(g, slots, P, d_mv_all, pred))
        report("mc_frames_batched", t, P, nmb * 384 * 2, f"{P} frames per launch")

        def run_res():
            for p in range(P):
                ctx.residual_frame(g, sl(slots, p + 1), sl(pred, p), args.qp, lv, nz, cbp[p])
        run_res()                                # warm-up (first launch of the kernel); fills cbp[]
        run_mc()
        if not args.batched_only:
            t = timed(run_res, warm=0, reps=1)   # in place: one pass over freshly predicted frames
            report("residual_frame", t, P, nmb * (384 * 3 + 784 + 29), "one frame per launch; SURVEY 8(d): ~2.0 kB/MB")
        lv_all = torch.zeros((P, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
        nz_all = torch.zeros((P, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
        cbp_b = torch.zeros((P, nmb), dtype=torch.int16, device="cuda")
        ctx.mc_frames(g, slots, P, d_mv_all, pred)
        t = timed(lambda: ctx.residual_frames(g, slots[g.slot_bytes:], pred, P, args.qp, lv_all, nz_all, cbp_b), warm=0, reps=1)
        report("residual_frames_batched", t, P, nmb * (384 * 3 + 784 + 29), f"{P} frames per launch")

        # deblock inputs: P_L0 16x16 everywhere, bS from a random-but-plausible field (timing only; parity is tests/)
        rng = np.random.RandomState(1)
        mb_type = torch.from_numpy(np.full((P, nmb), 4, np.int8)).cuda()
        part = torch.from_numpy(np.full((P, nmb), 16, np.uint8)).cuda()
        bs_h = (rng.rand(P, nmb, 2, 8, 4) < 0.35).astype(np.uint8) * rng.randint(1, 3, (P, nmb, 2, 8, 4)).astype(np.uint8)
        bs = torch.from_numpy(bs_h).cuda()
        cbp_all = torch.stack(cbp)

        def run_db1():
            for p in range(P):
                ctx.deblock_frame(g, sl(pred, p), mb_type[p], part[p], cbp[p], bs[p], args.qp, 0, 0)
        if not args.batched_only:
            run_db1()                            # warm-up; deblocking an already deblocked frame is still a full pass
            t = timed(run_db1, warm=0, reps=1)
            report("deblock_frame_one_per_launch", t, P, nmb * (768 + 64),
                   "one frame per launch: bound by the wavefront critical path (mb_w + 2 mb_h macroblock times)")
        else:
            ctx.deblock_frames(g, pred, P, mb_type, part, cbp_all, bs, args.qp, 0, 0)
        t = timed(lambda: ctx.deblock_frames(g, pred, P, mb_type, part, cbp_all, bs, args.qp, 0, 0), warm=0, reps=1)
        report("deblock_frames_batched", t, P, nmb * (768 + 64), f"{P} frames per launch; SURVEY 8(d): 768 B rd+wr + 64 B bS per MB")
        # one frame per call is a 2 us kernel behind a launch; a clip's macroblocks go in one call
        nb = min(P, 48) * nmb
        nnz = torch.from_numpy((rng.rand(nb, 120) < 0.3).astype(np.uint8)).cuda()
        ref = torch.from_numpy(rng.randint(-1, 2, (nb, 2, 40)).astype(np.int8)).cuda()
        mvs = torch.from_numpy(rng.randint(-6, 7, (nb, 2, 40, 2)).astype(np.int16)).cuda()
        bs2 = torch.zeros((nb, 2, 8, 4), dtype=torch.uint8, device="cuda")
        t = timed(lambda: ctx.deblock_strength(nmb, nnz, ref, mvs, bs2), reps=10)
        report("deblock_strength_one_frame_per_call", t, 1, nmb * (120 + 80 + 320 + 64), "scan8-layout inputs, 64 B out per MB")
        t = timed(lambda: ctx.deblock_strength(nb, nnz, ref, mvs, bs2), reps=10)
        report("deblock_strength", t, nb // nmb, nmb * (120 + 80 + 320 + 64), f"{nb // nmb} frames per call")

    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
