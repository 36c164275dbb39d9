#!/usr/bin/env python
"""bench.py -- the x264-dsp hot path on B200: 1080p lowres-lookahead motion estimation.

Workload (BASELINE.json configs[1]): 1080p synthetic clips of 8 frames; for every clip the lowres
planes of all 8 frames are built (x264_frame_init_lowres) and x264_slicetype_frame_cost is run
intra-only on frame 0 and as a P analysis (DIA + SAD full-pel, half-pel refine, SATD re-cost, 3-mode
intra SATD) on frames 1..7.  One step = `--clips` independent clips per GPU (default 128: the step's
input, 2.1 GB of luma, is far larger than the 126 MB L2, and the wavefronts of 896 frame pairs keep
every warp slot of the 148 SMs busy).  Frames/s counts all frames.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                      the reference's own C path on host cores

Under torchrun (N > 1) every rank drives its own GPU on its own clips (weak scaling, no collective
on the data path); time = max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOOKAHEAD_BYTES_PER_PAIR = None   # filled from the geometry: 5 lowres planes in + 8 B per block out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--clips", type=int, default=256, help="independent 8-frame clips per GPU per step")
    ap.add_argument("--clip-len", type=int, default=8)
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-me", action="store_true", help="skip the configs[2] / configs[3] measurements (ME search, recon)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="2 = the headline (BASELINE.json configs[1], 1080p lookahead, + configs[2] and [3] as sub-objects); "
                         "5 = configs[4]: ONE 4K sequence of --seq-frames frames, lookahead + 16x16 / 8x8 motion search, "
                         "split by frame range over the ranks, gathered with NCCL and compared with rank 0's own single-GPU run")
    ap.add_argument("--seq-frames", type=int, default=64)
    return ap.parse_args()


def load_package():
    import __graft_entry__ as ge
    return ge.load_package()


def make_clips(pkg, w, h, clips, clip_len, first_clip=0, out=None):
    """luma of `clips` consecutive synthetic clips, [clips*clip_len, h*w] uint8"""
    n = clips * clip_len
    luma = np.empty((n, w * h), np.uint8) if out is None else out
    for i in range(n):
        y = luma[i]
        rc = pkg.lib().x264dsp_synth_frame(w, h, first_clip * clip_len + i, -1, y.ctypes.data_as(pkg.u8p), None, None)
        assert rc == 0
    return luma


# --------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """summary of the samples received between wall-clock times t0 and t1 (all if None)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for stamp, line in self.lines:
            if t0 is not None and not (t0 <= stamp <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# secondary measurement (BASELINE.json configs[2]): full-resolution motion search of every partition
# size, HEX + subme 5 + qpel refine, on the first `pairs` frame pairs of clip 0

ME_SIZES = ("16x16", "16x8", "8x16", "8x8", "8x4", "4x8", "4x4")


def me_search_measure(pkg, ctx, torch, g, luma_dev, la_mvs, pairs, reps=3, qp=26):
    """returns (per-size ms per frame, blocks arrays of pair 0, the device slots) -- CUDA events on the
    context's stream; one launch per size covers all `pairs` frame pairs"""
    stream = ctx.torch_stream()
    nf = pairs + 1
    slots = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_luma(g, luma_dev, slots, nf)
    ctx.frame_expand_border(g, slots, nf)
    ctx.frame_filter(g, slots, nf)
    prm = pkg.MeParams(pkg.ME_HEX, 5, 16, qp, 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms, blocks0 = {}, {}
    for size, name in enumerate(ME_SIZES):
        blocks = [pkg.tiling_blocks(g, size, la_mvs[p + 1]) for p in range(pairs)]
        nb = len(blocks[0])
        d_blocks = torch.from_numpy(np.concatenate(blocks).view(np.uint8)).cuda()
        d_res = torch.zeros(pairs * nb * pkg.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        run = lambda: ctx.me_search_sized_frames(g, slots[g.slot_bytes:], slots, pairs, prm, size, nb, d_blocks, d_res)
        run()
        torch.cuda.synchronize()
        ev[0].record(stream)
        for _ in range(reps):
            run()
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms[name] = ev[0].elapsed_time(ev[1]) / reps / pairs
        blocks0[name] = (blocks[0], d_res[: nb * pkg.ME_RESULT_DTYPE.itemsize].cpu().numpy())
    return ms, blocks0, slots


def me_search_cpu_counts(pkg, g, slots, blocks0, qp=26):
    """cpu_baseline leg: the oracle's x264_me_search_ref restatement on frame pair 0 (one core), which
    also counts the pixel comparisons the reference issues and checks the GPU results"""
    import cpu_checkers as cc
    o = cc.oracle()
    go = cc.oracle_geom(g.width, g.height)
    host = slots[: 2 * g.slot_bytes].cpu().numpy()
    out = {}
    for name, (blocks, gpu_res) in blocks0.items():
        nb = len(blocks)
        want = np.zeros(nb, cc.ME_RESULT_DTYPE)
        c0, c1 = (C.c_int64 * 4)(), (C.c_int64 * 4)()
        prm = cc.MeParams(1, 5, 16, qp, 1)
        o.xo_work_counters(c0, 1)
        t0 = time.perf_counter()
        o.xo_me_search_batch(C.byref(go), cc.ptr(host[g.slot_bytes:]), cc.ptr(host[: g.slot_bytes]), C.byref(prm), nb,
                             blocks.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
        dt = time.perf_counter() - t0
        o.xo_work_counters(c1, 0)
        out[name] = {"sad_pix": int(c1[0]), "satd_pix": int(c1[1]), "cpu_s": dt,
                     "bit_exact": bool(np.array_equal(gpu_res.view(cc.ME_RESULT_DTYPE), want))}
    return out


# --------------------------------------------------------------------------------------------
# secondary measurement (BASELINE.json configs[3]): motion compensation from 16x16 MVs, residual coding
# (DCT / quant / dequant / IDCT / decimation) and in-loop deblocking of whole frames, QP 26

def recon_measure(pkg, ctx, torch, g, w, h, n_frames, first_frame=0, reps=3, qp=26):
    """n_frames frames coded against their predecessors, frame-batched launches; returns the per-kernel
    ms per frame and what the cpu leg needs to check frame 0 bit for bit"""
    stream = ctx.torch_stream()
    nf = n_frames + 1
    pics = np.stack([pkg.synth_frame(w, h, first_frame + i) for i in range(nf)])
    i420 = torch.from_numpy(pics).cuda()
    slots = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, i420, slots, nf)
    ctx.frame_expand_border(g, slots, nf)
    ctx.frame_filter(g, slots, nf)
    nmb = g.mb_count
    # 16x16 MVs: the true pan of the synthetic clip (3, 2 luma samples per frame) plus +-1 qpel of jitter
    rng = np.random.RandomState(5)
    mv = (np.array([12, 8]) + rng.randint(-1, 2, (n_frames, nmb, 2))).astype(np.int16)
    d_mv = torch.from_numpy(mv).cuda()
    pred = torch.zeros(n_frames * g.slot_bytes, dtype=torch.uint8, device="cuda")
    lv = torch.zeros((n_frames, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
    nz = torch.zeros((n_frames, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
    cbp = torch.zeros((n_frames, nmb), dtype=torch.int16, device="cuda")
    mb_type = torch.from_numpy(np.full((n_frames, nmb), 4, np.int8)).cuda()          # P_L0
    part = torch.from_numpy(np.full((n_frames, nmb), 16, np.uint8)).cuda()           # D_16x16
    bs_h = (rng.rand(n_frames, nmb, 2, 8, 4) < 0.35).astype(np.uint8) * rng.randint(1, 3, (n_frames, nmb, 2, 8, 4)).astype(np.uint8)
    bs = torch.from_numpy(bs_h).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ms = {"mc": 0.0, "residual": 0.0, "deblock": 0.0}
    for rep in range(reps + 1):
        ev[0].record(stream)
        ctx.mc_frames(g, slots, n_frames, d_mv, pred)
        ev[1].record(stream)
        ctx.residual_frames(g, slots[g.slot_bytes:], pred, n_frames, qp, lv, nz, cbp)
        ev[2].record(stream)
        if rep == 0:
            torch.cuda.synchronize()                       # the clone runs on torch's stream, the kernels on the context's
            recon0 = pred[: g.slot_bytes].clone()          # frame 0 before deblocking, for the check
            torch.cuda.synchronize()
        ctx.deblock_frames(g, pred, n_frames, mb_type, part, cbp, bs, qp, 0, 0)
        ev[3].record(stream)
        torch.cuda.synchronize()
        if rep:                                            # rep 0 is the warm-up
            ms["mc"] += ev[0].elapsed_time(ev[1]) / reps / n_frames
            ms["residual"] += ev[1].elapsed_time(ev[2]) / reps / n_frames
            ms["deblock"] += ev[2].elapsed_time(ev[3]) / reps / n_frames
    check = {"slots": slots[: 2 * g.slot_bytes].cpu().numpy(), "mv": mv[0], "bs": bs_h[0], "qp": qp,
             "recon": recon0.cpu().numpy(), "deblocked": pred[: g.slot_bytes].cpu().numpy(),
             "levels": lv[0].cpu().numpy(), "nnz": nz[0].cpu().numpy(), "cbp": cbp[0].cpu().numpy()}
    return ms, check


def recon_cpu_check(pkg, g, check):
    """cpu_baseline leg: frame 0 of the measurement through the oracle (one core), compared bit for bit"""
    import cpu_checkers as cc
    from cpu_checkers import ptr, i16p, i8p
    o = cc.oracle()
    go = cc.oracle_geom(g.width, g.height)
    nmb = g.mb_count
    host = check["slots"]
    t0 = time.perf_counter()
    pred = np.zeros(go.slot_bytes, np.uint8)
    o.xo_mc_frame(C.byref(go), ptr(host[: g.slot_bytes]), ptr(check["mv"], i16p), ptr(pred))
    lv = np.zeros((nmb, pkg.RES_LEVELS_PER_MB), np.int16)
    nz = np.zeros((nmb, pkg.RES_NNZ_PER_MB), np.uint8)
    cbp = np.zeros(nmb, np.int16)
    o.xo_residual_frame(C.byref(go), ptr(host[g.slot_bytes:]), ptr(pred), check["qp"], ptr(lv, i16p), ptr(nz), ptr(cbp, i16p))
    ok = (np.array_equal(pred, check["recon"]) and np.array_equal(lv, check["levels"])
          and np.array_equal(nz, check["nnz"]) and np.array_equal(cbp, check["cbp"]))
    mb_type, part = np.full(nmb, 4, np.int8), np.full(nmb, 16, np.uint8)
    o.xo_deblock_frame(C.byref(go), ptr(pred), ptr(mb_type, i8p), ptr(part), ptr(cbp, i16p), ptr(check["bs"]), check["qp"], 0, 0)
    dt = time.perf_counter() - t0
    ok = ok and np.array_equal(pred, check["deblocked"])
    return bool(ok), dt, int(np.count_nonzero(cbp))


# --------------------------------------------------------------------------------------------
# CPU baselines of configs[2] and configs[3] from the UNMODIFIED reference (oracle/_ref): its x264_me_search_ref +
# x264_me_refine_qpel over the same block lists, its x264_mb_mc + x264_macroblock_encode + x264_frame_deblock_row over
# the same frames, on all host threads (one encoder instance per thread), for the -O2 and the -O3 build

def _run_threads(fn, threads):
    """fn(t) on `threads` threads behind a barrier; returns the seconds between the common start and the last finish"""
    barrier = threading.Barrier(threads + 1)
    done = threading.Barrier(threads + 1)

    def run(t):
        barrier.wait()
        fn(t)
        done.wait()
    ths = [threading.Thread(target=run, args=(t,), daemon=True) for t in range(threads)]
    for th in ths:
        th.start()
    barrier.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for th in ths:
        th.join()
    return dt


def cpu_me_reference(which, w, h, pics, blocks_by_size, threads, qp=26, target_s=4.0):
    """frame pair 0 of the ME measurement through the reference: every thread searches a contiguous share of every
    partition size's block list.  Returns (frames/s for all sizes together, results by size of the last pass)"""
    import cpu_checkers as cc
    lib = cc.ref_o3() if which == "O3" else cc.ref()
    if lib is None:
        return None, None
    state = []
    for t in range(threads):
        enc = cc.RefEncoder(w, h, me=1, subme=5, me_range=16, qp=qp, lib=lib)
        fref, fenc = enc.new_frame(True), enc.new_frame(False)
        enc.load(fref, pics[0])
        enc.load(fenc, pics[1])
        lib.xref_frame_filter_all(enc.h, fref)
        state.append((enc, fenc, fref))
    out = {name: np.zeros(len(b), cc.ME_RESULT_DTYPE) for name, b in blocks_by_size.items()}
    shares = {name: np.linspace(0, len(b), threads + 1).astype(int) for name, b in blocks_by_size.items()}
    sizes = {name: i for i, name in enumerate(ME_SIZES)}

    def work(t):
        enc, fenc, fref = state[t]
        for name, b in blocks_by_size.items():
            lo, hi = shares[name][t], shares[name][t + 1]
            if hi > lo:
                lib.xref_me_search_batch(enc.h, fenc, fref, qp, 1, 5, 16, 1, b[lo:hi].ctypes.data_as(C.c_void_p), int(hi - lo),
                                         out[name][lo:hi].ctypes.data_as(C.c_void_p))
    dt = _run_threads(work, threads)                       # warm-up and a first estimate
    reps = max(1, min(64, int(target_s / max(dt, 1e-3))))
    dt = _run_threads(lambda t: [work(t) for _ in range(reps)], threads)
    return reps / dt, out


def cpu_recon_reference(which, w, h, pics, mv, bs, threads, qp=26, target_s=4.0):
    """frame 0 of the recon measurement through the reference, one frame per thread and pass.  Returns
    (frames/s, (levels, nnz, cbp, deblocked luma plane area, deblocked chroma plane area) of thread 0)"""
    import cpu_checkers as cc
    from cpu_checkers import ptr, i16p, i8p
    lib = cc.ref_o3() if which == "O3" else cc.ref()
    if lib is None:
        return None, None
    g = cc.oracle_geom(w, h)
    n = g.mb_count
    mb_type, part = np.full(n, 4, np.int8), np.full(n, 16, np.uint8)
    state = []
    for t in range(threads):
        enc = cc.RefEncoder(w, h, me=1, subme=5, me_range=16, qp=qp, lib=lib)
        fref, fdec, fenc = enc.new_frame(True), enc.new_frame(True), enc.new_frame(False)
        enc.load(fref, pics[0])
        enc.load(fenc, pics[1])
        enc.load(fdec, pics[1])
        lib.xref_frame_filter_all(enc.h, fref)
        state.append((enc, fenc, fref, fdec, np.zeros((n, 392), np.int16), np.zeros((n, 27), np.uint8), np.zeros(n, np.int16)))

    def work(t):
        enc, fenc, fref, fdec, lv, nz, cbp = state[t]
        lib.xref_recon_frame(enc.h, fenc, fref, fdec, ptr(mv, i16p), qp, ptr(lv, i16p), ptr(nz), ptr(cbp, i16p))
        lib.xref_deblock_frame(enc.h, fdec, ptr(mb_type, i8p), ptr(part), ptr(cbp, i16p), ptr(bs), qp, 0, 0)
    dt = _run_threads(work, threads)
    reps = max(1, min(64, int(target_s / max(dt, 1e-3))))
    dt = _run_threads(lambda t: [work(t) for _ in range(reps)], threads)
    enc, fenc, fref, fdec, lv, nz, cbp = state[0]
    y = enc.buffer(fdec, 10, 4 * g.luma_plane_size)[g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
    c = enc.buffer(fdec, 11, g.chroma_plane_size)[g.chroma_origin:][: (g.luma_h // 2) * g.chroma_stride].reshape(
        g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
    return threads * reps / dt, (lv.copy(), nz.copy(), cbp.copy(), y.copy(), c.copy())


PF_FRAMES = 384
PF_E2E_FRAMES = 384
GOPS_N, GOPS_LEN = 64, 12            # closed GOPs per call, frames per GOP (the `gops` entry)
# (name, (me_method, subme, analyse.inter != 0))
PF_SETTINGS = (("dia_subme1", (0, 1, 0)), ("hex_subme5", (1, 5, 0)), ("hex_subme5_psub16x16", (1, 5, 1)))


def pframe_measure(pkg, ctx, torch, g, w, h, n_frames, me, subme, qp=26, reps=3, part=0):
    """BASELINE north_star's closed loop, SURVEY 8(f) N2: x264_macroblock_analyse + x264_macroblock_encode for every macroblock
    of n_frames independent 1080p P frames in ONE launch (x264dsp_p_frames_dev).  Inputs as an encoder has them: reference
    frame border-expanded and half-pel filtered, the lookahead's vectors of each pair as first search candidate.  Returns
    (ms per frame, ms for one frame alone, skipped share, (host slots of pair 0, lowres mvs, results of frame 0))"""
    stream = ctx.torch_stream()
    nmb = g.mb_count
    distinct = min(n_frames, 24) + 1
    frames = np.stack([pkg.synth_frame(w, h, i) for i in range(distinct)])
    one = torch.zeros(distinct * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(frames).cuda(), one, distinct)
    ctx.frame_expand_border(g, one, distinct)
    ctx.frame_filter(g, one, distinct)
    ctx.frame_init_lowres(g, one, distinct)
    src = torch.zeros((n_frames + 1) * g.slot_bytes, dtype=torch.uint8, device="cuda")
    for k in range(n_frames + 1):                      # pair k = slot k -> slot k + 1, cycling through the distinct frames
        j = k % distinct
        src[k * g.slot_bytes:(k + 1) * g.slot_bytes] = one[j * g.slot_bytes:(j + 1) * g.slot_bytes]
    torch.cuda.synchronize()                           # torch's copies are on its own stream; the context's streams do not wait for it
    del one
    b = np.arange(1, n_frames + 1, dtype=np.int32)
    d_lmv = torch.zeros((n_frames, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n_frames, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n_frames, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, src, b, b - 1, np.ones(n_frames, np.uint8), d_lmv, d_lc, d_ls)
    o = dict(mb_type=torch.zeros((n_frames, nmb), dtype=torch.int8, device="cuda"),
             mv=torch.zeros((n_frames, nmb, 4, 2) if part else (n_frames, nmb, 2), dtype=torch.int16, device="cuda"),
             mvr=torch.zeros((n_frames, nmb, 2), dtype=torch.int16, device="cuda"),
             levels=torch.zeros((n_frames, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
             nnz=torch.zeros((n_frames, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
             cbp=torch.zeros((n_frames, nmb), dtype=torch.int16, device="cuda"))
    if part:
        o["partition"] = torch.zeros((n_frames, nmb), dtype=torch.uint8, device="cuda")
    recon = torch.zeros(n_frames * g.slot_bytes, dtype=torch.uint8, device="cuda")
    prm = pkg.PFrameParams(me, subme, 16, qp, 512, 1, 0, part)

    def run(n):
        if part:
            ctx.p_frames_part(g, src[g.slot_bytes:], src, recon, n, prm, d_lmv[:n], None, o["mb_type"], o["partition"], o["mv"],
                              o["mvr"], o["levels"], o["nnz"], o["cbp"])
            return
        ctx.p_frames(g, src[g.slot_bytes:], src, recon, n, prm, d_lmv[:n], None, o["mb_type"], o["mv"], o["mvr"], o["levels"],
                     o["nnz"], o["cbp"])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for n in (n_frames, 1):
        run(n)
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(reps):
            run(n)
        ev1.record(stream)
        torch.cuda.synchronize()
        out.append(ev0.elapsed_time(ev1) / reps / n)
    run(n_frames)
    torch.cuda.synchronize()
    types = o["mb_type"].cpu().numpy()
    check = (src[: 2 * g.slot_bytes].cpu().numpy(), d_lmv[0].cpu().numpy(),
             {k: v[0].cpu().numpy() for k, v in o.items()}, recon[: g.slot_bytes].cpu().numpy(), (me, subme, qp, part))
    return out[0], out[1], float((types == pkg.MB_P_SKIP).mean()), check


def pframe_e2e(pkg, ctx, g, w, h, n_frames, me, subme, qp=26, reps=3, dense=True):
    """the same through the host-memory door (x264dsp_p_frames_host_packed): pinned pictures in; types, vectors, mvd, nnz, cbp and
    the levels as the compact stream the entropy coder reads out (the reconstruction stays on the device, where an encoder needs
    it as the next reference); every copy inside the timed region.  Returns (seconds per call, h2d bytes, d2h bytes,
    seconds per call of the dense door x264dsp_p_frames_host with the reconstruction, its d2h bytes)"""
    nmb = g.mb_count
    pics = ctx.pinned_empty((n_frames + 1, w * h * 3 // 2), np.uint8)
    for i in range(n_frames + 1):
        pics[i] = pkg.synth_frame(w, h, i % 25)
    o = {"mb_type": ctx.pinned_empty((n_frames, nmb), np.int8), "mv": ctx.pinned_empty((n_frames, nmb, 2), np.int16),
         "mvr": ctx.pinned_empty((n_frames, nmb, 2), np.int16), "mvd": ctx.pinned_empty((n_frames, nmb, 2), np.int16),
         "nnz": ctx.pinned_empty((n_frames, nmb, pkg.RES_NNZ_PER_MB), np.uint8), "cbp": ctx.pinned_empty((n_frames, nmb), np.int16),
         "mb_offset": ctx.pinned_empty((n_frames, nmb), np.int32)}
    packed = ctx.pinned_empty((n_frames * nmb * pkg.RES_LEVELS_PER_MB // 4,), np.int16)
    f_off = np.zeros(n_frames + 1, np.int64)
    prm = pkg.PFrameParams(me, subme, 16, qp, 512, 1, 0)

    def run():
        ctx.p_frames_host_packed(w, h, n_frames, pics, prm, o["mb_type"], None, o["mv"], o["mvr"], o["mvd"], packed, f_off,
                                 o["mb_offset"], o["nnz"], o["cbp"])
    run()
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    dt = (time.perf_counter() - t0) / reps
    d2h = int(sum(a.nbytes for a in o.values()) + 2 * int(f_off[-1]))
    if not dense:
        return dt, int(pics.nbytes), d2h, None, None
    # the dense door for comparison: all 392 levels of every macroblock and the reconstructed pictures come back as well
    levels = ctx.pinned_empty((n_frames, nmb, pkg.RES_LEVELS_PER_MB), np.int16)
    recon = ctx.pinned_empty((n_frames, w * h * 3 // 2), np.uint8)

    def run_dense():
        ctx.p_frames_host(w, h, n_frames, pics, prm, o["mb_type"], o["mv"], o["mvr"], o["mvd"], levels, o["nnz"], o["cbp"], recon)
    run_dense()
    t0 = time.perf_counter()
    for _ in range(reps):
        run_dense()
    dt_dense = (time.perf_counter() - t0) / reps
    d2h_dense = int(sum(a.nbytes for k, a in o.items() if k != "mb_offset") + levels.nbytes + recon.nbytes)
    return dt, int(pics.nbytes), d2h, dt_dense, d2h_dense


def gops_measure(pkg, ctx, torch, g, w, h, n_gops, gop_len, me, subme, qp=26):
    """closed GOPs (an I frame and gop_len - 1 P frames, in-loop filter on) coded entirely on the device
    (x264dsp_gops_encode_dev: one launch of every stage per GOP position over the frames of all GOPs), and the same through host
    memory (x264dsp_gops_encode_host: pinned pictures in, the entropy coder's input out).  Returns (device ms per call, host
    seconds per call, h2d bytes, d2h bytes)"""
    stream = ctx.torch_stream()
    nmb, n = g.mb_count, n_gops * gop_len
    prm = pkg.GopEncodeParams(me, subme, 16, qp - 3, qp, 512, 1, 0, 1, 0, 0)
    distinct = [pkg.synth_frame(w, h, i) for i in range(25)]
    one = torch.zeros(25 * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(np.stack(distinct)).cuda(), one, 25)
    ctx.frame_expand_border(g, one, 25)
    ctx.frame_init_lowres(g, one, 25)
    fenc = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    for t in range(gop_len):
        for gop in range(n_gops):
            k, j = t * n_gops + gop, (3 * gop + t) % 25
            fenc[k * g.slot_bytes:(k + 1) * g.slot_bytes] = one[j * g.slot_bytes:(j + 1) * g.slot_bytes]
    torch.cuda.synchronize()                           # as in pframe_measure
    del one
    b = np.arange(n_gops, n, dtype=np.int32)
    d_lmv = torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, fenc, b, b - n_gops, np.zeros(b.size, np.uint8), d_lmv[n_gops:], d_lc[n_gops:], d_ls[n_gops:])
    shapes = {"mb_type": ((n, nmb), np.int8), "partition": ((n, nmb), np.uint8), "mv8": ((n, nmb, 4, 2), np.int16),
              "mvr": ((n, nmb, 2), np.int16), "mvd8": ((n, nmb, 4, 2), np.int16), "nnz": ((n, nmb, 27), np.uint8), "cbp": ((n, nmb), np.int16),
              "mode16": ((n_gops, nmb), np.uint8), "chroma_mode": ((n_gops, nmb), np.uint8), "modes4": ((n_gops, nmb, 16), np.uint8),
              "luma_dc": ((n_gops, nmb, 16), np.int16)}
    dev = {k: torch.zeros(s, dtype=getattr(torch, np.dtype(t).name), device="cuda") for k, (s, t) in shapes.items()}
    dev["levels"] = torch.zeros((n, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
    recon = torch.zeros_like(fenc)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.gops_encode(g, fenc, recon, n_gops, gop_len, prm, d_lmv, dev)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(2):
        ctx.gops_encode(g, fenc, recon, n_gops, gop_len, prm, d_lmv, dev)
    ev1.record(stream)
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / 2
    del fenc, recon, dev, d_lmv, d_lc, d_ls
    torch.cuda.empty_cache()
    pics = ctx.pinned_empty((n, w * h * 3 // 2), np.uint8)
    for gop in range(n_gops):
        for t in range(gop_len):
            pics[gop * gop_len + t] = distinct[(3 * gop + t) % 25]
    out = {k: ctx.pinned_empty(s, t) for k, (s, t) in shapes.items()}
    packed = ctx.pinned_empty((n * nmb * pkg.RES_LEVELS_PER_MB // 3,), np.int16)
    f_off, f_size = np.zeros(n, np.int64), np.zeros(n, np.int32)
    mb_off = ctx.pinned_empty((n, nmb), np.int32)
    run = lambda: ctx.gops_encode_host(w, h, n_gops, gop_len, pics, prm, out, packed, f_off, f_size, mb_off)
    run()
    t0 = time.perf_counter()
    for _ in range(2):
        run()
    dt = (time.perf_counter() - t0) / 2
    d2h = int(sum(a.nbytes for a in out.values()) + mb_off.nbytes + 2 * int(f_size.sum()))
    return dev_ms, dt, int(pics.nbytes), d2h


def pframe_oracle_check(g, check):
    """frame 0 of the P-frame measurement against the CPU oracle's xo_p_frame (pinned to the running reference encoder)"""
    import cpu_checkers as cc
    from cpu_checkers import ptr
    slots, lmv, got, recon, (me, subme, qp, part) = check
    o = cc.oracle()
    go = cc.oracle_geom(g.width, g.height)
    nmb = g.mb_count

    class P(C.Structure):
        _fields_ = [(n, C.c_int32) for n in ("me_method", "subpel_refine", "me_range", "qp", "mv_range", "fast_pskip", "mvc_scale", "analyse_inter")]
    want = {"mb_type": np.zeros(nmb, np.int8), "mv": np.zeros((nmb, 2), np.int16), "mvr": np.zeros((nmb, 2), np.int16),
            "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
    wrecon = np.zeros(g.slot_bytes, np.uint8)
    prm = P(me, subme, 16, qp, 512, 1, 0, part)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    if part:
        want["mv"] = np.zeros((nmb, 4, 2), np.int16)
        want["partition"] = np.zeros(nmb, np.uint8)
        mvd8 = np.zeros((nmb, 4, 2), np.int16)
        o.xo_p_frame_part(C.byref(go), ptr(slots[g.slot_bytes:]), ptr(slots[: g.slot_bytes]), ptr(wrecon), C.byref(prm), vp(lmv), None,
                          vp(want["mb_type"]), vp(want["partition"]), vp(want["mv"]), vp(want["mvr"]), vp(mvd8), vp(want["levels"]),
                          vp(want["nnz"]), vp(want["cbp"]))
    else:
        o.xo_p_frame(C.byref(go), ptr(slots[g.slot_bytes:]), ptr(slots[: g.slot_bytes]), ptr(wrecon), C.byref(prm), vp(lmv), None,
                     vp(want["mb_type"]), vp(want["mv"]), vp(want["mvr"]), None, vp(want["levels"]), vp(want["nnz"]), vp(want["cbp"]))
    ok = all(np.array_equal(got[k], want[k]) for k in want)
    lo = g.luma_origin
    a = recon[lo:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
    bb = wrecon[lo:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
    return bool(ok and np.array_equal(a, bb))


def cpu_pframe_reference(which, w, h, threads, me, subme, qp=26, n_frames=5, psub=0):
    """the reference's own x264_macroblock_analyse + x264_macroblock_encode on P slices: every host thread encodes its own
    1080p clip with its own encoder instance (unmodified reference, oracle/_ref); the doors in front of the two functions
    keep the time each thread spends inside them on P slices.  Returns (P frames/s of that path summed over the threads,
    whole-encoder frames/s by wall clock, share of the encoder's time the path takes)"""
    import cpu_checkers as cc
    from cpu_checkers import ptr
    lib = cc.ref_o3() if which == "O3" else cc.ref()
    if lib is None or not hasattr(lib, "xref_open_ex"):
        return None
    lib.xref_open_ex.restype = C.c_void_p
    clips = [np.concatenate([cc.synth_frame(w, h, 3 * t + i) for i in range(n_frames)]) for t in range(min(threads, 4))]
    encs = [C.c_void_p(lib.xref_open_ex(w, h, me, subme, 16, qp, psub, 1)) for _ in range(threads)]
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    rates, secs = [0.0] * threads, [0.0] * threads
    lib.xref_set_door_timing(1)

    def work(t):
        out = np.zeros(8 << 20, np.uint8)
        t0 = time.perf_counter()
        size = lib.xref_encode_clip(encs[t], ptr(clips[t % len(clips)]), n_frames, ptr(out), out.size)
        secs[t] = time.perf_counter() - t0
        acc = (C.c_double * 3)()
        lib.xref_door_seconds_read(acc)
        assert size > 0 and acc[2] > 0
        rates[t] = (acc[2] / mbs) / (acc[0] + acc[1])          # P frames per second of analyse + encode on this thread
        secs[t] = (acc[0] + acc[1]) / secs[t]
    try:
        dt = _run_threads(work, threads)
    finally:
        lib.xref_set_door_timing(0)
    return sum(rates), threads * n_frames / dt, float(np.mean(secs))


def cli_measure(pkg, w, h, frames=48, lead=4):
    """the reference's own CLI against the same CLI with the library behind its drivers (glue/_build/x264ref_gpu: every source
    of the reference compiled unmodified + glue/*.c; lowres, lookahead, in-loop filter and the macroblock loop of every slice on
    the device, the entropy coder on the host): wall-clock frames/s of an encode, start-up taken out by differencing a short
    and a long clip.  Returns None when a binary is missing."""
    import subprocess
    import tempfile
    import cpu_checkers as cc
    gpu_cli = os.path.join(ROOT, "glue", "_build", "x264ref_gpu")
    if not (os.path.exists(gpu_cli) and os.path.exists(cc.REF_CLI)):
        return None
    out = {}
    with tempfile.TemporaryDirectory() as td:
        clip = np.concatenate([cc.synth_frame(w, h, i) for i in range(lead + frames)])
        srcs = {}
        for tag, n in (("short", lead), ("long", lead + frames)):
            d = os.path.join(td, tag)
            os.makedirs(d)
            srcs[tag] = os.path.join(d, f"syn_{w}x{h}.yuv")
            clip[: n * w * h * 3 // 2].tofile(srcs[tag])
        streams = {}
        for name, exe in (("reference", cc.REF_CLI), ("gpu", gpu_cli)):
            t = {}
            for tag in ("short", "long"):
                dst = os.path.join(td, f"{name}_{tag}.264")
                t[tag] = 1e30
                for _ in range(3):
                    # The GPU binary's start-up is CUDA context creation, 0.8 .. 2.3 s from run to run -- more than the encode
                    # itself, so differencing wall clocks does not remove it.  The glue keeps its own clock (X264DSP_GLUE_STATS):
                    # time since the doors were installed (before main) minus the time spent creating the context.
                    env = dict(os.environ)
                    stats = os.path.join(td, "stats.json")
                    if name == "gpu":
                        env["X264DSP_GLUE_STATS"] = stats
                    t0 = time.perf_counter()
                    subprocess.run([exe, srcs[tag], dst], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
                    wall = time.perf_counter() - t0
                    if name == "gpu":
                        sec = json.load(open(stats))["seconds"]
                        wall = sec["since_install"] - sec["open"]
                    t[tag] = min(t[tag], wall)
                if tag == "long":
                    streams[name] = np.fromfile(dst, np.uint8)
            out[name] = {"frames_per_s": frames / max(t["long"] - t["short"], 1e-9), "wall_s_long_clip": t["long"],
                         "wall_s_short_clip": t["short"],
                         "clock": "wall clock of the process" if name == "reference" else
                                  "the glue's own clock from before main() to exit, minus CUDA context creation (0.8 .. 2.3 s, varies)"}
        out["bitstreams_identical"] = bool(np.array_equal(streams["reference"], streams["gpu"]))
        out["bitstream_bytes"] = int(streams["reference"].size)
    out["frames"] = frames
    out["speedup"] = out["gpu"]["frames_per_s"] / out["reference"]["frames_per_s"]
    return out


def me_search_e2e(pkg, ctx, g, w, h, luma_host, la_mvs, pairs, reps=3, qp=26):
    """configs[2] end to end through x264dsp_me_search_frames_host: pinned host pictures and block lists in, pinned
    results out.  Returns (seconds per call, h2d bytes, d2h bytes, results of the last call by size)"""
    prm = pkg.MeParams(pkg.ME_HEX, 5, 16, qp, 1)
    luma = ctx.pinned_empty((pairs + 1, w * h), np.uint8)
    luma[:] = luma_host[: pairs + 1]
    blocks, results = [], []
    for size in range(len(ME_SIZES)):
        b = np.concatenate([pkg.tiling_blocks(g, size, la_mvs[p + 1]) for p in range(pairs)])
        pb = ctx.pinned_empty((len(b),), pkg.ME_BLOCK_DTYPE)
        pb[:] = b
        blocks.append(pb)
        results.append(ctx.pinned_empty((len(b),), pkg.ME_RESULT_DTYPE))
    sizes = list(range(len(ME_SIZES)))
    ctx.me_search_frames_host(w, h, luma, prm, sizes, blocks, results)          # warm-up (allocations)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.me_search_frames_host(w, h, luma, prm, sizes, blocks, results)
    dt = (time.perf_counter() - t0) / reps
    h2d = luma.nbytes + sum(b.nbytes for b in blocks)
    d2h = sum(r.nbytes for r in results)
    return dt, h2d, d2h, results


def recon_e2e(pkg, ctx, g, w, h, n_frames, reps=3, qp=26):
    """configs[3] end to end through x264dsp_recon_frames_host (pinned buffers).  Same inputs as recon_measure."""
    nmb = g.mb_count
    rng = np.random.RandomState(5)
    i420 = ctx.pinned_empty((n_frames + 1, w * h * 3 // 2), np.uint8)
    for i in range(n_frames + 1):
        i420[i] = pkg.synth_frame(w, h, i)
    mv = ctx.pinned_empty((n_frames, nmb, 2), np.int16)
    mv[:] = (np.array([12, 8]) + rng.randint(-1, 2, (n_frames, nmb, 2))).astype(np.int16)
    bs = ctx.pinned_empty((n_frames, nmb, 64), np.uint8)
    bs[:] = ((rng.rand(n_frames, nmb, 2, 8, 4) < 0.35).astype(np.uint8)
             * rng.randint(1, 3, (n_frames, nmb, 2, 8, 4)).astype(np.uint8)).reshape(n_frames, nmb, 64)
    mb_type = ctx.pinned_empty((n_frames, nmb), np.int8)
    mb_type[:] = 4
    part = ctx.pinned_empty((n_frames, nmb), np.uint8)
    part[:] = 16
    out = (ctx.pinned_empty((n_frames, nmb, pkg.RES_LEVELS_PER_MB), np.int16), ctx.pinned_empty((n_frames, nmb, pkg.RES_NNZ_PER_MB), np.uint8),
           ctx.pinned_empty((n_frames, nmb), np.int16), ctx.pinned_empty((n_frames, w * h * 3 // 2), np.uint8))
    ctx.recon_frames_host(w, h, i420, mv, qp, mb_type, part, bs, out)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.recon_frames_host(w, h, i420, mv, qp, mb_type, part, bs, out)
    dt = (time.perf_counter() - t0) / reps
    h2d = i420.nbytes + mv.nbytes + bs.nbytes + mb_type.nbytes + part.nbytes
    d2h = sum(o.nbytes for o in out)
    return dt, h2d, d2h, out


def lookahead_oracle_check(pkg, w, h, clip_len, luma_clip, mvs, costs, sums):
    """the headline step's first clip through the CPU oracle (one core, a fraction of a second): True when the MVs, the
    block costs and the frame sums the GPU returned are bit-exact"""
    import cpu_checkers as cc
    o = cc.oracle()
    g = cc.oracle_geom(w, h)
    slots = [np.zeros(g.slot_bytes, np.uint8) for _ in range(clip_len)]
    chroma = np.full(w * h // 2, 128, np.uint8)
    for i in range(clip_len):
        pic = np.concatenate([luma_clip[i], chroma])
        o.xo_frame_load_i420(C.byref(g), cc.ptr(pic), cc.ptr(slots[i]))
        o.xo_frame_init_lowres(C.byref(g), cc.ptr(slots[i]))
    ok = True
    for i in range(clip_len):
        mv_o, c_o, s_o = np.zeros((g.mb_count, 2), np.int16), np.zeros(g.mb_count, np.int32), np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(g), cc.ptr(slots[i]), cc.ptr(slots[i - 1]) if i else None, 1,
                                  cc.ptr(mv_o, cc.i16p), cc.ptr(c_o, cc.i32p), cc.ptr(s_o, cc.i32p), None)
        if i:
            ok = ok and np.array_equal(mvs[i], mv_o) and np.array_equal(costs[i], c_o)
        ok = ok and np.array_equal(sums[i][:5], s_o[:5])
    return bool(ok)


def int_pipe_roofline(sad_px, satd_px, seconds, kernels):
    """algorithmic integer instructions (SAD 0.25 per pixel comparison = one VABSDIFF4.U8.ACC per 4 pixels,
    SATD 3.5 packed instructions per pixel) / time / the measured ALU-pipe peak (tools/int_pipe_peak.cu on this pool's
    B200: 64 thread-instructions per clock per SM for VABSDIFF4 / IADD3 / LOP3 / PRMT / VIADD.16x2 / IDP.4A)"""
    path = os.path.join(ROOT, "profiles", "int_pipe_peak.json")
    per_clk_sm, sms, khz = 64.0, 148, 1965000
    src = "fallback: 64 thread-inst/clk/SM x 148 SMs x 1965 MHz"
    if os.path.exists(path):
        d = json.load(open(path))
        per_clk_sm = next(o["thread_inst_per_clk_per_sm"] for o in d["ops"] if o["op"].startswith("VABSDIFF4.U8.ACC"))
        sms, khz = d["sms"], d["max_clock_khz"]
        src = "measured (profiles/int_pipe_peak.json)"
    peak = per_clk_sm * sms * khz * 1e3
    inst = 0.25 * sad_px + 3.5 * satd_px
    return {"kernels": kernels, "algorithmic_thread_inst": inst, "achieved_tinst_per_s": inst / seconds / 1e12,
            "peak_tinst_per_s": peak / 1e12, "frac": inst / seconds / peak, "peak_source": src,
            "per_pixel": {"sad": 0.25, "satd": 3.5}}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own C path (oracle/_ref) or, if that build is absent, the oracle port

def cpu_lookahead(w, h, clip_len, luma_clips, threads, repeats, lib="O2"):
    """runs the lookahead pass of len(luma_clips) clips on `threads` host threads, one clip at a
    time per thread; returns (seconds of the timed region, kind, frames processed).  lib: "O2" = oracle/_ref as the
    reference's recipe builds it, "O3" = the same sources at -O3 -march=x86-64-v3."""
    import cpu_checkers as cc
    have = os.path.exists(cc.REF_SO) or os.path.isdir(cc.REFERENCE_TREE)
    lib = (cc.ref_o3() if lib == "O3" else cc.ref()) if have else None
    kind = "reference" if lib is not None else "port"
    n_clips = len(luma_clips)
    chroma = np.full(w * h // 2, 128, np.uint8)
    # I420 pictures are assembled once, outside the timed region (distinct clips only: the lists may repeat)
    pics = {}
    for clip in luma_clips:
        key = clip.__array_interface__["data"][0]
        if key not in pics:
            pics[key] = [np.concatenate([clip[i], chroma]) for i in range(clip_len)]
    work = [[] for _ in range(threads)]
    for c in range(n_clips):
        work[c % threads].append(c)
    state = []
    if kind == "reference":
        for t in range(threads):
            if not work[t]:
                state.append(None)
                continue
            enc = cc.RefEncoder(w, h, lib=lib)
            frames = [enc.new_frame(False) for _ in range(clip_len)]
            state.append((enc, frames))
    else:
        o = cc.oracle()
        g = cc.oracle_geom(w, h)
        for t in range(threads):
            state.append([np.zeros(g.slot_bytes, np.uint8) for _ in range(clip_len)] if work[t] else None)

    barrier = threading.Barrier(threads + 1)
    done = threading.Barrier(threads + 1)

    def run(t):
        barrier.wait()
        if state[t] is not None:
            for _ in range(repeats):
                for c in work[t]:
                    clip_pics = pics[luma_clips[c].__array_interface__["data"][0]]
                    if kind == "reference":
                        enc, frames = state[t]
                        for i in range(clip_len):
                            # luma part of x264_frame_copy_picture + mod-16 padding: what the GPU arm's e2e uploads
                            enc.lib.xref_frame_load_luma(enc.h, frames[i], cc.ptr(clip_pics[i]))
                        arr = (C.c_void_p * clip_len)(*[f.value for f in frames])
                        costs = (C.c_int * clip_len)()
                        enc.lib.xref_time_lookahead(enc.h, arr, clip_len, costs)
                    else:
                        slots = state[t]
                        for i in range(clip_len):
                            o.xo_frame_load_i420(C.byref(g), cc.ptr(clip_pics[i]), cc.ptr(slots[i]))
                            o.xo_frame_init_lowres(C.byref(g), cc.ptr(slots[i]))
                        mv = np.zeros((g.mb_count, 2), np.int16)
                        cs = np.zeros(g.mb_count, np.int32)
                        sm = np.zeros(8, np.int32)
                        for i in range(clip_len):
                            o.xo_lookahead_frame_cost(C.byref(g), cc.ptr(slots[i]), cc.ptr(slots[i - 1]) if i else None, 1,
                                                      cc.ptr(mv, cc.i16p), cc.ptr(cs, cc.i32p), cc.ptr(sm, cc.i32p), None)
        done.wait()

    ths = [threading.Thread(target=run, args=(t,), daemon=True) for t in range(threads)]
    for th in ths:
        th.start()
    barrier.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for th in ths:
        th.join()
    return dt, kind, n_clips * clip_len * repeats


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """the reference's own C path (oracle/_ref, unmodified sources) on all host threads; nothing of the product is loaded:
    the synthetic pictures come from oracle/_build/libx264dsp_synth.so (the generator source every consumer shares)"""
    import cpu_checkers as cc
    w, h = args.width, args.height
    threads = args.cpu_threads or host_cores()
    # four clips per thread and step: a bounded sample of the GPU arm's step (args.clips clips per GPU)
    n_clips = threads * 4
    distinct = min(n_clips, 8)
    base = np.stack([cc.synth_frame(w, h, i, -1, True) for i in range(distinct * args.clip_len)])
    luma_clips = [base.reshape(-1, args.clip_len, w * h)[c % distinct] for c in range(n_clips)]
    for _ in range(max(args.warmup, 1)):
        cpu_lookahead(w, h, args.clip_len, luma_clips[: threads], threads, 1)
    times = []
    kind = "reference"
    for _ in range(args.steps):
        dt, kind, frames = cpu_lookahead(w, h, args.clip_len, luma_clips, threads, 1)
        times.append(dt)
    total = sum(times)
    frames_per_step = n_clips * args.clip_len
    value = frames_per_step * args.steps / total
    o3 = None
    if cc.ref_o3() is not None:
        cpu_lookahead(w, h, args.clip_len, luma_clips[: threads], threads, 1, lib="O3")
        dt3, _, frames3 = cpu_lookahead(w, h, args.clip_len, luma_clips, threads, max(1, min(args.steps, 5)), lib="O3")
        o3 = frames3 / dt3
    line = {
        "impl": "reference",
        "metric": "1080p ME frames/sec" if (w, h) == (1920, 1080) else f"{w}x{h} ME frames/sec",
        "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, frames_per_step),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{n_clips} clips x {args.clip_len} frames per step x {args.steps} steps, four clips per host "
                                   f"thread (a bounded sample of the GPU arm's {args.clips} clips per step; throughput metric)",
                         "build": "-O2, the reference's recipe (oracle/Makefile)",
                         "value_O3_x86-64-v3": o3,
                         "timed": "luma plane copy + mod-16 padding, x264_frame_init_lowres, x264_slicetype_frame_cost per frame"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, frames_per_step, sample=None):
    cfg = {
        "workload": (f"{args.width}x{args.height} lowres lookahead: x264_frame_init_lowres + "
                     f"x264_slicetype_frame_cost (DIA/SAD full-pel, hpel refine, SATD) over {args.clip_len}-frame clips"),
        "clips_per_gpu_per_step": args.clips, "clip_len": args.clip_len, "frames_per_step_per_gpu": args.clips * args.clip_len,
        "l2_policy": "step input larger than L2 (luma %.0f MB per GPU per step)" % (args.clips * args.clip_len * args.width * args.height / 1e6),
        "step": ("per frame: the four half-resolution planes from the w x h luma picture (the picture is read as "
                 "frame->plane[0], mod-16 replication included), then intra cost of every frame and inter cost of "
                 "frames 1.. against their predecessor; both arms return lowres MVs, MV costs and frame costs"),
        "parallelism": f"frame-range sharding, {args.gpus} GPU(s), no data-path collective",
    }
    return cfg


# --------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: keep this process (and the pinned buffers it first-touches) on the CPU cores NVML
    reports as local to its GPU, so that eight ranks' host->device copies do not cross sockets.  Best effort:
    returns a short description, or None when NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cores &= allowed
        if cores and cores != allowed:
            os.sched_setaffinity(0, cores)
            return f"{len(cores)} of {len(allowed)} cores (NVML affinity of GPU {gpu_index})"
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: ONE sequence split by frame range over the ranks

def run_sharded(args, pkg, rank, world, local_rank):
    """One 3840x2160 (--width/--height) sequence of --seq-frames frames: every rank takes the contiguous range
    x264dsp_frame_range gives it, uploads it plus the one overlap frame (the reference of its first frame), runs the lowres
    lookahead and the full-resolution 16x16 + 8x8 motion search (HEX, subme 5; mvp from the lookahead) of its frames and
    returns the results to its host; the per-frame cost tables are then all-gathered with NCCL and rank 0 compares the
    stitched tables bit for bit with its own single-GPU pass over the whole sequence.  Strong scaling: the sequence is fixed."""
    import torch
    import torch.distributed as dist
    w, h, n_seq = args.width, args.height, args.seq_frames
    torch.cuda.set_device(local_rank)
    if world > 1:
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    ctx = pkg.Context(local_rank)
    g = pkg.geometry(w, h)
    mbc = g.mb_count
    prm = pkg.MeParams(pkg.ME_HEX, 5, 16, 26, 1)
    sizes = (0, 3)                                            # 16x16 and 8x8: the two searches the reference's P analysis always runs
    chunk = 8                                                 # frame pairs per launch group

    def analyse(first, count, need_prev, timed):
        """frames [first, first+count) of the sequence -> (la_mvs, la_costs, la_sums, me results by size) for exactly those
        frames (frame 0 of the sequence: intra-only lookahead, no search); returns also the seconds of the second,
        timed pass (pinned host pictures in, host results out)"""
        lo = first - 1 if need_prev else first
        nf = first + count - lo
        luma = ctx.pinned_empty((nf, w * h), np.uint8)
        for i in range(nf):
            luma[i] = pkg.synth_frame(w, h, lo + i, -1, True)
        la_m = ctx.pinned_empty((nf, mbc, 2), np.int16)
        la_c = ctx.pinned_empty((nf, mbc), np.int32)
        la_s = ctx.pinned_empty((nf, pkg.LA_SUMS), np.int32)
        n_pairs = nf - 1

        def lookahead():
            # the whole range as one clip: frame lo intra-only, every other frame against its predecessor
            ctx.lookahead_clips_host(w, h, 1, nf, luma, la_m, la_c, la_s)
        lookahead()
        blocks, results = [], []
        for sz in sizes:
            b = np.concatenate([pkg.tiling_blocks(g, sz, la_m[p + 1]) for p in range(n_pairs)]) if n_pairs else np.zeros(0, pkg.ME_BLOCK_DTYPE)
            pb = ctx.pinned_empty((max(len(b), 1),), pkg.ME_BLOCK_DTYPE)[: len(b)]
            pb[:] = b
            blocks.append(pb)
            results.append(ctx.pinned_empty((max(len(b), 1),), pkg.ME_RESULT_DTYPE)[: len(b)])

        def search():
            for p0 in range(0, n_pairs, chunk):
                k = min(chunk, n_pairs - p0)
                nb = [len(b) // n_pairs for b in blocks]
                ctx.me_search_frames_host(w, h, luma[p0: p0 + k + 1], prm, list(sizes),
                                          [b[p0 * n: (p0 + k) * n] for b, n in zip(blocks, nb)],
                                          [r[p0 * n: (p0 + k) * n] for r, n in zip(results, nb)])
        if n_pairs:
            search()
        dt = 0.0
        if timed:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            lookahead()
            if n_pairs:
                search()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        skip = 1 if need_prev else 0                          # the overlap frame's own (intra-only) result is not ours
        nb = [len(b) // max(n_pairs, 1) for b in blocks]
        me = []
        for r, n in zip(results, nb):
            arr = np.zeros((count, n), pkg.ME_RESULT_DTYPE)   # row i = sequence frame first + i; frame 0 has no search
            got = r.reshape(n_pairs, n) if n_pairs else r.reshape(0, n)
            arr[count - n_pairs:] = got
            me.append(arr)
        return la_m[skip:].copy(), la_c[skip:].copy(), la_s[skip:].copy(), me, dt

    first, count, need_prev = pkg.frame_range(n_seq, rank, world)
    launches0 = ctx.launches
    mvs, costs, sums, me, dt = analyse(first, count, need_prev, True)
    launches = ctx.launches - launches0

    # ---- gather: every rank's tables, padded to the largest range, with NCCL
    t_gather = 0.0
    full = None
    if world > 1:
        cmax = (n_seq + world - 1) // world

        def gather(a):
            pad = np.zeros((cmax,) + a.shape[1:], a.dtype)
            pad[: len(a)] = a
            mine = torch.from_numpy(pad.view(np.uint8).reshape(cmax, -1)).cuda()
            out = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(out, mine)
            parts = []
            for r in range(world):
                _, c, _ = pkg.frame_range(n_seq, r, world)
                parts.append(out[r][:c].cpu().numpy())
            return np.concatenate(parts).view(a.dtype).reshape((n_seq,) + a.shape[1:])
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        full = [gather(mvs), gather(costs), gather(sums)] + [gather(m) for m in me]
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - t0
    else:
        full = [mvs, costs, sums] + me
    tt = torch.tensor([dt, t_gather], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt, t_gather = float(tt[0]), float(tt[1])

    if rank == 0:
        same = None
        if world > 1:
            ref = analyse(0, n_seq, False, False)             # rank 0 alone over the whole sequence
            want = [ref[0], ref[1], ref[2]] + ref[3]
            same = all(np.array_equal(a, b) for k, (a, b) in enumerate(zip(full, want)) if k != 2)
            same = bool(same and np.array_equal(full[2][:, :5], want[2][:, :5]))     # sums: costs / counts (5.. are spare)
        line = {
            "metric": f"{w}x{h} ME + lookahead frames/sec (one {n_seq}-frame sequence, frame-range sharded)",
            "value": n_seq / dt, "unit": "frames/s", "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[4]: {w}x{h}, {n_seq} frames: lowres lookahead of every frame + 16x16 and 8x8 "
                                   "x264_me_search_ref (HEX, subme 5, qpel refine) of every P frame, split into contiguous frame ranges "
                                   "with a one-frame overlap (x264dsp_frame_range)",
                       "frames_per_rank": count, "timed": "pinned host pictures -> results in host memory, per rank; max over ranks",
                       "parallelism": f"frame-range sharding over {world} GPU(s); NCCL all-gather of the cost tables after the timed region"},
            "e2e": {"value": n_seq / dt, "unit": "frames/s", "h2d_bytes_per_step": int((count + (1 if need_prev else 0)) * w * h),
                    "d2h_bytes_per_step": int(mvs.nbytes + costs.nbytes + sums.nbytes + sum(m.nbytes for m in me))},
            "gather_ms": t_gather * 1e3, "gather": "torch.distributed all_gather (NCCL) of lookahead MVs / costs / sums and the ME results",
            "sharded_equals_single_gpu": same, "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and args.impl != "reference" else None
    pkg = load_package()
    sampler = ClockSampler(local_rank)
    if rank == 0 and args.impl != "reference":
        sampler.start()                    # nvidia-smi takes a while to produce its first line: start it early

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args)
        return 0
    if args.config == 5:
        if (args.width, args.height) == (1920, 1080):
            args.width, args.height = 3840, 2160
        sampler.stop()
        return run_sharded(args, pkg, rank, world, local_rank)

    import torch
    import torch.distributed as dist
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created (NCCL_DEBUG=VERSION on this image);
        # stdout carries exactly one JSON line, so the banner goes to stderr
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    torch.cuda.set_device(local_rank)
    ctx = pkg.Context(local_rank)
    w, h, clips, clip_len = args.width, args.height, args.clips, args.clip_len
    g = pkg.geometry(w, h)
    n = clips * clip_len
    mbc = g.mb_count

    # ---- synthetic input in pinned host memory (each rank gets its own clips)
    luma_host = ctx.pinned_empty((n, w * h), np.uint8)
    make_clips(pkg, w, h, clips, clip_len, first_clip=rank * clips, out=luma_host)
    mvs_host = ctx.pinned_empty((n, mbc, 2), np.int16)
    costs_host = ctx.pinned_empty((n, mbc), np.int32)
    sums_host = ctx.pinned_empty((n, pkg.LA_SUMS), np.int32)

    # ---- device-resident arm: raw luma already in HBM when the timed region starts
    stream = ctx.torch_stream()
    luma_dev = torch.from_numpy(luma_host).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    d_mvs = torch.zeros((n, mbc, 2), dtype=torch.int16, device="cuda")
    d_costs = torch.zeros((n, mbc), dtype=torch.int32, device="cuda")
    d_sums = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    b = np.arange(n, dtype=np.int32)
    p0 = np.where(b % clip_len == 0, -1, b - 1).astype(np.int32)
    wi = np.ones(n, np.uint8)
    torch.cuda.synchronize()

    def step_dev():
        ctx.frame_lowres_from_luma(g, luma_dev, slots, n)          # x264_frame_init_lowres, the picture being plane[0]
        ctx.lookahead_frame_cost(g, slots, b, p0, wi, d_mvs, d_costs, d_sums)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sync_all()
    ctx.profile_enable(True)
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_region0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step_dev()
    ev1.record(stream)
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    inter_ms, inter_n = ctx.profile_read(pkg.PROF_LA_INTER)
    prof = {name: ctx.profile_read(kind) for name, kind in (
        ("load", pkg.PROF_LOAD), ("lowres", pkg.PROF_LOWRES), ("border", pkg.PROF_BORDER),
        ("la_intra", pkg.PROF_LA_INTRA), ("la_inter", pkg.PROF_LA_INTER))}
    ctx.profile_enable(False)
    sums_np = d_sums.cpu().numpy()

    # ---- end-to-end arm: host buffers in, host results out, through the C ABI
    for _ in range(2):
        ctx.lookahead_clips_host(w, h, clips, clip_len, luma_host, mvs_host, costs_host, sums_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.lookahead_clips_host(w, h, clips, clip_len, luma_host, mvs_host, costs_host, sums_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(sums_host, sums_np), "host and device arms disagree"
    # ---- copy-only probe: the same call with its kernels switched off -- same pinned buffers, same streams, same copy
    # order on every rank at the same time -- i.e. what the host's DMA side alone allows at this rank count
    ctx.debug_copies_only(True)
    mv_keep, cost_keep, sum_keep = mvs_host.copy(), costs_host.copy(), sums_host.copy()
    ctx.lookahead_clips_host(w, h, clips, clip_len, luma_host, mvs_host, costs_host, sums_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.lookahead_clips_host(w, h, clips, clip_len, luma_host, mvs_host, costs_host, sums_host)
    torch.cuda.synchronize()
    copy_s = time.perf_counter() - t0
    ctx.debug_copies_only(False)
    mvs_host[:], costs_host[:], sums_host[:] = mv_keep, cost_keep, sum_keep
    del mv_keep, cost_keep, sum_keep
    clocks = sampler.stop(t_region0, time.time()) if rank == 0 else None

    # ---- single clip latency (exactly the 8-frame configuration), device resident
    one = clip_len
    def step_one():
        ctx.frame_lowres_from_luma(g, luma_dev, slots, one)
        ctx.lookahead_frame_cost(g, slots, b[:one], p0[:one], wi[:one], d_mvs, d_costs, d_sums)
    for _ in range(3):
        step_one()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step_one()
    ev1.record(stream)
    torch.cuda.synchronize()
    one_ms = ev0.elapsed_time(ev1) / args.steps

    # ---- configs[2] and configs[3] on EVERY rank (each on its own frames): device-resident launches and the
    # host-memory doors; times are max-reduced over the ranks below
    me_ms = me_blocks0 = me_slots = rc_ms = rc_check = None
    me_e2e = rc_e2e = None
    me_pairs = min(8, clip_len - 1)
    rc_frames = 96 if w * h <= 1920 * 1088 else 16
    sec = [0.0, 0.0, 0.0, 0.0]                 # ME device ms per frame, ME e2e s per call, recon device ms per frame, recon e2e s
    if not args.no_me:
        la_mvs_host = d_mvs[: clip_len].cpu().numpy()
        me_ms, me_blocks0, me_slots = me_search_measure(pkg, ctx, torch, g, luma_dev, la_mvs_host, me_pairs)
        me_e2e = me_search_e2e(pkg, ctx, g, w, h, luma_host, la_mvs_host, me_pairs)
        rc_ms, rc_check = recon_measure(pkg, ctx, torch, g, w, h, rc_frames, first_frame=rank * 7)
        rc_e2e = recon_e2e(pkg, ctx, g, w, h, rc_frames)
        sec = [sum(me_ms.values()), me_e2e[0], sum(rc_ms.values()), rc_e2e[0]]
    rc_ms_big = None
    if not args.no_me and world == 1 and rank == 0 and w * h <= 1920 * 1088:
        rc_ms_big, _ = recon_measure(pkg, ctx, torch, g, w, h, 384, reps=2)
    # ---- the closed loop of a P frame on the device, every rank on its own frames
    pf = None
    if not args.no_me and w * h <= 1920 * 1088:
        pf = {}
        for name, (pme, psub, ppart) in PF_SETTINGS:
            pf[name] = pframe_measure(pkg, ctx, torch, g, w, h, PF_FRAMES if not ppart else PF_FRAMES // 2, pme, psub, part=ppart)
        # several ranks share one host: half the frames per call and no dense-door comparison (3.6 GB of pinned memory per rank)
        pf_e2e_frames = PF_E2E_FRAMES if world == 1 else PF_E2E_FRAMES // 2
        pf_e2e = pframe_e2e(pkg, ctx, g, w, h, pf_e2e_frames, 0, 1, dense=world == 1)
        gops = gops_measure(pkg, ctx, torch, g, w, h, GOPS_N, GOPS_LEN, 0, 1) if world == 1 else None
        sec += [pf["dia_subme1"][0], pf["hex_subme5"][0], pf_e2e[0], pf["hex_subme5_psub16x16"][0]]
    else:
        sec += [0.0, 0.0, 0.0, 0.0]

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3] + sec + [copy_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    me_dev_ms_max, me_e2e_s_max, rc_dev_ms_max, rc_e2e_s_max = (float(x) for x in t[2:6])
    pf_ms_max = (float(t[6]), float(t[7]), float(t[9]))
    pf_e2e_s_max = float(t[8])
    copy_ms = float(t[10])

    if rank == 0:
        frames_total = world * n * args.steps
        value = frames_total / (dev_ms / 1e3)
        e2e_value = frames_total / (e2e_ms / 1e3)
        pairs_per_launch = clips * (clip_len - 1)
        bytes_per_pair = 5 * g.lowres_w * g.lowres_h + 8 * mbc
        inter_avg_s = inter_ms / max(inter_n, 1) / 1e3
        achieved = bytes_per_pair * pairs_per_launch / inter_avg_s / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("xd_la_multi_kernel_bytes_per_pair")
            if traffic is not None:
                traffic = traffic * pairs_per_launch
        sad_px = int(sums_np[:, pkg.LA_SAD_EVALS].sum()) * 64
        satd_px = int(sums_np[:, pkg.LA_SATD_EVALS].sum()) * 64
        step_s = dev_ms / 1e3 / args.steps
        line = {
            "metric": "1080p ME frames/sec" if (w, h) == (1920, 1080) else f"{w}x{h} ME frames/sec",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, n),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(n * w * h),
                    "d2h_bytes_per_step": int(mvs_host.nbytes + costs_host.nbytes + sums_host.nbytes),
                    "ms_per_step": e2e_ms / args.steps, "api": "x264dsp_lookahead_clips_host (pinned host buffers)",
                    # the same call with its kernels switched off (x264dsp_debug_copies_only): same pinned buffers, same
                    # 16 streams, same copy order, all ranks at once -- the DMA-only time of a step at this rank count
                    "copy_only": {"ms_per_step": copy_ms / args.steps,
                                  "h2d_gbs_per_rank": n * w * h / (copy_ms / args.steps / 1e3) / 1e9,
                                  "h2d_gbs_all_ranks": world * n * w * h / (copy_ms / args.steps / 1e3) / 1e9,
                                  "frames_per_s_if_copies_were_all": world * n / (copy_ms / args.steps / 1e3)},
                    "copy_share_of_step": copy_ms / e2e_ms,
                    "bound": ("host DMA: the copies alone take %.0f %% of the step (kernels are hidden behind them)"
                              % (100 * copy_ms / e2e_ms)) if copy_ms >= 0.85 * e2e_ms else "kernels + copies"},
            "gpu_launches": int(launches),
            "host_affinity": numa,
            "clocks": clocks,
            "bit_exact_vs_oracle": lookahead_oracle_check(pkg, w, h, clip_len, luma_host[:clip_len], mvs_host[:clip_len],
                                                          costs_host[:clip_len], sums_host[:clip_len]),
            "roofline": {"kernel": "xd_la_multi_kernel<4>", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(bytes_per_pair * pairs_per_launch),
                         "avg_launch_ms": inter_avg_s * 1e3, "launches_timed": inter_n,
                         "int_pipe": int_pipe_roofline(sad_px, satd_px, (prof["la_inter"][0] + prof["la_intra"][0]) / args.steps / 1e3,
                                                       "xd_la_multi_kernel<4> + xd_la_intra_kernel"),
                         "note": "dependent-search wavefront kernel (SURVEY 8(d) config 2): HBM % reported as required; the "
                                 "kernel is bound by instruction issue and the ALU pipe, not by DRAM (ncu summaries under "
                                 "profiles/, newest round first); DRAM traffic is below the algorithmic bytes because "
                                 "consecutive pairs share a reference frame in L2"},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "pixel_cmp": {"sad_gpix_per_s": sad_px * world / step_s / 1e9, "satd_gpix_per_s": satd_px * world / step_s / 1e9,
                          "sad_pix_per_step_per_gpu": sad_px, "satd_pix_per_step_per_gpu": satd_px,
                          "counted_by": "evaluations the reference issues on the same input (kernel work counters == oracle counters)"},
            "single_clip": {"frames": one, "ms": one_ms, "frames_per_s": one / (one_ms / 1e3),
                            "note": "exactly one 8-frame clip in flight: latency bound"},
        }
        if not args.no_cpu_baseline and world == 1:
            threads = args.cpu_threads or host_cores()
            sample_clips = threads
            luma_clips = [luma_host.reshape(clips, clip_len, w * h)[c % clips] for c in range(sample_clips)]
            cpu_lookahead(w, h, clip_len, luma_clips, threads, 1)          # warm-up (page-in, allocations)
            dt, kind, frames = cpu_lookahead(w, h, clip_len, luma_clips, threads, 24)
            line["cpu_baseline"] = {"value": frames / dt, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": f"{sample_clips} clips x {clip_len} frames, 24 passes, one clip per thread "
                                              f"({frames} frames in {dt:.2f} s)"}
        else:
            line["cpu_baseline"] = None
        threads = args.cpu_threads or host_cores()
        baseline_ok = not args.no_cpu_baseline and world == 1
        if me_ms is not None:
            tot_ms = me_dev_ms_max                          # ms per frame, all 7 sizes, slowest rank
            me = {"workload": f"{w}x{h} x264_me_search_ref + x264_me_refine_qpel, HEX, subme 5, range 16, QP 26, one block list "
                              "per partition size tiling the frame (7 sizes), mvp from the lowres MVs",
                  "value": world * 1e3 / tot_ms, "unit": "frames/s (all 7 sizes)", "n_gpus": world,
                  "frames_per_s_all_sizes": world * 1e3 / tot_ms, "ms_per_frame_by_size": me_ms,
                  "launch": f"x264dsp_me_search_sized_frames_dev, {me_pairs} frame pairs per launch, every rank on its own frames",
                  "e2e": {"value": world * me_pairs / me_e2e_s_max, "unit": "frames/s (all 7 sizes)",
                          "h2d_bytes_per_step": int(me_e2e[1]), "d2h_bytes_per_step": int(me_e2e[2]),
                          "api": "x264dsp_me_search_frames_host (pinned host pictures, block lists and results)",
                          "note": "a block description is 116 bytes: the lists of the 7 sizes are 18x the bytes of the pictures, "
                                  "so this door is bound by the host-to-device copy of the lists"}}
            if baseline_ok:
                cnt = me_search_cpu_counts(pkg, g, me_slots, me_blocks0)
                sad = sum(c["sad_pix"] for c in cnt.values())
                satd = sum(c["satd_pix"] for c in cnt.values())
                ok = all(c["bit_exact"] for c in cnt.values())
                e2e_ok = all(np.array_equal(me_e2e[3][i][: len(me_blocks0[nm][0])].view(np.uint8), me_blocks0[nm][1])
                             for i, nm in enumerate(ME_SIZES))
                me.update({"sad_gpix_per_s": sad / (tot_ms / 1e3) / 1e9, "satd_gpix_per_s": satd / (tot_ms / 1e3) / 1e9,
                           "sad_pix_per_frame": sad, "satd_pix_per_frame": satd,
                           "counted_by": "the CPU oracle's instrumented run of the same block lists (frame pair 0)",
                           "bit_exact_vs_oracle": ok, "e2e_equals_device": bool(e2e_ok),
                           "roofline": dict(int_pipe_roofline(sad, satd, tot_ms / 1e3, "xd_me_sized_kernel<W,H>, 7 sizes"), bound="int_pipe")})
                pics = [pkg.synth_frame(w, h, rank * clips * clip_len + i) for i in range(2)]
                blocks_by_size = {nm: me_blocks0[nm][0] for nm in ME_SIZES}
                cb = {"unit": "frames/s (all 7 sizes)", "cores": threads, "kind": "reference",
                      "sample": "the block lists of frame pair 0, every host thread a contiguous share of each list, repeated for ~4 s",
                      "port_1core": 1.0 / sum(c["cpu_s"] for c in cnt.values())}
                for which in ("O2", "O3"):
                    fps, res = cpu_me_reference(which, w, h, pics, blocks_by_size, threads)
                    if fps is not None:
                        cb["value" if which == "O2" else "value_O3_x86-64-v3"] = fps
                        same = all(np.array_equal(res[nm].view(np.uint8), me_blocks0[nm][1]) for nm in ME_SIZES)
                        cb["gpu_equals_reference" + ("" if which == "O2" else "_O3")] = bool(same)
                me["cpu_baseline"] = cb
            line["me_search"] = me
        if rc_ms is not None:
            tot = rc_dev_ms_max
            hb = (w, h) == (1920, 1080)
            rc = {"workload": f"{w}x{h} inter macroblocks: x264_mb_mc 16x16 + x264_macroblock_encode (4x4 DCT, quant, dequant, "
                              "IDCT, decimation, chroma DC) + x264_frame_deblock_row, QP 26, frame-batched launches",
                  "value": world * 1e3 / tot, "unit": "frames/s", "n_gpus": world,
                  "frames_per_s": world * 1e3 / tot, "ms_per_frame": rc_ms, "frames_per_launch": rc_frames,
                  "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "algorithmic_MB_per_frame": {"mc": 6.27, "residual": 16.5, "deblock": 6.8},
                               "frac": {"mc": 6.27e6 / (rc_ms["mc"] / 1e3) / 1e9 / peak,
                                        "residual": 16.5e6 / (rc_ms["residual"] / 1e3) / 1e9 / peak,
                                        "deblock": 6.8e6 / (rc_ms["deblock"] / 1e3) / 1e9 / peak}} if hb else None,
                  "e2e": {"value": world * rc_frames / rc_e2e_s_max, "unit": "frames/s",
                          "h2d_bytes_per_step": int(rc_e2e[1]), "d2h_bytes_per_step": int(rc_e2e[2]),
                          "api": "x264dsp_recon_frames_host (pinned pictures, MVs, bS in; levels, nnz, cbp, deblocked I420 out)"}}
            if rc_ms_big is not None and rc["roofline"]:
                rc["ms_per_frame_at_384_frames_per_launch"] = rc_ms_big
                rc["roofline"]["frac_at_384_frames_per_launch"] = {
                    "mc": 6.27e6 / (rc_ms_big["mc"] / 1e3) / 1e9 / peak, "residual": 16.5e6 / (rc_ms_big["residual"] / 1e3) / 1e9 / peak,
                    "deblock": 6.8e6 / (rc_ms_big["deblock"] / 1e3) / 1e9 / peak}
            if baseline_ok:
                ok, cpu_s, coded = recon_cpu_check(pkg, g, rc_check)
                rc.update({"bit_exact_vs_oracle": ok, "coded_mbs_frame0": coded})
                pics = [pkg.synth_frame(w, h, rank * 7 + i) for i in range(2)]
                cb = {"unit": "frames/s", "cores": threads, "kind": "reference", "port_1core": 1.0 / cpu_s,
                      "sample": "frame 0 of the measurement (x264_mb_mc + x264_macroblock_encode per macroblock, then "
                                "x264_frame_deblock_row), one frame per host thread and pass, repeated for ~4 s"}
                for which in ("O2", "O3"):
                    fps, res = cpu_recon_reference(which, w, h, pics, rc_check["mv"], rc_check["bs"], threads)
                    if fps is not None:
                        cb["value" if which == "O2" else "value_O3_x86-64-v3"] = fps
                        go = rc_check
                        gy = go["deblocked"][g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
                        gc = go["deblocked"][g.slot_chroma_off + g.chroma_origin:][: (g.luma_h // 2) * g.chroma_stride].reshape(
                            g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
                        same = (np.array_equal(res[0], go["levels"]) and np.array_equal(res[1], go["nnz"]) and np.array_equal(res[2], go["cbp"])
                                and np.array_equal(res[3], gy) and np.array_equal(res[4], gc))
                        cb["gpu_equals_reference" + ("" if which == "O2" else "_O3")] = bool(same)
                rc["cpu_baseline"] = cb
            line["recon"] = rc
        if pf is not None:
            pfl = {"workload": f"{w}x{h} P frames, the whole macroblock loop on the device: x264_macroblock_analyse (MV prediction, "
                               "fast / early P_SKIP probe, 16x16 search with the reference's candidate list, refine_qpel) + "
                               "x264_macroblock_encode (mc, residual, forced P_SKIP) for every macroblock as a wavefront; "
                               "QP 26, one reference frame, analyse.inter = 0 (the reference's default); the host keeps CABAC",
                   "launch": f"x264dsp_p_frames_dev, {PF_FRAMES} independent frames per launch (the rows of all frames share one ticket queue; the "
                             "more frames, the less of the time the rows wait for each other: 96 per launch cost 91 / 158 us per "
                             "frame, one frame alone is the wavefront's critical path), every rank on its own frames",
                   "unit": "frames/s", "n_gpus": world, "settings": {}}
            for k, (name, (pme, psub, ppart)) in enumerate(PF_SETTINGS):
                ms_f, ms_one, skipped, check = pf[name]
                ent = {"value": world * 1e3 / pf_ms_max[k], "ms_per_frame": pf_ms_max[k], "ms_one_frame_alone": ms_one,
                       "skipped_mb_share": skipped, "me_method": pme, "subme": psub,
                       "analyse_inter": "PSUB16x16 (P8x8 / P16x8 / P8x16 analysed as well: x264dsp_p_frames_part_dev)" if ppart else 0}
                if baseline_ok:
                    ent["bit_exact_vs_oracle"] = pframe_oracle_check(g, check)
                    cb = {"unit": "frames/s", "cores": threads, "kind": "reference",
                          "sample": "every host thread encodes its own 5-frame 1080p clip with its own instance of the unmodified "
                                    "reference; the doors in front of x264_macroblock_analyse and x264_macroblock_encode keep the "
                                    "time each thread spends inside them on P slices; value = sum over the threads of P frames per "
                                    "second of that path"}
                    for which in ("O2", "O3"):
                        r = cpu_pframe_reference(which, w, h, threads, pme, psub, psub=ppart)
                        if r is not None:
                            sfx = "" if which == "O2" else "_O3_x86-64-v3"
                            cb["value" + sfx] = r[0]
                            cb["whole_encoder_frames_per_s" + sfx] = r[1]
                            cb["path_share_of_encoder_time" + sfx] = r[2]
                    ent["cpu_baseline"] = cb
                pfl["settings"][name] = ent
            if baseline_ok:
                cli = cli_measure(pkg, w, h)
                if cli is not None:
                    pfl["encoder_cli"] = dict(cli, note="one stream, one process each: the reference's CLI (single-threaded by "
                                              "construction) against glue/_build/x264ref_gpu, whose host side is the same code "
                                              "minus the doors' work plus synchronous copies per frame; CABAC, rate "
                                              "control and file I/O stay on the host in both (about 6.5 ms of the GPU CLI's 15 ms per "
                                              "1080p frame: glue/x264dsp_glue.c's own clock, X264DSP_GLUE_STATS)")
            pfl["value"] = pfl["settings"]["dia_subme1"]["value"]
            pfl["e2e"] = {"value": world * pf_e2e_frames / pf_e2e_s_max, "unit": "frames/s", "h2d_bytes_per_step": pf_e2e[1],
                          "d2h_bytes_per_step": pf_e2e[2], "setting": f"dia_subme1, {pf_e2e_frames} frames per call",
                          "api": "x264dsp_p_frames_host_packed (pinned I420 pictures in; types, vectors, mvd, nnz, cbp and the levels "
                                 "as the compact stream the entropy coder reads out, the reconstruction stays on the device; reference "
                                 "planes, lowres planes and the lookahead of every pair built on the device inside the call)",
                          "dense_door": None if pf_e2e[3] is None else
                                        {"value": pf_e2e_frames / pf_e2e[3], "d2h_bytes_per_step": pf_e2e[4],
                                         "api": "x264dsp_p_frames_host: all 392 levels per macroblock and the reconstructed "
                                                "pictures come back too"}}
            if gops is not None:
                frames_g = GOPS_N * GOPS_LEN
                pfl["gops"] = {"workload": f"{GOPS_N} closed GOPs of {GOPS_LEN} 1080p frames (I + P, DIA / subme 1, QP 23 / 26, in-loop filter on), every "
                                           "stage on the device: slice kernels, boundary strengths, deblocking, border, half-pel planes; each "
                                           "frame predicts from the device's own previous reconstruction (tests/test_gpu_gop_chain.py: identical "
                                           "to the running reference encoder frame by frame)",
                               "value": 1e3 * frames_g / gops[0], "unit": "frames/s", "us_per_frame": 1e3 * gops[0] / frames_g,
                               "note": "an I frame costs about 0.6 ms, a P frame with its filter stages about 0.11 ms at this batch size",
                               "e2e": {"value": frames_g / gops[1], "unit": "frames/s", "h2d_bytes_per_step": gops[2],
                                       "d2h_bytes_per_step": gops[3], "api": "x264dsp_gops_encode_host"}}
            line["pframe"] = pfl
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
