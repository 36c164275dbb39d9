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
    ap.add_argument("--no-me", action="store_true", help="skip the secondary full-resolution ME measurement")
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

def cpu_lookahead(w, h, clip_len, luma_clips, threads, repeats):
    """runs the lookahead pass of len(luma_clips) clips on `threads` host threads, one clip at a
    time per thread; returns (seconds of the timed region, kind, frames processed)"""
    import cpu_checkers as cc
    lib = cc.ref() if os.path.exists(cc.REF_SO) or os.path.isdir(cc.REFERENCE_TREE) else None
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
            enc = cc.RefEncoder(w, h)
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
                            enc.load(frames[i], clip_pics[i])   # x264_frame_copy_picture (the GPU arm's e2e has its H2D copy)
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


def run_reference_arm(args, pkg):
    w, h = args.width, args.height
    threads = args.cpu_threads or host_cores()
    # four clips per thread and step: a bounded sample of the GPU arm's step (args.clips clips per GPU)
    n_clips = threads * 4
    base = make_clips(pkg, w, h, min(n_clips, 8), args.clip_len)
    luma_clips = [base.reshape(-1, args.clip_len, w * h)[c % min(n_clips, 8)] for c in range(n_clips)]
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
    line = {
        "impl": "reference",
        "metric": "1080p ME frames/sec" if (w, h) == (1920, 1080) else f"{w}x{h} ME frames/sec",
        "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, frames_per_step, sample=f"{n_clips} clips of {args.clip_len} frames per step, "
                                  f"four clips per host thread"),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{n_clips} clips x {args.clip_len} frames per step x {args.steps} steps"},
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
    if sample:
        cfg["reference_sample"] = sample
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
            run_reference_arm(args, pkg)
        return 0

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

    # ---- secondary: full-resolution motion search (configs[2]), rank 0 only
    me_ms = me_blocks0 = me_slots = None
    if rank == 0 and not args.no_me:
        me_pairs = min(8, clip_len - 1)
        me_ms, me_blocks0, me_slots = me_search_measure(pkg, ctx, torch, g, luma_dev, d_mvs[: clip_len].cpu().numpy(), me_pairs)
    rc_ms = rc_check = None
    if rank == 0 and not args.no_me:
        rc_frames = 96 if w * h <= 1920 * 1088 else 16
        rc_ms, rc_check = recon_measure(pkg, ctx, torch, g, w, h, rc_frames)

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

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
                    "ms_per_step": e2e_ms / args.steps, "api": "x264dsp_lookahead_clips_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "host_affinity": numa,
            "clocks": clocks,
            "roofline": {"kernel": "xd_la_multi_kernel<4>", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(bytes_per_pair * pairs_per_launch),
                         "avg_launch_ms": inter_avg_s * 1e3, "launches_timed": inter_n,
                         "int_pipe": int_pipe_roofline(sad_px, satd_px, (prof["la_inter"][0] + prof["la_intra"][0]) / args.steps / 1e3,
                                                       "xd_la_multi_kernel<4> + xd_la_intra_kernel"),
                         "note": "dependent-search wavefront kernel (SURVEY 8(d) config 2): HBM % reported as required; ncu "
                                 "(profiles/r01k) shows instruction issue (56 % of 4 warp-inst/clk/SM), the ALU pipe (60 % of "
                                 "its measured 2 warp-inst/clk/SM, profiles/int_pipe_peak.json) and the L1 data pipe (61 % of "
                                 "peak wavefronts) as the limiters; DRAM traffic is below the algorithmic bytes because "
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
        if me_ms is not None:
            tot_ms = sum(me_ms.values())
            me = {"workload": f"{w}x{h} x264_me_search_ref + x264_me_refine_qpel, HEX, subme 5, range 16, QP 26, one block list "
                              "per partition size tiling the frame (7 sizes), mvp from the lowres MVs",
                  "frames_per_s_all_sizes": 1e3 / tot_ms, "ms_per_frame_by_size": me_ms,
                  "launch": f"x264dsp_me_search_sized_frames_dev, {min(8, clip_len - 1)} frame pairs per launch"}
            if not args.no_cpu_baseline and world == 1:
                cnt = me_search_cpu_counts(pkg, g, me_slots, me_blocks0)
                sad = sum(c["sad_pix"] for c in cnt.values())
                satd = sum(c["satd_pix"] for c in cnt.values())
                me.update({"sad_gpix_per_s": sad / (tot_ms / 1e3) / 1e9, "satd_gpix_per_s": satd / (tot_ms / 1e3) / 1e9,
                           "sad_pix_per_frame": sad, "satd_pix_per_frame": satd,
                           "counted_by": "the CPU oracle's instrumented run of the same block lists (frame pair 0)",
                           "bit_exact_vs_oracle": all(c["bit_exact"] for c in cnt.values()),
                           "cpu_port_1core_frames_per_s": 1.0 / sum(c["cpu_s"] for c in cnt.values()),
                           "int_pipe": int_pipe_roofline(sad, satd, tot_ms / 1e3, "xd_me_sized_kernel<W,H>, 7 sizes")})
            line["me_search"] = me
        if rc_ms is not None:
            tot = sum(rc_ms.values())
            rc = {"workload": f"{w}x{h} inter macroblocks: x264_mb_mc 16x16 + x264_macroblock_encode (4x4 DCT, quant, dequant, "
                              "IDCT, decimation, chroma DC) + x264_frame_deblock_row, QP 26, frame-batched launches",
                  "frames_per_s": 1e3 / tot, "ms_per_frame": rc_ms,
                  "hbm_frac": {"residual": 16.5e6 / (rc_ms["residual"] / 1e3) / 1e9 / peak,
                               "deblock": 6.8e6 / (rc_ms["deblock"] / 1e3) / 1e9 / peak} if (w, h) == (1920, 1080) else None}
            if not args.no_cpu_baseline and world == 1:
                ok, cpu_s, coded = recon_cpu_check(pkg, g, rc_check)
                rc.update({"bit_exact_vs_oracle": ok, "coded_mbs_frame0": coded, "cpu_port_1core_frames_per_s": 1.0 / cpu_s})
            line["recon"] = rc
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
