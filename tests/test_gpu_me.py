"""GPU parity: x264dsp_me_search_batch_dev against the CPU oracle (itself pinned to the reference's
x264_me_search_ref / x264_me_refine_qpel in tests/test_oracle_vs_ref.py), all partition sizes,
DIA and HEX, subme 1..5, with and without the final qpel refinement.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from test_oracle_vs_ref import make_me_blocks

pytestmark = pytest.mark.gpu


def prepare(pkg, ctx, w, h):
    import torch
    frames = [pkg.synth_frame(w, h, i) for i in range(2)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    o = cc.oracle()
    host = []
    for f in frames:
        s = np.zeros(go.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(go), cc.ptr(f), cc.ptr(s))
        o.xo_frame_expand_border(C.byref(go), cc.ptr(s))
        o.xo_frame_filter(C.byref(go), cc.ptr(s))
        host.append(s)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(2 * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, 2)
    ctx.frame_expand_border(g, slots, 2)
    ctx.frame_filter(g, slots, 2)
    ctx.sync()
    return g, go, host, slots


@pytest.mark.parametrize("me,subme,refine", [(0, 1, 1), (0, 2, 0), (1, 2, 1), (1, 3, 0), (1, 4, 1), (1, 5, 1), (0, 5, 0), (1, 1, 0)])
def test_me_search_batch(pkg, ctx, me, subme, refine):
    import torch
    w, h = 352, 288
    g, go, host, slots = prepare(pkg, ctx, w, h)
    o = cc.oracle()
    rng = np.random.RandomState(500 + me * 10 + subme)
    ref_slot, enc_slot = slots[: g.slot_bytes], slots[g.slot_bytes:]
    for size in range(8):
        for qp, mv_scale in ((26, 24), (38, 90)):
            n = 400
            blocks = make_me_blocks(go, rng, size, n, mv_scale)
            want = np.zeros(n, cc.ME_RESULT_DTYPE)
            prm = cc.MeParams(me, subme, 16, qp, refine)
            o.xo_me_search_batch(C.byref(go), cc.ptr(host[1]), cc.ptr(host[0]), C.byref(prm), n,
                                 blocks.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
            d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
            d_res = torch.zeros(n * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            p = pkg.MeParams(me, subme, 16, qp, refine)
            ctx.me_search_batch(g, enc_slot, ref_slot, p, n, d_blocks, d_res)
            ctx.sync()
            got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)
            bad = [i for i in range(n) if got[i] != want[i]]
            assert not bad, (f"me {me} subme {subme} refine {refine} size {size} qp {qp}: {len(bad)}/{n} differ; "
                             f"first {bad[0]}: gpu {got[bad[0]]} oracle {want[bad[0]]} block {blocks[bad[0]]}")
            # the size-specialised kernel (G lanes per block) must agree as well
            d_res.zero_()
            torch.cuda.synchronize()
            ctx.me_search_sized(g, enc_slot, ref_slot, p, size, n, d_blocks, d_res)
            ctx.sync()
            got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)
            bad = [i for i in range(n) if got[i] != want[i]]
            assert not bad, (f"SIZED me {me} subme {subme} refine {refine} size {size} qp {qp}: {len(bad)}/{n} differ; "
                             f"first {bad[0]}: gpu {got[bad[0]]} oracle {want[bad[0]]} block {blocks[bad[0]]}")


def test_me_search_1080p_tiling(pkg, ctx):
    """config 3 shape: every 16x16 and 8x8 block of a 1080p frame, HEX + subme 5 + qpel refine,
    mvp/mvc as a harness would derive them"""
    import torch
    w, h = 1920, 1080
    g, go, host, slots = prepare(pkg, ctx, w, h)
    o = cc.oracle()
    rng = np.random.RandomState(9)
    for size, step in ((0, 16), (3, 8)):
        xs, ys = np.meshgrid(np.arange(0, g.luma_w, step), np.arange(0, g.luma_h, step))
        n = xs.size
        blocks = np.zeros(n, cc.ME_BLOCK_DTYPE)
        blocks["i_pixel"] = size
        blocks["bx"] = xs.ravel()
        blocks["by"] = ys.ravel()
        mbx, mby = blocks["bx"] // 16, blocks["by"] // 16
        fmv = 512 << 2
        for k, (mb, nmb) in enumerate(((mbx, g.mb_w), (mby, g.mb_h))):
            smin = np.clip((-(mb << 4) - 24) << 2, -fmv, fmv - 1)
            smax = np.clip((((nmb - mb - 1) << 4) + 24) << 2, -fmv, fmv - 1)
            blocks["mv_min_spel"][:, k], blocks["mv_max_spel"][:, k] = smin, smax
            blocks["mv_min_fpel"][:, k], blocks["mv_max_fpel"][:, k] = (smin >> 2) + 6, (smax >> 2) - 6
        blocks["mvp"] = rng.randint(-16, 17, (n, 2))
        blocks["i_mvc"] = 2
        blocks["mvc"][:, 0] = blocks["mvp"]
        blocks["mvc"][:, 1] = 0
        keep = rng.choice(n, 4000, replace=False)           # the oracle is a scalar CPU loop
        sub = np.ascontiguousarray(blocks[keep])
        want = np.zeros(len(sub), cc.ME_RESULT_DTYPE)
        prm = cc.MeParams(1, 5, 16, 26, 1)
        o.xo_me_search_batch(C.byref(go), cc.ptr(host[1]), cc.ptr(host[0]), C.byref(prm), len(sub),
                             sub.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
        d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
        d_res = torch.zeros(n * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.me_search_batch(g, slots[g.slot_bytes:], slots[: g.slot_bytes], pkg.MeParams(1, 5, 16, 26, 1), n, d_blocks, d_res)
        ctx.sync()
        got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)[keep]
        assert np.array_equal(got, want), f"size {size}: {np.count_nonzero(got != want)} of {len(sub)} differ"
        d_res.zero_()
        torch.cuda.synchronize()
        ctx.me_search_sized(g, slots[g.slot_bytes:], slots[: g.slot_bytes], pkg.MeParams(1, 5, 16, 26, 1), size, n, d_blocks, d_res)
        ctx.sync()
        got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)[keep]
        assert np.array_equal(got, want), f"sized, size {size}: {np.count_nonzero(got != want)} of {len(sub)} differ"


@pytest.mark.parametrize("me,subme,refine", [(2, 2, 0), (2, 5, 1), (3, 1, 1), (3, 4, 0), (4, 2, 1), (4, 5, 0), (4, 1, 1)])
def test_me_search_umh_esa_tesa(pkg, ctx, me, subme, refine):
    """me = UMH / ESA / TESA do what the reference does with them (encoder/me.c:389-394: no pattern search; TESA with
    subme >= 2: SATD as the full-pel metric) -- oracle pinned in tests/test_oracle_vs_ref.py::test_me_search_umh_esa_tesa"""
    import torch
    w, h = 352, 288
    g, go, host, slots = prepare(pkg, ctx, w, h)
    o = cc.oracle()
    rng = np.random.RandomState(40 + me * 10 + subme)
    ref_slot, enc_slot = slots[: g.slot_bytes], slots[g.slot_bytes:]
    for size in range(8):
        n = 300
        blocks = make_me_blocks(go, rng, size, n, 40)
        want = np.zeros(n, cc.ME_RESULT_DTYPE)
        prm = cc.MeParams(me, subme, 16, 30, refine)
        o.xo_me_search_batch(C.byref(go), cc.ptr(host[1]), cc.ptr(host[0]), C.byref(prm), n,
                             blocks.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
        d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
        p = pkg.MeParams(me, subme, 16, 30, refine)
        for which in ("batch", "sized"):
            d_res = torch.zeros(n * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            if which == "batch":
                ctx.me_search_batch(g, enc_slot, ref_slot, p, n, d_blocks, d_res)
            else:
                ctx.me_search_sized(g, enc_slot, ref_slot, p, size, n, d_blocks, d_res)
            ctx.sync()
            got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)
            assert np.array_equal(got, want), f"{which} me {me} subme {subme} size {size}: {np.count_nonzero(got != want)}/{n} differ"


@pytest.mark.parametrize("me,subme", [(1, 2), (1, 4), (1, 5), (0, 3), (2, 5)])
def test_me_halfpel_thresh_and_refdupe(pkg, ctx, me, subme):
    """x264_me_search_ref with p_halfpel_thresh, x264_me_refine_qpel_refdupe and x264_me_refine_qpel alone
    (encoder/me.c:421, 426-440, 526-539): x264dsp_me_search_batch_ex_dev against the oracle (pinned to the reference in
    tests/test_oracle_vs_ref.py::test_me_halfpel_thresh_and_refdupe)"""
    import torch
    w, h = 352, 288
    g, go, host, slots = prepare(pkg, ctx, w, h)
    o = cc.oracle()
    rng = np.random.RandomState(77 + me * 10 + subme)
    ref_slot, enc_slot = slots[: g.slot_bytes], slots[g.slot_bytes:]
    for size in range(8):
        n = 300
        blocks = make_me_blocks(go, rng, size, n, 40)
        prm = cc.MeParams(me, subme, 16, 28, 0)
        p = pkg.MeParams(me, subme, 16, 28, 0)
        base = np.zeros(n, cc.ME_RESULT_DTYPE)
        o.xo_me_search_batch(C.byref(go), cc.ptr(host[1]), cc.ptr(host[0]), C.byref(prm), n,
                             blocks.ctypes.data_as(C.c_void_p), base.ctypes.data_as(C.c_void_p))
        thresh0 = (base["cost"] * rng.choice([0.5, 0.8, 0.95, 1.0, 1.3, 4.0], n)).astype(np.int32)
        thresh0[rng.rand(n) < 0.1] = 2**31 - 1
        d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
        for mode in (0, 1, 2):
            start = base.copy()
            if mode:
                start["mv"] = (start["mv"] & ~3) if mode == 1 else start["mv"]
                start["mv"] += rng.randint(-1, 2, (n, 2)) * 4
                start["cost"] += rng.randint(0, 50, n)
            want, t_want = start.copy(), thresh0.copy()
            use_t = mode != 2
            o.xo_me_search_batch_ex(C.byref(go), cc.ptr(host[1]), cc.ptr(host[0]), C.byref(prm), n,
                                    blocks.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), mode,
                                    cc.ptr(t_want, cc.i32p) if use_t else None)
            d_res = torch.from_numpy(start.copy().view(np.uint8)).cuda()
            d_t = torch.from_numpy(thresh0.copy()).cuda() if use_t else None
            torch.cuda.synchronize()
            ctx.me_search_batch_ex(g, enc_slot, ref_slot, p, n, d_blocks, d_res, mode, d_t)
            ctx.sync()
            got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)
            assert np.array_equal(got, want), f"mode {mode} size {size}: {np.count_nonzero(got != want)}/{n} results differ"
            if use_t:
                assert np.array_equal(d_t.cpu().numpy(), t_want), f"mode {mode} size {size}: thresholds differ"


def test_me_search_sized_frames_1080p(pkg, ctx):
    """x264dsp_me_search_sized_frames_dev -- the frame-batched call bench.py times for configs[2] -- at 1080p: all seven
    partition sizes tiling the frame, three frame pairs per launch, HEX + subme 5 + qpel refine, EVERY block of every
    pair compared with the oracle (0.2 s per frame and size on one core)."""
    import torch
    w, h, nf = 1920, 1080, 3
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    o = cc.oracle()
    frames = [pkg.synth_frame(w, h, i) for i in range(nf + 1)]
    host = []
    for f in frames:
        s = np.zeros(go.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(go), cc.ptr(f), cc.ptr(s))
        o.xo_frame_expand_border(C.byref(go), cc.ptr(s))
        o.xo_frame_filter(C.byref(go), cc.ptr(s))
        host.append(s)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros((nf + 1) * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, nf + 1)
    ctx.frame_expand_border(g, slots, nf + 1)
    ctx.frame_filter(g, slots, nf + 1)
    ctx.sync()
    rng = np.random.RandomState(31)
    prm_t = (1, 5, 16, 26, 1)
    for size in range(7):
        # pair f: frame f+1 searched in frame f; the MB-level MVs differ per pair so that the block lists differ
        per_pair = [pkg.tiling_blocks(g, size, rng.randint(-12, 13, (g.mb_count, 2)).astype(np.int16)) for _ in range(nf)]
        n = len(per_pair[0])
        blocks = np.concatenate(per_pair)
        d_blocks = torch.from_numpy(blocks.view(np.uint8)).cuda()
        d_res = torch.zeros(nf * n * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.me_search_sized_frames(g, slots[g.slot_bytes:], slots, nf, pkg.MeParams(*prm_t), size, n, d_blocks, d_res)
        ctx.sync()
        got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE).reshape(nf, n)
        for f in range(nf):
            want = np.zeros(n, cc.ME_RESULT_DTYPE)
            o.xo_me_search_batch(C.byref(go), cc.ptr(host[f + 1]), cc.ptr(host[f]), C.byref(cc.MeParams(*prm_t)), n,
                                 per_pair[f].ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
            bad = np.nonzero(got[f] != want)[0]
            assert len(bad) == 0, (f"size {size} pair {f}: {len(bad)} of {n} blocks differ; first {bad[0]}: "
                                   f"gpu {got[f][bad[0]]} oracle {want[bad[0]]}")


def test_predict_mv_batch(pkg, ctx):
    """x264_mb_predict_mv_16x16 / x264_mb_predict_mv_pskip for a batch of neighbourhoods against the oracle (pinned to the
    reference's functions, tests/test_oracle_vs_ref.py::test_predict_mv_16x16_and_pskip)"""
    import torch
    o = cc.oracle()
    rng = np.random.RandomState(162)
    n = 20000
    ref = rng.choice([-2, -1, 0, 0, 0, 1], (n, 4)).astype(np.int8)
    mv = rng.randint(-40, 41, (n, 4, 2)).astype(np.int16)
    mv[rng.rand(n, 4) < 0.25] = 0
    same = rng.rand(n) < 0.2
    mv[same] = mv[same][:, :1]
    i_ref = rng.choice([0, 0, 1], n).astype(np.int8)
    nb = np.zeros((n, 20), np.uint8)
    nb[:, :4] = ref.view(np.uint8)
    nb[:, 4:] = mv.view(np.uint8).reshape(n, 16)
    shape = (rng.randint(0, 5, n) + 8 * (rng.rand(n) < 0.3)).astype(np.uint8)
    want_p, want_s = np.zeros((n, 2), np.int16), np.zeros((n, 2), np.int16)
    for i in range(n):
        o.xo_predict_mv_part(cc.ptr(nb[i]), int(i_ref[i]), int(shape[i] & 7), int(shape[i] >> 3), cc.ptr(want_p[i], cc.i16p))
        o.xo_predict_mv_pskip(cc.ptr(nb[i]), cc.ptr(want_s[i], cc.i16p))
    d_p = torch.full((n, 2), 77, dtype=torch.int16, device="cuda")
    d_s = torch.full((n, 2), 77, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.predict_mv_batch(n, torch.from_numpy(nb).cuda(), torch.from_numpy(i_ref).cuda(), d_p, d_s, shape=torch.from_numpy(shape).cuda())
    ctx.sync()
    bad = np.nonzero((d_p.cpu().numpy() != want_p).any(1))[0]
    assert len(bad) == 0, f"mvp differs at {bad[:5]}: shape {shape[bad[:5]]} ref {ref[bad[:5]].tolist()}"
    d_p.fill_(77)
    ctx.predict_mv_batch(n, torch.from_numpy(nb).cuda(), None, d_p, None)          # defaults: reference 0, shape 0
    ctx.sync()
    w0 = np.zeros((n, 2), np.int16)
    for i in range(0, n, 7):
        o.xo_predict_mv_16x16(cc.ptr(nb[i]), 0, cc.ptr(w0[i], cc.i16p))
        assert np.array_equal(d_p[i].cpu().numpy(), w0[i]), f"default path, macroblock {i}"
    assert np.array_equal(d_s.cpu().numpy(), want_s), "pskip mv"
    assert (want_s == 0).all(1).any() and (want_s != 0).any()


@pytest.mark.parametrize("mb_w,mb_h,nf", [(22, 18, 3), (120, 68, 2), (1, 1, 1), (13, 2, 2)])
def test_predict_mvc_16x16_frames(pkg, ctx, mb_w, mb_h, nf):
    """x264_mb_predict_mv_ref16x16 for whole frames against the oracle (pinned to the reference's function,
    tests/test_oracle_vs_ref.py::test_predict_mvc_16x16_frame): with and without lookahead MVs / temporal candidates"""
    import torch
    o = cc.oracle()
    rng = np.random.RandomState(mb_w * 7 + nf)
    n = mb_w * mb_h
    for variant in range(4):
        lowres = rng.randint(-300, 301, (nf, n, 2)).astype(np.int16) if variant & 1 else None
        if lowres is not None:
            lowres[rng.rand(nf, n) < 0.2] = [-17000, 16500]
            if nf > 1:
                lowres[1, 0, 0] = 0x7FFF                         # this frame pair has not been analysed
        mvr = rng.randint(-200, 201, (nf, n, 2)).astype(np.int16)
        l0 = rng.randint(-200, 201, (nf, n, 2)).astype(np.int16) if variant & 2 else None
        scale = [128, 256, 77, 385][variant]
        want_c, want_n = np.full((nf, n, 9, 2), 999, np.int16), np.zeros((nf, n), np.int32)
        for f in range(nf):
            o.xo_predict_mvc_16x16_frame(mb_w, mb_h, cc.ptr(lowres[f], cc.i16p) if lowres is not None else None,
                                         cc.ptr(mvr[f], cc.i16p), cc.ptr(l0[f], cc.i16p) if l0 is not None else None, scale,
                                         cc.ptr(want_c[f], cc.i16p), cc.ptr(want_n[f], cc.i32p))
        d_c = torch.full((nf, n, 9, 2), 999, dtype=torch.int16, device="cuda")
        d_n = torch.zeros((nf, n), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.predict_mvc_16x16_frames(mb_w, mb_h, nf, torch.from_numpy(lowres).cuda() if lowres is not None else None,
                                     torch.from_numpy(mvr).cuda(), torch.from_numpy(l0).cuda() if l0 is not None else None,
                                     scale, d_c, d_n)
        ctx.sync()
        assert np.array_equal(d_n.cpu().numpy(), want_n), f"variant {variant}: counts"
        assert np.array_equal(d_c.cpu().numpy(), want_c), f"variant {variant}: candidates"
