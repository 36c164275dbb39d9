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
