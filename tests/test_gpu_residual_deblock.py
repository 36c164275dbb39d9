"""GPU parity: motion compensation, residual coding (DCT / quant / dequant / IDCT / decimation /
chroma DC) and in-loop deblocking of whole frames, CUDA through the C ABI against the CPU oracle
(pinned to the reference's x264_macroblock_encode / x264_frame_deblock_row).  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr, i16p, i8p

pytestmark = pytest.mark.gpu


def _slots(pkg, ctx, w, h, n, filtered):
    import torch
    frames = [pkg.synth_frame(w, h, i) for i in range(n)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    o = cc.oracle()
    host = []
    for f in frames:
        s = np.zeros(go.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(go), ptr(f), ptr(s))
        if filtered:
            o.xo_frame_expand_border(C.byref(go), ptr(s))
            o.xo_frame_filter(C.byref(go), ptr(s))
        host.append(s)
    dev = torch.from_numpy(np.concatenate(host)).cuda()
    return g, go, host, dev


@pytest.mark.parametrize("w,h,qp", [(352, 288, 26), (200, 120, 20), (352, 288, 38), (1920, 1080, 26), (352, 288, 12)])
def test_mc_and_residual_frame(pkg, ctx, w, h, qp):
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, 2, True)
    o = cc.oracle()
    rng = np.random.RandomState(qp + w)
    n = g.mb_count
    # MVs around the true pan (3,2 px/frame -> 12,8 qpel) with outliers, incl. out-of-range ones that get clipped
    mv = (np.array([12, 8]) + rng.randint(-6, 7, (n, 2))).astype(np.int16)
    mv[rng.rand(n) < 0.05] = rng.randint(-3000, 3000, 2)
    mv[rng.rand(n) < 0.1] = 0

    pred_o = np.zeros(go.slot_bytes, np.uint8)
    o.xo_mc_frame(C.byref(go), ptr(host[0]), ptr(mv, i16p), ptr(pred_o))
    ref_slot, enc_slot = dev[: g.slot_bytes], dev[g.slot_bytes:]
    pred = torch.zeros(g.slot_bytes, dtype=torch.uint8, device="cuda")
    d_mv = torch.from_numpy(mv).cuda()
    torch.cuda.synchronize()
    ctx.mc_frame(g, ref_slot, d_mv, pred)
    ctx.sync()
    assert np.array_equal(pred.cpu().numpy(), pred_o), "mc_frame"

    lv_o = np.zeros((n, pkg.RES_LEVELS_PER_MB), np.int16)
    nz_o = np.zeros((n, pkg.RES_NNZ_PER_MB), np.uint8)
    cbp_o = np.zeros(n, np.int16)
    o.xo_residual_frame(C.byref(go), ptr(host[1]), ptr(pred_o), qp, ptr(lv_o, i16p), ptr(nz_o), ptr(cbp_o, i16p))
    lv = torch.full((n, pkg.RES_LEVELS_PER_MB), -1, dtype=torch.int16, device="cuda")
    nz = torch.full((n, pkg.RES_NNZ_PER_MB), 77, dtype=torch.uint8, device="cuda")
    cbp = torch.full((n,), -1, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.residual_frame(g, enc_slot, pred, qp, lv, nz, cbp)
    ctx.sync()
    bad = np.nonzero(cbp.cpu().numpy() != cbp_o)[0]
    assert len(bad) == 0, f"cbp differs at MBs {bad[:5]}: {cbp.cpu().numpy()[bad[:5]]} vs {cbp_o[bad[:5]]}"
    assert np.array_equal(nz.cpu().numpy(), nz_o), "nnz"
    bad = np.nonzero((lv.cpu().numpy() != lv_o).any(1))[0]
    assert len(bad) == 0, f"levels differ at MBs {bad[:5]}"
    assert np.array_equal(pred.cpu().numpy(), pred_o), "reconstruction"
    # the test must exercise coded, decimated and skipped macroblocks
    assert (cbp_o == 0).any() or qp < 30
    assert (cbp_o != 0).any()


@pytest.mark.parametrize("w,h,nf,qp", [(352, 288, 4, 26), (1920, 1080, 3, 30), (200, 120, 5, 18)])
def test_mc_and_residual_frames_batch(pkg, ctx, w, h, nf, qp):
    """nf frame pairs in one launch each (x264dsp_mc_frames_dev / x264dsp_residual_frames_dev) against the
    oracle frame by frame"""
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, nf + 1, True)
    o = cc.oracle()
    rng = np.random.RandomState(qp + nf)
    n = g.mb_count
    mv = (np.array([12, 8]) + rng.randint(-6, 7, (nf, n, 2))).astype(np.int16)
    mv[rng.rand(nf, n) < 0.05] = rng.randint(-3000, 3000, 2)
    pred_o, lv_o, nz_o, cbp_o = [], [], [], []
    for f in range(nf):
        po = np.zeros(go.slot_bytes, np.uint8)
        o.xo_mc_frame(C.byref(go), ptr(host[f]), ptr(mv[f], i16p), ptr(po))
        lv, nz, cb = np.zeros((n, pkg.RES_LEVELS_PER_MB), np.int16), np.zeros((n, pkg.RES_NNZ_PER_MB), np.uint8), np.zeros(n, np.int16)
        mc_only = po.copy()
        o.xo_residual_frame(C.byref(go), ptr(host[f + 1]), ptr(po), qp, ptr(lv, i16p), ptr(nz), ptr(cb, i16p))
        pred_o.append((mc_only, po)); lv_o.append(lv); nz_o.append(nz); cbp_o.append(cb)
    pred = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    d_mv = torch.from_numpy(mv).cuda()
    torch.cuda.synchronize()
    ctx.mc_frames(g, dev[: nf * g.slot_bytes], nf, d_mv, pred)
    ctx.sync()
    got = pred.cpu().numpy().reshape(nf, -1)
    for f in range(nf):
        assert np.array_equal(got[f], pred_o[f][0]), f"mc frame {f}"
    lv = torch.full((nf, n, pkg.RES_LEVELS_PER_MB), -1, dtype=torch.int16, device="cuda")
    nz = torch.full((nf, n, pkg.RES_NNZ_PER_MB), 77, dtype=torch.uint8, device="cuda")
    cbp = torch.full((nf, n), -1, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.residual_frames(g, dev[g.slot_bytes:], pred, nf, qp, lv, nz, cbp)
    ctx.sync()
    got = pred.cpu().numpy().reshape(nf, -1)
    for f in range(nf):
        assert np.array_equal(cbp.cpu().numpy()[f], cbp_o[f]), f"cbp frame {f}"
        assert np.array_equal(nz.cpu().numpy()[f], nz_o[f]), f"nnz frame {f}"
        assert np.array_equal(lv.cpu().numpy()[f], lv_o[f]), f"levels frame {f}"
        assert np.array_equal(got[f], pred_o[f][1]), f"recon frame {f}"


@pytest.mark.parametrize("w,h,qp,aoff,boff", [(352, 288, 26, 0, 0), (200, 120, 38, 0, 0), (352, 288, 20, 3, -2),
                                              (1920, 1080, 30, 0, 0), (64, 64, 51, 0, 0), (352, 288, 14, 0, 0)])
def test_deblock_frame(pkg, ctx, w, h, qp, aoff, boff):
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, 1, False)
    o = cc.oracle()
    rng = np.random.RandomState(qp * 7 + w)
    n = g.mb_count
    mb_type = rng.choice([0, 2, 4, 5, 6], n, p=[0.05, 0.05, 0.5, 0.2, 0.2]).astype(np.int8)
    partition = rng.choice([13, 14, 15, 16], n).astype(np.uint8)
    cbp = (rng.randint(0, 48, n) * (rng.rand(n) < 0.6)).astype(np.int16)
    bs = rng.randint(0, 4, (n, 2, 8, 4)).astype(np.uint8)
    bs[rng.rand(n) < 0.2] = 0
    want = host[0].copy()
    o.xo_deblock_frame(C.byref(go), ptr(want), ptr(mb_type, i8p), ptr(partition), ptr(cbp, i16p), ptr(bs), qp, aoff, boff)
    slot = dev.clone()
    d = [torch.from_numpy(a).cuda() for a in (mb_type, partition, cbp, bs)]
    torch.cuda.synchronize()
    for rep in range(2):                       # twice: the row counters must reset between launches
        slot.copy_(dev)
        torch.cuda.synchronize()
        ctx.deblock_frame(g, slot, d[0], d[1], d[2], d[3], qp, aoff, boff)
        ctx.sync()
        got = slot.cpu().numpy()
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, f"rep {rep}: {len(bad)} bytes differ, first at offset {bad[:4]}"
    if qp >= 20:
        assert np.count_nonzero(want != host[0]) > 0, "the test must exercise the filter"


@pytest.mark.parametrize("w,h,nf,qp", [(352, 288, 5, 28), (1920, 1080, 3, 32), (64, 64, 7, 40)])
def test_deblock_frames_batch(pkg, ctx, w, h, nf, qp):
    """several independent frames in one launch (rows of all frames share the ticket queue)"""
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, nf, False)
    o = cc.oracle()
    rng = np.random.RandomState(qp + nf)
    n = g.mb_count
    mb_type = rng.choice([0, 2, 4, 5, 6], (nf, n), p=[0.05, 0.05, 0.5, 0.2, 0.2]).astype(np.int8)
    partition = rng.choice([13, 14, 15, 16], (nf, n)).astype(np.uint8)
    cbp = (rng.randint(0, 48, (nf, n)) * (rng.rand(nf, n) < 0.6)).astype(np.int16)
    bs = rng.randint(0, 4, (nf, n, 2, 8, 4)).astype(np.uint8)
    bs[rng.rand(nf, n) < 0.2] = 0
    want = []
    for f in range(nf):
        s = host[f].copy()
        o.xo_deblock_frame(C.byref(go), ptr(s), ptr(mb_type[f], i8p), ptr(partition[f]), ptr(cbp[f], i16p), ptr(bs[f]), qp, 0, 0)
        want.append(s)
    want = np.concatenate(want)
    slots = dev.clone()
    d = [torch.from_numpy(a).cuda() for a in (mb_type, partition, cbp, bs)]
    torch.cuda.synchronize()
    ctx.deblock_frames(g, slots, nf, d[0], d[1], d[2], d[3], qp)
    ctx.sync()
    got = slots.cpu().numpy()
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"{len(bad)} bytes differ, first at {bad[:4]} (slot {bad[0] // g.slot_bytes})"


@pytest.mark.parametrize("n,skew", [(5000, 0), (4999, 0), (1, 0), (7, 0), (8161, 0), (333, 4)])
def test_deblock_strength(pkg, ctx, n, skew):
    """odd counts exercise the staged kernel's tail, skew != 0 the direct kernel (inputs not 16-byte aligned)"""
    import torch
    rng = np.random.RandomState(3 + n)
    nnz = (rng.rand(n, 120) < 0.3).astype(np.uint8)
    ref = rng.randint(-1, 2, (n, 2, 40)).astype(np.int8)
    mv = rng.randint(-6, 7, (n, 2, 40, 2)).astype(np.int16)
    want = np.zeros((n, 2, 8, 4), np.uint8)
    cc.oracle().xo_deblock_strength(n, ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(want))
    d = []
    for a in (nnz, ref, mv):
        raw = torch.zeros(a.nbytes + 64, dtype=torch.uint8, device="cuda")
        raw[skew: skew + a.nbytes] = torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()
        d.append(raw[skew: skew + a.nbytes])
    bs = torch.zeros((n, 2, 8, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.deblock_strength(n, d[0], d[1], d[2], bs)
    ctx.sync()
    got = bs.cpu().numpy()
    assert np.array_equal(got[:, :, :4], want[:, :, :4])


@pytest.mark.parametrize("n,skew", [(5000, 0), (4999, 0), (1, 0), (8161, 0), (333, 4)])
def test_macroblock_deblock_strength_with_intra(pkg, ctx, n, skew):
    """x264_macroblock_deblock_strength (common/macroblock.c:677-691): intra macroblocks mixed in -- inner edges 3, edge 0
    untouched; bs starts from a random pattern so that untouched bytes are compared as well"""
    import torch
    rng = np.random.RandomState(13 + n)
    mb_type = rng.choice([0, 1, 2, 3, 4, 5, 6], n).astype(np.int8)
    nnz = (rng.rand(n, 120) < 0.3).astype(np.uint8)
    ref = rng.randint(-1, 2, (n, 2, 40)).astype(np.int8)
    mv = rng.randint(-6, 7, (n, 2, 40, 2)).astype(np.int16)
    start = rng.randint(0, 256, (n, 2, 8, 4)).astype(np.uint8)
    want = start.copy()
    cc.oracle().xo_macroblock_deblock_strength(n, ptr(mb_type, i8p), ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(want))
    assert (want[mb_type < 4][:, :, 1:4] == 3).all() and np.array_equal(want[:, :, 4:], start[:, :, 4:])
    d = []
    for a in (nnz, ref, mv):
        raw = torch.zeros(a.nbytes + 64, dtype=torch.uint8, device="cuda")
        raw[skew: skew + a.nbytes] = torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()
        d.append(raw[skew: skew + a.nbytes])
    bs = torch.from_numpy(start.copy()).cuda()
    torch.cuda.synchronize()
    ctx.macroblock_deblock_strength(n, torch.from_numpy(mb_type).cuda(), d[0], d[1], d[2], bs)
    ctx.sync()
    assert np.array_equal(bs.cpu().numpy(), want)


@pytest.mark.parametrize("w,h,nf,qp,intra_share", [(352, 288, 2, 26, 0.5), (200, 120, 3, 18, 1.0), (352, 288, 2, 38, 0.3),
                                                    (1920, 1080, 2, 30, 0.5), (352, 288, 2, 12, 1.0), (208, 160, 2, 51, 0.5)])
def test_residual_frames_typed_inter_i16x16_i4x4(pkg, ctx, w, h, nf, qp, intra_share):
    """x264dsp_residual_frames_typed_dev: inter macroblocks (P slice rules), I16x16 macroblocks (I slice rules: intra
    tables, luma DC block, no decimation) and I4x4 macroblocks (sixteen serial predict / transform / reconstruct steps
    on random mode sets; placed on the even-even lattice so that none neighbours another) mixed in one launch, against
    the oracle (pinned to the reference's x264_macroblock_encode for all three kinds, tests/test_oracle_vs_ref.py)"""
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, nf + 1, False)
    o = cc.oracle()
    rng = np.random.RandomState(qp + nf + w)
    n = g.mb_count
    kind = (rng.rand(nf, n) < intra_share).astype(np.uint8)
    xs, ys = np.arange(n) % g.mb_w, np.arange(n) // g.mb_w
    lattice = (xs % 2 == 0) & (ys % 2 == 0)
    i4 = lattice[None, :] & (rng.rand(nf, n) < 0.7)
    kind[i4] = 2
    kind[i4 & (rng.rand(nf, n) < 0.3)] = 6                   # a row above but no top-right macroblock
    modes = rng.randint(0, 12, (nf, n, 16)).astype(np.uint8)
    # prediction: the previous frame (zero-motion inter prediction); intra macroblocks get a flat (DC-like) luma block
    # or a vertical / horizontal extension of the source's own first row / column, and flat chroma
    preds = []
    for f in range(nf):
        p = host[f].copy()
        src = host[f + 1]
        luma = p[go.luma_origin:].reshape(-1)[: (16 * go.mb_h) * go.luma_stride].reshape(16 * go.mb_h, go.luma_stride)
        sl = src[go.luma_origin:].reshape(-1)[: (16 * go.mb_h) * go.luma_stride].reshape(16 * go.mb_h, go.luma_stride)
        co = go.slot_chroma_off + go.chroma_origin
        chroma = p[co:].reshape(-1)[: (8 * go.mb_h) * go.chroma_stride].reshape(8 * go.mb_h, go.chroma_stride)
        sc = src[co:].reshape(-1)[: (8 * go.mb_h) * go.chroma_stride].reshape(8 * go.mb_h, go.chroma_stride)
        for xy in np.nonzero(kind[f] == 1)[0]:
            mx, my = xy % go.mb_w, xy // go.mb_w
            blk = sl[16 * my: 16 * my + 16, 16 * mx: 16 * mx + 16]
            mode = rng.randint(4)
            if mode == 0:
                luma[16 * my: 16 * my + 16, 16 * mx: 16 * mx + 16] = int(blk.mean())
            elif mode == 1:
                luma[16 * my: 16 * my + 16, 16 * mx: 16 * mx + 16] = blk[0][None, :]
            elif mode == 2:
                luma[16 * my: 16 * my + 16, 16 * mx: 16 * mx + 16] = blk[:, 0][:, None]
            else:                                            # nearly perfect prediction: DC-only / empty macroblocks
                luma[16 * my: 16 * my + 16, 16 * mx: 16 * mx + 16] = np.clip(blk.astype(int) + rng.randint(-2, 3), 0, 255)
            cb = sc[8 * my: 8 * my + 8, 16 * mx: 16 * mx + 16]
            chroma[8 * my: 8 * my + 8, 16 * mx: 16 * mx + 16: 2] = int(cb[:, 0::2].mean())
            chroma[8 * my: 8 * my + 8, 16 * mx + 1: 16 * mx + 16: 2] = int(cb[:, 1::2].mean())
        preds.append(p)
    lv_o = np.zeros((nf, n, pkg.RES_LEVELS_PER_MB), np.int16)
    dc_o = np.zeros((nf, n, 16), np.int16)
    nz_o = np.zeros((nf, n, pkg.RES_NNZ_PER_MB), np.uint8)
    cbp_o = np.zeros((nf, n), np.int16)
    rec_o = []
    for f in range(nf):
        po = preds[f].copy()
        o.xo_residual_frame_typed(C.byref(go), ptr(host[f + 1]), ptr(po), qp, ptr(kind[f]), ptr(modes[f]), ptr(lv_o[f], i16p),
                                  ptr(dc_o[f], i16p), ptr(nz_o[f]), ptr(cbp_o[f], i16p))
        rec_o.append(po)
    pred = torch.from_numpy(np.concatenate(preds)).cuda()
    d_kind = torch.from_numpy(kind).cuda()
    lv = torch.full((nf, n, pkg.RES_LEVELS_PER_MB), -1, dtype=torch.int16, device="cuda")
    dc = torch.full((nf, n, 16), -1, dtype=torch.int16, device="cuda")
    nz = torch.full((nf, n, pkg.RES_NNZ_PER_MB), 77, dtype=torch.uint8, device="cuda")
    cbp = torch.full((nf, n), -1, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.residual_frames_typed(g, dev[g.slot_bytes:], pred, nf, qp, d_kind, lv, dc, nz, cbp, i4_modes=torch.from_numpy(modes).cuda())
    ctx.sync()
    got = pred.cpu().numpy().reshape(nf, -1)
    for f in range(nf):
        bad = np.nonzero(cbp[f].cpu().numpy() != cbp_o[f])[0]
        assert len(bad) == 0, f"frame {f}: cbp differs at MBs {bad[:5]} kinds {kind[f][bad[:5]]}: {cbp[f].cpu().numpy()[bad[:5]]} vs {cbp_o[f][bad[:5]]}"
        assert np.array_equal(nz[f].cpu().numpy(), nz_o[f]), f"frame {f}: nnz"
        assert np.array_equal(lv[f].cpu().numpy(), lv_o[f]), f"frame {f}: levels"
        assert np.array_equal(dc[f].cpu().numpy(), dc_o[f]), f"frame {f}: luma DC levels"
        assert np.array_equal(got[f], rec_o[f]), f"frame {f}: reconstruction"
    assert ((cbp_o & 15)[(kind & 3) == 2] != 0).any(), "no coded I4x4 macroblock"
    if intra_share > 0:
        i16 = kind == 1
        assert (nz_o[..., 24][i16] != 0).any(), "no I16x16 macroblock with a coded DC block"
        assert ((cbp_o & 15)[i16] == 15).any() and ((cbp_o & 15)[i16] == 0).any(), "I16x16: both luma cbp cases wanted"


@pytest.mark.parametrize("w,h,nf,qp", [(352, 288, 3, 26), (1920, 1080, 2, 30), (200, 120, 3, 20), (352, 288, 2, 40), (352, 288, 2, 14)])
def test_probe_pskip_frames(pkg, ctx, w, h, nf, qp):
    """x264_macroblock_probe_pskip for whole frames: P_SKIP predictions from x264dsp_mc_frames_dev at MVs around the
    clip's true pan (most macroblocks are skippable, the outliers are not), decision per macroblock against the oracle
    (pinned to the reference's function, tests/test_oracle_vs_ref.py::test_probe_pskip_mb)"""
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, nf + 1, True)
    o = cc.oracle()
    rng = np.random.RandomState(qp + nf + h)
    n = g.mb_count
    mv = (np.array([12, 8]) + rng.randint(-1, 2, (nf, n, 2))).astype(np.int16)
    mv[rng.rand(nf, n) < 0.25] = rng.randint(-24, 25, 2)
    want = np.zeros((nf, n), np.uint8)
    for f in range(nf):
        po = np.zeros(go.slot_bytes, np.uint8)
        o.xo_mc_frame(C.byref(go), ptr(host[f]), ptr(mv[f], i16p), ptr(po))
        o.xo_probe_pskip_frame(C.byref(go), ptr(host[f + 1]), ptr(po), qp, ptr(want[f]))
    pred = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    skip = torch.full((nf, n), 7, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.mc_frames(g, dev[: nf * g.slot_bytes], nf, torch.from_numpy(mv).cuda(), pred)
    before = pred.clone()
    ctx.probe_pskip_frames(g, dev[g.slot_bytes:], pred, nf, qp, skip)
    ctx.sync()
    got = skip.cpu().numpy()
    bad = np.argwhere(got != want)
    assert len(bad) == 0, f"{len(bad)} decisions differ, first (frame, mb) {bad[:5].tolist()}: {got[tuple(bad[0])]} vs {want[tuple(bad[0])]}"
    assert torch.equal(before, pred), "the probe must not modify the prediction"
    assert not want.all() and (want.any() or qp < 26), f"one-sided test: {int(want.sum())} of {want.size} skippable"


@pytest.mark.parametrize("w,h,nf", [(352, 288, 2), (200, 120, 3), (1920, 1080, 1)])
def test_mc_frames_part(pkg, ctx, w, h, nf):
    """one MV per 8x8 block (every partition of x264_mb_mc) against the oracle; with four equal MVs per macroblock the
    oracle's partition-wise composition must reproduce its 16x16 one (which is what the reference's mc tables pin)"""
    import torch
    g, go, host, dev = _slots(pkg, ctx, w, h, nf, True)
    o = cc.oracle()
    rng = np.random.RandomState(w + nf)
    n = g.mb_count
    mv = (np.array([12, 8]) + rng.randint(-9, 10, (nf, n, 4, 2))).astype(np.int16)
    mv[rng.rand(nf, n) < 0.05] = rng.randint(-3000, 3000, 2)             # clipped to the macroblock's limits
    same = rng.rand(nf, n) < 0.3
    mv[same] = mv[same][:, :1]                                            # 16x16 macroblocks
    want = []
    for f in range(nf):
        po = np.zeros(go.slot_bytes, np.uint8)
        o.xo_mc_frame_part(C.byref(go), ptr(host[f]), ptr(mv[f], i16p), ptr(po))
        want.append(po)
    mv16 = np.ascontiguousarray(mv[0][:, 0])
    a, b = np.zeros(go.slot_bytes, np.uint8), np.zeros(go.slot_bytes, np.uint8)
    o.xo_mc_frame(C.byref(go), ptr(host[0]), ptr(mv16, i16p), ptr(a))
    o.xo_mc_frame_part(C.byref(go), ptr(host[0]), ptr(np.ascontiguousarray(np.repeat(mv16[:, None], 4, 1)), i16p), ptr(b))
    assert np.array_equal(a, b), "oracle: partition-wise MC of a 16x16 macroblock differs from the 16x16 call"
    pred = torch.zeros(nf * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.mc_frames_part(g, dev, nf, torch.from_numpy(mv).cuda(), pred)
    ctx.sync()
    got = pred.cpu().numpy().reshape(nf, -1)
    for f in range(nf):
        assert np.array_equal(got[f], want[f]), f"frame {f}"
