"""GPU parity of the host-memory doors of the full-resolution paths (x264-dsp_b200/csrc/host_paths.cu), the calls bench.py
times end to end for configs[2] and configs[3]: pictures and side information in host memory in, results in host memory
out, compared with the oracle chain on the same input."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr, i8p, i16p

pytestmark = pytest.mark.gpu


def oracle_planes(go, frames, hpel):
    o = cc.oracle()
    out = []
    for f in frames:
        s = np.zeros(go.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(go), ptr(f), ptr(s))
        o.xo_frame_expand_border(C.byref(go), ptr(s))
        if hpel:
            o.xo_frame_filter(C.byref(go), ptr(s))
        out.append(s)
    return out


@pytest.mark.parametrize("w,h,pairs", [(352, 288, 3), (208, 160, 2)])
def test_me_search_frames_host(pkg, ctx, w, h, pairs):
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    o = cc.oracle()
    frames = [pkg.synth_frame(w, h, i) for i in range(pairs + 1)]
    luma = np.stack([f[: w * h] for f in frames])
    host = oracle_planes(go, frames, True)
    rng = np.random.RandomState(w + pairs)
    sizes = list(range(7))
    blocks = []
    for s in sizes:
        per_pair = [pkg.tiling_blocks(g, s, rng.randint(-10, 11, (g.mb_count, 2)).astype(np.int16)) for _ in range(pairs)]
        blocks.append(np.concatenate(per_pair))
    prm_t = (1, 5, 16, 26, 1)
    res = ctx.me_search_frames_host(w, h, luma, pkg.MeParams(*prm_t), sizes, blocks)
    for s in sizes:
        n = len(blocks[s]) // pairs
        for p in range(pairs):
            want = np.zeros(n, cc.ME_RESULT_DTYPE)
            o.xo_me_search_batch(C.byref(go), ptr(host[p + 1]), ptr(host[p]), C.byref(cc.MeParams(*prm_t)), n,
                                 np.ascontiguousarray(blocks[s][p * n:(p + 1) * n]).ctypes.data_as(C.c_void_p),
                                 want.ctypes.data_as(C.c_void_p))
            got = res[s][p * n:(p + 1) * n]
            assert np.array_equal(got, want), f"size {s} pair {p}: {np.count_nonzero(got != want)} of {n} differ"


@pytest.mark.parametrize("w,h,n,qp", [(352, 288, 9, 26), (208, 160, 3, 20), (200, 120, 5, 34)])
def test_recon_frames_host(pkg, ctx, w, h, n, qp):
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    o = cc.oracle()
    frames = [pkg.synth_frame(w, h, i) for i in range(n + 1)]
    i420 = np.stack(frames)
    host = oracle_planes(go, frames, True)
    rng = np.random.RandomState(qp + n)
    nmb = g.mb_count
    mv = (np.array([12, 8]) + rng.randint(-5, 6, (n, nmb, 2))).astype(np.int16)
    mb_type = np.full((n, nmb), 4, np.int8)
    part = np.full((n, nmb), 16, np.uint8)
    bs = ((rng.rand(n, nmb, 64) < 0.35) * rng.randint(1, 4, (n, nmb, 64))).astype(np.uint8)
    lv, nz, cbp, rec = ctx.recon_frames_host(w, h, i420, mv, qp, mb_type, part, bs)
    for f in range(n):
        pred = np.zeros(go.slot_bytes, np.uint8)
        o.xo_mc_frame(C.byref(go), ptr(host[f]), ptr(mv[f], i16p), ptr(pred))
        lv_o, nz_o, cbp_o = np.zeros((nmb, pkg.RES_LEVELS_PER_MB), np.int16), np.zeros((nmb, pkg.RES_NNZ_PER_MB), np.uint8), np.zeros(nmb, np.int16)
        o.xo_residual_frame(C.byref(go), ptr(host[f + 1]), ptr(pred), qp, ptr(lv_o, i16p), ptr(nz_o), ptr(cbp_o, i16p))
        assert np.array_equal(lv[f], lv_o) and np.array_equal(nz[f], nz_o) and np.array_equal(cbp[f], cbp_o), f"frame {f}: levels / nnz / cbp"
        o.xo_deblock_frame(C.byref(go), ptr(pred), ptr(mb_type[f], i8p), ptr(part[f]), ptr(cbp_o, i16p), ptr(bs[f]), qp, 0, 0)
        y = pred[go.luma_origin:][: go.luma_h * go.luma_stride].reshape(go.luma_h, go.luma_stride)[:h, :w]
        c = pred[go.slot_chroma_off + go.chroma_origin:][: (go.luma_h // 2) * go.chroma_stride].reshape(go.luma_h // 2, go.chroma_stride)[: h // 2, :w]
        want = np.concatenate([y.ravel(), c[:, 0::2].ravel(), c[:, 1::2].ravel()])
        assert np.array_equal(rec[f], want), f"frame {f}: deblocked reconstruction ({np.count_nonzero(rec[f] != want)} bytes differ)"


@pytest.mark.parametrize("w,h,n,me,subme,qp", [(352, 288, 9, 0, 1, 28), (208, 160, 17, 1, 4, 26)])
def test_p_frames_host_matches_oracle(pkg, ctx, w, h, n, me, subme, qp, monkeypatch):
    """x264dsp_p_frames_host: pictures in host memory in, the coded P frames out -- reference planes, half-resolution planes,
    the lookahead's vectors and the macroblock loop all on the device, several stream groups -- against the oracle's
    xo_p_frame fed with the oracle's own planes and lookahead vectors"""
    import ctypes as C
    monkeypatch.setenv("X264DSP_PF_HOST_GROUPS", "3")        # the default is one group per ~96 frames
    from cpu_checkers import ptr, i16p, i32p
    o = cc.oracle()
    go = cc.oracle_geom(w, h)
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    pics = np.stack([pkg.synth_frame(w, h, i) for i in range(n + 1)])
    out = {"mb_type": np.zeros((n, nmb), np.int8), "mv": np.zeros((n, nmb, 2), np.int16), "mvr": np.zeros((n, nmb, 2), np.int16),
           "mvd": np.zeros((n, nmb, 2), np.int16), "levels": np.zeros((n, nmb, pkg.RES_LEVELS_PER_MB), np.int16),
           "nnz": np.zeros((n, nmb, pkg.RES_NNZ_PER_MB), np.uint8), "cbp": np.zeros((n, nmb), np.int16)}
    recon = np.zeros((n, w * h * 3 // 2), np.uint8)
    prm = pkg.PFrameParams(me, subme, 16, qp, 128, 1, 0)
    ctx.p_frames_host(w, h, n, pics, prm, out["mb_type"], out["mv"], out["mvr"], out["mvd"], out["levels"], out["nnz"],
                      out["cbp"], recon)

    class P(C.Structure):
        _fields_ = [(k, C.c_int32) for k in ("me_method", "subpel_refine", "me_range", "qp", "mv_range", "fast_pskip", "mvc_scale", "analyse_inter")]
    slots = [np.zeros(g.slot_bytes, np.uint8) for _ in range(n + 1)]
    for i in range(n + 1):
        o.xo_frame_load_i420(C.byref(go), ptr(pics[i]), ptr(slots[i]))
        o.xo_frame_expand_border(C.byref(go), ptr(slots[i]))
        o.xo_frame_filter(C.byref(go), ptr(slots[i]))
        o.xo_frame_init_lowres(C.byref(go), ptr(slots[i]))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in range(n):
        lmv = np.zeros((nmb, 2), np.int16)
        lc = np.zeros(nmb, np.int32)
        ls = np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(go), ptr(slots[k + 1]), ptr(slots[k]), 0, ptr(lmv, i16p), ptr(lc, i32p), ptr(ls, i32p), None)
        want = {"mb_type": np.zeros(nmb, np.int8), "mv": np.zeros((nmb, 2), np.int16), "mvr": np.zeros((nmb, 2), np.int16),
                "mvd": np.zeros((nmb, 2), np.int16), "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8),
                "cbp": np.zeros(nmb, np.int16)}
        wrec = np.zeros(g.slot_bytes, np.uint8)
        p = P(me, subme, 16, qp, 128, 1, 0)
        o.xo_p_frame(C.byref(go), ptr(slots[k + 1]), ptr(slots[k]), ptr(wrec), C.byref(p), vp(lmv), None, vp(want["mb_type"]),
                     vp(want["mv"]), vp(want["mvr"]), vp(want["mvd"]), vp(want["levels"]), vp(want["nnz"]), vp(want["cbp"]))
        for key in want:
            assert np.array_equal(out[key][k], want[key]), f"frame {k + 1}: {key} differs"
        wy = wrec[g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:h, :w]
        assert np.array_equal(recon[k][: w * h].reshape(h, w), wy), f"frame {k + 1}: luma reconstruction differs"
        co = g.slot_chroma_off + g.chroma_origin
        wc = wrec[co:][: (g.luma_h // 2) * g.chroma_stride].reshape(g.luma_h // 2, g.chroma_stride)[: h // 2, :w]
        assert np.array_equal(recon[k][w * h: w * h + w * h // 4].reshape(h // 2, w // 2), wc[:, 0::2]), f"frame {k + 1}: U differs"
        assert np.array_equal(recon[k][w * h + w * h // 4:].reshape(h // 2, w // 2), wc[:, 1::2]), f"frame {k + 1}: V differs"


@pytest.mark.parametrize("w,h,n,me,subme,qp", [(352, 288, 9, 1, 5, 24), (208, 160, 17, 0, 2, 28)])
def test_p_frames_part_host_matches_oracle(pkg, ctx, w, h, n, me, subme, qp, monkeypatch):
    """x264dsp_p_frames_part_host (analyse.inter = PSUB16x16) against xo_p_frame_part on the oracle's own planes and vectors"""
    import ctypes as C
    monkeypatch.setenv("X264DSP_PF_HOST_GROUPS", "2")
    from cpu_checkers import ptr, i16p, i32p
    o = cc.oracle()
    go = cc.oracle_geom(w, h)
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    pics = np.stack([pkg.synth_frame(w, h, i, cut_frame=4) for i in range(n + 1)])
    shapes = {"mb_type": ((nmb,), np.int8), "partition": ((nmb,), np.uint8), "mv8": ((nmb, 4, 2), np.int16), "mvr": ((nmb, 2), np.int16),
              "mvd8": ((nmb, 4, 2), np.int16), "levels": ((nmb, 392), np.int16), "nnz": ((nmb, 27), np.uint8), "cbp": ((nmb,), np.int16)}
    out = {k: np.zeros((n,) + s, t) for k, (s, t) in shapes.items()}
    recon = np.zeros((n, w * h * 3 // 2), np.uint8)
    ctx.p_frames_part_host(w, h, n, pics, pkg.PFrameParams(me, subme, 16, qp, 128, 1, 0, 1), out["mb_type"], out["partition"], out["mv8"],
                           out["mvr"], out["mvd8"], out["levels"], out["nnz"], out["cbp"], recon)

    class P(C.Structure):
        _fields_ = [(k, C.c_int32) for k in ("me_method", "subpel_refine", "me_range", "qp", "mv_range", "fast_pskip", "mvc_scale", "analyse_inter")]
    slots = [np.zeros(g.slot_bytes, np.uint8) for _ in range(n + 1)]
    for i in range(n + 1):
        o.xo_frame_load_i420(C.byref(go), ptr(pics[i]), ptr(slots[i]))
        o.xo_frame_expand_border(C.byref(go), ptr(slots[i]))
        o.xo_frame_filter(C.byref(go), ptr(slots[i]))
        o.xo_frame_init_lowres(C.byref(go), ptr(slots[i]))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    seen = set()
    for k in range(n):
        lmv = np.zeros((nmb, 2), np.int16)
        lc = np.zeros(nmb, np.int32)
        ls = np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(go), ptr(slots[k + 1]), ptr(slots[k]), 0, ptr(lmv, i16p), ptr(lc, i32p), ptr(ls, i32p), None)
        want = {key: np.zeros(s, t) for key, (s, t) in shapes.items()}
        wrec = np.zeros(g.slot_bytes, np.uint8)
        p = P(me, subme, 16, qp, 128, 1, 0, 1)
        o.xo_p_frame_part(C.byref(go), ptr(slots[k + 1]), ptr(slots[k]), ptr(wrec), C.byref(p), vp(lmv), None,
                          *[vp(want[key]) for key in ("mb_type", "partition", "mv8", "mvr", "mvd8", "levels", "nnz", "cbp")])
        for key in want:
            assert np.array_equal(out[key][k], want[key]), f"frame {k + 1}: {key} differs"
        wy = wrec[g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:h, :w]
        assert np.array_equal(recon[k][: w * h].reshape(h, w), wy), f"frame {k + 1}: luma reconstruction differs"
        seen |= set(np.unique(want["partition"]).tolist())
    assert seen >= {13, 16}, seen


HO_FLAG = list(range(16)) + [25, 26] + list(range(16, 24))                 # nnz index of unit u
HO_DENSE = [16 * u for u in range(16)] + [256, 260] + [264 + 16 * k for k in range(8)]
HO_LEN = [16] * 16 + [4, 4] + [16] * 8


def expand_packed(stream, mb_offset, nnz):
    """the compact hand-off of one frame back to the dense [mb][392] form (units without a flag stay zero)"""
    nmb = nnz.shape[0]
    dense = np.zeros((nmb, 392), np.int16)
    for mb in range(nmb):
        at = int(mb_offset[mb])
        for u in range(26):
            if nnz[mb, HO_FLAG[u]]:
                dense[mb, HO_DENSE[u]: HO_DENSE[u] + HO_LEN[u]] = stream[at: at + HO_LEN[u]]
                at += HO_LEN[u]
    return dense


def mask_dense(levels, nnz):
    """the dense levels with every unit the entropy coder does not read (flag 0) cleared"""
    out = np.zeros_like(levels)
    for u in range(26):
        on = nnz[:, HO_FLAG[u]] != 0
        out[on, HO_DENSE[u]: HO_DENSE[u] + HO_LEN[u]] = levels[on, HO_DENSE[u]: HO_DENSE[u] + HO_LEN[u]]
    return out


def test_levels_pack_matches_dense(pkg, ctx):
    """x264dsp_levels_pack_dev on random flags / levels: offsets are the running sum of the present units' lengths, the stream
    holds exactly those units in order"""
    import torch
    rng = np.random.RandomState(5)
    n, nmb = 3, 700
    nnz = (rng.rand(n, nmb, 27) < 0.3).astype(np.uint8) * rng.randint(1, 17, (n, nmb, 27)).astype(np.uint8)
    nnz[1, :50] = 0                                                   # macroblocks with nothing coded
    nnz[2, 100:130] = 3                                               # and with everything coded
    levels = rng.randint(-300, 300, (n, nmb, 392)).astype(np.int16)
    stride = nmb * 392
    d_packed = torch.full((n * stride,), 12345, dtype=torch.int16, device="cuda")
    d_off = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_tot = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.levels_pack(n, nmb, torch.from_numpy(levels).cuda(), torch.from_numpy(nnz).cuda(), d_packed, stride, d_off, d_tot)
    ctx.sync()
    packed, off, tot = d_packed.cpu().numpy().reshape(n, stride), d_off.cpu().numpy(), d_tot.cpu().numpy()
    for f in range(n):
        sizes = sum((nnz[f][:, HO_FLAG[u]] != 0).astype(np.int64) * HO_LEN[u] for u in range(26))
        assert np.array_equal(off[f], np.concatenate([[0], np.cumsum(sizes)[:-1]])), f"frame {f}: offsets"
        assert tot[f] == sizes.sum()
        assert np.array_equal(expand_packed(packed[f], off[f], nnz[f]), mask_dense(levels[f], nnz[f])), f"frame {f}: stream"
        assert (packed[f][tot[f]:] == 12345).all(), "wrote past the stream's end"


@pytest.mark.parametrize("w,h,n,me,subme,qp,part", [(352, 288, 9, 1, 5, 24, 0), (208, 160, 11, 0, 2, 28, 1)])
def test_p_frames_host_packed_matches_dense(pkg, ctx, w, h, n, me, subme, qp, part, monkeypatch):
    """x264dsp_p_frames_host_packed == x264dsp_p_frames_host (which is checked against the oracle) with the levels compacted
    and the reconstruction left on the device"""
    monkeypatch.setenv("X264DSP_PF_HOST_GROUPS", "3")
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    nv = 4 if part else 1
    pics = np.stack([pkg.synth_frame(w, h, i, cut_frame=4) for i in range(n + 1)])
    prm = pkg.PFrameParams(me, subme, 16, qp, 128, 1, 0, part)
    mk = lambda: {"mb_type": np.zeros((n, nmb), np.int8), "partition": np.zeros((n, nmb), np.uint8),
                  "mv": np.zeros((n, nmb, nv, 2), np.int16), "mvr": np.zeros((n, nmb, 2), np.int16),
                  "mvd": np.zeros((n, nmb, nv, 2), np.int16), "nnz": np.zeros((n, nmb, 27), np.uint8), "cbp": np.zeros((n, nmb), np.int16)}
    a, b = mk(), mk()
    levels = np.zeros((n, nmb, 392), np.int16)
    recon = np.zeros((n, w * h * 3 // 2), np.uint8)
    if part:
        ctx.p_frames_part_host(w, h, n, pics, prm, a["mb_type"], a["partition"], a["mv"], a["mvr"], a["mvd"], levels, a["nnz"], a["cbp"], recon)
    else:
        ctx.p_frames_host(w, h, n, pics, prm, a["mb_type"], a["mv"], a["mvr"], a["mvd"], levels, a["nnz"], a["cbp"], recon)
    packed = np.full(n * nmb * 392, 777, np.int16)
    f_off = np.zeros(n + 1, np.int64)
    mb_off = np.zeros((n, nmb), np.int32)
    ctx.p_frames_host_packed(w, h, n, pics, prm, b["mb_type"], b["partition"] if part else None, b["mv"], b["mvr"], b["mvd"], packed,
                             f_off, mb_off, b["nnz"], b["cbp"])
    for key in a:
        if key == "partition" and not part:
            continue
        assert np.array_equal(a[key], b[key]), key
    assert f_off[0] == 0 and (np.diff(f_off) >= 0).all() and f_off[-1] < packed.size // 2, f_off
    for f in range(n):
        got = expand_packed(packed[f_off[f]: f_off[f + 1]], mb_off[f], b["nnz"][f])
        assert np.array_equal(got, mask_dense(levels[f], a["nnz"][f])), f"frame {f}: compact stream != dense levels"
    assert (packed[f_off[-1]:] == 777).all()
    # a buffer that cannot hold the content is refused, not overrun
    small = np.zeros(max(int(f_off[-1]) // 2, 4), np.int16)
    with pytest.raises(pkg.X264DspError):
        ctx.p_frames_host_packed(w, h, n, pics, prm, b["mb_type"], b["partition"] if part else None, b["mv"], b["mvr"], b["mvd"], small,
                                 f_off, mb_off, b["nnz"], b["cbp"])
