"""x264dsp_p_frames_dev (x264_macroblock_analyse + x264_macroblock_encode for every macroblock of a P frame,
as a wavefront on the device) against the CPU oracle's xo_p_frame, which tests/test_oracle_pframe.py pins to the running
reference encoder.  Inputs are built on the device the way an encoder would have them: source frames staged from I420,
the reference frame border-expanded and half-pel filtered, the lookahead's vectors of the pair as the first search
candidate, and -- from the second frame of a chain on -- the previous frame's 16x16 vectors as temporal candidates.
Compared bit for bit: macroblock types, final vectors, mvr, cbp, levels, nnz and the reconstruction."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr

pytestmark = pytest.mark.gpu


class OPFrameParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("me_method", "subpel_refine", "me_range", "qp", "mv_range", "fast_pskip", "mvc_scale", "analyse_inter")]


def vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.mark.parametrize("w,h,n,me,subme,qp,cut", [
    (176, 144, 3, 0, 1, 26, -1), (352, 288, 3, 1, 2, 30, -1), (208, 160, 4, 1, 5, 24, 2), (352, 288, 3, 0, 3, 36, -1),
    (1920, 1080, 2, 1, 4, 28, -1)])
def test_p_frames_match_oracle(pkg, ctx, w, h, n, me, subme, qp, cut):
    import torch
    o = cc.oracle()
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    nmb = g.mb_count
    frames = np.stack([pkg.synth_frame(w, h, i, cut_frame=cut) for i in range(n + 1)])
    i420 = torch.from_numpy(frames).cuda()
    slots = torch.zeros((n + 1) * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, i420, slots, n + 1)
    ctx.frame_expand_border(g, slots, n + 1)
    ctx.frame_filter(g, slots, n + 1)
    ctx.frame_init_lowres(g, slots, n + 1)
    # lookahead vectors of every pair (frame k+1 against frame k)
    b = np.arange(1, n + 1, dtype=np.int32)
    p0 = b - 1
    d_lmv = torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, slots, b, p0, np.ones(n, np.uint8), d_lmv, d_lc, d_ls)
    ctx.sync()
    host_slots = slots.cpu().numpy().reshape(n + 1, g.slot_bytes)
    lmv = d_lmv.cpu().numpy()

    def device_run(first, count, l0):
        out = {"mb_type": torch.full((count, nmb), -1, dtype=torch.int8, device="cuda"),
               "mv": torch.zeros((count, nmb, 2), dtype=torch.int16, device="cuda"),
               "mvr": torch.zeros((count, nmb, 2), dtype=torch.int16, device="cuda"),
               "mvd": torch.full((count, nmb, 2), 77, dtype=torch.int16, device="cuda"),
               "levels": torch.ones((count, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
               "nnz": torch.ones((count, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
               "cbp": torch.full((count, nmb), -1, dtype=torch.int16, device="cuda")}
        recon = torch.zeros(count * g.slot_bytes, dtype=torch.uint8, device="cuda")
        prm = pkg.PFrameParams(me, subme, 16, qp, 128, 1, 256 if l0 is not None else 0)
        ctx.p_frames(g, slots[(first + 1) * g.slot_bytes:], slots[first * g.slot_bytes:], recon, count, prm,
                     d_lmv[first:first + count].contiguous(), l0, out["mb_type"], out["mv"], out["mvr"], out["levels"],
                     out["nnz"], out["cbp"], mvd=out["mvd"])
        ctx.sync()
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res["recon"] = recon.cpu().numpy().reshape(count, g.slot_bytes)
        return res

    def oracle_run(k, l0):
        res = {"mb_type": np.zeros(nmb, np.int8), "mv": np.zeros((nmb, 2), np.int16), "mvr": np.zeros((nmb, 2), np.int16),
               "mvd": np.zeros((nmb, 2), np.int16), "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
        recon = np.zeros(g.slot_bytes, np.uint8)
        prm = OPFrameParams(me, subme, 16, qp, 128, 1, 256 if l0 is not None else 0)
        o.xo_p_frame(C.byref(go), ptr(host_slots[k + 1]), ptr(host_slots[k]), ptr(recon), C.byref(prm), vp(lmv[k]), vp(l0),
                     vp(res["mb_type"]), vp(res["mv"]), vp(res["mvr"]), vp(res["mvd"]), vp(res["levels"]), vp(res["nnz"]), vp(res["cbp"]))
        res["recon"] = recon
        return res

    def compare(got, want, tag):
        for key in ("mb_type", "mv", "mvr", "mvd", "cbp", "nnz", "levels"):
            if not np.array_equal(got[key], want[key]):
                d = np.flatnonzero((got[key].reshape(nmb, -1) != want[key].reshape(nmb, -1)).any(1))
                raise AssertionError(f"{tag}: {key} differs at macroblocks {d[:8]} ({d.size} in all): "
                                     f"{got[key][d[0]].ravel()[:8]} vs {want[key][d[0]].ravel()[:8]}")
        lo, co = g.luma_origin, g.slot_chroma_off + g.chroma_origin
        gy = got["recon"][lo:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
        wy = want["recon"][lo:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
        gc = got["recon"][co:][: (g.luma_h // 2) * g.chroma_stride].reshape(g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
        wc = want["recon"][co:][: (g.luma_h // 2) * g.chroma_stride].reshape(g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
        assert np.array_equal(gy, wy), f"{tag}: luma reconstruction differs"
        assert np.array_equal(gc, wc), f"{tag}: chroma reconstruction differs"

    # (1) all n pairs as independent frames of ONE launch, no temporal candidates
    got = device_run(0, n, None)
    skipped = coded = 0
    wants = []
    for k in range(n):
        want = oracle_run(k, None)
        wants.append(want)
        compare({key: v[k] for key, v in got.items()}, want, f"{w}x{h} frame {k + 1} (batched launch)")
        skipped += int((want["mb_type"] == pkg.MB_P_SKIP).sum())
        coded += int((want["mb_type"] == pkg.MB_P_L0).sum())
    assert coded > 0, "no coded macroblock at all"
    if qp >= 28:
        assert skipped > 0, "no skipped macroblock at all"
    # (2) a chain: frame 2 takes frame 1's 16x16 vectors as temporal candidates (scale 256 = the reference's POC step)
    if n >= 2:
        l0_dev = torch.from_numpy(wants[0]["mvr"]).cuda().reshape(1, nmb, 2)
        got2 = device_run(1, 1, l0_dev)
        want2 = oracle_run(1, wants[0]["mvr"])
        compare({key: v[0] for key, v in got2.items()}, want2, f"{w}x{h} frame 2 with temporal candidates")


@pytest.mark.parametrize("w,h,n,me,subme,qp,cut,inter", [
    (176, 144, 3, 0, 1, 26, -1, 1), (352, 288, 3, 1, 2, 30, 2, 1), (208, 160, 4, 1, 5, 22, 2, 1), (352, 288, 3, 0, 3, 36, -1, 1),
    (1920, 1080, 2, 1, 4, 28, -1, 1), (208, 160, 3, 1, 5, 26, -1, 0)])
def test_p_frames_with_partitions_match_oracle(pkg, ctx, w, h, n, me, subme, qp, cut, inter):
    """x264dsp_p_frames_part_dev (analyse.inter = PSUB16x16: P8x8 / P16x8 / P8x16 as well) against xo_p_frame_part, which
    tests/test_oracle_pframe.py pins to the running reference encoder"""
    import torch
    o = cc.oracle()
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    nmb = g.mb_count
    frames = np.stack([pkg.synth_frame(w, h, i, cut_frame=cut) for i in range(n + 1)])
    slots = torch.zeros((n + 1) * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(frames).cuda(), slots, n + 1)
    ctx.frame_expand_border(g, slots, n + 1)
    ctx.frame_filter(g, slots, n + 1)
    ctx.frame_init_lowres(g, slots, n + 1)
    b = np.arange(1, n + 1, dtype=np.int32)
    d_lmv = torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, slots, b, b - 1, np.ones(n, np.uint8), d_lmv, d_lc, d_ls)
    ctx.sync()
    host_slots = slots.cpu().numpy().reshape(n + 1, g.slot_bytes)
    lmv = d_lmv.cpu().numpy()
    keys = ("mb_type", "partition", "mv8", "mvr", "mvd8", "levels", "nnz", "cbp")

    def device_run(first, count, l0):
        out = {"mb_type": torch.full((count, nmb), -1, dtype=torch.int8, device="cuda"),
               "partition": torch.zeros((count, nmb), dtype=torch.uint8, device="cuda"),
               "mv8": torch.zeros((count, nmb, 4, 2), dtype=torch.int16, device="cuda"),
               "mvr": torch.zeros((count, nmb, 2), dtype=torch.int16, device="cuda"),
               "mvd8": torch.full((count, nmb, 4, 2), 77, dtype=torch.int16, device="cuda"),
               "levels": torch.ones((count, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
               "nnz": torch.ones((count, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
               "cbp": torch.full((count, nmb), -1, dtype=torch.int16, device="cuda")}
        recon = torch.zeros(count * g.slot_bytes, dtype=torch.uint8, device="cuda")
        prm = pkg.PFrameParams(me, subme, 16, qp, 128, 1, 256 if l0 is not None else 0, inter)
        ctx.p_frames_part(g, slots[(first + 1) * g.slot_bytes:], slots[first * g.slot_bytes:], recon, count, prm,
                          d_lmv[first:first + count].contiguous(), l0, out["mb_type"], out["partition"], out["mv8"], out["mvr"],
                          out["levels"], out["nnz"], out["cbp"], mvd8=out["mvd8"])
        ctx.sync()
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res["recon"] = recon.cpu().numpy().reshape(count, g.slot_bytes)
        return res

    def oracle_run(k, l0):
        res = {"mb_type": np.zeros(nmb, np.int8), "partition": np.zeros(nmb, np.uint8), "mv8": np.zeros((nmb, 4, 2), np.int16),
               "mvr": np.zeros((nmb, 2), np.int16), "mvd8": np.zeros((nmb, 4, 2), np.int16),
               "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
        recon = np.zeros(g.slot_bytes, np.uint8)
        prm = OPFrameParams(me, subme, 16, qp, 128, 1, 256 if l0 is not None else 0, inter)
        o.xo_p_frame_part(C.byref(go), ptr(host_slots[k + 1]), ptr(host_slots[k]), ptr(recon), C.byref(prm), vp(lmv[k]), vp(l0),
                          *[vp(res[key]) for key in keys])
        res["recon"] = recon
        return res

    def compare(got, want, tag):
        for key in keys:
            if not np.array_equal(got[key], want[key]):
                d = np.flatnonzero((got[key].reshape(nmb, -1) != want[key].reshape(nmb, -1)).any(1))
                raise AssertionError(f"{tag}: {key} differs at macroblocks {d[:8]} ({d.size} in all): "
                                     f"{got[key][d[0]].ravel()[:8]} vs {want[key][d[0]].ravel()[:8]}")
        lo, co = g.luma_origin, g.slot_chroma_off + g.chroma_origin
        for off, rows, stride in ((lo, g.luma_h, g.luma_stride), (co, g.luma_h // 2, g.chroma_stride)):
            a = got["recon"][off:][: rows * stride].reshape(rows, stride)[:, : g.luma_w]
            bb = want["recon"][off:][: rows * stride].reshape(rows, stride)[:, : g.luma_w]
            assert np.array_equal(a, bb), f"{tag}: reconstruction differs"

    got = device_run(0, n, None)
    seen = set()
    wants = []
    for k in range(n):
        want = oracle_run(k, None)
        wants.append(want)
        compare({key: v[k] for key, v in got.items()}, want, f"{w}x{h} frame {k + 1} (batched launch)")
        seen |= set(np.unique(want["partition"][want["mb_type"] != pkg.MB_P_SKIP]).tolist())
    if inter:
        assert seen >= {13, 16} and (seen & {14, 15}), f"partitions exercised: {seen}"
    else:
        assert seen == {16}
    if n >= 2:
        l0_dev = torch.from_numpy(wants[0]["mvr"]).cuda().reshape(1, nmb, 2)
        got2 = device_run(1, 1, l0_dev)
        compare({key: v[0] for key, v in got2.items()}, oracle_run(1, wants[0]["mvr"]), f"{w}x{h} frame 2 with temporal candidates")


def test_p_frames_dev_refuses_partitions(pkg, ctx):
    """one vector per macroblock cannot describe a partitioned macroblock: analyse_inter != 0 goes through _part_dev"""
    import torch
    g = pkg.geometry(64, 48)
    z = torch.zeros(2 * g.slot_bytes, dtype=torch.uint8, device="cuda")
    t = torch.zeros(g.mb_count * 400, dtype=torch.int16, device="cuda")
    with pytest.raises(pkg.X264DspError):
        ctx.p_frames(g, z, z, z, 1, pkg.PFrameParams(1, 2, 16, 26, 128, 1, 0, 1), None, None, t, t, t, t, t, t)
