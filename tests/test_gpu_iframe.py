"""x264dsp_i_frames_dev (x264_mb_analyse_intra + the I16x16 / I4x4 decision + x264_mb_analyse_intra_chroma +
x264_macroblock_encode's intra branches for every macroblock of an I frame, as a wavefront on the device) against the CPU
oracle's xo_i_frame, which tests/test_oracle_iframe.py pins to the running reference encoder.  Compared bit for bit:
macroblock types, 16x16 / 4x4 / chroma modes, levels, luma DC levels, nnz, cbp and the reconstruction."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,n,qp,cut", [(176, 144, 2, 26, -1), (352, 288, 3, 18, 1), (208, 160, 2, 38, -1), (112, 96, 3, 30, -1),
                                          (1920, 1080, 1, 24, -1)])
def test_i_frames_match_oracle(pkg, ctx, w, h, n, qp, cut):
    import torch
    o = cc.oracle()
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    nmb = g.mb_count
    frames = np.stack([pkg.synth_frame(w, h, 3 * i, cut_frame=cut) for i in range(n)])
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(frames).cuda(), slots, n)
    ctx.sync()
    host_slots = slots.cpu().numpy().reshape(n, g.slot_bytes)
    out = {"mb_type": torch.full((n, nmb), -1, dtype=torch.int8, device="cuda"),
           "mode16": torch.full((n, nmb), 99, dtype=torch.uint8, device="cuda"),
           "chroma_mode": torch.full((n, nmb), 99, dtype=torch.uint8, device="cuda"),
           "modes4": torch.full((n, nmb, 16), 99, dtype=torch.uint8, device="cuda"),
           "levels": torch.ones((n, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
           "luma_dc": torch.ones((n, nmb, 16), dtype=torch.int16, device="cuda"),
           "nnz": torch.ones((n, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
           "cbp": torch.full((n, nmb), -1, dtype=torch.int16, device="cuda")}
    recon = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.i_frames(g, slots, recon, n, qp, out["mb_type"], out["mode16"], out["chroma_mode"], out["modes4"], out["levels"],
                 out["luma_dc"], out["nnz"], out["cbp"])
    ctx.sync()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    grec = recon.cpu().numpy().reshape(n, g.slot_bytes)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n4 = n16 = 0
    for k in range(n):
        want = {"mb_type": np.zeros(nmb, np.int8), "mode16": np.zeros(nmb, np.uint8), "chroma_mode": np.zeros(nmb, np.uint8),
                "modes4": np.zeros((nmb, 16), np.uint8), "levels": np.zeros((nmb, 392), np.int16),
                "luma_dc": np.zeros((nmb, 16), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
        wrec = np.zeros(g.slot_bytes, np.uint8)
        o.xo_i_frame(C.byref(go), ptr(host_slots[k]), ptr(wrec), qp, vp(want["mb_type"]), vp(want["mode16"]), vp(want["chroma_mode"]),
                     vp(want["modes4"]), vp(want["levels"]), vp(want["luma_dc"]), vp(want["nnz"]), vp(want["cbp"]))
        for key in ("mb_type", "mode16", "chroma_mode", "modes4", "cbp", "nnz", "luma_dc", "levels"):
            if not np.array_equal(got[key][k], want[key]):
                d = np.flatnonzero((got[key][k].reshape(nmb, -1) != want[key].reshape(nmb, -1)).any(1))
                raise AssertionError(f"{w}x{h} frame {k}: {key} differs at macroblocks {d[:8]} ({d.size} in all): "
                                     f"{got[key][k][d[0]].ravel()[:16]} vs {want[key][d[0]].ravel()[:16]}")
        lo, co = g.luma_origin, g.slot_chroma_off + g.chroma_origin
        for (off, rows, tag) in ((lo, g.luma_h, "luma"), (co, g.luma_h // 2, "chroma")):
            stride = g.luma_stride
            a = grec[k][off:][: rows * stride].reshape(rows, stride)[:, : g.luma_w]
            b = wrec[off:][: rows * stride].reshape(rows, stride)[:, : g.luma_w]
            assert np.array_equal(a, b), f"{w}x{h} frame {k}: {tag} reconstruction differs"
        n4 += int((want["mb_type"] == 0).sum())
        n16 += int((want["mb_type"] == 2).sum())
    assert n4 > 0 and n16 > 0, f"one-sided input: {n4} I4x4, {n16} I16x16"
