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
