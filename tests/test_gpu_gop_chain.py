"""x264dsp_gops_encode_dev: closed GOPs coded on the device from the I frame's first macroblock to the last frame's reference
planes -- slice kernels, boundary strengths, in-loop filter, border, half-pel planes, one launch of each per GOP position --
against the RUNNING reference encoder with its in-loop filter on: real clips are encoded by oracle/_ref (unmodified reference),
the observer of tests/test_oracle_pframe.py captures every frame's decisions and, with the next frame, the previous frame's
FINAL reference planes (deblocked, expanded, filtered).  The device chain gets the pictures, the slice QPs and the lookahead's
vectors and must reproduce all of it, frame after frame, each frame predicting from the device's own previous output."""
import numpy as np
import pytest

import cpu_checkers as cc
from test_oracle_pframe import capture_encode

pytestmark = pytest.mark.gpu


def padded(g, plane, stride, origin, w, rows, pad_h, pad_v):
    """the picture area of a plane with pad_v rows / pad_h bytes of its border on every side"""
    r0 = origin // stride - pad_v
    c0 = origin % stride - pad_h
    return plane[: (r0 + rows + 2 * pad_v) * stride].reshape(-1, stride)[r0: r0 + rows + 2 * pad_v, c0: c0 + w + 2 * pad_h]


@pytest.mark.parametrize("w,h,n,me,subme,qp,psub,n_gops", [
    (176, 144, 5, 0, 1, 26, 0, 1), (352, 288, 5, 1, 5, 28, 0, 2), (208, 160, 6, 1, 4, 24, 1, 2), (352, 288, 4, 1, 2, 32, 1, 1)])
def test_gop_chain_reproduces_the_encoder(pkg, ctx, w, h, n, me, subme, qp, psub, n_gops):
    import torch
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    g, frames, got = capture_encode(w, h, n, -1, me, subme, qp, 1, psub=psub)
    assert len(got) == n and got[0]["slice_type"] != 0 and all(d["slice_type"] == 0 for d in got[1:]), [d["slice_type"] for d in got]
    nmb = g.mb_count
    dg = pkg.geometry(w, h)
    # position-major staging: frame t of every GOP next to each other (the GOPs are copies of the same clip)
    pics = np.stack([frames[d["i_frame"]] for d in got for _ in range(n_gops)])
    fenc = torch.zeros(n * n_gops * dg.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(dg, torch.from_numpy(pics).cuda(), fenc, n * n_gops)
    recon = torch.zeros_like(fenc)
    lmv = np.zeros((n, n_gops, nmb, 2), np.int16)
    for t, d in enumerate(got):
        if t == 0 or d["lowres_mv"] is None:
            lmv[t, :, 0, 0] = 0x7fff                                   # "the lookahead has no vectors for this pair"
        else:
            lmv[t, :] = d["lowres_mv"].reshape(nmb, 2)
    N = n * n_gops
    out = {"mb_type": torch.full((N, nmb), -1, dtype=torch.int8, device="cuda"),
           "partition": torch.zeros((N, nmb), dtype=torch.uint8, device="cuda"),
           "mv8": torch.zeros((N, nmb, 4, 2), dtype=torch.int16, device="cuda"), "mvr": torch.zeros((N, nmb, 2), dtype=torch.int16, device="cuda"),
           "mvd8": torch.zeros((N, nmb, 4, 2), dtype=torch.int16, device="cuda"),
           "levels": torch.zeros((N, nmb, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda"),
           "nnz": torch.zeros((N, nmb, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda"),
           "cbp": torch.zeros((N, nmb), dtype=torch.int16, device="cuda"),
           "mode16": torch.zeros((n_gops, nmb), dtype=torch.uint8, device="cuda"), "chroma_mode": torch.zeros((n_gops, nmb), dtype=torch.uint8, device="cuda"),
           "modes4": torch.zeros((n_gops, nmb, 16), dtype=torch.uint8, device="cuda"), "luma_dc": torch.zeros((n_gops, nmb, 16), dtype=torch.int16, device="cuda")}
    prm = pkg.GopEncodeParams(me, subme, 16, got[0]["qp"], got[1]["qp"], got[1]["mv_range"], got[1]["fast_pskip"], psub, 1, 0, 0)
    ctx.gops_encode(dg, fenc, recon, n_gops, n, prm, torch.from_numpy(lmv).cuda(), out)
    ctx.sync()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    rec = recon.cpu().numpy().reshape(n, n_gops, dg.slot_bytes)
    lps, cps = g.luma_plane_size, g.chroma_plane_size
    pad_v, pad_h = g.luma_origin // g.luma_stride, g.luma_origin % g.luma_stride
    seen_parts = set()
    for t, d in enumerate(got):
        for gop in range(n_gops):
            k = t * n_gops + gop
            tag = f"{w}x{h} me={me} subme={subme} psub={psub}: frame {t} of GOP {gop}"
            assert np.array_equal(res["mb_type"][k], d["mb_type"]), f"{tag}: macroblock types differ at {np.flatnonzero(res['mb_type'][k] != d['mb_type'])[:8]}"
            if t > 0:
                coded = d["mb_type"] != 6
                assert np.array_equal(res["mv8"][k], d["mv8"]), f"{tag}: vectors differ at {np.flatnonzero((res['mv8'][k] != d['mv8']).any((1, 2)))[:8]}"
                assert np.array_equal(res["mvr"][k], d["mvr"]), f"{tag}: mvr differs"
                assert np.array_equal(res["partition"][k][coded], d["partition"][coded]), f"{tag}: partitions differ"
                assert np.array_equal(res["cbp"][k][coded], d["cbp"][coded]), f"{tag}: cbp differs"
                seen_parts |= set(np.unique(d["partition"][coded]).tolist())
            if t + 1 < n:
                # the encoder's reference for frame t + 1 is what the device must hold for frame t: all four luma planes with
                # their borders, and the chroma plane
                want = got[t + 1]["fref_slot"]
                for p, name in enumerate("N H V HV".split()):
                    a = padded(g, rec[t, gop][p * lps: (p + 1) * lps], g.luma_stride, g.luma_origin, g.luma_w, g.luma_h, pad_h, pad_v)
                    b = padded(g, want[p * lps: (p + 1) * lps], g.luma_stride, g.luma_origin, g.luma_w, g.luma_h, pad_h, pad_v)
                    assert np.array_equal(a, b), f"{tag}: reference plane {name} differs in {np.count_nonzero(a != b)} bytes"
                co = g.slot_chroma_off
                cpad_v, cpad_h = g.chroma_origin // g.chroma_stride, g.chroma_origin % g.chroma_stride
                a = padded(g, rec[t, gop][co: co + cps], g.chroma_stride, g.chroma_origin, g.luma_w, g.luma_h // 2, cpad_h, cpad_v)
                b = padded(g, want[co: co + cps], g.chroma_stride, g.chroma_origin, g.luma_w, g.luma_h // 2, cpad_h, cpad_v)
                assert np.array_equal(a, b), f"{tag}: chroma reference plane differs in {np.count_nonzero(a != b)} bytes"
    if psub:
        assert seen_parts >= {13, 16}, seen_parts


@pytest.mark.parametrize("w,h,n_gops,gop_len,me,subme,psub", [(208, 160, 5, 4, 1, 5, 1), (352, 288, 3, 3, 0, 1, 0)])
def test_gops_encode_host_matches_the_device_chain(pkg, ctx, w, h, n_gops, gop_len, me, subme, psub):
    """x264dsp_gops_encode_host (pictures in host memory, [gop][t]; one GOP position after the other through the copy / kernel
    pipeline; compact levels) against the same stages called one by one on device memory"""
    import torch
    from test_gpu_host_paths import expand_packed, mask_dense
    g = pkg.geometry(w, h)
    nmb = g.mb_count
    n = n_gops * gop_len
    pics = np.stack([pkg.synth_frame(w, h, 7 * gop + t, cut_frame=-1) for gop in range(n_gops) for t in range(gop_len)])     # [gop][t]
    prm = pkg.GopEncodeParams(me, subme, 16, 23, 26, 128, 1, psub, 1, 0, 0)
    shapes = {"mb_type": ((n, nmb), np.int8), "partition": ((n, nmb), np.uint8), "mv8": ((n, nmb, 4, 2), np.int16),
              "mvr": ((n, nmb, 2), np.int16), "mvd8": ((n, nmb, 4, 2), np.int16), "nnz": ((n, nmb, 27), np.uint8), "cbp": ((n, nmb), np.int16),
              "mode16": ((n_gops, nmb), np.uint8), "chroma_mode": ((n_gops, nmb), np.uint8), "modes4": ((n_gops, nmb, 16), np.uint8),
              "luma_dc": ((n_gops, nmb, 16), np.int16)}
    out = {k: np.zeros(s, t) for k, (s, t) in shapes.items()}
    packed = np.full(n * nmb * 392 // 2, 999, np.int16)
    f_off, f_size, mb_off = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros((n, nmb), np.int32)
    ctx.gops_encode_host(w, h, n_gops, gop_len, pics, prm, out, packed, f_off, f_size, mb_off)
    # ---- the same, stage by stage, position-major
    order = [gop * gop_len + t for t in range(gop_len) for gop in range(n_gops)]
    fenc = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(pics[order]).cuda(), fenc, n)
    ctx.frame_expand_border(g, fenc, n)
    ctx.frame_init_lowres(g, fenc, n)
    b = np.arange(n_gops, n, dtype=np.int32)
    d_lmv = torch.zeros((n, nmb, 2), dtype=torch.int16, device="cuda")
    d_lc = torch.zeros((n, nmb), dtype=torch.int32, device="cuda")
    d_ls = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    ctx.lookahead_frame_cost(g, fenc, b, b - n_gops, np.zeros(b.size, np.uint8), d_lmv[n_gops:], d_lc[n_gops:], d_ls[n_gops:])
    dev = {k: torch.zeros(s, dtype=getattr(torch, np.dtype(t).name), device="cuda") for k, (s, t) in shapes.items()}
    dev["levels"] = torch.zeros((n, nmb, 392), dtype=torch.int16, device="cuda")
    recon = torch.zeros_like(fenc)
    ctx.gops_encode(g, fenc, recon, n_gops, gop_len, prm, d_lmv, dev)
    ctx.sync()
    for k in shapes:
        assert np.array_equal(out[k], dev[k].cpu().numpy()), k
    levels, nnz = dev["levels"].cpu().numpy(), out["nnz"]
    assert (out["mb_type"][:n_gops] <= 3).all() and (out["mb_type"][n_gops:] >= 4).all()          # position 0 intra, the rest inter
    used = np.zeros(packed.size, bool)
    for k in range(n):
        stream = packed[f_off[k]: f_off[k] + f_size[k]]
        assert np.array_equal(expand_packed(stream, mb_off[k], nnz[k]), mask_dense(levels[k], nnz[k])), f"frame {k}: compact levels"
        assert not used[f_off[k]: f_off[k] + f_size[k]].any(), "two frames share bytes of the stream"
        used[f_off[k]: f_off[k] + f_size[k]] = True
    assert (packed[~used] == 999).all()
