"""Pins the oracle's P-slice macroblock loop (oracle/xo_pframe.c: x264_macroblock_analyse + x264_macroblock_encode for
every macroblock of a P frame) against the RUNNING reference encoder: real clips are encoded by
oracle/_ref (unmodified reference), an observer at the end of every frame's macroblock loop captures what the encoder
worked from (reference frame planes, lookahead vectors, the reference frame's 16x16 vectors, slice QP, POCs) and what
it decided (macroblock types, vectors, mvr, cbp, reconstruction), and xo_p_frame must reproduce every P frame from the
captured inputs.  With the in-loop filter off the reconstruction is compared as well."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr

OBS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class Capture(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("slice_type", "qp", "poc", "ref_poc", "inv_ref_poc", "ref_is_inter", "mv_range",
                                         "b4_stride", "have_lowres_mv", "mb_count", "fast_pskip", "i_frame", "frame_type")] + \
               [(n, C.c_void_p) for n in ("fenc", "fref", "fdec", "mb_type", "mvr", "cbp", "mv4x4", "lowres_mv", "l0_mv16",
                                          "partition", "nnz", "mvd")] + \
               [(n, C.c_int32) for n in ("keyint_max", "keyint_min", "scenecut", "icost", "pcost", "pad1")] + \
               [(n, C.c_void_p) for n in ("i4_edge_modes", "chroma_pred_mode")]


class PFrameParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("me_method", "subpel_refine", "me_range", "qp", "mv_range", "fast_pskip", "mvc_scale", "analyse_inter")]


def view(addr, count, dtype):
    nbytes = count * np.dtype(dtype).itemsize
    return np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype).copy()


def capture_encode(w, h, n, cut, me, subme, qp, deblock, keyint=None, light=False, psub=0):
    """encode the clip with the reference; returns (geometry, clip frames, [captured frame dicts])"""
    lib = cc.ref()
    assert lib is not None, "oracle/_ref/libx264ref.so not built"
    lib.xref_open_ex.restype = C.c_void_p
    lib.xref_frame_ptr.restype = C.c_void_p
    lib.xref_frame_ptr.argtypes = [C.c_void_p, C.c_int]
    g = cc.oracle_geom(w, h)
    frames = [cc.synth_frame(w, h, i, cut_frame=cut) for i in range(n)]
    clip = np.concatenate(frames)
    lib.xref_set_keyint(*(keyint or (0, 0, 0)))
    enc = C.c_void_p(lib.xref_open_ex(w, h, me, subme, 16, qp, psub, deblock))
    lib.xref_set_keyint(0, 0, 0)
    assert enc.value
    got = []
    lps, cps = g.luma_plane_size, g.chroma_plane_size

    @OBS_CB
    def observe(hv, frame):
        c = Capture()
        lib.xref_capture_frame(C.c_void_p(hv), C.byref(c))
        nmb = c.mb_count
        d = {k: getattr(c, k) for k in ("slice_type", "qp", "poc", "ref_poc", "inv_ref_poc", "ref_is_inter", "mv_range",
                                        "have_lowres_mv", "fast_pskip", "i_frame", "frame_type")}
        d.update(keyint_max=c.keyint_max, keyint_min=c.keyint_min, scenecut=c.scenecut, icost=c.icost, pcost=c.pcost)
        if light:                       # frame-level facts only (tests/test_gop.py)
            got.append(d)
            return
        d["mb_type"] = view(c.mb_type, nmb, np.int8)
        d["i4_edge_modes"] = view(c.i4_edge_modes, nmb * 8, np.int8).reshape(nmb, 8)
        d["chroma_pred_mode"] = view(c.chroma_pred_mode, nmb, np.int8)
        d["cbp"] = view(c.cbp, nmb, np.int16)
        d["mvr"] = view(c.mvr, nmb * 2, np.int16).reshape(nmb, 2)
        mv4 = view(c.mv4x4, c.b4_stride * g.mb_h * 4 * 2, np.int16).reshape(g.mb_h * 4, c.b4_stride, 2)
        d["mv"] = mv4[0::4, 0:4 * g.mb_w:4].reshape(nmb, 2).copy()
        d["mv4_uniform"] = all(np.array_equal(mv4[dy::4, dx:4 * g.mb_w:4].reshape(nmb, 2), d["mv"])
                               for dy in range(4) for dx in range(4))
        d["mv8"] = np.stack([mv4[2 * (k >> 1)::4, 2 * (k & 1):4 * g.mb_w:4].reshape(nmb, 2) for k in range(4)], axis=1)
        d["mv8_uniform"] = all(np.array_equal(mv4[2 * (k >> 1) + dy::4, 2 * (k & 1) + dx:4 * g.mb_w:4].reshape(nmb, 2), d["mv8"][:, k])
                               for k in range(4) for dy in range(2) for dx in range(2))
        d["partition"] = view(c.partition, nmb, np.uint8)
        d["lowres_mv"] = view(c.lowres_mv, nmb * 2, np.int16) if c.have_lowres_mv else None
        d["mvd_ctx"] = view(c.mvd, nmb * 16, np.uint8).reshape(nmb, 8, 2) if c.mvd else None
        d.update(keyint_max=c.keyint_max, keyint_min=c.keyint_min, scenecut=c.scenecut, icost=c.icost, pcost=c.pcost)
        if c.fref:
            slot = np.zeros(g.slot_bytes, np.uint8)
            slot[: 4 * lps] = view(lib.xref_frame_ptr(c.fref, 10), 4 * lps, np.uint8)
            slot[g.slot_chroma_off: g.slot_chroma_off + cps] = view(lib.xref_frame_ptr(c.fref, 11), cps, np.uint8)
            d["fref_slot"] = slot
            d["l0_mv16"] = view(c.l0_mv16, nmb * 2, np.int16) if c.ref_is_inter else None
        d["recon_y"] = view(lib.xref_frame_ptr(c.fdec, 10), lps, np.uint8)
        d["recon_c"] = view(lib.xref_frame_ptr(c.fdec, 11), cps, np.uint8)
        got.append(d)

    lib.xref_set_observer(observe)
    try:
        out = np.zeros(1 << 22, np.uint8)
        size = lib.xref_encode_clip(enc, ptr(clip), n, ptr(out), out.size)
        assert size > 0
    finally:
        lib.xref_set_observer(OBS_CB())
    return g, frames, got


def interior(g, luma_plane, chroma_plane):
    y = luma_plane[g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
    c = chroma_plane[g.chroma_origin:][: (g.luma_h // 2) * g.chroma_stride].reshape(g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
    return y, c


def run_oracle_pframe(g, frames, d, me, subme):
    o = cc.oracle()
    nmb = g.mb_count
    fenc = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frames[d["i_frame"]]), ptr(fenc))
    recon = np.zeros(g.slot_bytes, np.uint8)
    prm = PFrameParams(me, subme, 16, d["qp"], d["mv_range"], d["fast_pskip"],
                       (d["poc"] - d["ref_poc"]) * d["inv_ref_poc"] if d["l0_mv16"] is not None else 0)
    res = {"mb_type": np.zeros(nmb, np.int8), "mv": np.zeros((nmb, 2), np.int16), "mvr": np.zeros((nmb, 2), np.int16),
           "mvd": np.zeros((nmb, 2), np.int16),
           "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
    lm, l0 = d["lowres_mv"], d["l0_mv16"]
    o.xo_p_frame(C.byref(g), ptr(fenc), ptr(d["fref_slot"]), ptr(recon), C.byref(prm),
                 lm.ctypes.data_as(C.c_void_p) if lm is not None else None,
                 l0.ctypes.data_as(C.c_void_p) if l0 is not None else None,
                 res["mb_type"].ctypes.data_as(C.c_void_p), res["mv"].ctypes.data_as(C.c_void_p),
                 res["mvr"].ctypes.data_as(C.c_void_p), res["mvd"].ctypes.data_as(C.c_void_p),
                 res["levels"].ctypes.data_as(C.c_void_p),
                 res["nnz"].ctypes.data_as(C.c_void_p), res["cbp"].ctypes.data_as(C.c_void_p))
    res["recon"] = recon
    return res


SLICE_TYPE_P = 0


@pytest.mark.parametrize("w,h,n,cut,me,subme,qp,deblock", [
    (176, 144, 6, -1, 0, 1, 26, 0), (352, 288, 8, 5, 1, 2, 26, 0), (208, 160, 7, 3, 1, 5, 30, 0),
    (176, 144, 6, 4, 0, 3, 22, 0), (352, 288, 6, -1, 1, 4, 36, 0), (208, 160, 6, 2, 0, 1, 26, 1), (352, 288, 6, -1, 1, 5, 28, 1)])
def test_p_frame_oracle_reproduces_the_encoder(w, h, n, cut, me, subme, qp, deblock):
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    g, frames, got = capture_encode(w, h, n, cut, me, subme, qp, deblock)
    assert len(got) == n
    p_frames = [d for d in got if d["slice_type"] == SLICE_TYPE_P]
    assert len(p_frames) >= n - 2, [d["slice_type"] for d in got]
    n_skip = n_l0 = 0
    for d in p_frames:
        assert d["mv4_uniform"], "a macroblock with more than one vector in a P16x16-only encode"
        assert set(np.unique(d["mb_type"])) <= {4, 6}, np.unique(d["mb_type"])
        res = run_oracle_pframe(g, frames, d, me, subme)
        tag = f"frame {d['i_frame']} ({w}x{h} me={me} subme={subme} qp={d['qp']} deblock={deblock})"
        bad = np.flatnonzero(res["mb_type"] != d["mb_type"])
        assert bad.size == 0, f"{tag}: type differs at macroblocks {bad[:8]}: {res['mb_type'][bad[:8]]} vs {d['mb_type'][bad[:8]]}"
        assert np.array_equal(res["mv"], d["mv"]), f"{tag}: final vectors differ at {np.flatnonzero((res['mv'] != d['mv']).any(1))[:8]}"
        assert np.array_equal(res["mvr"], d["mvr"]), f"{tag}: mvr differs at {np.flatnonzero((res['mvr'] != d['mvr']).any(1))[:8]}"
        # the entropy coder's hand-off (8(f) N3): what x264_cabac_mvd wrote and kept as context for the neighbours
        want_ctx = np.minimum(np.abs(res["mvd"].astype(np.int32)), 66).astype(np.uint8)
        for k in range(7):              # bottom row (4) and right column (3) of the macroblock; the eighth entry is padding
            assert np.array_equal(d["mvd_ctx"][:, k, :], want_ctx), f"{tag}: mvd context differs"
        coded = d["mb_type"] != 6
        assert np.array_equal(res["cbp"][coded], d["cbp"][coded]), f"{tag}: cbp differs"
        if not deblock:
            ry, rc = interior(g, res["recon"][: g.luma_plane_size], res["recon"][g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size])
            wy, wc = interior(g, d["recon_y"], d["recon_c"])
            assert np.array_equal(ry, wy), f"{tag}: luma reconstruction differs"
            assert np.array_equal(rc, wc), f"{tag}: chroma reconstruction differs"
        n_skip += int((d["mb_type"] == 6).sum())
        n_l0 += int((d["mb_type"] == 4).sum())
    assert n_skip > 0 and n_l0 > 0, f"one-sided clip: {n_skip} skipped, {n_l0} coded macroblocks"


def run_oracle_pframe_part(g, frames, d, me, subme, inter=1):
    o = cc.oracle()
    nmb = g.mb_count
    fenc = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frames[d["i_frame"]]), ptr(fenc))
    recon = np.zeros(g.slot_bytes, np.uint8)
    prm = PFrameParams(me, subme, 16, d["qp"], d["mv_range"], d["fast_pskip"],
                       (d["poc"] - d["ref_poc"]) * d["inv_ref_poc"] if d["l0_mv16"] is not None else 0, inter)
    res = {"mb_type": np.zeros(nmb, np.int8), "partition": np.zeros(nmb, np.uint8), "mv8": np.zeros((nmb, 4, 2), np.int16),
           "mvr": np.zeros((nmb, 2), np.int16), "mvd8": np.zeros((nmb, 4, 2), np.int16),
           "levels": np.zeros((nmb, 392), np.int16), "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
    lm, l0 = d["lowres_mv"], d["l0_mv16"]
    o.xo_p_frame_part(C.byref(g), ptr(fenc), ptr(d["fref_slot"]), ptr(recon), C.byref(prm),
                      lm.ctypes.data_as(C.c_void_p) if lm is not None else None,
                      l0.ctypes.data_as(C.c_void_p) if l0 is not None else None,
                      *[res[k].ctypes.data_as(C.c_void_p) for k in ("mb_type", "partition", "mv8", "mvr", "mvd8", "levels", "nnz", "cbp")])
    res["recon"] = recon
    return res


# h->mb.mvd[mb][0..6] holds the bottom row (4x4 cells (0..3, 3)) and the right column (cells (3, 0..2)) of the macroblock:
# the 8x8 blocks those cells lie in
MVD_CTX_BLOCK = [2, 2, 3, 3, 1, 1, 3]


def check_part_frame(g, d, res, tag, deblock):
    bad = np.flatnonzero(res["mb_type"] != d["mb_type"])
    assert bad.size == 0, f"{tag}: type differs at macroblocks {bad[:8]}: {res['mb_type'][bad[:8]]} vs {d['mb_type'][bad[:8]]}"
    coded = d["mb_type"] != 6
    assert np.array_equal(res["partition"][coded], d["partition"][coded]), \
        f"{tag}: partition differs at {np.flatnonzero(coded & (res['partition'] != d['partition']))[:8]}"
    assert np.array_equal(res["mv8"], d["mv8"]), f"{tag}: final vectors differ at {np.flatnonzero((res['mv8'] != d['mv8']).any((1, 2)))[:8]}"
    assert np.array_equal(res["mvr"], d["mvr"]), f"{tag}: mvr differs"
    want_ctx = np.minimum(np.abs(res["mvd8"].astype(np.int32)), 66).astype(np.uint8)
    for k, blk in enumerate(MVD_CTX_BLOCK):
        assert np.array_equal(d["mvd_ctx"][:, k, :], want_ctx[:, blk, :]), f"{tag}: mvd context {k} differs"
    assert np.array_equal(res["cbp"][coded], d["cbp"][coded]), f"{tag}: cbp differs"
    if not deblock:
        ry, rc = interior(g, res["recon"][: g.luma_plane_size], res["recon"][g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size])
        wy, wc = interior(g, d["recon_y"], d["recon_c"])
        assert np.array_equal(ry, wy), f"{tag}: luma reconstruction differs"
        assert np.array_equal(rc, wc), f"{tag}: chroma reconstruction differs"


@pytest.mark.parametrize("w,h,n,cut,me,subme,qp,deblock", [
    (176, 144, 6, -1, 0, 1, 26, 0), (352, 288, 8, 5, 1, 2, 26, 0), (208, 160, 7, 3, 1, 5, 30, 0),
    (176, 144, 6, 4, 0, 3, 22, 0), (352, 288, 6, -1, 1, 4, 20, 0), (352, 288, 6, -1, 1, 5, 28, 1)])
def test_p_frame_partitions_reproduce_the_encoder(w, h, n, cut, me, subme, qp, deblock):
    """analyse.inter = PSUB16x16: P8x8 / P16x8 / P8x16 as well (xo_p_frame_part)"""
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    g, frames, got = capture_encode(w, h, n, cut, me, subme, qp, deblock, psub=1)
    p_frames = [d for d in got if d["slice_type"] == SLICE_TYPE_P]
    assert len(p_frames) >= n - 2
    seen = set()
    for d in p_frames:
        assert d["mv8_uniform"], "a sub-8x8 partition in a PSUB16x16 encode"
        assert set(np.unique(d["mb_type"])) <= {4, 5, 6}, np.unique(d["mb_type"])
        res = run_oracle_pframe_part(g, frames, d, me, subme)
        check_part_frame(g, d, res, f"frame {d['i_frame']} ({w}x{h} me={me} subme={subme} qp={d['qp']} deblock={deblock})", deblock)
        seen |= set(np.unique(d["partition"][d["mb_type"] != 6]).tolist())
    assert seen >= {13, 16} and (seen & {14, 15}), f"clip exercises partitions {seen} only"


def test_p_frame_part_without_partitions_is_the_16x16_loop():
    """xo_p_frame_part( analyse_inter = 0 ) == xo_p_frame"""
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    g, frames, got = capture_encode(208, 160, 5, 3, 1, 5, 28, 0)
    for d in [d for d in got if d["slice_type"] == SLICE_TYPE_P]:
        a = run_oracle_pframe(g, frames, d, 1, 5)
        b = run_oracle_pframe_part(g, frames, d, 1, 5, inter=0)
        assert np.array_equal(a["mb_type"], b["mb_type"]) and np.array_equal(a["cbp"], b["cbp"])
        assert all(np.array_equal(b["mv8"][:, k], a["mv"]) for k in range(4))
        assert all(np.array_equal(b["mvd8"][:, k], a["mvd"]) for k in range(4))
        assert np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["nnz"], b["nnz"]) and np.array_equal(a["recon"], b["recon"])
