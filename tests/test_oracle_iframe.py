"""Pins the oracle's I-slice macroblock loop (oracle/xo_iframe.c: x264_mb_analyse_intra + x264_mb_analyse_intra_chroma + the
I-slice decision + x264_macroblock_encode's intra branches for every macroblock) against the RUNNING reference
encoder: the observer of tests/test_oracle_pframe.py captures every I frame of real encodes -- macroblock types, the 4x4
modes each macroblock shows its neighbours, chroma modes, cbp and (in-loop filter off) the reconstruction -- and xo_i_frame
must reproduce them from the source frame alone."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr
from test_oracle_pframe import capture_encode, interior

SLICE_TYPE_I = 2


def run_oracle_iframe(g, frame, qp):
    o = cc.oracle()
    nmb = g.mb_count
    fenc = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frame), ptr(fenc))
    res = {"recon": np.zeros(g.slot_bytes, np.uint8), "mb_type": np.zeros(nmb, np.int8), "mode16": np.zeros(nmb, np.uint8),
           "chroma_mode": np.zeros(nmb, np.uint8), "modes4": np.zeros((nmb, 16), np.uint8),
           "levels": np.zeros((nmb, 392), np.int16), "luma_dc": np.zeros((nmb, 16), np.int16),
           "nnz": np.zeros((nmb, 27), np.uint8), "cbp": np.zeros(nmb, np.int16)}
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    o.xo_i_frame(C.byref(g), ptr(fenc), ptr(res["recon"]), qp, vp(res["mb_type"]), vp(res["mode16"]), vp(res["chroma_mode"]),
                 vp(res["modes4"]), vp(res["levels"]), vp(res["luma_dc"]), vp(res["nnz"]), vp(res["cbp"]))
    return res


@pytest.mark.parametrize("w,h,n,cut,subme,qp,deblock,keyint", [
    (176, 144, 3, -1, 1, 26, 0, None), (352, 288, 5, 2, 2, 30, 0, (2, 1, 40)), (208, 160, 4, -1, 5, 22, 0, (3, 1, 0)),
    (352, 288, 3, -1, 4, 38, 0, None), (176, 144, 4, -1, 3, 18, 1, (2, 2, 0))])
def test_i_frame_oracle_reproduces_the_encoder(w, h, n, cut, subme, qp, deblock, keyint):
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    g, frames, got = capture_encode(w, h, n, cut, 1, subme, qp, deblock, keyint=keyint)
    i_frames = [d for d in got if d["slice_type"] == SLICE_TYPE_I]
    assert i_frames, "no I frame in the clip"
    n4 = n16 = 0
    for d in i_frames:
        res = run_oracle_iframe(g, frames[d["i_frame"]], d["qp"])
        tag = f"frame {d['i_frame']} ({w}x{h} subme={subme} qp={d['qp']} deblock={deblock})"
        bad = np.flatnonzero(res["mb_type"] != d["mb_type"])
        assert bad.size == 0, f"{tag}: type differs at macroblocks {bad[:8]}: {res['mb_type'][bad[:8]]} vs {d['mb_type'][bad[:8]]}"
        edge = res["modes4"][:, [10, 11, 14, 15, 5, 7, 13]].astype(np.int8)
        bad = np.flatnonzero((edge != d["i4_edge_modes"][:, :7]).any(1))
        assert bad.size == 0, f"{tag}: 4x4 modes differ at {bad[:8]}: {edge[bad[:3]]} vs {d['i4_edge_modes'][bad[:3], :7]}"
        fix = np.array([0, 1, 2, 3, 0, 0, 0], np.int8)
        assert np.array_equal(fix[res["chroma_mode"]], d["chroma_pred_mode"]), f"{tag}: chroma modes differ"
        assert np.array_equal(res["cbp"], d["cbp"]), f"{tag}: cbp differs at {np.flatnonzero(res['cbp'] != d['cbp'])[:8]}"
        if not deblock:
            ry, rc = interior(g, res["recon"][: g.luma_plane_size], res["recon"][g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size])
            wy, wc = interior(g, d["recon_y"], d["recon_c"])
            assert np.array_equal(ry, wy), f"{tag}: luma reconstruction differs"
            assert np.array_equal(rc, wc), f"{tag}: chroma reconstruction differs"
        n4 += int((d["mb_type"] == 0).sum())
        n16 += int((d["mb_type"] == 2).sum())
    assert n4 > 0 and n16 > 0, f"one-sided clip: {n4} I4x4, {n16} I16x16 macroblocks"
