#!/usr/bin/env python
"""Generates tests/golden/golden_r01_intra.npz from the UNMODIFIED reference (oracle/_ref/libx264ref.so): the
round's additions to the path --

  * x264_macroblock_encode for I16x16 macroblocks of an I slice (x264_mb_encode_i16x16 + intra chroma) on stored
    source blocks and predictions (oracle/ref_shim/harness.c: xref_encode_intra16_mb),
  * all 26 intra predictors of x264_predict_16x16_init / _8x8c_init / _4x4_init on stored neighbourhoods.

Run in the dev container only:  python tests/golden/make_golden_intra.py
Every OUTPUT array comes from the reference's own functions; tests/test_golden.py checks the CPU oracle and the
CUDA path against the file (the GPU box has no /root/reference)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.path.insert(0, os.path.dirname(TESTS))

import cpu_checkers as cc          # noqa: E402
from cpu_checkers import ptr, i16p  # noqa: E402

OUT = os.path.join(HERE, "golden_r01_intra.npz")
PRED_T = C.CFUNCTYPE(None, C.c_void_p)


def main():
    lib = cc.ref()
    assert lib is not None, "make -C oracle ref first"
    enc = cc.RefEncoder(128, 96, me=1, subme=5, me_range=16, qp=26)
    lib.xref_encode_intra16_mb.restype = C.c_int
    G = {}
    # ------------------------------------------------------------ I16x16 macroblocks, 48 per QP (= one 128x96 frame)
    r = np.random.RandomState(1616)
    RQ = [12, 20, 26, 34, 44]
    NMB = 48
    fy, fc = np.zeros((len(RQ), NMB, 16, 16), np.uint8), np.zeros((len(RQ), NMB, 8, 16), np.uint8)
    py, pc = np.zeros((len(RQ), NMB, 16, 32), np.uint8), np.zeros((len(RQ), NMB, 8, 32), np.uint8)
    ry, rc = np.zeros_like(py), np.zeros_like(pc)
    lv, dc = np.zeros((len(RQ), NMB, 392), np.int16), np.zeros((len(RQ), NMB, 16), np.int16)
    nz, cbp = np.zeros((len(RQ), NMB, 27), np.uint8), np.zeros((len(RQ), NMB), np.int32)
    for qi, qp in enumerate(RQ):
        for t in range(NMB):
            amp = [1, 3, 8, 25, 80][t % 5]
            kind = t % 4
            if kind == 0:
                py[qi, t, :, :16] = r.randint(20, 236)
                pc[qi, t, :, :8], pc[qi, t, :, 16:24] = r.randint(20, 236), r.randint(20, 236)
            elif kind == 1:
                py[qi, t, :, :16] = r.randint(0, 256, 16)[None, :]
                pc[qi, t, :, :8], pc[qi, t, :, 16:24] = r.randint(0, 256, 8)[None, :], r.randint(0, 256, 8)[None, :]
            elif kind == 2:
                py[qi, t, :, :16] = r.randint(0, 256, 16)[:, None]
                pc[qi, t, :, :8], pc[qi, t, :, 16:24] = r.randint(0, 256, 8)[:, None], r.randint(0, 256, 8)[:, None]
            else:
                py[qi, t, :, :16] = r.randint(0, 256, (16, 16))
                pc[qi, t, :, :8], pc[qi, t, :, 16:24] = r.randint(0, 256, (8, 8)), r.randint(0, 256, (8, 8))
            f = np.clip(py[qi, t, :, :16].astype(int) + r.randint(-amp, amp + 1, (16, 16)), 0, 255)
            if t % 6 == 0:
                f = np.clip(py[qi, t, :, :16].astype(int) + r.randint(-6, 7), 0, 255)      # DC-only luma change
            fy[qi, t] = f
            fc[qi, t, :, :8] = np.clip(pc[qi, t, :, :8].astype(int) + r.randint(-amp, amp + 1, (8, 8)), 0, 255)
            fc[qi, t, :, 8:] = np.clip(pc[qi, t, :, 16:24].astype(int) + r.randint(-amp, amp + 1, (8, 8)), 0, 255)
            y1, c1 = py[qi, t].copy(), pc[qi, t].copy()
            cbp[qi, t] = lib.xref_encode_intra16_mb(enc.h, ptr(np.ascontiguousarray(fy[qi, t])), ptr(np.ascontiguousarray(fc[qi, t])),
                                                    ptr(y1), ptr(c1), qp, ptr(lv[qi, t], i16p), ptr(dc[qi, t], i16p), ptr(nz[qi, t]))
            ry[qi, t], rc[qi, t] = y1, c1
    G["i16_qps"] = np.array(RQ)
    G["i16_fenc_y"], G["i16_fenc_c"], G["i16_pred_y"], G["i16_pred_c"] = fy, fc, py[..., :16], pc[..., :24]
    G["i16_recon_y"], G["i16_recon_c"] = ry[..., :16], rc[..., :24]
    G["i16_levels"], G["i16_luma_dc"], G["i16_nnz"], G["i16_cbp"] = lv, dc, nz, cbp

    # ------------------------------------------------------------ I4x4 macroblocks, 48 per QP, each with its own neighbourhood
    lib.xref_encode_intra4_mb.restype = C.c_int
    r4 = np.random.RandomState(4444)
    nbh = np.zeros((len(RQ), NMB, 17, 32), np.uint8)          # row 0 = the row above, origin (1, 8)
    f4y, f4c = np.zeros((len(RQ), NMB, 16, 16), np.uint8), np.zeros((len(RQ), NMB, 8, 16), np.uint8)
    p4c, r4c = np.zeros((len(RQ), NMB, 8, 32), np.uint8), np.zeros((len(RQ), NMB, 8, 32), np.uint8)
    r4y = np.zeros((len(RQ), NMB, 16, 16), np.uint8)
    m4, rep = np.zeros((len(RQ), NMB, 16), np.uint8), np.zeros((len(RQ), NMB), np.uint8)
    l4, n4, c4 = np.zeros((len(RQ), NMB, 392), np.int16), np.zeros((len(RQ), NMB, 27), np.uint8), np.zeros((len(RQ), NMB), np.int32)
    for qi, qp in enumerate(RQ):
        for t in range(NMB):
            nb = r4.randint(0, 256, (17, 32)).astype(np.uint8)
            if t % 4 == 0:
                nb[:] = r4.randint(40, 200)
            base = r4.randint(0, 256)
            f4y[qi, t] = (r4.randint(0, 256, (16, 16)) if t % 3 == 0 else
                          np.clip(base + r4.randint(-30, 31, (16, 16)) + np.arange(16)[None, :] * r4.randint(-3, 4), 0, 255))
            p4c[qi, t, :, :8], p4c[qi, t, :, 16:24] = r4.randint(20, 236), r4.randint(20, 236)
            f4c[qi, t, :, :8] = np.clip(p4c[qi, t, :, :8].astype(int) + r4.randint(-9, 10, (8, 8)), 0, 255)
            f4c[qi, t, :, 8:] = np.clip(p4c[qi, t, :, 16:24].astype(int) + r4.randint(-9, 10, (8, 8)), 0, 255)
            m4[qi, t] = r4.randint(0, 12, 16)
            rep[qi, t] = t % 2
            nbh[qi, t] = nb
            y1, c1 = nb.copy(), p4c[qi, t].copy()
            c4[qi, t] = lib.xref_encode_intra4_mb(enc.h, ptr(np.ascontiguousarray(f4y[qi, t])), ptr(np.ascontiguousarray(f4c[qi, t])),
                                                  C.cast(y1.ctypes.data + 32 + 8, C.c_void_p), ptr(c1), qp,
                                                  ptr(np.ascontiguousarray(m4[qi, t])), int(rep[qi, t]), ptr(l4[qi, t], i16p), ptr(n4[qi, t]))
            r4y[qi, t], r4c[qi, t] = y1[1:, 8:24], c1
    G["i4_nbh"], G["i4_fenc_y"], G["i4_fenc_c"], G["i4_pred_c"] = nbh, f4y, f4c, p4c[..., :24]
    G["i4_modes"], G["i4_replicate5"] = m4, rep
    G["i4_recon_y"], G["i4_recon_c"], G["i4_levels"], G["i4_nnz"], G["i4_cbp"] = r4y, r4c[..., :24], l4, n4, c4

    # ------------------------------------------------------------ x264_macroblock_probe_pskip, 96 macroblocks per QP
    lib.xref_probe_pskip_mb.restype = C.c_int
    rp = np.random.RandomState(9292)
    PQ = [14, 22, 26, 32, 40]
    NP = 96
    sk_fy, sk_fc = np.zeros((len(PQ), NP, 16, 16), np.uint8), np.zeros((len(PQ), NP, 8, 16), np.uint8)
    sk_py, sk_pc = np.zeros((len(PQ), NP, 16, 32), np.uint8), np.zeros((len(PQ), NP, 8, 32), np.uint8)
    sk = np.zeros((len(PQ), NP), np.uint8)
    for qi, qp in enumerate(PQ):
        scale = max(1, (qp - 14) // 5)
        for t in range(NP):
            p_y = rp.randint(0, 256, (16, 32)).astype(np.uint8)
            p_c = rp.randint(0, 256, (8, 32)).astype(np.uint8)
            if t % 3 == 0:
                p_y[:], p_c[:] = rp.randint(30, 220), rp.randint(30, 220)
            p_y[:, 16:], p_c[:, 8:16], p_c[:, 24:] = 0, 0, 0
            amp, camp = [0, 1, 2, 3, 5, 8, 14][t % 7] * scale, [0, 1, 2, 4, 9][t % 5] * scale
            f_y = np.clip(p_y[:, :16].astype(int) + rp.randint(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
            f_c = np.zeros((8, 16), np.uint8)
            f_c[:, :8] = np.clip(p_c[:, :8].astype(int) + rp.randint(-camp, camp + 1, (8, 8)), 0, 255)
            f_c[:, 8:] = np.clip(p_c[:, 16:24].astype(int) + rp.randint(-camp, camp + 1, (8, 8)), 0, 255)
            if t % 13 == 0:
                f_c[:, :8] = np.clip(p_c[:, :8].astype(int) + rp.randint(-4, 5), 0, 255)
            sk[qi, t] = lib.xref_probe_pskip_mb(enc.h, ptr(f_y), ptr(f_c), ptr(p_y), ptr(p_c), qp)
            sk_fy[qi, t], sk_fc[qi, t], sk_py[qi, t], sk_pc[qi, t] = f_y, f_c, p_y, p_c
    G["pskip_qps"] = np.array(PQ)
    G["pskip_fenc_y"], G["pskip_fenc_c"], G["pskip_pred_y"], G["pskip_pred_c"], G["pskip_skip"] = sk_fy, sk_fc, sk_py[..., :16], sk_pc[..., :24], sk

    # ------------------------------------------------------------ the 26 predictors on 10 neighbourhoods each
    tabs = [(PRED_T * 7)(), (PRED_T * 7)(), (PRED_T * 12)()]
    lib.x264_predict_16x16_init(0, tabs[0])
    lib.x264_predict_8x8c_init(0, tabs[1])
    lib.x264_predict_4x4_init(0, tabs[2])
    NT = 10
    src = r.randint(0, 256, (NT, 40, 32)).astype(np.uint8)
    src[0], src[1] = 255, 0
    src[2] = np.clip(np.add.outer(np.arange(40) * 17, np.arange(32) * 23) - 300, 0, 255)
    G["pred_src"] = src
    for ti, (name, size) in enumerate((("p16", 16), ("p8c", 8), ("p4", 4))):
        out = np.zeros((len(tabs[ti]), NT, size, size), np.uint8)
        for mode in range(len(tabs[ti])):
            for t in range(NT):
                b = src[t].copy()
                tabs[ti][mode](C.cast(b.ctypes.data + 8 * 32 + 8, C.c_void_p))
                assert np.array_equal(np.delete(b.reshape(-1), [(8 + y) * 32 + 8 + x for y in range(size) for x in range(size)]),
                                      np.delete(src[t].reshape(-1), [(8 + y) * 32 + 8 + x for y in range(size) for x in range(size)]))
                out[mode, t] = b[8:8 + size, 8:8 + size]
        G["pred_" + name] = out
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.0f} KiB, {len(G)} arrays")


if __name__ == "__main__":
    main()
