#!/usr/bin/env python
"""Generates tests/golden/golden_r01.npz from the UNMODIFIED reference (oracle/_ref/libx264ref.so,
compiled in place from /root/reference by `make -C oracle ref`).  Run in the dev container only:

    python tests/golden/make_golden.py

Every OUTPUT array in the file was produced by the reference's own functions (the six
function-pointer tables of x264_encoder_open, x264_frame_filter / x264_frame_init_lowres /
x264_frame_expand_border*, the static x264_slicetype_frame_cost, x264_me_search_ref +
x264_me_refine_qpel, x264_macroblock_encode, x264_frame_deblock_row) through
oracle/ref_shim/harness.c.  Nothing from oracle/xo_*.c or from the CUDA library touches an output;
the only product code used is the deterministic synthetic-picture generator (inputs; their
SHA-256 is stored, small inputs are stored verbatim).

tests/test_golden.py checks (a) the CPU oracle and (b) the CUDA path against this file, so the GPU
box -- which has no /root/reference -- still pins both to the reference.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, TESTS)
sys.path.insert(0, ROOT)

import cpu_checkers as cc          # noqa: E402
import ref_tables as rt            # noqa: E402
from cpu_checkers import ptr, i16p, i32p, u16p, i8p   # noqa: E402

OUT = os.path.join(HERE, "golden_r01.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def table(enc, name, cls):
    return C.cast(getattr(enc.lib, "xref_" + name)(enc.h), C.POINTER(cls)).contents


def me_blocks(g, rng, size, n, mv_scale):
    """as tests/test_oracle_vs_ref.py:make_me_blocks (analyse.c-style limits), kept here so the
    stored block lists do not depend on test code"""
    blocks = np.zeros(n, cc.ME_BLOCK_DTYPE)
    bw, bh = cc.BLOCK_W[size], cc.BLOCK_H[size]
    fmv = 512 << 2
    for i in range(n):
        mb_x, mb_y = rng.randint(0, g.mb_w), rng.randint(0, g.mb_h)
        b = blocks[i]
        b["i_pixel"] = size
        b["bx"] = mb_x * 16 + rng.randint(0, (16 - bw) // 4 + 1) * 4
        b["by"] = mb_y * 16 + rng.randint(0, (16 - bh) // 4 + 1) * 4
        lim = [((-(mb_x << 4) - 24) << 2, (((g.mb_w - mb_x - 1) << 4) + 24) << 2),
               ((-(mb_y << 4) - 24) << 2, (((g.mb_h - mb_y - 1) << 4) + 24) << 2)]
        for k in range(2):
            smin = int(np.clip(lim[k][0], -fmv, fmv - 1))
            smax = int(np.clip(lim[k][1], -fmv, fmv - 1))
            b["mv_min_spel"][k], b["mv_max_spel"][k] = smin, smax
            b["mv_min_fpel"][k], b["mv_max_fpel"][k] = (smin >> 2) + 6, (smax >> 2) - 6
        b["mvp"] = rng.randint(-mv_scale, mv_scale + 1, 2)
        b["i_mvc"] = rng.randint(0, 9)
        b["mvc"][: b["i_mvc"]] = rng.randint(-mv_scale, mv_scale + 1, (b["i_mvc"], 2))
        if rng.rand() < 0.3:
            b["mvc"][0] = 0
        if rng.rand() < 0.3 and b["i_mvc"] > 1:
            b["mvc"][1] = b["mvp"]
    return blocks


def main():
    import __graft_entry__ as ge
    pkg = ge.load_package()                     # synthetic pictures only (host C++ generator)
    assert cc.ref() is not None, "the reference build is required (make -C oracle ref)"
    G = {}

    # ------------------------------------------------------------ 1. constant tables
    enc = cc.RefEncoder(352, 288, me=1, subme=5, me_range=16, qp=26)
    lib = enc.lib
    cost_mv = np.zeros((52, 8193), np.uint16)
    lam, cqp = np.zeros(52, np.int32), np.zeros(52, np.int32)
    qmf, qbias = np.zeros((4, 52, 16), np.uint16), np.zeros((4, 52, 16), np.uint16)
    for qp in range(52):
        p = lib.xref_cost_mv(enc.h, qp)
        cost_mv[qp] = np.ctypeslib.as_array(C.cast(C.addressof(p.contents) - 2 * 4096, C.POINTER(C.c_uint16)),
                                            shape=(8193,))
        lam[qp], cqp[qp] = lib.xref_lambda(qp), lib.xref_chroma_qp(enc.h, qp)
        for cat in range(4):
            lib.xref_quant_tables(enc.h, cat, qp, ptr(qmf[cat, qp], u16p), ptr(qbias[cat, qp], u16p))
    dq = np.zeros((6, 16), np.int32)
    lib.xref_dequant_table(enc.h, 0, ptr(dq, i32p))
    # cost_mv is 852 KB raw: keep the rows of the distinct lambdas plus a hash of every row
    lam_rows = sorted(set(int(x) for x in lam))
    G["tab_lambda"], G["tab_chroma_qp"], G["tab_quant_mf"], G["tab_quant_bias"], G["tab_dequant"] = lam, cqp, qmf, qbias, dq
    G["tab_cost_mv_sha"] = np.array([sha(cost_mv[qp]) for qp in range(52)])
    G["tab_cost_mv_qp26"] = cost_mv[26]
    G["tab_cost_mv_qp51_head"] = cost_mv[51][4096 - 64: 4096 + 65]
    del lam_rows

    # ------------------------------------------------------------ 2. pixel metrics (tables of x264_pixel_init)
    pix = table(enc, "pixf", rt.PixelTable)
    rng = np.random.RandomState(20261018)
    s1, s2, rows = 16, 96, 80
    a = rng.randint(0, 256, s1 * rows).astype(np.uint8)
    b = rng.randint(0, 256, s2 * rows).astype(np.uint8)
    q1, q2 = s1 * 20, s2 * 20
    a[:q1], b[:q2] = 0, 255                           # maximal differences
    b[q2:2 * q2].reshape(20, s2)[:, :16] = a[q1:2 * q1].reshape(20, s1)   # identical where xb == 0
    a[2 * q1:3 * q1] = (np.arange(q1) & 1) * 255      # alternating extremes
    cases = []
    for size in range(8):
        bw, bh = cc.BLOCK_W[size], cc.BLOCK_H[size]
        for t in range(40):
            ya, yb = rng.randint(0, rows - bh), rng.randint(0, rows - bh)
            xb = 0 if t % 8 == 0 else rng.randint(0, s2 - bw)
            pa, pb = a[ya * s1:], b[yb * s2 + xb:]
            cases.append([size, ya, yb, xb] + [getattr(pix, nm)[size](ptr(pa), s1, ptr(pb), s2)
                                               for nm in ("sad", "ssd", "satd")])
    G["pix_a"], G["pix_b"], G["pix_cases"] = a, b, np.array(cases, np.int64)
    # sad_x4 / satd_x3 / var / var2 / intra x3 8x8c
    x_cases = []
    for size in range(7):
        bw, bh = cc.BLOCK_W[size], cc.BLOCK_H[size]
        for t in range(6):
            ya = rng.randint(0, rows - bh)
            offs = [rng.randint(0, rows - bh) * s2 + rng.randint(0, s2 - bw) for _ in range(4)]
            r4, r3 = (C.c_int * 4)(), (C.c_int * 3)()
            pix.sad_x4[size](ptr(a[ya * s1:]), *[ptr(b[o:]) for o in offs], s2, r4)
            pix.satd_x3[size](ptr(a[ya * s1:]), *[ptr(b[o:]) for o in offs[:3]], s2, r3)
            x_cases.append([size, ya] + offs + list(r4) + list(r3))
    G["pix_x_cases"] = np.array(x_cases, np.int64)
    v_cases = []
    intra_fenc = rng.randint(0, 256, (30, 16 * 8)).astype(np.uint8)
    intra_fdec = rng.randint(0, 256, (30, 32 * 10)).astype(np.uint8)
    intra_res = np.zeros((30, 2, 3), np.int32)
    intra_out = np.zeros((30, 2, 32 * 10), np.uint8)
    for t in range(30):
        o1, o2 = rng.randint(0, s1 * (rows - 16)) & ~15, rng.randint(0, s2 * (rows - 16))
        ssd = C.c_int()
        v2 = pix.var2[3](ptr(a[o1:]), 16, ptr(b[o2:]), s2, C.byref(ssd))
        v_cases.append([o1, o2, pix.var[0](ptr(b[o2:]), s2), pix.var[3](ptr(b[o2:]), s2), v2, ssd.value])
        for k, nm in enumerate(("intra_satd_x3_8x8c", "intra_sad_x3_8x8c")):
            f = intra_fdec[t].copy()
            r = (C.c_int * 3)()
            getattr(pix, nm)(ptr(intra_fenc[t]), ptr(f[32 + 8:]), r)
            intra_res[t, k], intra_out[t, k] = list(r), f
    G["pix_var_cases"] = np.array(v_cases, np.uint64)
    G["intra_fenc"], G["intra_fdec"], G["intra_res"], G["intra_out_sha"] = intra_fenc, intra_fdec, intra_res, np.array(sha(intra_out))

    # ------------------------------------------------------------ 3. transform / quant leaves
    dct = table(enc, "dctf", rt.DctTable)
    zz = table(enc, "zigzagf", rt.ZigzagTable)
    qf = table(enc, "quantf", rt.QuantTable)
    T = 24
    fenc = rng.randint(0, 256, (T, 16 * 16)).astype(np.uint8)
    fdec = rng.randint(0, 256, (T, 32 * 16)).astype(np.uint8)
    fenc[0], fdec[0] = 255, 0
    fenc[1], fdec[1] = 0, 255
    coef = rng.randint(-2000, 2000, (T, 256)).astype(np.int16)
    coef[:4] = rng.randint(-32768, 32768, (4, 256)).astype(np.int16)
    G["dct_fenc"], G["dct_fdec"], G["dct_coef"] = fenc, fdec, coef
    for name, n in (("sub4x4_dct", 16), ("sub8x8_dct", 64), ("sub16x16_dct", 256), ("sub8x8_dct_dc", 4)):
        out = np.zeros((T, n), np.int16)
        for t in range(T):
            getattr(dct, name)(ptr(out[t], i16p), ptr(fenc[t]), ptr(fdec[t]))
        G["dct_" + name] = out
    for name, n in (("add4x4_idct", 16), ("add8x8_idct", 64), ("add16x16_idct", 256), ("add8x8_idct_dc", 4),
                    ("add16x16_idct_dc", 16)):
        out = np.zeros((T, 32 * 16), np.uint8)
        for t in range(T):
            d, c = fdec[t].copy(), coef[t, :n].copy()
            getattr(dct, name)(ptr(d), ptr(c, i16p))
            out[t] = d
        G["dct_" + name] = out
    for name in ("dct4x4dc", "idct4x4dc"):
        out = coef[:, :16].copy()
        for t in range(T):
            getattr(dct, name)(ptr(out[t], i16p))
        G["dct_" + name] = out
    out = np.zeros((T, 16), np.int16)
    for t in range(T):
        zz.scan_4x4(ptr(out[t], i16p), ptr(coef[t], i16p))
    G["dct_zigzag"] = out
    qps = [0, 5, 12, 18, 22, 23, 24, 26, 30, 35, 36, 42, 51]
    qin = np.stack([rng.randint(-s, s + 1, (len(qps), 2, 16)) for s in (4, 40, 400, 4000, 30000)]).astype(np.int16)
    lvl = rng.randint(-40, 41, (len(qps), 16)).astype(np.int16)
    small = rng.randint(-6, 7, (len(qps), 8, 4)).astype(np.int16)
    q_out = np.zeros(qin.shape + (3,), np.int16)            # quant_4x4 / 4x4_dc / 2x2_dc (first 4)
    q_nz = np.zeros(qin.shape[:3] + (3,), np.int32)
    dq_out = np.zeros((len(qps), 2, 16), np.int16)
    oc_out, oc_nz = np.zeros_like(small), np.zeros(small.shape[:2], np.int32)
    for qi, qp in enumerate(qps):
        for inter in (0, 1):
            mf, bias = qmf[inter, qp], qbias[inter, qp]     # CQM_4IY = 0, CQM_4PY = 1
            for s in range(5):
                c = qin[s, qi, inter].copy()
                q_nz[s, qi, inter, 0] = qf.quant_4x4(ptr(c, i16p), ptr(mf, u16p), ptr(bias, u16p))
                q_out[s, qi, inter, :, 0] = c
                c = qin[s, qi, inter].copy()
                q_nz[s, qi, inter, 1] = qf.quant_4x4_dc(ptr(c, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                q_out[s, qi, inter, :, 1] = c
                c = qin[s, qi, inter, :4].copy()
                q_nz[s, qi, inter, 2] = qf.quant_2x2_dc(ptr(c, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                q_out[s, qi, inter, :4, 2] = c
        for k, name in enumerate(("dequant_4x4", "dequant_4x4_dc")):
            c = lvl[qi].copy()
            getattr(qf, name)(ptr(c, i16p), ptr(dq, i32p), qp)
            dq_out[qi, k] = c
        dmf = int(dq[qp % 6][0]) << (qp // 6)
        for t in range(8):
            c = small[qi, t].copy()
            oc_nz[qi, t] = qf.optimize_chroma_2x2_dc(ptr(c, i16p), dmf)
            oc_out[qi, t] = c
    G["q_qps"], G["q_in"], G["q_out"], G["q_nz"] = np.array(qps), qin, q_out, q_nz
    G["q_lvl"], G["q_dequant"], G["q_small"], G["q_optdc"], G["q_optdc_nz"] = lvl, dq_out, small, oc_out, oc_nz
    dec_in = (rng.randint(-2, 3, (200, 16)) * (rng.rand(200, 16) < 0.35)).astype(np.int16)
    G["q_dec_in"] = dec_in
    G["q_dec_out"] = np.array([[qf.decimate_score15(ptr(l, i16p)), qf.decimate_score16(ptr(l, i16p)),
                                qf.coeff_last[2](ptr(l, i16p))] for l in dec_in], np.int32)

    # ------------------------------------------------------------ 4. mc / hpel / lowres leaves
    mc = table(enc, "mcf", rt.McTable)
    stride, mrows = 96, 64
    planes = np.stack([rng.randint(0, 256, stride * mrows).astype(np.uint8) for _ in range(4)])
    org = 20 * stride + 24
    srcs = (rt.u8p * 4)(*[ptr(p[org:]) for p in planes])
    mc_cases, mc_out = [], []
    for t in range(110):
        w, h = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (16, 17), (20, 16), (12, 8), (8, 9)][t % 11]
        mvx, mvy = rng.randint(-40, 41), rng.randint(-40, 41)
        d = np.zeros(32 * 24, np.uint8)
        mc.mc_luma(ptr(d), 32, srcs, stride, mvx, mvy, w, h, None)
        mc_cases.append([w, h, mvx, mvy])
        mc_out.append(d)
    G["mc_planes"], G["mc_cases"], G["mc_luma_out"] = planes, np.array(mc_cases, np.int32), np.stack(mc_out)
    chroma = rng.randint(0, 256, stride * mrows).astype(np.uint8)
    cc_cases, cc_out = [], []
    for t in range(60):
        w, h = [(8, 8), (8, 4), (4, 8), (4, 4)][t % 4]
        mvx, mvy = rng.randint(-60, 61), rng.randint(-60, 61)
        u, v = np.zeros(32 * 8, np.uint8), np.zeros(32 * 8, np.uint8)
        mc.mc_chroma(ptr(u), ptr(v), 32, ptr(chroma[org:]), stride, mvx, mvy, w, h)
        cc_cases.append([w, h, mvx, mvy])
        cc_out.append(np.stack([u, v]))
    G["mc_chroma_plane"], G["mc_chroma_cases"], G["mc_chroma_out"] = chroma, np.array(cc_cases, np.int32), np.stack(cc_out)
    src = planes[0].copy()
    src[: stride * 8] = 255
    src[stride * 8: stride * 16] = 0
    houts = np.zeros((3, stride * mrows), np.uint8)
    buf = np.zeros(stride + 48, np.int16)
    o8 = 8 * stride + 8
    mc.hpel_filter(ptr(houts[0][o8:]), ptr(houts[1][o8:]), ptr(houts[2][o8:]), ptr(src[o8:]), stride, 64, 40, ptr(buf, i16p))
    lo = np.zeros((4, 64 * 32), np.uint8)
    mc.frame_init_lowres_core(ptr(src), *[ptr(x) for x in lo], stride, 64, 40, 24)
    G["hpel_src"], G["hpel_out"], G["lowres_core_out"] = src, houts, lo

    # ------------------------------------------------------------ 5. deblock leaves
    lf = table(enc, "loopf", rt.DeblockTable)
    dstride = 32
    db_in, db_par, db_out = [], [], []
    for t in range(64):
        base, spread = rng.randint(0, 256), [2, 6, 20, 80][t % 4]
        p = np.clip(base + rng.randint(-spread, spread + 1, dstride * 24), 0, 255).astype(np.uint8)
        alpha, beta = rng.randint(0, 60), rng.randint(0, 19)
        tc0 = np.array([rng.randint(-1, 10) for _ in range(4)], np.int8)
        dorg = 4 * dstride + 8
        outs = []
        for d in (0, 1):
            for name, intra in (("deblock_luma", 0), ("deblock_chroma", 0), ("deblock_luma_intra", 1),
                                ("deblock_chroma_intra", 1)):
                q = p.copy()
                if intra:
                    getattr(lf, name)[d](ptr(q[dorg:]), dstride, alpha, beta)
                else:
                    getattr(lf, name)[d](ptr(q[dorg:]), dstride, alpha, beta, ptr(tc0, i8p))
                outs.append(q)
        db_in.append(p)
        db_par.append([alpha, beta] + list(tc0))
        db_out.append(np.stack(outs))
    G["db_in"], G["db_par"], G["db_out"] = np.stack(db_in), np.array(db_par, np.int32), np.stack(db_out)
    nnz = (rng.rand(60, 120) < 0.3).astype(np.uint8)
    refi = rng.randint(-1, 2, (60, 2, 40)).astype(np.int8)
    mvv = rng.randint(-6, 7, (60, 2, 40, 2)).astype(np.int16)
    bs_out = np.zeros((60, 2, 8, 4), np.uint8)
    for t in range(60):
        lf.deblock_strength(ptr(nnz[t]), ptr(refi[t], i8p), ptr(mvv[t], i16p), ptr(bs_out[t]))
    G["bs_nnz"], G["bs_ref"], G["bs_mv"], G["bs_out"] = nnz, refi, mvv, bs_out

    # ------------------------------------------------------------ 6. whole frames: planes, lookahead, ME, deblock
    for tag, w, h, nfr, cut, stored in (("s", 128, 96, 4, 3, True), ("r", 72, 72, 2, -1, True), ("cif", 352, 288, 4, 3, False)):
        e = cc.RefEncoder(w, h, me=1, subme=5, me_range=16, qp=26)
        gE = e.geom
        frames = [pkg.synth_frame(w, h, i, cut_frame=cut) for i in range(nfr)]
        G[f"{tag}_wh"] = np.array([w, h, nfr, cut])
        G[f"{tag}_in_sha"] = np.array([sha(f) for f in frames])
        if stored:
            G[f"{tag}_in"] = np.stack(frames)
        lps, cps = gE[8], None
        # geometry: mb_w, mb_h, luma stride/w/h, lowres stride/w/h, luma plane size, luma origin, chroma origin
        G[f"{tag}_geom"] = np.array(gE[:11], np.int64)
        og = cc.oracle_geom(w, h)          # sizes of the raw allocations only (checked against xref_geometry above)
        assert (og.mb_w, og.mb_h, og.luma_stride, og.luma_plane_size) == (gE[0], gE[1], gE[2], gE[8])
        planes_sha, lowres_sha = [], []
        fdecs, fencs = [], []
        for i in range(nfr):
            fd = e.new_frame(True)
            e.load(fd, frames[i])
            e.lib.xref_frame_filter_all(e.h, fd)
            luma4 = e.buffer(fd, 10, 4 * og.luma_plane_size)
            planes_sha.append([sha(luma4[k * og.luma_plane_size:(k + 1) * og.luma_plane_size]) for k in range(4)]
                              + [sha(e.buffer(fd, 11, og.chroma_plane_size))])
            fdecs.append(fd)
            fe = e.new_frame(False)
            e.load(fe, frames[i])
            e.lib.xref_frame_init_lowres(e.h, fe)
            lo4 = e.buffer(fe, 12, 4 * og.lowres_plane_size)
            lowres_sha.append([sha(lo4[k * og.lowres_plane_size:(k + 1) * og.lowres_plane_size]) for k in range(4)]
                              + [sha(e.buffer(fe, 10, og.luma_plane_size))])
            fencs.append(fe)
            if stored and i == 0:
                G[f"{tag}_f0_luma4"] = luma4.copy()
                G[f"{tag}_f0_lowres4"] = lo4.copy()
        G[f"{tag}_planes_sha"], G[f"{tag}_lowres_sha"] = np.array(planes_sha), np.array(lowres_sha)
        # lookahead: frame i against i-1 (P, p1 == b), intra for frame 0
        arr = (C.c_void_p * nfr)(*[f.value for f in fencs])
        mcnt = og.mb_count
        la_mv, la_cost, la_sums = np.zeros((nfr, mcnt, 2), np.int16), np.zeros((nfr, mcnt), np.int32), np.zeros((nfr, 5), np.int32)
        for i in range(nfr):
            p0 = max(i - 1, 0)
            e.lib.xref_frame_cost(e.h, arr, p0, i, i)
            e.lib.xref_frame_lowres_results(e.h, fencs[i], i - p0, ptr(la_mv[i], i16p), ptr(la_cost[i], i32p),
                                            ptr(la_sums[i], i32p))
        G[f"{tag}_la_mv"], G[f"{tag}_la_cost"], G[f"{tag}_la_sums"] = la_mv, la_cost, la_sums
        # motion search: fenc = frame 1 (source), fref = frame 0 (filtered planes)
        if tag != "r":
            r2 = np.random.RandomState(w + h)
            for label, me, subme, refine in (("hex5q", 1, 5, 1), ("dia2", 0, 2, 0), ("hex3q", 1, 3, 1), ("dia1q", 0, 1, 1)):
                e2 = cc.RefEncoder(w, h, me=me, subme=subme, me_range=16, qp=26)
                fr, fe2 = e2.new_frame(True), e2.new_frame(False)
                e2.load(fr, frames[0])
                e2.load(fe2, frames[1])
                e2.lib.xref_frame_filter_all(e2.h, fr)
                bl, rs = [], []
                for size in range(7):
                    for qp, scale in ((26, 24), (38, 80)):
                        n = 40 if tag == "s" else 60
                        blocks = me_blocks(og, r2, size, n, scale)
                        res = np.zeros(n, cc.ME_RESULT_DTYPE)
                        e2.lib.xref_me_search_batch(e2.h, fe2, fr, qp, me, subme, 16, refine,
                                                    blocks.ctypes.data_as(C.c_void_p), n, res.ctypes.data_as(C.c_void_p))
                        bl.append(blocks.view(np.uint8).reshape(n, -1))
                        rs.append(res.view(np.uint8).reshape(n, -1))
                G[f"{tag}_me_{label}_blocks"], G[f"{tag}_me_{label}_res"] = np.stack(bl), np.stack(rs)
                G[f"{tag}_me_{label}_prm"] = np.array([me, subme, 16, refine])
        # deblock of frame 0 (unfiltered fdec planes) with a random MB field
        for qp, ao, bo in ((30, 0, 0), (20, 3, -2), (44, 0, 0)):
            r3 = np.random.RandomState(qp * 31 + w)
            fd = e.new_frame(True)
            e.load(fd, frames[0])
            mb_type = r3.choice([0, 2, 4, 5, 6], mcnt, p=[0.05, 0.05, 0.5, 0.2, 0.2]).astype(np.int8)
            part = r3.choice([13, 14, 15, 16], mcnt).astype(np.uint8)
            cbp = (r3.randint(0, 48, mcnt) * (r3.rand(mcnt) < 0.6)).astype(np.int16)
            bs = r3.randint(0, 4, (mcnt, 2, 8, 4)).astype(np.uint8)
            bs[r3.rand(mcnt) < 0.2] = 0
            e.lib.xref_deblock_frame(e.h, fd, ptr(mb_type, i8p), ptr(part), ptr(cbp, i16p), ptr(bs), qp, ao, bo)
            k = f"{tag}_db{qp}"
            G[k + "_type"], G[k + "_part"], G[k + "_cbp"], G[k + "_bs"] = mb_type, part, cbp, bs
            G[k + "_par"] = np.array([qp, ao, bo])
            G[k + "_sha"] = np.array([sha(e.buffer(fd, 10, og.luma_plane_size)), sha(e.buffer(fd, 11, og.chroma_plane_size))])
            if stored and qp == 30:
                G[k + "_luma"] = e.buffer(fd, 10, og.luma_plane_size).copy()
                G[k + "_chroma"] = e.buffer(fd, 11, og.chroma_plane_size).copy()

    # ------------------------------------------------------------ 7. residual: x264_macroblock_encode, one P_L0 MB per call
    r4 = np.random.RandomState(4242)
    RQ = [12, 20, 26, 34, 44]
    NMB = 48                                                   # = the MB count of the 128x96 geometry
    res_fy, res_fc = np.zeros((len(RQ), NMB, 16, 16), np.uint8), np.zeros((len(RQ), NMB, 8, 16), np.uint8)
    res_py, res_pc = np.zeros((len(RQ), NMB, 16, 32), np.uint8), np.zeros((len(RQ), NMB, 8, 32), np.uint8)
    res_ry, res_rc = np.zeros_like(res_py), np.zeros_like(res_pc)
    res_lv, res_nz, res_cbp = np.zeros((len(RQ), NMB, 392), np.int16), np.zeros((len(RQ), NMB, 27), np.uint8), np.zeros((len(RQ), NMB), np.int32)
    for qi, qp in enumerate(RQ):
        for t in range(NMB):
            amp = [1, 3, 8, 25, 80][t % 5]
            py = r4.randint(0, 256, (16, 32)).astype(np.uint8)
            pc = r4.randint(0, 256, (8, 32)).astype(np.uint8)
            if t % 3 == 0:
                py[:], pc[:] = r4.randint(30, 220), r4.randint(30, 220)
            py[:, 16:], pc[:, 8:16], pc[:, 24:] = 0, 0, 0
            fy = np.clip(py[:, :16].astype(int) + r4.randint(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
            fc = np.zeros((8, 16), np.uint8)
            fc[:, :8] = np.clip(pc[:, :8].astype(int) + r4.randint(-amp, amp + 1, (8, 8)), 0, 255)
            fc[:, 8:] = np.clip(pc[:, 16:24].astype(int) + r4.randint(-amp, amp + 1, (8, 8)), 0, 255)
            if t % 7 == 0:
                fc[:, :8] = np.clip(fc[:, :8].astype(int) + r4.randint(-3, 4), 0, 255)
            y1, c1 = py.copy(), pc.copy()
            res_cbp[qi, t] = lib.xref_encode_inter_mb(enc.h, ptr(fy), ptr(fc), ptr(y1), ptr(c1), qp,
                                                      ptr(res_lv[qi, t], i16p), ptr(res_nz[qi, t]))
            res_fy[qi, t], res_fc[qi, t], res_py[qi, t], res_pc[qi, t], res_ry[qi, t], res_rc[qi, t] = fy, fc, py, pc, y1, c1
    G["res_qps"] = np.array(RQ)
    G["res_fenc_y"], G["res_fenc_c"], G["res_pred_y"], G["res_pred_c"] = res_fy, res_fc, res_py[..., :16], res_pc[..., :24]
    G["res_recon_y"], G["res_recon_c"], G["res_levels"], G["res_nnz"], G["res_cbp"] = res_ry[..., :16], res_rc[..., :24], res_lv, res_nz, res_cbp

    # ------------------------------------------------------------ 8. config 1: the reference CLI end to end
    import subprocess
    import tempfile
    w, h, nfr = 352, 288, 30
    clip = np.concatenate([pkg.synth_frame(w, h, i, cut_frame=17) for i in range(nfr)])
    with tempfile.TemporaryDirectory() as td:
        src_path, out_path = os.path.join(td, f"syn_{w}x{h}.yuv"), os.path.join(td, "out.264")
        clip.tofile(src_path)
        subprocess.run([cc.REF_CLI, src_path, out_path], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        bits = np.fromfile(out_path, np.uint8)
    G["cli_cif30"] = np.array([sha(clip), sha(bits), str(bits.size)])

    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.0f} KiB, {len(G)} arrays; CLI bitstream {bits.size} bytes")


if __name__ == "__main__":
    main()
