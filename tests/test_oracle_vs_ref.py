"""Pins the CPU oracle (oracle/xo_*.c) to the UNMODIFIED reference C path (oracle/_ref, compiled
from /root/reference by `make -C oracle ref`).  CPU only.  Skipped when the reference build is
neither present nor buildable (then tests/test_golden.py still pins the oracle to vectors that
were generated from this same build)."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
import ref_tables as rt
from cpu_checkers import ptr, i16p, i32p, u16p, i8p, i64p

pytestmark = pytest.mark.skipif(not cc.ref_available(), reason="reference build unavailable")


@pytest.fixture(scope="module")
def enc():
    return cc.RefEncoder(352, 288, me=1, subme=5, me_range=16, qp=26)


def table(enc, name, cls):
    return C.cast(getattr(enc.lib, "xref_" + name)(enc.h), C.POINTER(cls)).contents


def test_table_struct_sizes(enc):
    sizes = (C.c_int * 8)()
    enc.lib.xref_table_sizes(sizes)
    for cls, s in zip(rt.TABLES, sizes):
        assert C.sizeof(cls) == s, cls.__name__
    assert C.sizeof(cc.MeBlock) == 116


# ------------------------------------------------------------------ constant tables

def test_cost_mv_and_quant_tables(enc):
    o = cc.oracle()
    for qp in range(52):
        want = np.ctypeslib.as_array(C.cast(C.addressof(enc.lib.xref_cost_mv(enc.h, qp).contents) - 2 * 4096,
                                            C.POINTER(C.c_uint16)), shape=(8193,))
        got = np.zeros(8193, np.uint16)
        o.xo_cost_mv_table(qp, ptr(got, u16p))
        assert np.array_equal(got, want), f"cost_mv qp {qp}"
        assert o.xo_lambda(qp) == enc.lib.xref_lambda(qp)
        assert o.xo_chroma_qp(qp) == enc.lib.xref_chroma_qp(enc.h, qp)
        for cat in range(4):
            mf_r, b_r = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
            mf_o, b_o = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
            enc.lib.xref_quant_tables(enc.h, cat, qp, ptr(mf_r, u16p), ptr(b_r, u16p))
            o.xo_quant_tables(cat & 1, qp, ptr(mf_o, u16p), ptr(b_o, u16p))
            assert np.array_equal(mf_r, mf_o) and np.array_equal(b_r, b_o), f"quant tables cat {cat} qp {qp}"
    dq_r, dq_o = np.zeros((6, 16), np.int32), np.zeros((6, 16), np.int32)
    enc.lib.xref_dequant_table(enc.h, 0, ptr(dq_r, i32p))
    o.xo_dequant_table(ptr(dq_o, i32p))
    assert np.array_equal(dq_r, dq_o)


# ------------------------------------------------------------------ pixel metrics

def adversarial_planes(rng, stride, rows):
    a = rng.randint(0, 256, stride * rows).astype(np.uint8)
    b = rng.randint(0, 256, stride * rows).astype(np.uint8)
    q = stride * (rows // 4)
    a[:q] = 0
    b[:q] = 255                                    # maximal differences
    b[q:2 * q] = a[q:2 * q]                        # identical
    a[2 * q:3 * q] = (np.arange(q) & 1) * 255      # checkerboard-ish
    return a, b


def test_sad_ssd_satd_all_sizes(enc):
    o = cc.oracle()
    pix = table(enc, "pixf", rt.PixelTable)
    rng = np.random.RandomState(1)
    s1, s2, rows = 16, 96, 80
    a, b = adversarial_planes(rng, s1, rows)
    _, b = adversarial_planes(rng, s2, rows)
    for size in range(8):
        bw, bh = cc.BLOCK_W[size], cc.BLOCK_H[size]
        for trial in range(60):
            ya, yb = rng.randint(0, rows - bh), rng.randint(0, rows - bh)
            xa = 0
            xb = rng.randint(0, s2 - bw)            # unaligned second operand
            pa = a[ya * s1 + xa:]
            pb = b[yb * s2 + xb:]
            for name, cmp in (("sad", 0), ("ssd", 1), ("satd", 2)):
                want = getattr(pix, name)[size](ptr(pa), s1, ptr(pb), s2)
                got = o.xo_cmp(cmp, size, ptr(pa), C.c_ssize_t(s1), ptr(pb), C.c_ssize_t(s2))
                assert got == want, f"{name} size {size} trial {trial}: {got} != {want}"


def test_var_var2_intra_x3(enc):
    o = cc.oracle()
    pix = table(enc, "pixf", rt.PixelTable)
    rng = np.random.RandomState(2)
    for trial in range(50):
        p = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        q = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        assert pix.var[0](ptr(p), 32) == o.xo_var(0, ptr(p), C.c_ssize_t(32))
        assert pix.var[3](ptr(p), 32) == o.xo_var(3, ptr(p), C.c_ssize_t(32))
        sr, so = C.c_int(), C.c_int()
        vr = pix.var2[3](ptr(p), 16, ptr(q), 32, C.byref(sr))
        vo = o.xo_var2_8x8(ptr(p), C.c_ssize_t(16), ptr(q), C.c_ssize_t(32), C.byref(so))
        assert (vr, sr.value) == (vo, so.value)
        # intra x3 8x8c: fenc @16, fdec @32 with top row / left column neighbours present
        fenc = rng.randint(0, 256, 16 * 8).astype(np.uint8)
        fdec = rng.randint(0, 256, 32 * 10).astype(np.uint8)
        for name, satd in (("intra_satd_x3_8x8c", 1), ("intra_sad_x3_8x8c", 0)):
            f1, f2 = fdec.copy(), fdec.copy()
            r1, r2 = (C.c_int * 3)(), (C.c_int * 3)()
            getattr(pix, name)(ptr(fenc), ptr(f1[32 + 8:]), r1)
            o.xo_intra_x3_8x8c(satd, ptr(fenc), ptr(f2[32 + 8:]), r2)
            assert list(r1) == list(r2) and np.array_equal(f1, f2), name


# ------------------------------------------------------------------ transform / quant

def test_dct_idct_zigzag(enc):
    o = cc.oracle()
    dct = table(enc, "dctf", rt.DctTable)
    zz = table(enc, "zigzagf", rt.ZigzagTable)
    rng = np.random.RandomState(3)
    for trial in range(40):
        fenc = rng.randint(0, 256, 16 * 16).astype(np.uint8)
        fdec = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        if trial == 0:
            fenc[:] = 255
            fdec[:] = 0
        if trial == 1:
            fenc[:] = 0
            fdec[:] = 255
        for name, n in (("sub4x4_dct", 16), ("sub8x8_dct", 64), ("sub16x16_dct", 256), ("sub8x8_dct_dc", 4)):
            r, g = np.zeros(n, np.int16), np.zeros(n, np.int16)
            getattr(dct, name)(ptr(r, i16p), ptr(fenc), ptr(fdec))
            getattr(o, "xo_" + name)(ptr(g, i16p), ptr(fenc), ptr(fdec))
            assert np.array_equal(r, g), name
        coef = rng.randint(-2000, 2000, 256).astype(np.int16)
        if trial < 4:
            coef = rng.randint(-32768, 32768, 256).astype(np.int16)      # int16 wrap-around paths
        for name, n in (("add4x4_idct", 16), ("add8x8_idct", 64), ("add16x16_idct", 256),
                        ("add8x8_idct_dc", 4), ("add16x16_idct_dc", 16)):
            d1, d2 = fdec.copy(), fdec.copy()
            c1, c2 = coef[:n].copy(), coef[:n].copy()
            getattr(dct, name)(ptr(d1), ptr(c1, i16p))
            getattr(o, "xo_" + name)(ptr(d2), ptr(c2, i16p))
            assert np.array_equal(d1, d2), name
        for name in ("dct4x4dc", "idct4x4dc"):
            c1, c2 = coef[:16].copy(), coef[:16].copy()
            getattr(dct, name)(ptr(c1, i16p))
            getattr(o, "xo_" + name)(ptr(c2, i16p))
            assert np.array_equal(c1, c2), name
        l1, l2 = np.zeros(16, np.int16), np.zeros(16, np.int16)
        zz.scan_4x4(ptr(l1, i16p), ptr(coef, i16p))
        o.xo_zigzag_4x4(ptr(l2, i16p), ptr(coef, i16p))
        assert np.array_equal(l1, l2)


def test_quant_dequant_decimate(enc):
    o = cc.oracle()
    qf = table(enc, "quantf", rt.QuantTable)
    rng = np.random.RandomState(4)
    dq = np.zeros((6, 16), np.int32)
    o.xo_dequant_table(ptr(dq, i32p))
    for qp in list(range(0, 52, 3)) + [22, 23, 24, 35, 36, 51]:
        for inter in (0, 1):
            mf, bias = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
            o.xo_quant_tables(inter, qp, ptr(mf, u16p), ptr(bias, u16p))
            for trial in range(12):
                scale = [4, 40, 400, 4000, 30000][trial % 5]
                coef = rng.randint(-scale, scale + 1, 16).astype(np.int16)
                c1, c2 = coef.copy(), coef.copy()
                n1 = qf.quant_4x4(ptr(c1, i16p), ptr(mf, u16p), ptr(bias, u16p))
                n2 = o.xo_quant_4x4(ptr(c2, i16p), ptr(mf, u16p), ptr(bias, u16p))
                assert n1 == n2 and np.array_equal(c1, c2), f"quant_4x4 qp {qp}"
                c1, c2 = coef.copy(), coef.copy()
                n1 = qf.quant_4x4_dc(ptr(c1, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                n2 = o.xo_quant_4x4_dc(ptr(c2, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                assert n1 == n2 and np.array_equal(c1, c2), "quant_4x4_dc"
                c1, c2 = coef[:4].copy(), coef[:4].copy()
                n1 = qf.quant_2x2_dc(ptr(c1, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                n2 = o.xo_quant_2x2_dc(ptr(c2, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                assert n1 == n2 and np.array_equal(c1, c2), "quant_2x2_dc"
                lv = rng.randint(-40, 41, 16).astype(np.int16)
                for name in ("dequant_4x4", "dequant_4x4_dc"):
                    c1, c2 = lv.copy(), lv.copy()
                    getattr(qf, name)(ptr(c1, i16p), ptr(dq, i32p), qp)
                    getattr(o, "xo_" + name)(ptr(c2, i16p), ptr(dq, i32p), qp)
                    assert np.array_equal(c1, c2), f"{name} qp {qp}"
                dmf = int(dq[qp % 6][0]) << (qp // 6)
                small = rng.randint(-6, 7, 4).astype(np.int16)
                c1, c2 = small.copy(), small.copy()
                n1 = qf.optimize_chroma_2x2_dc(ptr(c1, i16p), dmf)
                n2 = o.xo_optimize_chroma_2x2_dc(ptr(c2, i16p), dmf)
                assert n1 == n2 and np.array_equal(c1, c2), f"optimize_chroma_2x2_dc qp {qp}"
    for trial in range(300):
        lv = (rng.randint(-2, 3, 16) * (rng.rand(16) < 0.35)).astype(np.int16)
        assert qf.decimate_score15(ptr(lv, i16p)) == o.xo_decimate_score15(ptr(lv, i16p))
        assert qf.decimate_score16(ptr(lv, i16p)) == o.xo_decimate_score16(ptr(lv, i16p))
        assert qf.coeff_last[2](ptr(lv, i16p)) == o.xo_coeff_last(ptr(lv, i16p), 16)


# ------------------------------------------------------------------ motion compensation

def test_mc_luma_get_ref_chroma_hpel_lowres(enc):
    o = cc.oracle()
    mc = table(enc, "mcf", rt.McTable)
    rng = np.random.RandomState(5)
    stride, rows = 96, 64
    planes = [rng.randint(0, 256, stride * rows).astype(np.uint8) for _ in range(4)]
    org = 20 * stride + 24
    srcs = (rt.u8p * 4)(*[ptr(p[org:]) for p in planes])
    for trial in range(200):
        w, h = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (16, 17), (20, 16), (12, 8), (8, 9)][trial % 11]
        mvx, mvy = rng.randint(-40, 41), rng.randint(-40, 41)
        d1, d2 = np.zeros(32 * 24, np.uint8), np.zeros(32 * 24, np.uint8)
        mc.mc_luma(ptr(d1), 32, srcs, stride, mvx, mvy, w, h, None)
        o.xo_mc_luma(ptr(d2), C.c_ssize_t(32), srcs, C.c_ssize_t(stride), mvx, mvy, w, h)
        assert np.array_equal(d1, d2), "mc_luma"
        d1[:] = 0
        d2[:] = 0
        s1, s2 = C.c_ssize_t(32), C.c_ssize_t(32)
        o.xo_get_ref.restype = C.c_void_p
        r1 = mc.get_ref(ptr(d1), C.byref(s1), srcs, stride, mvx, mvy, w, h, None)
        r2 = o.xo_get_ref(ptr(d2), C.byref(s2), srcs, C.c_ssize_t(stride), mvx, mvy, w, h)
        assert s1.value == s2.value and np.array_equal(d1, d2), "get_ref"
        assert (r1 - d1.ctypes.data) == (r2 - d2.ctypes.data) or (r1 == r2), "get_ref pointer"
    chroma = rng.randint(0, 256, stride * rows).astype(np.uint8)
    for trial in range(100):
        w, h = [(8, 8), (8, 4), (4, 8), (4, 4)][trial % 4]
        mvx, mvy = rng.randint(-60, 61), rng.randint(-60, 61)
        u1, v1, u2, v2 = (np.zeros(32 * 8, np.uint8) for _ in range(4))
        mc.mc_chroma(ptr(u1), ptr(v1), 32, ptr(chroma[org:]), stride, mvx, mvy, w, h)
        o.xo_mc_chroma(ptr(u2), ptr(v2), C.c_ssize_t(32), ptr(chroma[org:]), C.c_ssize_t(stride), mvx, mvy, w, h)
        assert np.array_equal(u1, u2) and np.array_equal(v1, v2), "mc_chroma"
    src = planes[0].copy()
    src[: stride * 8] = 255
    src[stride * 8: stride * 16] = 0                # hard edges: exercises the clipping
    outs = [[np.zeros(stride * rows, np.uint8) for _ in range(3)] for _ in range(2)]
    buf = np.zeros(stride + 48, np.int16)
    o8 = 8 * stride + 8
    mc.hpel_filter(ptr(outs[0][0][o8:]), ptr(outs[0][1][o8:]), ptr(outs[0][2][o8:]), ptr(src[o8:]), stride, 64, 40,
                   ptr(buf, i16p))
    o.xo_hpel_filter(ptr(outs[1][0][o8:]), ptr(outs[1][1][o8:]), ptr(outs[1][2][o8:]), ptr(src[o8:]),
                     C.c_ssize_t(stride), 64, 40)
    for k in range(3):
        assert np.array_equal(outs[0][k], outs[1][k]), f"hpel_filter plane {k}"
    lo = [[np.zeros(64 * 32, np.uint8) for _ in range(4)] for _ in range(2)]
    mc.frame_init_lowres_core(ptr(src), *[ptr(x) for x in lo[0]], stride, 64, 40, 24)
    o.xo_lowres_core(ptr(src), *[ptr(x) for x in lo[1]], C.c_ssize_t(stride), C.c_ssize_t(64), 40, 24)
    for k in range(4):
        assert np.array_equal(lo[0][k], lo[1][k]), f"lowres core plane {k}"


# ------------------------------------------------------------------ deblock

def test_deblock_edge_filters(enc):
    o = cc.oracle()
    lf = table(enc, "loopf", rt.DeblockTable)
    rng = np.random.RandomState(6)
    stride = 64
    for trial in range(400):
        base = rng.randint(0, 256)
        spread = [2, 6, 20, 80][trial % 4]
        pix = np.clip(base + rng.randint(-spread, spread + 1, stride * 40), 0, 255).astype(np.uint8)
        alpha, beta = rng.randint(0, 60), rng.randint(0, 19)
        tc0 = np.array([rng.randint(-1, 10) for _ in range(4)], np.int8)
        org = 12 * stride + 16
        for d in (0, 1):
            for name, oname, intra in (("deblock_luma", "xo_deblock_luma", 0), ("deblock_chroma", "xo_deblock_chroma", 0),
                                       ("deblock_luma_intra", "xo_deblock_luma_intra", 1),
                                       ("deblock_chroma_intra", "xo_deblock_chroma_intra", 1)):
                p1, p2 = pix.copy(), pix.copy()
                if intra:
                    getattr(lf, name)[d](ptr(p1[org:]), stride, alpha, beta)
                    getattr(o, oname)(ptr(p2[org:]), C.c_ssize_t(stride), d, alpha, beta)
                else:
                    getattr(lf, name)[d](ptr(p1[org:]), stride, alpha, beta, ptr(tc0, i8p))
                    getattr(o, oname)(ptr(p2[org:]), C.c_ssize_t(stride), d, alpha, beta, ptr(tc0, i8p))
                assert np.array_equal(p1, p2), f"{name}[{d}] trial {trial}"


def test_deblock_strength(enc):
    o = cc.oracle()
    lf = table(enc, "loopf", rt.DeblockTable)
    rng = np.random.RandomState(8)
    for trial in range(100):
        nnz = (rng.rand(120) < 0.3).astype(np.uint8)
        ref = rng.randint(-1, 2, (2, 40)).astype(np.int8)
        mv = rng.randint(-6, 7, (2, 40, 2)).astype(np.int16)
        b1 = np.zeros((2, 8, 4), np.uint8)
        b2 = np.zeros((2, 8, 4), np.uint8)
        lf.deblock_strength(ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(b1))
        o.xo_deblock_strength(1, ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(b2))
        assert np.array_equal(b1, b2)


def test_macroblock_deblock_strength(enc):
    """x264_macroblock_deblock_strength (common/macroblock.c:677-691): every mb_type, bs pre-filled with a pattern so
    that the bytes the reference leaves alone (edge 0 of an intra macroblock, rows 4..7) are checked too"""
    o = cc.oracle()
    rng = np.random.RandomState(81)
    for trial in range(300):
        mb_type = int(rng.choice([0, 1, 2, 3, 4, 5, 6]))
        nnz = (rng.rand(120) < 0.3).astype(np.uint8)
        ref = rng.randint(-1, 2, (2, 40)).astype(np.int8)
        mv = rng.randint(-6, 7, (2, 40, 2)).astype(np.int16)
        b1 = rng.randint(0, 256, (2, 8, 4)).astype(np.uint8)
        b2 = b1.copy()
        enc.lib.xref_macroblock_deblock_strength(enc.h, mb_type, ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(b1))
        t = np.array([mb_type], np.int8)
        o.xo_macroblock_deblock_strength(1, ptr(t, i8p), ptr(nnz), ptr(ref, i8p), ptr(mv, i16p), ptr(b2))
        assert np.array_equal(b1, b2), f"mb_type {mb_type}"


@pytest.mark.parametrize("qp,aoff,boff", [(26, 0, 0), (38, 0, 0), (20, 3, -2), (51, 0, 0), (12, 0, 0)])
def test_deblock_frame(enc, qp, aoff, boff):
    o = cc.oracle()
    w, h = 352, 288
    g = cc.oracle_geom(w, h)
    rng = np.random.RandomState(qp)
    frame = cc.synth_clip(w, h, 1, seed=qp)[0]
    # blocky content so that edges actually filter
    f = enc.new_frame(True)
    enc.load(f, frame)
    slot = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frame), ptr(slot))
    n = g.mb_count
    mb_type = rng.choice([0, 2, 4, 5, 6], n, p=[0.05, 0.05, 0.5, 0.2, 0.2]).astype(np.int8)
    partition = rng.choice([13, 14, 15, 16], n).astype(np.uint8)
    cbp = (rng.randint(0, 48, n) * (rng.rand(n) < 0.6)).astype(np.int16)
    bs = rng.randint(0, 4, (n, 2, 8, 4)).astype(np.uint8)
    bs[rng.rand(n) < 0.2] = 0
    enc.lib.xref_deblock_frame(enc.h, f, ptr(mb_type, i8p), ptr(partition), ptr(cbp, i16p), ptr(bs), qp, aoff, boff)
    o.xo_deblock_frame(C.byref(g), ptr(slot), ptr(mb_type, i8p), ptr(partition), ptr(cbp, i16p), ptr(bs), qp, aoff, boff)
    assert np.array_equal(enc.buffer(f, 10, g.luma_plane_size), slot[: g.luma_plane_size]), "luma"
    assert np.array_equal(enc.buffer(f, 11, g.chroma_plane_size),
                          slot[g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size]), "chroma"
    changed = np.count_nonzero(slot[: g.luma_plane_size] !=
                               np.frombuffer(_loaded(g, frame), np.uint8)[: g.luma_plane_size])
    if qp >= 20:
        assert changed > 0, "the test must exercise the filter"


def _loaded(g, frame):
    s = np.zeros(g.slot_bytes, np.uint8)
    cc.oracle().xo_frame_load_i420(C.byref(g), ptr(frame), ptr(s))
    return s.tobytes()


# ------------------------------------------------------------------ frames, lookahead

@pytest.mark.parametrize("w,h", [(352, 288), (200, 120), (368, 304), (960, 540), (64, 48)])
def test_frame_planes(w, h):
    o = cc.oracle()
    enc = cc.RefEncoder(w, h)
    g = cc.oracle_geom(w, h)
    G = enc.geom
    assert (G[0], G[1], G[2], G[3], G[4]) == (g.mb_w, g.mb_h, g.luma_stride, g.luma_w, g.luma_h)
    assert (G[5], G[6], G[7], G[8], G[9], G[10]) == (g.lowres_stride, g.lowres_w, g.lowres_h, g.luma_plane_size,
                                                   g.luma_origin, g.chroma_origin)
    frame = cc.synth_clip(w, h, 1)[0]
    f = enc.new_frame(True)
    enc.load(f, frame)
    slot = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frame), ptr(slot))
    enc.lib.xref_frame_filter_all(enc.h, f)
    o.xo_frame_expand_border(C.byref(g), ptr(slot))
    o.xo_frame_filter(C.byref(g), ptr(slot))
    assert np.array_equal(enc.buffer(f, 10, 4 * g.luma_plane_size), slot[: 4 * g.luma_plane_size]), "luma N/H/V/HV"
    assert np.array_equal(enc.buffer(f, 11, g.chroma_plane_size),
                          slot[g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size]), "chroma"
    f2 = enc.new_frame(False)
    enc.load(f2, frame)
    enc.lib.xref_frame_init_lowres(enc.h, f2)
    slot2 = np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(frame), ptr(slot2))
    o.xo_frame_init_lowres(C.byref(g), ptr(slot2))
    assert np.array_equal(enc.buffer(f2, 12, 4 * g.lowres_plane_size),
                          slot2[g.slot_lowres_off: g.slot_lowres_off + 4 * g.lowres_plane_size]), "lowres"
    assert np.array_equal(enc.buffer(f2, 10, g.luma_plane_size), slot2[: g.luma_plane_size]), "source side effect"


@pytest.mark.parametrize("w,h,n,cut", [(352, 288, 5, 3), (200, 120, 3, -1), (64, 48, 3, -1)])
def test_lookahead_frame_cost(w, h, n, cut):
    o = cc.oracle()
    enc = cc.RefEncoder(w, h)
    g = cc.oracle_geom(w, h)
    clip = cc.synth_clip(w, h, n, cut_frame=cut)
    frames = [enc.new_frame(False) for _ in range(n)]
    slots = [np.zeros(g.slot_bytes, np.uint8) for _ in range(n)]
    for i in range(n):
        enc.load(frames[i], clip[i])
        enc.lib.xref_frame_init_lowres(enc.h, frames[i])
        o.xo_frame_load_i420(C.byref(g), ptr(clip[i]), ptr(slots[i]))
        o.xo_frame_init_lowres(C.byref(g), ptr(slots[i]))
    arr = (C.c_void_p * n)(*[f.value for f in frames])
    mc = g.mb_count
    for i in range(n):
        p0 = max(i - 1, 0)
        enc.lib.xref_frame_cost(enc.h, arr, p0, i, i)
        mv_r, c_r, s_r = np.zeros((mc, 2), np.int16), np.zeros(mc, np.int32), np.zeros(5, np.int32)
        enc.lib.xref_frame_lowres_results(enc.h, frames[i], i - p0, ptr(mv_r, i16p), ptr(c_r, i32p), ptr(s_r, i32p))
        mv_o, c_o, s_o = np.zeros((mc, 2), np.int16), np.zeros(mc, np.int32), np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(g), ptr(slots[i]), ptr(slots[p0]) if i else None, 1,
                                  ptr(mv_o, i16p), ptr(c_o, i32p), ptr(s_o, i32p), None)
        if i:
            assert np.array_equal(mv_r, mv_o) and np.array_equal(c_r, c_o), f"frame {i}"
            assert (s_r[0], s_r[2], s_r[3]) == (s_o[0], s_o[1], s_o[2]), f"frame {i} sums"
        else:
            assert s_r[2] == s_o[1]


# ------------------------------------------------------------------ motion search

def make_me_blocks(g, rng, size, n, mv_scale):
    """blocks tiling random positions with analyse.c-style MV limits (fpel border 6, mv_range 512)"""
    blocks = np.zeros(n, cc.ME_BLOCK_DTYPE)
    bw, bh = cc.BLOCK_W[size], cc.BLOCK_H[size]
    for i in range(n):
        mb_x, mb_y = rng.randint(0, g.mb_w), rng.randint(0, g.mb_h)
        bx = mb_x * 16 + rng.randint(0, (16 - bw) // 4 + 1) * 4
        by = mb_y * 16 + rng.randint(0, (16 - bh) // 4 + 1) * 4
        fmv = 512 << 2
        lim = [((-(mb_x << 4) - 24) << 2, (((g.mb_w - mb_x - 1) << 4) + 24) << 2),
               ((-(mb_y << 4) - 24) << 2, (((g.mb_h - mb_y - 1) << 4) + 24) << 2)]
        b = blocks[i]
        b["i_pixel"], b["bx"], b["by"] = size, bx, by
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


@pytest.mark.parametrize("me,subme,refine", [(0, 1, 1), (0, 2, 0), (1, 2, 1), (1, 3, 0), (1, 4, 1), (1, 5, 1), (0, 5, 0)])
def test_me_search(me, subme, refine):
    o = cc.oracle()
    w, h = 352, 288
    enc = cc.RefEncoder(w, h, me=me, subme=max(subme, 1), me_range=16, qp=26)
    g = cc.oracle_geom(w, h)
    clip = cc.synth_clip(w, h, 2)
    fref = enc.new_frame(True)
    fenc = enc.new_frame(False)
    enc.load(fref, clip[0])
    enc.load(fenc, clip[1])
    enc.lib.xref_frame_filter_all(enc.h, fref)
    slot_ref, slot_enc = np.zeros(g.slot_bytes, np.uint8), np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(clip[0]), ptr(slot_ref))
    o.xo_frame_expand_border(C.byref(g), ptr(slot_ref))
    o.xo_frame_filter(C.byref(g), ptr(slot_ref))
    o.xo_frame_load_i420(C.byref(g), ptr(clip[1]), ptr(slot_enc))
    rng = np.random.RandomState(100 + me * 10 + subme)
    for size in range(7):
        for qp, mv_scale in ((26, 24), (38, 80)):
            n = 150
            blocks = make_me_blocks(g, rng, size, n, mv_scale)
            r_ref = np.zeros(n, cc.ME_RESULT_DTYPE)
            r_ora = np.zeros(n, cc.ME_RESULT_DTYPE)
            enc.lib.xref_me_search_batch(enc.h, fenc, fref, qp, me, subme, 16, refine,
                                         blocks.ctypes.data_as(C.c_void_p), n, r_ref.ctypes.data_as(C.c_void_p))
            prm = cc.MeParams(me, subme, 16, qp, refine)
            o.xo_me_search_batch(C.byref(g), ptr(slot_enc), ptr(slot_ref), C.byref(prm), n,
                                 blocks.ctypes.data_as(C.c_void_p), r_ora.ctypes.data_as(C.c_void_p))
            bad = [i for i in range(n) if r_ref[i] != r_ora[i]]
            assert not bad, (f"me {me} subme {subme} size {size} qp {qp}: {len(bad)} differ, "
                             f"e.g. {bad[0]}: ref {r_ref[bad[0]]} oracle {r_ora[bad[0]]} block {blocks[bad[0]]}")


def _me_pair(w, h, me, subme):
    """a reference encoder + the oracle's slots of the same two frames"""
    o = cc.oracle()
    enc = cc.RefEncoder(w, h, me=me, subme=max(subme, 1), me_range=16, qp=26)
    g = cc.oracle_geom(w, h)
    clip = cc.synth_clip(w, h, 2)
    fref = enc.new_frame(True)
    fenc = enc.new_frame(False)
    enc.load(fref, clip[0])
    enc.load(fenc, clip[1])
    enc.lib.xref_frame_filter_all(enc.h, fref)
    slot_ref, slot_enc = np.zeros(g.slot_bytes, np.uint8), np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(clip[0]), ptr(slot_ref))
    o.xo_frame_expand_border(C.byref(g), ptr(slot_ref))
    o.xo_frame_filter(C.byref(g), ptr(slot_ref))
    o.xo_frame_load_i420(C.byref(g), ptr(clip[1]), ptr(slot_enc))
    return enc, g, fenc, fref, slot_enc, slot_ref


@pytest.mark.parametrize("me,subme,refine", [(2, 2, 0), (2, 5, 1), (3, 1, 1), (3, 4, 0), (4, 2, 1), (4, 5, 0), (4, 1, 1)])
def test_me_search_umh_esa_tesa(me, subme, refine):
    """me = UMH / ESA / TESA: the reference accepts the parameter (encoder.c:251-259) and its search has no case for
    them (me.c:389-394) -- predictors + sub-pel refinement; TESA with subme >= 2 makes fpelcmp SATD (encoder.c:429-432;
    with subme <= 1 the parameter check turns it into ESA).  The oracle must do the same, whatever that is."""
    o = cc.oracle()
    enc, g, fenc, fref, slot_enc, slot_ref = _me_pair(352, 288, me, subme)
    assert enc.geom[15] == (3 if me == 4 and subme <= 1 else me)
    rng = np.random.RandomState(900 + me * 10 + subme)
    for size in range(7):
        n = 120
        blocks = make_me_blocks(g, rng, size, n, 40)
        r_ref = np.zeros(n, cc.ME_RESULT_DTYPE)
        r_ora = np.zeros(n, cc.ME_RESULT_DTYPE)
        enc.lib.xref_me_search_batch(enc.h, fenc, fref, 30, enc.geom[15], subme, 16, refine,
                                     blocks.ctypes.data_as(C.c_void_p), n, r_ref.ctypes.data_as(C.c_void_p))
        prm = cc.MeParams(me, subme, 16, 30, refine)
        o.xo_me_search_batch(C.byref(g), ptr(slot_enc), ptr(slot_ref), C.byref(prm), n,
                             blocks.ctypes.data_as(C.c_void_p), r_ora.ctypes.data_as(C.c_void_p))
        assert np.array_equal(r_ref, r_ora), f"me {me} subme {subme} size {size}: {np.count_nonzero(r_ref != r_ora)} differ"


@pytest.mark.parametrize("me,subme", [(1, 2), (1, 4), (1, 5), (0, 3), (2, 5)])
def test_me_halfpel_thresh_and_refdupe(me, subme):
    """x264_me_search_ref with p_halfpel_thresh, x264_me_refine_qpel_refdupe and x264_me_refine_qpel alone
    (me.c:421, 426-440, 526-539) against the oracle's xo_me_search_batch_ex"""
    o = cc.oracle()
    enc, g, fenc, fref, slot_enc, slot_ref = _me_pair(352, 288, me, subme)
    rng = np.random.RandomState(77 + me * 10 + subme)
    for size in (0, 1, 3, 4, 6):
        n = 150
        blocks = make_me_blocks(g, rng, size, n, 40)
        prm = cc.MeParams(me, subme, 16, 28, 0)
        # plain search first: gives realistic costs to derive thresholds and starting points from
        base = np.zeros(n, cc.ME_RESULT_DTYPE)
        o.xo_me_search_batch(C.byref(g), ptr(slot_enc), ptr(slot_ref), C.byref(prm), n,
                             blocks.ctypes.data_as(C.c_void_p), base.ctypes.data_as(C.c_void_p))
        thresh0 = (base["cost"] * rng.choice([0.5, 0.8, 0.95, 1.0, 1.3, 4.0], n)).astype(np.int32)
        thresh0[rng.rand(n) < 0.1] = 2**31 - 1
        for mode in (0, 1, 2):
            start = base.copy()
            if mode:
                start["mv"] = (start["mv"] & ~3) if mode == 1 else start["mv"]
                start["mv"] += rng.randint(-1, 2, (n, 2)) * 4
                start["cost"] += rng.randint(0, 50, n)
            r_ref, r_ora = start.copy(), start.copy()
            t_ref, t_ora = thresh0.copy(), thresh0.copy()
            use_t = mode != 2
            enc.lib.xref_me_search_batch_ex(enc.h, fenc, fref, 28, me, subme, 16, 0, blocks.ctypes.data_as(C.c_void_p), n,
                                            r_ref.ctypes.data_as(C.c_void_p), mode, ptr(t_ref, cc.i32p) if use_t else None)
            o.xo_me_search_batch_ex(C.byref(g), ptr(slot_enc), ptr(slot_ref), C.byref(prm), n,
                                    blocks.ctypes.data_as(C.c_void_p), r_ora.ctypes.data_as(C.c_void_p), mode,
                                    ptr(t_ora, cc.i32p) if use_t else None)
            assert np.array_equal(r_ref, r_ora), f"mode {mode} size {size}: {np.count_nonzero(r_ref != r_ora)} results differ"
            assert np.array_equal(t_ref, t_ora), f"mode {mode} size {size}: thresholds differ"
            if mode == 0:
                cut = np.count_nonzero((r_ref["mv"] != base["mv"]).any(1) | (r_ref["cost"] != base["cost"]))
                assert cut > 0 or subme < 4, "the early exit never fired"
                assert np.count_nonzero(t_ref != thresh0) > 0


# ------------------------------------------------------------------ residual

@pytest.mark.parametrize("qp", [20, 26, 34])
def test_recon_frame_through_reference_functions(qp):
    """xref_recon_frame (x264_mb_mc + x264_macroblock_encode over a whole frame, then x264_frame_deblock_row) -- the CPU
    baseline of config 4 in bench.py -- against the oracle's xo_mc_frame + xo_residual_frame + xo_deblock_frame"""
    o = cc.oracle()
    w, h = 352, 288
    enc = cc.RefEncoder(w, h, me=1, subme=5, qp=qp)
    g = cc.oracle_geom(w, h)
    clip = cc.synth_clip(w, h, 2)
    fref, fdec, fenc = enc.new_frame(True), enc.new_frame(True), enc.new_frame(False)
    enc.load(fref, clip[0])
    enc.load(fenc, clip[1])
    enc.load(fdec, clip[1])                                   # any content: every macroblock is overwritten
    enc.lib.xref_frame_filter_all(enc.h, fref)
    slot_ref, slot_enc = np.zeros(g.slot_bytes, np.uint8), np.zeros(g.slot_bytes, np.uint8)
    o.xo_frame_load_i420(C.byref(g), ptr(clip[0]), ptr(slot_ref))
    o.xo_frame_expand_border(C.byref(g), ptr(slot_ref))
    o.xo_frame_filter(C.byref(g), ptr(slot_ref))
    o.xo_frame_load_i420(C.byref(g), ptr(clip[1]), ptr(slot_enc))
    rng = np.random.RandomState(qp)
    n = g.mb_count
    mv = (np.array([12, 8]) + rng.randint(-6, 7, (n, 2))).astype(np.int16)
    mv[rng.rand(n) < 0.05] = [-300, 250]                      # clipped by mv_min / mv_max near the borders
    lv_r, nz_r, cbp_r = np.zeros((n, 392), np.int16), np.zeros((n, 27), np.uint8), np.zeros(n, np.int16)
    enc.lib.xref_recon_frame(enc.h, fenc, fref, fdec, ptr(mv, i16p), qp, ptr(lv_r, i16p), ptr(nz_r), ptr(cbp_r, i16p))
    pred = np.zeros(g.slot_bytes, np.uint8)
    o.xo_mc_frame(C.byref(g), ptr(slot_ref), ptr(mv, i16p), ptr(pred))
    lv_o, nz_o, cbp_o = np.zeros((n, 392), np.int16), np.zeros((n, 27), np.uint8), np.zeros(n, np.int16)
    o.xo_residual_frame(C.byref(g), ptr(slot_enc), ptr(pred), qp, ptr(lv_o, i16p), ptr(nz_o), ptr(cbp_o, i16p))
    assert np.array_equal(lv_r, lv_o) and np.array_equal(nz_r, nz_o) and np.array_equal(cbp_r, cbp_o)

    def planes(buf_y, buf_c):
        y = buf_y[g.luma_origin:][: g.luma_h * g.luma_stride].reshape(g.luma_h, g.luma_stride)[:, : g.luma_w]
        c = buf_c[g.chroma_origin:][: (g.luma_h // 2) * g.chroma_stride].reshape(g.luma_h // 2, g.chroma_stride)[:, : g.luma_w]
        return y, c
    ry, rc = planes(enc.buffer(fdec, 10, 4 * g.luma_plane_size), enc.buffer(fdec, 11, g.chroma_plane_size))
    oy, oc = planes(pred, pred[g.slot_chroma_off:])
    assert np.array_equal(ry, oy) and np.array_equal(rc, oc), "reconstruction before the in-loop filter"
    mb_type, part = np.full(n, 4, np.int8), np.full(n, 16, np.uint8)
    bs = (rng.rand(n, 2, 8, 4) < 0.35).astype(np.uint8) * rng.randint(1, 3, (n, 2, 8, 4)).astype(np.uint8)
    enc.lib.xref_deblock_frame(enc.h, fdec, ptr(mb_type, i8p), ptr(part), ptr(cbp_r, i16p), ptr(bs), qp, 0, 0)
    o.xo_deblock_frame(C.byref(g), ptr(pred), ptr(mb_type, i8p), ptr(part), ptr(cbp_o, i16p), ptr(bs), qp, 0, 0)
    ry, rc = planes(enc.buffer(fdec, 10, 4 * g.luma_plane_size), enc.buffer(fdec, 11, g.chroma_plane_size))
    oy, oc = planes(pred, pred[g.slot_chroma_off:])
    assert np.array_equal(ry, oy) and np.array_equal(rc, oc), "deblocked reconstruction"


@pytest.mark.parametrize("qp", [12, 18, 22, 26, 32, 40, 51])
def test_residual_inter_mb(enc, qp):
    o = cc.oracle()
    rng = np.random.RandomState(qp)
    o.xo_encode_inter_mb.restype = C.c_int
    for trial in range(300):
        amp = [1, 3, 8, 25, 80][trial % 5]
        pred_y = rng.randint(0, 256, (16, 32)).astype(np.uint8)
        pred_c = rng.randint(0, 256, (8, 32)).astype(np.uint8)
        if trial % 3 == 0:                                   # smooth prediction, small residual
            pred_y[:] = rng.randint(30, 220)
            pred_c[:] = rng.randint(30, 220)
        fenc_y = np.clip(pred_y[:, :16].astype(int) + rng.randint(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
        fenc_c = np.zeros((8, 16), np.uint8)
        fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-amp, amp + 1, (8, 8)), 0, 255)
        fenc_c[:, 8:] = np.clip(pred_c[:, 16:24].astype(int) + rng.randint(-amp, amp + 1, (8, 8)), 0, 255)
        if trial % 7 == 0:
            fenc_c[:, :8] = np.clip(fenc_c[:, :8].astype(int) + rng.randint(-3, 4), 0, 255)   # DC-only chroma change
        y1, c1, y2, c2 = pred_y.copy(), pred_c.copy(), pred_y.copy(), pred_c.copy()
        l1, l2 = np.zeros(392, np.int16), np.zeros(392, np.int16)
        n1, n2 = np.zeros(27, np.uint8), np.zeros(27, np.uint8)
        cbp1 = enc.lib.xref_encode_inter_mb(enc.h, ptr(fenc_y), ptr(fenc_c), ptr(y1), ptr(c1), qp, ptr(l1, i16p), ptr(n1))
        cbp2 = o.xo_encode_inter_mb(ptr(fenc_y), ptr(fenc_c), ptr(y2), ptr(c2), qp, ptr(l2, i16p), ptr(n2))
        assert cbp1 == cbp2, f"cbp trial {trial}: {cbp1:#x} vs {cbp2:#x}"
        assert np.array_equal(n1, n2), f"nnz trial {trial}"
        assert np.array_equal(y1[:, :16], y2[:, :16]), f"luma recon trial {trial}"
        assert np.array_equal(c1[:, :8], c2[:, :8]) and np.array_equal(c1[:, 16:24], c2[:, 16:24]), f"chroma recon {trial}"
        # levels: the reference leaves blocks it did not code untouched (zero here); compare where coded
        luma_ok = all(np.array_equal(l1[i * 16:(i + 1) * 16], l2[i * 16:(i + 1) * 16])
                      for i in range(16) if l1[i * 16:(i + 1) * 16].any() or n1[i])
        assert luma_ok, f"luma levels trial {trial}"
        assert np.array_equal(l1[256:264], l2[256:264]), f"chroma dc levels trial {trial}: {l1[256:264]} {l2[256:264]}"
        for i in range(8):
            a, b = l1[264 + i * 16: 280 + i * 16], l2[264 + i * 16: 280 + i * 16]
            if a.any():
                assert np.array_equal(a, b), f"chroma ac levels trial {trial} blk {i}"


@pytest.mark.parametrize("qp", [12, 18, 22, 26, 32, 40, 51])
def test_residual_intra16_mb(enc, qp):
    """x264_mb_encode_i16x16 + intra chroma (I slice, no decimation) on caller-supplied predictions: the reference
    (its predictors swapped for no-ops by the harness) against the oracle's restatement"""
    o = cc.oracle()
    rng = np.random.RandomState(100 + qp)
    o.xo_encode_intra16_mb.restype = C.c_int
    enc.lib.xref_encode_intra16_mb.restype = C.c_int
    for trial in range(300):
        amp = [1, 3, 8, 25, 80][trial % 5]
        pred_y = np.zeros((16, 32), np.uint8)
        pred_c = np.zeros((8, 32), np.uint8)
        kind = trial % 4
        if kind == 0:                                        # DC prediction
            pred_y[:] = rng.randint(20, 236)
            pred_c[:, :16] = rng.randint(20, 236)
            pred_c[:, 16:] = rng.randint(20, 236)
        elif kind == 1:                                      # vertical
            pred_y[:, :16] = rng.randint(0, 256, 16)[None, :]
            pred_c[:, :8] = rng.randint(0, 256, 8)[None, :]
            pred_c[:, 16:24] = rng.randint(0, 256, 8)[None, :]
        elif kind == 2:                                      # horizontal
            pred_y[:, :16] = rng.randint(0, 256, 16)[:, None]
            pred_c[:, :8] = rng.randint(0, 256, 8)[:, None]
            pred_c[:, 16:24] = rng.randint(0, 256, 8)[:, None]
        else:                                                # anything
            pred_y[:] = rng.randint(0, 256, (16, 32))
            pred_c[:] = rng.randint(0, 256, (8, 32))
        fenc_y = np.clip(pred_y[:, :16].astype(int) + rng.randint(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
        if trial % 6 == 0:                                   # DC-only luma change: exercises add16x16_idct_dc
            fenc_y = np.clip(pred_y[:, :16].astype(int) + rng.randint(-6, 7), 0, 255).astype(np.uint8)
        fenc_c = np.zeros((8, 16), np.uint8)
        fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-amp, amp + 1, (8, 8)), 0, 255)
        fenc_c[:, 8:] = np.clip(pred_c[:, 16:24].astype(int) + rng.randint(-amp, amp + 1, (8, 8)), 0, 255)
        if trial % 7 == 0:
            fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-3, 4), 0, 255)
        y1, c1, y2, c2 = pred_y.copy(), pred_c.copy(), pred_y.copy(), pred_c.copy()
        l1, l2 = np.zeros(392, np.int16), np.zeros(392, np.int16)
        d1, d2 = np.zeros(16, np.int16), np.zeros(16, np.int16)
        n1, n2 = np.zeros(27, np.uint8), np.zeros(27, np.uint8)
        cbp1 = enc.lib.xref_encode_intra16_mb(enc.h, ptr(fenc_y), ptr(fenc_c), ptr(y1), ptr(c1), qp, ptr(l1, i16p),
                                              ptr(d1, i16p), ptr(n1))
        cbp2 = o.xo_encode_intra16_mb(ptr(fenc_y), ptr(fenc_c), ptr(y2), ptr(c2), qp, ptr(l2, i16p), ptr(d2, i16p), ptr(n2))
        assert cbp1 == cbp2, f"cbp trial {trial}: {cbp1:#x} vs {cbp2:#x}"
        assert np.array_equal(n1, n2), f"nnz trial {trial}: {n1} {n2}"
        assert np.array_equal(y1[:, :16], y2[:, :16]), f"luma recon trial {trial}"
        assert np.array_equal(c1[:, :8], c2[:, :8]) and np.array_equal(c1[:, 16:24], c2[:, 16:24]), f"chroma recon {trial}"
        for i in range(16):
            if n1[i]:
                assert np.array_equal(l1[i * 16:(i + 1) * 16], l2[i * 16:(i + 1) * 16]), f"luma levels trial {trial} blk {i}"
        if n1[24]:
            assert np.array_equal(d1, d2), f"luma dc levels trial {trial}: {d1} {d2}"
        for ch in range(2):
            if n1[25 + ch]:
                assert np.array_equal(l1[256 + 4 * ch:260 + 4 * ch], l2[256 + 4 * ch:260 + 4 * ch]), f"chroma dc {trial}"
        for i in range(8):
            if n1[16 + i]:
                assert np.array_equal(l1[264 + i * 16: 280 + i * 16], l2[264 + i * 16: 280 + i * 16]), f"chroma ac {trial} blk {i}"


def test_predict_4x4_oracle(enc):
    """the oracle's per-pixel restatement of the twelve 4x4 predictors against x264_predict_4x4_init's functions"""
    o = cc.oracle()
    PRED_T = C.CFUNCTYPE(None, C.c_void_p)
    tab = (PRED_T * 12)()
    enc.lib.x264_predict_4x4_init(0, tab)
    rng = np.random.RandomState(44)
    for mode in range(12):
        for trial in range(40):
            buf = rng.randint(0, 256, (12, 32)).astype(np.uint8)
            if trial < 2:
                buf[:] = 255 * trial
            a, b = buf.copy(), buf.copy()
            tab[mode](C.cast(a.ctypes.data + 4 * 32 + 8, C.c_void_p))
            o.xo_predict_4x4(mode, C.cast(b.ctypes.data + 4 * 32 + 8, C.c_void_p))
            assert np.array_equal(a, b), f"mode {mode} trial {trial}"


@pytest.mark.parametrize("qp", [12, 20, 26, 34, 44, 51])
def test_residual_intra4_mb(enc, qp):
    """x264_macroblock_encode on I4x4 macroblocks (sixteen predict / transform / reconstruct steps, each predicting
    from the blocks before it) with random mode sets: reference against oracle"""
    o = cc.oracle()
    rng = np.random.RandomState(400 + qp)
    o.xo_encode_intra4_mb.restype = C.c_int
    enc.lib.xref_encode_intra4_mb.restype = C.c_int
    for trial in range(200):
        nb = rng.randint(0, 256, (17, 32)).astype(np.uint8)         # row 0 = the row above; origin at (1, 8)
        if trial % 4 == 0:
            nb[:] = rng.randint(40, 200)
        base = rng.randint(0, 256) if trial % 3 else None
        fenc_y = (rng.randint(0, 256, (16, 16)) if base is None else
                  np.clip(base + rng.randint(-30, 31, (16, 16)) + np.arange(16)[None, :] * rng.randint(-3, 4), 0, 255)).astype(np.uint8)
        pred_c = np.zeros((8, 32), np.uint8)
        pred_c[:, :8], pred_c[:, 16:24] = rng.randint(20, 236), rng.randint(20, 236)
        fenc_c = np.zeros((8, 16), np.uint8)
        fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-9, 10, (8, 8)), 0, 255)
        fenc_c[:, 8:] = np.clip(pred_c[:, 16:24].astype(int) + rng.randint(-9, 10, (8, 8)), 0, 255)
        modes = rng.randint(0, 12, 16).astype(np.uint8)
        rep5 = int(trial % 2)
        y1, y2, c1, c2 = nb.copy(), nb.copy(), pred_c.copy(), pred_c.copy()
        l1, l2 = np.zeros(392, np.int16), np.zeros(392, np.int16)
        n1, n2 = np.zeros(27, np.uint8), np.zeros(27, np.uint8)
        org1 = C.cast(y1.ctypes.data + 32 + 8, C.c_void_p)
        org2 = C.cast(y2.ctypes.data + 32 + 8, C.c_void_p)
        cbp1 = enc.lib.xref_encode_intra4_mb(enc.h, ptr(fenc_y), ptr(fenc_c), org1, ptr(c1), qp, ptr(modes), rep5,
                                             ptr(l1, i16p), ptr(n1))
        cbp2 = o.xo_encode_intra4_mb(ptr(fenc_y), ptr(fenc_c), org2, ptr(c2), qp, ptr(modes), rep5, ptr(l2, i16p), ptr(n2))
        assert cbp1 == cbp2, f"cbp trial {trial}: {cbp1:#x} vs {cbp2:#x}"
        assert np.array_equal(n1, n2), f"nnz trial {trial}"
        assert np.array_equal(y1[1:, 8:24], y2[1:, 8:24]), f"luma recon trial {trial} modes {modes}"
        assert np.array_equal(c1[:, :8], c2[:, :8]) and np.array_equal(c1[:, 16:24], c2[:, 16:24]), f"chroma recon {trial}"
        for i in range(16):
            if n1[i]:
                assert np.array_equal(l1[i * 16:(i + 1) * 16], l2[i * 16:(i + 1) * 16]), f"luma levels trial {trial} blk {i}"


@pytest.mark.parametrize("qp", [12, 18, 22, 26, 32, 40, 51])
def test_probe_pskip_mb(enc, qp):
    """x264_macroblock_probe_pskip on supplied P_SKIP predictions (its motion compensation stubbed by the harness):
    reference against oracle, around the decision thresholds"""
    o = cc.oracle()
    rng = np.random.RandomState(700 + qp)
    o.xo_probe_pskip_mb.restype = C.c_int
    enc.lib.xref_probe_pskip_mb.restype = C.c_int
    seen = [0, 0]
    for trial in range(600):
        pred_y = rng.randint(0, 256, (16, 32)).astype(np.uint8)
        pred_c = rng.randint(0, 256, (8, 32)).astype(np.uint8)
        if trial % 3 == 0:
            pred_y[:], pred_c[:] = rng.randint(30, 220), rng.randint(30, 220)
        scale = max(1, (qp - 14) // 5)                       # the thresholds grow with the quantiser
        amp = [0, 1, 2, 3, 5, 8, 14][trial % 7] * scale
        camp = [0, 1, 2, 4, 9][trial % 5] * scale
        fenc_y = np.clip(pred_y[:, :16].astype(int) + rng.randint(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
        if trial % 11 == 0:                                  # one hot 4x4 in an otherwise perfect prediction
            fenc_y = pred_y[:, :16].copy()
            bx, by = rng.randint(4) * 4, rng.randint(4) * 4
            fenc_y[by:by + 4, bx:bx + 4] = np.clip(fenc_y[by:by + 4, bx:bx + 4].astype(int) + rng.randint(-40, 41, (4, 4)), 0, 255)
        fenc_c = np.zeros((8, 16), np.uint8)
        fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-camp, camp + 1, (8, 8)), 0, 255)
        fenc_c[:, 8:] = np.clip(pred_c[:, 16:24].astype(int) + rng.randint(-camp, camp + 1, (8, 8)), 0, 255)
        if trial % 13 == 0:
            fenc_c[:, :8] = np.clip(pred_c[:, :8].astype(int) + rng.randint(-4, 5), 0, 255)       # DC shift only
        r1 = enc.lib.xref_probe_pskip_mb(enc.h, ptr(fenc_y), ptr(fenc_c), ptr(pred_y), ptr(pred_c), qp)
        r2 = o.xo_probe_pskip_mb(ptr(fenc_y), ptr(fenc_c), ptr(pred_y), ptr(pred_c), qp)
        assert r1 == r2, f"trial {trial}: reference {r1} oracle {r2}"
        seen[r1] += 1
    assert seen[0] > 20 and seen[1] > 20, seen


def test_predict_mv_16x16_and_pskip(enc):
    """x264_mb_predict_mv_16x16 / x264_mb_predict_mv_pskip on random neighbourhoods (every availability pattern, equal and
    different references, zero vectors) against the oracle"""
    o = cc.oracle()
    rng = np.random.RandomState(161)
    for trial in range(4000):
        ref = rng.choice([-2, -1, 0, 0, 0, 1], 4).astype(np.int8)
        mv = rng.randint(-40, 41, (4, 2)).astype(np.int16)
        mv[rng.rand(4) < 0.25] = 0
        if trial % 5 == 0:
            mv[:] = mv[0]
        i_ref = int(rng.choice([0, 0, 1]))
        a, b = np.zeros(2, np.int16), np.zeros(2, np.int16)
        enc.lib.xref_predict_mv(enc.h, ptr(ref, i8p), ptr(mv, i16p), i_ref, ptr(a, i16p), ptr(b, i16p))
        nb = np.zeros(20, np.uint8)
        nb[:4] = ref.view(np.uint8)
        nb[4:] = mv.view(np.uint8).reshape(-1)
        c, d = np.zeros(2, np.int16), np.zeros(2, np.int16)
        o.xo_predict_mv_16x16(ptr(nb), i_ref, ptr(c, i16p))
        o.xo_predict_mv_pskip(ptr(nb), ptr(d, i16p))
        assert np.array_equal(a, c), f"mvp trial {trial}: ref {ref} mv {mv.tolist()} i_ref {i_ref}: {a} vs {c}"
        assert np.array_equal(b, d), f"pskip trial {trial}: ref {ref} mv {mv.tolist()}: {b} vs {d}"


def test_predict_mv_partitions(enc):
    """x264_mb_predict_mv for every partition shape the reference analyses (16x16, 16x8 upper / lower, 8x16 left / right,
    the four 8x8s) against the oracle's shape / c_unreachable formulation"""
    o = cc.oracle()
    rng = np.random.RandomState(163)
    # (partition type, idx, width in 4-pixel units, oracle shape)
    cases = [(16, 0, 4, 0), (14, 0, 4, 1), (14, 8, 4, 2), (15, 0, 2, 3), (15, 4, 2, 4),
             (13, 0, 2, 0), (13, 4, 2, 0), (13, 8, 2, 0), (13, 12, 2, 0)]
    for trial in range(6000):
        part, idx, width, shape = cases[trial % len(cases)]
        ref = rng.choice([-2, -1, 0, 0, 0, 1], 4).astype(np.int8)
        mv = rng.randint(-40, 41, (4, 2)).astype(np.int16)
        mv[rng.rand(4) < 0.2] = 0
        i_ref = int(rng.choice([0, 0, 1]))
        a = np.zeros(2, np.int16)
        enc.lib.xref_predict_mv_part(enc.h, ptr(ref, i8p), ptr(mv, i16p), i_ref, part, idx, width, ptr(a, i16p))
        nb = np.zeros(20, np.uint8)
        nb[:4] = ref.view(np.uint8)
        nb[4:] = mv.view(np.uint8).reshape(-1)
        c_unreachable = int((idx & 3) >= 2 + (width & 1))
        c = np.zeros(2, np.int16)
        o.xo_predict_mv_part(ptr(nb), i_ref, shape, c_unreachable, ptr(c, i16p))
        assert np.array_equal(a, c), f"trial {trial} case {cases[trial % len(cases)]}: ref {ref} mv {mv.tolist()} i_ref {i_ref}: {a} vs {c}"


def test_predict_mvc_16x16_frame(enc):
    """x264_mb_predict_mv_ref16x16 for every macroblock of a frame (lowres candidate, the four spatial neighbours with
    the frame-edge rule, the three scaled temporal candidates) against the oracle"""
    o = cc.oracle()
    lib = enc.lib
    lib.xref_frame_new.restype = C.c_void_p
    lib.xref_frame_new.argtypes = [C.c_void_p, C.c_int]
    g = cc.oracle_geom(352, 288)
    W, H, n = g.mb_w, g.mb_h, g.mb_count
    fenc, fref, fdec = lib.xref_frame_new(enc.h, 0), lib.xref_frame_new(enc.h, 1), lib.xref_frame_new(enc.h, 1)
    assert fenc and fref and fdec
    rng = np.random.RandomState(164)
    for trial in range(24):
        lowres = rng.randint(-300, 301, (n, 2)).astype(np.int16) if trial % 3 else None
        if lowres is not None and trial % 6 == 1:
            lowres[rng.rand(n) < 0.3] = [-17000, 16500]          # doubling wraps in 16 bits
        mvr = rng.randint(-200, 201, (n, 2)).astype(np.int16)
        l0 = rng.randint(-200, 201, (n, 2)).astype(np.int16) if trial % 2 else None
        curpoc, refpoc = 2 * (trial + 3), 2 * (trial + 2 - trial % 3)
        inv = (256 + 1) // 2 if trial % 4 else 77
        m1, m2 = np.full((n, 9, 2), 999, np.int16), np.full((n, 9, 2), 999, np.int16)
        n1, n2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib.xref_predict_mvc_frame(enc.h, C.c_void_p(fenc), C.c_void_p(fref), C.c_void_p(fdec),
                                   ptr(lowres, i16p) if lowres is not None else None, ptr(mvr, i16p),
                                   ptr(l0, i16p) if l0 is not None else None, curpoc, refpoc, inv, ptr(m1, i16p), ptr(n1, i32p))
        o.xo_predict_mvc_16x16_frame(W, H, ptr(lowres, i16p) if lowres is not None else None, ptr(mvr, i16p),
                                     ptr(l0, i16p) if l0 is not None else None, (curpoc - refpoc) * inv, ptr(m2, i16p), ptr(n2, i32p))
        assert np.array_equal(n1, n2), f"trial {trial}: counts differ at {np.nonzero(n1 != n2)[0][:5]}: {n1[n1 != n2][:5]} vs {n2[n1 != n2][:5]}"
        assert np.array_equal(m1, m2), f"trial {trial}: candidates differ at mb {np.nonzero((m1 != m2).any((1, 2)))[0][:5]}"
        assert n1.min() >= 4 and n1.max() == 4 + (lowres is not None) + 3 * (l0 is not None)


def test_o3_build_of_the_reference_agrees_with_o2():
    """the -O3 -march=x86-64-v3 build bench.py also times is the same code: identical lookahead and search results"""
    lib3 = cc.ref_o3()
    if lib3 is None:
        pytest.skip("oracle/_ref/o3 not built")
    w, h = 352, 288
    clip = cc.synth_clip(w, h, 3)
    res = []
    for lib in (cc.ref(), lib3):
        enc = cc.RefEncoder(w, h, me=1, subme=5, lib=lib)
        frames = [enc.new_frame(False) for _ in range(3)]
        for f, pic in zip(frames, clip):
            enc.load(f, pic)
        arr = (C.c_void_p * 3)(*[f.value for f in frames])
        costs = (C.c_int * 3)()
        lib.xref_time_lookahead(enc.h, arr, 3, costs)
        g = cc.oracle_geom(w, h)
        fref = enc.new_frame(True)
        enc.load(fref, clip[0])
        lib.xref_frame_filter_all(enc.h, fref)
        blocks = make_me_blocks(g, np.random.RandomState(4), 3, 200, 30)
        out = np.zeros(200, cc.ME_RESULT_DTYPE)
        lib.xref_me_search_batch(enc.h, frames[1], fref, 26, 1, 5, 16, 1, blocks.ctypes.data_as(C.c_void_p), 200,
                                 out.ctypes.data_as(C.c_void_p))
        res.append((list(costs), out.copy()))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
