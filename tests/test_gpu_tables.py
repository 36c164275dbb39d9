"""GPU parity of the drop-in function-pointer tables (include/x264dsp_tables.h): every member that
x264_pixel_init / x264_dct_init / x264_zigzag_init / x264_mc_init / x264_quant_init /
x264_deblock_init of libx264dsp_b200.so fill is called on random operands next to the same member
of the UNMODIFIED reference's table (oracle/_ref/libx264ref.so) and compared bit for bit,
including in-place side effects (fdec after intra_x3, *i_dst_stride after get_ref).  Members the
reference leaves NULL must be NULL here too."""
import ctypes as C
import os

import numpy as np
import pytest

import cpu_checkers as cc
import ref_tables as rt
from cpu_checkers import ptr, i16p, i32p, u16p, i8p

pytestmark = pytest.mark.gpu

BW, BH = cc.BLOCK_W, cc.BLOCK_H


def at(arr, off, t=rt.u8p):
    """pointer `off` BYTES into a numpy array (negative neighbours stay inside the array)"""
    return C.cast(arr.ctypes.data + off, t)


@pytest.fixture(scope="module")
def ref_enc():
    assert cc.ref() is not None, "oracle/_ref/libx264ref.so must travel to the GPU box (make -C oracle ref)"
    return cc.RefEncoder(352, 288, me=1, subme=5, me_range=16, qp=26)


@pytest.fixture(scope="module")
def tabs(pkg, ref_enc):
    if os.environ.get("X264DSP_TABLES_SELFTEST"):
        # harness self-check on a CPU box: compare the reference's tables with themselves
        t = {name: C.cast(getattr(ref_enc.lib, "xref_" + name)(ref_enc.h), C.POINTER(cls)).contents
             for name, cls in (("pixf", rt.PixelTable), ("dctf", rt.DctTable), ("zigzagf", rt.ZigzagTable),
                               ("mcf", rt.McTable), ("loopf", rt.DeblockTable), ("quantf", rt.QuantTable))}
        return t, t
    lib = pkg.lib()
    ours = {}
    for name, cls, init in (("pixf", rt.PixelTable, "x264_pixel_init"), ("dctf", rt.DctTable, "x264_dct_init"),
                            ("zigzagf", rt.ZigzagTable, "x264_zigzag_init"), ("mcf", rt.McTable, "x264_mc_init"),
                            ("loopf", rt.DeblockTable, "x264_deblock_init")):
        t = cls()
        getattr(lib, init)(0, C.byref(t))
        ours[name] = t
    q = rt.QuantTable()
    lib.x264_quant_init(None, 0, C.byref(q))
    ours["quantf"] = q
    theirs = {name: C.cast(getattr(ref_enc.lib, "xref_" + name)(ref_enc.h), C.POINTER(type(t))).contents
              for name, t in ours.items()}
    return ours, theirs


def fptr(f):
    return C.cast(f, C.c_void_p).value


def test_same_members_filled(tabs):
    ours, theirs = tabs
    post_init = {"mbcmp", "mbcmp_unaligned", "fpelcmp", "fpelcmp_x3", "fpelcmp_x4", "intra_mbcmp_x3_16x16",
                 "intra_mbcmp_x3_4x4", "intra_mbcmp_x4_4x4_h", "intra_mbcmp_x4_4x4_v", "intra_mbcmp_x3_chroma",
                 "intra_mbcmp_x3_8x8c", "intra_mbcmp_x9_4x4", "intra_satd_x3_chroma", "intra_sad_x3_chroma", "prefetch_fenc"}
    for name in ours:
        for field, ftype in type(ours[name])._fields_:
            if field in post_init:
                continue      # aliases the ENCODER sets after init (encoder.c:412-457), not the init functions
            a, b = getattr(ours[name], field), getattr(theirs[name], field)
            if hasattr(a, "__len__"):
                for i in range(len(a)):
                    if (field, i) in (("coeff_last", 3), ("coeff_level_run", 3)):
                        continue      # DCT_CHROMA_DC: set by chroma_dsp_init (encoder.c:449-450)
                    assert bool(fptr(a[i])) == bool(fptr(b[i])), f"{name}.{field}[{i}]"
            else:
                av = a if isinstance(a, (int, type(None))) else fptr(a)
                bv = b if isinstance(b, (int, type(None))) else fptr(b)
                assert bool(av) == bool(bv), f"{name}.{field}"


def test_pixel_cmp(tabs):
    ours, theirs = tabs
    po, pr = ours["pixf"], theirs["pixf"]
    rng = np.random.RandomState(11)
    s1, s2, rows = 16, 96, 64
    a, _ = None, None
    from test_oracle_vs_ref import adversarial_planes
    a, _ = adversarial_planes(rng, s1, rows)
    _, b = adversarial_planes(rng, s2, rows)
    for size in range(8):
        bw, bh = BW[size], BH[size]
        for trial in range(6):
            ya, yb, xb = rng.randint(0, rows - bh), rng.randint(0, rows - bh), rng.randint(0, s2 - bw - 3)
            pa, pb = at(a, ya * s1), at(b, yb * s2 + xb)
            for name in ("sad", "ssd", "satd", "sad_aligned"):
                assert getattr(po, name)[size](pa, s1, pb, s2) == getattr(pr, name)[size](pa, s1, pb, s2), (name, size)
            if size < 7:
                refs = [at(b, yb * s2 + xb + k) for k in range(4)]
                for name in ("sad_x3", "satd_x3"):
                    r1, r2 = (C.c_int * 3)(), (C.c_int * 3)()
                    getattr(po, name)[size](pa, *refs[:3], s2, r1)
                    getattr(pr, name)[size](pa, *refs[:3], s2, r2)
                    assert list(r1) == list(r2), (name, size)
                for name in ("sad_x4", "satd_x4"):
                    r1, r2 = (C.c_int * 4)(), (C.c_int * 4)()
                    getattr(po, name)[size](pa, *refs, s2, r1)
                    getattr(pr, name)[size](pa, *refs, s2, r2)
                    assert list(r1) == list(r2), (name, size)


def test_pixel_var_intra(tabs):
    ours, theirs = tabs
    po, pr = ours["pixf"], theirs["pixf"]
    rng = np.random.RandomState(12)
    for trial in range(8):
        p = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        q = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        if trial == 0:
            p[:] = 255
            q[:] = 0
        assert po.var[0](ptr(p), 32) == pr.var[0](ptr(p), 32)
        assert po.var[3](ptr(p), 32) == pr.var[3](ptr(p), 32)
        s1, s2 = C.c_int(), C.c_int()
        assert po.var2[3](ptr(p), 16, ptr(q), 32, C.byref(s1)) == pr.var2[3](ptr(p), 16, ptr(q), 32, C.byref(s2))
        assert s1.value == s2.value
        fenc = rng.randint(0, 256, 16 * 16).astype(np.uint8)
        fdec = rng.randint(0, 256, 32 * 20).astype(np.uint8)
        for name, n in (("intra_sad_x3_4x4", 3), ("intra_satd_x3_4x4", 3), ("intra_sad_x3_8x8c", 3),
                        ("intra_satd_x3_8x8c", 3), ("intra_sad_x3_16x16", 3), ("intra_satd_x3_16x16", 3),
                        ("intra_satd_x4_4x4_h", 9), ("intra_satd_x4_4x4_v", 9)):
            f1, f2 = fdec.copy(), fdec.copy()
            r1, r2 = (C.c_int * 9)(*([-7] * 9)), (C.c_int * 9)(*([-7] * 9))
            getattr(po, name)(ptr(fenc), at(f1, 32 * 2 + 8), r1)
            getattr(pr, name)(ptr(fenc), at(f2, 32 * 2 + 8), r2)
            assert list(r1) == list(r2), name
            assert np.array_equal(f1, f2), name + " fdec side effect"


def test_dct_tables(tabs):
    ours, theirs = tabs
    do, dr = ours["dctf"], theirs["dctf"]
    rng = np.random.RandomState(13)
    for trial in range(8):
        fenc = rng.randint(0, 256, 16 * 16).astype(np.uint8)
        fdec = rng.randint(0, 256, 32 * 16).astype(np.uint8)
        if trial == 0:
            fenc[:], fdec[:] = 255, 0
        if trial == 1:
            fenc[:], fdec[:] = 0, 255
        for name, n in (("sub4x4_dct", 16), ("sub8x8_dct", 64), ("sub16x16_dct", 256), ("sub8x8_dct_dc", 4)):
            r, g = np.zeros(n, np.int16), np.zeros(n, np.int16)
            getattr(dr, name)(ptr(r, i16p), ptr(fenc), ptr(fdec))
            getattr(do, name)(ptr(g, i16p), ptr(fenc), ptr(fdec))
            assert np.array_equal(r, g), name
        coef = rng.randint(-2000, 2000, 256).astype(np.int16)
        if trial < 3:
            coef = rng.randint(-32768, 32768, 256).astype(np.int16)
        for name, n in (("add4x4_idct", 16), ("add8x8_idct", 64), ("add16x16_idct", 256),
                        ("add8x8_idct_dc", 4), ("add16x16_idct_dc", 16)):
            d1, d2 = fdec.copy(), fdec.copy()
            c1, c2 = coef[:n].copy(), coef[:n].copy()
            getattr(dr, name)(ptr(d1), ptr(c1, i16p))
            getattr(do, name)(ptr(d2), ptr(c2, i16p))
            assert np.array_equal(d1, d2), name
        for name in ("dct4x4dc", "idct4x4dc"):
            c1, c2 = coef[:16].copy(), coef[:16].copy()
            getattr(dr, name)(ptr(c1, i16p))
            getattr(do, name)(ptr(c2, i16p))
            assert np.array_equal(c1, c2), name
        l1, l2 = np.zeros(16, np.int16), np.zeros(16, np.int16)
        theirs["zigzagf"].scan_4x4(ptr(l1, i16p), ptr(coef, i16p))
        ours["zigzagf"].scan_4x4(ptr(l2, i16p), ptr(coef, i16p))
        assert np.array_equal(l1, l2)


def test_quant_tables(pkg, tabs):
    ours, theirs = tabs
    qo, qr = ours["quantf"], theirs["quantf"]
    rng = np.random.RandomState(14)
    dq = pkg.dequant_table().astype(np.int32)
    for qp in (0, 5, 12, 20, 22, 23, 26, 35, 36, 51):
        mf, bias = pkg.quant_tables(qp & 1, qp)
        for trial in range(5):
            scale = [4, 40, 400, 4000, 30000][trial % 5]
            coef = rng.randint(-scale, scale + 1, 16).astype(np.int16)
            c1, c2 = coef.copy(), coef.copy()
            assert qr.quant_4x4(ptr(c1, i16p), ptr(mf, u16p), ptr(bias, u16p)) == qo.quant_4x4(ptr(c2, i16p), ptr(mf, u16p), ptr(bias, u16p))
            assert np.array_equal(c1, c2), f"quant_4x4 qp {qp}"
            for name, n in (("quant_4x4_dc", 16), ("quant_2x2_dc", 4)):
                c1, c2 = coef[:n].copy(), coef[:n].copy()
                a = getattr(qr, name)(ptr(c1, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                b = getattr(qo, name)(ptr(c2, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                assert a == b and np.array_equal(c1, c2), name
            lv = rng.randint(-40, 41, 16).astype(np.int16)
            for name in ("dequant_4x4", "dequant_4x4_dc"):
                c1, c2 = lv.copy(), lv.copy()
                getattr(qr, name)(ptr(c1, i16p), ptr(dq, i32p), qp)
                getattr(qo, name)(ptr(c2, i16p), ptr(dq, i32p), qp)
                assert np.array_equal(c1, c2), f"{name} qp {qp}"
            dmf = int(dq[qp % 6][0]) << (qp // 6)
            small = rng.randint(-6, 7, 4).astype(np.int16)
            c1, c2 = small.copy(), small.copy()
            assert qr.optimize_chroma_2x2_dc(ptr(c1, i16p), dmf) == qo.optimize_chroma_2x2_dc(ptr(c2, i16p), dmf)
            assert np.array_equal(c1, c2), f"optimize_chroma_2x2_dc qp {qp}"
    for size in (16, 64):
        d = rng.randint(-300, 300, size).astype(np.int16)
        off = rng.randint(0, 50, size).astype(np.uint16)
        s1 = rng.randint(0, 1000, size).astype(np.uint32)
        s2, d1, d2 = s1.copy(), d.copy(), d.copy()
        qr.denoise_dct(ptr(d1, i16p), ptr(s1, C.POINTER(C.c_uint32)), ptr(off, u16p), size)
        qo.denoise_dct(ptr(d2, i16p), ptr(s2, C.POINTER(C.c_uint32)), ptr(off, u16p), size)
        assert np.array_equal(d1, d2) and np.array_equal(s1, s2), "denoise_dct"
    for trial in range(40):
        lv = (rng.randint(-2, 3, 64) * (rng.rand(64) < 0.35)).astype(np.int16)
        if trial == 0:
            lv[:] = 0
        assert qr.decimate_score15(ptr(lv, i16p)) == qo.decimate_score15(ptr(lv, i16p))
        assert qr.decimate_score16(ptr(lv, i16p)) == qo.decimate_score16(ptr(lv, i16p))
        for cat in range(14):
            if fptr(qr.coeff_last[cat]) and cat != 3:      # [3] is filled by the encoder after init
                assert qr.coeff_last[cat](ptr(lv, i16p)) == qo.coeff_last[cat](ptr(lv, i16p)), f"coeff_last[{cat}]"
        assert qr.coeff_last4(ptr(lv, i16p)) == qo.coeff_last4(ptr(lv, i16p))
        assert qr.coeff_last8(ptr(lv, i16p)) == qo.coeff_last8(ptr(lv, i16p))
        fns = [(qr.coeff_level_run[c], qo.coeff_level_run[c], 15 if c in (1, 4, 7, 11) else 16) for c in range(13)
               if fptr(qr.coeff_level_run[c]) and c != 3]
        fns += [(qr.coeff_level_run4, qo.coeff_level_run4, 4), (qr.coeff_level_run8, qo.coeff_level_run8, 8)]
        for fr, fo, n in fns:
            if not lv[:n].any():
                continue                       # the reference reads dct[-1] on an all-zero block (callers never do)
            r1, r2 = rt.RunLevel(), rt.RunLevel()
            t1, t2 = fr(ptr(lv, i16p), C.byref(r1)), fo(ptr(lv, i16p), C.byref(r2))
            assert t1 == t2 and r1.last == r2.last and r1.mask == r2.mask, f"level_run{n}"
            assert list(r1.level)[:t1] == list(r2.level)[:t2], f"level_run{n} levels"


def test_mc_tables(tabs):
    ours, theirs = tabs
    mo, mr = ours["mcf"], theirs["mcf"]
    rng = np.random.RandomState(15)
    stride, rows = 128, 96
    planes = [rng.randint(0, 256, stride * rows).astype(np.uint8) for _ in range(4)]
    org = 32 * stride + 40
    src = (rt.u8p * 4)(*[at(p, org) for p in planes])
    for trial in range(40):
        mvx, mvy = int(rng.randint(-40, 41)), int(rng.randint(-40, 41))
        if trial < 16:
            mvx, mvy = (trial & 3) + 4 * int(rng.randint(-3, 4)), (trial >> 2) + 4 * int(rng.randint(-3, 4))
        w, h = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (20, 18)][trial % 8]
        d1, d2 = np.full(32 * 32, 9, np.uint8), np.full(32 * 32, 9, np.uint8)
        mr.mc_luma(ptr(d1), 32, src, stride, mvx, mvy, w, h, None)
        mo.mc_luma(ptr(d2), 32, src, stride, mvx, mvy, w, h, None)
        assert np.array_equal(d1, d2), f"mc_luma mv ({mvx},{mvy}) {w}x{h}"
        d1[:], d2[:] = 7, 7
        s1, s2 = C.c_ssize_t(32), C.c_ssize_t(32)
        p1 = mr.get_ref(ptr(d1), C.byref(s1), src, stride, mvx, mvy, w, h, None)
        p2 = mo.get_ref(ptr(d2), C.byref(s2), src, stride, mvx, mvy, w, h, None)
        assert s1.value == s2.value, "get_ref stride"
        assert (p1 - d1.ctypes.data) == (p2 - d2.ctypes.data) if p1 == d1.ctypes.data else p1 == p2, "get_ref pointer"
        assert np.array_equal(d1, d2), f"get_ref mv ({mvx},{mvy}) {w}x{h}"
        # chroma: NV12 plane, 1/8 pel
        cw, ch = [(8, 8), (8, 4), (4, 8), (4, 4)][trial % 4]
        u1, v1, u2, v2 = (np.full(32 * 16, 3, np.uint8) for _ in range(4))
        mr.mc_chroma(ptr(u1), ptr(v1), 32, at(planes[0], org), stride, mvx, mvy, cw, ch)
        mo.mc_chroma(ptr(u2), ptr(v2), 32, at(planes[0], org), stride, mvx, mvy, cw, ch)
        assert np.array_equal(u1, u2) and np.array_equal(v1, v2), f"mc_chroma ({mvx},{mvy}) {cw}x{ch}"
    for idx, w in ((0, 16), (3, 8), (6, 4)):
        d1, d2 = np.zeros(32 * 16, np.uint8), np.zeros(32 * 16, np.uint8)
        mr.copy[idx](ptr(d1), 32, at(planes[1], org + 3), stride, 16)
        mo.copy[idx](ptr(d2), 32, at(planes[1], org + 3), stride, 16)
        assert np.array_equal(d1, d2) and d1.any(), f"copy w{w}"
    # chroma (de)interleave helpers on fenc / fdec shaped buffers
    for name, dst_n in (("load_deinterleave_chroma_fenc", 16 * 8), ("load_deinterleave_chroma_fdec", 32 * 8)):
        d1, d2 = np.zeros(dst_n, np.uint8), np.zeros(dst_n, np.uint8)
        getattr(mr, name)(ptr(d1), at(planes[2], org), stride, 8)
        getattr(mo, name)(ptr(d2), at(planes[2], org), stride, 8)
        assert np.array_equal(d1, d2) and d1.any(), name
    fd = rng.randint(0, 256, 32 * 8).astype(np.uint8)
    d1, d2 = np.zeros(64 * 8, np.uint8), np.zeros(64 * 8, np.uint8)
    mr.store_interleave_chroma(ptr(d1), 64, ptr(fd), at(fd, 16), 8)
    mo.store_interleave_chroma(ptr(d2), 64, ptr(fd), at(fd, 16), 8)
    assert np.array_equal(d1, d2) and d1.any(), "store_interleave_chroma"
    w, h = 52, 10
    d1, d2 = np.zeros(stride * h, np.uint8), np.zeros(stride * h, np.uint8)
    mr.plane_copy(ptr(d1), stride, at(planes[3], org), stride, w, h)
    mo.plane_copy(ptr(d2), stride, at(planes[3], org), stride, w, h)
    assert np.array_equal(d1, d2) and d1.any(), "plane_copy"
    d1[:], d2[:] = 0, 0
    mr.plane_copy_interleave(ptr(d1), stride, at(planes[0], org), stride, at(planes[1], org), stride, w, h)
    mo.plane_copy_interleave(ptr(d2), stride, at(planes[0], org), stride, at(planes[1], org), stride, w, h)
    assert np.array_equal(d1, d2) and d1.any(), "plane_copy_interleave"
    a1, b1, a2, b2 = (np.zeros(stride * h, np.uint8) for _ in range(4))
    mr.plane_copy_deinterleave(ptr(a1), stride, ptr(b1), stride, at(planes[0], org), stride, w, h)
    mo.plane_copy_deinterleave(ptr(a2), stride, ptr(b2), stride, at(planes[0], org), stride, w, h)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and a1.any(), "plane_copy_deinterleave"


def test_hpel_and_lowres_tables(tabs):
    ours, theirs = tabs
    mo, mr = ours["mcf"], theirs["mcf"]
    rng = np.random.RandomState(16)
    stride, rows, width, height = 160, 64, 112, 16
    src = rng.randint(0, 256, stride * rows).astype(np.uint8)
    src[: stride * 20] = np.where(rng.rand(stride * 20) < 0.5, 0, 255)      # clipping paths of the six-tap
    org = 16 * stride + 24
    outs = []
    for m in (mr, mo):
        h_, v_, c_ = (np.zeros(stride * rows, np.uint8) for _ in range(3))
        buf = np.zeros(width + 16, np.int16)
        m.hpel_filter(at(h_, org), at(v_, org), at(c_, org), at(src, org), stride, width, height, ptr(buf, i16p))
        outs.append((h_, v_, c_))
    for a, b, name in zip(outs[0], outs[1], ("dsth", "dstv", "dstc")):
        assert np.array_equal(a, b) and a.any(), "hpel_filter " + name
    lw, lh, ls = 40, 20, 64
    outs = []
    for m in (mr, mo):
        d = [np.zeros(ls * lh, np.uint8) for _ in range(4)]
        m.frame_init_lowres_core(at(src, org), ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), stride, ls, lw, lh)
        outs.append(d)
    for k in range(4):
        assert np.array_equal(outs[0][k], outs[1][k]) and outs[0][k].any(), f"frame_init_lowres_core plane {k}"


def test_deblock_tables(tabs):
    ours, theirs = tabs
    lo, lr = ours["loopf"], theirs["loopf"]
    rng = np.random.RandomState(17)
    stride = 64
    for trial in range(30):
        # smooth base + a step at the edge so that the filters actually fire, plus noise
        base = rng.randint(40, 200)
        img = (base + rng.randint(-3, 4, (32, stride))).astype(np.int32)
        step = int(rng.randint(-12, 13))
        img[16:, :] += step
        img[:, 16:] += step
        img = np.clip(img, 0, 255).astype(np.uint8).ravel()
        alpha, beta = int(rng.randint(0, 60)), int(rng.randint(0, 19))
        tc0 = np.array([rng.randint(-1, 6) for _ in range(4)], np.int8)
        org = 16 * stride + 16
        for k in range(2):
            for name, inter in (("deblock_luma", 1), ("deblock_chroma", 1), ("deblock_luma_intra", 0), ("deblock_chroma_intra", 0)):
                i1, i2 = img.copy(), img.copy()
                if inter:
                    getattr(lr, name)[k](at(i1, org), stride, alpha, beta, ptr(tc0, i8p))
                    getattr(lo, name)[k](at(i2, org), stride, alpha, beta, ptr(tc0, i8p))
                else:
                    getattr(lr, name)[k](at(i1, org), stride, alpha, beta)
                    getattr(lo, name)[k](at(i2, org), stride, alpha, beta)
                assert np.array_equal(i1, i2), f"{name}[{k}] alpha {alpha} beta {beta} tc0 {tc0}"
    for trial in range(20):
        nnz = (rng.rand(120) < 0.2).astype(np.uint8)
        refi = rng.randint(-1, 2, (2, 40)).astype(np.int8)
        mv = rng.randint(-6, 7, (2, 40, 2)).astype(np.int16)
        b1, b2 = np.full((2, 8, 4), 9, np.uint8), np.full((2, 8, 4), 9, np.uint8)
        lr.deblock_strength(ptr(nnz), ptr(refi, i8p), ptr(mv, i16p), ptr(b1))
        lo.deblock_strength(ptr(nnz), ptr(refi, i8p), ptr(mv, i16p), ptr(b2))
        assert np.array_equal(b1, b2), "deblock_strength"


PRED_T = C.CFUNCTYPE(None, C.c_void_p)


def predict_tables(lib):
    t = [(PRED_T * 7)(), (PRED_T * 7)(), (PRED_T * 12)()]
    lib.x264_predict_16x16_init(0, t[0])
    lib.x264_predict_8x8c_init(0, t[1])
    lib.x264_predict_4x4_init(0, t[2])
    return t


def test_predict_tables(pkg, ref_enc):
    """all 26 intra predictors (x264_predict_16x16_init / _8x8c_init / _4x4_init, common/predict.c:474-546) against the
    reference's own functions on random neighbourhoods; the whole buffer is compared, so nothing outside the block
    may change"""
    ours, theirs = predict_tables(pkg.lib()), predict_tables(ref_enc.lib)
    rng = np.random.RandomState(11)
    for t, size in ((0, 16), (1, 8), (2, 4)):
        for mode in range(len(ours[t])):
            for trial in range(12):
                buf = rng.randint(0, 256, (40, 32)).astype(np.uint8)
                if trial == 0:
                    buf[:] = 255
                elif trial == 1:
                    buf[:] = 0
                elif trial == 2:                              # steep gradients: the plane predictors must clip
                    buf[:] = np.clip(np.add.outer(np.arange(40) * 17, np.arange(32) * 23) - 300, 0, 255)
                a, b = buf.copy(), buf.copy()
                off = 8 * 32 + 8
                theirs[t][mode](at(a, off))
                ours[t][mode](at(b, off))
                assert np.array_equal(a, b), f"predict size {size} mode {mode} trial {trial}"
                assert not np.array_equal(a, buf) or trial < 2, f"predict size {size} mode {mode}: nothing written"


@pytest.mark.parametrize("w,h,n,me,subme,qp", [(96, 64, 4, 1, 5, 26), (64, 48, 3, 0, 2, 20), (176, 144, 5, 1, 4, 32)])
def test_reference_encoder_runs_on_our_tables(pkg, ctx, w, h, n, me, subme, qp):
    """THE drop-in check: the unmodified reference encoder (x264_encoder_encode: lookahead, analysis,
    ME, residual, deblock, CABAC) with its six tables and its three intra predictor tables replaced by ours
    must emit the same bitstream, byte for byte, as with its own tables."""
    lib = cc.ref()
    assert lib is not None
    clip = np.concatenate(cc.synth_clip(w, h, n, seed=77, cut_frame=2))
    outs = []
    for use_ours in (False, True):
        enc = cc.RefEncoder(w, h, me=me, subme=subme, me_range=16, qp=qp, psub16x16=1)
        if use_ours:
            t = [rt.PixelTable(), rt.DctTable(), rt.ZigzagTable(), rt.McTable(), rt.QuantTable(), rt.DeblockTable()]
            plib = pkg.lib()
            plib.x264_pixel_init(0, C.byref(t[0]))
            plib.x264_dct_init(0, C.byref(t[1]))
            plib.x264_zigzag_init(0, C.byref(t[2]))
            plib.x264_mc_init(0, C.byref(t[3]))
            plib.x264_quant_init(None, 0, C.byref(t[4]))
            plib.x264_deblock_init(0, C.byref(t[5]))
            lib.xref_install_tables(enc.h, *[C.byref(x) for x in t])
            pt = predict_tables(plib)
            lib.xref_install_predict_tables(enc.h, pt[0], pt[1], pt[2])
        out = np.zeros(1 << 20, np.uint8)
        launches0 = ctx.launches
        size = lib.xref_encode_clip(enc.h, ptr(clip), n, ptr(out), out.size)
        assert size > 0, size
        outs.append(out[:size].copy())
    assert outs[0].size == outs[1].size and np.array_equal(outs[0], outs[1]), \
        f"bitstreams differ: {outs[0].size} vs {outs[1].size} bytes"
    shim_ctx_launches = pkg.lib().x264dsp_launch_count(C.c_void_p(pkg.lib().x264dsp_tables_context()))
    assert shim_ctx_launches > 1000, "the encode must have gone through the CUDA shims"
