"""The PRODUCT's constant tables (x264-dsp_b200/csrc/ctx.cu: lambda, cost_mv, chroma QP, flat-CQM quant / dequant)
against the reference build oracle/_ref for every QP -- encoder/analyse.c:98-111, 171-315; common/set.c:265-353;
common/macroblock.h:251-266.  The host copies need no GPU; the device copy of cost_mv is read back in a `gpu` test."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc

u16p = C.POINTER(C.c_uint16)


@pytest.fixture(scope="module")
def enc():
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return cc.RefEncoder(352, 288)


def ref_cost_mv(enc, qp):
    base = C.addressof(enc.lib.xref_cost_mv(enc.h, qp).contents) - 2 * 4096
    return np.ctypeslib.as_array(C.cast(base, u16p), shape=(8193,))


def test_product_host_tables_match_reference(pkg, enc):
    lib = pkg.lib()
    for qp in range(52):
        assert lib.x264dsp_lambda(qp) == enc.lib.xref_lambda(qp), f"lambda qp {qp}"
        assert lib.x264dsp_chroma_qp(qp) == enc.lib.xref_chroma_qp(enc.h, qp), f"chroma qp {qp}"
        assert np.array_equal(pkg.cost_mv_table(qp), ref_cost_mv(enc, qp)), f"cost_mv qp {qp}"
        for cat in range(4):                       # CQM_4IY, 4PY, 4IC, 4PC (common/set.h:61-64): odd = inter
            mf_r, b_r = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
            enc.lib.xref_quant_tables(enc.h, cat, qp, cc.ptr(mf_r, u16p), cc.ptr(b_r, u16p))
            mf, bias = pkg.quant_tables(cat & 1, qp)
            assert np.array_equal(mf, mf_r) and np.array_equal(bias, b_r), f"quant tables cat {cat} qp {qp}"
    dq_r = np.zeros((6, 16), np.int32)
    for cat in range(4):
        enc.lib.xref_dequant_table(enc.h, cat, cc.ptr(dq_r, cc.i32p))
        assert np.array_equal(pkg.dequant_table(), dq_r), f"dequant cat {cat}"


def test_product_host_tables_match_oracle(pkg):
    """the same against the oracle restatement (runs wherever the oracle builds)"""
    o = cc.oracle()
    for qp in range(52):
        t = np.zeros(8193, np.uint16)
        o.xo_cost_mv_table(qp, cc.ptr(t, u16p))
        assert np.array_equal(pkg.cost_mv_table(qp), t), f"cost_mv qp {qp}"
        assert pkg.lib().x264dsp_lambda(qp) == o.xo_lambda(qp)
        assert pkg.lib().x264dsp_chroma_qp(qp) == o.xo_chroma_qp(qp)


@pytest.mark.gpu
def test_product_device_cost_mv_matches_reference(pkg, ctx, enc):
    """cost_mv[qp] as the kernels see it (one device table per distinct lambda, shared between QPs)"""
    for qp in range(52):
        t = np.zeros(8193, np.uint16)
        pkg.check(pkg.lib().x264dsp_cost_mv_table_dev(ctx._h, qp, cc.ptr(t, u16p)), "x264dsp_cost_mv_table_dev")
        assert np.array_equal(t, ref_cost_mv(enc, qp)), f"device cost_mv qp {qp}"
