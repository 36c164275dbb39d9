"""CPU check of device leaf routines that are written as plain C++ (x264-dsp_b200/csrc/dbfilter.cuh): compiled with g++
and compared line by line with the oracle's deblocking filters (pinned to the reference in tests/test_oracle_vs_ref.py).
The kernels that call them are compared with the oracle on the GPU (tests/test_gpu_residual_deblock.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr, i8p

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def chk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostleaf") / "libdbfilter_check.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(HERE, "host", "dbfilter_check.cpp"), "-o", so],
                   check=True)
    return C.CDLL(so)


def lines(rng, n, width):
    """sample lines around an edge: smooth, stepped and random ones so that every branch of the filters is taken"""
    base = rng.randint(0, 256, (n, 1))
    step = rng.randint(-40, 41, (n, 1)) * (rng.rand(n, 1) < 0.7)
    noise = rng.randint(-6, 7, (n, width)) * (rng.rand(n, 1) < 0.8)
    ramp = np.arange(width)[None, :] * rng.randint(-3, 4, (n, 1))
    s = base + noise + ramp
    s[:, width // 2:] += step
    rnd = rng.rand(n) < 0.1
    s[rnd] = rng.randint(0, 256, (rnd.sum(), width))
    return np.clip(s, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("alpha,beta", [(4, 2), (13, 4), (40, 10), (127, 15), (255, 18), (0, 5), (20, 0)])
def test_luma_filters_match_oracle(chk, alpha, beta):
    o = cc.oracle()
    rng = np.random.RandomState(alpha * 31 + beta)
    n = 4000
    L = lines(rng, n, 8)
    for tc0 in (-1, 0, 1, 2, 5, 13, 25):
        want = L.copy()
        for i in range(0, n, 16):                      # the oracle filters 16 lines at a time, tc0 per 4 lines
            blk = np.ascontiguousarray(want[i:i + 16])
            t = np.full(4, tc0, np.int8)
            o.xo_deblock_luma(ptr(blk[0:, 4:]), C.c_ssize_t(8), 0, alpha, beta, ptr(t, i8p))
            want[i:i + 16] = blk
        got = L.astype(np.int32)
        for i in range(n):
            s = np.ascontiguousarray(got[i])
            chk.chk_luma_normal(s.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta, tc0, int(tc0 >= 0))
            got[i] = s
        assert np.array_equal(got, want), f"bS<4 luma, tc0 {tc0}: {np.count_nonzero((got != want).any(1))} lines differ"
        assert tc0 < 0 or alpha == 0 or beta == 0 or (want != L).any(), "no line was filtered: the case checks nothing"
    want = L.copy()
    for i in range(0, n, 16):
        blk = np.ascontiguousarray(want[i:i + 16])
        o.xo_deblock_luma_intra(ptr(blk[0:, 4:]), C.c_ssize_t(8), 0, alpha, beta)
        want[i:i + 16] = blk
    got = L.astype(np.int32)
    for i in range(n):
        s = np.ascontiguousarray(got[i])
        chk.chk_luma_intra(s.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta, 1)
        got[i] = s
    assert np.array_equal(got, want), f"bS=4 luma: {np.count_nonzero((got != want).any(1))} lines differ"
    # act = 0 leaves everything alone
    s = np.ascontiguousarray(L[0].astype(np.int32))
    chk.chk_luma_intra(s.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta, 0)
    chk.chk_luma_normal(s.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta, 3, 0)
    assert np.array_equal(s, L[0])


@pytest.mark.parametrize("alpha,beta", [(4, 2), (13, 4), (40, 10), (255, 18), (0, 5)])
def test_chroma_filters_match_oracle(chk, alpha, beta):
    """NV12 lines: the oracle filters 8 rows of U/V pairs across a vertical edge (dir 0: xstride 2)"""
    o = cc.oracle()
    rng = np.random.RandomState(alpha * 7 + beta)
    n = 4000                                           # lines of one component
    Lc = lines(rng, n, 4)
    for tc0 in (0, 1, 4, 12, None):                    # None: the bS = 4 filter
        # pack as rows of interleaved U,V: row r holds lines 2r (U) and 2r+1 (V), 4 samples each
        nv = np.zeros((n // 2, 8), np.uint8)
        nv[:, 0::2], nv[:, 1::2] = Lc[0::2], Lc[1::2]
        want = nv.copy()
        for i in range(0, n // 2, 8):
            blk = np.ascontiguousarray(want[i:i + 8])
            if tc0 is None:
                o.xo_deblock_chroma_intra(ptr(blk[0:, 4:]), C.c_ssize_t(8), 0, alpha, beta)
            else:
                t = np.full(4, tc0 + 1, np.int8)
                o.xo_deblock_chroma(ptr(blk[0:, 4:]), C.c_ssize_t(8), 0, alpha, beta, ptr(t, i8p))
            want[i:i + 8] = blk
        got = Lc.astype(np.int32)
        for i in range(n):
            s = np.ascontiguousarray(got[i])
            chk.chk_chroma(s.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta, 0 if tc0 is None else tc0 + 1, int(tc0 is None), 1)
            got[i] = s
        back = np.zeros_like(nv)
        back[:, 0::2], back[:, 1::2] = got[0::2], got[1::2]
        assert np.array_equal(back, want), f"chroma tc0 {tc0}: {np.count_nonzero((back != want).any(1))} rows differ"
