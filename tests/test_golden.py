"""Golden vectors of the UNMODIFIED reference (tests/golden/golden_r01.npz, produced by
tests/golden/make_golden.py from oracle/_ref in the dev container).  Two consumers:

  * CPU (`-m "not gpu"`): the oracle (oracle/xo_*.c) must reproduce every vector -- this is what
    pins the oracle on a machine without /root/reference;
  * GPU (`-m gpu`): the CUDA path, called through the C ABI (frame-batched entry points and the
    drop-in function-pointer tables), must reproduce the same vectors bit for bit.
"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import cpu_checkers as cc
import ref_tables as rt
from cpu_checkers import ptr, i16p, i32p, u16p, i8p

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r01.npz")
SIZES_WITH_FRAMES = ("s", "r", "cif")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


def frames_of(G, pkg, tag):
    w, h, n, cut = (int(x) for x in G[f"{tag}_wh"])
    if f"{tag}_in" in G.files:
        frames = [np.ascontiguousarray(f) for f in G[f"{tag}_in"]]
    else:
        frames = [pkg.synth_frame(w, h, i, cut_frame=cut) for i in range(n)]
    assert [sha(f) for f in frames] == list(G[f"{tag}_in_sha"]), "synthetic input drifted: regenerate the golden file"
    return w, h, frames


def at(arr, off, t=rt.u8p):
    return C.cast(arr.ctypes.data + off, t)


# =============================================================================================
# CPU: oracle vs golden
# =============================================================================================

def test_golden_file_is_complete(G):
    assert len(G.files) >= 150
    assert G["cli_cif30"][2] == "117267"


def test_oracle_tables(G):
    o = cc.oracle()
    for qp in range(52):
        t = np.zeros(8193, np.uint16)
        o.xo_cost_mv_table(qp, ptr(t, u16p))
        assert sha(t) == G["tab_cost_mv_sha"][qp], f"cost_mv qp {qp}"
        assert o.xo_lambda(qp) == G["tab_lambda"][qp] and o.xo_chroma_qp(qp) == G["tab_chroma_qp"][qp]
        for cat in range(4):
            mf, b = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
            o.xo_quant_tables(cat & 1, qp, ptr(mf, u16p), ptr(b, u16p))
            assert np.array_equal(mf, G["tab_quant_mf"][cat, qp]) and np.array_equal(b, G["tab_quant_bias"][cat, qp])
    dq = np.zeros((6, 16), np.int32)
    o.xo_dequant_table(ptr(dq, i32p))
    assert np.array_equal(dq, G["tab_dequant"])


def pixel_cases(G):
    a, b = np.ascontiguousarray(G["pix_a"]), np.ascontiguousarray(G["pix_b"])
    return a, b, 16, 96


def test_oracle_pixel(G):
    o = cc.oracle()
    a, b, s1, s2 = pixel_cases(G)
    for size, ya, yb, xb, sad, ssd, satd in G["pix_cases"]:
        pa, pb = a[ya * s1:], b[yb * s2 + xb:]
        got = [o.xo_cmp(k, int(size), ptr(pa), C.c_ssize_t(s1), ptr(pb), C.c_ssize_t(s2)) for k in range(3)]
        assert got == [sad, ssd, satd], f"size {size} ya {ya} yb {yb} xb {xb}"
    for row in G["pix_x_cases"]:
        size, ya, offs, r4, r3 = int(row[0]), int(row[1]), row[2:6], row[6:10], row[10:13]
        pa = a[ya * s1:]
        assert [o.xo_sad(size, ptr(pa), C.c_ssize_t(s1), ptr(b[k:]), C.c_ssize_t(s2)) for k in offs] == list(r4)
        assert [o.xo_satd(size, ptr(pa), C.c_ssize_t(s1), ptr(b[k:]), C.c_ssize_t(s2)) for k in offs[:3]] == list(r3)
    for o1, o2, v16, v8, v2, ssd in G["pix_var_cases"]:
        o1, o2 = int(o1), int(o2)
        assert o.xo_var(0, ptr(b[o2:]), C.c_ssize_t(s2)) == v16 and o.xo_var(3, ptr(b[o2:]), C.c_ssize_t(s2)) == v8
        s = C.c_int()
        assert o.xo_var2_8x8(ptr(a[o1:]), C.c_ssize_t(16), ptr(b[o2:]), C.c_ssize_t(s2), C.byref(s)) == np.int64(v2).astype(np.int32)
        assert s.value == int(ssd)
    outs = np.zeros((30, 2, 320), np.uint8)
    for t in range(30):
        for k in range(2):
            f = G["intra_fdec"][t].copy()
            r = (C.c_int * 3)()
            o.xo_intra_x3_8x8c(1 - k, ptr(np.ascontiguousarray(G["intra_fenc"][t])), ptr(f[40:]), r)
            assert list(r) == list(G["intra_res"][t, k])
            outs[t, k] = f
    assert sha(outs) == str(G["intra_out_sha"])


def test_oracle_dct_quant(G):
    o = cc.oracle()
    fenc, fdec, coef = (np.ascontiguousarray(G[k]) for k in ("dct_fenc", "dct_fdec", "dct_coef"))
    T = len(fenc)
    for name, n in (("sub4x4_dct", 16), ("sub8x8_dct", 64), ("sub16x16_dct", 256), ("sub8x8_dct_dc", 4)):
        for t in range(T):
            out = np.zeros(n, np.int16)
            getattr(o, "xo_" + name)(ptr(out, i16p), ptr(fenc[t]), ptr(fdec[t]))
            assert np.array_equal(out, G["dct_" + name][t]), name
    for name, n in (("add4x4_idct", 16), ("add8x8_idct", 64), ("add16x16_idct", 256), ("add8x8_idct_dc", 4),
                    ("add16x16_idct_dc", 16)):
        for t in range(T):
            d, c = fdec[t].copy(), coef[t, :n].copy()
            getattr(o, "xo_" + name)(ptr(d), ptr(c, i16p))
            assert np.array_equal(d, G["dct_" + name][t]), name
    for name in ("dct4x4dc", "idct4x4dc"):
        for t in range(T):
            c = coef[t, :16].copy()
            getattr(o, "xo_" + name)(ptr(c, i16p))
            assert np.array_equal(c, G["dct_" + name][t]), name
    for t in range(T):
        l = np.zeros(16, np.int16)
        o.xo_zigzag_4x4(ptr(l, i16p), ptr(coef[t], i16p))
        assert np.array_equal(l, G["dct_zigzag"][t])
    dq = np.ascontiguousarray(G["tab_dequant"])
    for qi, qp in enumerate(G["q_qps"]):
        qp = int(qp)
        for inter in (0, 1):
            mf = np.ascontiguousarray(G["tab_quant_mf"][inter, qp])
            bias = np.ascontiguousarray(G["tab_quant_bias"][inter, qp])
            for s in range(5):
                c = G["q_in"][s, qi, inter].copy()
                nz = o.xo_quant_4x4(ptr(c, i16p), ptr(mf, u16p), ptr(bias, u16p))
                assert nz == G["q_nz"][s, qi, inter, 0] and np.array_equal(c, G["q_out"][s, qi, inter, :, 0])
                c = G["q_in"][s, qi, inter].copy()
                nz = o.xo_quant_4x4_dc(ptr(c, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                assert nz == G["q_nz"][s, qi, inter, 1] and np.array_equal(c, G["q_out"][s, qi, inter, :, 1])
                c = G["q_in"][s, qi, inter, :4].copy()
                nz = o.xo_quant_2x2_dc(ptr(c, i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
                assert nz == G["q_nz"][s, qi, inter, 2] and np.array_equal(c, G["q_out"][s, qi, inter, :4, 2])
        for k, name in enumerate(("dequant_4x4", "dequant_4x4_dc")):
            c = G["q_lvl"][qi].copy()
            getattr(o, "xo_" + name)(ptr(c, i16p), ptr(dq, i32p), qp)
            assert np.array_equal(c, G["q_dequant"][qi, k]), f"{name} qp {qp}"
        dmf = int(dq[qp % 6][0]) << (qp // 6)
        for t in range(8):
            c = G["q_small"][qi, t].copy()
            assert o.xo_optimize_chroma_2x2_dc(ptr(c, i16p), dmf) == G["q_optdc_nz"][qi, t]
            assert np.array_equal(c, G["q_optdc"][qi, t])
    for l, (d15, d16, last) in zip(np.ascontiguousarray(G["q_dec_in"]), G["q_dec_out"]):
        assert (o.xo_decimate_score15(ptr(l, i16p)), o.xo_decimate_score16(ptr(l, i16p)),
                o.xo_coeff_last(ptr(l, i16p), 16)) == (d15, d16, last)


def test_oracle_mc_hpel_lowres(G):
    o = cc.oracle()
    stride = 96
    org = 20 * stride + 24
    planes = np.ascontiguousarray(G["mc_planes"])
    srcs = (rt.u8p * 4)(*[ptr(p[org:]) for p in planes])
    for (w, h, mvx, mvy), want in zip(G["mc_cases"], G["mc_luma_out"]):
        d = np.zeros(32 * 24, np.uint8)
        o.xo_mc_luma(ptr(d), C.c_ssize_t(32), srcs, C.c_ssize_t(stride), int(mvx), int(mvy), int(w), int(h))
        assert np.array_equal(d, want), f"mc_luma {w}x{h} mv {mvx},{mvy}"
    chroma = np.ascontiguousarray(G["mc_chroma_plane"])
    for (w, h, mvx, mvy), want in zip(G["mc_chroma_cases"], G["mc_chroma_out"]):
        u, v = np.zeros(32 * 8, np.uint8), np.zeros(32 * 8, np.uint8)
        o.xo_mc_chroma(ptr(u), ptr(v), C.c_ssize_t(32), ptr(chroma[org:]), C.c_ssize_t(stride), int(mvx), int(mvy), int(w), int(h))
        assert np.array_equal(u, want[0]) and np.array_equal(v, want[1]), "mc_chroma"
    src = np.ascontiguousarray(G["hpel_src"])
    outs = np.zeros((3, stride * 64), np.uint8)
    o8 = 8 * stride + 8
    o.xo_hpel_filter(ptr(outs[0][o8:]), ptr(outs[1][o8:]), ptr(outs[2][o8:]), ptr(src[o8:]), C.c_ssize_t(stride), 64, 40)
    assert np.array_equal(outs, G["hpel_out"])
    lo = np.zeros((4, 64 * 32), np.uint8)
    o.xo_lowres_core(ptr(src), *[ptr(x) for x in lo], C.c_ssize_t(stride), C.c_ssize_t(64), 40, 24)
    assert np.array_equal(lo, G["lowres_core_out"])


DB_NAMES = (("deblock_luma", 0), ("deblock_chroma", 0), ("deblock_luma_intra", 1), ("deblock_chroma_intra", 1))


def test_oracle_deblock_leaves(G):
    o = cc.oracle()
    dstride, dorg = 32, 4 * 32 + 8
    for p, par, outs in zip(G["db_in"], G["db_par"], G["db_out"]):
        alpha, beta = int(par[0]), int(par[1])
        tc0 = np.ascontiguousarray(par[2:6].astype(np.int8))
        k = 0
        for d in (0, 1):
            for name, intra in DB_NAMES:
                q = p.copy()
                if intra:
                    getattr(o, "xo_" + name)(ptr(q[dorg:]), C.c_ssize_t(dstride), d, alpha, beta)
                else:
                    getattr(o, "xo_" + name)(ptr(q[dorg:]), C.c_ssize_t(dstride), d, alpha, beta, ptr(tc0, i8p))
                assert np.array_equal(q, outs[k]), f"{name}[{d}]"
                k += 1
    n = len(G["bs_nnz"])
    bs = np.zeros((n, 2, 8, 4), np.uint8)
    o.xo_deblock_strength(n, ptr(np.ascontiguousarray(G["bs_nnz"])), ptr(np.ascontiguousarray(G["bs_ref"]), i8p),
                          ptr(np.ascontiguousarray(G["bs_mv"]), i16p), ptr(bs))
    assert np.array_equal(bs[:, :, :4], G["bs_out"][:, :, :4])      # the reference writes rows 0..3 of each direction


def oracle_slots(G, pkg, tag, filtered, lowres):
    o = cc.oracle()
    w, h, frames = frames_of(G, pkg, tag)
    g = cc.oracle_geom(w, h)
    slots = []
    for f in frames:
        s = np.zeros(g.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(g), ptr(f), ptr(s))
        if filtered:
            o.xo_frame_expand_border(C.byref(g), ptr(s))
            o.xo_frame_filter(C.byref(g), ptr(s))
        if lowres:
            o.xo_frame_init_lowres(C.byref(g), ptr(s))
        slots.append(s)
    return w, h, g, slots


def check_planes(G, tag, g, slot_f, slot_l, i):
    L, Lo = g.luma_plane_size, g.lowres_plane_size
    want = G[f"{tag}_planes_sha"][i]
    for k in range(4):
        assert sha(slot_f[k * L:(k + 1) * L]) == want[k], f"{tag} frame {i} luma plane {'NHVC'[k]}"
    assert sha(slot_f[g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size]) == want[4], "chroma"
    want = G[f"{tag}_lowres_sha"][i]
    for k in range(4):
        assert sha(slot_l[g.slot_lowres_off + k * Lo: g.slot_lowres_off + (k + 1) * Lo]) == want[k], f"lowres plane {k}"
    assert sha(slot_l[:L]) == want[4], "source plane after lowres init (side effect, mc.c:412-415)"


@pytest.mark.parametrize("tag", SIZES_WITH_FRAMES)
def test_oracle_frames_and_lookahead(G, pkg, tag):
    o = cc.oracle()
    w, h, g, slots_f = oracle_slots(G, pkg, tag, True, False)
    _, _, _, slots_l = oracle_slots(G, pkg, tag, False, True)
    geom = G[f"{tag}_geom"]
    assert (g.mb_w, g.mb_h, g.luma_stride, g.luma_w, g.luma_h, g.lowres_stride, g.lowres_w, g.lowres_h,
            g.luma_plane_size, g.luma_origin, g.chroma_origin) == tuple(int(x) for x in geom)
    for i in range(len(slots_f)):
        check_planes(G, tag, g, slots_f[i], slots_l[i], i)
    if f"{tag}_f0_luma4" in G.files:
        assert np.array_equal(slots_f[0][: 4 * g.luma_plane_size], G[f"{tag}_f0_luma4"])
    for i in range(len(slots_l)):
        mv, c, s = np.zeros((g.mb_count, 2), np.int16), np.zeros(g.mb_count, np.int32), np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(g), ptr(slots_l[i]), ptr(slots_l[i - 1]) if i else None, 1,
                                  ptr(mv, i16p), ptr(c, i32p), ptr(s, i32p), None)
        check_lookahead(G, tag, i, mv, c, s)


def check_lookahead(G, tag, i, mv, cost, sums):
    """sums: ours = {inter, intra, intra_mbs, ...}; golden = {i_cost_est[d][0], _aq, i_cost_est[0][0], i_intra_mbs[d], ..}"""
    want = G[f"{tag}_la_sums"][i]
    if i:
        assert np.array_equal(mv, G[f"{tag}_la_mv"][i]), f"{tag} frame {i}: lowres MVs"
        assert np.array_equal(cost, G[f"{tag}_la_cost"][i]), f"{tag} frame {i}: lowres MV costs"
        assert (sums[0], sums[1], sums[2]) == (want[0], want[2], want[3]), f"{tag} frame {i}: sums {sums[:3]} vs {want}"
    else:
        assert sums[1] == want[2], f"{tag} frame 0 intra cost"


ME_LABELS = ("hex5q", "dia2", "hex3q", "dia1q")


@pytest.mark.parametrize("tag", ("s", "cif"))
def test_oracle_me_search(G, pkg, tag):
    o = cc.oracle()
    w, h, g, slots = oracle_slots(G, pkg, tag, True, False)
    for label in ME_LABELS:
        me, subme, rng_, refine = (int(x) for x in G[f"{tag}_me_{label}_prm"])
        for k, (bl, rs) in enumerate(zip(G[f"{tag}_me_{label}_blocks"], G[f"{tag}_me_{label}_res"])):
            blocks = np.ascontiguousarray(bl).view(cc.ME_BLOCK_DTYPE).ravel()
            want = np.ascontiguousarray(rs).view(cc.ME_RESULT_DTYPE).ravel()
            got = np.zeros(len(blocks), cc.ME_RESULT_DTYPE)
            prm = cc.MeParams(me, subme, rng_, (26, 38)[k & 1], refine)
            o.xo_me_search_batch(C.byref(g), ptr(slots[1]), ptr(slots[0]), C.byref(prm), len(blocks),
                                 blocks.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, want), f"{tag} {label} list {k}"


def deblock_cases(G, tag):
    for qp in (30, 20, 44):
        k = f"{tag}_db{qp}"
        yield k, [np.ascontiguousarray(G[k + s]) for s in ("_type", "_part", "_cbp", "_bs")], [int(x) for x in G[k + "_par"]]


@pytest.mark.parametrize("tag", SIZES_WITH_FRAMES)
def test_oracle_deblock_frame(G, pkg, tag):
    o = cc.oracle()
    w, h, g, slots = oracle_slots(G, pkg, tag, False, False)
    for k, (t, p, cbp, bs), (qp, ao, bo) in deblock_cases(G, tag):
        s = slots[0].copy()
        o.xo_deblock_frame(C.byref(g), ptr(s), ptr(t, i8p), ptr(p), ptr(cbp, i16p), ptr(bs), qp, ao, bo)
        check_deblocked(G, k, g, s)


def check_deblocked(G, k, g, s):
    luma, chroma = s[: g.luma_plane_size], s[g.slot_chroma_off: g.slot_chroma_off + g.chroma_plane_size]
    if k + "_luma" in G.files:
        assert np.array_equal(luma, G[k + "_luma"]) and np.array_equal(chroma, G[k + "_chroma"]), k
    assert [sha(luma), sha(chroma)] == list(G[k + "_sha"]), k


def check_residual(G, qi, t, cbp, nz, lv, ry, rc):
    """compare one macroblock against x264_macroblock_encode's outputs.  The reference leaves the
    level arrays of blocks it did not code untouched, so levels are compared where nnz says coded."""
    assert cbp == G["res_cbp"][qi, t], f"cbp qp#{qi} mb {t}: {cbp:#x} vs {int(G['res_cbp'][qi, t]):#x}"
    assert np.array_equal(nz, G["res_nnz"][qi, t]), f"nnz qp#{qi} mb {t}"
    assert np.array_equal(ry, G["res_recon_y"][qi, t]), f"luma recon qp#{qi} mb {t}"
    assert np.array_equal(rc[:, :8], G["res_recon_c"][qi, t][:, :8]) and np.array_equal(rc[:, 16:24], G["res_recon_c"][qi, t][:, 16:24])
    want = G["res_levels"][qi, t]
    for i in range(16):
        if want[i * 16:(i + 1) * 16].any() or G["res_nnz"][qi, t][i]:
            assert np.array_equal(lv[i * 16:(i + 1) * 16], want[i * 16:(i + 1) * 16]), f"luma levels qp#{qi} mb {t} blk {i}"
    assert np.array_equal(lv[256:264], want[256:264]), f"chroma dc qp#{qi} mb {t}"
    for i in range(8):
        if want[264 + i * 16: 280 + i * 16].any():
            assert np.array_equal(lv[264 + i * 16: 280 + i * 16], want[264 + i * 16: 280 + i * 16]), f"chroma ac {i}"


def test_oracle_residual_mb(G):
    o = cc.oracle()
    o.xo_encode_inter_mb.restype = C.c_int
    for qi, qp in enumerate(G["res_qps"]):
        for t in range(G["res_fenc_y"].shape[1]):
            y = np.zeros((16, 32), np.uint8)
            c = np.zeros((8, 32), np.uint8)
            y[:, :16], c[:, :24] = G["res_pred_y"][qi, t], G["res_pred_c"][qi, t]
            lv, nz = np.zeros(392, np.int16), np.zeros(27, np.uint8)
            cbp = o.xo_encode_inter_mb(ptr(np.ascontiguousarray(G["res_fenc_y"][qi, t])),
                                       ptr(np.ascontiguousarray(G["res_fenc_c"][qi, t])), ptr(y), ptr(c), int(qp),
                                       ptr(lv, i16p), ptr(nz))
            check_residual(G, qi, t, cbp, nz, lv, y[:, :16], c[:, :24])


def test_reference_cli_bitstream(G, pkg):
    """config 1 anchor: the unmodified reference CLI on the CIF clip gives the recorded bitstream
    (runs wherever oracle/_ref/x264ref exists; the built binary travels to the GPU box)"""
    import subprocess
    import tempfile
    if not os.path.exists(cc.REF_CLI):
        pytest.skip("oracle/_ref/x264ref not built")
    clip = np.concatenate([pkg.synth_frame(352, 288, i, cut_frame=17) for i in range(30)])
    assert sha(clip) == G["cli_cif30"][0]
    with tempfile.TemporaryDirectory() as td:
        src, out = os.path.join(td, "syn_352x288.yuv"), os.path.join(td, "out.264")
        clip.tofile(src)
        subprocess.run([cc.REF_CLI, src, out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert sha(np.fromfile(out, np.uint8)) == G["cli_cif30"][1]


# =============================================================================================
# GPU: the CUDA path (C ABI) vs golden
# =============================================================================================

def gpu_slots(G, pkg, ctx, tag, filtered, lowres):
    import torch
    w, h, frames = frames_of(G, pkg, tag)
    g = pkg.geometry(w, h)
    n = len(frames)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    if filtered:
        ctx.frame_expand_border(g, slots, n)
        ctx.frame_filter(g, slots, n)
    if lowres:
        ctx.frame_init_lowres(g, slots, n)
        ctx.frame_export_lowres(g, slots, n)      # row-major lowres[0..3] for the plane hashes
    ctx.sync()
    return w, h, g, n, slots


@pytest.mark.gpu
@pytest.mark.parametrize("tag", SIZES_WITH_FRAMES)
def test_gpu_frames_and_lookahead(G, pkg, ctx, tag):
    import torch
    w, h, g, n, dev_f = gpu_slots(G, pkg, ctx, tag, True, False)
    _, _, _, _, dev_l = gpu_slots(G, pkg, ctx, tag, False, True)
    hf, hl = dev_f.cpu().numpy(), dev_l.cpu().numpy()
    for i in range(n):
        check_planes(G, tag, g, hf[i * g.slot_bytes:(i + 1) * g.slot_bytes], hl[i * g.slot_bytes:(i + 1) * g.slot_bytes], i)
    b = np.arange(n, dtype=np.int32)
    mvs = torch.zeros((n, g.mb_count, 2), dtype=torch.int16, device="cuda")
    costs = torch.zeros((n, g.mb_count), dtype=torch.int32, device="cuda")
    sums = torch.zeros((n, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.lookahead_frame_cost(g, dev_l, b, b - 1, np.ones(n, np.uint8), mvs, costs, sums)
    ctx.sync()
    mvs, costs, sums = mvs.cpu().numpy(), costs.cpu().numpy(), sums.cpu().numpy()
    for i in range(n):
        check_lookahead(G, tag, i, mvs[i], costs[i], sums[i])
    # the host-buffer entry point (the e2e path of bench.py) must agree as well
    luma = np.stack([f[: w * h] for f in frames_of(G, pkg, tag)[2]])
    m2, c2, s2 = ctx.lookahead_clip_host(w, h, luma)
    for i in range(n):
        check_lookahead(G, tag, i, m2[i], c2[i], s2[i])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("s", "cif"))
def test_gpu_me_search(G, pkg, ctx, tag):
    import torch
    w, h, g, n, slots = gpu_slots(G, pkg, ctx, tag, True, False)
    ref_slot, enc_slot = slots[: g.slot_bytes], slots[g.slot_bytes: 2 * g.slot_bytes]
    for label in ME_LABELS:
        me, subme, rng_, refine = (int(x) for x in G[f"{tag}_me_{label}_prm"])
        for k, (bl, rs) in enumerate(zip(G[f"{tag}_me_{label}_blocks"], G[f"{tag}_me_{label}_res"])):
            want = np.ascontiguousarray(rs).view(cc.ME_RESULT_DTYPE).ravel()
            nb = len(want)
            d_blocks = torch.from_numpy(np.ascontiguousarray(bl).ravel()).cuda()
            prm = pkg.MeParams(me, subme, rng_, (26, 38)[k & 1], refine)
            for sized in (False, True):
                d_res = torch.zeros(nb * cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                if sized:
                    ctx.me_search_sized(g, enc_slot, ref_slot, prm, k // 2, nb, d_blocks, d_res)
                else:
                    ctx.me_search_batch(g, enc_slot, ref_slot, prm, nb, d_blocks, d_res)
                ctx.sync()
                got = d_res.cpu().numpy().view(cc.ME_RESULT_DTYPE)
                assert np.array_equal(got, want), f"{tag} {label} list {k} sized={sized}"


@pytest.mark.gpu
@pytest.mark.parametrize("tag", SIZES_WITH_FRAMES)
def test_gpu_deblock_frame(G, pkg, ctx, tag):
    import torch
    w, h, g, n, slots = gpu_slots(G, pkg, ctx, tag, False, False)
    for k, arrs, (qp, ao, bo) in deblock_cases(G, tag):
        s = slots[: g.slot_bytes].clone()
        d = [torch.from_numpy(a).cuda() for a in arrs]
        torch.cuda.synchronize()
        ctx.deblock_frame(g, s, d[0], d[1], d[2], d[3], qp, ao, bo)
        ctx.sync()
        check_deblocked(G, k, g, s.cpu().numpy())


@pytest.mark.gpu
def test_gpu_residual_frame(G, pkg, ctx):
    """the 48 stored macroblocks of each QP laid out as one 128x96 frame (8x6 MBs)"""
    import torch
    g = pkg.geometry(128, 96)
    assert g.mb_count == G["res_fenc_y"].shape[1]
    cs = g.chroma_stride
    for qi, qp in enumerate(G["res_qps"]):
        fenc, pred = np.zeros(g.slot_bytes, np.uint8), np.zeros(g.slot_bytes, np.uint8)
        for t in range(g.mb_count):
            mx, my = t % g.mb_w, t // g.mb_w
            for slot, y, c, voff in ((fenc, G["res_fenc_y"][qi, t], G["res_fenc_c"][qi, t], 8),
                                     (pred, G["res_pred_y"][qi, t], G["res_pred_c"][qi, t], 16)):
                lo = g.luma_origin + my * 16 * g.luma_stride + mx * 16
                for r in range(16):
                    slot[lo + r * g.luma_stride: lo + r * g.luma_stride + 16] = y[r]
                co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
                for r in range(8):
                    slot[co + r * cs: co + r * cs + 16: 2] = c[r, :8]
                    slot[co + r * cs + 1: co + r * cs + 16: 2] = c[r, voff: voff + 8]
        d_fenc, d_pred = torch.from_numpy(fenc).cuda(), torch.from_numpy(pred).cuda()
        lv = torch.zeros((g.mb_count, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
        nz = torch.zeros((g.mb_count, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
        cbp = torch.zeros(g.mb_count, dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        ctx.residual_frame(g, d_fenc, d_pred, int(qp), lv, nz, cbp)
        ctx.sync()
        lv, nz, cbp, rec = lv.cpu().numpy(), nz.cpu().numpy(), cbp.cpu().numpy(), d_pred.cpu().numpy()
        for t in range(g.mb_count):
            mx, my = t % g.mb_w, t // g.mb_w
            lo = g.luma_origin + my * 16 * g.luma_stride + mx * 16
            ry = np.stack([rec[lo + r * g.luma_stride: lo + r * g.luma_stride + 16] for r in range(16)])
            co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
            rc = np.zeros((8, 24), np.uint8)
            for r in range(8):
                rc[r, :8] = rec[co + r * cs: co + r * cs + 16: 2]
                rc[r, 16:24] = rec[co + r * cs + 1: co + r * cs + 16: 2]
            check_residual(G, qi, t, int(cbp[t]), nz[t], lv[t], ry, rc)


@pytest.mark.gpu
def test_gpu_cost_batch(G, pkg, ctx):
    import torch
    a, b, s1, s2 = pixel_cases(G)
    cases = G["pix_cases"]
    n = len(cases)
    off1 = torch.from_numpy((cases[:, 1] * s1).astype(np.int64)).cuda()
    off2 = torch.from_numpy((cases[:, 2] * s2 + cases[:, 3]).astype(np.int64)).cuda()
    size = torch.from_numpy(cases[:, 0].astype(np.uint8)).cuda()
    pad = np.zeros(64, np.uint8)            # unaligned 8-pixel loads read whole words
    da, db = torch.from_numpy(np.concatenate([a, pad])).cuda(), torch.from_numpy(np.concatenate([b, pad])).cuda()
    for cmp in range(3):
        out = torch.zeros(n, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.cost_batch(cmp, n, da, off1, s1, db, off2, s2, size, out)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), cases[:, 4 + cmp].astype(np.int32)), ("sad", "ssd", "satd")[cmp]


@pytest.fixture(scope="module")
def tables(pkg):
    lib = pkg.lib()
    t = {}
    for name, cls, init in (("pixf", rt.PixelTable, "x264_pixel_init"), ("dctf", rt.DctTable, "x264_dct_init"),
                            ("zigzagf", rt.ZigzagTable, "x264_zigzag_init"), ("mcf", rt.McTable, "x264_mc_init"),
                            ("loopf", rt.DeblockTable, "x264_deblock_init")):
        t[name] = cls()
        getattr(lib, init)(0, C.byref(t[name]))
    t["quantf"] = rt.QuantTable()
    lib.x264_quant_init(None, 0, C.byref(t["quantf"]))
    return t


@pytest.mark.gpu
def test_gpu_tables_vs_golden(G, pkg, ctx, tables):
    """the drop-in function-pointer tables on the stored leaf vectors"""
    pix, dct, zz, mc, qf, lf = (tables[k] for k in ("pixf", "dctf", "zigzagf", "mcf", "quantf", "loopf"))
    a, b, s1, s2 = pixel_cases(G)
    for size, ya, yb, xb, sad, ssd, satd in G["pix_cases"][::4]:
        pa, pb = at(a, int(ya * s1)), at(b, int(yb * s2 + xb))
        assert [getattr(pix, nm)[int(size)](pa, s1, pb, s2) for nm in ("sad", "ssd", "satd")] == [sad, ssd, satd]
    for row in G["pix_x_cases"]:
        size, ya, offs, r4w, r3w = int(row[0]), int(row[1]), row[2:6], row[6:10], row[10:13]
        r4, r3 = (C.c_int * 4)(), (C.c_int * 3)()
        pix.sad_x4[size](at(a, ya * s1), *[at(b, int(k)) for k in offs], s2, r4)
        pix.satd_x3[size](at(a, ya * s1), *[at(b, int(k)) for k in offs[:3]], s2, r3)
        assert list(r4) == list(r4w) and list(r3) == list(r3w)
    for t in range(0, 30, 3):
        for k, nm in enumerate(("intra_satd_x3_8x8c", "intra_sad_x3_8x8c")):
            f = G["intra_fdec"][t].copy()
            r = (C.c_int * 3)()
            getattr(pix, nm)(ptr(np.ascontiguousarray(G["intra_fenc"][t])), at(f, 40), r)
            assert list(r) == list(G["intra_res"][t, k])
    fenc, fdec, coef = (np.ascontiguousarray(G[k]) for k in ("dct_fenc", "dct_fdec", "dct_coef"))
    for t in range(0, len(fenc), 3):
        for name, n in (("sub4x4_dct", 16), ("sub8x8_dct", 64), ("sub16x16_dct", 256), ("sub8x8_dct_dc", 4)):
            out = np.zeros(n, np.int16)
            getattr(dct, name)(ptr(out, i16p), ptr(fenc[t]), ptr(fdec[t]))
            assert np.array_equal(out, G["dct_" + name][t]), name
        for name, n in (("add4x4_idct", 16), ("add8x8_idct", 64), ("add16x16_idct", 256), ("add8x8_idct_dc", 4),
                        ("add16x16_idct_dc", 16)):
            d, c = fdec[t].copy(), coef[t, :n].copy()
            getattr(dct, name)(ptr(d), ptr(c, i16p))
            assert np.array_equal(d, G["dct_" + name][t]), name
        for name in ("dct4x4dc", "idct4x4dc"):
            c = coef[t, :16].copy()
            getattr(dct, name)(ptr(c, i16p))
            assert np.array_equal(c, G["dct_" + name][t]), name
        l = np.zeros(16, np.int16)
        zz.scan_4x4(ptr(l, i16p), ptr(coef[t], i16p))
        assert np.array_equal(l, G["dct_zigzag"][t])
    dq = np.ascontiguousarray(G["tab_dequant"])
    for qi, qp in enumerate(G["q_qps"]):
        qp = int(qp)
        mf, bias = np.ascontiguousarray(G["tab_quant_mf"][1, qp]), np.ascontiguousarray(G["tab_quant_bias"][1, qp])
        for s in range(5):
            c = G["q_in"][s, qi, 1].copy()
            nz = qf.quant_4x4(ptr(c, i16p), ptr(mf, u16p), ptr(bias, u16p))
            assert nz == G["q_nz"][s, qi, 1, 0] and np.array_equal(c, G["q_out"][s, qi, 1, :, 0])
        for k, name in enumerate(("dequant_4x4", "dequant_4x4_dc")):
            c = G["q_lvl"][qi].copy()
            getattr(qf, name)(ptr(c, i16p), ptr(dq, i32p), qp)
            assert np.array_equal(c, G["q_dequant"][qi, k]), f"{name} qp {qp}"
    stride, org = 96, 20 * 96 + 24
    planes = np.ascontiguousarray(G["mc_planes"])
    srcs = (rt.u8p * 4)(*[at(p, org) for p in planes])
    for (w, h, mvx, mvy), want in list(zip(G["mc_cases"], G["mc_luma_out"]))[::3]:
        d = np.zeros(32 * 24, np.uint8)
        mc.mc_luma(ptr(d), 32, srcs, stride, int(mvx), int(mvy), int(w), int(h), None)
        assert np.array_equal(d, want), f"mc_luma {w}x{h} mv {mvx},{mvy}"
    src = np.ascontiguousarray(G["hpel_src"])
    outs = np.zeros((3, stride * 64), np.uint8)
    buf = np.zeros(stride + 48, np.int16)
    o8 = 8 * stride + 8
    mc.hpel_filter(at(outs[0], o8), at(outs[1], o8), at(outs[2], o8), at(src, o8), stride, 64, 40, ptr(buf, i16p))
    assert np.array_equal(outs, G["hpel_out"])
    dstride, dorg = 32, 4 * 32 + 8
    for p, par, douts in list(zip(G["db_in"], G["db_par"], G["db_out"]))[::4]:
        alpha, beta = int(par[0]), int(par[1])
        tc0 = np.ascontiguousarray(par[2:6].astype(np.int8))
        k = 0
        for d in (0, 1):
            for name, intra in DB_NAMES:
                q = p.copy()
                if intra:
                    getattr(lf, name)[d](at(q, dorg), dstride, alpha, beta)
                else:
                    getattr(lf, name)[d](at(q, dorg), dstride, alpha, beta, ptr(tc0, i8p))
                assert np.array_equal(q, douts[k]), f"{name}[{d}]"
                k += 1


# ------------------------------------------------------------------ golden_r01_intra.npz: I16x16 macroblocks, predictors

@pytest.fixture(scope="module")
def GI():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r01_intra.npz")
    return np.load(path)


def check_intra16(GI, qi, t, cbp, nz, lv, dc, ry, rc):
    assert cbp == GI["i16_cbp"][qi, t], f"cbp qp#{qi} mb {t}: {cbp:#x} vs {int(GI['i16_cbp'][qi, t]):#x}"
    want_nz = GI["i16_nnz"][qi, t]
    assert np.array_equal(nz, want_nz), f"nnz qp#{qi} mb {t}"
    assert np.array_equal(ry, GI["i16_recon_y"][qi, t]), f"luma recon qp#{qi} mb {t}"
    assert np.array_equal(rc[:, :8], GI["i16_recon_c"][qi, t][:, :8]) and np.array_equal(rc[:, 16:24], GI["i16_recon_c"][qi, t][:, 16:24])
    want = GI["i16_levels"][qi, t]
    for i in range(16):
        if want_nz[i]:
            assert np.array_equal(lv[i * 16:(i + 1) * 16], want[i * 16:(i + 1) * 16]), f"luma levels qp#{qi} mb {t} blk {i}"
    if want_nz[24]:
        assert np.array_equal(dc, GI["i16_luma_dc"][qi, t]), f"luma dc levels qp#{qi} mb {t}"
    for ch in range(2):
        if want_nz[25 + ch]:
            assert np.array_equal(lv[256 + 4 * ch: 260 + 4 * ch], want[256 + 4 * ch: 260 + 4 * ch]), f"chroma dc qp#{qi} mb {t}"
    for i in range(8):
        if want_nz[16 + i]:
            assert np.array_equal(lv[264 + i * 16: 280 + i * 16], want[264 + i * 16: 280 + i * 16]), f"chroma ac {i}"


def test_oracle_intra16_mb(GI):
    o = cc.oracle()
    o.xo_encode_intra16_mb.restype = C.c_int
    coded_dc = 0
    for qi, qp in enumerate(GI["i16_qps"]):
        for t in range(GI["i16_fenc_y"].shape[1]):
            y = np.zeros((16, 32), np.uint8)
            c = np.zeros((8, 32), np.uint8)
            y[:, :16], c[:, :24] = GI["i16_pred_y"][qi, t], GI["i16_pred_c"][qi, t]
            lv, dc, nz = np.zeros(392, np.int16), np.zeros(16, np.int16), np.zeros(27, np.uint8)
            cbp = o.xo_encode_intra16_mb(ptr(np.ascontiguousarray(GI["i16_fenc_y"][qi, t])),
                                         ptr(np.ascontiguousarray(GI["i16_fenc_c"][qi, t])), ptr(y), ptr(c), int(qp),
                                         ptr(lv, i16p), ptr(dc, i16p), ptr(nz))
            check_intra16(GI, qi, t, cbp, nz, lv, dc, y[:, :16], c[:, :24])
            coded_dc += int(nz[24])
    assert coded_dc > 20


@pytest.mark.gpu
def test_gpu_intra16_frame(GI, pkg, ctx):
    """the 48 stored I16x16 macroblocks of each QP laid out as one 128x96 frame, x264dsp_residual_frames_typed_dev"""
    import torch
    g = pkg.geometry(128, 96)
    assert g.mb_count == GI["i16_fenc_y"].shape[1]
    cs = g.chroma_stride
    for qi, qp in enumerate(GI["i16_qps"]):
        fenc, pred = np.zeros(g.slot_bytes, np.uint8), np.zeros(g.slot_bytes, np.uint8)
        for t in range(g.mb_count):
            mx, my = t % g.mb_w, t // g.mb_w
            for slot, y, c, voff in ((fenc, GI["i16_fenc_y"][qi, t], GI["i16_fenc_c"][qi, t], 8),
                                     (pred, GI["i16_pred_y"][qi, t], GI["i16_pred_c"][qi, t], 16)):
                lo = g.luma_origin + my * 16 * g.luma_stride + mx * 16
                for r in range(16):
                    slot[lo + r * g.luma_stride: lo + r * g.luma_stride + 16] = y[r]
                co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
                for r in range(8):
                    slot[co + r * cs: co + r * cs + 16: 2] = c[r, :8]
                    slot[co + r * cs + 1: co + r * cs + 16: 2] = c[r, voff: voff + 8]
        d_fenc, d_pred = torch.from_numpy(fenc).cuda(), torch.from_numpy(pred).cuda()
        kind = torch.ones(g.mb_count, dtype=torch.uint8, device="cuda")
        lv = torch.zeros((g.mb_count, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
        dc = torch.zeros((g.mb_count, 16), dtype=torch.int16, device="cuda")
        nz = torch.zeros((g.mb_count, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
        cbp = torch.zeros(g.mb_count, dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        ctx.residual_frames_typed(g, d_fenc, d_pred, 1, int(qp), kind, lv, dc, nz, cbp)
        ctx.sync()
        lv, dc, nz, cbp, rec = lv.cpu().numpy(), dc.cpu().numpy(), nz.cpu().numpy(), cbp.cpu().numpy(), d_pred.cpu().numpy()
        for t in range(g.mb_count):
            mx, my = t % g.mb_w, t // g.mb_w
            lo = g.luma_origin + my * 16 * g.luma_stride + mx * 16
            ry = np.stack([rec[lo + r * g.luma_stride: lo + r * g.luma_stride + 16] for r in range(16)])
            co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
            rc = np.zeros((8, 24), np.uint8)
            for r in range(8):
                rc[r, :8] = rec[co + r * cs: co + r * cs + 16: 2]
                rc[r, 16:24] = rec[co + r * cs + 1: co + r * cs + 16: 2]
            check_intra16(GI, qi, t, int(cbp[t]), nz[t], lv[t], dc[t], ry, rc)


@pytest.mark.gpu
def test_gpu_predict_tables_golden(GI, pkg):
    """all 26 predictors of the library's x264_predict_*_init tables against the stored outputs of the reference's"""
    lib = pkg.lib()
    PRED_T = C.CFUNCTYPE(None, C.c_void_p)
    tabs = [(PRED_T * 7)(), (PRED_T * 7)(), (PRED_T * 12)()]
    lib.x264_predict_16x16_init(0, tabs[0])
    lib.x264_predict_8x8c_init(0, tabs[1])
    lib.x264_predict_4x4_init(0, tabs[2])
    for ti, (name, size) in enumerate((("p16", 16), ("p8c", 8), ("p4", 4))):
        want = GI["pred_" + name]
        for mode in range(len(tabs[ti])):
            for t in range(GI["pred_src"].shape[0]):
                b = GI["pred_src"][t].copy()
                tabs[ti][mode](C.cast(b.ctypes.data + 8 * 32 + 8, C.c_void_p))
                assert np.array_equal(b[8:8 + size, 8:8 + size], want[mode, t]), f"{name} mode {mode} case {t}"


def check_intra4(GI, qi, t, cbp, nz, lv, ry, rc):
    assert cbp == GI["i4_cbp"][qi, t], f"cbp qp#{qi} mb {t}: {cbp:#x} vs {int(GI['i4_cbp'][qi, t]):#x}"
    want_nz = GI["i4_nnz"][qi, t]
    assert np.array_equal(nz, want_nz), f"nnz qp#{qi} mb {t}"
    assert np.array_equal(ry, GI["i4_recon_y"][qi, t]), f"luma recon qp#{qi} mb {t} modes {GI['i4_modes'][qi, t]}"
    assert np.array_equal(rc[:, :8], GI["i4_recon_c"][qi, t][:, :8]) and np.array_equal(rc[:, 16:24], GI["i4_recon_c"][qi, t][:, 16:24])
    want = GI["i4_levels"][qi, t]
    for i in range(16):
        if want_nz[i]:
            assert np.array_equal(lv[i * 16:(i + 1) * 16], want[i * 16:(i + 1) * 16]), f"luma levels qp#{qi} mb {t} blk {i}"
    for i in range(8):
        if want_nz[16 + i]:
            assert np.array_equal(lv[264 + i * 16: 280 + i * 16], want[264 + i * 16: 280 + i * 16]), f"chroma ac {i}"


def test_oracle_intra4_mb(GI):
    o = cc.oracle()
    o.xo_encode_intra4_mb.restype = C.c_int
    for qi, qp in enumerate(GI["i16_qps"]):
        for t in range(GI["i4_fenc_y"].shape[1]):
            nb = GI["i4_nbh"][qi, t].copy()
            c = np.zeros((8, 32), np.uint8)
            c[:, :24] = GI["i4_pred_c"][qi, t]
            lv, nz = np.zeros(392, np.int16), np.zeros(27, np.uint8)
            cbp = o.xo_encode_intra4_mb(ptr(np.ascontiguousarray(GI["i4_fenc_y"][qi, t])), ptr(np.ascontiguousarray(GI["i4_fenc_c"][qi, t])),
                                        C.cast(nb.ctypes.data + 32 + 8, C.c_void_p), ptr(c), int(qp),
                                        ptr(np.ascontiguousarray(GI["i4_modes"][qi, t])), int(GI["i4_replicate5"][qi, t]),
                                        ptr(lv, i16p), ptr(nz))
            check_intra4(GI, qi, t, cbp, nz, lv, nb[1:, 8:24], c[:, :24])


@pytest.mark.gpu
def test_gpu_intra4_frame(GI, pkg, ctx):
    """the 48 stored I4x4 macroblocks of each QP on the even-even lattice of a 256x192 frame, each with its stored
    neighbourhood around it; the macroblocks in between are inter macroblocks with source == prediction (nothing
    coded, reconstruction == prediction), so the neighbourhoods are what the I4x4 kernel finds"""
    import torch
    g = pkg.geometry(256, 192)
    ls, cs = g.luma_stride, g.chroma_stride
    n = GI["i4_fenc_y"].shape[1]
    assert (g.mb_w // 2) * (g.mb_h // 2) == n
    for qi, qp in enumerate(GI["i16_qps"]):
        pred = np.zeros(g.slot_bytes, np.uint8)
        kind = np.zeros(g.mb_count, np.uint8)
        modes = np.zeros((g.mb_count, 16), np.uint8)
        where = []
        for t in range(n):
            mx, my = 2 * (t % (g.mb_w // 2)), 2 * (t // (g.mb_w // 2))
            where.append(my * g.mb_w + mx)
            lo = g.luma_origin + my * 16 * ls + mx * 16
            nb = GI["i4_nbh"][qi, t]
            pred[lo - ls - 1: lo - ls + 20] = nb[0, 7:28]
            for r in range(16):
                pred[lo + r * ls - 1] = nb[1 + r, 7]
            co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
            for r in range(8):
                pred[co + r * cs: co + r * cs + 16: 2] = GI["i4_pred_c"][qi, t][r, :8]
                pred[co + r * cs + 1: co + r * cs + 16: 2] = GI["i4_pred_c"][qi, t][r, 16:24]
            kind[where[-1]] = 2 + 4 * int(GI["i4_replicate5"][qi, t])
            modes[where[-1]] = GI["i4_modes"][qi, t]
        fenc = pred.copy()                                    # inter macroblocks: source == prediction
        for t in range(n):
            mx, my = 2 * (t % (g.mb_w // 2)), 2 * (t // (g.mb_w // 2))
            lo = g.luma_origin + my * 16 * ls + mx * 16
            for r in range(16):
                fenc[lo + r * ls: lo + r * ls + 16] = GI["i4_fenc_y"][qi, t][r]
            co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
            for r in range(8):
                fenc[co + r * cs: co + r * cs + 16: 2] = GI["i4_fenc_c"][qi, t][r, :8]
                fenc[co + r * cs + 1: co + r * cs + 16: 2] = GI["i4_fenc_c"][qi, t][r, 8:]
        d_fenc, d_pred = torch.from_numpy(fenc).cuda(), torch.from_numpy(pred).cuda()
        lv = torch.zeros((g.mb_count, pkg.RES_LEVELS_PER_MB), dtype=torch.int16, device="cuda")
        nz = torch.zeros((g.mb_count, pkg.RES_NNZ_PER_MB), dtype=torch.uint8, device="cuda")
        cbp = torch.zeros(g.mb_count, dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        ctx.residual_frames_typed(g, d_fenc, d_pred, 1, int(qp), torch.from_numpy(kind).cuda(), lv, None, nz, cbp,
                                  i4_modes=torch.from_numpy(modes).cuda())
        ctx.sync()
        lv, nz, cbp, rec = lv.cpu().numpy(), nz.cpu().numpy(), cbp.cpu().numpy(), d_pred.cpu().numpy()
        others = np.setdiff1d(np.arange(g.mb_count), where)
        assert not cbp[others].any(), "the filler macroblocks must not code anything"
        for t in range(n):
            mx, my = 2 * (t % (g.mb_w // 2)), 2 * (t // (g.mb_w // 2))
            lo = g.luma_origin + my * 16 * ls + mx * 16
            ry = np.stack([rec[lo + r * ls: lo + r * ls + 16] for r in range(16)])
            co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
            rc = np.zeros((8, 24), np.uint8)
            for r in range(8):
                rc[r, :8] = rec[co + r * cs: co + r * cs + 16: 2]
                rc[r, 16:24] = rec[co + r * cs + 1: co + r * cs + 16: 2]
            check_intra4(GI, qi, t, int(cbp[where[t]]), nz[where[t]], lv[where[t]], ry, rc)


def test_oracle_probe_pskip(GI):
    o = cc.oracle()
    o.xo_probe_pskip_mb.restype = C.c_int
    for qi, qp in enumerate(GI["pskip_qps"]):
        for t in range(GI["pskip_fenc_y"].shape[1]):
            y = np.zeros((16, 32), np.uint8)
            c = np.zeros((8, 32), np.uint8)
            y[:, :16], c[:, :24] = GI["pskip_pred_y"][qi, t], GI["pskip_pred_c"][qi, t]
            r = o.xo_probe_pskip_mb(ptr(np.ascontiguousarray(GI["pskip_fenc_y"][qi, t])), ptr(np.ascontiguousarray(GI["pskip_fenc_c"][qi, t])),
                                    ptr(y), ptr(c), int(qp))
            assert r == GI["pskip_skip"][qi, t], f"qp {qp} mb {t}: oracle {r} reference {GI['pskip_skip'][qi, t]}"


@pytest.mark.gpu
def test_gpu_probe_pskip_golden(GI, pkg, ctx):
    """the 96 stored macroblocks of each QP as two 128x96 frames, x264dsp_probe_pskip_frames_dev against the
    reference's stored decisions"""
    import torch
    g = pkg.geometry(128, 96)
    cs = g.chroma_stride
    nper = g.mb_count
    for qi, qp in enumerate(GI["pskip_qps"]):
        nfr = GI["pskip_fenc_y"].shape[1] // nper
        fenc, pred = np.zeros((nfr, g.slot_bytes), np.uint8), np.zeros((nfr, g.slot_bytes), np.uint8)
        for t in range(nfr * nper):
            f, m = divmod(t, nper)
            mx, my = m % g.mb_w, m // g.mb_w
            for slot, y, c, voff in ((fenc[f], GI["pskip_fenc_y"][qi, t], GI["pskip_fenc_c"][qi, t], 8),
                                     (pred[f], GI["pskip_pred_y"][qi, t], GI["pskip_pred_c"][qi, t], 16)):
                lo = g.luma_origin + my * 16 * g.luma_stride + mx * 16
                for r in range(16):
                    slot[lo + r * g.luma_stride: lo + r * g.luma_stride + 16] = y[r]
                co = g.slot_chroma_off + g.chroma_origin + my * 8 * cs + mx * 16
                for r in range(8):
                    slot[co + r * cs: co + r * cs + 16: 2] = c[r, :8]
                    slot[co + r * cs + 1: co + r * cs + 16: 2] = c[r, voff: voff + 8]
        skip = torch.full((nfr, nper), 9, dtype=torch.uint8, device="cuda")
        d_fenc, d_pred = torch.from_numpy(fenc.reshape(-1)).cuda(), torch.from_numpy(pred.reshape(-1)).cuda()
        torch.cuda.synchronize()
        ctx.probe_pskip_frames(g, d_fenc, d_pred, nfr, int(qp), skip)
        ctx.sync()
        assert np.array_equal(skip.cpu().numpy().reshape(-1), GI["pskip_skip"][qi][: nfr * nper]), f"qp {qp}"
