"""GPU parity: frame staging / borders / half-pel / lowres planes, block costs and the lowres
lookahead, CUDA (through the C ABI) against the CPU oracle on the same seeded input.
Bar: bit-exact, padding included."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc

pytestmark = pytest.mark.gpu

# sizes: CIF; not a multiple of 16; plane size without the 128-byte gap; stride = w+80; 1080p
SIZES = [(352, 288), (200, 120), (368, 304), (960, 540)]


def _torch():
    import torch
    return torch


def oracle_slots(g, frames, border=False, hpel=False, lowres=False):
    o = cc.oracle()
    out = []
    for f in frames:
        s = np.zeros(g.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(g), cc.ptr(f), cc.ptr(s))
        if border:
            o.xo_frame_expand_border(C.byref(g), cc.ptr(s))
        if hpel:
            o.xo_frame_filter(C.byref(g), cc.ptr(s))
        if lowres:
            o.xo_frame_init_lowres(C.byref(g), cc.ptr(s))
        out.append(s)
    return out


def diff_report(a, b, g):
    bad = np.nonzero(a != b)[0]
    if len(bad) == 0:
        return ""
    regions = []
    for off in bad[:8]:
        if off < 4 * g.luma_plane_size:
            k, r = divmod(int(off), g.luma_plane_size)
            regions.append(f"luma{k} row {r // g.luma_stride - 32} col {r % g.luma_stride - 32}")
        elif off < g.slot_lowres_off:
            r = int(off) - g.slot_chroma_off
            regions.append(f"chroma row {r // g.chroma_stride - 16} col {r % g.chroma_stride - 32}")
        elif off < g.slot_tiled_off:
            k, r = divmod(int(off) - g.slot_lowres_off, g.lowres_plane_size)
            regions.append(f"lowres{k} row {r // g.lowres_stride - 32} col {r % g.lowres_stride - 32}")
        else:
            k, r = divmod(int(off) - g.slot_tiled_off, g.tiled_plane_size)
            regions.append(f"tiled{k} tile {r // 64} byte {r % 64}")
    return f"{len(bad)} bytes differ, first: " + "; ".join(regions)


# the benchmark geometries, plus widths whose right-edge remainder (n_units % 30, n_units = W16/8 + 1) lands in each of the
# hpel kernel's narrow tail strips: 1080p / 4K -> 1 unit (4 lanes), 256 -> 3 units (8 lanes), 320 -> 11 units (16 lanes),
# 352 / 368 -> 15 / 17 units (32 lanes), 200 -> a single partial strip, 720 -> 1 unit after three full strips
HPEL_SIZES = SIZES + [(1920, 1080), (3840, 2160), (256, 64), (320, 64), (720, 96), (1280, 720)]


@pytest.mark.parametrize("w,h", HPEL_SIZES)
def test_frame_planes_match_oracle(pkg, ctx, w, h):
    torch = _torch()
    n = 2
    frames = [pkg.synth_frame(w, h, i) for i in range(n)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    assert bytes(g) == bytes(go), "geometry differs from the oracle"

    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.sync()
    want = oracle_slots(go, frames)
    got = slots.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert diff_report(got[i], want[i], g) == "", "load_i420"

    ctx.frame_expand_border(g, slots, n)
    ctx.frame_filter(g, slots, n)
    ctx.sync()
    want = oracle_slots(go, frames, border=True, hpel=True)
    got = slots.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert diff_report(got[i], want[i], g) == "", "expand_border + hpel filter"


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (352, 288), (720, 96), (256, 64)])
def test_hpel_tma_variant_matches_oracle(pkg, ctx, w, h, monkeypatch):
    """xd_hpel_tma_kernel (the full strips' source rows through cp.async.bulk.tensor: a measured variant, off by default,
    X264DSP_HPEL_TMA=1) produces the same planes; widths below one 29-unit strip fall back to the cp.async kernel"""
    torch = _torch()
    monkeypatch.setenv("X264DSP_HPEL_TMA", "1")
    n = 3
    frames = [pkg.synth_frame(w, h, i) for i in range(n)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    ctx.frame_load_i420(g, torch.from_numpy(np.concatenate(frames)).cuda(), slots, n)
    ctx.frame_expand_border(g, slots, n)
    ctx.frame_filter(g, slots, n)
    ctx.sync()
    want = oracle_slots(go, frames, border=True, hpel=True)
    got = slots.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert diff_report(got[i], want[i], g) == "", "expand_border + hpel filter (TMA variant)"


@pytest.mark.parametrize("w,h", SIZES + [(1920, 1080)])
def test_lowres_planes_match_oracle(pkg, ctx, w, h):
    torch = _torch()
    n = 2
    frames = [pkg.synth_frame(w, h, i) for i in range(n)]
    g = pkg.geometry(w, h)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    ctx.frame_export_lowres(g, slots, n)      # the slot keeps the planes tiled; the row-major form is on request
    ctx.sync()
    want = oracle_slots(cc.oracle_geom(w, h), frames, lowres=True)
    got = slots.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert diff_report(got[i], want[i], g) == "", "init_lowres (incl. source-plane side effect, tiled copies)"


@pytest.mark.parametrize("cmp", [0, 1, 2])
def test_cost_batch_matches_oracle(pkg, ctx, cmp):
    torch = _torch()
    rng = np.random.RandomState(7 + cmp)
    stride1, stride2, rows = 80, 112, 64
    p1 = rng.randint(0, 256, stride1 * rows + 64).astype(np.uint8)
    p2 = rng.randint(0, 256, stride2 * rows + 64).astype(np.uint8)
    # adversarial rows: all 0 vs all 255, equal planes, checkerboard
    p1[: stride1 * 16] = 0
    p2[: stride2 * 16] = 255
    p1[stride1 * 16: stride1 * 32] = (np.arange(stride1 * 16) & 1) * 255
    n = 4000
    size = rng.randint(0, 8, n).astype(np.uint8)
    off1 = np.zeros(n, np.int64)
    off2 = np.zeros(n, np.int64)
    for i in range(n):
        bw, bh = cc.BLOCK_W[size[i]], cc.BLOCK_H[size[i]]
        off1[i] = rng.randint(0, rows - bh) * stride1 + rng.randint(0, stride1 - bw)
        off2[i] = rng.randint(0, rows - bh) * stride2 + rng.randint(0, stride2 - bw)
    want = np.zeros(n, np.int32)
    cc.oracle().xo_cost_batch(cmp, n, cc.ptr(p1), cc.ptr(off1, cc.i64p), stride1, cc.ptr(p2),
                              cc.ptr(off2, cc.i64p), stride2, cc.ptr(size), cc.ptr(want, cc.i32p))
    d = [torch.from_numpy(a).cuda() for a in (p1, off1, p2, off2, size)]
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.cost_batch(cmp, n, d[0], d[1], stride1, d[2], d[3], stride2, d[4], out)
    ctx.sync()
    got = out.cpu().numpy()
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"cmp {cmp}: {len(bad)} costs differ, e.g. block {bad[:5]} size {size[bad[:5]]}"


def oracle_lookahead(g, slots, i, want_intra=1, rows=False):
    o = cc.oracle()
    mv = np.zeros((g.mb_count, 2), np.int16)
    c = np.zeros(g.mb_count, np.int32)
    s = np.zeros(8, np.int32)
    r = np.zeros((2, g.mb_h), np.int32)
    o.xo_lookahead_frame_cost(C.byref(g), cc.ptr(slots[i]), cc.ptr(slots[i - 1]) if i else None, want_intra,
                              cc.ptr(mv, cc.i16p), cc.ptr(c, cc.i32p), cc.ptr(s, cc.i32p), cc.ptr(r, cc.i32p))
    return mv, c, s, r


@pytest.mark.parametrize("kernel", [1, 2, 3])     # block rows per warp: 1, 4, 8
@pytest.mark.parametrize("w,h,n,cut", [(352, 288, 5, 3), (200, 120, 3, -1), (64, 64, 3, -1), (1920, 1080, 3, -1),
                                       (96, 64, 2, -1), (80, 112, 4, 2), (64, 144, 3, -1), (3840, 2160, 2, -1)])
def test_lookahead_matches_oracle(pkg, ctx, w, h, n, cut, kernel):
    torch = _torch()
    ctx.lookahead_select_kernel(kernel)
    frames = [pkg.synth_frame(w, h, i, cut) for i in range(n)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    ref_slots = oracle_slots(go, frames, lowres=True)

    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    mvs = torch.full((n, g.mb_count, 2), -7, dtype=torch.int16, device="cuda")
    costs = torch.full((n, g.mb_count), -7, dtype=torch.int32, device="cuda")
    sums = torch.full((n, pkg.LA_SUMS), -7, dtype=torch.int32, device="cuda")
    rows = torch.full((n, 2, g.mb_h), -7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    b = np.arange(n)
    p0 = b - 1
    for rep in range(2):            # second run exercises the epoch logic of the sync words
        ctx.lookahead_frame_cost(g, slots, b, p0, np.ones(n, np.uint8), mvs, costs, sums, rows)
        ctx.sync()
        gm, gc, gs, gr = (t.cpu().numpy() for t in (mvs, costs, sums, rows))
        for i in range(n):
            mv_o, c_o, s_o, r_o = oracle_lookahead(go, ref_slots, i, rows=True)
            assert np.array_equal(gm[i], mv_o), f"rep {rep} frame {i}: mvs"
            assert np.array_equal(gc[i], c_o), f"rep {rep} frame {i}: block costs"
            assert np.array_equal(gs[i][:5], s_o[:5]), f"rep {rep} frame {i}: sums {gs[i]} vs {s_o}"
            assert np.array_equal(gr[i], r_o), f"rep {rep} frame {i}: row sums"
    ctx.lookahead_select_kernel(0)


@pytest.mark.parametrize("w,h", [(352, 288), (200, 120), (1920, 1080), (66, 70)])
def test_fused_staging_and_lowres_equals_the_two_calls(pkg, ctx, w, h):
    """x264dsp_frame_load_luma_lowres_dev leaves the slot exactly as load_luma + init_lowres do (incl.
    non-mod-16 sizes, where the staging has to replicate the last column / row)"""
    torch = _torch()
    n = 3
    luma = np.stack([pkg.synth_frame(w, h, i, luma_only=True) for i in range(n)])
    g = pkg.geometry(w, h)
    d_luma = torch.from_numpy(luma).cuda()
    a = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    b = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_luma(g, d_luma, a, n)
    ctx.frame_init_lowres(g, a, n)
    ctx.frame_load_luma_lowres(g, d_luma, b, n)
    ctx.sync()
    diff = torch.nonzero(a != b)
    assert diff.numel() == 0, f"{diff.numel()} bytes differ, first at slot offset {int(diff[0]) % g.slot_bytes}"
    # the lowres-only variant writes the same (tiled) lowres planes and nothing else
    c = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_lowres_from_luma(g, d_luma, c, n)
    ctx.sync()
    av, cv = a.view(n, -1), c.view(n, -1)
    assert torch.equal(av[:, g.slot_tiled_off:], cv[:, g.slot_tiled_off:]), "tiled lowres planes differ"
    assert int(cv[:, : g.slot_tiled_off].max()) == 0, "the lowres-only entry point must not touch the other planes"


def test_lowres_export_and_import_are_inverse(pkg, ctx):
    """row-major <-> tiled: export of the tiled planes equals the oracle's planes (above); importing those
    planes back rebuilds the same tiles"""
    torch = _torch()
    w, h, n = 200, 120, 2
    frames = [pkg.synth_frame(w, h, i) for i in range(n)]
    g = pkg.geometry(w, h)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    ctx.frame_export_lowres(g, slots, n)
    ctx.sync()
    before = slots.clone()
    view = slots.view(n, -1)
    view[:, g.slot_tiled_off: g.slot_tiled_off + 4 * g.tiled_plane_size] = 0
    ctx.frame_retile_lowres(g, slots, n)
    ctx.sync()
    assert torch.equal(slots, before)


def test_lookahead_batch_of_clips_both_kernels(pkg, ctx):
    """a launch large enough for the automatic choice to take the quad-row kernel (>= 48 pairs): many
    short clips, both mappings and the automatic one must agree with the oracle and with each other"""
    torch = _torch()
    w, h, clips, clip_len = 208, 160, 20, 4
    n = clips * clip_len
    frames = [pkg.synth_frame(w, h, 3 * (i // clip_len) + i % clip_len, 2 if (i // clip_len) % 5 == 0 else -1) for i in range(n)]
    g = pkg.geometry(w, h)
    go = cc.oracle_geom(w, h)
    ref_slots = oracle_slots(go, frames, lowres=True)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    b = np.arange(n)
    p0 = np.where(b % clip_len == 0, -1, b - 1)
    want = []
    for i in range(n):
        o = cc.oracle()
        mv, c, s = np.zeros((go.mb_count, 2), np.int16), np.zeros(go.mb_count, np.int32), np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(go), cc.ptr(ref_slots[i]), cc.ptr(ref_slots[i - 1]) if p0[i] >= 0 else None, 1,
                                  cc.ptr(mv, cc.i16p), cc.ptr(c, cc.i32p), cc.ptr(s, cc.i32p), None)
        want.append((mv, c, s))
    for kernel in (3, 2, 1, 0):
        ctx.lookahead_select_kernel(kernel)
        mvs = torch.full((n, g.mb_count, 2), -7, dtype=torch.int16, device="cuda")
        costs = torch.full((n, g.mb_count), -7, dtype=torch.int32, device="cuda")
        sums = torch.full((n, pkg.LA_SUMS), -7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.lookahead_frame_cost(g, slots, b, p0, np.ones(n, np.uint8), mvs, costs, sums)
        ctx.sync()
        gm, gc, gs = mvs.cpu().numpy(), costs.cpu().numpy(), sums.cpu().numpy()
        for i in range(n):
            if p0[i] >= 0:
                assert np.array_equal(gm[i], want[i][0]), f"kernel {kernel} frame {i}: mvs"
                assert np.array_equal(gc[i], want[i][1]), f"kernel {kernel} frame {i}: costs"
            assert np.array_equal(gs[i][:5], want[i][2][:5]), f"kernel {kernel} frame {i}: sums"
    ctx.lookahead_select_kernel(0)


def test_lookahead_without_intra_and_host_api(pkg, ctx):
    """second analysis of a frame (intra already known) and the host-buffer entry point"""
    torch = _torch()
    w, h, n = 352, 288, 4
    luma = np.stack([pkg.synth_frame(w, h, i, luma_only=True) for i in range(n)])
    mvs, costs, sums = ctx.lookahead_clip_host(w, h, luma)
    go = cc.oracle_geom(w, h)
    frames = [np.concatenate([luma[i], np.zeros(w * h // 2, np.uint8)]) for i in range(n)]
    ref_slots = oracle_slots(go, frames, lowres=True)
    for i in range(n):
        mv_o, c_o, s_o, _ = oracle_lookahead(go, ref_slots, i)
        assert np.array_equal(mvs[i], mv_o) and np.array_equal(costs[i], c_o)
        assert np.array_equal(sums[i][:5], s_o[:5])

    g = pkg.geometry(w, h)
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    d_mvs = torch.zeros((1, g.mb_count, 2), dtype=torch.int16, device="cuda")
    d_costs = torch.zeros((1, g.mb_count), dtype=torch.int32, device="cuda")
    d_sums = torch.zeros((1, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    ctx.lookahead_frame_cost(g, slots, [2], [1], [0], d_mvs, d_costs, d_sums)
    ctx.sync()
    mv_o, c_o, s_o, _ = oracle_lookahead(go, ref_slots, 2, want_intra=0)
    assert np.array_equal(d_mvs.cpu().numpy()[0], mv_o)
    assert np.array_equal(d_costs.cpu().numpy()[0], c_o)
    assert np.array_equal(d_sums.cpu().numpy()[0][:5], s_o[:5])


def test_wavefront_calls_on_two_streams_share_the_context(pkg, ctx):
    """two x264dsp_lookahead_frame_cost_dev calls in flight on DIFFERENT streams of one context: the second queues
    behind the first (they share the context's sync words and ticket) and both results equal the single-call ones"""
    torch = _torch()
    w, h, clips, clip_len = 208, 160, 14, 4
    n = clips * clip_len
    frames = [pkg.synth_frame(w, h, 5 * (i // clip_len) + i % clip_len) for i in range(n)]
    g = pkg.geometry(w, h)
    i420 = torch.from_numpy(np.concatenate(frames)).cuda()
    slots = torch.zeros(n * g.slot_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.frame_load_i420(g, i420, slots, n)
    ctx.frame_init_lowres(g, slots, n)
    ctx.sync()
    b = np.arange(n)
    p0 = np.where(b % clip_len == 0, -1, b - 1)
    half = n // 2

    def outputs(k):
        return (torch.full((k, g.mb_count, 2), -7, dtype=torch.int16, device="cuda"),
                torch.full((k, g.mb_count), -7, dtype=torch.int32, device="cuda"),
                torch.full((k, pkg.LA_SUMS), -7, dtype=torch.int32, device="cuda"))

    want = []
    for lo, hi in ((0, half), (half, n)):
        o = outputs(hi - lo)
        ctx.lookahead_frame_cost(g, slots, b[lo:hi], p0[lo:hi], np.ones(hi - lo, np.uint8), *o)
        ctx.sync()
        want.append([t.cpu().numpy() for t in o])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):
        o1, o2 = outputs(half), outputs(n - half)
        torch.cuda.synchronize()
        ctx.lookahead_frame_cost(g, slots, b[:half], p0[:half], np.ones(half, np.uint8), *o1, stream=s1.cuda_stream)
        ctx.lookahead_frame_cost(g, slots, b[half:], p0[half:], np.ones(n - half, np.uint8), *o2, stream=s2.cuda_stream)
        torch.cuda.synchronize()
        for got, exp in ((o1, want[0]), (o2, want[1])):
            for t, e in zip(got, exp):
                assert np.array_equal(t.cpu().numpy()[..., :5] if t.shape[-1] == pkg.LA_SUMS else t.cpu().numpy(),
                                      e[..., :5] if e.shape[-1] == pkg.LA_SUMS else e), f"rep {rep}"
