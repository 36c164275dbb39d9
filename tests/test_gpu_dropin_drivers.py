"""Driver-level drop-in proof: the UNMODIFIED reference encoder (oracle/_ref) encodes a clip with three
of its hot-path drivers served by libx264dsp_b200.so --

  x264_frame_init_lowres                        -> x264dsp_frame_init_lowres_dev
  x264_frame_filter + _expand_border_filtered   -> x264dsp_frame_filter_dev
  x264_frame_deblock_row + x264_frame_expand_border (second test: the whole in-loop filter)
                                                -> x264dsp_deblock_frame_dev + x264dsp_frame_expand_border_dev
  x264_slicetype_frame_cost (its per-frame cache, filled before x264_slicetype_decide runs)
                                                -> x264dsp_lookahead_frame_cost_dev
  x264_me_search_ref (every partition search of the main encode, last cases)
                                                -> x264dsp_me_search_batch_dev on frames kept resident on the device
  x264_mb_mc (motion compensation of every P macroblock, all partitions, last case) -> x264dsp_mc_frames_part_dev
  x264_macroblock_probe_pskip (the P_SKIP test of the P-slice analysis, last case)
                                                -> x264dsp_mc_frame_dev + x264dsp_probe_pskip_frames_dev
  x264_macroblock_encode (every inter macroblock of the P slices and every I16x16 / I4x4 macroblock of the I slices, last
  two cases: DCT, quant, zig-zag, dequant, decimation, luma / chroma DC, IDCT; levels / nnz / cbp handed to the
  reference's CABAC writer)                     -> x264dsp_residual_frames_typed_dev

through the doors of glue/x264dsp_doors.c (the glue INTEGRATION.md describes; glue/x264dsp_glue.c is the same device side in C), and must emit the
byte-identical bitstream.  Every plane the main encode searches in (half-pel planes of every
reconstructed frame), every lowres MV used as an MV candidate and every frame cost that drives scenecut
and rate control then comes from the CUDA path."""
import ctypes as C

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr

pytestmark = pytest.mark.gpu

FRAME_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
COST_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
FDEC_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_int, C.c_int, C.c_int)
ME_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p)
MBMC_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
PSKIP_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                       C.c_void_p, C.POINTER(C.c_int))
MBENC_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                       C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int))


@pytest.mark.parametrize("w,h,n,cut,me,subme,psub,inloop,mehook,mbenc", [
    (352, 288, 12, 7, 1, 5, 0, False, False, False), (208, 160, 8, -1, 0, 2, 0, False, False, False),
    (352, 288, 12, 7, 1, 5, 0, True, False, False), (208, 160, 8, 4, 1, 3, 0, True, False, False),
    (352, 288, 10, 6, 1, 5, 0, True, True, False), (208, 160, 8, 4, 0, 2, 1, True, True, False),
    (208, 160, 6, -1, 1, 4, 1, True, True, False),
    (208, 160, 6, 3, 1, 2, 0, True, False, True), (176, 144, 5, -1, 1, 5, 1, True, True, True)])
def test_encoder_bitstream_identical_with_gpu_drivers(pkg, ctx, w, h, n, cut, me, subme, psub, inloop, mehook, mbenc):
    import torch
    lib = cc.ref()
    assert lib is not None, "oracle/_ref/libx264ref.so must travel to the GPU box (make -C oracle ref)"
    lib.xref_frame_ptr.restype = C.c_void_p
    lib.xref_frame_ptr.argtypes = [C.c_void_p, C.c_int]
    g = pkg.geometry(w, h)
    clip = np.concatenate([pkg.synth_frame(w, h, i, cut_frame=cut) for i in range(n)])
    lps, wps = g.luma_plane_size, g.lowres_plane_size
    slots = torch.zeros(2 * g.slot_bytes, dtype=torch.uint8, device="cuda")
    stage = np.zeros(2 * g.slot_bytes, np.uint8)

    def host_view(addr, nbytes):
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def upload(slot_index, offset, addr, nbytes):
        stage[:nbytes] = host_view(addr, nbytes)
        slots[slot_index * g.slot_bytes + offset: slot_index * g.slot_bytes + offset + nbytes].copy_(
            torch.from_numpy(stage[:nbytes]))

    def download(slot_index, offset, addr, nbytes):
        host_view(addr, nbytes)[:] = slots[slot_index * g.slot_bytes + offset:
                                           slot_index * g.slot_bytes + offset + nbytes].cpu().numpy()

    resident = {}                         # x264_frame_t* -> device slot of the frame (for the ME hook)

    def keep_resident(frame):
        if not mehook:
            return
        if frame not in resident and len(resident) >= 8:
            resident.pop(next(iter(resident)))
        resident[frame] = slots[: g.slot_bytes].clone()

    @FRAME_CB
    def lowres_cb(hv, frame):
        # the source frame's padded luma plane in, four padded lowres planes (and the source plane with its
        # duplicated last column / row, mc.c:412-415) out
        upload(0, 0, lib.xref_frame_ptr(frame, 10), lps)
        if mehook:
            upload(0, g.slot_chroma_off, lib.xref_frame_ptr(frame, 11), g.chroma_plane_size)   # the P_SKIP probe reads it
        keep_resident(frame)              # the source samples as the main encode's searches will see them
        ctx.frame_init_lowres(g, slots, 1)
        ctx.frame_export_lowres(g, slots, 1)      # the reference wants its row-major lowres[0..3]
        ctx.sync()
        download(0, 0, lib.xref_frame_ptr(frame, 10), lps)
        download(0, g.slot_lowres_off, lib.xref_frame_ptr(frame, 12), 4 * wps)

    @FRAME_CB
    def filter_cb(hv, frame):
        # the reconstructed, deblocked, border-expanded plane N in; planes H, V, HV with their borders out
        upload(0, 0, lib.xref_frame_ptr(frame, 10), lps)
        ctx.frame_filter(g, slots, 1)
        ctx.sync()
        download(0, lps, lib.xref_frame_ptr(frame, 10) + lps, 3 * lps)

    @COST_CB
    def cost_cb(hv, p0, b, want_intra, mvs, costs, sums):
        upload(0, g.slot_lowres_off, lib.xref_frame_ptr(p0, 12), 4 * wps)
        upload(1, g.slot_lowres_off, lib.xref_frame_ptr(b, 12), 4 * wps)
        ctx.frame_retile_lowres(g, slots, 2)      # the planes were written by the caller, not by init_lowres
        d_mvs = torch.zeros((1, g.mb_count, 2), dtype=torch.int16, device="cuda")
        d_costs = torch.zeros((1, g.mb_count), dtype=torch.int32, device="cuda")
        d_sums = torch.zeros((1, pkg.LA_SUMS), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.lookahead_frame_cost(g, slots, [1], [0], [want_intra], d_mvs, d_costs, d_sums)
        ctx.sync()
        host_view(mvs, g.mb_count * 4)[:] = d_mvs.cpu().numpy().view(np.uint8).ravel()
        host_view(costs, g.mb_count * 4)[:] = d_costs.cpu().numpy().view(np.uint8).ravel()
        host_view(sums, 32)[:] = d_sums.cpu().numpy().view(np.uint8).ravel()[:32]

    deblocked = [0]

    @FDEC_CB
    def fdec_cb(hv, frame, do_deblock, mb_type, partition, cbp, bs, qp, aoff, boff):
        # the reconstructed frame in; deblocked planes with borders and the three half-pel planes out
        cps = g.chroma_plane_size
        upload(0, 0, lib.xref_frame_ptr(frame, 10), lps)
        upload(0, g.slot_chroma_off, lib.xref_frame_ptr(frame, 11), cps)
        if do_deblock:
            nmb = g.mb_count
            d = [torch.from_numpy(host_view(p_, nb).copy()).cuda() for p_, nb in
                 ((mb_type, nmb), (partition, nmb), (cbp, 2 * nmb), (bs, 64 * nmb))]
            torch.cuda.synchronize()
            ctx.deblock_frame(g, slots, d[0], d[1], d[2], d[3], qp, aoff, boff)
            deblocked[0] += 1
        ctx.frame_expand_border(g, slots, 1)
        ctx.frame_filter(g, slots, 1)
        ctx.sync()
        download(0, 0, lib.xref_frame_ptr(frame, 10), 4 * lps)
        download(0, g.slot_chroma_off, lib.xref_frame_ptr(frame, 11), cps)
        keep_resident(frame)              # this reconstructed frame is the next frame's reference

    me_calls = [0]

    @ME_CB
    def me_cb(hv, fenc, fref, blk, me_method, subme_, me_range, qp, out):
        if fenc not in resident or fref not in resident:
            return 1                      # declined: the reference's own code runs
        d_blk = torch.from_numpy(host_view(blk, cc.ME_BLOCK_DTYPE.itemsize).copy()).cuda()
        d_res = torch.zeros(cc.ME_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.me_search_batch(g, resident[fenc], resident[fref], pkg.MeParams(me_method, subme_, me_range, qp, 0), 1, d_blk, d_res)
        ctx.sync()
        host_view(out, cc.ME_RESULT_DTYPE.itemsize)[:] = d_res.cpu().numpy()
        me_calls[0] += 1
        return 0

    # one macroblock as a 16x16 "frame": source and prediction slots in the library's own plane layout
    g1 = pkg.geometry(16, 16)
    mb_slots = torch.zeros(2 * g1.slot_bytes, dtype=torch.uint8, device="cuda")
    mb_stage = np.zeros((2, g1.slot_bytes), np.uint8)
    d_lv = torch.zeros(pkg.RES_LEVELS_PER_MB, dtype=torch.int16, device="cuda")
    d_nz = torch.zeros(pkg.RES_NNZ_PER_MB, dtype=torch.uint8, device="cuda")
    d_cbp = torch.zeros(1, dtype=torch.int16, device="cuda")
    d_dc = torch.zeros(16, dtype=torch.int16, device="cuda")
    d_kind = torch.zeros(1, dtype=torch.uint8, device="cuda")
    d_modes = torch.zeros(16, dtype=torch.uint8, device="cuda")
    mbenc_calls = [0, 0, 0]               # inter, I16x16, I4x4

    def mb_planes(buf):
        luma = buf[g1.luma_origin:][: 16 * g1.luma_stride].reshape(16, g1.luma_stride)[:, :16]
        co = g1.slot_chroma_off + g1.chroma_origin
        chroma = buf[co:][: 8 * g1.chroma_stride].reshape(8, g1.chroma_stride)[:, :16]
        return luma, chroma

    @MBENC_CB
    def mbenc_cb(hv, fenc_y, fenc_c, fdec_y, fdec_c, qp, kind, i4_modes, levels, luma_dc, nnz, cbp):
        fy = host_view(fenc_y, 16 * 16).reshape(16, 16)
        fc = host_view(fenc_c, 8 * 16).reshape(8, 16)              # U at +0, V at +8
        dy = host_view(fdec_y, 16 * 32).reshape(16, 32)
        dc = host_view(fdec_c, 8 * 32).reshape(8, 32)              # U at +0, V at +16
        for k, (y_, u_, v_) in enumerate(((fy, fc[:, :8], fc[:, 8:16]), (dy[:, :16], dc[:, :8], dc[:, 16:24]))):
            luma, chroma = mb_planes(mb_stage[k])
            luma[:] = y_
            chroma[:, 0::2] = u_
            chroma[:, 1::2] = v_
        if kind & 2:
            # the reconstructed neighbourhood fdec_buf holds around the macroblock goes into the slot's padding:
            # the row above from column -1 to 19 and the column to the left
            nbh = host_view(fdec_y - 33, 17 * 32 + 1)
            lo = g1.luma_origin
            mb_stage[1][lo - g1.luma_stride - 1: lo - g1.luma_stride + 20] = nbh[:21]
            for r in range(16):
                mb_stage[1][lo + r * g1.luma_stride - 1] = nbh[33 + r * 32 - 1]
            d_modes.copy_(torch.from_numpy(host_view(i4_modes, 16).copy()))
        mb_slots.copy_(torch.from_numpy(mb_stage.reshape(-1)))
        d_kind.fill_(kind)
        torch.cuda.synchronize()
        ctx.residual_frames_typed(g1, mb_slots[: g1.slot_bytes], mb_slots[g1.slot_bytes:], 1, qp, d_kind, d_lv, d_dc,
                                  d_nz, d_cbp, i4_modes=d_modes)
        ctx.sync()
        rec = mb_slots[g1.slot_bytes:].cpu().numpy()
        luma, chroma = mb_planes(rec)
        dy[:, :16] = luma
        dc[:, :8] = chroma[:, 0::2]
        dc[:, 16:24] = chroma[:, 1::2]
        host_view(levels, 2 * pkg.RES_LEVELS_PER_MB)[:] = d_lv.cpu().numpy().view(np.uint8)
        host_view(nnz, pkg.RES_NNZ_PER_MB)[:] = d_nz.cpu().numpy()
        host_view(luma_dc, 32)[:] = d_dc.cpu().numpy().view(np.uint8)
        cbp[0] = int(d_cbp.cpu().numpy()[0])
        mbenc_calls[kind & 3] += 1
        return 0

    pskip_calls = [0, 0]                  # probes served, of which skippable
    d_pmv = torch.zeros((g.mb_count, 2), dtype=torch.int16, device="cuda")
    d_ppred = torch.zeros(g.slot_bytes, dtype=torch.uint8, device="cuda")
    d_pskip = torch.zeros(g.mb_count, dtype=torch.uint8, device="cuda")

    @PSKIP_CB
    def pskip_cb(hv, fenc, fref, mb_x, mb_y, mvx, mvy, qp, fdec_y, fdec_c, skip):
        if fenc not in resident or fref not in resident:
            return 1
        xy = mb_y * g.mb_w + mb_x
        d_pmv.zero_()
        d_pmv[xy, 0], d_pmv[xy, 1] = mvx, mvy
        torch.cuda.synchronize()
        ctx.mc_frame(g, resident[fref], d_pmv, d_ppred)                       # mc_luma + mc_chroma at the pskip mv
        ctx.probe_pskip_frames(g, resident[fenc], d_ppred, 1, qp, d_pskip)
        ctx.sync()
        pred = d_ppred.cpu().numpy()
        lo = g.luma_origin + mb_y * 16 * g.luma_stride + mb_x * 16
        co = g.slot_chroma_off + g.chroma_origin + mb_y * 8 * g.chroma_stride + mb_x * 16
        dy = host_view(fdec_y, 16 * 32).reshape(16, 32)
        dc = host_view(fdec_c, 8 * 32).reshape(8, 32)
        for r in range(16):
            dy[r, :16] = pred[lo + r * g.luma_stride: lo + r * g.luma_stride + 16]
        for r in range(8):
            row = pred[co + r * g.chroma_stride: co + r * g.chroma_stride + 16]
            dc[r, :8], dc[r, 16:24] = row[0::2], row[1::2]
        skip[0] = int(d_pskip[xy].item())
        pskip_calls[0] += 1
        pskip_calls[1] += skip[0]
        return 0

    mbmc_calls = [0]
    d_mv4 = torch.zeros((g.mb_count, 4, 2), dtype=torch.int16, device="cuda")

    @MBMC_CB
    def mbmc_cb(hv, fref, mb_x, mb_y, mv8x8, fdec_y, fdec_c):
        if fref not in resident:
            return 1
        xy = mb_y * g.mb_w + mb_x
        d_mv4.zero_()
        d_mv4[xy] = torch.from_numpy(np.frombuffer(host_view(mv8x8, 16).tobytes(), np.int16).reshape(4, 2).copy()).cuda()
        torch.cuda.synchronize()
        ctx.mc_frames_part(g, resident[fref], 1, d_mv4, d_ppred)
        ctx.sync()
        pred = d_ppred.cpu().numpy()
        lo = g.luma_origin + mb_y * 16 * g.luma_stride + mb_x * 16
        co = g.slot_chroma_off + g.chroma_origin + mb_y * 8 * g.chroma_stride + mb_x * 16
        dy = host_view(fdec_y, 16 * 32).reshape(16, 32)
        dc = host_view(fdec_c, 8 * 32).reshape(8, 32)
        for r in range(16):
            dy[r, :16] = pred[lo + r * g.luma_stride: lo + r * g.luma_stride + 16]
        for r in range(8):
            row = pred[co + r * g.chroma_stride: co + r * g.chroma_stride + 16]
            dc[r, :8], dc[r, 16:24] = row[0::2], row[1::2]
        mbmc_calls[0] += 1
        return 0

    outs, calls = [], (C.c_int * 3)()
    for use_gpu in (False, True):
        enc = cc.RefEncoder(w, h, me=me, subme=subme, me_range=16, qp=26, psub16x16=psub)
        if use_gpu:
            lib.xref_set_driver_hooks(lowres_cb, filter_cb, cost_cb)
            if inloop:
                lib.xref_set_fdec_hook(fdec_cb)
            if mehook:
                lib.xref_set_me_hook(me_cb)
            if mbenc:
                lib.xref_set_mbenc_hook(mbenc_cb)
            if mbenc and mehook:
                lib.xref_set_pskip_hook(pskip_cb)
                lib.xref_set_mbmc_hook(mbmc_cb)
        else:
            lib.xref_set_driver_hooks(None, None, None)
        out = np.zeros(1 << 20, np.uint8)
        launches0 = ctx.launches
        doors = (C.c_int * 12)()
        lib.xref_door_stats_reset()
        try:
            size = lib.xref_encode_clip(enc.h, ptr(clip), n, ptr(out), out.size)
        finally:
            lib.xref_driver_hook_calls(calls)
            lib.xref_door_stats_read(doors)
            lib.xref_set_driver_hooks(None, None, None)
            lib.xref_set_fdec_hook(None)
            lib.xref_set_me_hook(None)
            lib.xref_set_mbenc_hook(None)
            lib.xref_set_pskip_hook(None)
            lib.xref_set_mbmc_hook(None)
        assert size > 0, size
        outs.append(out[:size].copy())
        if use_gpu:
            # exact counts: every frame's lowres init, every reconstructed frame's filter pass (all n frames are
            # reference frames: no B frames), one lookahead cost per frame after the first (the reference's slicetype
            # decision analyses each new frame once against its predecessor)
            assert calls[0] == n, f"x264_frame_init_lowres hooked {calls[0]} times for {n} frames"
            assert calls[1] == n, f"x264_frame_filter / in-loop filter hooked {calls[1]} times for {n} frames"
            print(f"DOORCOUNTS {w}x{h} n={n} cut={cut} me={me} subme={subme} psub={psub}: calls={list(calls)} doors={list(doors)}")
            assert calls[2] == n - 1, f"lookahead cost hooked {calls[2]} times for {n} frames"
            assert ctx.launches - launches0 >= 2 * n, "the encode must have gone through the CUDA kernels"
            # per door {entered, eligible, served}: every eligible call must have been served by the device (a decline
            # falls back to the reference's code silently inside x264dsp_doors.c -- that must not happen), and the callbacks'
            # own counters must agree with the doors'
            st = {name: tuple(doors[3 * i: 3 * i + 3]) for i, name in enumerate(("me", "mbenc", "pskip", "mbmc"))}
            if mehook:
                assert st["me"][1] > 0 and st["me"][2] == st["me"][1] == me_calls[0], f"ME door {st['me']} vs {me_calls[0]}"
                # with the cost door installed the lookahead's own searches run on the device as well, so every call of
                # x264_me_search_ref that the encoder makes is a main-encode search: entered == eligible == served
                assert st["me"][0] == st["me"][1], f"ME door {st['me']}: calls the door did not consider eligible"
            if mbenc and mehook:
                assert st["mbmc"][2] == st["mbmc"][1] == mbmc_calls[0] and mbmc_calls[0] > 0, f"mb_mc door {st['mbmc']}"
                assert st["pskip"][2] == st["pskip"][1] == pskip_calls[0] and pskip_calls[0] > 0, f"pskip door {st['pskip']}"
                assert 0 < pskip_calls[1] < pskip_calls[0], f"one-sided probes: {pskip_calls}"
            if mbenc:
                assert st["mbenc"][2] == st["mbenc"][1] == sum(mbenc_calls), f"macroblock_encode door {st['mbenc']} vs {mbenc_calls}"
                assert mbenc_calls[0] >= g.mb_count, f"only {mbenc_calls[0]} inter macroblocks were coded on the device"
                assert mbenc_calls[1] > 0, "no I16x16 macroblock of the I frames was coded on the device"
                assert mbenc_calls[2] > 0, "no I4x4 macroblock of the I frames was coded on the device"
            if inloop:
                assert deblocked[0] >= n - 1, f"deblocking ran on the device for {deblocked[0]} of {n} frames"
    assert outs[0].size == outs[1].size and np.array_equal(outs[0], outs[1]), \
        f"bitstreams differ: {outs[0].size} vs {outs[1].size} bytes"
