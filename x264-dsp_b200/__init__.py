"""x264dsp_b200 -- ctypes binding of libx264dsp_b200.so (the B200-native x264-dsp hot path).

The directory name carries a hyphen, so load it with `load_package()` from `__graft_entry__.py`
(or importlib) under the module name `x264dsp_b200`.

There is no CPU implementation behind these calls.  If the shared library has not been built the
import fails; if no CUDA device can be opened, `Context()` raises.  PyTorch is used only to own
device memory and streams in tests and benchmarks.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libx264dsp_b200.so")

PIXEL_16x16, PIXEL_16x8, PIXEL_8x16, PIXEL_8x8, PIXEL_8x4, PIXEL_4x8, PIXEL_4x4, PIXEL_4x16 = range(8)
BLOCK_W = [16, 16, 8, 8, 8, 4, 4, 4]
BLOCK_H = [16, 8, 16, 8, 4, 8, 4, 16]
CMP_SAD, CMP_SSD, CMP_SATD = 0, 1, 2
ME_DIA, ME_HEX, ME_UMH, ME_ESA, ME_TESA = range(5)
ME_MODE_SEARCH, ME_MODE_REFDUPE, ME_MODE_REFINE_QPEL = range(3)
LA_COST_INTER, LA_COST_INTRA, LA_INTRA_MBS, LA_SAD_EVALS, LA_SATD_EVALS, LA_SUMS = 0, 1, 2, 3, 4, 8
(PROF_LOAD, PROF_LOWRES, PROF_LA_INTRA, PROF_LA_INTER, PROF_HPEL, PROF_BORDER, PROF_COST, PROF_ME, PROF_MC,
 PROF_RESIDUAL, PROF_DEBLOCK, PROF_LA_TILE) = range(12)
RES_LEVELS_PER_MB = 16 * 16 + 2 * 4 + 2 * 4 * 16
RES_NNZ_PER_MB = 16 + 8 + 3

u8p = C.POINTER(C.c_uint8)


class Geom(C.Structure):
    """x264dsp_geom_t"""
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "mb_w", "mb_h", "mb_count", "luma_w", "luma_h",
        "luma_stride", "luma_plane_size", "luma_origin",
        "chroma_stride", "chroma_h", "chroma_plane_size", "chroma_origin",
        "lowres_w", "lowres_h", "lowres_stride", "lowres_plane_size", "lowres_origin",
        "slot_chroma_off", "slot_lowres_off")] + [("slot_bytes", C.c_int64)] + [(n, C.c_int32) for n in (
        "tile_w", "tile_h", "tiled_plane_size", "slot_tiled_off")]


class MeParams(C.Structure):
    _fields_ = [("me_method", C.c_int32), ("subpel_refine", C.c_int32), ("me_range", C.c_int32),
                ("qp", C.c_int32), ("refine_qpel", C.c_int32)]


class PFrameParams(C.Structure):
    """x264dsp_pframe_params_t"""
    _fields_ = [("me_method", C.c_int32), ("subpel_refine", C.c_int32), ("me_range", C.c_int32), ("qp", C.c_int32),
                ("mv_range", C.c_int32), ("fast_pskip", C.c_int32), ("mvc_scale", C.c_int32), ("analyse_inter", C.c_int32)]


MB_P_L0, MB_P_8x8, MB_P_SKIP = 4, 5, 6

ME_BLOCK_DTYPE = np.dtype([("i_pixel", "<i4"), ("bx", "<i4"), ("by", "<i4"), ("mvp", "<i2", (2,)),
                           ("i_mvc", "<i4"), ("mvc", "<i2", (16, 2)),
                           ("mv_min_fpel", "<i4", (2,)), ("mv_max_fpel", "<i4", (2,)),
                           ("mv_min_spel", "<i4", (2,)), ("mv_max_spel", "<i4", (2,))], align=True)
ME_RESULT_DTYPE = np.dtype([("mv", "<i2", (2,)), ("cost", "<i4"), ("cost_mv", "<i4")], align=True)


class X264DspError(RuntimeError):
    pass


def build(force=False):
    """compile the CUDA library in-tree (nvcc, sm_100a)"""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-s", "-C", ROOT], check=True)
    return LIB_PATH


_lib = None


def lib():
    """the loaded shared library; raises when it does not exist (no fallback)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise X264DspError(
                f"{LIB_PATH} is missing: run `make` (nvcc, sm_100a). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        l.x264dsp_version.restype = C.c_char_p
        l.x264dsp_stream.restype = C.c_void_p
        l.x264dsp_launch_count.restype = C.c_int64
        l.x264dsp_tables_context.restype = C.c_void_p
        l.x264dsp_stream.argtypes = [C.c_void_p]
        l.x264dsp_launch_count.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        kind = {-1: "bad argument", -2: "no CUDA device (there is no CPU path)", -3: "out of memory"}.get(
            rc, f"cudaError {rc}")
        raise X264DspError(f"{what} failed: {kind}")


def geometry(width, height):
    g = Geom()
    check(lib().x264dsp_geometry(int(width), int(height), C.byref(g)), "x264dsp_geometry")
    return g


def synth_frame(width, height, n, cut_frame=-1, luma_only=False):
    """planar I420 (or luma-only) numpy array of synthetic frame n"""
    y = np.empty(width * height, np.uint8)
    if luma_only:
        check(lib().x264dsp_synth_frame(width, height, n, cut_frame, y.ctypes.data_as(u8p), None, None),
              "x264dsp_synth_frame")
        return y
    cw, ch = width // 2, height // 2
    buf = np.empty(width * height + 2 * cw * ch, np.uint8)
    yv = buf[: width * height]
    uv = buf[width * height: width * height + cw * ch]
    vv = buf[width * height + cw * ch:]
    check(lib().x264dsp_synth_frame(width, height, n, cut_frame, yv.ctypes.data_as(u8p),
                                    uv.ctypes.data_as(u8p), vv.ctypes.data_as(u8p)), "x264dsp_synth_frame")
    return buf


def tiling_blocks(g, i_pixel, mb_mv):
    """x264dsp_me_block_t list of one block per bw x bh tile of the frame:
    mvp = 2 x the lowres MV of the co-located macroblock (common/mvpred.c:172-184), mvc = {mvp, 0},
    MV limits as encoder/analyse.c:378-393 sets them (fpel border 6, mv_range 512)"""
    bw, bh = BLOCK_W[i_pixel], BLOCK_H[i_pixel]
    xs, ys = np.meshgrid(np.arange(0, g.luma_w, bw), np.arange(0, g.luma_h, bh))
    blocks = np.zeros(xs.size, ME_BLOCK_DTYPE)
    blocks["i_pixel"] = i_pixel
    blocks["bx"] = xs.ravel()
    blocks["by"] = ys.ravel()
    mbx, mby = blocks["bx"] // 16, blocks["by"] // 16
    fmv = 512 << 2
    for k, (mb, nmb) in enumerate(((mbx, g.mb_w), (mby, g.mb_h))):
        smin = np.clip((-(mb << 4) - 24) << 2, -fmv, fmv - 1)
        smax = np.clip((((nmb - mb - 1) << 4) + 24) << 2, -fmv, fmv - 1)
        blocks["mv_min_spel"][:, k], blocks["mv_max_spel"][:, k] = smin, smax
        blocks["mv_min_fpel"][:, k], blocks["mv_max_fpel"][:, k] = (smin >> 2) + 6, (smax >> 2) - 6
    blocks["mvp"] = mb_mv[mby * g.mb_w + mbx] * 2
    blocks["i_mvc"] = 2
    blocks["mvc"][:, 0] = blocks["mvp"]
    blocks["mvc"][:, 1] = 0
    return blocks


def frame_range(n_frames, rank, world):
    """(first, count, need_prev): the frames rank `rank` of `world` owns (x264dsp_frame_range)"""
    first, count, prev = C.c_int(), C.c_int(), C.c_int()
    check(lib().x264dsp_frame_range(int(n_frames), int(rank), int(world), C.byref(first), C.byref(count),
                                    C.byref(prev)), "x264dsp_frame_range")
    return first.value, count.value, bool(prev.value)


class GopEncodeParams(C.Structure):
    """x264dsp_gop_encode_params_t"""
    _fields_ = [(n, C.c_int32) for n in ("me_method", "subpel_refine", "me_range", "qp_i", "qp_p", "mv_range", "fast_pskip",
                                         "analyse_inter", "deblock", "alpha_c0_offset", "beta_offset")]


class GopParams(C.Structure):
    """x264dsp_gop_params_t"""
    _fields_ = [("keyint_max", C.c_int32), ("keyint_min", C.c_int32), ("scenecut_threshold", C.c_int32)]


TYPE_IDR, TYPE_I, TYPE_P = 1, 2, 3


def slicetype_decide(icost, pcost, keyint_max, keyint_min, scenecut_threshold):
    """frame types of a whole sequence from the lookahead's frame costs (x264dsp_slicetype_decide)"""
    ic = np.ascontiguousarray(icost, np.int32)
    pc = np.ascontiguousarray(pcost, np.int32)
    types = np.zeros(ic.size, np.uint8)
    prm = GopParams(keyint_max, keyint_min, scenecut_threshold)
    check(lib().x264dsp_slicetype_decide(int(ic.size), _hp(ic, C.c_int32), _hp(pc, C.c_int32), C.byref(prm), _hp(types)),
          "x264dsp_slicetype_decide")
    return types


def gop_ranges(types):
    """[(first frame, frame count)] of every GOP (x264dsp_gop_ranges)"""
    t = np.ascontiguousarray(types, np.uint8)
    first, count = np.zeros(max(t.size, 1), np.int32), np.zeros(max(t.size, 1), np.int32)
    n = C.c_int(0)
    check(lib().x264dsp_gop_ranges(int(t.size), _hp(t), _hp(first, C.c_int32), _hp(count, C.c_int32), C.byref(n)),
          "x264dsp_gop_ranges")
    return [(int(first[i]), int(count[i])) for i in range(n.value)]


def gop_shard(gops, rank, world):
    """the GOPs (a slice of `gops`) rank `rank` of `world` encodes (x264dsp_gop_shard)"""
    count = np.array([c for _, c in gops], np.int32)
    first, n = C.c_int(0), C.c_int(0)
    check(lib().x264dsp_gop_shard(len(gops), _hp(count, C.c_int32) if len(gops) else None, int(rank), int(world),
                                  C.byref(first), C.byref(n)), "x264dsp_gop_shard")
    return gops[first.value: first.value + n.value]


def lookahead_sharded(analyse, luma, rank, world, gather=None, mb_count=None):
    """Frame-range sharded lookahead pass of ONE sequence.

    `luma` [n, h*w] is the whole sequence (every rank sees the same host input; only its own range
    plus the one overlap frame is uploaded).  `analyse(frames)` is the single-GPU pass -- in the
    product `lambda f: ctx.lookahead_clip_host(w, h, f)` -- returning (mvs, costs, sums) with the
    first frame analysed intra-only and every other frame against its predecessor.
    Returns this rank's (first, mvs, costs, sums) for exactly the frames it owns; frame 0 of the
    sequence keeps its intra-only result.  If `gather` is given (a callable doing an all-gather of
    a numpy array along axis 0, e.g. over torch.distributed) every rank gets the full-sequence
    arrays instead -- the only collective of the path, and an optional one.
    `mb_count` (lowres blocks per frame) shapes the empty result of a rank that owns no frame (world > n); with a `gather`
    it must be given so that every rank contributes arrays of the same trailing shape.
    """
    n = luma.shape[0]
    first, count, need_prev = frame_range(n, rank, world)
    lo = first - 1 if need_prev else first
    if count == 0:
        if mb_count is None and gather is not None:
            raise ValueError("lookahead_sharded: a rank without frames needs mb_count to shape what it gathers")
        mvs = np.zeros((0, mb_count or 0, 2), np.int16)
        costs = np.zeros((0, mb_count or 0), np.int32)
        sums = np.zeros((0, LA_SUMS), np.int32)
    else:
        mvs, costs, sums = analyse(luma[lo:first + count])
        if need_prev:               # drop the overlap frame's own (intra-only) result
            mvs, costs, sums = mvs[1:], costs[1:], sums[1:]
    if gather is None:
        return first, mvs, costs, sums
    return 0, gather(mvs), gather(costs), gather(sums)


def cost_mv_table(qp):
    t = np.zeros(8193, np.uint16)
    check(lib().x264dsp_cost_mv_table(qp, t.ctypes.data_as(C.POINTER(C.c_uint16))), "x264dsp_cost_mv_table")
    return t


def quant_tables(b_inter, qp):
    mf = np.zeros(16, np.uint16)
    bias = np.zeros(16, np.uint16)
    check(lib().x264dsp_quant_tables(int(b_inter), qp, mf.ctypes.data_as(C.POINTER(C.c_uint16)),
                                     bias.ctypes.data_as(C.POINTER(C.c_uint16))), "x264dsp_quant_tables")
    return mf, bias


def dequant_table():
    t = np.zeros((6, 16), np.int32)
    check(lib().x264dsp_dequant_table(t.ctypes.data_as(C.POINTER(C.c_int))), "x264dsp_dequant_table")
    return t


def _dp(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def _hp(a, ctype=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(ctype))


class Context:
    """x264dsp_ctx_t: one per process / GPU"""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().x264dsp_create(int(device), C.byref(self._h)), "x264dsp_create")
        self.device = device

    def close(self):
        if self._h:
            for p in getattr(self, "_pinned", []):
                lib().x264dsp_host_free(self._h, p)
            self._pinned = []
            lib().x264dsp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        """cudaStream_t of the context as an int"""
        return lib().x264dsp_stream(self._h)

    def torch_stream(self):
        import torch
        return torch.cuda.ExternalStream(self.stream, device=f"cuda:{self.device}")

    def sync(self):
        check(lib().x264dsp_sync(self._h), "x264dsp_sync")

    @property
    def launches(self):
        return lib().x264dsp_launch_count(self._h)

    # ---- frame staging -------------------------------------------------------------------
    def frame_load_i420(self, g, i420_dev, slots_dev, n_frames):
        check(lib().x264dsp_frame_load_i420_dev(self._h, C.byref(g), _dp(i420_dev), _dp(slots_dev),
                                                int(n_frames), None), "x264dsp_frame_load_i420_dev")

    def frame_load_luma(self, g, luma_dev, slots_dev, n_frames):
        check(lib().x264dsp_frame_load_luma_dev(self._h, C.byref(g), _dp(luma_dev), _dp(slots_dev),
                                                int(n_frames), None), "x264dsp_frame_load_luma_dev")

    def frame_load_luma_lowres(self, g, luma_dev, slots_dev, n_frames):
        """frame_load_luma + frame_init_lowres fused (one pass over the picture)"""
        check(lib().x264dsp_frame_load_luma_lowres_dev(self._h, C.byref(g), _dp(luma_dev), _dp(slots_dev),
                                                       int(n_frames), None), "x264dsp_frame_load_luma_lowres_dev")

    def frame_lowres_from_luma(self, g, luma_dev, slots_dev, n_frames):
        """x264_frame_init_lowres reading the pictures directly; the slots' luma planes are not written"""
        check(lib().x264dsp_frame_lowres_from_luma_dev(self._h, C.byref(g), _dp(luma_dev), _dp(slots_dev),
                                                       int(n_frames), None), "x264dsp_frame_lowres_from_luma_dev")

    def frame_export_lowres(self, g, slots_dev, n_frames):
        """tiled lowres planes -> the reference's row-major lowres[0..3] in the slot's lowres region"""
        check(lib().x264dsp_frame_export_lowres_dev(self._h, C.byref(g), _dp(slots_dev), int(n_frames), None),
              "x264dsp_frame_export_lowres_dev")

    def frame_retile_lowres(self, g, slots_dev, n_frames):
        """rebuild the tiled lowres copies of slots whose row-major lowres planes were written by the caller"""
        check(lib().x264dsp_frame_retile_lowres_dev(self._h, C.byref(g), _dp(slots_dev), int(n_frames), None),
              "x264dsp_frame_retile_lowres_dev")

    def frame_expand_border(self, g, slots_dev, n_frames):
        check(lib().x264dsp_frame_expand_border_dev(self._h, C.byref(g), _dp(slots_dev), int(n_frames), None),
              "x264dsp_frame_expand_border_dev")

    def frame_filter(self, g, slots_dev, n_frames):
        check(lib().x264dsp_frame_filter_dev(self._h, C.byref(g), _dp(slots_dev), int(n_frames), None),
              "x264dsp_frame_filter_dev")

    def frame_init_lowres(self, g, slots_dev, n_frames):
        check(lib().x264dsp_frame_init_lowres_dev(self._h, C.byref(g), _dp(slots_dev), int(n_frames), None),
              "x264dsp_frame_init_lowres_dev")

    # ---- block costs ---------------------------------------------------------------------
    def cost_batch(self, cmp, n, pix1, off1, stride1, pix2, off2, stride2, size, out):
        check(lib().x264dsp_cost_batch_dev(self._h, int(cmp), int(n), _dp(pix1), _dp(off1), int(stride1),
                                           _dp(pix2), _dp(off2), int(stride2), _dp(size), _dp(out), None),
              "x264dsp_cost_batch_dev")

    # ---- lookahead -----------------------------------------------------------------------
    def lookahead_frame_cost(self, g, slots_dev, b, p0, want_intra, mvs, costs, sums, row_satds=None, stream=None):
        """stream: a cudaStream_t as an int (e.g. torch.cuda.Stream().cuda_stream); None = the context's stream"""
        b = np.ascontiguousarray(b, np.int32)
        p0 = np.ascontiguousarray(p0, np.int32)
        wi = np.ascontiguousarray(want_intra, np.uint8)
        check(lib().x264dsp_lookahead_frame_cost_dev(
            self._h, C.byref(g), _dp(slots_dev), len(b), _hp(b, C.c_int32), _hp(p0, C.c_int32), _hp(wi),
            _dp(mvs), _dp(costs), _dp(sums), _dp(row_satds), C.c_void_p(stream) if stream else None),
            "x264dsp_lookahead_frame_cost_dev")

    def debug_copies_only(self, on):
        """measurement aid: lookahead_clips_host issues its copies but none of its kernels"""
        check(lib().x264dsp_debug_copies_only(self._h, int(bool(on))), "x264dsp_debug_copies_only")

    def lookahead_select_kernel(self, mode):
        """0 = by batch size, 1 = one block row per warp, 2 = four rows per warp, 3 = eight rows per warp"""
        check(lib().x264dsp_lookahead_select_kernel(self._h, int(mode)), "x264dsp_lookahead_select_kernel")

    def lookahead_clip_host(self, width, height, luma_frames):
        """luma_frames: uint8 numpy [n, height*width] in ordinary host memory.
        Returns (mvs [n, mb_count, 2] int16, costs [n, mb_count] int32, sums [n, LA_SUMS] int32)."""
        g = geometry(width, height)
        luma = np.ascontiguousarray(luma_frames, np.uint8)
        n = luma.shape[0]
        mvs = np.empty((n, g.mb_count, 2), np.int16)
        costs = np.empty((n, g.mb_count), np.int32)
        sums = np.empty((n, LA_SUMS), np.int32)
        check(lib().x264dsp_lookahead_clip_host(self._h, int(width), int(height), int(n), _hp(luma),
                                                _hp(mvs, C.c_int16), _hp(costs, C.c_int32),
                                                _hp(sums, C.c_int32)), "x264dsp_lookahead_clip_host")
        return mvs, costs, sums

    def lookahead_clips_host(self, width, height, n_clips, clip_len, luma, mvs=None, costs=None, sums=None):
        """luma: uint8 numpy [n_clips*clip_len, height*width]; outputs are allocated unless given
        (pass arrays created with `pinned_empty` to skip the staging copies)."""
        g = geometry(width, height)
        n = n_clips * clip_len
        assert luma.dtype == np.uint8 and luma.size == n * width * height and luma.flags.c_contiguous
        mvs = np.empty((n, g.mb_count, 2), np.int16) if mvs is None else mvs
        costs = np.empty((n, g.mb_count), np.int32) if costs is None else costs
        sums = np.empty((n, LA_SUMS), np.int32) if sums is None else sums
        check(lib().x264dsp_lookahead_clips_host(self._h, int(width), int(height), int(n_clips), int(clip_len),
                                                 _hp(luma), _hp(mvs, C.c_int16), _hp(costs, C.c_int32),
                                                 _hp(sums, C.c_int32)), "x264dsp_lookahead_clips_host")
        return mvs, costs, sums

    def pinned_empty(self, shape, dtype):
        """numpy array backed by pinned host memory (x264dsp_host_alloc); freed with the context"""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(lib().x264dsp_host_alloc(self._h, C.c_size_t(max(nbytes, 1)), C.byref(p)), "x264dsp_host_alloc")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def profile_enable(self, on=True):
        check(lib().x264dsp_profile_enable(self._h, int(on)), "x264dsp_profile_enable")

    def profile_read(self, kind):
        """(total_ms, launches) of kernel class `kind` (PROF_*) since profile_enable"""
        ms, n = C.c_double(), C.c_int()
        check(lib().x264dsp_profile_read(self._h, int(kind), C.byref(ms), C.byref(n)), "x264dsp_profile_read")
        return ms.value, n.value

    # ---- motion search -------------------------------------------------------------------
    def me_search_batch(self, g, fenc_slot, fref_slot, params, n, blocks_dev, results_dev):
        check(lib().x264dsp_me_search_batch_dev(self._h, C.byref(g), _dp(fenc_slot), _dp(fref_slot),
                                                C.byref(params), int(n), _dp(blocks_dev), _dp(results_dev),
                                                None), "x264dsp_me_search_batch_dev")

    def me_search_batch_ex(self, g, fenc_slot, fref_slot, params, n, blocks_dev, results_dev, mode, thresh_dev=None):
        """x264_me_search_ref with p_halfpel_thresh (mode 0), x264_me_refine_qpel_refdupe (1), x264_me_refine_qpel (2);
        thresh_dev: int32[n] device tensor, read and updated (None = NULL)"""
        check(lib().x264dsp_me_search_batch_ex_dev(self._h, C.byref(g), _dp(fenc_slot), _dp(fref_slot),
                                                   C.byref(params), int(n), _dp(blocks_dev), _dp(results_dev),
                                                   int(mode), _dp(thresh_dev), None), "x264dsp_me_search_batch_ex_dev")

    def me_search_sized(self, g, fenc_slot, fref_slot, params, i_pixel, n, blocks_dev, results_dev):
        """uniform-size list: the size-specialised kernel (x264dsp_me_search_sized_dev)"""
        check(lib().x264dsp_me_search_sized_dev(self._h, C.byref(g), _dp(fenc_slot), _dp(fref_slot),
                                                C.byref(params), int(i_pixel), int(n), _dp(blocks_dev),
                                                _dp(results_dev), None), "x264dsp_me_search_sized_dev")

    def me_search_sized_frames(self, g, fenc_slots, fref_slots, n_frames, params, i_pixel, n, blocks_dev, results_dev):
        """n_frames frame pairs (consecutive slots), n blocks each, one launch"""
        check(lib().x264dsp_me_search_sized_frames_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(fref_slots),
                                                       int(n_frames), C.byref(params), int(i_pixel), int(n),
                                                       _dp(blocks_dev), _dp(results_dev), None),
              "x264dsp_me_search_sized_frames_dev")

    # ---- residual / MC / deblock ---------------------------------------------------------
    def mc_frame(self, g, fref_slot, mv_dev, pred_slot):
        check(lib().x264dsp_mc_frame_dev(self._h, C.byref(g), _dp(fref_slot), _dp(mv_dev), _dp(pred_slot), None),
              "x264dsp_mc_frame_dev")

    def residual_frame(self, g, fenc_slot, pred_slot, qp, levels, nnz, cbp):
        check(lib().x264dsp_residual_frame_dev(self._h, C.byref(g), _dp(fenc_slot), _dp(pred_slot), int(qp),
                                               _dp(levels), _dp(nnz), _dp(cbp), None),
              "x264dsp_residual_frame_dev")

    def mc_frames(self, g, fref_slots, n_frames, mv_dev, pred_slots):
        check(lib().x264dsp_mc_frames_dev(self._h, C.byref(g), _dp(fref_slots), int(n_frames), _dp(mv_dev),
                                          _dp(pred_slots), None), "x264dsp_mc_frames_dev")

    def mc_frames_part(self, g, fref_slots, n_frames, mv8x8_dev, pred_slots):
        """one MV per 8x8 block: int16[n][mb][4][2]"""
        check(lib().x264dsp_mc_frames_part_dev(self._h, C.byref(g), _dp(fref_slots), int(n_frames), _dp(mv8x8_dev),
                                               _dp(pred_slots), None), "x264dsp_mc_frames_part_dev")

    def p_frames(self, g, fenc_slots, fref_slots, recon_slots, n_frames, prm, lowres_mv, l0_mv16, mb_type, mv, mvr,
                 levels, nnz, cbp, mvd=None):
        """x264_macroblock_analyse + x264_macroblock_encode for every macroblock of n_frames independent P frames
        (x264dsp_p_frames_dev); lowres_mv / l0_mv16 may be None"""
        check(lib().x264dsp_p_frames_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(fref_slots), _dp(recon_slots),
                                         int(n_frames), C.byref(prm), _dp(lowres_mv) if lowres_mv is not None else None,
                                         _dp(l0_mv16) if l0_mv16 is not None else None, _dp(mb_type), _dp(mv), _dp(mvr),
                                         _dp(mvd), _dp(levels), _dp(nnz), _dp(cbp), None), "x264dsp_p_frames_dev")

    def p_frames_part(self, g, fenc_slots, fref_slots, recon_slots, n_frames, prm, lowres_mv, l0_mv16, mb_type, partition,
                      mv8, mvr, levels, nnz, cbp, mvd8=None):
        """the P-slice loop with the sub-16x16 partitions of analyse.inter = PSUB16x16 (x264dsp_p_frames_part_dev): vectors
        and differences per 8x8 block, int16[n][mb][4][2]"""
        check(lib().x264dsp_p_frames_part_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(fref_slots), _dp(recon_slots),
                                              int(n_frames), C.byref(prm), _dp(lowres_mv) if lowres_mv is not None else None,
                                              _dp(l0_mv16) if l0_mv16 is not None else None, _dp(mb_type), _dp(partition),
                                              _dp(mv8), _dp(mvr), _dp(mvd8), _dp(levels), _dp(nnz), _dp(cbp), None),
              "x264dsp_p_frames_part_dev")

    def i_frames(self, g, fenc_slots, recon_slots, n_frames, qp, mb_type, mode16, chroma_mode, modes4, levels, luma_dc, nnz, cbp):
        """intra analysis + coding of every macroblock of n_frames independent I frames (x264dsp_i_frames_dev)"""
        check(lib().x264dsp_i_frames_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(recon_slots), int(n_frames), int(qp),
                                         _dp(mb_type), _dp(mode16), _dp(chroma_mode), _dp(modes4), _dp(levels), _dp(luma_dc),
                                         _dp(nnz), _dp(cbp), None), "x264dsp_i_frames_dev")

    def p_frames_host(self, w, h, n_frames, i420, prm, mb_type, mv, mvr, mvd, levels, nnz, cbp, recon_i420):
        """x264dsp_p_frames_host: numpy (ideally pinned) arrays in and out"""
        check(lib().x264dsp_p_frames_host(self._h, int(w), int(h), int(n_frames), _hp(i420), C.byref(prm),
                                          mb_type.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p),
                                          mvr.ctypes.data_as(C.c_void_p), mvd.ctypes.data_as(C.c_void_p) if mvd is not None else None,
                                          levels.ctypes.data_as(C.c_void_p), _hp(nnz), cbp.ctypes.data_as(C.c_void_p),
                                          _hp(recon_i420)), "x264dsp_p_frames_host")

    def p_frames_part_host(self, w, h, n_frames, i420, prm, mb_type, partition, mv8, mvr, mvd8, levels, nnz, cbp, recon_i420):
        """x264dsp_p_frames_part_host: as p_frames_host with partition uint8[n][mb] and mv8 / mvd8 int16[n][mb][4][2]"""
        check(lib().x264dsp_p_frames_part_host(self._h, int(w), int(h), int(n_frames), _hp(i420), C.byref(prm),
                                               mb_type.ctypes.data_as(C.c_void_p), _hp(partition), mv8.ctypes.data_as(C.c_void_p),
                                               mvr.ctypes.data_as(C.c_void_p),
                                               mvd8.ctypes.data_as(C.c_void_p) if mvd8 is not None else None,
                                               levels.ctypes.data_as(C.c_void_p), _hp(nnz), cbp.ctypes.data_as(C.c_void_p),
                                               _hp(recon_i420)), "x264dsp_p_frames_part_host")

    def boundary_strength_frames(self, g, n_frames, mb_type, nnz, mv8, bs):
        check(lib().x264dsp_boundary_strength_frames_dev(self._h, C.byref(g), int(n_frames), _dp(mb_type), _dp(nnz), _dp(mv8), _dp(bs),
                                                         None), "x264dsp_boundary_strength_frames_dev")

    def gops_encode(self, g, fenc_slots, recon_slots, n_gops, gop_len, prm, lowres_mv, out):
        """x264dsp_gops_encode_dev; out: dict of device tensors mb_type, partition, mv8, mvr, mvd8, levels, nnz, cbp, mode16,
        chroma_mode, modes4, luma_dc (position-major: [t][gop][mb]...)"""
        check(lib().x264dsp_gops_encode_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(recon_slots), int(n_gops), int(gop_len),
                                            C.byref(prm), _dp(lowres_mv) if lowres_mv is not None else None, _dp(out["mb_type"]),
                                            _dp(out["partition"]), _dp(out["mv8"]), _dp(out["mvr"]), _dp(out.get("mvd8")),
                                            _dp(out["levels"]), _dp(out["nnz"]), _dp(out["cbp"]), _dp(out["mode16"]),
                                            _dp(out["chroma_mode"]), _dp(out["modes4"]), _dp(out["luma_dc"]), None),
              "x264dsp_gops_encode_dev")

    def gops_encode_host(self, w, h, n_gops, gop_len, i420, prm, out, packed, frame_offset, frame_size, mb_offset):
        """x264dsp_gops_encode_host: numpy (ideally pinned) arrays; out: dict mb_type, partition, mv8, mvr, mvd8, nnz, cbp, mode16,
        chroma_mode, modes4, luma_dc"""
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        check(lib().x264dsp_gops_encode_host(self._h, int(w), int(h), int(n_gops), int(gop_len), _hp(i420), C.byref(prm), vp(out["mb_type"]),
                                             vp(out["partition"]), vp(out["mv8"]), vp(out["mvr"]), vp(out.get("mvd8")), vp(out["nnz"]),
                                             vp(out["cbp"]), vp(out["mode16"]), vp(out["chroma_mode"]), vp(out["modes4"]),
                                             vp(out["luma_dc"]), vp(packed), C.c_int64(int(packed.size)), vp(frame_offset),
                                             vp(frame_size), vp(mb_offset)), "x264dsp_gops_encode_host")

    def levels_pack(self, n_frames, mb_count, levels, nnz, packed, packed_stride, mb_offset, frame_total):
        """x264dsp_levels_pack_dev: the dense levels of n_frames as the compact stream the entropy coder reads"""
        check(lib().x264dsp_levels_pack_dev(self._h, int(n_frames), int(mb_count), _dp(levels), _dp(nnz), _dp(packed),
                                            C.c_int64(int(packed_stride)), _dp(mb_offset), _dp(frame_total), None),
              "x264dsp_levels_pack_dev")

    def p_frames_host_packed(self, w, h, n_frames, i420, prm, mb_type, partition, mv, mvr, mvd, packed, frame_offset, mb_offset,
                             nnz, cbp, recon_i420=None):
        """x264dsp_p_frames_host_packed: numpy (ideally pinned) arrays; partition None = one vector per macroblock"""
        check(lib().x264dsp_p_frames_host_packed(self._h, int(w), int(h), int(n_frames), _hp(i420), C.byref(prm),
                                                 mb_type.ctypes.data_as(C.c_void_p), _hp(partition) if partition is not None else None,
                                                 mv.ctypes.data_as(C.c_void_p), mvr.ctypes.data_as(C.c_void_p),
                                                 mvd.ctypes.data_as(C.c_void_p) if mvd is not None else None,
                                                 packed.ctypes.data_as(C.c_void_p), C.c_int64(int(packed.size)),
                                                 frame_offset.ctypes.data_as(C.c_void_p), mb_offset.ctypes.data_as(C.c_void_p),
                                                 _hp(nnz), cbp.ctypes.data_as(C.c_void_p),
                                                 _hp(recon_i420) if recon_i420 is not None else None), "x264dsp_p_frames_host_packed")

    def residual_frames(self, g, fenc_slots, pred_slots, n_frames, qp, levels, nnz, cbp):
        check(lib().x264dsp_residual_frames_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(pred_slots), int(n_frames),
                                                int(qp), _dp(levels), _dp(nnz), _dp(cbp), None),
              "x264dsp_residual_frames_dev")

    def residual_frames_typed(self, g, fenc_slots, pred_slots, n_frames, qp, mb_kind, levels, luma_dc, nnz, cbp,
                              i4_modes=None):
        """mb_kind: uint8 per macroblock, 0 = inter (P slice), 1 = I16x16, 2 (6) = I4x4 with i4_modes uint8[n][mb][16]
        (I slice); luma_dc: int16[n][mb][16] or None"""
        check(lib().x264dsp_residual_frames_typed_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(pred_slots), int(n_frames),
                                                      int(qp), _dp(mb_kind) if mb_kind is not None else None,
                                                      _dp(i4_modes) if i4_modes is not None else None, _dp(levels),
                                                      _dp(luma_dc) if luma_dc is not None else None, _dp(nnz), _dp(cbp), None),
              "x264dsp_residual_frames_typed_dev")

    def predict_mv_batch(self, n, nb, i_ref, mvp, pskip_mv, shape=None):
        """nb: uint8[n][20] = {int8 ref[4], int16 mv[4][2]} (left, top, top-right, top-left); outputs int16[n][2];
        shape: uint8[n] partition rule (0 16x16 / 8x8, 1 / 2 16x8, 3 / 4 8x16, + 8: top-right not reachable)"""
        check(lib().x264dsp_predict_mv_batch_dev(self._h, int(n), _dp(nb), _dp(i_ref) if i_ref is not None else None,
                                                 _dp(shape) if shape is not None else None,
                                                 _dp(mvp) if mvp is not None else None,
                                                 _dp(pskip_mv) if pskip_mv is not None else None, None),
              "x264dsp_predict_mv_batch_dev")

    def predict_mvc_16x16_frames(self, mb_w, mb_h, n_frames, lowres_mv, mvr, l0_mv16, scale, mvc, n_mvc):
        check(lib().x264dsp_predict_mvc_16x16_frames_dev(self._h, int(mb_w), int(mb_h), int(n_frames),
                                                         _dp(lowres_mv) if lowres_mv is not None else None, _dp(mvr),
                                                         _dp(l0_mv16) if l0_mv16 is not None else None, int(scale),
                                                         _dp(mvc), _dp(n_mvc), None), "x264dsp_predict_mvc_16x16_frames_dev")

    def probe_pskip_frames(self, g, fenc_slots, pred_slots, n_frames, qp, skip):
        check(lib().x264dsp_probe_pskip_frames_dev(self._h, C.byref(g), _dp(fenc_slots), _dp(pred_slots), int(n_frames),
                                                   int(qp), _dp(skip), None), "x264dsp_probe_pskip_frames_dev")

    def deblock_frame(self, g, slot, mb_type, partition, cbp, bs, qp, alpha_off=0, beta_off=0):
        check(lib().x264dsp_deblock_frame_dev(self._h, C.byref(g), _dp(slot), _dp(mb_type), _dp(partition),
                                              _dp(cbp), _dp(bs), int(qp), int(alpha_off), int(beta_off), None),
              "x264dsp_deblock_frame_dev")

    def deblock_frames(self, g, slots, n_frames, mb_type, partition, cbp, bs, qp, alpha_off=0, beta_off=0, stream=None):
        """n_frames consecutive slots in one launch; the per-MB arrays hold n_frames x mb_count entries"""
        check(lib().x264dsp_deblock_frames_dev(self._h, C.byref(g), _dp(slots), int(n_frames), _dp(mb_type),
                                               _dp(partition), _dp(cbp), _dp(bs), int(qp), int(alpha_off),
                                               int(beta_off), C.c_void_p(stream) if stream else None),
              "x264dsp_deblock_frames_dev")

    def deblock_strength(self, n, nnz, ref, mv, bs):
        check(lib().x264dsp_deblock_strength_dev(self._h, int(n), _dp(nnz), _dp(ref), _dp(mv), _dp(bs), None),
              "x264dsp_deblock_strength_dev")

    def macroblock_deblock_strength(self, n, mb_type, nnz, ref, mv, bs):
        """x264_macroblock_deblock_strength: mb_type int8[n] (0..3 = intra -> inner edges 3); None = all inter"""
        check(lib().x264dsp_macroblock_deblock_strength_dev(self._h, int(n), _dp(mb_type), _dp(nnz), _dp(ref), _dp(mv),
                                                            _dp(bs), None), "x264dsp_macroblock_deblock_strength_dev")

    # ---- full-resolution paths from host memory ------------------------------------------
    def me_search_frames_host(self, width, height, luma, params, sizes, blocks, results=None):
        """luma: uint8 [n_pairs+1, h*w]; sizes: list of i_pixel; blocks: list of ME_BLOCK_DTYPE arrays [n_pairs*n] (pair-major).
        Returns the list of ME_RESULT_DTYPE arrays (pass pinned `results` to skip staging)."""
        n_pairs = luma.shape[0] - 1
        ns = len(sizes)
        nb = [len(b) // n_pairs for b in blocks]
        if results is None:
            results = [np.zeros(len(b), ME_RESULT_DTYPE) for b in blocks]
        ip = (C.c_int32 * ns)(*sizes)
        nbl = (C.c_int32 * ns)(*nb)
        bp = (C.c_void_p * ns)(*[b.ctypes.data for b in blocks])
        rp = (C.c_void_p * ns)(*[r.ctypes.data for r in results])
        check(lib().x264dsp_me_search_frames_host(self._h, int(width), int(height), int(n_pairs), _hp(luma), C.byref(params),
                                                  ns, ip, nbl, bp, rp), "x264dsp_me_search_frames_host")
        return results

    def recon_frames_host(self, width, height, i420, mv16, qp, mb_type, partition, bs, out=None, alpha_off=0, beta_off=0):
        """i420: uint8 [n+1, w*h*3/2]; mv16 int16 [n, mb, 2]; mb_type int8 [n, mb]; partition uint8 [n, mb]; bs uint8 [n, mb, 64].
        Returns (levels, nnz, cbp, recon_i420); `out` may hold those four arrays preallocated (e.g. pinned)."""
        g = geometry(width, height)
        n = i420.shape[0] - 1
        if out is None:
            out = (np.zeros((n, g.mb_count, RES_LEVELS_PER_MB), np.int16), np.zeros((n, g.mb_count, RES_NNZ_PER_MB), np.uint8),
                   np.zeros((n, g.mb_count), np.int16), np.zeros((n, width * height * 3 // 2), np.uint8))
        lv, nz, cbp, rec = out
        check(lib().x264dsp_recon_frames_host(self._h, int(width), int(height), int(n), _hp(i420), _hp(mv16, C.c_int16), int(qp),
                                              _hp(mb_type, C.c_int8), _hp(partition), _hp(bs), int(alpha_off), int(beta_off),
                                              _hp(lv, C.c_int16), _hp(nz), _hp(cbp, C.c_int16), _hp(rec)),
              "x264dsp_recon_frames_host")
        return out

    def frame_store_i420(self, g, slots_dev, i420_dev, n_frames):
        check(lib().x264dsp_frame_store_i420_dev(self._h, C.byref(g), _dp(slots_dev), _dp(i420_dev), int(n_frames), None),
              "x264dsp_frame_store_i420_dev")
