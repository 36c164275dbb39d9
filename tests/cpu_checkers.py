"""ctypes doors to the two CPU checkers (TEST INFRASTRUCTURE, never used by the product):

  oracle  -- oracle/_build/libx264dsp_oracle.so, our plain-C restatement (oracle/xo*.c)
  ref     -- oracle/_ref/libx264ref.so, the UNMODIFIED reference compiled from /root/reference
             (only buildable in the dev container; travels to the GPU box as a built file)

Both are built on demand with `make -C oracle`.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libx264dsp_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libx264ref.so")
REF_O3_SO = os.path.join(ORACLE_DIR, "_ref", "o3", "libx264ref.so")       # same sources at -O3 -march=x86-64-v3
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "x264ref")
REFERENCE_TREE = "/root/reference"

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i16p = C.POINTER(C.c_int16)
u16p = C.POINTER(C.c_uint16)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class Geom(C.Structure):
    """x264dsp_geom_t (include/x264dsp_b200.h)"""
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "mb_w", "mb_h", "mb_count", "luma_w", "luma_h",
        "luma_stride", "luma_plane_size", "luma_origin",
        "chroma_stride", "chroma_h", "chroma_plane_size", "chroma_origin",
        "lowres_w", "lowres_h", "lowres_stride", "lowres_plane_size", "lowres_origin",
        "slot_chroma_off", "slot_lowres_off")] + [("slot_bytes", C.c_int64)] + [(n, C.c_int32) for n in (
        "tile_w", "tile_h", "tiled_plane_size", "slot_tiled_off")]


class MeBlock(C.Structure):
    """x264dsp_me_block_t == xref_me_in_t"""
    _fields_ = [("i_pixel", C.c_int32), ("bx", C.c_int32), ("by", C.c_int32),
                ("mvp", C.c_int16 * 2), ("i_mvc", C.c_int32), ("mvc", (C.c_int16 * 2) * 16),
                ("mv_min_fpel", C.c_int32 * 2), ("mv_max_fpel", C.c_int32 * 2),
                ("mv_min_spel", C.c_int32 * 2), ("mv_max_spel", C.c_int32 * 2)]


class MeResult(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("cost", C.c_int32), ("cost_mv", C.c_int32)]


class MeParams(C.Structure):
    _fields_ = [("me_method", C.c_int32), ("subpel_refine", C.c_int32), ("me_range", C.c_int32),
                ("qp", C.c_int32), ("refine_qpel", C.c_int32)]


ME_BLOCK_DTYPE = np.dtype([("i_pixel", "<i4"), ("bx", "<i4"), ("by", "<i4"), ("mvp", "<i2", (2,)),
                           ("i_mvc", "<i4"), ("mvc", "<i2", (16, 2)),
                           ("mv_min_fpel", "<i4", (2,)), ("mv_max_fpel", "<i4", (2,)),
                           ("mv_min_spel", "<i4", (2,)), ("mv_max_spel", "<i4", (2,))], align=True)
ME_RESULT_DTYPE = np.dtype([("mv", "<i2", (2,)), ("cost", "<i4"), ("cost_mv", "<i4")], align=True)
assert ME_BLOCK_DTYPE.itemsize == C.sizeof(MeBlock)
assert ME_RESULT_DTYPE.itemsize == C.sizeof(MeResult)

BLOCK_W = [16, 16, 8, 8, 8, 4, 4, 4]
BLOCK_H = [16, 8, 16, 8, 4, 8, 4, 16]


def ptr(a, t=u8p):
    return a.ctypes.data_as(t)


def _make(target=None):
    cmd = ["make", "-s", "-C", ORACLE_DIR] + ([target] if target else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    """our C restatement"""
    global _oracle
    if _oracle is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        stale = (not os.path.exists(ORACLE_SO) or not os.path.exists(SYNTH_SO)
                 or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(s) for s in srcs))
        if stale:
            _make()
        lib = C.CDLL(ORACLE_SO)
        lib.xo_var.restype = C.c_uint64
        _oracle = lib
    return _oracle


SYNTH_SO = os.path.join(ORACLE_DIR, "_build", "libx264dsp_synth.so")
_synth = None


def synth_frame(width, height, n, cut_frame=-1, luma_only=False):
    """the shared synthetic generator (x264dsp_synth_frame) WITHOUT the product library: planar I420 or luma of frame n"""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_SO):
            _make()
        _synth = C.CDLL(SYNTH_SO)
    y = np.empty(width * height, np.uint8)
    if luma_only:
        rc = _synth.x264dsp_synth_frame(width, height, n, cut_frame, ptr(y), None, None)
        assert rc == 0
        return y
    cw, ch = width // 2, height // 2
    buf = np.empty(width * height + 2 * cw * ch, np.uint8)
    rc = _synth.x264dsp_synth_frame(width, height, n, cut_frame, ptr(buf), ptr(buf[width * height:]),
                                    ptr(buf[width * height + cw * ch:]))
    assert rc == 0
    return buf


def ref_available():
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_TREE)


def _load_ref(path):
    lib = C.CDLL(path)
    lib.xref_open.restype = C.c_void_p
    lib.xref_frame_new.restype = C.c_void_p
    lib.xref_frame_ptr.restype = C.c_void_p
    lib.xref_cost_mv.restype = C.POINTER(C.c_uint16)
    lib.xref_time_lookahead.restype = C.c_double
    for name in ("xref_pixf", "xref_dctf", "xref_zigzagf", "xref_mcf", "xref_quantf", "xref_loopf"):
        getattr(lib, name).restype = C.c_void_p
    return lib


def ref():
    """the unmodified reference; None when it can neither be found nor built"""
    global _ref
    if _ref is None:
        if os.path.isdir(REFERENCE_TREE):
            _make("ref")
        if not os.path.exists(REF_SO):
            return None
        _ref = _load_ref(REF_SO)
    return _ref


_ref_o3 = None


def ref_o3():
    """the same reference sources built at -O3 -march=x86-64-v3; None when absent"""
    global _ref_o3
    if _ref_o3 is None:
        if ref() is None or not os.path.exists(REF_O3_SO):
            return None
        _ref_o3 = _load_ref(REF_O3_SO)
    return _ref_o3


class RefEncoder:
    """an encoder instance of the unmodified reference (x264_encoder_open)"""

    def __init__(self, width, height, me=0, subme=2, me_range=16, qp=26, psub16x16=0, lib=None):
        self.lib = lib if lib is not None else ref()
        assert self.lib is not None
        self.h = C.c_void_p(self.lib.xref_open(width, height, me, subme, me_range, qp, psub16x16))
        assert self.h.value
        self.width, self.height = width, height
        g = (C.c_int * 16)()
        self.lib.xref_geometry(self.h, g)
        self.geom = list(g)

    def new_frame(self, fdec):
        f = C.c_void_p(self.lib.xref_frame_new(self.h, int(fdec)))
        assert f.value
        return f

    def load(self, f, i420):
        w, h = self.width, self.height
        y = i420[: w * h]
        u = i420[w * h: w * h + (w // 2) * (h // 2)]
        v = i420[w * h + (w // 2) * (h // 2):]
        self.lib.xref_frame_load_i420(self.h, f, ptr(y), ptr(u), ptr(v))

    def buffer(self, f, which, nbytes):
        """numpy view of one of the frame's raw allocations (10 luma, 11 chroma, 12 lowres)"""
        p = self.lib.xref_frame_ptr(f, which)
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(nbytes,))

    def close(self):
        pass  # encoders are leaked on purpose: x264_encoder_close prints statistics


def oracle_geom(width, height):
    g = Geom()
    oracle().xo_geometry(width, height, C.byref(g))
    return g


def synth_clip(width, height, n_frames, seed=1234, cut_frame=-1):
    """Small numpy synthetic clip for CPU tests: smooth texture panning (3,2) px/frame plus two
    moving gradient squares plus +-2 noise; optional scene cut.  Returns a list of I420 arrays."""
    rng = np.random.RandomState(seed)

    def texture(r):
        t = r.randint(0, 256, size=(height + 64, width + 64)).astype(np.float32)
        k = 5
        c = np.cumsum(np.cumsum(np.pad(t, ((k, 0), (k, 0)), mode="wrap"), 0), 1)
        t = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
        t = (t - t.min()) / max(1e-6, t.max() - t.min()) * 255.0
        return t

    tex = texture(rng)
    frames = []
    for n in range(n_frames):
        if n == cut_frame:
            tex = texture(np.random.RandomState(seed + 999))
        yy = (np.arange(height)[:, None] + 2 * n) % tex.shape[0]
        xx = (np.arange(width)[None, :] + 3 * n) % tex.shape[1]
        y = tex[yy, xx].copy()
        for (sx, sy, vx, vy) in ((width // 4, height // 4, 3, 2), (width // 2, height // 2, -2, 1)):
            x0 = (sx + vx * n) % max(1, width - 48)
            y0 = (sy + vy * n) % max(1, height - 48)
            hh = min(48, height - y0)
            ww = min(48, width - x0)
            grad = (np.arange(ww)[None, :] * 4 + np.arange(hh)[:, None] * 2) % 256
            y[y0:y0 + hh, x0:x0 + ww] = grad
        y = y + np.random.RandomState(seed + 17 * n + 1).randint(-2, 3, size=y.shape)
        y = np.clip(y, 0, 255).astype(np.uint8)
        u = (y[::2, ::2] // 2 + 64).astype(np.uint8)
        v = (191 - y[::2, ::2] // 2).astype(np.uint8)
        frames.append(np.concatenate([y.ravel(), u.ravel(), v.ravel()]))
    return frames
