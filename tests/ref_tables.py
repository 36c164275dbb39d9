"""ctypes mirrors of the reference's function-pointer tables (non-_DEBUG layout):
x264_pixel_function_t (common/pixel.h:55-116), x264_dct_function_t / x264_zigzag_function_t
(common/dct.h:8-33), x264_mc_functions_t (common/mc.h:31-79), x264_quant_function_t
(common/quant.h:8-31), x264_deblock_function_t (common/frame.h:206-216).

The same classes read the tables of the unmodified reference (oracle/_ref) and the tables filled
by the product's drop-in x264_*_init (include/x264dsp_tables.h)."""
import ctypes as C

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i16p = C.POINTER(C.c_int16)
u16p = C.POINTER(C.c_uint16)
intp = C.POINTER(C.c_int)
iptr = C.c_ssize_t          # intptr_t

CMP = C.CFUNCTYPE(C.c_int, u8p, iptr, u8p, iptr)
CMP_X3 = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, iptr, intp)
CMP_X4 = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, u8p, iptr, intp)
VAR = C.CFUNCTYPE(C.c_uint64, u8p, iptr)
VAR2 = C.CFUNCTYPE(C.c_int, u8p, iptr, u8p, iptr, intp)
INTRA_X3 = C.CFUNCTYPE(None, u8p, u8p, intp)
INTRA_X9 = C.CFUNCTYPE(C.c_int, u8p, u8p, u16p)


class PixelTable(C.Structure):
    _fields_ = [
        ("sad", CMP * 8), ("ssd", CMP * 8), ("satd", CMP * 8), ("mbcmp", CMP * 8),
        ("mbcmp_unaligned", CMP * 8), ("fpelcmp", CMP * 8), ("fpelcmp_x3", CMP_X3 * 7),
        ("fpelcmp_x4", CMP_X4 * 7), ("sad_aligned", CMP * 8), ("var", VAR * 4), ("var2", VAR2 * 4),
        ("sad_x3", CMP_X3 * 7), ("sad_x4", CMP_X4 * 7), ("satd_x3", CMP_X3 * 7), ("satd_x4", CMP_X4 * 7),
        ("intra_mbcmp_x3_16x16", INTRA_X3), ("intra_satd_x3_16x16", INTRA_X3), ("intra_sad_x3_16x16", INTRA_X3),
        ("intra_mbcmp_x3_4x4", INTRA_X3), ("intra_satd_x3_4x4", INTRA_X3), ("intra_sad_x3_4x4", INTRA_X3),
        ("intra_mbcmp_x4_4x4_h", INTRA_X3), ("intra_satd_x4_4x4_h", INTRA_X3), ("intra_sad_x4_4x4_h", INTRA_X3),
        ("intra_mbcmp_x4_4x4_v", INTRA_X3), ("intra_satd_x4_4x4_v", INTRA_X3), ("intra_sad_x4_4x4_v", INTRA_X3),
        ("intra_mbcmp_x3_chroma", INTRA_X3), ("intra_satd_x3_chroma", INTRA_X3), ("intra_sad_x3_chroma", INTRA_X3),
        ("intra_mbcmp_x3_8x8c", INTRA_X3), ("intra_satd_x3_8x8c", INTRA_X3), ("intra_sad_x3_8x8c", INTRA_X3),
        ("intra_mbcmp_x9_4x4", INTRA_X9), ("intra_satd_x9_4x4", INTRA_X9), ("intra_sad_x9_4x4", INTRA_X9),
    ]


SUB_DCT = C.CFUNCTYPE(None, i16p, u8p, u8p)
ADD_IDCT = C.CFUNCTYPE(None, u8p, i16p)
DC_FN = C.CFUNCTYPE(None, i16p)


class DctTable(C.Structure):
    _fields_ = [
        ("sub4x4_dct", SUB_DCT), ("add4x4_idct", ADD_IDCT),
        ("sub8x8_dct", SUB_DCT), ("sub8x8_dct_dc", SUB_DCT), ("add8x8_idct", ADD_IDCT), ("add8x8_idct_dc", ADD_IDCT),
        ("sub16x16_dct", SUB_DCT), ("add16x16_idct", ADD_IDCT), ("add16x16_idct_dc", ADD_IDCT),
        ("dct4x4dc", DC_FN), ("idct4x4dc", DC_FN),
    ]


class ZigzagTable(C.Structure):
    _fields_ = [("scan_4x4", C.CFUNCTYPE(None, i16p, i16p))]


MC_LUMA = C.CFUNCTYPE(None, u8p, iptr, C.POINTER(u8p), iptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p)
GET_REF = C.CFUNCTYPE(C.c_void_p, u8p, C.POINTER(iptr), C.POINTER(u8p), iptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p)
MC_CHROMA = C.CFUNCTYPE(None, u8p, u8p, iptr, u8p, iptr, C.c_int, C.c_int, C.c_int, C.c_int)
MC_COPY = C.CFUNCTYPE(None, u8p, iptr, u8p, iptr, C.c_int)
HPEL = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, iptr, C.c_int, C.c_int, i16p)
LOWRES = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, u8p, iptr, iptr, C.c_int, C.c_int)
VOIDP = C.c_void_p


class McTable(C.Structure):
    _fields_ = [
        ("mc_luma", MC_LUMA), ("get_ref", GET_REF), ("mc_chroma", MC_CHROMA), ("copy", MC_COPY * 7),
        ("store_interleave_chroma", C.CFUNCTYPE(None, u8p, iptr, u8p, u8p, C.c_int)),
        ("load_deinterleave_chroma_fenc", C.CFUNCTYPE(None, u8p, u8p, iptr, C.c_int)),
        ("load_deinterleave_chroma_fdec", C.CFUNCTYPE(None, u8p, u8p, iptr, C.c_int)),
        ("plane_copy", C.CFUNCTYPE(None, u8p, iptr, u8p, iptr, C.c_int, C.c_int)),
        ("plane_copy_interleave", C.CFUNCTYPE(None, u8p, iptr, u8p, iptr, u8p, iptr, C.c_int, C.c_int)),
        ("plane_copy_deinterleave", C.CFUNCTYPE(None, u8p, iptr, u8p, iptr, u8p, iptr, C.c_int, C.c_int)),
        ("plane_copy_deinterlace", VOIDP), ("plane_deinterlace", VOIDP),
        ("hpel_filter", HPEL),
        ("prefetch_fenc", VOIDP), ("prefetch_fenc_420", VOIDP), ("prefetch_ref", VOIDP),
        ("memcpy_aligned", VOIDP), ("memzero_aligned", VOIDP),
        ("frame_init_lowres_core", LOWRES),
    ]


QUANT4 = C.CFUNCTYPE(C.c_int, i16p, u16p, u16p)
QUANT_DC = C.CFUNCTYPE(C.c_int, i16p, C.c_int, C.c_int)
DEQUANT = C.CFUNCTYPE(None, i16p, C.POINTER(C.c_int), C.c_int)
COEF_INT = C.CFUNCTYPE(C.c_int, i16p)
DENOISE = C.CFUNCTYPE(None, i16p, C.POINTER(C.c_uint32), u16p, C.c_int)


class RunLevel(C.Structure):
    """x264_run_level_t (common/bitstream.h:33-38)"""
    _fields_ = [("last", C.c_int), ("mask", C.c_int), ("level", C.c_int16 * 16)]


LEVEL_RUN = C.CFUNCTYPE(C.c_int, i16p, C.POINTER(RunLevel))


class QuantTable(C.Structure):
    _fields_ = [
        ("quant_4x4", QUANT4), ("quant_4x4_dc", QUANT_DC), ("quant_2x2_dc", QUANT_DC),
        ("dequant_4x4", DEQUANT), ("dequant_4x4_dc", DEQUANT),
        ("optimize_chroma_2x2_dc", C.CFUNCTYPE(C.c_int, i16p, C.c_int)),
        ("denoise_dct", DENOISE),
        ("decimate_score15", COEF_INT), ("decimate_score16", COEF_INT),
        ("coeff_last", COEF_INT * 14), ("coeff_last4", COEF_INT), ("coeff_last8", COEF_INT),
        ("coeff_level_run", LEVEL_RUN * 13), ("coeff_level_run4", LEVEL_RUN), ("coeff_level_run8", LEVEL_RUN),
    ]


DEBLOCK_INTER = C.CFUNCTYPE(None, u8p, iptr, C.c_int, C.c_int, i8p)
DEBLOCK_INTRA = C.CFUNCTYPE(None, u8p, iptr, C.c_int, C.c_int)


class DeblockTable(C.Structure):
    _fields_ = [
        ("deblock_luma", DEBLOCK_INTER * 2), ("deblock_chroma", DEBLOCK_INTER * 2),
        ("deblock_luma_intra", DEBLOCK_INTRA * 2), ("deblock_chroma_intra", DEBLOCK_INTRA * 2),
        ("deblock_strength", C.CFUNCTYPE(None, u8p, i8p, i16p, u8p)),
    ]


TABLES = [PixelTable, DctTable, ZigzagTable, McTable, QuantTable, DeblockTable]
