/*
 * x264dsp_tables.h -- the reference's function-pointer tables, served by libx264dsp_b200.so.
 *
 * The six structs below have the layout of the reference's tables WITHOUT _DEBUG (the reference
 * embeds them by value in struct x264_t, common/common.h:1088-1098) and the six init functions
 * have the reference's signatures, so that x264_encoder_open (encoder/encoder.c:551-560) can be
 * linked against this library instead of common/{pixel,dct,mc,quant,deblock}.c:
 *
 *     x264_pixel_init   common/pixel.h:118      x264_mc_init      common/mc.h:81
 *     x264_dct_init     common/dct.h:35         x264_quant_init   common/quant.h:33
 *     x264_zigzag_init  common/dct.h:36         x264_deblock_init common/frame.h:232
 *
 * Every entry is a per-call shim: it stages the caller's HOST operands to the GPU, runs the
 * corresponding CUDA leaf routine and copies the result back.  That is a drop-in for correctness,
 * not for speed -- a 4x4 block is not worth a PCIe round trip; the frame-batched entry points of
 * x264dsp_b200.h are the fast path with identical per-block semantics.  The shims use a process
 * wide context on device $X264DSP_DEVICE (default 0), created by the first init call; if no CUDA
 * device can be opened the init functions abort() -- there is no CPU fallback to fall back to.
 * All shims stage through ONE buffer of that context: table members must be called from one thread at a time (the
 * reference itself has no slice or lookahead threads; two encoders in two threads of one process need a lock around their
 * table calls, or the frame-batched entry points, which take a context per caller).
 *
 * Members the hot path does not cover stay NULL, exactly as listed in DESIGN.md ("out of scope"):
 * coeff_level_run*, denoise_dct (entropy side / off by default), intra_*_x4_4x4_{h,v} and
 * intra_*_x9_4x4 (need the nine 4x4 predictors: SURVEY.md 8(f) N1), plane_copy_deinterlace /
 * plane_deinterlace (TI capture path), prefetch_* (no-ops in the reference, set to no-ops here).
 */
#ifndef X264DSP_TABLES_H
#define X264DSP_TABLES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef X264_COMMON_H            /* when built inside the reference these come from common.h */
typedef uint8_t  pixel;
typedef int16_t  dctcoef;
typedef uint16_t udctcoef;
struct x264_t;
typedef struct x264_t x264_t;
struct x264_weight_t;
typedef struct x264_weight_t x264_weight_t;
struct x264_run_level_t;
typedef struct x264_run_level_t x264_run_level_t;
#define X264DSP_OWN_TABLE_TYPES 1
#endif

#ifdef X264DSP_OWN_TABLE_TYPES

/* ---- common/pixel.h:10-12, 55-116 */
typedef int  (*x264_pixel_cmp_t)( pixel *, intptr_t, pixel *, intptr_t );
typedef void (*x264_pixel_cmp_x3_t)( pixel *, pixel *, pixel *, pixel *, intptr_t, int[3] );
typedef void (*x264_pixel_cmp_x4_t)( pixel *, pixel *, pixel *, pixel *, pixel *, intptr_t, int[4] );
typedef void (*x264_intra_cmp_t)( pixel *fenc, pixel *fdec, int res[] );
typedef int  (*x264_intra_x9_t)( pixel *fenc, pixel *fdec, uint16_t *bitcosts );

typedef struct
{
    x264_pixel_cmp_t    sad[8], ssd[8], satd[8];
    x264_pixel_cmp_t    mbcmp[8], mbcmp_unaligned[8], fpelcmp[8];    /* aliases set by the encoder (encoder.c:412-436) */
    x264_pixel_cmp_x3_t fpelcmp_x3[7];
    x264_pixel_cmp_x4_t fpelcmp_x4[7];
    x264_pixel_cmp_t    sad_aligned[8];
    uint64_t (*var[4])( pixel *pix, intptr_t stride );
    int      (*var2[4])( pixel *pix1, intptr_t stride1, pixel *pix2, intptr_t stride2, int *ssd );
    x264_pixel_cmp_x3_t sad_x3[7];
    x264_pixel_cmp_x4_t sad_x4[7];
    x264_pixel_cmp_x3_t satd_x3[7];
    x264_pixel_cmp_x4_t satd_x4[7];
    x264_intra_cmp_t intra_mbcmp_x3_16x16, intra_satd_x3_16x16, intra_sad_x3_16x16;
    x264_intra_cmp_t intra_mbcmp_x3_4x4, intra_satd_x3_4x4, intra_sad_x3_4x4;
    x264_intra_cmp_t intra_mbcmp_x4_4x4_h, intra_satd_x4_4x4_h, intra_sad_x4_4x4_h;
    x264_intra_cmp_t intra_mbcmp_x4_4x4_v, intra_satd_x4_4x4_v, intra_sad_x4_4x4_v;
    x264_intra_cmp_t intra_mbcmp_x3_chroma, intra_satd_x3_chroma, intra_sad_x3_chroma;
    x264_intra_cmp_t intra_mbcmp_x3_8x8c, intra_satd_x3_8x8c, intra_sad_x3_8x8c;
    x264_intra_x9_t  intra_mbcmp_x9_4x4, intra_satd_x9_4x4, intra_sad_x9_4x4;
} x264_pixel_function_t;

/* ---- common/dct.h:8-33 */
typedef struct
{
    void (*sub4x4_dct)( dctcoef dct[16], pixel *pix1, pixel *pix2 );
    void (*add4x4_idct)( pixel *p_dst, dctcoef dct[16] );
    void (*sub8x8_dct)( dctcoef dct[4][16], pixel *pix1, pixel *pix2 );
    void (*sub8x8_dct_dc)( dctcoef dct[4], pixel *pix1, pixel *pix2 );
    void (*add8x8_idct)( pixel *p_dst, dctcoef dct[4][16] );
    void (*add8x8_idct_dc)( pixel *p_dst, dctcoef dct[4] );
    void (*sub16x16_dct)( dctcoef dct[16][16], pixel *pix1, pixel *pix2 );
    void (*add16x16_idct)( pixel *p_dst, dctcoef dct[16][16] );
    void (*add16x16_idct_dc)( pixel *p_dst, dctcoef dct[16] );
    void (*dct4x4dc)( dctcoef d[16] );
    void (*idct4x4dc)( dctcoef d[16] );
} x264_dct_function_t;

typedef struct
{
    void (*scan_4x4)( dctcoef level[16], dctcoef dct[16] );
} x264_zigzag_function_t;

/* ---- common/mc.h:31-79 */
typedef struct
{
    void (*mc_luma)( pixel *dst, intptr_t i_dst, pixel **src, intptr_t i_src, int mvx, int mvy,
                     int i_width, int i_height, const x264_weight_t *weight );
    pixel *(*get_ref)( pixel *dst, intptr_t *i_dst, pixel **src, intptr_t i_src, int mvx, int mvy,
                       int i_width, int i_height, const x264_weight_t *weight );
    void (*mc_chroma)( pixel *dstu, pixel *dstv, intptr_t i_dst, pixel *src, intptr_t i_src,
                       int mvx, int mvy, int i_width, int i_height );
    void (*copy[7])( pixel *dst, intptr_t dst_stride, pixel *src, intptr_t src_stride, int i_height );
    void (*store_interleave_chroma)( pixel *dst, intptr_t i_dst, pixel *srcu, pixel *srcv, int height );
    void (*load_deinterleave_chroma_fenc)( pixel *dst, pixel *src, intptr_t i_src, int height );
    void (*load_deinterleave_chroma_fdec)( pixel *dst, pixel *src, intptr_t i_src, int height );
    void (*plane_copy)( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int w, int h );
    void (*plane_copy_interleave)( pixel *dst, intptr_t i_dst, pixel *srcu, intptr_t i_srcu,
                                   pixel *srcv, intptr_t i_srcv, int w, int h );
    void (*plane_copy_deinterleave)( pixel *dstu, intptr_t i_dstu, pixel *dstv, intptr_t i_dstv,
                                     pixel *src, intptr_t i_src, int w, int h );
    void (*plane_copy_deinterlace)( pixel *srcy, intptr_t i_srcy, pixel *dsty, intptr_t i_dsty,
                                    pixel *srcc, intptr_t i_srcc, pixel *dstc, intptr_t i_dstc,
                                    int i_width, int i_height );
    void (*plane_deinterlace)( pixel *pixy, intptr_t i_pixy, pixel *pixc, intptr_t i_pixc, int i_width, int i_height );
    void (*hpel_filter)( pixel *dsth, pixel *dstv, pixel *dstc, pixel *src, intptr_t i_stride,
                         int i_width, int i_height, int16_t *buf );
    void (*prefetch_fenc)( pixel *pix_y, intptr_t stride_y, pixel *pix_uv, intptr_t stride_uv, int mb_x );
    void (*prefetch_fenc_420)( pixel *pix_y, intptr_t stride_y, pixel *pix_uv, intptr_t stride_uv, int mb_x );
    void (*prefetch_ref)( pixel *pix, intptr_t stride, int parity );
    void *(*memcpy_aligned)( void *dst, const void *src, size_t n );
    void (*memzero_aligned)( void *dst, size_t n );
    void (*frame_init_lowres_core)( pixel *src0, pixel *dst0, pixel *dsth, pixel *dstv, pixel *dstc,
                                    intptr_t src_stride, intptr_t dst_stride, int width, int height );
} x264_mc_functions_t;

/* ---- common/quant.h:8-31 */
typedef struct
{
    int  (*quant_4x4)( dctcoef dct[16], udctcoef mf[16], udctcoef bias[16] );
    int  (*quant_4x4_dc)( dctcoef dct[16], int mf, int bias );
    int  (*quant_2x2_dc)( dctcoef dct[4], int mf, int bias );
    void (*dequant_4x4)( dctcoef dct[16], int dequant_mf[6][16], int i_qp );
    void (*dequant_4x4_dc)( dctcoef dct[16], int dequant_mf[6][16], int i_qp );
    int  (*optimize_chroma_2x2_dc)( dctcoef dct[4], int dequant_mf );
    void (*denoise_dct)( dctcoef *dct, uint32_t *sum, udctcoef *offset, int size );
    int  (*decimate_score15)( dctcoef *dct );
    int  (*decimate_score16)( dctcoef *dct );
    int  (*coeff_last[14])( dctcoef *dct );
    int  (*coeff_last4)( dctcoef *dct );
    int  (*coeff_last8)( dctcoef *dct );
    int  (*coeff_level_run[13])( dctcoef *dct, x264_run_level_t *runlevel );
    int  (*coeff_level_run4)( dctcoef *dct, x264_run_level_t *runlevel );
    int  (*coeff_level_run8)( dctcoef *dct, x264_run_level_t *runlevel );
} x264_quant_function_t;

/* ---- common/frame.h:206-216 */
typedef void (*x264_deblock_inter_t)( pixel *pix, intptr_t stride, int alpha, int beta, int8_t *tc0 );
typedef void (*x264_deblock_intra_t)( pixel *pix, intptr_t stride, int alpha, int beta );
typedef struct
{
    x264_deblock_inter_t deblock_luma[2];          /* [0] vertical edge (filters across x), [1] horizontal edge */
    x264_deblock_inter_t deblock_chroma[2];
    x264_deblock_intra_t deblock_luma_intra[2];
    x264_deblock_intra_t deblock_chroma_intra[2];
    void (*deblock_strength)( uint8_t nnz[120], int8_t ref[2][40], int16_t mv[2][40][2], uint8_t bs[2][8][4] );
} x264_deblock_function_t;

#endif /* X264DSP_OWN_TABLE_TYPES */

void x264_pixel_init( int cpu, x264_pixel_function_t *pixf );
void x264_dct_init( int cpu, x264_dct_function_t *dctf );
void x264_zigzag_init( int cpu, x264_zigzag_function_t *zigzagf );
void x264_mc_init( int cpu, x264_mc_functions_t *pf );
void x264_quant_init( x264_t *h, int cpu, x264_quant_function_t *pf );
void x264_deblock_init( int cpu, x264_deblock_function_t *pf );

/* The intra predictor tables: x264_predict_t (common/predict.h:8) arrays indexed by the reference's
 * enums -- I_PRED_16x16_{V,H,DC,P,DC_LEFT,DC_TOP,DC_128} (predict.h:27-37), I_PRED_CHROMA_{DC,H,V,P,DC_LEFT,DC_TOP,
 * DC_128} (predict.h:10-20), I_PRED_4x4_{V,H,DC,DDL,DDR,VR,HD,VL,HU,DC_LEFT,DC_TOP,DC_128} (predict.h:44-59).  Each
 * entry predicts the block at src (FDEC_STRIDE 32) in place from its row above / column to the left, like
 * x264_predict_16x16_init / _8x8c_init / _4x4_init (common/predict.c:474-546; called at encoder/encoder.c:551-553). */
typedef void (*x264_predict_t)( pixel *src );
void x264_predict_16x16_init( int cpu, x264_predict_t pf[7] );
void x264_predict_8x8c_init( int cpu, x264_predict_t pf[7] );
void x264_predict_4x4_init( int cpu, x264_predict_t pf[12] );

/* the context the shims run on (created on first use); NULL if no CUDA device could be opened */
struct x264dsp_ctx;
struct x264dsp_ctx *x264dsp_tables_context( void );

#ifdef __cplusplus
}
#endif
#endif /* X264DSP_TABLES_H */
