/*
 * xo.h -- CPU oracle: a plain-C restatement of the x264-dsp hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libx264dsp_b200.so) never links,
 * calls or falls back to it.
 *
 * Parity status: PINNED.  Every function here is checked against the UNMODIFIED reference C path
 * compiled from /root/reference (oracle/_ref/libx264ref.so, built by `make -C oracle ref`) in
 * tests/test_oracle_vs_ref.py, and against golden vectors generated from that build and committed
 * under tests/golden/ (tests/golden/make_golden.py).  Not pinned by the reference (it does not
 * implement them): UMH/ESA/TESA search, subme 0, sub-8x8 partition DECISION.
 *
 * Each function cites the reference file:line it follows.  The code is written from the
 * behaviour, not transcribed: same integers out, different structure.
 *
 * Data layout is the product's (include/x264dsp_b200.h): frame slots, x264dsp_geom_t,
 * x264dsp_me_block_t ... -- only the POD types are shared, no code.
 */
#ifndef XO_H
#define XO_H

#include <stdint.h>
#include <stddef.h>
#include "../include/x264dsp_b200.h"

typedef uint8_t pixel_t;
typedef int16_t coef_t;

#define XO_FENC_STRIDE 16     /* common/common.h:871 */
#define XO_FDEC_STRIDE 32     /* common/common.h:872 */

/* ---- tables (encoder/analyse.c:98-111,171-315; common/set.c:265-353; common/macroblock.h:251-266) */
int  xo_lambda( int qp );
void xo_cost_mv_table( int qp, uint16_t out8193[8193] );
void xo_quant_tables( int b_inter, int qp, uint16_t mf[16], uint16_t bias[16] );
void xo_dequant_table( int out[6][16] );
int  xo_chroma_qp( int qp );
int  xo_lambda2( int qp );

/* ---- geometry + frame staging (common/frame.c) */
void xo_geometry( int width, int height, x264dsp_geom_t *g );
void xo_frame_load_i420( const x264dsp_geom_t *g, const uint8_t *i420, uint8_t *slot );
void xo_frame_expand_border( const x264dsp_geom_t *g, uint8_t *slot );
void xo_frame_filter( const x264dsp_geom_t *g, uint8_t *slot );
void xo_frame_init_lowres( const x264dsp_geom_t *g, uint8_t *slot );
void xo_frame_retile_lowres( const x264dsp_geom_t *g, uint8_t *slot );

/* ---- pixel metrics (common/pixel.c) */
int  xo_block_w( int size );
int  xo_block_h( int size );
int  xo_sad( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb );
int  xo_ssd( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb );
int  xo_satd( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb );
int  xo_cmp( int cmp, int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb );
uint64_t xo_var( int size, const pixel_t *p, intptr_t stride );
int  xo_var2_8x8( const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb, int *ssd );
void xo_cost_batch( int cmp, int n, const pixel_t *pix1, const int64_t *off1, int stride1,
                    const pixel_t *pix2, const int64_t *off2, int stride2,
                    const uint8_t *size, int32_t *out );
/* x264_predict_8x8c_{dc,h,v}_c; mode 0 = DC, 1 = H, 2 = V; src at FDEC stride with neighbours loaded */
void xo_predict_8x8c( int mode, pixel_t *src );
/* x264_intra_{sad,satd}_x3_8x8c: res order DC,H,V; overwrites the 8x8 at fdec */
void xo_intra_x3_8x8c( int use_satd, const pixel_t *fenc, pixel_t *fdec, int res[3] );

/* ---- motion compensation (common/mc.c) */
void xo_hpel_filter( pixel_t *dsth, pixel_t *dstv, pixel_t *dstc, const pixel_t *src,
                     intptr_t stride, int width, int height );
void xo_mc_luma( pixel_t *dst, intptr_t dst_stride, const pixel_t *const src[4], intptr_t src_stride,
                 int mvx, int mvy, int w, int h );
/* returns the pointer the reference's get_ref would return; *dst_stride is rewritten likewise */
const pixel_t *xo_get_ref( pixel_t *dst, intptr_t *dst_stride, const pixel_t *const src[4],
                           intptr_t src_stride, int mvx, int mvy, int w, int h );
void xo_mc_chroma( pixel_t *dstu, pixel_t *dstv, intptr_t dst_stride, const pixel_t *src,
                   intptr_t src_stride, int mvx, int mvy, int w, int h );
void xo_lowres_core( const pixel_t *src0, pixel_t *dst0, pixel_t *dsth, pixel_t *dstv, pixel_t *dstc,
                     intptr_t src_stride, intptr_t dst_stride, int width, int height );

/* ---- transform + quant (common/dct.c, common/quant.c) */
void xo_sub4x4_dct( coef_t dct[16], const pixel_t *fenc, const pixel_t *fdec );
void xo_sub8x8_dct( coef_t dct[4][16], const pixel_t *fenc, const pixel_t *fdec );
void xo_sub16x16_dct( coef_t dct[16][16], const pixel_t *fenc, const pixel_t *fdec );
void xo_sub8x8_dct_dc( coef_t dc[4], const pixel_t *fenc, const pixel_t *fdec );
void xo_add4x4_idct( pixel_t *fdec, const coef_t dct[16] );
void xo_add8x8_idct( pixel_t *fdec, coef_t dct[4][16] );
void xo_add16x16_idct( pixel_t *fdec, coef_t dct[16][16] );
void xo_add8x8_idct_dc( pixel_t *fdec, const coef_t dc[4] );
void xo_add16x16_idct_dc( pixel_t *fdec, const coef_t dc[16] );
void xo_dct4x4dc( coef_t d[16] );
void xo_idct4x4dc( coef_t d[16] );
void xo_zigzag_4x4( coef_t level[16], const coef_t dct[16] );
int  xo_quant_4x4( coef_t dct[16], const uint16_t mf[16], const uint16_t bias[16] );
int  xo_quant_4x4_dc( coef_t dct[16], int mf, int bias );
int  xo_quant_2x2_dc( coef_t dct[4], int mf, int bias );
void xo_dequant_4x4( coef_t dct[16], int dequant_mf[6][16], int qp );
void xo_dequant_4x4_dc( coef_t dct[16], int dequant_mf[6][16], int qp );
int  xo_optimize_chroma_2x2_dc( coef_t dct[4], int dmf );
int  xo_decimate_score15( const coef_t *level );
int  xo_decimate_score16( const coef_t *level );
int  xo_coeff_last( const coef_t *level, int n );

/* ---- deblock (common/deblock.c) */
void xo_deblock_luma( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta, const int8_t tc0[4] );
void xo_deblock_chroma( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta, const int8_t tc0[4] );
void xo_deblock_luma_intra( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta );
void xo_deblock_chroma_intra( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta );
void xo_macroblock_deblock_strength( int n, const int8_t *mb_type, const uint8_t *nnz, const int8_t *ref,
                                     const int16_t *mv, uint8_t *bs );
void xo_deblock_strength( int n, const uint8_t *nnz, const int8_t *ref, const int16_t *mv, uint8_t *bs );
void xo_deblock_frame( const x264dsp_geom_t *g, uint8_t *slot, const int8_t *mb_type,
                       const uint8_t *partition, const int16_t *cbp, const uint8_t *bs,
                       int qp, int alpha_c0_offset, int beta_offset );

/* ---- motion search (encoder/me.c) */
void xo_me_search_batch_ex( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                            const x264dsp_me_params_t *prm, int n, const x264dsp_me_block_t *blocks,
                            x264dsp_me_result_t *results, int mode, int32_t *thresh );
void xo_me_search_batch( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                         const x264dsp_me_params_t *params, int n, const x264dsp_me_block_t *blocks,
                         x264dsp_me_result_t *results );

/* ---- lowres lookahead (encoder/slicetype.c) */
void xo_lookahead_frame_cost( const x264dsp_geom_t *g, const uint8_t *slot_b, const uint8_t *slot_p0,
                              int want_intra, int16_t *mvs, int32_t *costs, int32_t *sums,
                              int32_t *row_satds );

/* ---- residual + MC (encoder/macroblock.c, common/macroblock.c) */
void xo_mc_frame( const x264dsp_geom_t *g, const uint8_t *fref_slot, const int16_t *mv, uint8_t *pred_slot );
void xo_mc_frame_part( const x264dsp_geom_t *g, const uint8_t *fref_slot, const int16_t *mv, uint8_t *pred_slot );
void xo_residual_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                        int16_t *levels, uint8_t *nnz, int16_t *cbp );
void xo_residual_frame_typed( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                              const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels, int16_t *luma_dc,
                              uint8_t *nnz, int16_t *cbp );
int xo_probe_pskip_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, const pixel_t *fdec_y, const pixel_t *fdec_c, int qp );
void xo_probe_pskip_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *pred_slot, int qp, uint8_t *skip );
void xo_predict_mv_16x16( const x264dsp_mv_neighbours_t *nb, int i_ref, int16_t mvp[2] );
void xo_predict_mv_pskip( const x264dsp_mv_neighbours_t *nb, int16_t mv[2] );
void xo_predict_mvc_16x16_frame( int mb_w, int mb_h, const int16_t *lowres_mv, const int16_t *mvr, const int16_t *l0_mv16,
                                 int scale, int16_t *mvc, int32_t *n_mvc );
void xo_predict_mv_part( const x264dsp_mv_neighbours_t *nb, int i_ref, int shape, int c_unreachable, int16_t mvp[2] );
void xo_predict_4x4( int mode, pixel_t *src );
int xo_encode_luma_i4x4( const pixel_t *fenc, pixel_t *fdec, int qp, const uint8_t modes[16], int replicate5,
                         int16_t *levels, uint8_t *nnz );
int xo_encode_intra4_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                         int qp, const uint8_t modes[16], int replicate5, int16_t *levels, uint8_t *nnz );
/* x264_macroblock_encode, inter branch, on one macroblock in fenc_buf / fdec_buf form; returns h->mb.cbp (CABAC packing) */
int xo_encode_inter_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                        int qp, int16_t *levels, uint8_t *nnz );
int xo_encode_intra16_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                          int qp, int16_t *levels, int16_t *luma_dc, uint8_t *nnz );

/* the P-slice macroblock loop: x264_macroblock_analyse + x264_macroblock_encode for every macroblock (xo_pframe.c) */
void xo_p_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot, uint8_t *recon_slot,
                 const x264dsp_pframe_params_t *prm, const int16_t *lowres_mv, const int16_t *l0_mv16,
                 int8_t *mb_type, int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *levels, uint8_t *nnz, int16_t *cbp );

/* the same with every partition the reference analyses (analyse.inter = PSUB16x16): vectors per 8x8 block */
void xo_p_frame_part( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot, uint8_t *recon_slot,
                      const x264dsp_pframe_params_t *prm, const int16_t *lowres_mv, const int16_t *l0_mv16,
                      int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8, int16_t *levels,
                      uint8_t *nnz, int16_t *cbp );

/* the I-slice macroblock loop: intra analysis + coding of every macroblock (xo_iframe.c) */
void xo_predict_16x16( int mode, pixel_t *src );
void xo_predict_chroma( int mode, pixel_t *src );
void xo_i_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *recon_slot, int qp, int8_t *mb_type,
                 uint8_t *mode16, uint8_t *chroma_mode, uint8_t *modes4, int16_t *levels, int16_t *luma_dc, uint8_t *nnz,
                 int16_t *cbp );

/* work counters of the last xo_me_search_batch / xo_lookahead_frame_cost call on this thread:
 * counts[0] = pixel comparisons done by SAD, counts[1] = by SATD, [2] = SAD calls, [3] = SATD calls */
void xo_work_counters( int64_t counts[4], int reset );

#endif
