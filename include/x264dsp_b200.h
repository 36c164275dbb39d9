/*
 * x264dsp_b200.h -- C ABI of the B200-native x264-dsp hot path.
 *
 * One shared library, libx264dsp_b200.so (CUDA, sm_100a), stands behind two headers:
 *
 *   x264dsp_tables.h  the reference's own function-pointer tables
 *                     (x264_pixel_function_t, x264_dct_function_t, x264_zigzag_function_t,
 *                     x264_mc_functions_t, x264_quant_function_t, x264_deblock_function_t) and
 *                     their x264_*_init entry points, same layouts and signatures, so the
 *                     reference can be linked against this library unchanged for that path;
 *   x264dsp_b200.h    (this file) frame-batched entry points with identical per-block /
 *                     per-macroblock semantics -- the calls that are worth a kernel launch.
 *
 * Conventions
 *   - plain C, no CUDA or torch types.  `stream` arguments are a cudaStream_t passed as void*
 *     (NULL = the context's own stream).
 *   - *_dev functions take DEVICE pointers and only enqueue work on the stream; the *_host
 *     functions take HOST pointers, stage through the context's pinned buffers, and return
 *     when the results are in the caller's memory.
 *   - return value: 0 = ok, <0 = bad argument (X264DSP_E_*), >0 = cudaError_t of the failing call.
 *   - there is NO CPU fallback.  If no CUDA device can be opened, x264dsp_create fails and
 *     every other entry point refuses to run.
 *   - all arithmetic is integer; every output is bit-exact with the reference's portable C path.
 *
 * Each declaration cites the reference interface (file:line under the x264-dsp tree) it replaces.
 */
#ifndef X264DSP_B200_H
#define X264DSP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X264DSP_E_ARG     (-1)   /* NULL / out-of-range argument */
#define X264DSP_E_NOGPU   (-2)   /* no usable CUDA device (there is no CPU path) */
#define X264DSP_E_NOMEM   (-3)

/* block sizes: same numbering as the reference's enum (common/pixel.h:14-36) */
enum
{
    X264DSP_PIXEL_16x16 = 0, X264DSP_PIXEL_16x8 = 1, X264DSP_PIXEL_8x16 = 2, X264DSP_PIXEL_8x8 = 3,
    X264DSP_PIXEL_8x4 = 4, X264DSP_PIXEL_4x8 = 5, X264DSP_PIXEL_4x4 = 6, X264DSP_PIXEL_4x16 = 7
};

enum { X264DSP_CMP_SAD = 0, X264DSP_CMP_SSD = 1, X264DSP_CMP_SATD = 2 };
/* common/x264.h:117-121.  DIA and HEX are the two pattern searches of this reference; UMH, ESA and TESA pass its
 * parameter check (encoder/encoder.c:251-259) but have no case in x264_me_search_ref's switch (encoder/me.c:389-394),
 * so they evaluate the predictors and go straight to the sub-pel refinement -- which is what this library does too;
 * TESA with subme >= 2 also makes the full-pel metric SATD (mbcmp_init, encoder/encoder.c:429-432). */
enum { X264DSP_ME_DIA = 0, X264DSP_ME_HEX = 1, X264DSP_ME_UMH = 2, X264DSP_ME_ESA = 3, X264DSP_ME_TESA = 4 };

#define X264DSP_PADH 32                                   /* common/frame.h:9-10 */
#define X264DSP_PADV 32
#define X264DSP_LOOKAHEAD_QP 12                           /* common/common.h:46 */

/* ------------------------------------------------------------------ geometry
 * Frame layout of the reference (common/frame.c:22-57, 77-97, 126-133), reproduced exactly so
 * that planes can be compared byte for byte including padding.
 *
 * One "frame slot" in HBM is a single allocation:
 *     [ luma N | luma H | luma V | luma HV ]   4 x luma_plane_size      (frame.c:83-92)
 *     [ chroma NV12 ]                          chroma_plane_size        (frame.c:78-80)
 *     [ lowres N | H | V | HV ]                4 x lowres_plane_size    (frame.c:129-132)
 * luma_origin / chroma_origin / lowres_origin are the byte offsets of sample (0,0) inside a plane.
 */
typedef struct x264dsp_geom
{
    int32_t width, height;            /* picture size */
    int32_t mb_w, mb_h, mb_count;     /* 16x16 macroblock grid */
    int32_t luma_w, luma_h;           /* mb_w*16, mb_h*16 */
    int32_t luma_stride, luma_plane_size, luma_origin;
    int32_t chroma_stride, chroma_h, chroma_plane_size, chroma_origin;
    int32_t lowres_w, lowres_h, lowres_stride, lowres_plane_size, lowres_origin;
    int32_t slot_chroma_off, slot_lowres_off;
    int64_t slot_bytes;               /* bytes of one frame slot, multiple of 256 */
    /* 8x8-tiled copies of the four padded lowres planes (the lookahead's search layout): tile (tx,ty) of
     * the padded plane = 64 contiguous bytes (8 rows of 8 samples), tiles of a tile row back to back;
     * sample (x,y) of plane k lives at slot_tiled_off + k*tiled_plane_size
     *   + (((y+32)>>3) * tile_w + ((x+32)>>3)) * 64 + ((y+32)&7) * 8 + ((x+32)&7).
     * Written by x264dsp_frame_init_lowres_dev together with the row-major planes. */
    int32_t tile_w, tile_h, tiled_plane_size, slot_tiled_off;
} x264dsp_geom_t;

int x264dsp_geometry( int width, int height, x264dsp_geom_t *g );

/* ------------------------------------------------------------------ context */
typedef struct x264dsp_ctx x264dsp_ctx_t;

/* opens CUDA device `device`, creates a stream, uploads the constant tables (cost_mv for every
 * distinct lambda: encoder/analyse.c:243-315; flat-CQM quant/dequant tables: common/set.c:265-353) */
int  x264dsp_create( int device, x264dsp_ctx_t **out );
void x264dsp_destroy( x264dsp_ctx_t *ctx );
/* the context's stream as a cudaStream_t (so a caller can record events on it) */
void *x264dsp_stream( x264dsp_ctx_t *ctx );
int  x264dsp_sync( x264dsp_ctx_t *ctx );
/* number of kernel launches this context has enqueued so far (bench.py's gpu_launches) */
int64_t x264dsp_launch_count( const x264dsp_ctx_t *ctx );
const char *x264dsp_version( void );

/* host copies of the constant tables, for callers and tests
 * (encoder/analyse.c:98-111, 171-206, 243-315; common/set.c:265-353; common/macroblock.h:251-266) */
int x264dsp_lambda( int qp );
int x264dsp_cost_mv_table( int qp, uint16_t out8193[8193] );           /* index i+4096 <-> mv delta i */
/* the context's DEVICE copy of cost_mv[qp] (the table the search kernels index), read back to the host */
int x264dsp_cost_mv_table_dev( x264dsp_ctx_t *ctx, int qp, uint16_t out8193[8193] );
int x264dsp_quant_tables( int b_inter, int qp, uint16_t mf[16], uint16_t bias[16] );
int x264dsp_dequant_table( int out[6][16] );
int x264dsp_chroma_qp( int qp );

/* device memory helpers so that a plain-C caller needs no CUDA headers */
int x264dsp_dev_alloc( x264dsp_ctx_t *ctx, size_t bytes, void **dev );
int x264dsp_dev_free( x264dsp_ctx_t *ctx, void *dev );
int x264dsp_dev_zero( x264dsp_ctx_t *ctx, void *dev, size_t bytes, void *stream );
int x264dsp_h2d( x264dsp_ctx_t *ctx, void *dev, const void *host, size_t bytes, void *stream );
int x264dsp_d2h( x264dsp_ctx_t *ctx, void *host, const void *dev, size_t bytes, void *stream );

/* pinned host memory: buffers obtained here are copied from / to directly by the *_host entry points */
int x264dsp_host_alloc( x264dsp_ctx_t *ctx, size_t bytes, void **host );
int x264dsp_host_free( x264dsp_ctx_t *ctx, void *host );
/* ... or page-lock memory the caller already owns (the reference's x264_frame_t buffers: glue/x264dsp_glue.c) */
int x264dsp_host_register( x264dsp_ctx_t *ctx, void *host, size_t bytes );
int x264dsp_host_unregister( x264dsp_ctx_t *ctx, void *host );

/* per-kernel timing with CUDA events recorded around each launch on its own stream.
 * kind: one of X264DSP_PROF_*; total_ms / count cover the launches since x264dsp_profile_enable. */
enum
{
    X264DSP_PROF_LOAD = 0, X264DSP_PROF_LOWRES, X264DSP_PROF_LA_INTRA, X264DSP_PROF_LA_INTER, X264DSP_PROF_HPEL,
    X264DSP_PROF_BORDER, X264DSP_PROF_COST, X264DSP_PROF_ME, X264DSP_PROF_MC, X264DSP_PROF_RESIDUAL,
    X264DSP_PROF_DEBLOCK, X264DSP_PROF_LA_TILE, X264DSP_PROF_KINDS
};
int x264dsp_profile_enable( x264dsp_ctx_t *ctx, int on );
int x264dsp_profile_read( x264dsp_ctx_t *ctx, int kind, double *total_ms, int *count );

/* ------------------------------------------------------------------ synthetic input
 * Seeded synthetic YUV 4:2:0: panning blurred-noise texture, two moving gradient
 * squares, +-2 noise, scene cut at `cut_frame` (<0: none).  Host code; writes planar I420. */
int x264dsp_synth_frame( int width, int height, int frame_no, int cut_frame,
                         uint8_t *y, uint8_t *u, uint8_t *v );

/* ------------------------------------------------------------------ frame staging (8(f) N4)
 * x264_frame_copy_picture for I420 input + x264_frame_expand_border_mod16
 * (common/frame.c:198-232, 423-450): planar Y,U,V -> padded luma plane N + NV12 chroma plane.
 * i420 holds n_frames consecutive pictures (Y then U then V, tightly packed).
 * slots: n_frames frame slots. */
int x264dsp_frame_load_i420_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *i420,
                                 uint8_t *slots, int n_frames, void *stream );

/* luma-only staging for the lookahead path, which never touches chroma: n_frames pictures of
 * width*height bytes -> padded luma plane N of each slot (plane_copy + mod16 padding,
 * common/frame.c:227, 435-448). */
int x264dsp_frame_load_luma_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                 uint8_t *slots, int n_frames, void *stream );

/* x264_frame_expand_border for every MB row (common/frame.c:386-396): replicate luma N and chroma
 * into their 32 / 16 sample padding. */
int x264dsp_frame_expand_border_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                     int n_frames, void *stream );

/* x264_frame_filter + x264_frame_expand_border_filtered over the whole frame
 * (common/mc.c:144-167, 506-535; common/frame.c:398-413): H, V, HV half-pel planes from plane N
 * (whose border must already be expanded), padded from the last filtered sample. */
int x264dsp_frame_filter_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                              int n_frames, void *stream );

/* x264_frame_init_lowres (common/mc.c:404-456 + common/frame.c:415-421): duplicates the last
 * column/row INTO the source luma plane, builds the four half-resolution planes and pads them.
 * The planes are kept in the slot's TILED form (see x264dsp_geom_t), which is what every lookahead
 * entry point reads; x264dsp_frame_export_lowres_dev writes the reference's row-major lowres[0..3]
 * (padding included) into the slot's lowres region when a caller wants them. */
int x264dsp_frame_init_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                   int n_frames, void *stream );
/* x264dsp_frame_load_luma_dev + x264dsp_frame_init_lowres_dev in one pass over the picture (same final
 * slot contents: luma plane N incl. the duplicated column / row, four padded lowres planes) */
int x264dsp_frame_load_luma_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                        uint8_t *slots, int n_frames, void *stream );
/* x264_frame_init_lowres with the caller's width x height picture AS frame->plane[0]: only the (tiled)
 * lowres planes of the slot are written, its luma plane is left alone.  For callers that want nothing
 * but the lookahead of these pictures (x264dsp_lookahead_clips_host uses it internally). */
int x264dsp_frame_lowres_from_luma_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                        uint8_t *slots, int n_frames, void *stream );
/* tiled -> row-major: fills the slot's lowres region (lowres[0..3] of x264_frame_t, padding included) */
int x264dsp_frame_export_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                     int n_frames, void *stream );
/* row-major -> tiled: for a caller that has written row-major lowres planes into the slot's lowres
 * region itself (e.g. planes computed by the reference) and wants the lookahead to use them */
int x264dsp_frame_retile_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                     int n_frames, void *stream );

/* ------------------------------------------------------------------ block costs
 * x264_pixel_function_t::sad / ssd / satd  (common/pixel.c:44-102, 267-337) on n independent
 * block pairs.  Block i compares  pix1 + off1[i] (stride1)  with  pix2 + off2[i] (stride2);
 * size[i] is an X264DSP_PIXEL_* value.  off1/off2/size/out are device arrays.
 * The kernels assemble unaligned rows from aligned 32-bit words and always fetch the word after the last one they need:
 * pix1 and pix2 must have at least 4 readable bytes after the last byte of the last block (frame slots do: they are padded). */
int x264dsp_cost_batch_dev( x264dsp_ctx_t *ctx, int cmp, int n,
                            const uint8_t *pix1, const int64_t *off1, int stride1,
                            const uint8_t *pix2, const int64_t *off2, int stride2,
                            const uint8_t *size, int32_t *out, void *stream );

/* ------------------------------------------------------------------ lowres lookahead
 * x264_slicetype_frame_cost / x264_slicetype_mb_cost (encoder/slicetype.c:48-322) with the
 * reference's defaults (do_edges = 0, no B frames, lookahead QP 12, DIA + subme 2).
 *
 * b / p0 / want_intra are small HOST arrays describing the batch; everything else is device memory.
 * pair p analyses frame slot b[p] against reference slot p0[p] (P frame, p1 == b).
 * p0[p] < 0 means intra only (the reference's frame_cost(b,b,b) call).
 * want_intra[p] != 0 also produces the intra estimate (first analysis of that frame,
 * slicetype.c:145-180).
 *
 * outputs per pair:
 *   mvs   [p][mb_count][2] int16   lowres_mvs[0][0]        (border blocks stay 0)
 *   costs [p][mb_count]    int32   lowres_mv_costs[0][0]   (border blocks stay 0)
 *   sums  [p][X264DSP_LA_SUMS]     see enum below
 *   row_satds [p][2][mb_h] int32   inter / intra row sums (i_row_satds; may be NULL)
 */
enum
{
    X264DSP_LA_COST_INTER = 0,   /* i_cost_est[b-p0][0] */
    X264DSP_LA_COST_INTRA = 1,   /* i_cost_est[0][0] (valid when want_intra) */
    X264DSP_LA_INTRA_MBS  = 2,   /* i_intra_mbs[b-p0] */
    X264DSP_LA_SAD_EVALS  = 3,   /* work counters: 8x8 SAD evaluations the reference would issue */
    X264DSP_LA_SATD_EVALS = 4,   /* 8x8 SATD evaluations the reference would issue */
    X264DSP_LA_SUMS       = 8
};
int x264dsp_lookahead_frame_cost_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *slots,
                                      int n_pairs, const int32_t *b, const int32_t *p0,
                                      const uint8_t *want_intra,
                                      int16_t *mvs, int32_t *costs, int32_t *sums, int32_t *row_satds,
                                      void *stream );

/* Whole lookahead pass from HOST memory: n_clips independent clips of clip_len planar luma pictures
 * (width*height bytes each; chroma is not used by the lookahead) -> per-frame results in host
 * arrays laid out like the _dev outputs ([frame][mb_count]...).  The first frame of every clip is
 * analysed intra-only, every other frame against its predecessor.  This is the call bench.py
 * times end to end (H2D of the pictures and D2H of the results inside).  Groups of clips run on
 * separate streams so that copies and kernels overlap. */
int x264dsp_lookahead_clips_host( x264dsp_ctx_t *ctx, int width, int height, int n_clips, int clip_len,
                                  const uint8_t *luma, int16_t *mvs, int32_t *costs, int32_t *sums );
/* The inter search has two mappings with identical results: a warp per block row (shortest dependency
 * chain, best for a handful of frame pairs) and a warp per four or eight block rows (a quarter of the
 * instructions, best once a launch holds a few dozen pairs).  mode 0 = choose by batch size (default),
 * 1 = one row per warp, 2 = four rows per warp, 3 = eight rows per warp. */
int x264dsp_lookahead_select_kernel( x264dsp_ctx_t *ctx, int mode );
/* measurement aid: on != 0 makes x264dsp_lookahead_clips_host issue its copies (same buffers, streams and order) but none
 * of its kernels -- what the host<->device DMA alone costs; results are meaningless while it is on (bench.py e2e.copy_only) */
int x264dsp_debug_copies_only( x264dsp_ctx_t *ctx, int on );
/* debug aid: clock64() cycles the inter kernel's warps spent per phase, summed over all warps since the
 * last reset: out[0..9] = waiting on the row below, block setup, zero-mv SATD probe, predictor
 * candidates, diamond search, sub-pel refine + SATD, publish/accounting, blocks processed, first wait
 * of each row, rows processed. */
int x264dsp_debug_lookahead_timing( x264dsp_ctx_t *ctx, int enable, int reset, uint64_t out[10] );
/* single clip: same as n_clips = 1 */
int x264dsp_lookahead_clip_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames,
                                 const uint8_t *luma, int16_t *mvs, int32_t *costs, int32_t *sums );

/* ------------------------------------------------------------------ multi-GPU sharding
 * The path shards by frame range with no exchange step (each frame's lookahead / ME costs depend
 * only on source frames).  Rank `rank` of `world` owns frames [first, first + count) of an
 * n_frames sequence, contiguous and balanced to within one frame; to analyse its first frame as a
 * P frame it must also LOAD frame first-1, which *need_prev reports (0 for the rank that owns
 * frame 0, and for empty ranges).  Returns X264DSP_E_ARG for a bad rank / world. */
int x264dsp_frame_range( int n_frames, int rank, int world, int *first, int *count, int *need_prev );

/* ------------------------------------------------------------------ motion search
 * x264_me_search_ref (+ optional x264_me_refine_qpel) (encoder/me.c:129-435) on n independent
 * blocks of one source frame against one reference frame's N/H/V/HV planes. */
typedef struct x264dsp_me_block
{
    int32_t i_pixel;                       /* x264_me_t::i_pixel */
    int32_t bx, by;                        /* luma position of the block */
    int16_t mvp[2];                        /* x264_me_t::mvp */
    int32_t i_mvc;                         /* number of candidates, 0..16 */
    int16_t mvc[16][2];
    int32_t mv_min_fpel[2], mv_max_fpel[2];   /* h->mb.mv_{min,max}_fpel */
    int32_t mv_min_spel[2], mv_max_spel[2];   /* h->mb.mv_{min,max}_spel */
} x264dsp_me_block_t;

typedef struct x264dsp_me_result
{
    int16_t mv[2];                         /* quarter-pel */
    int32_t cost;
    int32_t cost_mv;
} x264dsp_me_result_t;

typedef struct x264dsp_me_params
{
    int32_t me_method;                     /* h->mb.i_me_method: X264DSP_ME_DIA .. _TESA */
    int32_t subpel_refine;                 /* h->mb.i_subpel_refine, 1..5 */
    int32_t me_range;                      /* h->param.analyse.i_me_range */
    int32_t qp;                            /* selects cost_mv[qp] */
    int32_t refine_qpel;                   /* also run x264_me_refine_qpel on each result */
} x264dsp_me_params_t;

int x264dsp_me_search_batch_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                 const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                 const x264dsp_me_params_t *params, int n,
                                 const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                 void *stream );

/* The other two callers of refine_subpel, and x264_me_search_ref with a non-NULL p_halfpel_thresh
 * (encoder/me.h:40-44, me.c:129, 426-440, 526-539; used by x264_mb_analyse_inter_p16x16 when several references are
 * searched, encoder/analyse.c:792-820):
 *   mode X264DSP_ME_MODE_SEARCH       x264_me_search_ref( h, m, mvc, i_mvc, p_halfpel_thresh )
 *   mode X264DSP_ME_MODE_REFDUPE      x264_me_refine_qpel_refdupe( h, m, p_halfpel_thresh ): results[i] holds m->mv,
 *                                     m->cost and m->cost_mv on entry and the refined values on return
 *   mode X264DSP_ME_MODE_REFINE_QPEL  x264_me_refine_qpel( h, m ) alone, results[i] in/out; the caller has already
 *                                     subtracted m->i_ref_cost for sizes <= 8x8 (me.c:431-432)
 * halfpel_thresh: device array [n], *p_halfpel_thresh of each block, read and updated (NULL = the reference's NULL).
 * When the early exit of me.c:529-536 fires, mv and cost are stored and cost_mv keeps its previous value. */
enum { X264DSP_ME_MODE_SEARCH = 0, X264DSP_ME_MODE_REFDUPE = 1, X264DSP_ME_MODE_REFINE_QPEL = 2 };
int x264dsp_me_search_batch_ex_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                    const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                    const x264dsp_me_params_t *params, int n,
                                    const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                    int mode, int32_t *halfpel_thresh, void *stream );

/* Same search for a list whose blocks all have partition size i_pixel (blocks[k].i_pixel is ignored):
 * the kernel is specialised on the size and packs 4 .. 32 blocks into a warp (one lane per 8x4 / 4x4
 * Hadamard tile), which is several times faster than the generic entry point on small partitions.
 * The reference's analysis visits one partition type at a time (encoder/analyse.c:787-1232), so its
 * call sites map onto uniform lists. */
int x264dsp_me_search_sized_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                 const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                 const x264dsp_me_params_t *params, int i_pixel, int n,
                                 const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                 void *stream );
/* the same for n_frames frame pairs in one launch: pair f searches fenc_slots + f*slot_bytes in
 * fref_slots + f*slot_bytes with blocks[f*n .. f*n+n) -> results[f*n ..).  One 1080p frame of 16x16
 * blocks is about half a wave on 148 SMs; a batch fills the machine. */
int x264dsp_me_search_sized_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                        const uint8_t *fenc_slots, const uint8_t *fref_slots, int n_frames,
                                        const x264dsp_me_params_t *params, int i_pixel, int n,
                                        const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                        void *stream );

/* The same from HOST memory (SURVEY 8(d) config 3 end to end): luma holds n_pairs + 1 pictures of width x height bytes,
 * frame f + 1 is searched in frame f, whose border expansion and half-pel planes (x264_frame_expand_border,
 * x264_frame_filter) are built on the device.  For each of the n_sizes partition sizes i_pixel[s], blocks[s] is a host
 * array of n_pairs x n_blocks[s] block descriptions (pair-major) and results[s] receives as many results.  Copies of
 * the pictures, the block lists and the results are inside the call; the sizes run on separate streams. */
int x264dsp_me_search_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_pairs, const uint8_t *luma,
                                   const x264dsp_me_params_t *params, int n_sizes, const int32_t *i_pixel,
                                   const int32_t *n_blocks, const x264dsp_me_block_t *const *blocks,
                                   x264dsp_me_result_t *const *results );

/* ------------------------------------------------------------------ residual
 * The inter-macroblock branch of x264_macroblock_encode + x264_mb_encode_chroma
 * (encoder/macroblock.c:175-305, 379-471) for every macroblock of a frame, b_dct_decimate = 1,
 * CABAC cbp packing, one slice QP.  pred_slot holds the motion-compensated prediction
 * (luma plane N + NV12 chroma) and is updated IN PLACE to the reconstruction, as p_fdec is.
 *   levels [mb][24+... ] see X264DSP_RES_*   zig-zagged quantised levels (h->dct.luma4x4 / chroma_dc)
 *   nnz    [mb][48+3] uint8                  non_zero_count per 4x4 (coding order) + luma DC + 2 chroma DC
 *   cbp    [mb] int16                        h->mb.cbp
 */
#define X264DSP_RES_LEVELS_PER_MB (16*16 + 2*4 + 2*4*16)   /* luma 4x4s, chroma DC u/v, chroma 4x4s */
#define X264DSP_RES_NNZ_PER_MB    (16 + 8 + 3)             /* luma, chroma u/v, luma DC, chroma DC u/v */
int x264dsp_residual_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                                int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream );
/* the same for n_frames consecutive slots in one launch; levels / nnz / cbp hold n_frames x mb_count
 * macroblocks back to back (one frame per launch is launch- and tail-bound, a batch streams) */
int x264dsp_residual_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                 const uint8_t *fenc_slots, uint8_t *pred_slots, int n_frames, int qp,
                                 int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream );

/* Macroblocks of two kinds in one launch: mb_kind[frame][mb] = 0 codes the macroblock as above (inter, P slice),
 * 1 codes it as an I16x16 macroblock of an I slice -- x264_mb_encode_i16x16 (encoder/macroblock.c:72-162: AC blocks
 * with the CQM_4IY tables, the sixteen DC terms through dct4x4dc / quant_4x4_dc / idct4x4dc / dequant_4x4_dc,
 * add16x16_idct or add16x16_idct_dc) and x264_mb_encode_chroma with b_inter = 0 (macroblock.c:175-305), both with
 * h->mb.b_dct_decimate = 0 as in an I slice.  pred_slot holds the intra prediction the caller chose
 * (h->predict_16x16[mode] / h->predict_chroma[mode] output), reconstruction in place.
 *   luma_dc [frame][mb][16] int16   zig-zagged levels of the luma DC block (h->dct.luma16x16_dc), zero for kind 0;
 *                                   may be NULL.  nnz[24] and bit 8 of cbp carry its non-zero flag.
 * mb_kind = 2 codes an I4x4 macroblock of an I slice (I_4x4 branch of x264_macroblock_encode, encoder/macroblock.c:
 * 355-377, + x264_mb_encode_i4x4, encoder/macroblock.h:37-61): i4_modes[frame][mb][16] holds the sixteen I_PRED_4x4_*
 * modes (h->mb.cache.intra4x4_pred_mode, coding order); each block is predicted from the RECONSTRUCTION around it, so
 * pred_slot must hold the final reconstruction of the macroblock's left, top-left, top and top-right neighbours (their
 * row / column next to the macroblock) -- no such neighbour may be an I4x4 macroblock of the same launch; a wavefront
 * caller launches per anti-diagonal.  Add 4 to the kind (6) when the macroblock has a row above it but no top-right
 * macroblock (block 5 then replicates its last top sample, macroblock.c:372-374).  The chroma prediction is the
 * caller's, as for kind 1.
 * mb_kind == NULL codes every macroblock as kind 0; i4_modes may be NULL when no macroblock has kind 2. */
int x264dsp_residual_frames_typed_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                       const uint8_t *fenc_slots, uint8_t *pred_slots, int n_frames, int qp,
                                       const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels,
                                       int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, void *stream );

/* ------------------------------------------------------------------ MV prediction (8(f) N2)
 * x264_mb_predict_mv_16x16 (common/mvpred.c:101-137) and x264_mb_predict_mv_pskip (mvpred.c:139-155) for n
 * macroblocks: the median / single-match / left-only rules over the neighbours h->mb.cache holds around X264_SCAN8_0,
 * and the P_SKIP vector (zero when a neighbour is missing or is a zero-vector reference-0 block).
 *   ref[k], mv[k]: A = left, B = top, C = top-right, D = top-left; ref -2 = not available, -1 = intra, >= 0 = index.
 * mvp[i] is the prediction for reference i_ref[i] (i_ref == NULL: reference 0 everywhere); either output may be NULL.
 * shape[i] (NULL: all 0) selects the partition rule of x264_mb_predict_mv (mvpred.c:22-99): 0 = 16x16 or 8x8 (the
 * rules above), 1 / 2 = upper / lower 16x8 (B resp. A wins outright when it uses i_ref), 3 / 4 = left / right 8x16 (A
 * resp. C); + 8 when the partition's top-right block comes later in scan order, so that D stands in for C.  The
 * neighbours are then those of the partition (the cells around x264_scan8[idx]). */
typedef struct x264dsp_mv_neighbours
{
    int8_t  ref[4];
    int16_t mv[4][2];
} x264dsp_mv_neighbours_t;
int x264dsp_predict_mv_batch_dev( x264dsp_ctx_t *ctx, int n, const x264dsp_mv_neighbours_t *nb, const int8_t *i_ref,
                                  const uint8_t *shape, int16_t *mvp, int16_t *pskip_mv, void *stream );

/* x264_mb_predict_mv_ref16x16 (common/mvpred.c:167-219; list 0, reference 0) for every macroblock of n_frames frames:
 * the candidate list of the 16x16 search -- the lookahead's MV doubled (lowres_mv[frame][mb][2], the MVs
 * x264dsp_lookahead_frame_cost_dev writes; NULL or a first entry of 0x7fff: none), the 16x16 MVs of the left, top,
 * top-left and top-right macroblocks (mvr[frame][mb][2] = h->mb.mvr[0][0]; outside the frame: (0,0)), and the reference
 * frame's 16x16 MVs at the same, right and lower macroblock scaled by scale = (curpoc - refpoc) * inv_ref_poc
 * ((mv * scale + 128) >> 8; l0_mv16 == NULL: the reference frame was intra, no temporal candidates).
 *   mvc [frame][mb][9][2] int16 (entries beyond the count untouched), n_mvc [frame][mb] int32 (4 .. 8).
 * The spatial candidates are the neighbours' FINAL vectors: a wavefront caller fills mvr as it goes. */
int x264dsp_predict_mvc_16x16_frames_dev( x264dsp_ctx_t *ctx, int mb_w, int mb_h, int n_frames,
                                          const int16_t *lowres_mv, const int16_t *mvr, const int16_t *l0_mv16,
                                          int scale, int16_t *mvc, int32_t *n_mvc, void *stream );

/* x264_macroblock_probe_pskip (encoder/macroblock.c:492-604) for every macroblock of n_frames frames: pred_slots hold
 * the P_SKIP prediction of each macroblock (x264dsp_mc_frames_dev at the clipped pskip MVs -- the function's own
 * mc_luma / mc_chroma calls), fenc_slots the source; skip[frame][mb] = 1 when the reference would return 1 (luma
 * decimate scores below 6 in total, each chroma plane passing its SSD / DC / AC-decimate ladder), else 0.  Nothing is
 * modified: as in the reference, the prediction doubles as the reconstruction of a macroblock that ends up skipped. */
int x264dsp_probe_pskip_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                                    const uint8_t *pred_slots, int n_frames, int qp, uint8_t *skip, void *stream );

/* x264_mb_mc for P_L0 16x16 macroblocks (common/macroblock.c:8-28; mc_luma common/mc.c:216-239,
 * mc_chroma common/mc.c:290-323): builds the prediction frame from one quarter-pel MV per MB. */
int x264dsp_mc_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slot,
                          const int16_t *mv, uint8_t *pred_slot, void *stream );
/* n_frames consecutive reference slots -> n_frames consecutive prediction slots, mv[n_frames][mb_count][2] */
int x264dsp_mc_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slots,
                           int n_frames, const int16_t *mv, uint8_t *pred_slots, void *stream );
/* The same with one MV per 8x8 block (mv8x8[frame][mb][4][2], raster order inside the macroblock): every partition
 * x264_mb_mc handles -- D_16x16, D_16x8, D_8x16, D_8x8 (common/macroblock.c:28-48) -- once the macroblock's MVs are
 * written out per 8x8, as h->mb.cache.mv holds them at x264_scan8[0], [4], [8], [12]. */
int x264dsp_mc_frames_part_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slots,
                                int n_frames, const int16_t *mv8x8, uint8_t *pred_slots, void *stream );

/* ------------------------------------------------------------------ I-slice analysis + coding (8(f) N1)
 * The I-slice branch of x264_macroblock_analyse (encoder/analyse.c:1079-1088: x264_mb_analyse_intra 565-763 -- the 16x16 modes,
 * the sixteen 4x4 blocks with their mode lists, shortcuts, early exits and in-place coding --, the I16x16 / I4x4 decision,
 * x264_mb_analyse_intra_chroma 509-563) followed by x264_macroblock_encode's intra branches, for EVERY macroblock of a frame.
 * A macroblock predicts from the reconstruction of its left, top-left, top and top-right neighbours and from the 4x4 modes
 * next to it, so the frame is a wavefront (row y may do column x once row y-1 has finished column x+1).  analyse.intra =
 * I4x4 (no 8x8 transform: the reference's build), SATD metric (every subme the reference accepts).
 *   recon_slots  n_frames output slots: luma plane N and chroma receive the reconstruction
 *   mb_type      [frame][mb] 0 = I_4x4, 2 = I_16x16 (the reference's enum values)
 *   mode16 / chroma_mode  [frame][mb] I_PRED_16x16_* / I_PRED_CHROMA_* as chosen (common/predict.h; mode16 is what an
 *                I_16x16 macroblock codes, the best 16x16 mode otherwise)
 *   modes4       [frame][mb][16] I_PRED_4x4_* in coding order (I_16x16 macroblocks: all 2 = DC, as their neighbours see them)
 *   levels / luma_dc / nnz / cbp  as x264dsp_residual_frames_typed_dev */
int x264dsp_i_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                          int n_frames, int qp, int8_t *mb_type, uint8_t *mode16, uint8_t *chroma_mode, uint8_t *modes4,
                          int16_t *levels, int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, void *stream );

/* ------------------------------------------------------------------ slice types, GOPs and their sharding (8(f) N4)
 * x264_slicetype_analyse + scenecut + the key-frame rules of x264_slicetype_decide (encoder/slicetype.c:322-435, 508-537;
 * no B frames, closed GOPs) for a WHOLE sequence at once: icost[k] / pcost[k] are frame k's intra estimate and its inter
 * estimate against frame k-1 (sums[X264DSP_LA_COST_INTRA] / [X264DSP_LA_COST_INTER] of the lookahead; pcost[0] is not read).
 * Both come from source frames only, so one lookahead pass decides every type: types[k] = X264DSP_TYPE_IDR / _I / _P
 * (the reference's X264_TYPE_* values, common/x264.h).  Host code, no device work. */
enum { X264DSP_TYPE_IDR = 1, X264DSP_TYPE_I = 2, X264DSP_TYPE_P = 3 };
typedef struct x264dsp_gop_params
{
    int32_t keyint_max, keyint_min;        /* h->param.i_keyint_max / i_keyint_min (after x264's validation) */
    int32_t scenecut_threshold;            /* h->param.i_scenecut_threshold, 0 = off */
} x264dsp_gop_params_t;
int x264dsp_slicetype_decide( int n_frames, const int32_t *icost, const int32_t *pcost,
                              const x264dsp_gop_params_t *p, uint8_t *types );
/* every I / IDR frame opens a GOP: gop_first[i], gop_count[i] for i < *n_gops (arrays of n_frames entries suffice) */
int x264dsp_gop_ranges( int n_frames, const uint8_t *types, int32_t *gop_first, int32_t *gop_count, int *n_gops );
/* the GOPs rank `rank` of `world` encodes: whole GOPs, in order, contiguous, balanced by frame count.  With one reference
 * frame nothing reaches across an I frame, so the ranks need no exchange until the bitstreams are concatenated. */
int x264dsp_gop_shard( int n_gops, const int32_t *gop_count, int rank, int world, int *first_gop, int *count );

/* ------------------------------------------------------------------ P-slice analysis + coding (8(f) N2)
 * x264_macroblock_analyse for a P slice (encoder/analyse.c:1059-1232: x264_mb_analyse_init 327-420, the fast P_SKIP
 * probe, x264_mb_analyse_inter_p16x16 787-860 with its early P_SKIP exit, x264_me_refine_qpel) followed by
 * x264_macroblock_encode (encoder/macroblock.c:310-485: x264_mb_mc, residual coding, the forced-P_SKIP rule) for EVERY
 * macroblock of a frame, with the reference's own data flow between macroblocks: the MV prediction, the P_SKIP vector,
 * the 16x16 search's candidate list and the fast-skip condition of a macroblock read the FINAL types and vectors of its
 * left, top, top-left and top-right neighbours, so the frame is a wavefront (row y may work on column x once row y-1 has
 * finished column x+1).  One reference frame, analyse.inter == 0 (P16x16 only: the reference's default; its P-slice
 * analysis has no intra candidates -- analyse.c:1206-1210 is compiled out), CABAC cbp packing.
 *
 *   fenc_slots   n_frames source frames (luma N + NV12 chroma)
 *   fref_slots   n_frames reference frames (reconstructed, border-expanded, with their H / V / HV planes)
 *   recon_slots  n_frames output slots: luma plane N and chroma receive the reconstruction (p_fdec)
 *   lowres_mv    [frame][mb][2] the lookahead's vectors of the pair (fenc->lowres_mvs[0][0]) or NULL
 *   l0_mv16      [frame][mb][2] the reference frame's own 16x16 vectors (its mvr output) for the temporal candidates,
 *                scaled by params->mvc_scale = (curpoc - refpoc) * inv_ref_poc; NULL: reference frame was intra
 * outputs, [frame][mb]...:
 *   mb_type      X264DSP_MB_P_L0 / X264DSP_MB_P_SKIP (the reference's enum values, common/macroblock.h:41-50)
 *   mv           [2] the macroblock's final vector (h->mb.cache.mv: the refined 16x16 vector, or the P_SKIP vector)
 *   mvr          [2] h->mb.mvr[0][0] = fdec->mv16x16: the 16x16 search result as later macroblocks and the next frame
 *                see it (zero for a macroblock skipped by the fast probe)
 *   mvd          [2] the vector difference the entropy coder writes (encoder/cabac.c:278-300: mv minus the prediction of
 *                x264_mb_predict_mv for the 16x16 partition; zero for P_SKIP); h->mb.mvd, the context the CABAC writer
 *                keeps for the neighbours, is min(|mvd|, 66) of it.  May be NULL.  (8(f) N3: the hand-off)
 *   levels / nnz / cbp  as x264dsp_residual_frames_dev (all zero for P_SKIP)
 * n_frames independent frames run as one launch (their wavefronts interleave); frames of one sequence depend on each
 * other through the reference frame and are launched one after the other. */
enum { X264DSP_MB_P_L0 = 4, X264DSP_MB_P_8x8 = 5, X264DSP_MB_P_SKIP = 6 };
typedef struct x264dsp_pframe_params
{
    int32_t me_method, subpel_refine, me_range;   /* as x264dsp_me_params_t */
    int32_t qp;                                    /* slice QP (h->sh.i_qp) */
    int32_t mv_range;                              /* h->param.analyse.i_mv_range, full-pel */
    int32_t fast_pskip;                            /* h->param.analyse.b_fast_pskip */
    int32_t mvc_scale;                             /* temporal candidates: (curpoc - refpoc) * inv_ref_poc */
    int32_t analyse_inter;                         /* h->param.analyse.inter: 0, or != 0 = X264_ANALYSE_PSUB16x16 (P8x8 / P16x8 /
                                                    * P8x16 are analysed as well: x264dsp_p_frames_part_dev) */
} x264dsp_pframe_params_t;
int x264dsp_p_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                          const uint8_t *fref_slots, uint8_t *recon_slots, int n_frames,
                          const x264dsp_pframe_params_t *params, const int16_t *lowres_mv, const int16_t *l0_mv16,
                          int8_t *mb_type, int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *levels, uint8_t *nnz,
                          int16_t *cbp, void *stream );

/* The same with the sub-16x16 partitions the reference analyses when params->analyse_inter has X264_ANALYSE_PSUB16x16:
 * x264_mb_analyse_inter_p8x8 (analyse.c:864-921), _p16x8 (923-990), _p8x16 (992-1054) with their early exits, the cost
 * comparison (1133-1172), x264_me_refine_qpel on every partition of the winner (1176-1203) and x264_mb_mc per partition.
 * Outputs as above except
 *   mb_type      also X264DSP_MB_P_8x8
 *   partition    h->mb.partition: 13 = D_8x8, 14 = D_16x8, 15 = D_8x16, 16 = D_16x16 (common/macroblock.h:92-101)
 *   mv8          [4][2] the final vector of each 8x8 block in raster order (all h->mb.cache.mv can hold without sub-8x8
 *                partitions; a 16x8 / 8x16 / 16x16 partition repeats its vector)
 *   mvd8         [4][2] the difference the entropy coder writes for the partition each 8x8 block lies in, predicted from the
 *                final vectors in coding order (encoder/cabac.c:352-412).  May be NULL.
 * params->analyse_inter == 0 gives x264dsp_p_frames_dev's decisions in this layout.
 * Slots, mv8 and mvd8 must start on 16-byte boundaries (the kernels use aligned vector accesses; X264DSP_E_ARG otherwise). */
int x264dsp_p_frames_part_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                               const uint8_t *fref_slots, uint8_t *recon_slots, int n_frames,
                               const x264dsp_pframe_params_t *params, const int16_t *lowres_mv, const int16_t *l0_mv16,
                               int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8,
                               int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream );

/* The same from HOST memory: i420 holds n_frames + 1 planar pictures, frame f + 1 is coded against picture f (the source
 * picture standing in for the reconstruction, as in x264dsp_recon_frames_host).  The reference planes (border, half-pel),
 * the half-resolution planes and the lookahead's vectors of every pair are built on the device; outputs are host arrays laid
 * out like the _dev ones, recon_i420 receives the n_frames reconstructions as planar I420.  All copies are inside the call. */
int x264dsp_p_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                           const x264dsp_pframe_params_t *params, int8_t *mb_type, int16_t *mv, int16_t *mvr, int16_t *mvd,
                           int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 );
/* ... with partitions (x264dsp_p_frames_part_dev's outputs: partition [frame][mb], mv8 / mvd8 [frame][mb][4][2]) */
int x264dsp_p_frames_part_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                const x264dsp_pframe_params_t *params, int8_t *mb_type, uint8_t *partition, int16_t *mv8,
                                int16_t *mvr, int16_t *mvd8, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 );

/* ------------------------------------------------------------------ closed GOPs on the device (8(f) N4 + the in-loop filter)
 * Boundary strengths from the slice kernels' own outputs: x264_macroblock_deblock_strength (common/macroblock.c:677-691) +
 * deblock_strength_c (common/deblock.c:297-323) for every macroblock of n_frames frames, one reference frame:
 *   mb_type [frame][mb], nnz [frame][mb][27] (entries 0..15: the luma 4x4 blocks in coding order), mv8 [frame][mb][4][2]
 *   (may be NULL when every macroblock is intra) -> bs [frame][mb][2][8][4] as x264dsp_deblock_frames_dev reads it.
 * Segments of edge 0 at the picture's left / top border are 0 (the reference never filters them). */
int x264dsp_boundary_strength_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, int n_frames, const int8_t *mb_type,
                                          const uint8_t *nnz, const int16_t *mv8, uint8_t *bs, void *stream );

/* n_gops closed GOPs of gop_len frames each (frame 0 an I frame, the rest P frames against their predecessor: the reference
 * without B frames, encoder/encoder.c:1180-1385) coded entirely on the device.  Slots and outputs are POSITION-major:
 * slot / entry [t * n_gops + gop] is frame t of GOP gop, so that one launch of each stage covers position t of every GOP:
 *   x264dsp_i_frames_dev (t = 0) or x264dsp_p_frames_dev / _part_dev against recon[t - 1]  ->  boundary strengths ->
 *   x264dsp_deblock_frames_dev -> x264dsp_frame_expand_border_dev -> x264dsp_frame_filter_dev on recon[t]
 * fenc_slots: gop_len * n_gops staged source frames; recon_slots: as many slots, which end up holding every frame's final
 * reference planes (N, H, V, HV, chroma).  lowres_mv [t][gop][mb][2]: the lookahead's vectors of frame t against frame t - 1
 * (entries of t = 0 unused) or NULL.  The previous frame's 16x16 vectors serve as temporal candidates from t = 2 on.
 * Outputs [t][gop][mb]...: mb_type, partition, mv8 / mvd8 [4][2] (one vector per 8x8 whatever params->analyse_inter), mvr [2],
 * levels, nnz, cbp as the slice kernels write them; mode16, chroma_mode, modes4 [16], luma_dc [16]: position 0 only
 * ([gop][mb]...).  Fixed QPs (qp_i for position 0, qp_p after it: constant-QP rate control). */
typedef struct x264dsp_gop_encode_params
{
    int32_t me_method, subpel_refine, me_range;
    int32_t qp_i, qp_p;
    int32_t mv_range, fast_pskip, analyse_inter;
    int32_t deblock, alpha_c0_offset, beta_offset;     /* h->param.b_deblocking_filter, h->sh.i_alpha_c0_offset, i_beta_offset */
} x264dsp_gop_encode_params_t;
int x264dsp_gops_encode_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                             int n_gops, int gop_len, const x264dsp_gop_encode_params_t *params, const int16_t *lowres_mv,
                             int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8, int16_t *levels,
                             uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode, uint8_t *modes4,
                             int16_t *luma_dc, void *stream );

/* Position t alone (same arrays and layout; positions in order: t reads the reconstruction and the 16x16 vectors of t - 1):
 * for a caller that feeds the positions as they arrive, as x264dsp_gops_encode_host does. */
int x264dsp_gops_encode_step_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                  int n_gops, int gop_len, int t, const x264dsp_gop_encode_params_t *params,
                                  const int16_t *lowres_mv, int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr,
                                  int16_t *mvd8, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode,
                                  uint8_t *modes4, int16_t *luma_dc, void *stream );

/* The same from HOST memory: i420 holds the GOPs one after the other ([gop][t] planar pictures); source staging, the
 * half-resolution planes and the lookahead's vectors of every (t - 1, t) pair are built on the device, then
 * x264dsp_gops_encode_dev, then x264dsp_levels_pack_dev.  Host outputs are position-major over all GOPs ([t][gop][mb]...,
 * index k = t * n_gops + gop) except mode16 / chroma_mode / modes4 / luma_dc ([gop][mb]...: the I frames); the compact
 * levels of frame k are packed_levels[frame_offset[k] .. + frame_size[k]), its macroblocks at mb_offset[k][mb] inside it
 * (X264DSP_E_ARG when packed_capacity, in int16 units, is too small for the content).  The reconstructions stay on the
 * device.  All copies are inside the call: position t + 1 is uploaded and position t - 1 downloaded while position t is coded. */
int x264dsp_gops_encode_host( x264dsp_ctx_t *ctx, int width, int height, int n_gops, int gop_len, const uint8_t *i420,
                              const x264dsp_gop_encode_params_t *params, int8_t *mb_type, uint8_t *partition, int16_t *mv8,
                              int16_t *mvr, int16_t *mvd8, uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode,
                              uint8_t *modes4, int16_t *luma_dc, int16_t *packed_levels, int64_t packed_capacity,
                              int64_t *frame_offset, int32_t *frame_size, int32_t *mb_offset );

/* ------------------------------------------------------------------ entropy hand-off, compact (8(f) N3)
 * The dense levels are 392 int16 per macroblock, 6.4 MB per 1080p frame, and the writer (x264_macroblock_write_cabac,
 * encoder/cabac.c:571-700) only reads a block whose non_zero_count flag is set.  The compact stream keeps exactly those units
 * in the writer's order, back to back, per macroblock:
 *   unit  0..15  luma 4x4 block k (16 levels)          present iff nnz[k]
 *   unit 16, 17  chroma DC of U, V (4 levels each)     present iff nnz[25], nnz[26]
 *   unit 18..25  U AC 0..3, V AC 0..3 (16 levels)      present iff nnz[16..19], nnz[20..23]
 * packed[frame * packed_stride + mb_offset[frame][mb] ...] = the macroblock's units; frame_total[frame] = the stream's length in
 * int16 units.  packed_stride >= mb_count * X264DSP_RES_LEVELS_PER_MB (the dense size is the worst case), a multiple of 4. */
int x264dsp_levels_pack_dev( x264dsp_ctx_t *ctx, int n_frames, int mb_count, const int16_t *levels, const uint8_t *nnz,
                             int16_t *packed, int64_t packed_stride, int32_t *mb_offset, int32_t *frame_total, void *stream );
/* x264dsp_p_frames_host / _part_host (partition != NULL: mv / mvd are [mb][4][2]) with that stream instead of the dense levels:
 * packed_levels (capacity in int16 units; X264DSP_E_ARG when the content does not fit) receives the frames back to back,
 * frame f at frame_offset[f] .. frame_offset[f + 1] (n_frames + 1 entries), macroblocks at mb_offset[f][mb] inside it.
 * recon_i420 may be NULL: the reconstruction then stays on the device. */
int x264dsp_p_frames_host_packed( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                  const x264dsp_pframe_params_t *params, int8_t *mb_type, uint8_t *partition, int16_t *mv,
                                  int16_t *mvr, int16_t *mvd, int16_t *packed_levels, int64_t packed_capacity,
                                  int64_t *frame_offset, int32_t *mb_offset, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 );

/* ------------------------------------------------------------------ deblock
 * x264_frame_deblock_row for every MB row (common/deblock.c:341-427) with the reference's
 * slice-QP rule.  mb_type / partition / cbp: per-MB; bs: [mb][2][8][4] boundary strengths.
 * Filters luma plane N and the NV12 chroma plane of `slot` in place. */
int x264dsp_deblock_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slot,
                               const int8_t *mb_type, const uint8_t *partition, const int16_t *cbp,
                               const uint8_t *bs, int qp, int alpha_c0_offset, int beta_offset,
                               void *stream );

/* the same for n_frames independent frames in one launch (consecutive slots; mb_type / partition /
 * cbp / bs hold n_frames x mb_count entries).  One frame alone is bound by the wavefront's critical
 * path (mb_w + 2*mb_h macroblock times); batching frames is what fills the machine. */
int x264dsp_deblock_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots, int n_frames,
                                const int8_t *mb_type, const uint8_t *partition, const int16_t *cbp,
                                const uint8_t *bs, int qp, int alpha_c0_offset, int beta_offset,
                                void *stream );

/* SURVEY 8(d) config 4 from HOST memory: n_frames P frames of P_L0 16x16 macroblocks.  i420 holds n_frames + 1 planar
 * pictures; frame f + 1 is predicted from frame f with mv16[f][mb][2] (x264_mb_mc), coded (x264_macroblock_encode: levels,
 * nnz, cbp as in x264dsp_residual_frames_dev) and its reconstruction deblocked (x264_frame_deblock_row with the caller's
 * mb_type / partition / bs per macroblock, the cbp just computed); recon_i420 receives the n_frames deblocked pictures
 * as planar I420 (width x height).  All copies are inside the call; groups of frames run on separate streams. */
int x264dsp_recon_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                               const int16_t *mv16, int qp, const int8_t *mb_type, const uint8_t *partition,
                               const uint8_t *bs, int alpha_c0_offset, int beta_offset,
                               int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 );
/* slot -> planar I420 (picture area only): the inverse of x264dsp_frame_load_i420_dev */
int x264dsp_frame_store_i420_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *slots,
                                  uint8_t *i420, int n_frames, void *stream );

/* deblock_strength_c (common/deblock.c:297-323) for n macroblocks:
 * nnz [n][120], ref [n][2][40], mv [n][2][40][2] -> bs [n][2][8][4] (scan8 layout). */
int x264dsp_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const uint8_t *nnz, const int8_t *ref,
                                  const int16_t *mv, uint8_t *bs, void *stream );

/* x264_macroblock_deblock_strength (common/macroblock.c:677-691) for n macroblocks: an intra macroblock (mb_type[i]
 * = I_4x4 .. I_PCM = 0 .. 3, common/macroblock.h:41-52) gets bS 3 on its three inner edges in both directions and its
 * bs[dir][0] is left as it was (x264_frame_deblock_row filters the outer edges of an intra macroblock with bS 4 whatever
 * is stored); every other macroblock goes through deblock_strength_c as above.  mb_type == NULL: all inter. */
int x264dsp_macroblock_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const int8_t *mb_type, const uint8_t *nnz,
                                             const int8_t *ref, const int16_t *mv, uint8_t *bs, void *stream );

#ifdef __cplusplus
}
#endif
#endif /* X264DSP_B200_H */
