/* x264dsp_glue.c -- the device side of the doors in x264dsp_doors.c, in plain C.
 *
 * This is the file a maintainer of the reference adds next to glue/x264dsp_doors.c and
 * glue/x264dsp_door_slicetype.c: it is compiled against the reference's own headers (x264_t, x264_frame_t) and calls
 * nothing but the C ABI of libx264dsp_b200.so (include/x264dsp_b200.h).  x264dsp_glue_install() puts one function
 * behind every door:
 *
 *   x264_frame_init_lowres (common/mc.c:404)                         -> x264dsp_frame_init_lowres_dev
 *   x264_frame_deblock_row + x264_frame_expand_border + x264_frame_filter + x264_frame_expand_border_filtered
 *       (common/deblock.c:341, common/frame.c:386-413, common/mc.c:506; encoder/encoder.c:1359-1385)
 *                                                                    -> x264dsp_deblock_frame_dev,
 *                                                                       x264dsp_frame_expand_border_dev, x264dsp_frame_filter_dev
 *   x264_slicetype_frame_cost (encoder/slicetype.c:223)              -> x264dsp_lookahead_frame_cost_dev
 *   x264_me_search_ref (encoder/me.c:129)                            -> x264dsp_me_search_batch_dev
 *   x264_mb_mc (common/macroblock.c:28)                              -> x264dsp_mc_frames_part_dev
 *   x264_macroblock_probe_pskip (encoder/macroblock.c:492)           -> x264dsp_mc_frame_dev + x264dsp_probe_pskip_frames_dev
 *   x264_macroblock_encode (encoder/macroblock.c:310)                -> x264dsp_residual_frames_typed_dev
 *   x264_macroblock_analyse + x264_macroblock_encode of a whole P slice (encoder/analyse.c:1059, encoder.c:1523-1578)
 *                                                                    -> x264dsp_p_frames_dev, once per frame: the host
 *                                                                       keeps the entropy coder (x264dsp_glue_install_pframe)
 *
 * The reference calls these doors one macroblock (or one frame) at a time, so this glue is a CORRECTNESS drop-in -- every
 * call is a round trip to the device -- and not the fast path (the frame-batched entry points are; bench.py times those).
 * What it proves is the boundary: the unmodified encoder, linked with these three files and the library, writes the
 * byte-identical bitstream (glue/Makefile builds it as x264ref_gpu; tests/test_gpu_glue_cli.py runs it).
 *
 * Frames the encoder will search in or read source samples from stay resident on the device (a small cache keyed by
 * the x264_frame_t pointer, filled by the lowres door for source frames and by the in-loop-filter door for
 * reconstructed frames).  A door whose frame is not resident declines (returns 1) and the reference's own code runs;
 * the door statistics (xref_door_stats_read) count that, so a caller can insist on zero declines.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common/common.h"
#include "x264dsp_b200.h"
#include "x264dsp_glue.h"

/* the doors (x264dsp_doors.c, x264dsp_door_slicetype.c) */
typedef void (*xref_frame_cb)( void *h, void *frame );
typedef void (*xref_cost_cb)( void *h, void *p0, void *b, int want_intra, int16_t *mvs, int *costs, int *sums );
typedef void (*xref_fdec_cb)( void *h, void *frame, int do_deblock, const int8_t *mb_type, const uint8_t *partition,
                              const int16_t *cbp, const uint8_t *bs, int qp, int alpha_off, int beta_off );
typedef int (*xref_me_cb)( void *h, void *fenc, void *fref, const x264dsp_me_block_t *in, int me_method, int subme,
                           int me_range, int qp, x264dsp_me_result_t *out );
typedef int (*xref_mbenc_cb)( void *h, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y, uint8_t *fdec_c,
                              int qp, int kind, const uint8_t *i4_modes, int16_t *levels, int16_t *luma_dc, uint8_t *nnz,
                              int *cbp );
typedef int (*xref_pskip_cb)( void *h, void *fenc, void *fref, int mb_x, int mb_y, int mvx, int mvy, int qp,
                              uint8_t *fdec_y, uint8_t *fdec_c, int *skip );
typedef int (*xref_mbmc_cb)( void *h, void *fref, int mb_x, int mb_y, const int16_t *mv8x8, uint8_t *fdec_y, uint8_t *fdec_c );
void xref_set_driver_hooks( xref_frame_cb lowres, xref_frame_cb filter, xref_cost_cb cost );
void xref_set_fdec_hook( xref_fdec_cb cb );
void xref_set_me_hook( xref_me_cb cb );
void xref_set_mbenc_hook( xref_mbenc_cb cb );
void xref_set_pskip_hook( xref_pskip_cb cb );
void xref_set_mbmc_hook( xref_mbmc_cb cb );
typedef struct
{
    const int8_t *mb_type;
    const int16_t *mv, *mvr;
    const int16_t *cbp;
    const int16_t *levels;
    const uint8_t *nnz;
    const uint8_t *recon_y, *recon_c;
    int stride_y, stride_c;
    const uint8_t *mode16, *chroma_mode, *modes4;
    const int16_t *luma_dc;
    const uint8_t *partition;
} xref_pframe_out_t;
typedef int (*xref_pframe_cb)( void *h, xref_pframe_out_t *out );
void xref_set_pframe_hook( xref_pframe_cb cb );
void xref_pframe_stats_read( int out[3] );
void xref_set_iframe_hook( xref_pframe_cb cb );
void xref_iframe_stats_read( int out[3] );
void xref_driver_hook_calls( int out[3] );
void xref_door_stats_read( int out[12] );

#define GLUE_RESIDENT 8

static struct
{
    x264dsp_ctx_t *ctx;
    x264dsp_geom_t g, g1;                 /* the picture; one macroblock as a 16x16 "frame" (macroblock_encode door) */
    uint8_t *pool;                        /* GLUE_RESIDENT resident slots + 2 work slots + 1 prediction slot */
    void *res_frame[GLUE_RESIDENT];       /* x264_frame_t* held by each resident slot */
    uint8_t res_stale[GLUE_RESIDENT];     /* the host's copy of this frame's H / V / HV planes has not been brought up to date */
    int res_next;
    int slice_hooks;                      /* both slice loops are on the device: the host never reads the half-pel planes */
    int lazy_planes;                      /* frames whose half-pel planes were fetched after all (a slice fell back to the host) */
    /* device scratch */
    int8_t *d_mb_type;
    uint8_t *d_partition, *d_bs, *d_skip, *d_mb_slots, *d_kind, *d_modes, *d_nnz;
    int16_t *d_cbp, *d_mvs, *d_pmv, *d_mv4, *d_levels, *d_luma_dc, *d_cbp1;
    int32_t *d_costs, *d_sums;
    x264dsp_me_block_t *d_blk;
    x264dsp_me_result_t *d_res;
    /* whole-P-frame results: device, and their host copies */
    int8_t *d_pf_type, *h_pf_type;
    int16_t *d_pf_mv, *d_pf_mvr, *d_pf_cbp, *d_pf_levels, *d_pf_lmv, *d_pf_l0;
    int16_t *h_pf_mv, *h_pf_mvr, *h_pf_cbp, *h_pf_levels;
    uint8_t *d_pf_nnz, *h_pf_nnz, *h_pf_luma, *h_pf_chroma, *d_pf_part, *h_pf_part;
    uint8_t *d_if_mode16, *d_if_cmode, *d_if_modes4, *h_if_mode16, *h_if_cmode, *h_if_modes4;
    int16_t *d_if_dc, *h_if_dc;
    /* host scratch */
    uint8_t *mb_stage;                    /* two 16x16 slots */
    uint8_t *rows;                        /* 16 luma rows of a macroblock as one linear piece of the plane */
    int64_t launches0;
    int calls[10];                        /* lowres, fdec, cost, me, mbenc, pskip, mbmc, deblocked frames, P frames, I frames */
    double seconds[10];                   /* wall time inside the hooks, same order (copies and waits included) */
    double t_install, seconds_open;       /* seconds_open: context creation + allocations, inside the first hook call */
} G;

#include <time.h>
static double glue_clock( void )
{
    struct timespec ts;
    clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define GLUE_CHECK( call ) do { int rc_ = ( call ); if( rc_ ) { fprintf( stderr, "x264dsp glue: %s failed (%d) at %s:%d\n", \
                                #call, rc_, __FILE__, __LINE__ ); abort(); } } while( 0 )

static void *glue_dev( size_t bytes )
{
    void *p = NULL;
    GLUE_CHECK( x264dsp_dev_alloc( G.ctx, bytes, &p ) );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, p, bytes, NULL ) );
    return p;
}

/* first call: the picture size is known from the encoder handle */
static void glue_open( x264_t *h )
{
    if( G.ctx )
        return;
    const double t_open = glue_clock();
    GLUE_CHECK( x264dsp_create( 0, &G.ctx ) );
    GLUE_CHECK( x264dsp_geometry( h->param.i_width, h->param.i_height, &G.g ) );
    GLUE_CHECK( x264dsp_geometry( 16, 16, &G.g1 ) );
    const size_t nmb = G.g.mb_count;
    G.pool = glue_dev( (size_t)( GLUE_RESIDENT + 3 ) * G.g.slot_bytes );
    G.d_mb_type = glue_dev( nmb );
    G.d_partition = glue_dev( nmb );
    G.d_cbp = glue_dev( 2 * nmb );
    G.d_bs = glue_dev( 64 * nmb );
    G.d_skip = glue_dev( nmb );
    G.d_mvs = glue_dev( 4 * nmb );
    G.d_pmv = glue_dev( 4 * nmb );
    G.d_mv4 = glue_dev( 16 * nmb );
    G.d_costs = glue_dev( 4 * nmb );
    G.d_sums = glue_dev( X264DSP_LA_SUMS * sizeof(int32_t) );
    G.d_blk = glue_dev( sizeof(x264dsp_me_block_t) );
    G.d_res = glue_dev( sizeof(x264dsp_me_result_t) );
    G.d_mb_slots = glue_dev( 2 * (size_t)G.g1.slot_bytes );
    G.d_kind = glue_dev( 16 );
    G.d_modes = glue_dev( 16 );
    G.d_levels = glue_dev( X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t) );
    G.d_luma_dc = glue_dev( 16 * sizeof(int16_t) );
    G.d_nnz = glue_dev( 32 );
    G.d_cbp1 = glue_dev( 16 );
    G.mb_stage = calloc( 2, (size_t)G.g1.slot_bytes );
    G.rows = malloc( (size_t)16 * G.g.luma_stride );
    G.d_pf_type = glue_dev( nmb );
    G.d_pf_mv = glue_dev( 16 * nmb );                          /* [mb][4][2] when partitions are analysed */
    G.d_pf_part = glue_dev( nmb );
    G.d_pf_mvr = glue_dev( 4 * nmb );
    G.d_pf_cbp = glue_dev( 2 * nmb );
    G.d_pf_levels = glue_dev( nmb * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t) );
    G.d_pf_nnz = glue_dev( nmb * X264DSP_RES_NNZ_PER_MB );
    G.d_pf_lmv = glue_dev( 4 * nmb );
    G.d_pf_l0 = glue_dev( 4 * nmb );
    G.d_if_mode16 = glue_dev( nmb );
    G.d_if_cmode = glue_dev( nmb );
    G.d_if_modes4 = glue_dev( 16 * nmb );
    G.d_if_dc = glue_dev( 32 * nmb );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb, (void **)&G.h_if_mode16 ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb, (void **)&G.h_if_cmode ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, 16 * nmb, (void **)&G.h_if_modes4 ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, 32 * nmb, (void **)&G.h_if_dc ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb, (void **)&G.h_pf_type ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, 16 * nmb, (void **)&G.h_pf_mv ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb, (void **)&G.h_pf_part ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, 4 * nmb, (void **)&G.h_pf_mvr ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, 2 * nmb, (void **)&G.h_pf_cbp ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t), (void **)&G.h_pf_levels ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, nmb * X264DSP_RES_NNZ_PER_MB, (void **)&G.h_pf_nnz ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, G.g.luma_plane_size, (void **)&G.h_pf_luma ) );
    GLUE_CHECK( x264dsp_host_alloc( G.ctx, G.g.chroma_plane_size, (void **)&G.h_pf_chroma ) );
    G.launches0 = x264dsp_launch_count( G.ctx );
    G.seconds_open = glue_clock() - t_open;
}

/* page-lock the encoder's own frame buffers the first time they are seen (the reference allocates its frames once and
 * recycles them): pageable copies run at a fraction of the link's rate and were a third of the glue's time per frame */
#define GLUE_PINNED_MAX 256
static struct { void *p; } glue_pinned[GLUE_PINNED_MAX];
static int glue_n_pinned;
static void glue_pin( void *p, size_t bytes )
{
    for( int i = 0; i < glue_n_pinned; i++ )
        if( glue_pinned[i].p == p )
            return;
    if( glue_n_pinned == GLUE_PINNED_MAX )
        return;
    glue_pinned[glue_n_pinned++].p = p;
    if( x264dsp_host_register( G.ctx, p, bytes ) )
        fprintf( stderr, "x264dsp glue: could not page-lock a frame buffer (copies from it stay pageable)\n" );
}
static void glue_pin_frame( x264_frame_t *f )
{
    const x264dsp_geom_t *g = &G.g;
    static int on = -1;
    if( on < 0 )
        on = getenv( "X264DSP_GLUE_PIN" ) && atoi( getenv( "X264DSP_GLUE_PIN" ) );
    if( !on )
        return;
    glue_pin( f->buffer[0], 4 * (size_t)g->luma_plane_size );
    glue_pin( f->buffer[1], g->chroma_plane_size );
    if( f->buffer_lowres[0] )
        glue_pin( f->buffer_lowres[0], 4 * (size_t)g->lowres_plane_size );
}

static uint8_t *glue_slot( int i )      { return G.pool + (size_t)i * G.g.slot_bytes; }
static uint8_t *glue_work( int i )      { return glue_slot( GLUE_RESIDENT + i ); }
static uint8_t *glue_pred( void )       { return glue_slot( GLUE_RESIDENT + 2 ); }

static uint8_t *glue_resident( const void *frame )
{
    for( int i = 0; i < GLUE_RESIDENT; i++ )
        if( G.res_frame[i] == frame )
            return glue_slot( i );
    return NULL;
}

/* the slot that will hold `frame` from now on (the oldest entry makes room) */
static uint8_t *glue_make_resident( void *frame )
{
    uint8_t *s = glue_resident( frame );
    if( s )
        return s;
    const int i = G.res_next;
    G.res_next = ( G.res_next + 1 ) % GLUE_RESIDENT;
    G.res_frame[i] = frame;
    G.res_stale[i] = 0;
    return glue_slot( i );
}

/* ---- x264_frame_init_lowres: padded source plane in; four padded half-resolution planes out, and the source plane with
 *      its duplicated last column / row (mc.c:412-415).  buffer[0] = planes N | H | V | HV, buffer[1] = NV12 chroma,
 *      buffer_lowres[0] = the four lowres planes -- the layout of one frame slot (x264dsp_geom_t). */
static void glue_lowres( void *hv, void *fv )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    x264_frame_t *f = fv;
    glue_open( h );
    const x264dsp_geom_t *g = &G.g;
    uint8_t *slot = glue_make_resident( f );
    glue_pin_frame( f );
    GLUE_CHECK( x264dsp_h2d( G.ctx, slot, f->buffer[0], g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_h2d( G.ctx, slot + g->slot_chroma_off, f->buffer[1], g->chroma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_frame_init_lowres_dev( G.ctx, g, slot, 1, NULL ) );
    GLUE_CHECK( x264dsp_frame_export_lowres_dev( G.ctx, g, slot, 1, NULL ) );    /* the reference wants row-major lowres[0..3] */
    GLUE_CHECK( x264dsp_d2h( G.ctx, f->buffer[0], slot, g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, f->buffer_lowres[0], slot + g->slot_lowres_off, 4 * (size_t)g->lowres_plane_size, NULL ) );
    G.seconds[0] += glue_clock() - t0_;
    G.calls[0]++;
}

/* ---- the in-loop filter of a reconstructed frame, once per frame: deblock -> expand_border -> hpel planes */
static void glue_fdec( void *hv, void *fv, int do_deblock, const int8_t *mb_type, const uint8_t *partition,
                       const int16_t *cbp, const uint8_t *bs, int qp, int alpha_off, int beta_off )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    x264_frame_t *f = fv;
    glue_open( h );
    const x264dsp_geom_t *g = &G.g;
    const size_t nmb = g->mb_count;
    uint8_t *slot = glue_make_resident( f );                 /* this reconstruction is the next frame's reference */
    glue_pin_frame( f );
    GLUE_CHECK( x264dsp_h2d( G.ctx, slot, f->buffer[0], g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_h2d( G.ctx, slot + g->slot_chroma_off, f->buffer[1], g->chroma_plane_size, NULL ) );
    if( do_deblock )
    {
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_mb_type, mb_type, nmb, NULL ) );
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_partition, partition, nmb, NULL ) );
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_cbp, cbp, 2 * nmb, NULL ) );
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_bs, bs, 64 * nmb, NULL ) );
        GLUE_CHECK( x264dsp_deblock_frame_dev( G.ctx, g, slot, G.d_mb_type, G.d_partition, G.d_cbp, G.d_bs, qp, alpha_off,
                                               beta_off, NULL ) );
        G.calls[7]++;
    }
    GLUE_CHECK( x264dsp_frame_expand_border_dev( G.ctx, g, slot, 1, NULL ) );
    GLUE_CHECK( x264dsp_frame_filter_dev( G.ctx, g, slot, 1, NULL ) );
    /* With the macroblock loops of both slice types on the device the host never reads a reference frame's half-pel planes
     * (no search, no motion compensation on its side): only the reconstruction itself goes back (PSNR, dumps), the other three
     * planes -- two thirds of the bytes -- stay behind and are fetched by glue_fetch_planes should a slice ever fall back. */
    /* ... and only while every P slice is certain to be offered to the device: the door in front of x264_macroblock_analyse
     * does not even ask when the stream has several references or slices, trellis or noise reduction (x264dsp_doors.c) */
    const int lazy = G.slice_hooks && h->param.i_frame_reference == 1 && !h->param.analyse.i_trellis && !h->param.analyse.i_noise_reduction
                     && !( h->param.analyse.inter & ~X264_ANALYSE_PSUB16x16 ) && h->param.i_slice_count <= 1 && !h->param.i_slice_max_size
                     && !h->param.i_slice_max_mbs;
    GLUE_CHECK( x264dsp_d2h( G.ctx, f->buffer[0], slot, ( lazy ? 1 : 4 ) * (size_t)g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, f->buffer[1], slot + g->slot_chroma_off, g->chroma_plane_size, NULL ) );
    for( int i = 0; i < GLUE_RESIDENT; i++ )
        if( G.res_frame[i] == f )
            G.res_stale[i] = (uint8_t)lazy;
    G.seconds[1] += glue_clock() - t0_;
    G.calls[1]++;
}

/* ---- x264_slicetype_frame_cost( p0, b, b ): the lowres planes as the caller holds them (row-major) go into two work
 *      slots, are brought into the lookahead's tiled layout, and the per-block MVs / costs and the frame sums come back
 *      into the reference's own arrays */
static void glue_cost( void *hv, void *p0v, void *bv, int want_intra, int16_t *mvs, int *costs, int *sums )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    x264_frame_t *p0 = p0v, *b = bv;
    glue_open( h );
    const x264dsp_geom_t *g = &G.g;
    const size_t wps4 = 4 * (size_t)g->lowres_plane_size, nmb = g->mb_count;
    const int32_t bi[1] = { GLUE_RESIDENT + 1 }, pi[1] = { GLUE_RESIDENT };
    const uint8_t wi[1] = { (uint8_t)( want_intra != 0 ) };
    int32_t s[X264DSP_LA_SUMS];
    GLUE_CHECK( x264dsp_h2d( G.ctx, glue_work( 0 ) + g->slot_lowres_off, p0->buffer_lowres[0], wps4, NULL ) );
    GLUE_CHECK( x264dsp_h2d( G.ctx, glue_work( 1 ) + g->slot_lowres_off, b->buffer_lowres[0], wps4, NULL ) );
    GLUE_CHECK( x264dsp_frame_retile_lowres_dev( G.ctx, g, glue_work( 0 ), 2, NULL ) );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, G.d_mvs, 4 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, G.d_costs, 4 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, G.d_sums, sizeof(s), NULL ) );
    GLUE_CHECK( x264dsp_lookahead_frame_cost_dev( G.ctx, g, G.pool, 1, bi, pi, wi, G.d_mvs, G.d_costs, G.d_sums, NULL, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, mvs, G.d_mvs, 4 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, costs, G.d_costs, 4 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, s, G.d_sums, sizeof(s), NULL ) );
    memcpy( sums, s, 8 * sizeof(int) );
    G.seconds[2] += glue_clock() - t0_;
    G.calls[2]++;
}

/* ---- x264_me_search_ref for one partition: a one-block list on the resident source / reference frames */
static int glue_me( void *hv, void *fenc, void *fref, const x264dsp_me_block_t *in, int me_method, int subme, int me_range,
                    int qp, x264dsp_me_result_t *out )
{
    const double t0_ = glue_clock();
    const uint8_t *se = glue_resident( fenc ), *sr = glue_resident( fref );
    if( !G.ctx || !se || !sr )
        return 1;
    const x264dsp_me_params_t prm = { me_method, subme, me_range, qp, 0 };
    GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_blk, in, sizeof(*in), NULL ) );
    GLUE_CHECK( x264dsp_me_search_batch_dev( G.ctx, &G.g, se, sr, &prm, 1, G.d_blk, G.d_res, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, out, G.d_res, sizeof(*out), NULL ) );
    G.seconds[3] += glue_clock() - t0_;
    G.calls[3]++;
    return 0;
}

/* the 16x16 luma and 8x8 U / V samples of macroblock (mb_x, mb_y) of a prediction slot -> fdec (stride 32, U at +0 and
 * V at +16 of the chroma rows, common/macroblock.c:242-265) */
static void glue_fetch_mb( const uint8_t *slot, int mb_x, int mb_y, uint8_t *fdec_y, uint8_t *fdec_c )
{
    const x264dsp_geom_t *g = &G.g;
    const size_t lo = (size_t)g->luma_origin + (size_t)mb_y * 16 * g->luma_stride + mb_x * 16;
    const size_t co = (size_t)g->slot_chroma_off + g->chroma_origin + (size_t)mb_y * 8 * g->chroma_stride + mb_x * 16;
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.rows, slot + lo, (size_t)15 * g->luma_stride + 16, NULL ) );
    for( int r = 0; r < 16; r++ )
        memcpy( fdec_y + r * 32, G.rows + (size_t)r * g->luma_stride, 16 );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.rows, slot + co, (size_t)7 * g->chroma_stride + 16, NULL ) );
    for( int r = 0; r < 8; r++ )
        for( int x = 0; x < 8; x++ )
        {
            fdec_c[r * 32 + x] = G.rows[(size_t)r * g->chroma_stride + 2 * x];
            fdec_c[r * 32 + 16 + x] = G.rows[(size_t)r * g->chroma_stride + 2 * x + 1];
        }
}

/* ---- x264_mb_mc: the four 8x8 MVs of the macroblock -> prediction into fdec */
static int glue_mbmc( void *hv, void *fref, int mb_x, int mb_y, const int16_t *mv8x8, uint8_t *fdec_y, uint8_t *fdec_c )
{
    const double t0_ = glue_clock();
    const uint8_t *sr = glue_resident( fref );
    if( !G.ctx || !sr )
        return 1;
    const size_t xy = (size_t)mb_y * G.g.mb_w + mb_x;
    /* the frame kernel predicts every macroblock; only this one's vectors matter (the others stay zero) */
    GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_mv4 + xy * 8, mv8x8, 16, NULL ) );
    GLUE_CHECK( x264dsp_mc_frames_part_dev( G.ctx, &G.g, sr, 1, G.d_mv4, glue_pred(), NULL ) );
    glue_fetch_mb( glue_pred(), mb_x, mb_y, fdec_y, fdec_c );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, G.d_mv4 + xy * 8, 16, NULL ) );
    G.seconds[6] += glue_clock() - t0_;
    G.calls[6]++;
    return 0;
}

/* ---- x264_macroblock_probe_pskip: mc at the clipped pskip mv, then the "would anything be coded" test */
static int glue_pskip( void *hv, void *fenc, void *fref, int mb_x, int mb_y, int mvx, int mvy, int qp, uint8_t *fdec_y,
                       uint8_t *fdec_c, int *skip )
{
    const double t0_ = glue_clock();
    const uint8_t *se = glue_resident( fenc ), *sr = glue_resident( fref );
    if( !G.ctx || !se || !sr )
        return 1;
    const size_t xy = (size_t)mb_y * G.g.mb_w + mb_x;
    const int16_t mv[2] = { (int16_t)mvx, (int16_t)mvy };
    uint8_t s = 0;
    GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_pmv + xy * 2, mv, 4, NULL ) );
    GLUE_CHECK( x264dsp_mc_frame_dev( G.ctx, &G.g, sr, G.d_pmv, glue_pred(), NULL ) );
    GLUE_CHECK( x264dsp_probe_pskip_frames_dev( G.ctx, &G.g, se, glue_pred(), 1, qp, G.d_skip, NULL ) );
    glue_fetch_mb( glue_pred(), mb_x, mb_y, fdec_y, fdec_c );
    GLUE_CHECK( x264dsp_d2h( G.ctx, &s, G.d_skip + xy, 1, NULL ) );
    GLUE_CHECK( x264dsp_dev_zero( G.ctx, G.d_pmv + xy * 2, 4, NULL ) );
    *skip = s;
    G.seconds[5] += glue_clock() - t0_;
    G.calls[5]++;
    return 0;
}

/* ---- x264_macroblock_encode: one macroblock as a 16x16 "frame" -- source and prediction slots in the library's own
 *      plane layout, levels / nnz / cbp back in the layout the door hands to the reference's entropy coder */
static uint8_t *glue_mb_luma( uint8_t *slot )   { return slot + G.g1.luma_origin; }
static uint8_t *glue_mb_chroma( uint8_t *slot ) { return slot + G.g1.slot_chroma_off + G.g1.chroma_origin; }

static int glue_mbenc( void *hv, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y, uint8_t *fdec_c, int qp,
                       int kind, const uint8_t *i4_modes, int16_t *levels, int16_t *luma_dc, uint8_t *nnz, int *cbp )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    glue_open( h );
    const x264dsp_geom_t *g1 = &G.g1;
    const int ls = g1->luma_stride, cs = g1->chroma_stride;
    uint8_t *src = G.mb_stage, *prd = G.mb_stage + g1->slot_bytes;
    const uint8_t k8 = (uint8_t)kind;
    int16_t c16 = 0;
    for( int r = 0; r < 16; r++ )
    {
        memcpy( glue_mb_luma( src ) + r * ls, fenc_y + r * 16, 16 );          /* FENC_STRIDE 16 */
        memcpy( glue_mb_luma( prd ) + r * ls, fdec_y + r * 32, 16 );          /* FDEC_STRIDE 32 */
    }
    for( int r = 0; r < 8; r++ )
        for( int x = 0; x < 8; x++ )
        {
            glue_mb_chroma( src )[r * cs + 2 * x] = fenc_c[r * 16 + x];       /* U at +0, V at +8 */
            glue_mb_chroma( src )[r * cs + 2 * x + 1] = fenc_c[r * 16 + 8 + x];
            glue_mb_chroma( prd )[r * cs + 2 * x] = fdec_c[r * 32 + x];       /* U at +0, V at +16 */
            glue_mb_chroma( prd )[r * cs + 2 * x + 1] = fdec_c[r * 32 + 16 + x];
        }
    if( kind & 2 )
    {
        /* I4x4: the reconstructed neighbourhood fdec_buf holds around the macroblock goes into the slot's padding --
         * the row above from column -1 to 19 and the column to the left */
        memcpy( glue_mb_luma( prd ) - ls - 1, fdec_y - 33, 21 );
        for( int r = 0; r < 16; r++ )
            glue_mb_luma( prd )[r * ls - 1] = fdec_y[r * 32 - 1];
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_modes, i4_modes, 16, NULL ) );
    }
    GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_mb_slots, G.mb_stage, 2 * (size_t)g1->slot_bytes, NULL ) );
    GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_kind, &k8, 1, NULL ) );
    GLUE_CHECK( x264dsp_residual_frames_typed_dev( G.ctx, g1, G.d_mb_slots, G.d_mb_slots + g1->slot_bytes, 1, qp, G.d_kind,
                                                   G.d_modes, G.d_levels, G.d_luma_dc, G.d_nnz, G.d_cbp1, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, prd, G.d_mb_slots + g1->slot_bytes, g1->slot_bytes, NULL ) );
    for( int r = 0; r < 16; r++ )
        memcpy( fdec_y + r * 32, glue_mb_luma( prd ) + r * ls, 16 );
    for( int r = 0; r < 8; r++ )
        for( int x = 0; x < 8; x++ )
        {
            fdec_c[r * 32 + x] = glue_mb_chroma( prd )[r * cs + 2 * x];
            fdec_c[r * 32 + 16 + x] = glue_mb_chroma( prd )[r * cs + 2 * x + 1];
        }
    GLUE_CHECK( x264dsp_d2h( G.ctx, levels, G.d_levels, X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t), NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, luma_dc, G.d_luma_dc, 16 * sizeof(int16_t), NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, nnz, G.d_nnz, X264DSP_RES_NNZ_PER_MB, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, &c16, G.d_cbp1, sizeof(c16), NULL ) );
    *cbp = c16;
    G.seconds[4] += glue_clock() - t0_;
    G.calls[4]++;
    return 0;
}

/* ---- the macroblock loop of a whole P slice: x264dsp_p_frames_dev on the resident source and reference frame, with the
 *      lookahead's vectors of the pair and the reference frame's 16x16 vectors as the search's extra candidates exactly
 *      as x264_mb_predict_mv_ref16x16 takes them (common/mvpred.c:167-219) */
/* the slice is going back to the host's own loop: its reference frame needs the half-pel planes the in-loop filter left behind */
static void glue_fetch_planes( x264_frame_t *f )
{
    for( int i = 0; i < GLUE_RESIDENT; i++ )
        if( G.res_frame[i] == f && G.res_stale[i] )
        {
            GLUE_CHECK( x264dsp_d2h( G.ctx, f->buffer[0] + G.g.luma_plane_size, glue_slot( i ) + G.g.luma_plane_size,
                                     3 * (size_t)G.g.luma_plane_size, NULL ) );
            G.res_stale[i] = 0;
            G.lazy_planes++;
        }
}

static int glue_pframe( void *hv, xref_pframe_out_t *out )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    x264_frame_t *fref = h->fref[0][0];
    const uint8_t *se = glue_resident( h->fenc ), *sr = glue_resident( fref );
    if( !G.ctx || !se || !sr )
    {
        if( G.ctx )
            glue_fetch_planes( fref );
        return 1;
    }
    const x264dsp_geom_t *g = &G.g;
    const size_t nmb = g->mb_count;
    const int idx = h->fenc->i_frame - fref->i_frame - 1;
    const int have_lowres = h->frames.b_have_lowres && idx >= 0 && idx <= h->param.i_bframe
                            && h->fenc->lowres_mvs[0][idx][0][0] != 0x7fff;
    const int have_l0 = fref->i_ref[0] > 0;
    x264dsp_pframe_params_t prm;
    prm.me_method = h->mb.i_me_method;
    prm.subpel_refine = h->mb.i_subpel_refine;
    prm.me_range = h->param.analyse.i_me_range;
    prm.qp = h->sh.i_qp;
    prm.mv_range = h->param.analyse.i_mv_range;
    prm.fast_pskip = h->param.analyse.b_fast_pskip;
    prm.mvc_scale = have_l0 ? ( h->fdec->i_poc - fref->i_poc ) * fref->inv_ref_poc[0] : 0;
    prm.analyse_inter = h->param.analyse.inter & X264_ANALYSE_PSUB16x16;
    const int by_part = prm.analyse_inter != 0;
    if( have_lowres )
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_pf_lmv, h->fenc->lowres_mvs[0][idx], 4 * nmb, NULL ) );
    if( have_l0 )
        GLUE_CHECK( x264dsp_h2d( G.ctx, G.d_pf_l0, fref->mv16x16, 4 * nmb, NULL ) );
    if( by_part ? x264dsp_p_frames_part_dev( G.ctx, g, se, sr, glue_pred(), 1, &prm, have_lowres ? G.d_pf_lmv : NULL,
                                             have_l0 ? G.d_pf_l0 : NULL, G.d_pf_type, G.d_pf_part, G.d_pf_mv, G.d_pf_mvr, NULL,
                                             G.d_pf_levels, G.d_pf_nnz, G.d_pf_cbp, NULL )
                : x264dsp_p_frames_dev( G.ctx, g, se, sr, glue_pred(), 1, &prm, have_lowres ? G.d_pf_lmv : NULL, have_l0 ? G.d_pf_l0 : NULL,
                                        G.d_pf_type, G.d_pf_mv, G.d_pf_mvr, NULL, G.d_pf_levels, G.d_pf_nnz, G.d_pf_cbp, NULL ) )
    {
        glue_fetch_planes( fref );
        return 1;                                              /* parameters the device path does not take: the host's own loop */
    }
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_type, G.d_pf_type, nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_mv, G.d_pf_mv, ( by_part ? 16 : 4 ) * nmb, NULL ) );
    if( by_part )
        GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_part, G.d_pf_part, nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_mvr, G.d_pf_mvr, 4 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_cbp, G.d_pf_cbp, 2 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_levels, G.d_pf_levels, nmb * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t), NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_nnz, G.d_pf_nnz, nmb * X264DSP_RES_NNZ_PER_MB, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_luma, glue_pred(), g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_chroma, glue_pred() + g->slot_chroma_off, g->chroma_plane_size, NULL ) );
    memset( out, 0, sizeof(*out) );
    out->mb_type = G.h_pf_type;
    out->mv = G.h_pf_mv;
    out->partition = by_part ? G.h_pf_part : NULL;
    out->mvr = G.h_pf_mvr;
    out->cbp = G.h_pf_cbp;
    out->levels = G.h_pf_levels;
    out->nnz = G.h_pf_nnz;
    out->recon_y = G.h_pf_luma + g->luma_origin;
    out->recon_c = G.h_pf_chroma + g->chroma_origin;
    out->stride_y = g->luma_stride;
    out->stride_c = g->chroma_stride;
    G.seconds[8] += glue_clock() - t0_;
    G.calls[8]++;
    return 0;
}

/* ---- the macroblock loop of a whole I slice: x264dsp_i_frames_dev on the resident source frame */
static int glue_iframe( void *hv, xref_pframe_out_t *out )
{
    const double t0_ = glue_clock();
    x264_t *h = hv;
    const uint8_t *se = glue_resident( h->fenc );
    if( !G.ctx || !se )
        return 1;
    const x264dsp_geom_t *g = &G.g;
    const size_t nmb = g->mb_count;
    if( x264dsp_i_frames_dev( G.ctx, g, se, glue_pred(), 1, h->sh.i_qp, G.d_pf_type, G.d_if_mode16, G.d_if_cmode, G.d_if_modes4,
                              G.d_pf_levels, G.d_if_dc, G.d_pf_nnz, G.d_pf_cbp, NULL ) )
        return 1;
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_type, G.d_pf_type, nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_if_mode16, G.d_if_mode16, nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_if_cmode, G.d_if_cmode, nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_if_modes4, G.d_if_modes4, 16 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_if_dc, G.d_if_dc, 32 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_cbp, G.d_pf_cbp, 2 * nmb, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_levels, G.d_pf_levels, nmb * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t), NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_nnz, G.d_pf_nnz, nmb * X264DSP_RES_NNZ_PER_MB, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_luma, glue_pred(), g->luma_plane_size, NULL ) );
    GLUE_CHECK( x264dsp_d2h( G.ctx, G.h_pf_chroma, glue_pred() + g->slot_chroma_off, g->chroma_plane_size, NULL ) );
    memset( out, 0, sizeof(*out) );
    out->mb_type = G.h_pf_type;
    out->cbp = G.h_pf_cbp;
    out->levels = G.h_pf_levels;
    out->nnz = G.h_pf_nnz;
    out->recon_y = G.h_pf_luma + g->luma_origin;
    out->recon_c = G.h_pf_chroma + g->chroma_origin;
    out->stride_y = g->luma_stride;
    out->stride_c = g->chroma_stride;
    out->mode16 = G.h_if_mode16;
    out->chroma_mode = G.h_if_cmode;
    out->modes4 = G.h_if_modes4;
    out->luma_dc = G.h_if_dc;
    G.seconds[9] += glue_clock() - t0_;
    G.calls[9]++;
    return 0;
}

/* the whole P-slice macroblock loop on the device (on top of x264dsp_glue_install) */
void x264dsp_glue_install_pframe( void )
{
    xref_set_pframe_hook( glue_pframe );
}

/* the whole I-slice macroblock loop on the device as well: with both installed the host runs no analysis and no
 * macroblock coding at all, only the entropy coder */
void x264dsp_glue_install_iframe( void )
{
    xref_set_iframe_hook( glue_iframe );
    G.slice_hooks = 1;                                         /* installed on top of the P-slice hook (x264dsp_glue_auto.c) */
}

void x264dsp_glue_install( void )
{
    G.t_install = glue_clock();
    xref_set_driver_hooks( glue_lowres, NULL, glue_cost );
    xref_set_fdec_hook( glue_fdec );
    xref_set_me_hook( glue_me );
    xref_set_mbenc_hook( glue_mbenc );
    xref_set_pskip_hook( glue_pskip );
    xref_set_mbmc_hook( glue_mbmc );
}

void x264dsp_glue_uninstall( void )
{
    xref_set_driver_hooks( NULL, NULL, NULL );
    xref_set_fdec_hook( NULL );
    xref_set_me_hook( NULL );
    xref_set_mbenc_hook( NULL );
    xref_set_pskip_hook( NULL );
    xref_set_mbmc_hook( NULL );
    xref_set_pframe_hook( NULL );
    xref_set_iframe_hook( NULL );
}

/* one JSON object: how often each door was served by the device, what the doors themselves counted
 * ({entered, eligible, served} per door: eligible != served means a silent fallback), kernel launches */
int x264dsp_glue_report( FILE *out )
{
    int hook[3], doors[12], pf[3], iff[3];
    xref_driver_hook_calls( hook );
    xref_pframe_stats_read( pf );
    xref_iframe_stats_read( iff );
    xref_door_stats_read( doors );
    return fprintf( out, "{\"lowres\": %d, \"inloop_filter\": %d, \"deblocked_frames\": %d, \"lookahead_cost\": %d, "
                    "\"me_search\": %d, \"macroblock_encode\": %d, \"probe_pskip\": %d, \"mb_mc\": %d, "
                    "\"door_me\": [%d, %d, %d], \"door_mbenc\": [%d, %d, %d], \"door_pskip\": [%d, %d, %d], "
                    "\"door_mbmc\": [%d, %d, %d], \"hook_calls\": [%d, %d, %d], \"p_frames\": %d, "
                    "\"p_slices\": [%d, %d, %d], \"i_frames\": %d, \"i_slices\": [%d, %d, %d], \"kernel_launches\": %lld, "
                    "\"seconds\": {\"lowres\": %.4f, \"inloop_filter\": %.4f, \"lookahead_cost\": %.4f, \"p_frames\": %.4f, "
                    "\"i_frames\": %.4f, \"per_macroblock_doors\": %.4f, \"open\": %.4f, \"since_install\": %.4f}, "
                    "\"half_pel_planes_fetched_late\": %d}\n",
                    G.calls[0], G.calls[1], G.calls[7], G.calls[2], G.calls[3], G.calls[4], G.calls[5], G.calls[6],
                    doors[0], doors[1], doors[2], doors[3], doors[4], doors[5], doors[6], doors[7], doors[8], doors[9],
                    doors[10], doors[11], hook[0], hook[1], hook[2], G.calls[8], pf[0], pf[1], pf[2], G.calls[9], iff[0], iff[1], iff[2],
                    G.ctx ? (long long)( x264dsp_launch_count( G.ctx ) - G.launches0 ) : 0LL,
                    G.seconds[0] - G.seconds_open, G.seconds[1], G.seconds[2], G.seconds[8], G.seconds[9],
                    G.seconds[3] + G.seconds[4] + G.seconds[5] + G.seconds[6], G.seconds_open, glue_clock() - G.t_install, G.lazy_planes );
}
