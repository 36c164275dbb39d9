/* Driver-level doors of the drop-in: the reference-side half of the glue (the device side is x264dsp_glue.c).
 *
 * Three drivers of the hot path are compiled under other names (see oracle/Makefile: the UNMODIFIED
 * sources, renamed with -D on the command line) so that the definitions below take their place:
 *
 *   x264_frame_init_lowres              common/mc.c:404      called from encoder/encoder.c:1770
 *   x264_frame_filter                   common/mc.c:506      called from encoder/encoder.c:1382
 *   x264_frame_expand_border_filtered   common/frame.c:398   called from encoder/encoder.c:1383
 *
 *   x264_frame_deblock_row              common/deblock.c:341 called from encoder/encoder.c:1370
 *   x264_frame_expand_border            common/frame.c:386   called from encoder/encoder.c:1376
 *
 *   x264_me_search_ref                  encoder/me.c:129     called from encoder/analyse.c:820 ... (every partition)
 *
 * and encoder/slicetype.c's x264_slicetype_decide is wrapped in x264dsp_door_slicetype.c.  With no hooks
 * installed every definition forwards to the original, so the library and the CLI behave exactly like
 * the reference (tests/test_golden.py::test_reference_cli_bitstream pins that).  With hooks installed
 * (tests/test_gpu_dropin_drivers.py) the planes and the lookahead costs come from libx264dsp_b200.so;
 * this is the glue a maintainer of the reference would add.
 */
#include "common/common.h"
#include "encoder/macroblock.h"
#include "encoder/me.h"

typedef void (*xref_frame_cb)( void *h, void *frame );
typedef void (*xref_cost_cb)( void *h, void *p0, void *b, int want_intra, int16_t *mvs, int *costs, int *sums );

/* whole in-loop filter of a reconstructed frame: deblock (if do_deblock) -> expand_border -> hpel planes */
typedef void (*xref_fdec_cb)( void *h, void *frame, int do_deblock, const int8_t *mb_type, const uint8_t *partition,
                              const int16_t *cbp, const uint8_t *bs, int qp, int alpha_off, int beta_off );

xref_frame_cb xref_hook_lowres = NULL, xref_hook_filter = NULL;
xref_cost_cb xref_hook_cost = NULL;
xref_fdec_cb xref_hook_fdec = NULL;
int xref_hook_calls[3] = { 0, 0, 0 };
/* per door { calls that entered with the hook installed, of those eligible for the device (main-encode calls the
 * door is meant for), of those served by the device }: the test requires served == eligible -- no silent fallback */
enum { XREF_DOOR_ME = 0, XREF_DOOR_MBENC, XREF_DOOR_PSKIP, XREF_DOOR_MBMC, XREF_DOORS };
int xref_door_stats[XREF_DOORS][3];
void xref_door_stats_read( int out[XREF_DOORS * 3] ) { memcpy( out, xref_door_stats, sizeof(xref_door_stats) ); }
void xref_door_stats_reset( void ) { memset( xref_door_stats, 0, sizeof(xref_door_stats) ); }
static uint8_t *xref_bs_stash = NULL;       /* [mb_h][mb_w][2][8][4], filled row by row */
static int xref_bs_rows = 0;

/* supersedes the filter hook: deblocking and border expansion move to the end of the frame as well */
void xref_set_fdec_hook( xref_fdec_cb cb )
{
    xref_hook_fdec = cb;
    xref_bs_rows = 0;
}

void xref_set_driver_hooks( xref_frame_cb lowres, xref_frame_cb filter, xref_cost_cb cost )
{
    xref_hook_lowres = lowres;
    xref_hook_filter = filter;
    xref_hook_cost = cost;
    xref_hook_calls[0] = xref_hook_calls[1] = xref_hook_calls[2] = 0;
}

void xref_driver_hook_calls( int out[3] )
{
    out[0] = xref_hook_calls[0];
    out[1] = xref_hook_calls[1];
    out[2] = xref_hook_calls[2];
}

void xref_orig_frame_init_lowres( x264_t *h, x264_frame_t *frame );
void xref_orig_frame_filter( x264_t *h, x264_frame_t *frame, int mb_y, int b_end );
void xref_orig_frame_expand_border_filtered( x264_t *h, x264_frame_t *frame, int mb_y, int b_end );

void x264_frame_init_lowres( x264_t *h, x264_frame_t *frame )
{
    int x, y;
    if( !xref_hook_lowres )
    {
        xref_orig_frame_init_lowres( h, frame );
        return;
    }
    /* the four padded half-resolution planes and the duplicated last column / row of the source plane */
    xref_hook_lowres( h, frame );
    xref_hook_calls[0]++;
    /* per-frame bookkeeping the driver also does (mc.c:421-431) */
    memset( frame->i_cost_est, -1, sizeof(frame->i_cost_est) );
    for( y = 0; y < h->param.i_bframe + 2; y++ )
        for( x = 0; x < h->param.i_bframe + 2; x++ )
            frame->i_row_satds[y][x][0] = -1;
    for( y = 0; y <= !!h->param.i_bframe; y++ )
        for( x = 0; x <= h->param.i_bframe; x++ )
            frame->lowres_mvs[y][x][0][0] = 0x7FFF;
}

/* the reference filters MB row by MB row as the rows are reconstructed; the hooked version filters the
 * whole frame once, when the last row arrives (the planes are first read by the next frame's search) */
void x264_frame_filter( x264_t *h, x264_frame_t *frame, int mb_y, int b_end )
{
    if( !xref_hook_filter && !xref_hook_fdec )
        xref_orig_frame_filter( h, frame, mb_y, b_end );
}

void xref_orig_frame_deblock_row( x264_t *h, int mb_y );
void xref_orig_frame_expand_border( x264_t *h, x264_frame_t *frame, int mb_y );

/* the reference deblocks row mb_y here; the hooked version only keeps the row's boundary strengths
 * (they live in a two-row ring, common/common.h:1085) and filters the whole frame at its end.  Intra
 * prediction reads the unfiltered samples either way (intra_border_backup), so the encode of the
 * current frame does not notice. */
void x264_frame_deblock_row( x264_t *h, int mb_y )
{
    if( !xref_hook_fdec )
    {
        xref_orig_frame_deblock_row( h, mb_y );
        return;
    }
    if( !xref_bs_stash )
        xref_bs_stash = malloc( (size_t)h->mb.i_mb_count * 64 );
    memcpy( xref_bs_stash + (size_t)mb_y * h->mb.i_mb_width * 64, h->deblock_strength[mb_y&1], (size_t)h->mb.i_mb_width * 64 );
    xref_bs_rows++;
}

void x264_frame_expand_border( x264_t *h, x264_frame_t *frame, int mb_y )
{
    if( !xref_hook_fdec )
        xref_orig_frame_expand_border( h, frame, mb_y );
}

/* observer: called when the last macroblock row of a frame has been reconstructed, before anything else happens to it
 * (the capture point of the P-slice analysis pin, tests/test_oracle_pframe.py); changes nothing */
typedef void (*xref_observe_cb)( void *h, void *frame );
xref_observe_cb xref_hook_observe = NULL;
void xref_set_observer( xref_observe_cb cb ) { xref_hook_observe = cb; }

void x264_frame_expand_border_filtered( x264_t *h, x264_frame_t *frame, int mb_y, int b_end )
{
    if( b_end && xref_hook_observe )
        xref_hook_observe( h, frame );
    if( xref_hook_fdec )
    {
        if( b_end )
        {
            xref_hook_fdec( h, frame, xref_bs_rows == h->mb.i_mb_height, h->mb.type, h->mb.partition, h->mb.cbp,
                            xref_bs_stash, h->sh.i_qp, h->sh.i_alpha_c0_offset, h->sh.i_beta_offset );
            xref_bs_rows = 0;
            xref_hook_calls[1]++;
        }
        return;
    }
    if( !xref_hook_filter )
    {
        xref_orig_frame_expand_border_filtered( h, frame, mb_y, b_end );
        return;
    }
    if( b_end )
    {
        xref_hook_filter( h, frame );
        xref_hook_calls[1]++;
    }
}


/* ------------------------------------------------------------------ motion search
 * x264_me_search_ref for one partition of the main encode, served by x264dsp_me_search_batch_dev with a
 * one-block list.  The block description is exactly the x264_me_t inputs plus the MV limits the analysis
 * has put into h->mb (the layout of x264dsp_me_block_t / xref_me_in_t in harness.c). */
typedef struct
{
    int32_t i_pixel;
    int32_t bx, by;
    int16_t mvp[2];
    int32_t i_mvc;
    int16_t mvc[16][2];
    int32_t mv_min_fpel[2], mv_max_fpel[2];
    int32_t mv_min_spel[2], mv_max_spel[2];
} xref_hook_me_in_t;

typedef struct
{
    int16_t mv[2];
    int32_t cost;
    int32_t cost_mv;
} xref_hook_me_out_t;

typedef int (*xref_me_cb)( void *h, void *fenc, void *fref, const xref_hook_me_in_t *in, int me_method, int subme,
                           int me_range, int qp, xref_hook_me_out_t *out );
xref_me_cb xref_hook_me = NULL;
int xref_hook_me_calls = 0;

void xref_set_me_hook( xref_me_cb cb )
{
    xref_hook_me = cb;
    xref_hook_me_calls = 0;
}
int xref_me_hook_calls( void ) { return xref_hook_me_calls; }

void xref_orig_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh );

void x264_me_search_ref( x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh )
{
    x264_frame_t *fref = h->fref[0][0];
    xref_hook_me_in_t in;
    xref_hook_me_out_t out;
    int k, qp;
    if( xref_hook_me )
        xref_door_stats[XREF_DOOR_ME][0]++;
    /* only the main encode's searches in the newest reference frame; the lowres lookahead (other planes,
     * other stride) and multi-reference early termination keep the reference's own code */
    if( !xref_hook_me || p_halfpel_thresh || !fref || m->i_ref != 0 || m->i_stride[0] != fref->i_stride[0]
        || i_mvc > 16 || m->p_fref[0] < fref->filtered[0][0]
        || m->p_fref[0] >= fref->filtered[0][0] + (intptr_t)fref->i_stride[0] * fref->i_lines[0] )
    {
        xref_orig_me_search_ref( h, m, mvc, i_mvc, p_halfpel_thresh );
        return;
    }
    for( qp = 0; qp < 52 && h->cost_mv[qp] != m->p_cost_mv; qp++ )
        ;
    if( qp == 52 )
    {
        xref_orig_me_search_ref( h, m, mvc, i_mvc, p_halfpel_thresh );
        return;
    }
    xref_door_stats[XREF_DOOR_ME][1]++;
    {
        const intptr_t off = m->p_fref[0] - fref->filtered[0][0];
        in.by = (int32_t)( off / fref->i_stride[0] );
        in.bx = (int32_t)( off - (intptr_t)in.by * fref->i_stride[0] );
    }
    in.i_pixel = m->i_pixel;
    in.mvp[0] = m->mvp[0];
    in.mvp[1] = m->mvp[1];
    in.i_mvc = i_mvc;
    memset( in.mvc, 0, sizeof(in.mvc) );
    for( k = 0; k < i_mvc; k++ )
    {
        in.mvc[k][0] = mvc[k][0];
        in.mvc[k][1] = mvc[k][1];
    }
    for( k = 0; k < 2; k++ )
    {
        in.mv_min_fpel[k] = h->mb.mv_min_fpel[k];
        in.mv_max_fpel[k] = h->mb.mv_max_fpel[k];
        in.mv_min_spel[k] = h->mb.mv_min_spel[k];
        in.mv_max_spel[k] = h->mb.mv_max_spel[k];
    }
    if( xref_hook_me( h, h->fenc, fref, &in, h->mb.i_me_method, h->mb.i_subpel_refine, h->param.analyse.i_me_range, qp, &out ) )
    {
        xref_orig_me_search_ref( h, m, mvc, i_mvc, p_halfpel_thresh );     /* the hook declined (frame not resident) */
        return;
    }
    xref_hook_me_calls++;
    xref_door_stats[XREF_DOOR_ME][2]++;
    m->mv[0] = out.mv[0];
    m->mv[1] = out.mv[1];
    m->cost = out.cost;
    m->cost_mv = out.cost_mv;
}

/* ------------------------------------------------------------------------------------------------
 * x264_macroblock_analyse     encoder/analyse.c:1059       called from encoder/encoder.c:1526 (slice loop)
 *
 * The whole P-slice macroblock loop on the device: when the first macroblock of a P slice arrives, the
 * hook runs x264dsp_p_frames_dev on the resident source / reference frames and hands back, for every macroblock of the
 * frame, what x264_macroblock_analyse and x264_macroblock_encode would have left behind -- type, vector, 16x16 search
 * vector, levels, nnz, cbp, reconstruction.  This door then only installs macroblock xy's decisions where the rest of the
 * slice loop reads them (what x264_mb_analyse_init and x264_analyse_update_cache write, analyse.c:327-420, 1236-1300),
 * and the x264_macroblock_encode door below serves the macroblock's coded data from the same frame result: the host is
 * left with the entropy coder.  Eligible: one reference frame, analyse.inter == 0 or X264_ANALYSE_PSUB16x16 (the only
 * inter flag the reference's analysis reads), no trellis / noise reduction.
 * Both doors also keep the time the reference spends in its own two functions (bench.py's cpu_baseline for this path). */
typedef struct
{
    const int8_t *mb_type;          /* [mb] */
    const int16_t *mv, *mvr;        /* [mb][2] */
    const int16_t *cbp;             /* [mb] */
    const int16_t *levels;          /* [mb][392] */
    const uint8_t *nnz;             /* [mb][27] */
    const uint8_t *recon_y, *recon_c;   /* sample (0,0) of the reconstructed luma / NV12 chroma planes */
    int stride_y, stride_c;
    /* I slices (x264dsp_i_frames_dev): */
    const uint8_t *mode16, *chroma_mode;    /* [mb] */
    const uint8_t *modes4;                  /* [mb][16], coding order */
    const int16_t *luma_dc;                 /* [mb][16] */
    /* P slices with analyse.inter = PSUB16x16 (x264dsp_p_frames_part_dev): h->mb.partition per macroblock, and mv is then
     * [mb][4][2], one vector per 8x8 block; NULL: 16x16 only */
    const uint8_t *partition;
} xref_pframe_out_t;
typedef int (*xref_pframe_cb)( void *h, xref_pframe_out_t *out );
xref_pframe_cb xref_hook_pframe = NULL, xref_hook_iframe = NULL;
static xref_pframe_out_t xref_pframe;
static int xref_pframe_live = 0;            /* the current slice is served from xref_pframe */
int xref_pframe_stats[3];                   /* P slices seen with the hook installed, served, macroblocks served */
/* time this THREAD has spent inside the reference's own x264_macroblock_analyse / x264_macroblock_encode on P slices,
 * and the macroblocks it covers (every encoder instance runs on its own thread in the benchmark) */
static __thread double xref_door_seconds[2];
static __thread long xref_door_mbs;
int xref_door_timing = 0;

void xref_set_pframe_hook( xref_pframe_cb cb )
{
    xref_hook_pframe = cb;
    xref_pframe_live = 0;
    xref_pframe_stats[0] = xref_pframe_stats[1] = xref_pframe_stats[2] = 0;
}
void xref_pframe_stats_read( int out[3] ) { memcpy( out, xref_pframe_stats, sizeof(xref_pframe_stats) ); }
/* the same for I slices (x264_mb_analyse_intra and the intra branches of x264_macroblock_encode on the device) */
int xref_iframe_stats[3];
void xref_set_iframe_hook( xref_pframe_cb cb )
{
    xref_hook_iframe = cb;
    xref_pframe_live = 0;
    xref_iframe_stats[0] = xref_iframe_stats[1] = xref_iframe_stats[2] = 0;
}
void xref_iframe_stats_read( int out[3] ) { memcpy( out, xref_iframe_stats, sizeof(xref_iframe_stats) ); }
void xref_set_door_timing( int on ) { xref_door_timing = on; }
/* read and reset the calling thread's counters: out = { seconds in analyse, seconds in encode, macroblocks } */
void xref_door_seconds_read( double out[3] )
{
    out[0] = xref_door_seconds[0]; out[1] = xref_door_seconds[1]; out[2] = (double)xref_door_mbs;
    xref_door_seconds[0] = xref_door_seconds[1] = 0;
    xref_door_mbs = 0;
}

#include <time.h>
static double xref_clock( void )
{
    struct timespec ts;
    clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void xref_orig_macroblock_analyse( x264_t *h );

void x264_macroblock_analyse( x264_t *h )
{
    if( h->sh.i_type == SLICE_TYPE_P && xref_hook_pframe && h->mb.i_mb_xy == h->sh.i_first_mb )
    {
        xref_pframe_stats[0]++;
        xref_pframe_live = h->i_ref[0] == 1 && !( h->param.analyse.inter & ~X264_ANALYSE_PSUB16x16 ) && !h->param.analyse.i_trellis
                           && !h->param.analyse.i_noise_reduction && h->sh.i_first_mb == 0 && h->sh.i_qp <= QP_MAX_SPEC
                           && !xref_hook_pframe( h, &xref_pframe );
        xref_pframe_stats[1] += xref_pframe_live;
    }
    else if( h->sh.i_type == SLICE_TYPE_I && xref_hook_iframe && h->mb.i_mb_xy == h->sh.i_first_mb )
    {
        xref_iframe_stats[0]++;
        xref_pframe_live = ( h->param.analyse.intra & X264_ANALYSE_I4x4 ) && !h->param.analyse.b_transform_8x8
                           && !h->param.analyse.i_trellis && !h->param.analyse.i_noise_reduction && h->sh.i_first_mb == 0
                           && h->sh.i_qp <= QP_MAX_SPEC && !xref_hook_iframe( h, &xref_pframe );
        xref_iframe_stats[1] += xref_pframe_live;
    }
    else if( h->mb.i_mb_xy == h->sh.i_first_mb )
        xref_pframe_live = 0;
    if( !xref_pframe_live )
    {
        if( xref_door_timing && h->sh.i_type == SLICE_TYPE_P )
        {
            const double t0 = xref_clock();
            xref_orig_macroblock_analyse( h );
            xref_door_seconds[0] += xref_clock() - t0;
            xref_door_mbs++;
        }
        else
            xref_orig_macroblock_analyse( h );
        return;
    }
    if( h->sh.i_type == SLICE_TYPE_I )
    {
        /* what x264_mb_analyse_init, x264_mb_analyse_intra and x264_analyse_update_cache leave behind for an intra macroblock */
        const int xy = h->mb.i_mb_xy;
        int i;
        h->mb.i_qp = h->sh.i_qp;
        h->mb.i_chroma_qp = h->chroma_qp_table[h->sh.i_qp];
        h->mb.b_transform_8x8 = 0;
        h->mb.b_noise_reduction = 0;
        h->mb.b_trellis = 0;
        h->mb.i_skip_intra = 0;
        h->mb.i_type = xref_pframe.mb_type[xy];
        if( h->mb.i_type == I_4x4 )
            for( i = 0; i < 16; i++ )
                h->mb.cache.intra4x4_pred_mode[x264_scan8[i]] = (int8_t)xref_pframe.modes4[16 * xy + i];
        else
            h->mb.i_intra16x16_pred_mode = xref_pframe.mode16[xy];
        h->mb.i_chroma_pred_mode = xref_pframe.chroma_mode[xy];
        xref_iframe_stats[2]++;
        return;
    }
    {
        const int xy = h->mb.i_mb_xy;
        const int type = xref_pframe.mb_type[xy];
        /* x264_mb_analyse_init */
        h->mb.i_qp = h->sh.i_qp;
        h->mb.i_chroma_qp = h->chroma_qp_table[h->sh.i_qp];
        h->mb.b_transform_8x8 = 0;
        h->mb.b_noise_reduction = 0;
        h->mb.b_trellis = 0;
        h->mb.b_skip_mc = 1;                 /* the prediction never has to be built on the host */
        h->mb.i_skip_intra = 1;
        h->mb.mv_min[0] = ( -( h->mb.i_mb_x << 4 ) - 24 ) << 2;
        h->mb.mv_max[0] = ( ( ( h->mb.i_mb_width - h->mb.i_mb_x - 1 ) << 4 ) + 24 ) << 2;
        h->mb.mv_min[1] = ( -( h->mb.i_mb_y << 4 ) - 24 ) << 2;
        h->mb.mv_max[1] = ( ( ( h->mb.i_mb_height - h->mb.i_mb_y - 1 ) << 4 ) + 24 ) << 2;
        /* x264_analyse_update_cache: P_L0 (16x16, 16x8, 8x16), P_8x8 with four 8x8 sub-partitions, or P_SKIP; reference 0 */
        h->mb.i_type = type;
        x264_macroblock_cache_ref( h, 0, 0, 4, 4, 0, 0 );
        if( xref_pframe.partition )
        {
            int k;
            h->mb.i_partition = xref_pframe.partition[xy];
            for( k = 0; k < 4; k++ )
            {
                x264_macroblock_cache_mv_ptr( h, 2 * ( k & 1 ), 2 * ( k >> 1 ), 2, 2, 0, (int16_t *)( xref_pframe.mv + 8 * xy + 2 * k ) );
                h->mb.i_sub_partition[k] = D_L0_8x8;
            }
        }
        else
        {
            h->mb.i_partition = D_16x16;
            x264_macroblock_cache_mv_ptr( h, 0, 0, 4, 4, 0, (int16_t *)( xref_pframe.mv + 2 * xy ) );
        }
        CP32( h->mb.mvr[0][0][xy], xref_pframe.mvr + 2 * xy );
        xref_pframe_stats[2]++;
    }
}

/* ------------------------------------------------------------------------------------------------
 * x264_macroblock_encode      encoder/macroblock.c:310     called from encoder/encoder.c (slice loop)
 *
 * Inter macroblocks of a P slice (P_L0, P_8x8; any partition -- the residual is one 16x16 transform
 * whatever the motion partition) and I16x16 macroblocks of an I slice: the reference's own x264_mb_mc /
 * h->predict_16x16[] / h->predict_chroma[] build the prediction in fdec, the
 * hook turns source + prediction into levels / nnz / cbp / reconstruction (x264dsp_residual_frame_dev
 * on the device), and what comes back is laid out where x264_macroblock_write_cabac / _cavlc and
 * x264_macroblock_cache_save read it (the entropy coder's hand-off format):
 *   levels  -> h->dct.luma4x4[0..15], h->dct.chroma_dc[0..1], h->dct.luma4x4[16..19] (U AC), [32..35] (V AC)
 *   nnz     -> h->mb.cache.non_zero_count[x264_scan8[..]]
 *   cbp     -> h->mb.i_cbp_luma, h->mb.i_cbp_chroma, h->mb.cbp[mb_xy] (CABAC: with the DC flags in bits 8..10)
 * followed by the forced-P_SKIP rule of macroblock.c:465-485.  Buffer shapes are xref_encode_inter_mb's. */
typedef int (*xref_mbenc_cb)( void *h, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y, uint8_t *fdec_c,
                              int qp, int kind /* 0 inter, 1 I16x16, 2 / 6 I4x4 (mb_kind of the library) */,
                              const uint8_t *i4_modes, int16_t *levels, int16_t *luma_dc, uint8_t *nnz, int *cbp );
xref_mbenc_cb xref_hook_mbenc = NULL;
int xref_hook_mbenc_calls = 0;

void xref_set_mbenc_hook( xref_mbenc_cb cb )
{
    xref_hook_mbenc = cb;
    xref_hook_mbenc_calls = 0;
}
int xref_mbenc_hook_calls( void ) { return xref_hook_mbenc_calls; }

void xref_orig_macroblock_encode( x264_t *h );

void x264_macroblock_encode( x264_t *h )
{
    int16_t levels[392], luma_dc[16];
    uint8_t nnz[27];
    int cbp = 0, i;
    /* I16x16 macroblocks of I slices (x264_mb_encode_i16x16 + intra chroma, no decimation): the reference's own
     * predictors fill fdec, the device does the rest (x264dsp_residual_frames_typed_dev, kind 1) */
    const int i16 = h->mb.i_type == I_16x16 && h->sh.i_type == SLICE_TYPE_I && !h->mb.b_dct_decimate;
    /* I4x4 macroblocks: fdec_buf holds the reconstructed neighbours the sixteen predictions start from; the modes go
     * over in coding order.  (h->mb.i_skip_intra only says that analysis has coded the blocks already -- the same
     * deterministic steps, so coding all sixteen again gives what the reference's shortcut copies back.) */
    const int i4 = h->mb.i_type == I_4x4 && h->sh.i_type == SLICE_TYPE_I && !h->mb.b_dct_decimate;
    uint8_t i4_modes[16];
    int kind = i16;
    const int inter = !IS_INTRA( h->mb.i_type ) && h->mb.i_type != P_SKIP && h->sh.i_type == SLICE_TYPE_P && h->mb.b_dct_decimate;
    if( xref_pframe_live && ( h->sh.i_type == SLICE_TYPE_P || h->sh.i_type == SLICE_TYPE_I ) )
    {
        /* the macroblock was coded on the device with the rest of its frame: reconstruction into fdec, coded data where
         * the entropy coder reads them (layout as below) */
        const int xy = h->mb.i_mb_xy;
        const uint8_t *ry = xref_pframe.recon_y + (intptr_t)( h->mb.i_mb_y << 4 ) * xref_pframe.stride_y + ( h->mb.i_mb_x << 4 );
        const uint8_t *rc = xref_pframe.recon_c + (intptr_t)( h->mb.i_mb_y << 3 ) * xref_pframe.stride_c + ( h->mb.i_mb_x << 4 );
        const int16_t *lv = xref_pframe.levels + (size_t)xy * 392;
        const uint8_t *nz = xref_pframe.nnz + (size_t)xy * 27;
        int x, y;
        cbp = xref_pframe.cbp[xy];
        for( y = 0; y < 16; y++ )
            memcpy( h->mb.pic.p_fdec[0] + y * FDEC_STRIDE, ry + (intptr_t)y * xref_pframe.stride_y, 16 );
        for( y = 0; y < 8; y++ )
            for( x = 0; x < 8; x++ )
            {
                h->mb.pic.p_fdec[1][y * FDEC_STRIDE + x] = rc[(intptr_t)y * xref_pframe.stride_c + 2 * x];
                h->mb.pic.p_fdec[2][y * FDEC_STRIDE + x] = rc[(intptr_t)y * xref_pframe.stride_c + 2 * x + 1];
            }
        memcpy( h->dct.luma4x4[0], lv, 16*16*sizeof(int16_t) );
        memcpy( h->dct.chroma_dc[0], lv + 256, 4*sizeof(int16_t) );
        memcpy( h->dct.chroma_dc[1], lv + 260, 4*sizeof(int16_t) );
        memcpy( h->dct.luma4x4[16], lv + 264, 4*16*sizeof(int16_t) );
        memcpy( h->dct.luma4x4[32], lv + 328, 4*16*sizeof(int16_t) );
        for( i = 0; i < 16; i++ )
            h->mb.cache.non_zero_count[x264_scan8[i]] = nz[i];
        for( i = 0; i < 4; i++ )
        {
            h->mb.cache.non_zero_count[x264_scan8[16+i]] = nz[16+i];
            h->mb.cache.non_zero_count[x264_scan8[32+i]] = nz[20+i];
        }
        h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]] = h->mb.i_type == I_16x16 ? nz[24] : 0;
        if( h->mb.i_type == I_16x16 )
            memcpy( h->dct.luma16x16_dc[0], xref_pframe.luma_dc + 16 * xy, 16*sizeof(int16_t) );
        h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]] = nz[25];
        h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]] = nz[26];
        h->mb.i_cbp_luma = cbp & 15;
        h->mb.i_cbp_chroma = ( cbp >> 4 ) & 3;
        h->mb.cbp[xy] = h->param.b_cabac ? cbp : ( cbp & 0x3f );
        return;
    }
    if( xref_hook_mbenc )
        xref_door_stats[XREF_DOOR_MBENC][0]++;
    if( xref_door_timing && !xref_hook_mbenc && h->sh.i_type == SLICE_TYPE_P )
    {
        const double t0 = xref_clock();
        xref_orig_macroblock_encode( h );
        xref_door_seconds[1] += xref_clock() - t0;
        return;
    }
    if( !xref_hook_mbenc || !( i16 || i4 || inter ) || h->mb.b_noise_reduction || h->mb.b_transform_8x8 || h->mb.b_lossless
        || h->mb.i_chroma_qp != h->chroma_qp_table[h->mb.i_qp] )
    {
        xref_orig_macroblock_encode( h );
        return;
    }
    xref_door_stats[XREF_DOOR_MBENC][1]++;
    h->mb.i_cbp_luma = 0;
    h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]] = 0;
    if( i16 || i4 )
    {
        if( i16 )
            h->predict_16x16[h->mb.i_intra16x16_pred_mode]( h->mb.pic.p_fdec[0] );
        h->predict_chroma[h->mb.i_chroma_pred_mode]( h->mb.pic.p_fdec[1] );
        h->predict_chroma[h->mb.i_chroma_pred_mode]( h->mb.pic.p_fdec[2] );
    }
    if( i4 )
    {
        for( i = 0; i < 16; i++ )
            i4_modes[i] = (uint8_t)h->mb.cache.intra4x4_pred_mode[x264_scan8[i]];
        kind = 2 + ( ( h->mb.i_neighbour4[5] & (MB_TOPRIGHT|MB_TOP) ) == MB_TOP ? 4 : 0 );
    }
    else if( inter && !h->mb.b_skip_mc )
        x264_mb_mc( h );
    memset( levels, 0, sizeof(levels) );
    memset( luma_dc, 0, sizeof(luma_dc) );
    if( xref_hook_mbenc( h, h->mb.pic.p_fenc[0], h->mb.pic.p_fenc[1], h->mb.pic.p_fdec[0], h->mb.pic.p_fdec[1],
                         h->mb.i_qp, kind, i4_modes, levels, luma_dc, nnz, &cbp ) )
    {
        /* declined: the prediction is already in fdec; an intra macroblock's predictors simply run again */
        if( inter )
            h->mb.b_skip_mc = 1;
        xref_orig_macroblock_encode( h );
        return;
    }
    xref_hook_mbenc_calls++;
    xref_door_stats[XREF_DOOR_MBENC][2]++;
    memcpy( h->dct.luma4x4[0], levels, 16*16*sizeof(int16_t) );
    memcpy( h->dct.chroma_dc[0], levels + 256, 4*sizeof(int16_t) );
    memcpy( h->dct.chroma_dc[1], levels + 260, 4*sizeof(int16_t) );
    memcpy( h->dct.luma4x4[16], levels + 264, 4*16*sizeof(int16_t) );
    memcpy( h->dct.luma4x4[32], levels + 328, 4*16*sizeof(int16_t) );
    if( i16 )
        memcpy( h->dct.luma16x16_dc[0], luma_dc, 16*sizeof(int16_t) );
    for( i = 0; i < 16; i++ )
        h->mb.cache.non_zero_count[x264_scan8[i]] = nnz[i];
    for( i = 0; i < 4; i++ )
    {
        h->mb.cache.non_zero_count[x264_scan8[16+i]] = nnz[16+i];
        h->mb.cache.non_zero_count[x264_scan8[32+i]] = nnz[20+i];
    }
    h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]] = nnz[24];
    h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]] = nnz[25];
    h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]] = nnz[26];
    h->mb.i_cbp_luma = cbp & 15;
    h->mb.i_cbp_chroma = ( cbp >> 4 ) & 3;
    h->mb.cbp[h->mb.i_mb_xy] = h->param.b_cabac ? cbp : ( cbp & 0x3f );
    if( h->mb.i_type == P_L0 && h->mb.i_partition == D_16x16 && !( h->mb.i_cbp_luma | h->mb.i_cbp_chroma )
        && M32( h->mb.cache.mv[0][x264_scan8[0]] ) == M32( h->mb.cache.pskip_mv )
        && h->mb.cache.ref[0][x264_scan8[0]] == 0 )
        h->mb.i_type = P_SKIP;
}

/* ------------------------------------------------------------------------------------------------
 * x264_macroblock_probe_pskip   encoder/macroblock.c:492     called from encoder/analyse.c (P-slice analysis)
 *
 * The P_SKIP probe of a macroblock: motion compensation at the clipped pskip mv, then the "would anything be
 * coded" test.  The hook gets the frames (kept resident on the device by the test), the macroblock position, the
 * clipped mv and the QP; it leaves the prediction in fdec, as the reference does, and returns the decision
 * (x264dsp_mc_frame_dev + x264dsp_probe_pskip_frames_dev). */
typedef int (*xref_pskip_cb)( void *h, void *fenc, void *fref, int mb_x, int mb_y, int mvx, int mvy, int qp,
                              uint8_t *fdec_y, uint8_t *fdec_c, int *skip );
xref_pskip_cb xref_hook_pskip = NULL;
int xref_hook_pskip_calls = 0;

void xref_set_pskip_hook( xref_pskip_cb cb )
{
    xref_hook_pskip = cb;
    xref_hook_pskip_calls = 0;
}
int xref_pskip_hook_calls( void ) { return xref_hook_pskip_calls; }

int xref_orig_macroblock_probe_pskip( x264_t *h );

int x264_macroblock_probe_pskip( x264_t *h )
{
    int skip = 0, mvx, mvy;
    if( xref_hook_pskip )
        xref_door_stats[XREF_DOOR_PSKIP][0]++;
    if( !xref_hook_pskip || h->mb.b_noise_reduction || !h->fref[0][0]
        || h->mb.i_chroma_qp != h->chroma_qp_table[h->mb.i_qp] )
        return xref_orig_macroblock_probe_pskip( h );
    xref_door_stats[XREF_DOOR_PSKIP][1]++;
    mvx = x264_clip3( h->mb.cache.pskip_mv[0], h->mb.mv_min[0], h->mb.mv_max[0] );
    mvy = x264_clip3( h->mb.cache.pskip_mv[1], h->mb.mv_min[1], h->mb.mv_max[1] );
    if( xref_hook_pskip( h, h->fenc, h->fref[0][0], h->mb.i_mb_x, h->mb.i_mb_y, mvx, mvy, h->mb.i_qp,
                         h->mb.pic.p_fdec[0], h->mb.pic.p_fdec[1], &skip ) )
        return xref_orig_macroblock_probe_pskip( h );          /* declined: a frame is not resident */
    xref_hook_pskip_calls++;
    xref_door_stats[XREF_DOOR_PSKIP][2]++;
    if( skip )
        h->mb.b_skip_mc = 1;                                   /* the prediction in fdec is the reconstruction */
    return skip;
}

/* ------------------------------------------------------------------------------------------------
 * x264_mb_mc                    common/macroblock.c:28       called from x264_macroblock_encode (and analysis)
 *
 * Motion compensation of a P macroblock with any partition the reference analyses: the four 8x8 MVs of the cache go
 * to x264dsp_mc_frames_part_dev on the resident reference frame, the prediction comes back into fdec. */
typedef int (*xref_mbmc_cb)( void *h, void *fref, int mb_x, int mb_y, const int16_t *mv8x8, uint8_t *fdec_y, uint8_t *fdec_c );
xref_mbmc_cb xref_hook_mbmc = NULL;
int xref_hook_mbmc_calls = 0;

void xref_set_mbmc_hook( xref_mbmc_cb cb )
{
    xref_hook_mbmc = cb;
    xref_hook_mbmc_calls = 0;
}
int xref_mbmc_hook_calls( void ) { return xref_hook_mbmc_calls; }

void xref_orig_mb_mc( x264_t *h );

void x264_mb_mc( x264_t *h )
{
    static const int8_t cell[4] = { 0, 2, 16, 18 };            /* x264_scan8[0] + x + (y<<3) for the four 8x8 corners */
    int16_t mv[4][2];
    int k;
    if( xref_hook_mbmc )
        xref_door_stats[XREF_DOOR_MBMC][0]++;
    if( !xref_hook_mbmc || h->sh.i_type != SLICE_TYPE_P || !h->fref[0][0] )
    {
        xref_orig_mb_mc( h );
        return;
    }
    for( k = 0; k < 4; k++ )
    {
        if( h->mb.cache.ref[0][x264_scan8[0] + cell[k]] != 0 )
        {
            xref_orig_mb_mc( h );
            return;
        }
        mv[k][0] = h->mb.cache.mv[0][x264_scan8[0] + cell[k]][0];
        mv[k][1] = h->mb.cache.mv[0][x264_scan8[0] + cell[k]][1];
    }
    xref_door_stats[XREF_DOOR_MBMC][1]++;
    if( xref_hook_mbmc( h, h->fref[0][0], h->mb.i_mb_x, h->mb.i_mb_y, &mv[0][0], h->mb.pic.p_fdec[0], h->mb.pic.p_fdec[1] ) )
    {
        xref_orig_mb_mc( h );
        return;
    }
    xref_hook_mbmc_calls++;
    xref_door_stats[XREF_DOOR_MBMC][2]++;
}
