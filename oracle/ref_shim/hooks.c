/* Driver-level doors of the drop-in proof.  TEST INFRASTRUCTURE / integration example.
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
 * and encoder/slicetype.c's x264_slicetype_decide is wrapped in wrap_slicetype.c.  With no hooks
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

void x264_frame_expand_border_filtered( x264_t *h, x264_frame_t *frame, int mb_y, int b_end )
{
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
    m->mv[0] = out.mv[0];
    m->mv[1] = out.mv[1];
    m->cost = out.cost;
    m->cost_mv = out.cost_mv;
}
