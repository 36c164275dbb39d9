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
 * and encoder/slicetype.c's x264_slicetype_decide is wrapped in wrap_slicetype.c.  With no hooks
 * installed every definition forwards to the original, so the library and the CLI behave exactly like
 * the reference (tests/test_golden.py::test_reference_cli_bitstream pins that).  With hooks installed
 * (tests/test_gpu_dropin_drivers.py) the planes and the lookahead costs come from libx264dsp_b200.so;
 * this is the glue a maintainer of the reference would add.
 */
#include "common/common.h"

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
