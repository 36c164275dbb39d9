/* Test harness around the UNMODIFIED reference (colin121/x264-dsp) C path.
 *
 * Compiled together with the reference sources where they lie under
 * /root/reference into oracle/_ref/libx264ref.so (see oracle/Makefile).  It opens a
 * real encoder instance with x264_encoder_open, so every table (pixf, dctf,
 * quantf, mc, loopf, cost_mv, quant4_mf ...) is the one the reference itself
 * would use, and then exposes plain C entry points that drive the hot-path
 * functions on caller supplied data.  None of the reference's arithmetic is
 * restated here: this file only moves bytes in and out.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs; never by the product library.
 */
#include "common/common.h"
#include "encoder/macroblock.h"
#include "encoder/me.h"
#include "encoder/analyse.h"
#include <time.h>

int xref_slicetype_frame_cost( x264_t *h, x264_frame_t **frames, int p0, int p1, int b );

static void xref_quiet_log( void *priv, int level, const char *fmt, va_list ap )
{
    (void)priv; (void)level; (void)fmt; (void)ap;
}

/* ------------------------------------------------------------------ open/close */

void *xref_open( int width, int height, int me_method, int subme, int me_range, int qp, int psub16x16 )
{
    x264_param_t param;
    x264_param_default( &param );
    param.i_width = width;
    param.i_height = height;
    param.i_csp = X264_CSP_I420;
    param.pf_log = xref_quiet_log;
    param.i_log_level = X264_LOG_NONE;
    param.analyse.i_me_method = me_method;
    param.analyse.i_subpel_refine = subme;
    param.analyse.i_me_range = me_range;
    if( psub16x16 )
        param.analyse.inter |= X264_ANALYSE_PSUB16x16;
    param.rc.i_rc_method = X264_RC_CQP;
    param.rc.i_qp_constant = qp;
    /* keep lowres planes even under CQP: scenecut stays at its default (>0) */
    return x264_encoder_open( &param );
}

/* key-frame parameters of the next xref_open_ex (0 = the reference's defaults) */
static int xref_open_keyint_max = 0, xref_open_keyint_min = 0, xref_open_scenecut = 0;
void xref_set_keyint( int keyint_max, int keyint_min, int scenecut )
{
    xref_open_keyint_max = keyint_max;
    xref_open_keyint_min = keyint_min;
    xref_open_scenecut = scenecut;
}

/* the same with the in-loop deblocking filter switched on or off (a frame's reconstruction is then final when its last
 * macroblock row is coded, which is what the P-slice analysis pin compares) */
void *xref_open_ex( int width, int height, int me_method, int subme, int me_range, int qp, int psub16x16, int deblock )
{
    x264_param_t param;
    x264_param_default( &param );
    param.i_width = width;
    param.i_height = height;
    param.i_csp = X264_CSP_I420;
    param.pf_log = xref_quiet_log;
    param.i_log_level = X264_LOG_NONE;
    param.analyse.i_me_method = me_method;
    param.analyse.i_subpel_refine = subme;
    param.analyse.i_me_range = me_range;
    if( psub16x16 )
        param.analyse.inter |= X264_ANALYSE_PSUB16x16;
    param.rc.i_rc_method = X264_RC_CQP;
    param.rc.i_qp_constant = qp;
    param.b_deblocking_filter = deblock;
    if( xref_open_keyint_max > 0 )
    {
        param.i_keyint_max = xref_open_keyint_max;
        param.i_keyint_min = xref_open_keyint_min;
        param.i_scenecut_threshold = xref_open_scenecut;
    }
    return x264_encoder_open( &param );
}

/* What the encoder holds at the end of a frame's macroblock loop (call from the observer of x264dsp_doors.c): the inputs
 * x264_macroblock_analyse / x264_macroblock_encode worked from and the decisions they left behind. */
typedef struct
{
    int32_t slice_type, qp, poc, ref_poc, inv_ref_poc, ref_is_inter, mv_range, b4_stride, have_lowres_mv, mb_count;
    int32_t fast_pskip, i_frame, frame_type;      /* frame_type: fenc->i_type (X264_TYPE_IDR / I / P) */
    void *fenc, *fref, *fdec;
    const int8_t *mb_type;          /* h->mb.type */
    const int16_t *mvr;             /* h->mb.mvr[0][0] = fdec->mv16x16 */
    const int16_t *cbp;             /* h->mb.cbp */
    const int16_t *mv4x4;           /* fdec->mv[0], one per 4x4, row stride b4_stride */
    const int16_t *lowres_mv;       /* fenc->lowres_mvs[0][0] */
    const int16_t *l0_mv16;         /* fref->mv16x16 */
    const uint8_t *partition;       /* h->mb.partition */
    const uint8_t *nnz;             /* h->mb.non_zero_count, 48 per macroblock */
    const uint8_t *mvd;             /* h->mb.mvd[0]: [mb][8][2], min(|mvd|, 66) of the bottom row / right column (CABAC) */
    int32_t keyint_max, keyint_min, scenecut, icost, pcost, pad1;
    const int8_t *i4_edge_modes;    /* h->mb.intra4x4_pred_mode: [mb][8] = blocks 10 11 14 15 | 5 7 13 | pad */
    const int8_t *chroma_pred_mode; /* h->mb.chroma_pred_mode (CABAC) */
} xref_frame_capture_t;

void xref_capture_frame( void *hv, xref_frame_capture_t *o )
{
    x264_t *h = hv;
    x264_frame_t *fref = h->i_ref[0] > 0 ? h->fref[0][0] : NULL;
    memset( o, 0, sizeof(*o) );
    o->slice_type = h->sh.i_type;
    o->qp = h->sh.i_qp;
    o->poc = h->fdec->i_poc;
    o->mv_range = h->param.analyse.i_mv_range;
    o->b4_stride = h->mb.i_b4_stride;
    o->mb_count = h->mb.i_mb_count;
    o->fast_pskip = h->param.analyse.b_fast_pskip;
    o->i_frame = h->fenc->i_frame;
    o->frame_type = h->fenc->i_type;
    o->fenc = h->fenc;
    o->fdec = h->fdec;
    o->fref = fref;
    o->mb_type = h->mb.type;
    o->mvr = &h->mb.mvr[0][0][0][0];
    o->cbp = h->mb.cbp;
    o->mv4x4 = &h->fdec->mv[0][0][0];
    o->partition = h->mb.partition;
    o->nnz = &h->mb.non_zero_count[0][0];
    o->mvd = h->param.b_cabac ? &h->mb.mvd[0][0][0][0] : NULL;
    o->i4_edge_modes = &h->mb.intra4x4_pred_mode[0][0];
    o->chroma_pred_mode = h->mb.chroma_pred_mode;
    o->keyint_max = h->param.i_keyint_max;
    o->keyint_min = h->param.i_keyint_min;
    o->scenecut = h->param.i_scenecut_threshold;
    o->icost = h->fenc->i_cost_est[0][0];
    o->pcost = h->fenc->i_cost_est[1][0];
    o->lowres_mv = &h->fenc->lowres_mvs[0][0][0][0];
    if( fref )
    {
        o->ref_poc = fref->i_poc;
        o->inv_ref_poc = fref->inv_ref_poc[0];
        o->ref_is_inter = fref->i_ref[0] > 0;
        o->l0_mv16 = &fref->mv16x16[0][0];
        o->have_lowres_mv = h->frames.b_have_lowres && h->fenc->i_frame - fref->i_frame - 1 <= h->param.i_bframe
                            && h->fenc->lowres_mvs[0][h->fenc->i_frame - fref->i_frame - 1][0][0] != 0x7fff;
    }
}

void xref_close( void *hv )
{
    /* x264_encoder_close prints stats and frees; leaking a test encoder is harmless
     * but closing keeps valgrind/ASan runs clean. */
    x264_encoder_close( (x264_t *)hv );
}

/* out[0..15]: geometry the oracle's own layout code must reproduce */
void xref_geometry( void *hv, int *out )
{
    x264_t *h = hv;
    x264_frame_t *f = h->fdec;
    out[0] = h->mb.i_mb_width;
    out[1] = h->mb.i_mb_height;
    out[2] = f->i_stride[0];
    out[3] = f->i_width[0];
    out[4] = f->i_lines[0];
    out[5] = f->i_stride_lowres;
    out[6] = f->i_width_lowres;
    out[7] = f->i_lines_lowres;
    out[8] = (int)(f->filtered[0][1] - f->filtered[0][0]);   /* luma_plane_size */
    out[9] = (int)(f->plane[0] - f->buffer[0]);              /* origin offset in the buffer */
    out[10] = (int)(f->plane[1] - f->buffer[1]);
    out[11] = f->i_stride[1];
    out[12] = f->i_lines[1];
    out[13] = h->param.analyse.i_me_range;
    out[14] = h->param.analyse.i_subpel_refine;
    out[15] = h->param.analyse.i_me_method;
}

/* ------------------------------------------------------------------ tables */

void *xref_pixf( void *hv )    { return &((x264_t *)hv)->pixf; }
void *xref_dctf( void *hv )    { return &((x264_t *)hv)->dctf; }
void *xref_zigzagf( void *hv ) { return &((x264_t *)hv)->zigzagf; }
void *xref_mcf( void *hv )     { return &((x264_t *)hv)->mc; }
void *xref_quantf( void *hv )  { return &((x264_t *)hv)->quantf; }
void *xref_loopf( void *hv )   { return &((x264_t *)hv)->loopf; }
int xref_table_sizes( int *out )
{
    out[0] = sizeof(x264_pixel_function_t);
    out[1] = sizeof(x264_dct_function_t);
    out[2] = sizeof(x264_zigzag_function_t);
    out[3] = sizeof(x264_mc_functions_t);
    out[4] = sizeof(x264_quant_function_t);
    out[5] = sizeof(x264_deblock_function_t);
    out[6] = sizeof(x264_me_t);
    out[7] = sizeof(x264_weight_t);
    return 8;
}

/* centre pointer of cost_mv[qp]; valid for indices -4096..4096 */
const uint16_t *xref_cost_mv( void *hv, int qp ) { return ((x264_t *)hv)->cost_mv[qp]; }
int xref_lambda( int qp ) { return x264_lambda_tab[qp]; }
int xref_chroma_qp( void *hv, int qp ) { return ((x264_t *)hv)->chroma_qp_table[qp]; }

void xref_quant_tables( void *hv, int cat, int qp, uint16_t *mf, uint16_t *bias )
{
    x264_t *h = hv;
    memcpy( mf, h->quant4_mf[cat][qp], 16 * sizeof(uint16_t) );
    memcpy( bias, h->quant4_bias[cat][qp], 16 * sizeof(uint16_t) );
}
void xref_dequant_table( void *hv, int cat, int *out )
{
    x264_t *h = hv;
    memcpy( out, h->dequant4_mf[cat], 6 * 16 * sizeof(int) );
}

/* x264_predict_8x8c_{dc,h,v}_c (common/predict.c:224-288), mode 0=DC 1=H 2=V */
void xref_predict_8x8c( void *hv, int mode, uint8_t *src_fdec )
{
    x264_t *h = hv;
    static const int map[3] = { I_PRED_CHROMA_DC, I_PRED_CHROMA_H, I_PRED_CHROMA_V };
    h->predict_8x8c[map[mode]]( src_fdec );
}

/* ------------------------------------------------------------------ frames */

void *xref_frame_new( void *hv, int b_fdec )
{
    x264_t *h = hv;
    x264_frame_t *f = x264_frame_pop_unused( h, b_fdec );
    int i;
    if( !f )
        return NULL;
    /* the reference mallocs planes without clearing them; zero them so that every
     * byte a comparison can see (alignment gaps included) is deterministic */
    {
        int luma_rows = f->i_lines[0] + 2*PADV;
        int64_t plane_size = (int64_t)f->i_stride[0] * luma_rows;
        if( !(plane_size & 1023) ) plane_size += 128;
        memset( f->buffer[0], 0, (f->filtered[0][1] ? 4 : 1) * plane_size );
        memset( f->buffer[1], 0, (size_t)f->i_stride[1] * (f->i_lines[1] + PADV) );
        if( f->buffer_lowres[0] )
        {
            int64_t lsize = (int64_t)f->i_stride_lowres * (f->i_lines_lowres + 2*PADV);
            if( !(lsize & 1023) ) lsize += 128;
            memset( f->buffer_lowres[0], 0, 4 * lsize );
        }
    }
    if( !b_fdec && f->lowres_mv_costs[0][0] )
        for( i = 0; i < h->mb.i_mb_count; i++ )
            f->lowres_mv_costs[0][0][i] = 0;
    return f;
}

void xref_frame_release( void *hv, void *fv )
{
    x264_frame_push_unused( (x264_t *)hv, (x264_frame_t *)fv );
}

/* planar I420 in, exactly as x264_encoder_encode does it (encoder.c:1745-1770 region):
 * x264_frame_copy_picture then x264_frame_expand_border_mod16 */
void xref_frame_load_i420( void *hv, void *fv, uint8_t *y, uint8_t *u, uint8_t *v )
{
    x264_t *h = hv;
    x264_frame_t *f = fv;
    x264_picture_t pic;
    x264_picture_init( &pic );
    pic.img.i_csp = X264_CSP_I420;
    pic.img.i_plane = 3;
    pic.img.plane[0] = y;
    pic.img.plane[1] = u;
    pic.img.plane[2] = v;
    pic.img.i_stride[0] = h->param.i_width;
    pic.img.i_stride[1] = h->param.i_width >> 1;
    pic.img.i_stride[2] = h->param.i_width >> 1;
    x264_frame_copy_picture( h, f, &pic );
    if( h->param.i_width & 15 || h->param.i_height & 15 )
        x264_frame_expand_border_mod16( h, f );
}

/* luma only, for callers that want nothing but the lookahead of a picture (it never reads chroma): the luma part of
 * x264_frame_copy_picture (h->mc.plane_copy, common/frame.c:227) + x264_frame_expand_border_mod16.  This is what the
 * GPU arm's end-to-end call uploads, so the reference arm of bench.py is not charged for a chroma copy it does not need. */
void xref_frame_load_luma( void *hv, void *fv, uint8_t *y )
{
    x264_t *h = hv;
    x264_frame_t *f = fv;
    h->mc.plane_copy( f->plane[0], f->i_stride[0], y, h->param.i_width, h->param.i_width, h->param.i_height );
    if( h->param.i_width & 15 || h->param.i_height & 15 )
        x264_frame_expand_border_mod16( h, f );
}

void xref_frame_init_lowres( void *hv, void *fv )
{
    x264_frame_init_lowres( (x264_t *)hv, (x264_frame_t *)fv );
}

/* which: 0 plane[0]  1 plane[1]  2..4 filtered[0][1..3]  5..8 lowres[0..3]
 *        10 buffer[0]  11 buffer[1]  12 buffer_lowres[0] */
uint8_t *xref_frame_ptr( void *fv, int which )
{
    x264_frame_t *f = fv;
    switch( which )
    {
        case 0:  return f->plane[0];
        case 1:  return f->plane[1];
        case 2: case 3: case 4: return f->filtered[0][which-1];
        case 5: case 6: case 7: case 8: return f->lowres[which-5];
        case 10: return f->buffer[0];
        case 11: return f->buffer[1];
        case 12: return f->buffer_lowres[0];
    }
    return NULL;
}

/* The in-loop filter sequence of x264_fdec_filter_row (encoder/encoder.c:1359-1385)
 * WITHOUT deblocking, row by row over the whole frame:
 * expand_border -> frame_filter (hpel) -> expand_border_filtered */
void xref_frame_filter_all( void *hv, void *fv )
{
    x264_t *h = hv;
    x264_frame_t *f = fv;
    int mb_y;
    for( mb_y = 1; mb_y <= h->mb.i_mb_height; mb_y++ )
    {
        int min_y = mb_y - 1;
        int end = mb_y == h->mb.i_mb_height;
        x264_frame_expand_border( h, f, min_y );
        x264_frame_filter( h, f, min_y, end );
        x264_frame_expand_border_filtered( h, f, min_y, end );
    }
}

/* border expansion of the unfiltered planes only (luma + NV12 chroma) */
void xref_frame_expand_border_all( void *hv, void *fv )
{
    x264_t *h = hv;
    int mb_y;
    for( mb_y = 0; mb_y < h->mb.i_mb_height; mb_y++ )
        x264_frame_expand_border( h, (x264_frame_t *)fv, mb_y );
}

/* ------------------------------------------------------------------ lookahead */

int xref_frame_cost( void *hv, void **frames, int p0, int p1, int b )
{
    return xref_slicetype_frame_cost( (x264_t *)hv, (x264_frame_t **)frames, p0, p1, b );
}

/* results of frame_cost for distance d = b-p0 (P frames: p1 == b):
 * mvs[mb_count][2], costs[mb_count], sums = { i_cost_est[d][0], i_cost_est_aq[d][0],
 * i_cost_est[0][0], i_intra_mbs[d], b_intra_calculated } */
void xref_frame_lowres_results( void *hv, void *fv, int d, int16_t *mvs, int *costs, int *sums )
{
    x264_t *h = hv;
    x264_frame_t *f = fv;
    if( d > 0 )
    {
        memcpy( mvs, f->lowres_mvs[0][d-1], 2 * h->mb.i_mb_count * sizeof(int16_t) );
        memcpy( costs, f->lowres_mv_costs[0][d-1], h->mb.i_mb_count * sizeof(int) );
    }
    sums[0] = f->i_cost_est[d][0];
    sums[1] = f->i_cost_est_aq[d][0];
    sums[2] = f->i_cost_est[0][0];
    sums[3] = f->i_intra_mbs[d];
    sums[4] = f->b_intra_calculated;
}

/* ------------------------------------------------------------------ motion search */

typedef struct
{
    int32_t i_pixel;          /* PIXEL_16x16 .. PIXEL_4x4 */
    int32_t bx, by;           /* luma position of the block's top-left sample */
    int16_t mvp[2];
    int32_t i_mvc;
    int16_t mvc[16][2];
    int32_t mv_min_fpel[2], mv_max_fpel[2];
    int32_t mv_min_spel[2], mv_max_spel[2];
} xref_me_in_t;

typedef struct
{
    int16_t mv[2];
    int32_t cost;
    int32_t cost_mv;
} xref_me_out_t;

/* x264_me_search_ref (encoder/me.c:129) on a list of blocks.
 * fenc: any frame (source samples from plane[0]); fref: an fdec frame whose
 * filtered[0][0..3] planes have been produced by xref_frame_filter_all.
 * refine != 0 additionally runs x264_me_refine_qpel on each result (me.c:426). */
void xref_me_search_batch_ex( void *hv, void *fencv, void *frefv, int qp, int me_method, int subme,
                              int me_range, int refine, const xref_me_in_t *in, int n, xref_me_out_t *out,
                              int mode, int *thresh );
void xref_me_search_batch( void *hv, void *fencv, void *frefv, int qp, int me_method, int subme,
                           int me_range, int refine, const xref_me_in_t *in, int n, xref_me_out_t *out )
{
    xref_me_search_batch_ex( hv, fencv, frefv, qp, me_method, subme, me_range, refine, in, n, out, 0, NULL );
}

/* mode 0: x264_me_search_ref with p_halfpel_thresh = &thresh[i] (thresh == NULL: NULL);
 * mode 1: x264_me_refine_qpel_refdupe (me.c:437) and mode 2: x264_me_refine_qpel (me.c:426), both starting
 * from the mv / cost / cost_mv found in out[i]. */
void xref_me_search_batch_ex( void *hv, void *fencv, void *frefv, int qp, int me_method, int subme,
                              int me_range, int refine, const xref_me_in_t *in, int n, xref_me_out_t *out,
                              int mode, int *thresh )
{
    x264_t *h = hv;
    x264_frame_t *fenc = fencv, *fref = frefv;
    int stride = fref->i_stride[0];
    int i, k, y;
    int keep_range = h->param.analyse.i_me_range;
    h->param.analyse.i_me_range = me_range;
    h->mb.i_me_method = me_method;
    h->mb.i_subpel_refine = subme;
    h->mb.b_chroma_me = 0;
    for( i = 0; i < n; i++ )
    {
        x264_me_t m;
        int16_t mvc[16][2];
        int bw = x264_pixel_size[in[i].i_pixel].w;
        int bh = x264_pixel_size[in[i].i_pixel].h;
        pixel *src = fenc->plane[0] + in[i].by * fenc->i_stride[0] + in[i].bx;
        pixel *dst = h->mb.pic.fenc_buf;
        memset( &m, 0, sizeof(m) );
        for( y = 0; y < bh; y++ )
            memcpy( dst + y*FENC_STRIDE, src + y*fenc->i_stride[0], bw );
        for( k = 0; k < 2; k++ )
        {
            h->mb.mv_min_fpel[k] = in[i].mv_min_fpel[k];
            h->mb.mv_max_fpel[k] = in[i].mv_max_fpel[k];
            h->mb.mv_min_spel[k] = in[i].mv_min_spel[k];
            h->mb.mv_max_spel[k] = in[i].mv_max_spel[k];
        }
        m.i_pixel = in[i].i_pixel;
        m.p_cost_mv = h->cost_mv[qp];
        m.i_ref_cost = 0;
        m.i_ref = 0;
        m.weight = x264_weight_none;
        for( k = 0; k < 4; k++ )
            m.p_fref[k] = fref->filtered[0][k] + in[i].by * stride + in[i].bx;
        m.p_fref_w = m.p_fref[0];
        m.p_fenc[0] = dst;
        m.i_stride[0] = stride;
        m.mvp[0] = in[i].mvp[0];
        m.mvp[1] = in[i].mvp[1];
        memcpy( mvc, in[i].mvc, sizeof(mvc) );
        if( mode )
        {
            m.mv[0] = out[i].mv[0];
            m.mv[1] = out[i].mv[1];
            m.cost = out[i].cost;
            m.cost_mv = out[i].cost_mv;
            if( mode == 1 )
                x264_me_refine_qpel_refdupe( h, &m, thresh ? &thresh[i] : NULL );
            else
                x264_me_refine_qpel( h, &m );
        }
        else
        {
            x264_me_search_ref( h, &m, mvc, in[i].i_mvc, thresh ? &thresh[i] : NULL );
            if( refine )
                x264_me_refine_qpel( h, &m );
        }
        out[i].mv[0] = m.mv[0];
        out[i].mv[1] = m.mv[1];
        out[i].cost = m.cost;
        out[i].cost_mv = m.cost_mv;
    }
    h->param.analyse.i_me_range = keep_range;
}

/* ------------------------------------------------------------------ deblock */

/* x264_macroblock_deblock_strength (common/macroblock.c:677) on a caller-filled cache; bs is read and written */
void xref_macroblock_deblock_strength( void *hv, int mb_type, const uint8_t *nnz, const int8_t *ref, const int16_t *mv,
                                       uint8_t *bs )
{
    x264_t *h = hv;
    uint8_t (*keep)[8][4] = h->mb.cache.deblock_strength;
    h->mb.i_type = mb_type;
    memcpy( h->mb.cache.non_zero_count, nnz, sizeof(h->mb.cache.non_zero_count) );
    memcpy( h->mb.cache.ref, ref, sizeof(h->mb.cache.ref) );
    memcpy( h->mb.cache.mv, mv, sizeof(h->mb.cache.mv) );
    h->mb.cache.deblock_strength = (uint8_t (*)[8][4])bs;
    x264_macroblock_deblock_strength( h );
    h->mb.cache.deblock_strength = keep;
}

/* x264_frame_deblock_row (common/deblock.c:341) over every MB row of frame f.
 * mb_type/partition/cbp: per-MB arrays (raster); bs: [mb][2][8][4] as produced by
 * deblock_strength / x264_macroblock_deblock_strength. */
void xref_deblock_frame( void *hv, void *fv, const int8_t *mb_type, const uint8_t *partition,
                         const int16_t *cbp, const uint8_t *bs, int qp, int alpha_off, int beta_off )
{
    x264_t *h = hv;
    x264_frame_t *keep = h->fdec;
    int8_t *keep_type = h->mb.type;
    uint8_t *keep_part = h->mb.partition;
    int mb_y, n = h->mb.i_mb_count;
    h->fdec = fv;
    h->mb.i_mb_stride = h->mb.i_mb_width;
    h->mb.type = h->fdec->mb_type;
    h->mb.partition = h->fdec->mb_partition;
    memcpy( h->mb.type, mb_type, n );
    memcpy( h->mb.partition, partition, n );
    memcpy( h->mb.cbp, cbp, n * sizeof(int16_t) );
    h->sh.i_qp = qp;
    h->sh.i_alpha_c0_offset = alpha_off;
    h->sh.i_beta_offset = beta_off;
    for( mb_y = 0; mb_y < h->mb.i_mb_height; mb_y++ )
    {
        memcpy( h->deblock_strength[mb_y&1], bs + (size_t)mb_y * h->mb.i_mb_width * 64,
                (size_t)h->mb.i_mb_width * 64 );
        x264_frame_deblock_row( h, mb_y );
    }
    h->fdec = keep;
    h->mb.type = keep_type;
    h->mb.partition = keep_part;
}

/* ------------------------------------------------------------------ residual

 * x264_macroblock_encode (encoder/macroblock.c:310) for one P_L0 16x16 macroblock whose
 * prediction is already in p_fdec (b_skip_mc = 1).  Buffers use the reference's fenc_buf /
 * fdec_buf shapes: fenc_y 16 rows @16, fenc_c 8 rows @16 (U at +0, V at +8);
 * fdec_y 16 rows @32, fdec_c 8 rows @32 (U at +0, V at +16); fdec is prediction in, recon out.
 * levels: luma4x4[0..15][16], chroma_dc[0..1][4], luma4x4[16..19][16], luma4x4[32..35][16].
 * nnz: 16 luma (coding order), 4 U, 4 V, luma DC, U DC, V DC.  Returns h->mb.cbp. */
int xref_encode_inter_mb( void *hv, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y,
                          uint8_t *fdec_c, int qp, int16_t *levels, uint8_t *nnz )
{
    x264_t *h = hv;
    int i, y;
    h->sh.i_type = SLICE_TYPE_P;
    x264_macroblock_thread_init( h );
    h->mb.i_type = P_L0;
    h->mb.i_partition = D_16x16;
    h->mb.b_skip_mc = 1;
    h->mb.b_noise_reduction = 0;
    h->mb.b_transform_8x8 = 0;
    h->nr_count = h->nr_count_buf[0];
    h->mb.i_qp = qp;
    h->mb.i_chroma_qp = h->chroma_qp_table[qp];
    h->mb.i_mb_xy = 0;
    h->mb.cache.mv[0][x264_scan8[0]][0] = 1;       /* != pskip_mv: keep the type P_L0 */
    h->mb.cache.mv[0][x264_scan8[0]][1] = 1;
    M32( h->mb.cache.pskip_mv ) = 0;
    memset( &h->dct, 0, sizeof(h->dct) );
    memset( h->mb.cache.non_zero_count, 0, sizeof(h->mb.cache.non_zero_count) );
    for( y = 0; y < 16; y++ )
    {
        memcpy( h->mb.pic.p_fenc[0] + y*FENC_STRIDE, fenc_y + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, fdec_y + y*32, 16 );
    }
    for( y = 0; y < 8; y++ )
    {
        memcpy( h->mb.pic.p_fenc[1] + y*FENC_STRIDE, fenc_c + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, fdec_c + y*32, 8 );
        memcpy( h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, fdec_c + y*32 + 16, 8 );
    }
    x264_macroblock_encode( h );
    for( y = 0; y < 16; y++ )
        memcpy( fdec_y + y*32, h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, 16 );
    for( y = 0; y < 8; y++ )
    {
        memcpy( fdec_c + y*32, h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, 8 );
        memcpy( fdec_c + y*32 + 16, h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, 8 );
    }
    memcpy( levels, h->dct.luma4x4[0], 16*16*sizeof(int16_t) );
    memcpy( levels + 256, h->dct.chroma_dc[0], 4*sizeof(int16_t) );
    memcpy( levels + 260, h->dct.chroma_dc[1], 4*sizeof(int16_t) );
    memcpy( levels + 264, h->dct.luma4x4[16], 4*16*sizeof(int16_t) );
    memcpy( levels + 328, h->dct.luma4x4[32], 4*16*sizeof(int16_t) );
    for( i = 0; i < 16; i++ )
        nnz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
    for( i = 0; i < 4; i++ )
    {
        nnz[16+i] = h->mb.cache.non_zero_count[x264_scan8[16+i]];
        nnz[20+i] = h->mb.cache.non_zero_count[x264_scan8[32+i]];
    }
    nnz[24] = h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]];
    nnz[25] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]];
    nnz[26] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]];
    return h->mb.cbp[0];
}

/* A whole P frame of P_L0 16x16 macroblocks through the reference's own functions -- x264_mb_mc (mc_luma + mc_chroma,
 * common/macroblock.c:8-28), x264_macroblock_encode (encoder/macroblock.c:310) -- the CPU side of SURVEY 8(d) config 4:
 * fenc = source frame, fref = reference frame with its filtered planes, fdec receives the reconstruction (luma plane +
 * NV12 chroma, not yet deblocked; xref_deblock_frame does that).  mv16 [mb][2]; levels [mb][392], nnz [mb][27],
 * cbp [mb] as in xref_encode_inter_mb. */
void xref_recon_frame( void *hv, void *fencv, void *frefv, void *fdecv, const int16_t *mv16, int qp,
                       int16_t *levels, uint8_t *nnz, int16_t *cbp )
{
    x264_t *h = hv;
    x264_frame_t *fenc = fencv, *fref = frefv, *fdec = fdecv;
    const int W = h->mb.i_mb_width, H = h->mb.i_mb_height;
    const int ls = fref->i_stride[0], cs = fref->i_stride[1];
    int mb_x, mb_y, i, y;
    h->sh.i_type = SLICE_TYPE_P;
    x264_macroblock_thread_init( h );
    h->mb.b_noise_reduction = 0;
    h->mb.b_transform_8x8 = 0;
    h->nr_count = h->nr_count_buf[0];
    h->mb.i_qp = qp;
    h->mb.i_chroma_qp = h->chroma_qp_table[qp];
    h->mb.pic.i_stride[0] = ls;
    h->mb.pic.i_stride[1] = cs;
    for( mb_y = 0; mb_y < H; mb_y++ )
        for( mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            const intptr_t oy = (intptr_t)( mb_y << 4 ) * ls + ( mb_x << 4 ), oc = (intptr_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
            int16_t *lv = levels + (size_t)xy * 392;
            uint8_t *nz = nnz + (size_t)xy * 27;
            h->mb.i_mb_x = mb_x;
            h->mb.i_mb_y = mb_y;
            h->mb.i_mb_xy = 0;                      /* h->mb.cbp[] etc. are per-frame arrays: slot 0 is scratch here */
            h->mb.i_type = P_L0;
            h->mb.i_partition = D_16x16;
            h->mb.b_skip_mc = 0;
            h->mb.mv_min[0] = ( -( mb_x << 4 ) - 24 ) << 2;                                   /* analyse.c:378-393 */
            h->mb.mv_max[0] = ( ( ( W - mb_x - 1 ) << 4 ) + 24 ) << 2;
            h->mb.mv_min[1] = ( -( mb_y << 4 ) - 24 ) << 2;
            h->mb.mv_max[1] = ( ( ( H - mb_y - 1 ) << 4 ) + 24 ) << 2;
            for( i = 0; i < 4; i++ )
                h->mb.pic.p_fref[0][0][i] = fref->filtered[0][i] + oy;
            h->mb.pic.p_fref[0][0][4] = fref->plane[1] + oc;
            h->mb.cache.ref[0][x264_scan8[0]] = 0;
            h->mb.cache.mv[0][x264_scan8[0]][0] = mv16[2*xy];
            h->mb.cache.mv[0][x264_scan8[0]][1] = mv16[2*xy+1];
            h->mb.cache.pskip_mv[0] = (int16_t)( mv16[2*xy] + 1 );      /* != mv: keep the type P_L0 */
            h->mb.cache.pskip_mv[1] = mv16[2*xy+1];
            memset( &h->dct, 0, sizeof(h->dct) );
            memset( h->mb.cache.non_zero_count, 0, sizeof(h->mb.cache.non_zero_count) );
            for( y = 0; y < 16; y++ )
                memcpy( h->mb.pic.p_fenc[0] + y*FENC_STRIDE, fenc->plane[0] + oy + (intptr_t)y * fenc->i_stride[0], 16 );
            for( y = 0; y < 8; y++ )
            {
                /* NV12 -> the reference's fenc chroma layout: U at +0, V at +8 (x264_macroblock_load_pic_pointers) */
                const pixel *src = fenc->plane[1] + oc + (intptr_t)y * fenc->i_stride[1];
                for( i = 0; i < 8; i++ )
                {
                    h->mb.pic.p_fenc[1][y*FENC_STRIDE + i] = src[2*i];
                    h->mb.pic.p_fenc[2][y*FENC_STRIDE + i] = src[2*i+1];
                }
            }
            x264_mb_mc( h );
            h->mb.b_skip_mc = 1;
            x264_macroblock_encode( h );
            for( y = 0; y < 16; y++ )
                memcpy( fdec->plane[0] + oy + (intptr_t)y * ls, h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, 16 );
            for( y = 0; y < 8; y++ )
            {
                pixel *dst = fdec->plane[1] + oc + (intptr_t)y * cs;
                for( i = 0; i < 8; i++ )
                {
                    dst[2*i] = h->mb.pic.p_fdec[1][y*FDEC_STRIDE + i];
                    dst[2*i+1] = h->mb.pic.p_fdec[2][y*FDEC_STRIDE + i];
                }
            }
            memcpy( lv, h->dct.luma4x4[0], 16*16*sizeof(int16_t) );
            memcpy( lv + 256, h->dct.chroma_dc[0], 4*sizeof(int16_t) );
            memcpy( lv + 260, h->dct.chroma_dc[1], 4*sizeof(int16_t) );
            memcpy( lv + 264, h->dct.luma4x4[16], 4*16*sizeof(int16_t) );
            memcpy( lv + 328, h->dct.luma4x4[32], 4*16*sizeof(int16_t) );
            for( i = 0; i < 16; i++ )
                nz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
            for( i = 0; i < 4; i++ )
            {
                nz[16+i] = h->mb.cache.non_zero_count[x264_scan8[16+i]];
                nz[20+i] = h->mb.cache.non_zero_count[x264_scan8[32+i]];
            }
            nz[24] = h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]];
            nz[25] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]];
            nz[26] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]];
            cbp[xy] = h->mb.cbp[0];
        }
}

/* x264_macroblock_encode for one I16x16 macroblock of an I slice whose prediction (luma and chroma) is
 * already in p_fdec: the predictors the function would call are swapped for no-ops during the call (the
 * tables are plain data members of x264_t), everything else is the reference's own code.  Same buffer
 * shapes as xref_encode_inter_mb; luma_dc receives h->dct.luma16x16_dc[0]. */
static void xref_predict_noop( pixel *src ) { (void)src; }

int xref_encode_intra16_mb( void *hv, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y,
                            uint8_t *fdec_c, int qp, int16_t *levels, int16_t *luma_dc, uint8_t *nnz )
{
    x264_t *h = hv;
    int i, y;
    x264_predict_t keep16 = h->predict_16x16[0], keepc = h->predict_chroma[0];
    h->sh.i_type = SLICE_TYPE_I;
    x264_macroblock_thread_init( h );
    h->mb.i_type = I_16x16;
    h->mb.i_intra16x16_pred_mode = 0;
    h->mb.i_chroma_pred_mode = 0;
    h->mb.b_dct_decimate = 0;
    h->mb.b_noise_reduction = 0;
    h->mb.b_transform_8x8 = 0;
    h->nr_count = h->nr_count_buf[0];
    h->mb.i_qp = qp;
    h->mb.i_chroma_qp = h->chroma_qp_table[qp];
    h->mb.i_mb_xy = 0;
    memset( &h->dct, 0, sizeof(h->dct) );
    memset( h->mb.cache.non_zero_count, 0, sizeof(h->mb.cache.non_zero_count) );
    for( y = 0; y < 16; y++ )
    {
        memcpy( h->mb.pic.p_fenc[0] + y*FENC_STRIDE, fenc_y + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, fdec_y + y*32, 16 );
    }
    for( y = 0; y < 8; y++ )
    {
        memcpy( h->mb.pic.p_fenc[1] + y*FENC_STRIDE, fenc_c + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, fdec_c + y*32, 8 );
        memcpy( h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, fdec_c + y*32 + 16, 8 );
    }
    h->predict_16x16[0] = xref_predict_noop;
    h->predict_chroma[0] = xref_predict_noop;
    x264_macroblock_encode( h );
    h->predict_16x16[0] = keep16;
    h->predict_chroma[0] = keepc;
    for( y = 0; y < 16; y++ )
        memcpy( fdec_y + y*32, h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, 16 );
    for( y = 0; y < 8; y++ )
    {
        memcpy( fdec_c + y*32, h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, 8 );
        memcpy( fdec_c + y*32 + 16, h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, 8 );
    }
    memcpy( levels, h->dct.luma4x4[0], 16*16*sizeof(int16_t) );
    memcpy( levels + 256, h->dct.chroma_dc[0], 4*sizeof(int16_t) );
    memcpy( levels + 260, h->dct.chroma_dc[1], 4*sizeof(int16_t) );
    memcpy( levels + 264, h->dct.luma4x4[16], 4*16*sizeof(int16_t) );
    memcpy( levels + 328, h->dct.luma4x4[32], 4*16*sizeof(int16_t) );
    memcpy( luma_dc, h->dct.luma16x16_dc[0], 16*sizeof(int16_t) );
    for( i = 0; i < 16; i++ )
        nnz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
    for( i = 0; i < 4; i++ )
    {
        nnz[16+i] = h->mb.cache.non_zero_count[x264_scan8[16+i]];
        nnz[20+i] = h->mb.cache.non_zero_count[x264_scan8[32+i]];
    }
    nnz[24] = h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]];
    nnz[25] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]];
    nnz[26] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]];
    return h->mb.cbp[0];
}

/* x264_macroblock_encode for one I4x4 macroblock of an I slice.  fdec_y points at the macroblock origin inside a
 * caller buffer of stride 32 that holds the reconstructed neighbours (row -1 from column -1 to 19, column -1 of rows
 * 0..15), exactly what the reference's fdec_buf holds at that point; the 4x4 predictors and everything else are the
 * reference's own, the chroma prediction is the caller's (predict_chroma swapped for a no-op as above).
 * modes[16] = h->mb.cache.intra4x4_pred_mode in coding order; replicate5: block 5 lacks its top-right samples. */
int xref_encode_intra4_mb( void *hv, const uint8_t *fenc_y, const uint8_t *fenc_c, uint8_t *fdec_y, uint8_t *fdec_c,
                           int qp, const uint8_t *modes, int replicate5, int16_t *levels, uint8_t *nnz )
{
    x264_t *h = hv;
    int i, y;
    const int all = MB_LEFT|MB_TOP|MB_TOPLEFT|MB_TOPRIGHT;
    x264_predict_t keepc = h->predict_chroma[0];
    h->sh.i_type = SLICE_TYPE_I;
    x264_macroblock_thread_init( h );
    h->mb.i_type = I_4x4;
    h->mb.i_skip_intra = 0;
    h->mb.i_chroma_pred_mode = 0;
    h->mb.b_dct_decimate = 0;
    h->mb.b_noise_reduction = 0;
    h->mb.b_transform_8x8 = 0;
    h->nr_count = h->nr_count_buf[0];
    h->mb.i_qp = qp;
    h->mb.i_chroma_qp = h->chroma_qp_table[qp];
    h->mb.i_mb_xy = 0;
    h->mb.i_neighbour4[0] = h->mb.i_neighbour4[1] = h->mb.i_neighbour4[2] = h->mb.i_neighbour4[4] =
    h->mb.i_neighbour4[8] = h->mb.i_neighbour4[10] = all;
    h->mb.i_neighbour4[5] = replicate5 ? MB_LEFT|MB_TOP|MB_TOPLEFT : all;
    /* what x264_macroblock_slice_init sets once per slice (common/macroblock.c:217-225) */
    h->mb.i_neighbour4[6] = h->mb.i_neighbour4[9] = h->mb.i_neighbour4[12] = h->mb.i_neighbour4[14] = all;
    h->mb.i_neighbour4[3] = h->mb.i_neighbour4[7] = h->mb.i_neighbour4[11] = h->mb.i_neighbour4[13] =
    h->mb.i_neighbour4[15] = MB_LEFT|MB_TOP|MB_TOPLEFT;
    for( i = 0; i < 16; i++ )
        h->mb.cache.intra4x4_pred_mode[x264_scan8[i]] = (int8_t)modes[i];
    memset( &h->dct, 0, sizeof(h->dct) );
    memset( h->mb.cache.non_zero_count, 0, sizeof(h->mb.cache.non_zero_count) );
    memcpy( h->mb.pic.p_fdec[0] - FDEC_STRIDE - 1, fdec_y - 32 - 1, 21 );
    for( y = 0; y < 16; y++ )
    {
        memcpy( h->mb.pic.p_fenc[0] + y*FENC_STRIDE, fenc_y + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[0] + y*FDEC_STRIDE - 1, fdec_y + y*32 - 1, 17 );
    }
    for( y = 0; y < 8; y++ )
    {
        memcpy( h->mb.pic.p_fenc[1] + y*FENC_STRIDE, fenc_c + y*16, 16 );
        memcpy( h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, fdec_c + y*32, 8 );
        memcpy( h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, fdec_c + y*32 + 16, 8 );
    }
    h->predict_chroma[0] = xref_predict_noop;
    x264_macroblock_encode( h );
    h->predict_chroma[0] = keepc;
    for( y = 0; y < 16; y++ )
        memcpy( fdec_y + y*32, h->mb.pic.p_fdec[0] + y*FDEC_STRIDE, 16 );
    for( y = 0; y < 8; y++ )
    {
        memcpy( fdec_c + y*32, h->mb.pic.p_fdec[1] + y*FDEC_STRIDE, 8 );
        memcpy( fdec_c + y*32 + 16, h->mb.pic.p_fdec[2] + y*FDEC_STRIDE, 8 );
    }
    memcpy( levels, h->dct.luma4x4[0], 16*16*sizeof(int16_t) );
    memcpy( levels + 256, h->dct.chroma_dc[0], 4*sizeof(int16_t) );
    memcpy( levels + 260, h->dct.chroma_dc[1], 4*sizeof(int16_t) );
    memcpy( levels + 264, h->dct.luma4x4[16], 4*16*sizeof(int16_t) );
    memcpy( levels + 328, h->dct.luma4x4[32], 4*16*sizeof(int16_t) );
    for( i = 0; i < 16; i++ )
        nnz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
    for( i = 0; i < 4; i++ )
    {
        nnz[16+i] = h->mb.cache.non_zero_count[x264_scan8[16+i]];
        nnz[20+i] = h->mb.cache.non_zero_count[x264_scan8[32+i]];
    }
    nnz[24] = h->mb.cache.non_zero_count[x264_scan8[LUMA_DC]];
    nnz[25] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC]];
    nnz[26] = h->mb.cache.non_zero_count[x264_scan8[CHROMA_DC+1]];
    return h->mb.cbp[0];
}

/* x264_macroblock_probe_pskip (encoder/macroblock.c:492) on a caller-supplied P_SKIP prediction: the function's
 * own motion compensation calls (h->mc.mc_luma / mc_chroma, plain data members of x264_t) are pointed at two stand-ins
 * that deliver the stashed prediction, everything after that -- transforms, quantisation, decimation, the chroma
 * SSD / DC / AC ladder -- is the reference's code.  Buffer shapes as in xref_encode_inter_mb.  Returns its result. */
static const uint8_t *xref_stash_y, *xref_stash_c;
static void xref_stash_mc_luma( pixel *dst, intptr_t i_dst, pixel **src, intptr_t i_src, int mvx, int mvy, int w, int ht,
                                const x264_weight_t *weight )
{
    int y;
    (void)src; (void)i_src; (void)mvx; (void)mvy; (void)weight;
    for( y = 0; y < ht; y++ )
        memcpy( dst + y*i_dst, xref_stash_y + y*32, w );
}
static void xref_stash_mc_chroma( pixel *dstu, pixel *dstv, intptr_t i_dst, pixel *src, intptr_t i_src, int mvx, int mvy,
                                  int w, int ht )
{
    int y;
    (void)src; (void)i_src; (void)mvx; (void)mvy;
    for( y = 0; y < ht; y++ )
    {
        memcpy( dstu + y*i_dst, xref_stash_c + y*32, w );
        memcpy( dstv + y*i_dst, xref_stash_c + y*32 + 16, w );
    }
}

int xref_probe_pskip_mb( void *hv, const uint8_t *fenc_y, const uint8_t *fenc_c, const uint8_t *pred_y,
                         const uint8_t *pred_c, int qp )
{
    x264_t *h = hv;
    x264_mc_functions_t keep = h->mc;
    static pixel dummy[64];
    int y, r;
    h->sh.i_type = SLICE_TYPE_P;
    x264_macroblock_thread_init( h );
    h->mb.b_noise_reduction = 0;
    h->mb.i_qp = qp;
    h->mb.i_chroma_qp = h->chroma_qp_table[qp];
    h->mb.cache.pskip_mv[0] = 4;                   /* non-zero: the chroma prediction comes through mc_chroma */
    h->mb.cache.pskip_mv[1] = 4;
    h->mb.mv_min[0] = h->mb.mv_min[1] = -64;
    h->mb.mv_max[0] = h->mb.mv_max[1] = 64;
    for( y = 0; y < 6; y++ )
        h->mb.pic.p_fref[0][0][y] = dummy;
    for( y = 0; y < 16; y++ )
        memcpy( h->mb.pic.p_fenc[0] + y*FENC_STRIDE, fenc_y + y*16, 16 );
    for( y = 0; y < 8; y++ )
        memcpy( h->mb.pic.p_fenc[1] + y*FENC_STRIDE, fenc_c + y*16, 16 );
    xref_stash_y = pred_y;
    xref_stash_c = pred_c;
    h->mc.mc_luma = xref_stash_mc_luma;
    h->mc.mc_chroma = xref_stash_mc_chroma;
    r = x264_macroblock_probe_pskip( h );
    h->mc = keep;
    return r;
}

/* x264_mb_predict_mv_16x16 / x264_mb_predict_mv_pskip (common/mvpred.c:101-155) on a caller-supplied neighbourhood:
 * ref[4] / mv[4][2] = left, top, top-right, top-left as they sit in h->mb.cache around X264_SCAN8_0 */
void xref_predict_mv( void *hv, const int8_t *ref, const int16_t *mv, int i_ref, int16_t *mvp, int16_t *pskip )
{
    x264_t *h = hv;
    static const int cell[4] = { X264_SCAN8_0 - 1, X264_SCAN8_0 - 8, X264_SCAN8_0 - 8 + 4, X264_SCAN8_0 - 8 - 1 };
    int k;
    for( k = 0; k < 4; k++ )
    {
        h->mb.cache.ref[0][cell[k]] = ref[k];
        h->mb.cache.mv[0][cell[k]][0] = mv[2*k];
        h->mb.cache.mv[0][cell[k]][1] = mv[2*k+1];
    }
    x264_mb_predict_mv_16x16( h, 0, i_ref, mvp );
    x264_mb_predict_mv_pskip( h, pskip );
}

/* x264_mb_predict_mv (common/mvpred.c:22) for partition `idx` of width `i_width` (in 4-pixel units) of a macroblock
 * whose partition type is `partition` (D_16x8 / D_8x16 / D_8x8 / D_16x16): the neighbours go into the cache cells the
 * function reads -- left, top, top-right, top-left of x264_scan8[idx] -- and the partition's own reference into its cell */
void xref_predict_mv_part( void *hv, const int8_t *ref, const int16_t *mv, int i_ref, int partition, int idx, int i_width,
                           int16_t *mvp )
{
    x264_t *h = hv;
    const int i8 = x264_scan8[idx];
    const int cell[4] = { i8 - 1, i8 - 8, i8 - 8 + i_width, i8 - 8 - 1 };
    int k;
    for( k = 0; k < 4; k++ )
    {
        h->mb.cache.ref[0][cell[k]] = ref[k];
        h->mb.cache.mv[0][cell[k]][0] = mv[2*k];
        h->mb.cache.mv[0][cell[k]][1] = mv[2*k+1];
    }
    h->mb.cache.ref[0][i8] = (int8_t)i_ref;
    h->mb.i_partition = partition;
    x264_mb_predict_mv( h, 0, idx, i_width, mvp );
}

/* x264_mb_predict_mv_ref16x16 (common/mvpred.c:167) for every macroblock of the encoder's frame size, on caller data:
 * lowres_mv (or NULL), mvr = the current frame's 16x16 MVs, l0_mv16 (or NULL: intra reference frame), curpoc / refpoc /
 * inv_ref_poc for the temporal scale.  Three scratch frames of the encoder carry the per-frame fields the function
 * reads; the neighbour indices are set per macroblock the way x264_macroblock_cache_load_neighbours does for a
 * single-slice frame (common/macroblock.c:304-367). */
void xref_predict_mvc_frame( void *hv, void *fenc_v, void *fref_v, void *fdec_v, const int16_t *lowres_mv, const int16_t *mvr,
                             const int16_t *l0_mv16, int curpoc, int refpoc, int inv_ref_poc, int16_t *mvc, int32_t *n_mvc )
{
    x264_t *h = hv;
    x264_frame_t *fenc = fenc_v, *fref = fref_v, *fdec = fdec_v;
    const int W = h->mb.i_mb_width, H = h->mb.i_mb_height, n = W * H;
    x264_frame_t *keep_fenc = h->fenc, *keep_fref = h->fref[0][0], *keep_fdec = h->fdec;
    int16_t (*keep_mvr)[2] = h->mb.mvr[0][0];
    int16_t (*mvr_buf)[2] = malloc( ( n + 1 ) * 4 );
    int x, y;
    memset( mvr_buf, 0, 4 );
    memcpy( mvr_buf + 1, mvr, n * 4 );
    h->fenc = fenc; h->fref[0][0] = fref; h->fdec = fdec;
    h->mb.mvr[0][0] = mvr_buf + 1;
    h->frames.b_have_lowres = 1;
    fenc->i_frame = 5; fref->i_frame = 4;                        /* idx = 0 <= i_bframe */
    if( lowres_mv )
        memcpy( fenc->lowres_mvs[0][0], lowres_mv, n * 4 );
    else
        fenc->lowres_mvs[0][0][0][0] = 0x7fff;
    fref->i_ref[0] = l0_mv16 ? 1 : 0;
    if( l0_mv16 )
        memcpy( fref->mv16x16, l0_mv16, n * 4 );
    fdec->i_poc = curpoc; fdec->i_delta_poc[0] = fdec->i_delta_poc[1] = 0;
    fref->i_poc = refpoc; fref->i_delta_poc[0] = fref->i_delta_poc[1] = 0;
    fref->inv_ref_poc[0] = fref->inv_ref_poc[1] = (int16_t)inv_ref_poc;
    for( y = 0; y < H; y++ )
        for( x = 0; x < W; x++ )
        {
            const int xy = y * W + x;
            h->mb.i_mb_x = x; h->mb.i_mb_y = y; h->mb.i_mb_xy = xy;
            h->mb.i_mb_left_xy[0] = x > 0 ? xy - 1 : -1;
            h->mb.i_mb_top_xy = y > 0 ? xy - W : -1;
            h->mb.i_mb_topleft_xy = ( x > 0 && y > 0 ) ? xy - W - 1 : -1;
            h->mb.i_mb_topright_xy = ( y > 0 && x < W - 1 ) ? xy - W + 1 : -1;
            x264_mb_predict_mv_ref16x16( h, 0, 0, (int16_t (*)[2])( mvc + (size_t)xy * 18 ), &n_mvc[xy] );
        }
    h->fenc = keep_fenc; h->fref[0][0] = keep_fref; h->fdec = keep_fdec;
    h->mb.mvr[0][0] = keep_mvr;
    free( mvr_buf );
}

/* ------------------------------------------------------------------ timing helpers
 * (cpu_baseline / --impl reference): loops over the reference functions with the
 * input already in memory; CLOCK_MONOTONIC around the loop; returns seconds. */

static double xref_now( void )
{
    struct timespec ts;
    clock_gettime( CLOCK_MONOTONIC, &ts );
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* lookahead leg for n frames: init_lowres on every frame, then frame_cost(i-1,i,i) for
 * i = 1..n-1 (frame 0 gets the intra-only call frame_cost(0,0,0)).
 * frames must have been loaded with xref_frame_load_i420. */
double xref_time_lookahead( void *hv, void **frames, int n, int *cost_out )
{
    x264_t *h = hv;
    double t0 = xref_now();
    int i;
    for( i = 0; i < n; i++ )
    {
        /* a recycled frame starts with b_intra_calculated = 0 (x264_frame_pop_unused, frame.c:511) */
        ((x264_frame_t *)frames[i])->b_intra_calculated = 0;
        x264_frame_init_lowres( h, (x264_frame_t *)frames[i] );
    }
    cost_out[0] = xref_slicetype_frame_cost( h, (x264_frame_t **)frames, 0, 0, 0 );
    for( i = 1; i < n; i++ )
        cost_out[i] = xref_slicetype_frame_cost( h, (x264_frame_t **)frames, i-1, i, i );
    return xref_now() - t0;
}

/* ------------------------------------------------------------------ drop-in table test
 * Replace the six function-pointer tables of an open encoder by caller-supplied ones (the product's
 * x264_*_init output) and redo the aliasing x264_encoder_open performs after the init calls
 * (mbcmp_init / chroma_dsp_init, encoder/encoder.c:412-457, both static there). */
void xref_install_tables( void *hv, const void *pixf, const void *dctf, const void *zigzagf,
                          const void *mcf, const void *quantf, const void *loopf )
{
    x264_t *h = hv;
    int satd = h->param.analyse.i_subpel_refine > 0;
    memcpy( &h->pixf, pixf, sizeof(h->pixf) );
    memcpy( &h->dctf, dctf, sizeof(h->dctf) );
    memcpy( &h->zigzagf, zigzagf, sizeof(h->zigzagf) );
    memcpy( &h->mc, mcf, sizeof(h->mc) );
    memcpy( &h->quantf, quantf, sizeof(h->quantf) );
    memcpy( &h->loopf, loopf, sizeof(h->loopf) );

    memcpy( h->pixf.mbcmp, satd ? h->pixf.satd : h->pixf.sad_aligned, sizeof(h->pixf.mbcmp) );
    memcpy( h->pixf.mbcmp_unaligned, satd ? h->pixf.satd : h->pixf.sad, sizeof(h->pixf.mbcmp_unaligned) );
    h->pixf.intra_mbcmp_x3_16x16 = satd ? h->pixf.intra_satd_x3_16x16 : h->pixf.intra_sad_x3_16x16;
    h->pixf.intra_mbcmp_x3_8x8c  = satd ? h->pixf.intra_satd_x3_8x8c  : h->pixf.intra_sad_x3_8x8c;
    h->pixf.intra_mbcmp_x3_4x4   = satd ? h->pixf.intra_satd_x3_4x4   : h->pixf.intra_sad_x3_4x4;
    h->pixf.intra_mbcmp_x4_4x4_h = satd ? h->pixf.intra_satd_x4_4x4_h : h->pixf.intra_sad_x4_4x4_h;
    h->pixf.intra_mbcmp_x4_4x4_v = satd ? h->pixf.intra_satd_x4_4x4_v : h->pixf.intra_sad_x4_4x4_v;
    h->pixf.intra_mbcmp_x9_4x4   = NULL;
    satd &= h->param.analyse.i_me_method == X264_ME_TESA;
    memcpy( h->pixf.fpelcmp, satd ? h->pixf.satd : h->pixf.sad, sizeof(h->pixf.fpelcmp) );
    memcpy( h->pixf.fpelcmp_x3, satd ? h->pixf.satd_x3 : h->pixf.sad_x3, sizeof(h->pixf.fpelcmp_x3) );
    memcpy( h->pixf.fpelcmp_x4, satd ? h->pixf.satd_x4 : h->pixf.sad_x4, sizeof(h->pixf.fpelcmp_x4) );

    h->mc.prefetch_fenc = h->mc.prefetch_fenc_420;
    h->pixf.intra_mbcmp_x3_chroma = h->pixf.intra_mbcmp_x3_8x8c;
    h->quantf.coeff_last[DCT_CHROMA_DC] = h->quantf.coeff_last4;
    h->quantf.coeff_level_run[DCT_CHROMA_DC] = h->quantf.coeff_level_run4;
}

/* the three intra predictor tables (encoder/encoder.c:551-553; predict_chroma is the copy made at encoder.c:447) */
void xref_install_predict_tables( void *hv, const x264_predict_t *p16, const x264_predict_t *p8c, const x264_predict_t *p4 )
{
    x264_t *h = hv;
    memcpy( h->predict_16x16, p16, sizeof(h->predict_16x16) );
    memcpy( h->predict_8x8c, p8c, sizeof(h->predict_8x8c) );
    memcpy( h->predict_chroma, p8c, sizeof(h->predict_chroma) );
    memcpy( h->predict_4x4, p4, sizeof(h->predict_4x4) );
}

/* encode n_frames tightly packed I420 pictures with x264_encoder_encode, flush, and concatenate every
 * NAL payload into out; returns the byte count, or <0 on error / overflow */
int xref_encode_clip( void *hv, uint8_t *i420, int n_frames, uint8_t *out, int out_cap )
{
    x264_t *h = hv;
    const int w = h->param.i_width, ht = h->param.i_height;
    const size_t pic_bytes = (size_t)w * ht * 3 / 2;
    int total = 0, i, k;
    for( i = 0; ; i++ )
    {
        x264_picture_t pic, pic_out;
        x264_nal_t *nal;
        int n_nal = 0, size;
        if( i < n_frames )
        {
            uint8_t *y = i420 + (size_t)i * pic_bytes;
            x264_picture_init( &pic );
            pic.img.i_csp = X264_CSP_I420;
            pic.img.i_plane = 3;
            pic.img.plane[0] = y;
            pic.img.plane[1] = y + (size_t)w * ht;
            pic.img.plane[2] = y + (size_t)w * ht + (size_t)( w / 2 ) * ( ht / 2 );
            pic.img.i_stride[0] = w;
            pic.img.i_stride[1] = pic.img.i_stride[2] = w / 2;
            pic.i_pts = i;
            size = x264_encoder_encode( h, &nal, &n_nal, &pic, &pic_out );
        }
        else
            size = x264_encoder_encode( h, &nal, &n_nal, NULL, &pic_out );
        if( size < 0 )
            return -1;
        for( k = 0; k < n_nal; k++ )
        {
            if( total + nal[k].i_payload > out_cap )
                return -2;
            memcpy( out + total, nal[k].p_payload, nal[k].i_payload );
            total += nal[k].i_payload;
        }
        if( i >= n_frames && ( size == 0 || i > 2 * n_frames + 8 ) )
            break;
    }
    return total;
}
