/* xo_me.c -- oracle: motion search and the lowres lookahead cost pass.
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * encoder/me.c:18-38 (iteration / pattern tables), 129-423 (x264_me_search_ref),
 * 426-435 (x264_me_refine_qpel), 466-587 (refine_subpel);
 * encoder/slicetype.c:48-200 (x264_slicetype_mb_cost), 223-322 (x264_slicetype_frame_cost);
 * common/common.h:247-261 (median), 283-293 (predictor round + clip).
 *
 * The search is restated as "evaluate a candidate list, keep the first strict minimum" instead of
 * the reference's packed (cost<<4)+code compares; the two are the same ordering because the codes
 * grow with evaluation order and the incumbent carries code 0.
 */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

#define COST_LIMIT (1 << 28)                      /* me.h:8 */

typedef struct
{
    int size, bw, bh;
    const pixel_t *ref[4];                        /* N, H, V, HV planes at the block origin */
    intptr_t stride;
    pixel_t fenc[16 * XO_FENC_STRIDE];
    const uint16_t *cost_mv;                      /* centre of the table */
    int mvp[2];
    int min_fpel[2], max_fpel[2], min_spel[2], max_spel[2];
    int method, subme, range;
    int fpel_satd;                                /* fpelcmp == satd: TESA with subme >= 2 (encoder.c:429-432) */
    int *thresh;                                  /* p_halfpel_thresh */
    /* result */
    int mv[2], cost, cost_mv_out;
} me_t;

/* { refine_hpel, refine_qpel, me_hpel, me_qpel } per subme (me.c:18-32) */
static const uint8_t subpel_iters[12][4] =
{
    {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2},
    {0,0,2,2}, {0,0,2,2}, {0,0,4,10}, {0,0,4,10}, {0,0,4,10}, {0,0,4,10}
};

static int clampi( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }

static int mv_bits( const me_t *m, int qx, int qy )
{
    return m->cost_mv[qx - m->mvp[0]] + m->cost_mv[qy - m->mvp[1]];
}

/* full-pel SAD + mv cost (COST_MV, me.c:48-52) */
static int fpel_sad( const me_t *m, int mx, int my )
{
    if( m->fpel_satd )
        return xo_satd( m->size, m->fenc, XO_FENC_STRIDE, m->ref[0] + my * m->stride + mx, m->stride );
    return xo_sad( m->size, m->fenc, XO_FENC_STRIDE, m->ref[0] + my * m->stride + mx, m->stride );
}
static int fpel_cost( const me_t *m, int mx, int my )
{
    return fpel_sad( m, mx, my ) + mv_bits( m, mx << 2, my << 2 );
}

/* cost at a quarter-pel position through get_ref (COST_MV_HPEL / COST_MV_SAD / COST_MV_SATD) */
static int qpel_cost( const me_t *m, int qx, int qy, int use_satd )
{
    pixel_t tmp[16 * 16];
    intptr_t ts = 16;
    const pixel_t *p = xo_get_ref( tmp, &ts, m->ref, m->stride, qx, qy, m->bw, m->bh );
    int c = use_satd || m->fpel_satd ? xo_satd( m->size, m->fenc, XO_FENC_STRIDE, p, ts )
                     : xo_sad( m->size, m->fenc, XO_FENC_STRIDE, p, ts );
    return c + mv_bits( m, qx, qy );
}

/* CHECK_MVRANGE (me.c:155-160): a packed compare, x in the high half, y in 15 bits */
static int fpel_in_range( const me_t *m, int mx, int my )
{
    uint32_t lo = ((uint32_t)(-m->min_fpel[0]) << 16) | ((uint32_t)(-m->min_fpel[1]) & 0x7FFF);
    uint32_t hi = ((uint32_t)m->max_fpel[0] << 16) | ((uint32_t)m->max_fpel[1] & 0x7FFF) | 0x8000;
    uint32_t v  = ((uint32_t)mx << 16) | ((uint32_t)my & 0x7FFF);
    return !( ((v + lo) | (hi - v)) & 0x80004000u );
}

static uint32_t pack_mv( int x, int y ) { return ((uint32_t)x & 0xFFFF) | ((uint32_t)y << 16); }

static void refine_subpel( me_t *m, int hpel_iters, int qpel_iters, int final_refine );

static void me_search( me_t *m, const int16_t (*mvc)[2], int n_mvc )
{
    int bmx = clampi( m->mvp[0], m->min_fpel[0] * 4, m->max_fpel[0] * 4 );
    int bmy = clampi( m->mvp[1], m->min_fpel[1] * 4, m->max_fpel[1] * 4 );
    const int pmx = (bmx + 2) >> 2, pmy = (bmy + 2) >> 2;
    int bcost = COST_LIMIT;
    int pred_mx = 0, pred_my = 0, pred_cost = COST_LIMIT;
    uint32_t pmv;
    int i, c;

    if( m->subme >= 3 )
    {
        /* me.c:176-193: sub-pel predictors, then start from the best one rounded to full-pel */
        pmv = pack_mv( bmx, bmy );
        if( n_mvc )
        {
            c = qpel_cost( m, bmx, bmy, 0 );
            if( c < pred_cost ) { pred_cost = c; pred_mx = bmx; pred_my = bmy; }
        }
        for( i = 0; i < n_mvc; i++ )
        {
            uint32_t raw = pack_mv( mvc[i][0], mvc[i][1] );
            if( raw && raw != pmv )
            {
                int qx = clampi( mvc[i][0], m->min_fpel[0] * 4, m->max_fpel[0] * 4 );
                int qy = clampi( mvc[i][1], m->min_fpel[1] * 4, m->max_fpel[1] * 4 );
                c = qpel_cost( m, qx, qy, 0 );
                if( c < pred_cost ) { pred_cost = c; pred_mx = qx; pred_my = qy; }
            }
        }
        bmx = (pred_mx + 2) >> 2;
        bmy = (pred_my + 2) >> 2;
        c = fpel_cost( m, bmx, bmy );
        if( c < bcost ) bcost = c;
    }
    else
    {
        /* me.c:194-229: rounded MVP without its mv cost, then the rounded + clipped candidates */
        bmx = pmx;
        bmy = pmy;
        bcost = fpel_sad( m, bmx, bmy );
        pmv = pack_mv( bmx, bmy );
        if( n_mvc > 0 )
        {
            int best = -1;
            int cand[16][2];
            for( i = 0; i < n_mvc; i++ )
            {
                cand[i][0] = clampi( (mvc[i][0] + 2) >> 2, m->min_fpel[0], m->max_fpel[0] );
                cand[i][1] = clampi( (mvc[i][1] + 2) >> 2, m->min_fpel[1], m->max_fpel[1] );
            }
            for( i = 0; i < n_mvc; i++ )
            {
                uint32_t v = pack_mv( cand[i][0], cand[i][1] );
                if( v && v != pmv )
                {
                    c = fpel_cost( m, cand[i][0], cand[i][1] );
                    if( c < bcost ) { bcost = c; best = i; }
                }
            }
            if( best >= 0 ) { bmx = cand[best][0]; bmy = cand[best][1]; }
        }
    }

    if( pmv )                                     /* me.c:231-233 */
    {
        c = fpel_cost( m, 0, 0 );
        if( c < bcost ) { bcost = c; bmx = 0; bmy = 0; }
    }

    if( m->method == X264DSP_ME_DIA )
    {
        /* me.c:237-274: small diamond, order up, down, left, right */
        static const int8_t dia[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
        int left = m->range;
        do
        {
            int pick = -1;
            for( i = 0; i < 4; i++ )
            {
                c = fpel_cost( m, bmx + dia[i][0], bmy + dia[i][1] );
                if( c < bcost ) { bcost = c; pick = i; }
            }
            if( pick < 0 )
                break;
            bmx += dia[pick][0];
            bmy += dia[pick][1];
        } while( --left && fpel_in_range( m, bmx, bmy ) );
    }
    else if( m->method == X264DSP_ME_HEX )
    {
        /* me.c:276-388: radius-2 hexagon walk with half-hexagon updates, then a 3x3 square */
        static const int8_t hex[8][2] = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
        static const int8_t sq[8][2] = { {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
        int pick = -1, dir;
        for( i = 0; i < 6; i++ )
        {
            c = fpel_cost( m, bmx + hex[i+1][0], bmy + hex[i+1][1] );
            if( c < bcost ) { bcost = c; pick = i; }
        }
        if( pick >= 0 )
        {
            int left;
            dir = pick;
            bmx += hex[dir+1][0];
            bmy += hex[dir+1][1];
            for( left = (m->range >> 1) - 1; left > 0 && fpel_in_range( m, bmx, bmy ); left-- )
            {
                pick = -1;
                for( i = 0; i < 3; i++ )
                {
                    c = fpel_cost( m, bmx + hex[dir+i][0], bmy + hex[dir+i][1] );
                    if( c < bcost ) { bcost = c; pick = i; }
                }
                if( pick < 0 )
                    break;
                dir = (dir + pick - 1 + 6) % 6;
                bmx += hex[dir+1][0];
                bmy += hex[dir+1][1];
            }
        }
        pick = -1;
        for( i = 0; i < 8; i++ )
        {
            c = fpel_cost( m, bmx + sq[i][0], bmy + sq[i][1] );
            if( c < bcost ) { bcost = c; pick = i; }
        }
        if( pick >= 0 ) { bmx += sq[pick][0]; bmy += sq[pick][1]; }
    }

    /* me.c:397-414 */
    if( pred_cost < bcost )
    {
        m->mv[0] = pred_mx; m->mv[1] = pred_my; m->cost = pred_cost;
    }
    else
    {
        m->mv[0] = bmx << 2; m->mv[1] = bmy << 2; m->cost = bcost;
    }
    m->cost_mv_out = mv_bits( m, m->mv[0], m->mv[1] );
    if( bmx == pmx && bmy == pmy && m->subme < 3 )
        m->cost += m->cost_mv_out;

    if( m->subme >= 2 )
        refine_subpel( m, subpel_iters[m->subme][2], subpel_iters[m->subme][3], 0 );
}

/* refine_subpel (me.c:466-587), p_halfpel_thresh == NULL */
static void refine_subpel( me_t *m, int hpel_iters, int qpel_iters, int final_refine )
{
    static const int8_t dq[4][2] = { {0,-1}, {0,1}, {-1,0}, {1,0} };
    int bmx = m->mv[0], bmy = m->mv[1], bcost = m->cost;
    int i, k, c;

    if( hpel_iters && m->subme < 3 )              /* me.c:483-490 */
    {
        int qx = clampi( m->mvp[0], m->min_spel[0] + 2, m->max_spel[0] - 2 );
        int qy = clampi( m->mvp[1], m->min_spel[1] + 2, m->max_spel[1] - 2 );
        if( qx != bmx || qy != bmy )
        {
            c = qpel_cost( m, qx, qy, 0 );
            if( c < bcost ) { bcost = c; bmx = qx; bmy = qy; }
        }
    }

    for( i = hpel_iters; i > 0; i-- )             /* me.c:492-517: half-pel diamond, SAD */
    {
        int ox = bmx, oy = bmy;
        for( k = 0; k < 4; k++ )
        {
            c = qpel_cost( m, ox + 2*dq[k][0], oy + 2*dq[k][1], 0 );
            if( c < bcost ) { bcost = c; bmx = ox + 2*dq[k][0]; bmy = oy + 2*dq[k][1]; }
        }
        if( bmx == ox && bmy == oy )
            break;
    }

    if( !final_refine && !m->fpel_satd )          /* me.c:519-524: re-cost the winner with SATD */
        bcost = qpel_cost( m, bmx, bmy, 1 );

    if( m->thresh )                               /* me.c:526-539 */
    {
        if( (bcost * 7) >> 3 > *m->thresh )
        {
            m->cost = bcost;
            m->mv[0] = bmx;
            m->mv[1] = bmy;
            return;                               /* cost_mv keeps its value */
        }
        else if( bcost < *m->thresh )
            *m->thresh = bcost;
    }

    if( m->subme != 1 )
    {
        int bdir = -1;                            /* me.c:541-564: quarter-pel diamond, SATD */
        for( i = qpel_iters; i > 0; i-- )
        {
            int ox = bmx, oy = bmy, odir = bdir;
            if( bmy <= m->min_spel[1] || bmy >= m->max_spel[1] || bmx <= m->min_spel[0] || bmx >= m->max_spel[0] )
                break;
            for( k = 0; k < 4; k++ )
            {
                if( !final_refine && (k ^ 1) == odir )
                    continue;                     /* the point we just came from */
                c = qpel_cost( m, ox + dq[k][0], oy + dq[k][1], 1 );
                if( c < bcost ) { bcost = c; bmx = ox + dq[k][0]; bmy = oy + dq[k][1]; bdir = k; }
            }
            if( bmx == ox && bmy == oy )
                break;
        }
    }
    else if( bmy > m->min_spel[1] && bmy < m->max_spel[1] && bmx > m->min_spel[0] && bmx < m->max_spel[0] )
    {
        int ox = bmx, oy = bmy;                   /* me.c:565-581: subme 1, one SAD quarter-pel round */
        for( k = 0; k < 4; k++ )
        {
            c = qpel_cost( m, ox + dq[k][0], oy + dq[k][1], 0 );
            if( c < bcost ) { bcost = c; bmx = ox + dq[k][0]; bmy = oy + dq[k][1]; }
        }
    }
    m->cost = bcost;
    m->mv[0] = bmx;
    m->mv[1] = bmy;
    m->cost_mv_out = mv_bits( m, bmx, bmy );
}

void xo_me_search_batch( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                         const x264dsp_me_params_t *prm, int n, const x264dsp_me_block_t *blocks,
                         x264dsp_me_result_t *results )
{
    xo_me_search_batch_ex( g, fenc_slot, fref_slot, prm, n, blocks, results, 0, NULL );
}

/* mode 0: x264_me_search_ref with p_halfpel_thresh = &thresh[i] (NULL: none); mode 1: x264_me_refine_qpel_refdupe
 * (me.c:437-440) and mode 2: x264_me_refine_qpel alone (me.c:426-435, i_ref_cost already subtracted), both on the
 * mv / cost / cost_mv found in results[i] */
void xo_me_search_batch_ex( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                            const x264dsp_me_params_t *prm, int n, const x264dsp_me_block_t *blocks,
                            x264dsp_me_result_t *results, int mode, int32_t *thresh )
{
    uint16_t *table = malloc( 8193 * sizeof(uint16_t) );
    int i, k, y;
    xo_cost_mv_table( prm->qp, table );
    for( i = 0; i < n; i++ )
    {
        const x264dsp_me_block_t *b = &blocks[i];
        me_t m;
        const pixel_t *src = fenc_slot + g->luma_origin + (ptrdiff_t)b->by * g->luma_stride + b->bx;
        memset( &m, 0, sizeof(m) );
        m.size = b->i_pixel;
        m.bw = xo_block_w( m.size );
        m.bh = xo_block_h( m.size );
        m.stride = g->luma_stride;
        for( k = 0; k < 4; k++ )
            m.ref[k] = fref_slot + (size_t)k * g->luma_plane_size + g->luma_origin
                     + (ptrdiff_t)b->by * g->luma_stride + b->bx;
        for( y = 0; y < m.bh; y++ )
            memcpy( m.fenc + y * XO_FENC_STRIDE, src + (ptrdiff_t)y * g->luma_stride, m.bw );
        m.cost_mv = table + 4096;
        m.mvp[0] = b->mvp[0];
        m.mvp[1] = b->mvp[1];
        for( k = 0; k < 2; k++ )
        {
            m.min_fpel[k] = b->mv_min_fpel[k]; m.max_fpel[k] = b->mv_max_fpel[k];
            m.min_spel[k] = b->mv_min_spel[k]; m.max_spel[k] = b->mv_max_spel[k];
        }
        m.method = prm->me_method;
        m.subme = prm->subpel_refine;
        m.range = prm->me_range;
        m.fpel_satd = prm->me_method == X264DSP_ME_TESA && prm->subpel_refine >= 2;
        m.thresh = thresh ? &thresh[i] : NULL;
        if( mode )
        {
            int q = subpel_iters[m.subme][3];
            m.mv[0] = results[i].mv[0];
            m.mv[1] = results[i].mv[1];
            m.cost = results[i].cost;
            m.cost_mv_out = results[i].cost_mv;
            if( mode == 1 )
                refine_subpel( &m, 0, q < 2 ? q : 2, 0 );
            else
            {
                m.thresh = NULL;
                refine_subpel( &m, subpel_iters[m.subme][0], subpel_iters[m.subme][1], 1 );
            }
        }
        else
        {
            me_search( &m, b->mvc, b->i_mvc );
            m.thresh = NULL;
            if( prm->refine_qpel )                /* x264_me_refine_qpel, me.c:426-435 (i_ref_cost = 0) */
                refine_subpel( &m, subpel_iters[m.subme][0], subpel_iters[m.subme][1], 1 );
        }
        results[i].mv[0] = (int16_t)m.mv[0];
        results[i].mv[1] = (int16_t)m.mv[1];
        results[i].cost = m.cost;
        results[i].cost_mv = m.cost_mv_out;
    }
    free( table );
}

/* ------------------------------------------------------------------ lowres lookahead */

static int median3( int a, int b, int c )
{
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : c > hi ? hi : c;
}

void xo_lookahead_frame_cost( const x264dsp_geom_t *g, const uint8_t *slot_b, const uint8_t *slot_p0,
                              int want_intra, int16_t *mvs, int32_t *costs, int32_t *sums,
                              int32_t *row_satds )
{
    const int W = g->mb_w, H = g->mb_h, ls = g->lowres_stride;
    const pixel_t *cur = slot_b + g->slot_lowres_off + g->lowres_origin;
    uint16_t *table = malloc( 8193 * sizeof(uint16_t) );
    int64_t cost_inter = 0, cost_intra = 0;
    int intra_mbs = 0, bx, by, k, r;
    int64_t before[4], after[4];

    xo_cost_mv_table( X264DSP_LOOKAHEAD_QP, table );
    xo_work_counters( before, 0 );
    if( slot_p0 )
    {
        memset( mvs, 0, (size_t)W * H * 2 * sizeof(int16_t) );
        memset( costs, 0, (size_t)W * H * sizeof(int32_t) );
    }
    if( row_satds )
        memset( row_satds, 0, (size_t)2 * H * sizeof(int32_t) );

    /* slicetype.c:285-293: reverse raster over the interior blocks (do_edges = 0) */
    for( by = H - 2; by >= 1; by-- )
        for( bx = W - 2; bx >= 1; bx-- )
        {
            const int xy = by * W + bx;
            const ptrdiff_t pel = ((ptrdiff_t)by * ls + bx) * 8;
            int bcost = COST_LIMIT, icost = COST_LIMIT, b_intra;
            me_t m;

            memset( &m, 0, sizeof(m) );
            for( r = 0; r < 8; r++ )
                memcpy( m.fenc + r * XO_FENC_STRIDE, cur + pel + (ptrdiff_t)r * ls, 8 );

            if( slot_p0 )
            {
                /* slicetype.c:79-101 */
                int16_t mvc[4][2];
                m.size = X264DSP_PIXEL_8x8; m.bw = m.bh = 8;
                m.stride = ls;
                for( k = 0; k < 4; k++ )
                    m.ref[k] = slot_p0 + g->slot_lowres_off + (size_t)k * g->lowres_plane_size
                             + g->lowres_origin + pel;
                m.cost_mv = table + 4096;
                m.min_fpel[0] = -(bx << 3) - 4;  m.max_fpel[0] = ((W - bx - 1) << 3) + 4;
                m.min_fpel[1] = -(by << 3) - 4;  m.max_fpel[1] = ((H - by - 1) << 3) + 4;
                for( k = 0; k < 2; k++ )
                {
                    m.min_spel[k] = (m.min_fpel[k] - 8) << 2;
                    m.max_spel[k] = (m.max_fpel[k] + 8) << 2;
                }
                m.method = X264DSP_ME_DIA;        /* slicetype.c:259-261 */
                m.subme = 2;
                m.range = 16;                     /* common/common.c default i_me_range */

                /* slicetype.c:105-113: right, below, below-left, below-right */
                memcpy( mvc[0], mvs + 2*(xy + 1), 4 );
                memcpy( mvc[1], mvs + 2*(xy + W), 4 );
                memcpy( mvc[2], mvs + 2*(xy + W - 1), 4 );
                memcpy( mvc[3], mvs + 2*(xy + W + 1), 4 );
                m.mvp[0] = median3( mvc[0][0], mvc[1][0], mvc[2][0] );
                m.mvp[1] = median3( mvc[0][1], mvc[1][1], mvc[2][1] );

                m.cost = -1;
                if( !m.mvp[0] && !m.mvp[1] )      /* slicetype.c:117-125 */
                {
                    int c0 = xo_satd( X264DSP_PIXEL_8x8, m.fenc, XO_FENC_STRIDE, m.ref[0], ls );
                    if( c0 < 64 )
                    {
                        m.cost = c0;
                        m.mv[0] = m.mv[1] = 0;
                    }
                }
                if( m.cost < 0 )
                {
                    me_search( &m, (const int16_t (*)[2])mvc, 4 );
                    m.cost -= 1;                  /* slicetype.c:128-130 */
                    if( m.mv[0] || m.mv[1] )
                        m.cost += 5;
                }
                mvs[2*xy] = (int16_t)m.mv[0];
                mvs[2*xy + 1] = (int16_t)m.mv[1];
                costs[xy] = m.cost;
                if( m.cost < bcost )
                    bcost = m.cost;
            }

            if( want_intra )
            {
                /* slicetype.c:145-180: neighbours come from the SOURCE lowres plane */
                pixel_t buf[9 * XO_FDEC_STRIDE];
                pixel_t *blk = buf + XO_FDEC_STRIDE + 8;
                const pixel_t *src = cur + pel;
                int res[3];
                memset( buf, 0, sizeof(buf) );
                memcpy( blk - XO_FDEC_STRIDE - 1, src - ls - 1, 17 );
                for( r = 0; r < 8; r++ )
                    blk[r * XO_FDEC_STRIDE - 1] = src[(ptrdiff_t)r * ls - 1];
                xo_intra_x3_8x8c( 1, m.fenc, blk, res );
                icost = res[0] < res[1] ? res[0] : res[1];
                if( res[2] < icost ) icost = res[2];
                icost += 5 + 4;
                cost_intra += icost;
                if( row_satds )
                    row_satds[H + by] += icost;
            }
            bcost += 4;
            b_intra = icost < bcost;
            if( b_intra )
                bcost = icost;
            intra_mbs += b_intra;
            if( slot_p0 )
            {
                cost_inter += bcost;
                if( row_satds )
                    row_satds[by] += bcost;
            }
        }

    xo_work_counters( after, 0 );
    memset( sums, 0, X264DSP_LA_SUMS * sizeof(int32_t) );
    sums[X264DSP_LA_COST_INTER] = (int32_t)cost_inter;
    sums[X264DSP_LA_COST_INTRA] = (int32_t)cost_intra;
    sums[X264DSP_LA_INTRA_MBS]  = intra_mbs;
    sums[X264DSP_LA_SAD_EVALS]  = (int32_t)( after[2] - before[2] );
    sums[X264DSP_LA_SATD_EVALS] = (int32_t)( after[3] - before[3] );
    free( table );
}
