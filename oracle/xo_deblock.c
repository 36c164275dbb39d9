/* xo_deblock.c -- oracle: in-loop deblocking filter.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/deblock.c:26-77 (alpha/beta/tc0 tables), 80-194 (bS<4 filters), 196-295 (bS=4 filters),
 * 297-323 (boundary strength), 325-427 (per-MB edge schedule, slice-QP rule).
 */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

/* H.264 table 8-16 / 8-17, indexA/indexB 0..51 */
static const uint8_t alpha_tab[52] =
{
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 4,4,5,6,7,8,9,10,12,13,15,17,20,22,
    25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255
};
static const uint8_t beta_tab[52] =
{
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 2,2,2,3,3,3,3,4,4,4,6,6,7,7,
    8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18
};
static const int8_t tc0_tab[52][3] =      /* bS = 1, 2, 3 */
{
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,1},{0,0,1},{0,0,1},
    {0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2},
    {1,1,2},{1,2,3},{1,2,3},{2,2,3},{2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6},
    {4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},{10,13,20},
    {11,15,23},{13,17,25}
};

static int clampi( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }
static int tab_index( int i ) { return clampi( i, 0, 51 ); }
/* the reference's tables carry 24 leading and 12 trailing guard entries holding the end values */
static int alpha_of( int index_a ) { return index_a < 0 ? 0 : alpha_tab[tab_index( index_a )]; }
static int beta_of( int index_b )  { return index_b < 0 ? 0 : beta_tab[tab_index( index_b )]; }
static int tc0_of( int index_a, int bs )
{
    if( bs == 0 ) return -1;
    return index_a < 0 ? 0 : tc0_tab[tab_index( index_a )][bs - 1];
}

static pixel_t clip_u8( int v ) { return v < 0 ? 0 : v > 255 ? 255 : (pixel_t)v; }

/* one line across an edge, bS < 4 (deblock.c:80-120).  xs = step across the edge. */
static void luma_line( pixel_t *pix, intptr_t xs, int alpha, int beta, int tc0 )
{
    int p2 = pix[-3*xs], p1 = pix[-2*xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2*xs];
    int tc, delta;
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    tc = tc0;
    if( abs( p2 - p0 ) < beta )
    {
        if( tc0 )
            pix[-2*xs] = (pixel_t)( p1 + clampi( ((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0 ) );
        tc++;
    }
    if( abs( q2 - q0 ) < beta )
    {
        if( tc0 )
            pix[xs] = (pixel_t)( q1 + clampi( ((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0 ) );
        tc++;
    }
    delta = clampi( (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc );
    pix[-xs] = clip_u8( p0 + delta );
    pix[0]   = clip_u8( q0 - delta );
}

/* dir_v = 0: vertical edge (filter across x, "deblock_h_luma"); dir_v = 1: horizontal edge.
 * 16 lines in four groups of four, one tc0 per group; tc0 < 0 skips the group (deblock.c:121-145) */
void xo_deblock_luma( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta, const int8_t tc0[4] )
{
    intptr_t xs = dir_v ? stride : 1, ys = dir_v ? 1 : stride;
    int i;
    for( i = 0; i < 16; i++ )
        if( tc0[i >> 2] >= 0 )
            luma_line( pix + i*ys, xs, alpha, beta, tc0[i >> 2] );
}

static void chroma_line( pixel_t *pix, intptr_t xs, int alpha, int beta, int tc )
{
    int p1 = pix[-2*xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs], delta;
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    delta = clampi( (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc );
    pix[-xs] = clip_u8( p0 + delta );
    pix[0]   = clip_u8( q0 - delta );
}

/* NV12 chroma, bS < 4 (deblock.c:147-194): 8 positions along the edge, U and V each;
 * groups of two positions share a tc; tc <= 0 skips the group. */
void xo_deblock_chroma( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta, const int8_t tc0[4] )
{
    /* across the edge: dir_v -> rows (stride); else -> next UV pair (2 bytes) */
    intptr_t xs = dir_v ? stride : 2, ys = dir_v ? 2 : stride;
    int i, c;
    for( i = 0; i < 8; i++ )
        if( tc0[i >> 1] > 0 )
            for( c = 0; c < 2; c++ )
                chroma_line( pix + i*ys + c, xs, alpha, beta, tc0[i >> 1] );
}

/* bS = 4 luma line (deblock.c:196-243) */
static void luma_intra_line( pixel_t *pix, intptr_t xs, int alpha, int beta )
{
    int p2 = pix[-3*xs], p1 = pix[-2*xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2*xs];
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    if( abs( p0 - q0 ) < ((alpha >> 2) + 2) )
    {
        if( abs( p2 - p0 ) < beta )
        {
            int p3 = pix[-4*xs];
            pix[-xs]   = (pixel_t)( (p2 + 2*p1 + 2*p0 + 2*q0 + q1 + 4) >> 3 );
            pix[-2*xs] = (pixel_t)( (p2 + p1 + p0 + q0 + 2) >> 2 );
            pix[-3*xs] = (pixel_t)( (2*p3 + 3*p2 + p1 + p0 + q0 + 4) >> 3 );
        }
        else
            pix[-xs] = (pixel_t)( (2*p1 + p0 + q1 + 2) >> 2 );
        if( abs( q2 - q0 ) < beta )
        {
            int q3 = pix[3*xs];
            pix[0]    = (pixel_t)( (p1 + 2*p0 + 2*q0 + 2*q1 + q2 + 4) >> 3 );
            pix[xs]   = (pixel_t)( (p0 + q0 + q1 + q2 + 2) >> 2 );
            pix[2*xs] = (pixel_t)( (2*q3 + 3*q2 + q1 + q0 + p0 + 4) >> 3 );
        }
        else
            pix[0] = (pixel_t)( (2*q1 + q0 + p1 + 2) >> 2 );
    }
    else
    {
        pix[-xs] = (pixel_t)( (2*p1 + p0 + q1 + 2) >> 2 );
        pix[0]   = (pixel_t)( (2*q1 + q0 + p1 + 2) >> 2 );
    }
}

void xo_deblock_luma_intra( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta )
{
    intptr_t xs = dir_v ? stride : 1, ys = dir_v ? 1 : stride;
    int i;
    for( i = 0; i < 16; i++ )
        luma_intra_line( pix + i*ys, xs, alpha, beta );
}

/* bS = 4 chroma (deblock.c:261-295) */
void xo_deblock_chroma_intra( pixel_t *pix, intptr_t stride, int dir_v, int alpha, int beta )
{
    intptr_t xs = dir_v ? stride : 2, ys = dir_v ? 2 : stride;
    int i, c;
    for( i = 0; i < 8; i++ )
        for( c = 0; c < 2; c++ )
        {
            pixel_t *p = pix + i*ys + c;
            int p1 = p[-2*xs], p0 = p[-xs], q0 = p[0], q1 = p[xs];
            if( abs( p0 - q0 ) < alpha && abs( p1 - p0 ) < beta && abs( q1 - q0 ) < beta )
            {
                p[-xs] = (pixel_t)( (2*p1 + p0 + q1 + 2) >> 2 );
                p[0]   = (pixel_t)( (2*q1 + q0 + p1 + 2) >> 2 );
            }
        }
}

/* deblock_strength_c (deblock.c:297-323) on the scan8 cache layout (common/common.h:136-186):
 * nnz [n][120], ref [n][2][40], mv [n][2][40][2] -> bs [n][2][8][4].  Only bs[dir][0..3] is written. */
void xo_deblock_strength( int n, const uint8_t *nnz, const int8_t *ref, const int16_t *mv, uint8_t *bs )
{
    int m, dir, edge, i;
    for( m = 0; m < n; m++, nnz += 120, ref += 80, mv += 160, bs += 64 )
        for( dir = 0; dir < 2; dir++ )
        {
            int along = dir ? 1 : 8, across = dir ? 8 : 1;
            for( edge = 0; edge < 4; edge++ )
                for( i = 0; i < 4; i++ )
                {
                    int cur = 12 + edge*across + i*along, nb = cur - across, s;
                    if( nnz[cur] || nnz[nb] )
                        s = 2;
                    else if( ref[cur] != ref[nb]
                          || abs( mv[2*cur] - mv[2*nb] ) >= 4 || abs( mv[2*cur+1] - mv[2*nb+1] ) >= 4 )
                        s = 1;
                    else
                        s = 0;
                    bs[dir*32 + edge*4 + i] = (uint8_t)s;
                }
        }
}

/* x264_macroblock_deblock_strength (common/macroblock.c:677-691): intra macroblocks (type 0..3) get bS 3 on the
 * inner edges and keep bs[dir][0]; the others go through deblock_strength_c */
void xo_macroblock_deblock_strength( int n, const int8_t *mb_type, const uint8_t *nnz, const int8_t *ref,
                                     const int16_t *mv, uint8_t *bs )
{
    int m;
    for( m = 0; m < n; m++, nnz += 120, ref += 80, mv += 160, bs += 64 )
    {
        if( mb_type && mb_type[m] >= 0 && mb_type[m] < 4 )
        {
            memset( bs + 4, 3, 12 );
            memset( bs + 32 + 4, 3, 12 );
        }
        else
            xo_deblock_strength( 1, nnz, ref, mv, bs );
    }
}

/* deblock_edge (deblock.c:325-339) */
static void edge_inter( pixel_t *pix, intptr_t stride, const uint8_t bs[4], int index_a, int alpha, int beta,
                        int chroma, int dir_v )
{
    int8_t tc[4];
    int i;
    if( !(bs[0] | bs[1] | bs[2] | bs[3]) || !alpha || !beta )
        return;
    for( i = 0; i < 4; i++ )
        tc[i] = (int8_t)( tc0_of( index_a, bs[i] ) + chroma );
    if( chroma )
        xo_deblock_chroma( pix, stride, dir_v, alpha, beta, tc );
    else
        xo_deblock_luma( pix, stride, dir_v, alpha, beta, tc );
}

/* x264_frame_deblock_row for every row (deblock.c:341-427); macroblocks in raster order, for each:
 * left edge, three inner vertical edges, top edge, three inner horizontal edges; chroma on
 * edges 0 and 2.  Every edge uses the slice QP (no per-MB QP averaging). */
void xo_deblock_frame( const x264dsp_geom_t *g, uint8_t *slot, const int8_t *mb_type,
                       const uint8_t *partition, const int16_t *cbp, const uint8_t *bs_all,
                       int qp, int a_off, int b_off )
{
    const int qpc = xo_chroma_qp( qp );
    const int ia = qp + a_off, ib = qp + b_off, iac = qpc + a_off, ibc = qpc + b_off;
    const int alpha = alpha_of( ia ), beta = beta_of( ib ), alphac = alpha_of( iac ), betac = beta_of( ibc );
    const int ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, e;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
        for( mb_x = 0; mb_x < g->mb_w; mb_x++ )
        {
            int xy = mb_y * g->mb_w + mb_x;
            const uint8_t (*bs)[8][4] = (const uint8_t (*)[8][4])( bs_all + (size_t)xy * 64 );
            pixel_t *py = slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            pixel_t *pc = slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            int intra = mb_type[xy] < 4;                    /* I_4x4, I_8x8, I_16x16, I_PCM */
            int first_only = partition[xy] == 16 && !cbp[xy] && !intra;   /* D_16x16 */
            if( mb_x > 0 )
            {
                if( intra || mb_type[xy - 1] < 4 )
                {
                    xo_deblock_luma_intra( py, ls, 0, alpha, beta );
                    xo_deblock_chroma_intra( pc, cs, 0, alphac, betac );
                }
                else
                {
                    edge_inter( py, ls, bs[0][0], ia, alpha, beta, 0, 0 );
                    edge_inter( pc, cs, bs[0][0], iac, alphac, betac, 1, 0 );
                }
            }
            if( !first_only )
            {
                for( e = 1; e < 4; e++ )
                    edge_inter( py + 4*e, ls, bs[0][e], ia, alpha, beta, 0, 0 );
                edge_inter( pc + 8, cs, bs[0][2], iac, alphac, betac, 1, 0 );
            }
            if( mb_y > 0 )
            {
                if( intra || mb_type[xy - g->mb_w] < 4 )
                {
                    xo_deblock_luma_intra( py, ls, 1, alpha, beta );
                    xo_deblock_chroma_intra( pc, cs, 1, alphac, betac );
                }
                else
                {
                    edge_inter( py, ls, bs[1][0], ia, alpha, beta, 0, 1 );
                    edge_inter( pc, cs, bs[1][0], iac, alphac, betac, 1, 1 );
                }
            }
            if( !first_only )
            {
                for( e = 1; e < 4; e++ )
                    edge_inter( py + (ptrdiff_t)4*e*ls, ls, bs[1][e], ia, alpha, beta, 0, 1 );
                edge_inter( pc + (ptrdiff_t)4*cs, cs, bs[1][2], iac, alphac, betac, 1, 1 );
            }
        }
}
