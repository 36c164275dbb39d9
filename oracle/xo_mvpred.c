/* xo_mvpred.c -- oracle: motion-vector prediction of a 16x16 partition and the P_SKIP vector.
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/mvpred.c:101-137 (x264_mb_predict_mv_16x16), 139-155 (x264_mb_predict_mv_pskip),
 * common/common.h:247-261 (x264_median_mv).  The neighbours are what h->mb.cache holds around X264_SCAN8_0:
 * A = left (-1), B = top (-8), C = top-right (-8+4), D = top-left (-8-1); ref -2 = not available, -1 = intra. */
#include "xo.h"

/* the middle one of three (x264_median, common.h:247-255, is the branch-free form of the same) */
static int median3( int a, int b, int c )
{
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : c > hi ? hi : c;
}

/* x264_mb_predict_mv (mvpred.c:22-99) for the partitions the reference analyses.  shape: 0 = 16x16 or 8x8 (median
 * rules only), 1 / 2 = upper / lower 16x8 (B resp. A wins outright when it uses the same reference), 3 / 4 = left /
 * right 8x16 (A resp. C).  c_unreachable: the partition's top-right block comes later in scan order
 * ((idx&3) >= 2 + (i_width&1) in the reference), so D stands in for C like for a missing macroblock. */
void xo_predict_mv_part( const x264dsp_mv_neighbours_t *nb, int i_ref, int shape, int c_unreachable, int16_t mvp[2] )
{
    x264dsp_mv_neighbours_t v = *nb;
    if( c_unreachable )
        v.ref[2] = -2;
    {
        const int refc = v.ref[2] == -2 ? v.ref[3] : v.ref[2];
        const int16_t *mvc = v.ref[2] == -2 ? v.mv[3] : v.mv[2];
        const int16_t *win = NULL;
        if( shape == 1 && v.ref[1] == i_ref ) win = v.mv[1];
        if( shape == 2 && v.ref[0] == i_ref ) win = v.mv[0];
        if( shape == 3 && v.ref[0] == i_ref ) win = v.mv[0];
        if( shape == 4 && refc == i_ref )     win = mvc;
        if( win )
        {
            mvp[0] = win[0]; mvp[1] = win[1];
            return;
        }
    }
    xo_predict_mv_16x16( &v, i_ref, mvp );
}

void xo_predict_mv_16x16( const x264dsp_mv_neighbours_t *nb, int i_ref, int16_t mvp[2] )
{
    int refa = nb->ref[0], refb = nb->ref[1], refc = nb->ref[2];
    const int16_t *mva = nb->mv[0], *mvb = nb->mv[1], *mvc = nb->mv[2];
    int count;
    if( refc == -2 )                                   /* no top-right macroblock: the top-left one stands in */
    {
        refc = nb->ref[3];
        mvc = nb->mv[3];
    }
    count = ( refa == i_ref ) + ( refb == i_ref ) + ( refc == i_ref );
    if( count == 1 )
    {
        const int16_t *m = refa == i_ref ? mva : refb == i_ref ? mvb : mvc;
        mvp[0] = m[0]; mvp[1] = m[1];
    }
    else if( count == 0 && refb == -2 && refc == -2 && refa != -2 )
    {
        mvp[0] = mva[0]; mvp[1] = mva[1];
    }
    else
    {
        mvp[0] = (int16_t)median3( mva[0], mvb[0], mvc[0] );
        mvp[1] = (int16_t)median3( mva[1], mvb[1], mvc[1] );
    }
}

void xo_predict_mv_pskip( const x264dsp_mv_neighbours_t *nb, int16_t mv[2] )
{
    const int refa = nb->ref[0], refb = nb->ref[1];
    if( refa == -2 || refb == -2
        || ( refa == 0 && nb->mv[0][0] == 0 && nb->mv[0][1] == 0 )
        || ( refb == 0 && nb->mv[1][0] == 0 && nb->mv[1][1] == 0 ) )
        mv[0] = mv[1] = 0;
    else
        xo_predict_mv_16x16( nb, 0, mv );
}

/* x264_mb_predict_mv_ref16x16 (mvpred.c:167-219), list 0, reference 0, for every macroblock of a frame: the candidate
 * list of the 16x16 search.
 *   lowres_mv  [mb][2] the lookahead's MVs of this frame pair, or NULL / first entry 0x7fff when there are none
 *              (doubled, mvpred.c:178-183)
 *   mvr        [mb][2] the 16x16 MVs of the current frame's macroblocks (h->mb.mvr[0][0]); a neighbour outside the
 *              frame contributes (0,0) -- the reference's index -1 (common/macroblock.c:87-89, 304-308)
 *   l0_mv16    [mb][2] the 16x16 MVs of the reference frame for the temporal candidates, scaled by
 *              scale = (curpoc - refpoc) * inv_ref_poc (mvpred.c:193-214); NULL: the reference frame was intra
 * Out: mvc [mb][9][2] (unused entries untouched), n_mvc [mb]. */
void xo_predict_mvc_16x16_frame( int mb_w, int mb_h, const int16_t *lowres_mv, const int16_t *mvr, const int16_t *l0_mv16,
                                 int scale, int16_t *mvc, int32_t *n_mvc )
{
    int x, y, k;
    for( y = 0; y < mb_h; y++ )
        for( x = 0; x < mb_w; x++ )
        {
            const int xy = y * mb_w + x;
            int16_t (*out)[2] = (int16_t (*)[2])( mvc + (size_t)xy * 18 );
            int i = 0;
            const int nb[4] = { x > 0 ? xy - 1 : -1, y > 0 ? xy - mb_w : -1, ( x > 0 && y > 0 ) ? xy - mb_w - 1 : -1,
                                ( y > 0 && x < mb_w - 1 ) ? xy - mb_w + 1 : -1 };
            if( lowres_mv && lowres_mv[0] != 0x7fff )
            {
                out[i][0] = (int16_t)( lowres_mv[2*xy] * 2 );
                out[i][1] = (int16_t)( lowres_mv[2*xy+1] * 2 );
                i++;
            }
            for( k = 0; k < 4; k++, i++ )
            {
                out[i][0] = nb[k] >= 0 ? mvr[2*nb[k]] : 0;
                out[i][1] = nb[k] >= 0 ? mvr[2*nb[k]+1] : 0;
            }
            if( l0_mv16 )
            {
                const int t[3] = { xy, x < mb_w - 1 ? xy + 1 : -1, y < mb_h - 1 ? xy + mb_w : -1 };
                for( k = 0; k < 3; k++ )
                    if( t[k] >= 0 )
                    {
                        out[i][0] = (int16_t)( ( l0_mv16[2*t[k]] * scale + 128 ) >> 8 );
                        out[i][1] = (int16_t)( ( l0_mv16[2*t[k]+1] * scale + 128 ) >> 8 );
                        i++;
                    }
            }
            n_mvc[xy] = i;
        }
}
