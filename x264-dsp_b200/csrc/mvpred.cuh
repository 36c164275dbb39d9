// mvpred.cuh -- x264_mb_predict_mv_16x16 / x264_mb_predict_mv / x264_mb_predict_mv_pskip (common/mvpred.c:22-155) as device
// routines on a neighbourhood in registers; shared by the batched kernels of mvpred.cu and the P-slice wavefront (pframe.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ int xd_median3( int a, int b, int c )
{
    return max( min( a, b ), min( max( a, b ), c ) );
}

// nb: ref[4] then mv[4][2] -- A left, B top, C top-right, D top-left; returns the packed prediction (x | y << 16)
__device__ __forceinline__ uint32_t xd_predict_mv_16x16( const x264dsp_mv_neighbours_t &nb, int i_ref )
{
    const int refa = nb.ref[0], refb = nb.ref[1];
    int refc = nb.ref[2], cx = nb.mv[2][0], cy = nb.mv[2][1];
    if( refc == -2 )                                       // no top-right macroblock: the top-left one stands in
    {
        refc = nb.ref[3];
        cx = nb.mv[3][0];
        cy = nb.mv[3][1];
    }
    const int ax = nb.mv[0][0], ay = nb.mv[0][1], bx = nb.mv[1][0], by = nb.mv[1][1];
    const int count = ( refa == i_ref ) + ( refb == i_ref ) + ( refc == i_ref );
    int x, y;
    if( count == 1 )
    {
        x = refa == i_ref ? ax : refb == i_ref ? bx : cx;
        y = refa == i_ref ? ay : refb == i_ref ? by : cy;
    }
    else if( count == 0 && refb == -2 && refc == -2 && refa != -2 )
    {
        x = ax;
        y = ay;
    }
    else
    {
        x = xd_median3( ax, bx, cx );
        y = xd_median3( ay, by, cy );
    }
    return ( (uint32_t)x & 0xFFFFu ) | ( (uint32_t)y << 16 );
}

// x264_mb_predict_mv (mvpred.c:22-99): shape 0 = 16x16 / 8x8, 1 / 2 = upper / lower 16x8, 3 / 4 = left / right 8x16;
// c_unreachable: the partition's top-right block comes later in scan order, D stands in for C
__device__ __forceinline__ uint32_t xd_predict_mv_part( x264dsp_mv_neighbours_t nb, int i_ref, int shape, bool c_unreachable )
{
    if( c_unreachable )
        nb.ref[2] = -2;
    const bool use_d = nb.ref[2] == -2;
    const int refc = use_d ? nb.ref[3] : nb.ref[2];
    const int k = shape == 1 ? 1 : ( shape == 2 || shape == 3 ) ? 0 : 2;            // the neighbour that may win outright
    const int refk = k == 2 ? refc : nb.ref[k];
    if( shape != 0 && refk == i_ref )
    {
        const int kk = k == 2 && use_d ? 3 : k;
        return ( (uint32_t)nb.mv[kk][0] & 0xFFFFu ) | ( (uint32_t)nb.mv[kk][1] << 16 );
    }
    return xd_predict_mv_16x16( nb, i_ref );
}

__device__ __forceinline__ uint32_t xd_predict_mv_pskip( const x264dsp_mv_neighbours_t &nb )
{
    const int refa = nb.ref[0], refb = nb.ref[1];
    if( refa == -2 || refb == -2 || ( refa == 0 && !( nb.mv[0][0] | nb.mv[0][1] ) ) || ( refb == 0 && !( nb.mv[1][0] | nb.mv[1][1] ) ) )
        return 0u;
    return xd_predict_mv_16x16( nb, 0 );
}

