// mvpred.cu -- motion-vector prediction of 16x16 partitions and the P_SKIP vector, sm_100a.
//
// Reference: x264_mb_predict_mv_16x16 (common/mvpred.c:101-137), x264_mb_predict_mv_pskip (mvpred.c:139-155),
// x264_median_mv (common/common.h:247-261).  One thread per macroblock; the device routines are what a macroblock
// wavefront calls with the MVs its neighbours have just produced -- here they are batched over
// caller-supplied neighbourhoods so that the rule itself is pinned.
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

__global__ void __launch_bounds__( 256 )
xd_predict_mv_kernel( int n, const x264dsp_mv_neighbours_t *__restrict__ nb, const int8_t *__restrict__ i_ref,
                      const uint8_t *__restrict__ shape, uint32_t *__restrict__ mvp, uint32_t *__restrict__ pskip )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n )
        return;
    // 20 bytes per macroblock, 4-byte aligned: five words
    const uint32_t *w = (const uint32_t *)( nb + i );
    uint32_t raw[5];
#pragma unroll
    for( int k = 0; k < 5; k++ )
        raw[k] = __ldg( w + k );
    x264dsp_mv_neighbours_t v;
    memcpy( &v, raw, sizeof( v ) );
    if( mvp )
    {
        const int sh = shape ? shape[i] : 0;
        mvp[i] = xd_predict_mv_part( v, i_ref ? i_ref[i] : 0, sh & 7, ( sh & 8 ) != 0 );
    }
    if( pskip )
        pskip[i] = xd_predict_mv_pskip( v );
}

extern "C" int x264dsp_predict_mv_batch_dev( x264dsp_ctx_t *ctx, int n, const x264dsp_mv_neighbours_t *nb, const int8_t *i_ref,
                                             const uint8_t *shape, int16_t *mvp, int16_t *pskip_mv, void *stream )
{
    if( !ctx || n < 0 || ( !mvp && !pskip_mv ) )
        return X264DSP_E_ARG;
    if( n == 0 )
        return 0;
    if( !nb || ( (uintptr_t)nb & 3 ) || ( (uintptr_t)mvp & 3 ) || ( (uintptr_t)pskip_mv & 3 ) )
        return X264DSP_E_ARG;
    xd_predict_mv_kernel<<<( n + 255 ) / 256, 256, 0, xd_stream( ctx, stream )>>>( n, nb, i_ref, shape, (uint32_t *)mvp, (uint32_t *)pskip_mv );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
