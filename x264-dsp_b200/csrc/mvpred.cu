// mvpred.cu -- motion-vector prediction of 16x16 partitions and the P_SKIP vector, sm_100a.
//
// Reference: x264_mb_predict_mv_16x16 (common/mvpred.c:101-137), x264_mb_predict_mv_pskip (mvpred.c:139-155),
// x264_median_mv (common/common.h:247-261).  One thread per macroblock; the device routines are what a macroblock
// wavefront calls with the MVs its neighbours have just produced -- here they are batched over
// caller-supplied neighbourhoods so that the rule itself is pinned.
#include "mvpred.cuh"

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

// x264_mb_predict_mv_ref16x16 (mvpred.c:167-219), list 0, reference 0: thread per macroblock, blockIdx.y = frame
__global__ void __launch_bounds__( 256 )
xd_predict_mvc_kernel( int mb_w, int mb_h, const int16_t *__restrict__ lowres_mv, const int16_t *__restrict__ mvr,
                       const int16_t *__restrict__ l0_mv16, int scale, int16_t *__restrict__ mvc, int32_t *__restrict__ n_mvc )
{
    const int n = mb_w * mb_h;
    const int xy = blockIdx.x * blockDim.x + threadIdx.x;
    if( xy >= n )
        return;
    const size_t f = (size_t)blockIdx.y * n;
    const uint32_t *mvr32 = (const uint32_t *)mvr + f;
    uint32_t *out = (uint32_t *)mvc + ( f + xy ) * 9;
    const int x = xy % mb_w, y = xy / mb_w;
    int i = 0;
    if( lowres_mv )
    {
        const uint32_t *lr = (const uint32_t *)lowres_mv + f;
        if( ( __ldg( lr ) & 0xFFFFu ) != 0x7FFFu )                 // the frame pair has been analysed (mc.c:427-429)
        {
            const uint32_t v = __ldg( lr + xy );
            out[i++] = ( ( v & 0xFFFFu ) * 2u & 0xFFFFu ) | ( ( ( v >> 16 ) * 2u ) << 16 );
        }
    }
    // spatial: left, top, top-left, top-right; a neighbour outside the frame is the reference's mvr[-1] = 0
    out[i++] = x > 0 ? __ldg( mvr32 + xy - 1 ) : 0u;
    out[i++] = y > 0 ? __ldg( mvr32 + xy - mb_w ) : 0u;
    out[i++] = ( x > 0 && y > 0 ) ? __ldg( mvr32 + xy - mb_w - 1 ) : 0u;
    out[i++] = ( y > 0 && x < mb_w - 1 ) ? __ldg( mvr32 + xy - mb_w + 1 ) : 0u;
    if( l0_mv16 )
    {
        const uint32_t *t = (const uint32_t *)l0_mv16 + f;
        const int idx[3] = { xy, x < mb_w - 1 ? xy + 1 : -1, y < mb_h - 1 ? xy + mb_w : -1 };
#pragma unroll
        for( int k = 0; k < 3; k++ )
            if( idx[k] >= 0 )
            {
                const uint32_t v = __ldg( t + idx[k] );
                const int tx = ( (int)(int16_t)( v & 0xFFFFu ) * scale + 128 ) >> 8, ty = ( (int)(int16_t)( v >> 16 ) * scale + 128 ) >> 8;
                out[i++] = ( (uint32_t)tx & 0xFFFFu ) | ( (uint32_t)ty << 16 );
            }
    }
    n_mvc[f + xy] = i;
}

extern "C" int x264dsp_predict_mvc_16x16_frames_dev( x264dsp_ctx_t *ctx, int mb_w, int mb_h, int n_frames,
                                                     const int16_t *lowres_mv, const int16_t *mvr, const int16_t *l0_mv16,
                                                     int scale, int16_t *mvc, int32_t *n_mvc, void *stream )
{
    if( !ctx || mb_w <= 0 || mb_h <= 0 || n_frames <= 0 || n_frames > 65535 || !mvr || !mvc || !n_mvc
        || ( ( (uintptr_t)lowres_mv | (uintptr_t)mvr | (uintptr_t)l0_mv16 | (uintptr_t)mvc ) & 3 ) )
        return X264DSP_E_ARG;
    const dim3 grid( ( mb_w * mb_h + 255 ) / 256, n_frames );
    xd_predict_mvc_kernel<<<grid, 256, 0, xd_stream( ctx, stream )>>>( mb_w, mb_h, lowres_mv, mvr, l0_mv16, scale, mvc, n_mvc );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
