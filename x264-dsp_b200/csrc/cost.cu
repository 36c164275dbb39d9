// cost.cu -- batched block costs: SAD / SSD / SATD for all eight block sizes, sm_100a.
//
// Reference: x264_pixel_sad_WxH, x264_pixel_ssd_WxH (common/pixel.c:44-102) and
// x264_pixel_satd_4x4 / _8x4 / PIXEL_SATD (common/pixel.c:267-337).
//
// Mapping: 8 lanes per block pair, 4 block pairs per warp.  A block is cut into "units":
//   SAD/SSD  8x1 pixel rows (4x1 for the 4-wide sizes)
//   SATD     8x4 pixel tiles (4x4 for the 4-wide sizes) -- exactly the reference's base blocks, so
//            the >>1 lands where the reference puts it
// and the units are dealt round-robin to the 8 lanes; three xor-shuffles finish the sum.
// Blocks may start at any byte address (pix2 is a motion-compensated position).
#include "common.cuh"
#include "leaf.cuh"


__global__ void __launch_bounds__( 256 )
xd_cost_batch_kernel( int cmp, int n, const uint8_t *__restrict__ pix1, const int64_t *__restrict__ off1, int stride1,
                      const uint8_t *__restrict__ pix2, const int64_t *__restrict__ off2, int stride2,
                      const uint8_t *__restrict__ size, int32_t *__restrict__ out )
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk = gtid >> 3, sub = gtid & 7;
    const bool live = blk < n;
    int acc = 0;
    if( live )
    {
        const int sz = size[blk] & 7;
        const int w = xd_blk_w[sz], h = xd_blk_h[sz];
        const uint8_t *a = pix1 + off1[blk], *b = pix2 + off2[blk];
        if( cmp != X264DSP_CMP_SATD )
        {
            const int per_row = w >= 8 ? w >> 3 : 1;
            const int units = h * per_row;
            for( int u = sub; u < units; u += 8 )
            {
                const int y = u / per_row, x = ( u % per_row ) * 8;
                const uint8_t *pa = a + (int64_t)y * stride1 + x, *pb = b + (int64_t)y * stride2 + x;
                if( w >= 8 )
                {
                    const uint2 va = xd_load8_unaligned( pa ), vb = xd_load8_unaligned( pb );
                    if( cmp == X264DSP_CMP_SAD )
                        acc += __vsadu4( va.x, vb.x ) + __vsadu4( va.y, vb.y );
                    else
                        acc += xd_sq4( va.x, vb.x ) + xd_sq4( va.y, vb.y );
                }
                else
                {
                    const uint32_t va = xd_load4_unaligned( pa ), vb = xd_load4_unaligned( pb );
                    acc += cmp == X264DSP_CMP_SAD ? __vsadu4( va, vb ) : xd_sq4( va, vb );
                }
            }
        }
        else
        {
            const int per_row = w >= 8 ? w >> 3 : 1;
            const int units = ( h >> 2 ) * per_row;
            for( int u = sub; u < units; u += 8 )
            {
                const int y = ( u / per_row ) * 4, x = ( u % per_row ) * 8;
                uint32_t ra[4], rb[4], sa[4], sb[4];
#pragma unroll
                for( int r = 0; r < 4; r++ )
                {
                    const uint8_t *pa = a + (int64_t)( y + r ) * stride1 + x, *pb = b + (int64_t)( y + r ) * stride2 + x;
                    if( w >= 8 )
                    {
                        const uint2 va = xd_load8_unaligned( pa ), vb = xd_load8_unaligned( pb );
                        ra[r] = va.x; sa[r] = va.y; rb[r] = vb.x; sb[r] = vb.y;
                    }
                    else
                    {
                        ra[r] = xd_load4_unaligned( pa );
                        rb[r] = xd_load4_unaligned( pb );
                        sa[r] = sb[r] = 0;
                    }
                }
                int s = xd_satd4x4( ra, rb );
                if( w >= 8 )
                    s += xd_satd4x4( sa, sb );
                acc += s;
            }
        }
    }
    acc += __shfl_xor_sync( 0xffffffffu, acc, 1 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 2 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 4 );
    if( live && sub == 0 )
        out[blk] = acc;
}

extern "C" int x264dsp_cost_batch_dev( x264dsp_ctx_t *ctx, int cmp, int n,
                                        const uint8_t *pix1, const int64_t *off1, int stride1,
                                        const uint8_t *pix2, const int64_t *off2, int stride2,
                                        const uint8_t *size, int32_t *out, void *stream )
{
    if( !ctx || n < 0 || cmp < X264DSP_CMP_SAD || cmp > X264DSP_CMP_SATD )
        return X264DSP_E_ARG;
    if( n == 0 )
        return 0;
    if( !pix1 || !pix2 || !off1 || !off2 || !size || !out )
        return X264DSP_E_ARG;
    const int64_t threads = (int64_t)n * 8;
    const int grid = (int)( ( threads + 255 ) / 256 );
    xd_cost_batch_kernel<<<grid, 256, 0, xd_stream( ctx, stream )>>>( cmp, n, pix1, off1, stride1, pix2, off2, stride2, size, out );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
