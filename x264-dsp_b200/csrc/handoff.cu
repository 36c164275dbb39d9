// handoff.cu -- the entropy coder's hand-off in the bytes it actually needs, sm_100a.
//
// The slice kernels write a macroblock's levels DENSE: 392 int16 per macroblock (16 luma 4x4 blocks, two chroma DC quads, eight
// chroma AC blocks -- encoder/macroblock.c's h->dct.luma4x4 / chroma_dc), 6.4 MB per 1080p frame, almost all of it zeros: the
// writer (x264_macroblock_write_cabac, encoder/cabac.c:571-700) only ever reads a block whose non_zero_count flag is set.  Over
// PCIe that array is two thirds of everything a coded frame sends back.  x264dsp_levels_pack_dev keeps exactly the units the
// writer reads, in its own order, back to back:
//
//   unit  0..15   luma 4x4 block (coding order)      16 levels   present iff nnz[k]
//   unit 16, 17   chroma DC of U, V                   4 levels   present iff nnz[25], nnz[26]
//   unit 18..21   U AC blocks, 22..25 V AC blocks    16 levels   present iff nnz[16..19], nnz[20..23]
//
// mb_offset[frame][mb] = where the macroblock's units start in the frame's stream (int16 units, a multiple of 4),
// frame_total[frame] = the stream's length.  Two kernels: sizes + an exclusive scan per frame (one CTA per frame), then a
// warp per macroblock copies the present units (lane = unit, its place from a ballot).
#include "common.cuh"

#define HO_UNITS 26

// nnz index of unit u
__device__ __forceinline__ int xd_ho_flag_index( int u )
{
    return u < 16 ? u : u < 18 ? 25 + ( u - 16 ) : u - 2;          // 18..25 -> 16..23
}
// where unit u lives in the dense 392-level record, and its length
__device__ __forceinline__ int xd_ho_dense_offset( int u )
{
    return u < 16 ? 16 * u : u < 18 ? 256 + 4 * ( u - 16 ) : 264 + 16 * ( u - 18 );
}

__device__ __forceinline__ uint32_t xd_ho_mask( const uint8_t *nnz )
{
    uint32_t m = 0;
#pragma unroll
    for( int u = 0; u < HO_UNITS; u++ )
        m |= ( nnz[xd_ho_flag_index( u )] != 0 ? 1u : 0u ) << u;
    return m;
}
__device__ __forceinline__ int xd_ho_size( uint32_t m )
{
    return 16 * __popc( m & ~( 3u << 16 ) ) + 4 * __popc( m & ( 3u << 16 ) );
}

__global__ void __launch_bounds__( 256 )
xd_levels_scan_kernel( int mb_count, const uint8_t *__restrict__ nnz, int32_t *__restrict__ mb_offset, int32_t *__restrict__ frame_total )
{
    __shared__ int s_warp[8];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    nnz += (size_t)f * mb_count * X264DSP_RES_NNZ_PER_MB;
    mb_offset += (size_t)f * mb_count;
    int carry = 0;
    for( int m0 = 0; m0 < mb_count; m0 += 256 )
    {
        const int mb = m0 + tid;
        const int size = mb < mb_count ? xd_ho_size( xd_ho_mask( nnz + (size_t)mb * X264DSP_RES_NNZ_PER_MB ) ) : 0;
        int incl = size;
#pragma unroll
        for( int o = 1; o < 32; o <<= 1 )
        {
            const int v = __shfl_up_sync( 0xffffffffu, incl, o );
            if( lane >= o )
                incl += v;
        }
        if( lane == 31 )
            s_warp[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for( int w = 0; w < 8; w++ )
        {
            before += w < warp ? s_warp[w] : 0;
            total += s_warp[w];
        }
        if( mb < mb_count )
            mb_offset[mb] = carry + before + incl - size;
        carry += total;
        __syncthreads();
    }
    if( tid == 0 )
        frame_total[f] = carry;
}

__global__ void __launch_bounds__( 256 )
xd_levels_copy_kernel( int mb_count, const int16_t *__restrict__ levels, const uint8_t *__restrict__ nnz,
                       const int32_t *__restrict__ mb_offset, int16_t *__restrict__ packed, size_t packed_stride )
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 8 + ( threadIdx.x >> 5 );
    if( mb >= mb_count )
        return;
    const size_t rec = (size_t)f * mb_count + mb;
    const bool present = lane < HO_UNITS && nnz[rec * X264DSP_RES_NNZ_PER_MB + xd_ho_flag_index( min( lane, HO_UNITS - 1 ) )] != 0;
    const uint32_t m = __ballot_sync( 0xffffffffu, present );
    if( !present )
        return;
    const int inner = xd_ho_size( m & ( ( 1u << lane ) - 1u ) );
    const uint2 *src = (const uint2 *)( levels + rec * X264DSP_RES_LEVELS_PER_MB + xd_ho_dense_offset( lane ) );
    uint2 *dst = (uint2 *)( packed + (size_t)f * packed_stride + mb_offset[rec] + inner );
    if( lane == 16 || lane == 17 )
        dst[0] = src[0];
    else
    {
        const uint2 a = src[0], b = src[1], c = src[2], d = src[3];
        dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
    }
}

extern "C" int x264dsp_levels_pack_dev( x264dsp_ctx_t *ctx, int n_frames, int mb_count, const int16_t *levels, const uint8_t *nnz,
                                        int16_t *packed, int64_t packed_stride, int32_t *mb_offset, int32_t *frame_total,
                                        void *stream )
{
    if( !ctx || !levels || !nnz || !packed || !mb_offset || !frame_total || n_frames <= 0 || n_frames > 65535 || mb_count <= 0
        || packed_stride < (int64_t)mb_count * X264DSP_RES_LEVELS_PER_MB || ( packed_stride & 3 ) )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    xd_levels_scan_kernel<<<n_frames, 256, 0, s>>>( mb_count, nnz, mb_offset, frame_total );
    const dim3 grid( ( mb_count + 7 ) / 8, n_frames );
    xd_levels_copy_kernel<<<grid, 256, 0, s>>>( mb_count, levels, nnz, mb_offset, packed, (size_t)packed_stride );
    ctx->launches += 2;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
