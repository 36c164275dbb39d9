// tables.cu -- the reference's six function-pointer tables, served by CUDA shims.  sm_100a.
//
// x264_pixel_init / x264_dct_init / x264_zigzag_init / x264_mc_init / x264_quant_init /
// x264_deblock_init (common/pixel.c:660-747, dct.c:290-355, mc.c:458-504, quant.c:303-353,
// deblock.c:429-456) fill the caller's table with the functions below.  Every function
//   1. gathers its HOST operands into a mapped pinned staging buffer (2-D blocks are packed at a
//      fixed stride, so the device code never sees the caller's strides),
//   2. launches ONE kernel (xd_shim_kernel) that runs the device leaf routine for that member --
//      the same register-level code the frame-batched kernels use (leaf.cuh) -- reading and
//      writing the staging buffer through its device alias,
//   3. synchronises and scatters the results back into the caller's memory, reproducing the side
//      effects callers rely on (intra_*_x3 leaves the last predicted mode in fdec, get_ref may
//      return a pointer into the caller's plane and rewrites *i_dst_stride, ...).
// This is a drop-in for correctness, not speed: the fast path is the frame-batched API.  There is
// no CPU implementation behind these pointers; if no CUDA device opens, the init functions abort.
// memcpy_aligned / memzero_aligned are libc in the reference too (mc.c:487-488) and stay libc.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "leaf.cuh"
#include "predict_warp.cuh"
#define X264DSP_OWN_TABLE_TYPES 1
#include "../../include/x264dsp_tables.h"

#define FENC_STRIDE 16          // common/common.h:871
#define FDEC_STRIDE 32          // common/common.h:872

enum
{
    OP_CMP = 1, OP_VAR, OP_VAR2, OP_INTRA, OP_PREDICT,
    OP_SUB_DCT, OP_SUB_DCT_DC, OP_ADD_IDCT, OP_ADD_IDCT_DC, OP_DCT4X4DC, OP_IDCT4X4DC, OP_ZIGZAG,
    OP_QUANT, OP_QUANT_DC, OP_DEQUANT, OP_DEQUANT_DC, OP_OPT_CHROMA_DC, OP_DENOISE, OP_DECIMATE, OP_COEFF_LAST,
    OP_LEVEL_RUN,
    OP_AVG, OP_COPY, OP_MC_CHROMA, OP_INTERLEAVE, OP_DEINTERLEAVE, OP_HPEL, OP_LOWRES,
    OP_DB_LUMA, OP_DB_CHROMA, OP_DB_STRENGTH
};
enum { CMP_SAD = 0, CMP_SSD = 1, CMP_SATD = 2 };

struct xd_shim_args
{
    int op;
    int a[15];
};

// ---------------------------------------------------------------------------------------------
// device side

__device__ __forceinline__ uint32_t xs_pack4( const uint8_t *p )
{
    return (uint32_t)p[0] | ( (uint32_t)p[1] << 8 ) | ( (uint32_t)p[2] << 16 ) | ( (uint32_t)p[3] << 24 );
}
__device__ __forceinline__ void xs_unpack4( uint8_t *p, uint32_t v )
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)( v >> 8 ); p[2] = (uint8_t)( v >> 16 ); p[3] = (uint8_t)( v >> 24 );
}

// warp-cooperative block cost (common/pixel.c:44-102, 267-337); every lane gets the total.
// SATD tiles are the reference's base blocks (8x4, or 4x4 for 4-wide), each halved once.
__device__ int xs_block_cost( int cmp, int w, int h, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int lane )
{
    int acc = 0;
    if( cmp == CMP_SATD )
    {
        const int tw = w >= 8 ? 8 : 4, tiles_x = w / tw, tiles = tiles_x * ( h >> 2 );
        for( int t = lane; t < tiles; t += 32 )
        {
            const int x = ( t % tiles_x ) * tw, y = ( t / tiles_x ) * 4;
            uint32_t a[4], b[4];
            int s = 0;
            for( int half = 0; half < tw; half += 4 )
            {
#pragma unroll
                for( int r = 0; r < 4; r++ )
                {
                    a[r] = xs_pack4( p1 + ( y + r ) * s1 + x + half );
                    b[r] = xs_pack4( p2 + ( y + r ) * s2 + x + half );
                }
                s += xd_satd4x4( a, b );
            }
            acc += s;
        }
    }
    else
    {
        const int wx = w >> 2, units = wx * h;
        for( int u = lane; u < units; u += 32 )
        {
            const int x = ( u % wx ) * 4, y = u / wx;
            const uint32_t a = xs_pack4( p1 + y * s1 + x ), b = xs_pack4( p2 + y * s2 + x );
            acc += cmp == CMP_SAD ? (int)__vsadu4( a, b ) : (int)xd_sq4( a, b );
        }
    }
    return (int)__reduce_add_sync( 0xffffffffu, (unsigned)acc );
}

// position of 4x4 block `i` (coding order) inside a 8x8 / 16x16 block (dct.c:152-166, 237-251)
__device__ __forceinline__ int xs_blk_x( int i ) { return 4 * ( ( i & 1 ) + 2 * ( ( i >> 2 ) & 1 ) ); }
__device__ __forceinline__ int xs_blk_y( int i ) { return 4 * ( ( ( i >> 1 ) & 1 ) + 2 * ( i >> 3 ) ); }

__device__ void xs_had4( int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3 )
{
    const int s01 = a + b, d01 = a - b, s23 = c + d, d23 = c - d;
    o0 = s01 + s23; o1 = s01 - s23; o2 = d01 - d23; o3 = d01 + d23;
}

__global__ void __launch_bounds__( 128 )
xd_shim_kernel( xd_shim_args A, uint8_t *__restrict__ b )
{
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int *a = A.a;
    switch( A.op )
    {
    case OP_CMP:        // a: cmp, w, h, nref, off_fenc, off_ref0, ref_pitch, off_out
    {
        if( tid >= 32 )
            return;
        for( int k = 0; k < a[3]; k++ )
        {
            const int c = xs_block_cost( a[0], a[1], a[2], b + a[4], 16, b + a[5] + k * a[6], 16, lane );
            if( lane == 0 )
                ( (int *)( b + a[7] ) )[k] = c;
        }
        break;
    }
    case OP_VAR:        // a: w, h, off_pix (stride 16), off_out (u64)      pixel.c:185-205
    {
        if( tid >= 32 )
            return;
        uint32_t sum = 0, sqr = 0;
        for( int i = lane; i < a[0] * a[1]; i += 32 )
        {
            const uint32_t v = b[a[2] + ( i / a[0] ) * 16 + i % a[0]];
            sum += v;
            sqr += v * v;
        }
        sum = __reduce_add_sync( 0xffffffffu, sum );
        sqr = __reduce_add_sync( 0xffffffffu, sqr );
        if( lane == 0 )
            *(unsigned long long *)( b + a[3] ) = sum + ( (unsigned long long)sqr << 32 );
        break;
    }
    case OP_VAR2:       // a: off_pix1, off_pix2 (8x8 at stride 16), off_out {var, ssd}      pixel.c:211-231
    {
        if( tid >= 32 )
            return;
        int sum = 0;
        uint32_t sqr = 0;
        for( int i = lane; i < 64; i += 32 )
        {
            const int d = (int)b[a[0] + ( i >> 3 ) * 16 + ( i & 7 )] - (int)b[a[1] + ( i >> 3 ) * 16 + ( i & 7 )];
            sum += d;
            sqr += d * d;
        }
        sum = (int)__reduce_add_sync( 0xffffffffu, (unsigned)sum );
        sqr = __reduce_add_sync( 0xffffffffu, sqr );
        if( lane == 0 )
        {
            const uint32_t s = (uint32_t)abs( sum );
            ( (int *)( b + a[2] ) )[0] = (int)( sqr - (uint32_t)( ( (unsigned long long)s * s ) >> 6 ) );
            ( (int *)( b + a[2] ) )[1] = (int)sqr;
        }
        break;
    }
    case OP_INTRA:      // a: cmp, size, n_modes, off_fenc, off_fdec (block origin), off_out, modes[4], res_index[4]
    {                   // pixel.c:489-555
        if( tid >= 32 )
            return;
        uint8_t *fd = b + a[4];
        for( int m = 0; m < a[2]; m++ )
        {
            __syncwarp();
            xs_predict( fd, a[1], a[6 + m], lane );
            __syncwarp();
            const int c = xs_block_cost( a[0], a[1], a[1], fd, FDEC_STRIDE, b + a[3], FENC_STRIDE, lane );
            if( lane == 0 )
                ( (int *)( b + a[5] ) )[a[10 + m]] = c;
        }
        break;
    }
    case OP_PREDICT:    // a: size, mode, off_fdec (block origin)                  predict.c:42-546
    {
        if( tid >= 32 )
            return;
        xs_predict( b + a[2], a[0], a[1], lane );
        break;
    }
    case OP_SUB_DCT:    // a: n_blocks (1,4,16), off_fenc, off_fdec, off_dct      dct.c:115-166
    {
        if( tid >= a[0] )
            return;
        const int x = xs_blk_x( tid ), y = xs_blk_y( tid );
        uint32_t f[4], p[4];
        int dct[16];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            f[r] = xs_pack4( b + a[1] + ( y + r ) * FENC_STRIDE + x );
            p[r] = xs_pack4( b + a[2] + ( y + r ) * FDEC_STRIDE + x );
        }
        xd_sub4x4_dct( dct, f, p );
        int16_t *o = (int16_t *)( b + a[3] ) + tid * 16;
#pragma unroll
        for( int i = 0; i < 16; i++ )
            o[i] = (int16_t)dct[i];
        break;
    }
    case OP_SUB_DCT_DC: // a: off_fenc, off_fdec, off_dct[4]      dct.c:168-195
    {
        if( tid >= 32 )
            return;
        int dc = 0;
        if( lane < 4 )
        {
            const int x = ( lane & 1 ) * 4, y = ( lane >> 1 ) * 4;
            for( int r = 0; r < 4; r++ )
                for( int c = 0; c < 4; c++ )
                    dc += (int)b[a[0] + ( y + r ) * FENC_STRIDE + x + c] - (int)b[a[1] + ( y + r ) * FDEC_STRIDE + x + c];
        }
        // the per-quadrant sums pass through dctcoef (int16) storage before the 2x2 transform
        const int q0 = (int16_t)__shfl_sync( 0xffffffffu, dc, 0 ), q1 = (int16_t)__shfl_sync( 0xffffffffu, dc, 1 );
        const int q2 = (int16_t)__shfl_sync( 0xffffffffu, dc, 2 ), q3 = (int16_t)__shfl_sync( 0xffffffffu, dc, 3 );
        if( lane == 0 )
        {
            int16_t *o = (int16_t *)( b + a[2] );
            const int d0 = q0 + q1, d1 = q2 + q3, d2 = q0 - q1, d3 = q2 - q3;
            o[0] = (int16_t)( d0 + d1 ); o[1] = (int16_t)( d0 - d1 ); o[2] = (int16_t)( d2 + d3 ); o[3] = (int16_t)( d2 - d3 );
        }
        break;
    }
    case OP_ADD_IDCT:   // a: n_blocks, off_fdec, off_dct      dct.c:197-251
    {
        if( tid >= a[0] )
            return;
        const int x = xs_blk_x( tid ), y = xs_blk_y( tid );
        const int16_t *in = (const int16_t *)( b + a[2] ) + tid * 16;
        int dct[16];
        uint32_t p[4];
#pragma unroll
        for( int i = 0; i < 16; i++ )
            dct[i] = in[i];
#pragma unroll
        for( int r = 0; r < 4; r++ )
            p[r] = xs_pack4( b + a[1] + ( y + r ) * FDEC_STRIDE + x );
        xd_add4x4_idct( p, dct );
#pragma unroll
        for( int r = 0; r < 4; r++ )
            xs_unpack4( b + a[1] + ( y + r ) * FDEC_STRIDE + x, p[r] );
        break;
    }
    case OP_ADD_IDCT_DC: // a: n_blocks (4: coding order 2x2; 16: RASTER 4x4), off_fdec, off_dct      dct.c:253-284
    {
        if( tid >= a[0] )
            return;
        const int x = a[0] == 4 ? ( tid & 1 ) * 4 : ( tid & 3 ) * 4, y = a[0] == 4 ? ( tid >> 1 ) * 4 : ( tid >> 2 ) * 4;
        uint32_t p[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
            p[r] = xs_pack4( b + a[1] + ( y + r ) * FDEC_STRIDE + x );
        xd_add4x4_dc( p, ( (const int16_t *)( b + a[2] ) )[tid] );
#pragma unroll
        for( int r = 0; r < 4; r++ )
            xs_unpack4( b + a[1] + ( y + r ) * FDEC_STRIDE + x, p[r] );
        break;
    }
    case OP_DCT4X4DC:   // a: off_d, inverse      dct.c:36-100 (intermediate and result stored as int16)
    case OP_IDCT4X4DC:
    {
        if( tid >= 32 )
            return;
        int16_t *d = (int16_t *)( b + a[0] );
        int t[4] = { 0, 0, 0, 0 };
        if( lane < 4 )      // first pass: lane = input row i, produces tmp[0..3][i]
            xs_had4( d[lane * 4], d[lane * 4 + 1], d[lane * 4 + 2], d[lane * 4 + 3], t[0], t[1], t[2], t[3] );
        int col[4];         // second pass: lane = row i of tmp = {tmp[i][0..3]} = t[i] of lanes 0..3
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            int v = 0;
#pragma unroll
            for( int j = 0; j < 4; j++ )
            {
                const int tj = (int16_t)__shfl_sync( 0xffffffffu, t[j], k );
                if( j == lane )
                    v = tj;
            }
            col[k] = v;
        }
        if( lane < 4 )
        {
            int o[4];
            xs_had4( col[0], col[1], col[2], col[3], o[0], o[1], o[2], o[3] );
            for( int k = 0; k < 4; k++ )
                d[lane * 4 + k] = (int16_t)( A.op == OP_DCT4X4DC ? ( o[k] + 1 ) >> 1 : o[k] );
        }
        break;
    }
    case OP_ZIGZAG:     // a: off_dct, off_level      dct.c:329-347
    {
        if( tid >= 32 )
            return;
        const int16_t *in = (const int16_t *)( b + a[0] );
        int16_t *out = (int16_t *)( b + a[1] );
        if( lane == 0 )
        {
            int q[16], lv[16];
            for( int i = 0; i < 16; i++ )
                q[i] = in[i];
            xd_zigzag( lv, q );
            for( int i = 0; i < 16; i++ )
                out[i] = (int16_t)lv[i];
        }
        break;
    }
    case OP_QUANT:      // a: n, off_dct, off_mf (u16[n]), off_bias (u16[n]), off_nz      quant.c:29-45
    case OP_QUANT_DC:   // a: n, off_dct, mf, bias, off_nz                               quant.c:47-62
    {
        if( tid >= 32 )
            return;
        int16_t *d = (int16_t *)( b + a[1] );
        int v = 0;
        if( lane < a[0] )
        {
            const int mf = A.op == OP_QUANT ? ( (const uint16_t *)( b + a[2] ) )[lane] : a[2];
            const int bias = A.op == OP_QUANT ? ( (const uint16_t *)( b + a[3] ) )[lane] : a[3];
            v = xd_quant1( d[lane], mf, bias );
            d[lane] = (int16_t)v;
        }
        const unsigned nz = __ballot_sync( 0xffffffffu, v != 0 );
        if( lane == 0 )
            *(int *)( b + a[4] ) = nz != 0;
        break;
    }
    case OP_DEQUANT:    // a: off_dct, off_dmf (int[16], row qp%6), qp      quant.c:64-81
    case OP_DEQUANT_DC: //                                                  quant.c:83-101
    {
        if( tid >= 16 )
            return;
        int16_t *d = (int16_t *)( b + a[0] );
        const int *dmf = (const int *)( b + a[1] );
        const int qbits = a[2] / 6 - ( A.op == OP_DEQUANT ? 4 : 6 );
        const int m = A.op == OP_DEQUANT ? dmf[tid] : dmf[0];
        int v = d[tid];
        if( qbits >= 0 )
            v = A.op == OP_DEQUANT ? ( v * m ) << qbits : v * ( m << qbits );
        else
            v = ( v * m + ( 1 << ( -qbits - 1 ) ) ) >> ( -qbits );
        d[tid] = (int16_t)v;
        break;
    }
    case OP_OPT_CHROMA_DC: // a: off_dct[4], dmf, off_ret      quant.c:133-192
    {
        if( tid != 0 )
            return;
        int16_t *d = (int16_t *)( b + a[0] );
        int dc[4] = { d[0], d[1], d[2], d[3] };
        const int nz = xd_optimize_chroma_dc( dc, a[1] );
        for( int i = 0; i < 4; i++ )
            d[i] = (int16_t)dc[i];
        *(int *)( b + a[2] ) = nz;
        break;
    }
    case OP_DENOISE:    // a: off_dct, off_sum (u32), off_offset (u16), size      quant.c:195-207
    {
        int16_t *d = (int16_t *)( b + a[0] );
        uint32_t *sum = (uint32_t *)( b + a[1] );
        const uint16_t *off = (const uint16_t *)( b + a[2] );
        for( int i = tid; i < a[3]; i += nthreads )
        {
            int level = d[i];
            const int sign = level >> 31;
            level = ( level + sign ) ^ sign;
            sum[i] += level;
            level -= off[i];
            d[i] = (int16_t)( level < 0 ? 0 : ( level ^ sign ) - sign );
        }
        break;
    }
    case OP_DECIMATE:   // a: off_dct, first (0 or 1), off_ret      quant.c:221-261
    {
        if( tid != 0 )
            return;
        const int16_t *d = (const int16_t *)( b + a[0] );
        int lv[16];
        for( int i = 0; i < 16; i++ )
            lv[i] = d[i];
        *(int *)( b + a[2] ) = xd_decimate( lv, a[1] );
        break;
    }
    case OP_COEFF_LAST: // a: off_dct, n, off_ret      quant.c:263-276
    {
        if( tid >= 32 )
            return;
        const int16_t *d = (const int16_t *)( b + a[0] );
        int last = -1;
        for( int i = lane; i < a[1]; i += 32 )
            if( d[i] )
                last = i;
        last = (int)__reduce_max_sync( 0xffffffffu, last );
        if( lane == 0 )
            *(int *)( b + a[2] ) = last;
        break;
    }
    case OP_LEVEL_RUN:  // a: off_dct, n, off_runlevel {int last, int mask, int16 level[16]}, off_ret      quant.c:278-301
    {
        if( tid >= 32 )
            return;
        const int16_t *d = (const int16_t *)( b + a[0] );
        const int v = lane < a[1] ? d[lane] : 0;
        const unsigned nzmask = __ballot_sync( 0xffffffffu, v != 0 );
        int *rl = (int *)( b + a[2] );
        int16_t *level = (int16_t *)( rl + 2 );
        // position of this coefficient among the non-zero ones, counted from the top
        const int rank = __popc( nzmask >> lane ) - 1;
        if( v != 0 )
            level[rank] = (int16_t)v;
        if( lane == 0 )
        {
            const int last = nzmask ? 31 - __clz( (int)nzmask ) : -1;
            rl[0] = last;
            if( !nzmask )
            {
                // the reference's do-while reads dct[-1] into level[0] and sets mask = 1 << -1 (undefined);
                // its callers never pass an empty block.  Report "no coefficients" with one zero level.
                level[0] = 0;
                rl[1] = 0;
            }
            else
                rl[1] = (int)nzmask;
            *(int *)( b + a[3] ) = nzmask ? __popc( nzmask ) : 1;
        }
        break;
    }
    case OP_AVG:        // a: off_src1, off_src2, off_dst, n_bytes      mc.c:74-87
        for( int i = tid; i < a[3]; i += nthreads )
            b[a[2] + i] = (uint8_t)( ( b[a[0] + i] + b[a[1] + i] + 1 ) >> 1 );
        break;
    case OP_COPY:       // a: off_src, off_dst, n_bytes      mc.c:92-103, 331-341
        for( int i = tid; i < a[2]; i += nthreads )
            b[a[1] + i] = b[a[0] + i];
        break;
    case OP_MC_CHROMA:  // a: off_src (pitch a[1], (h+1) rows of 2w+2 bytes), pitch, off_u, off_v, w, h, dx, dy      mc.c:290-323
    {
        const int w = a[4], h = a[5], dx = a[6], dy = a[7];
        const int cA = ( 8 - dx ) * ( 8 - dy ), cB = dx * ( 8 - dy ), cC = ( 8 - dx ) * dy, cD = dx * dy;
        for( int i = tid; i < w * h; i += nthreads )
        {
            const int x = i % w, y = i / w;
            const uint8_t *s = b + a[0] + y * a[1] + 2 * x, *sp = s + a[1];
            b[a[2] + i] = (uint8_t)( ( cA * s[0] + cB * s[2] + cC * sp[0] + cD * sp[2] + 32 ) >> 6 );
            b[a[3] + i] = (uint8_t)( ( cA * s[1] + cB * s[3] + cC * sp[1] + cD * sp[3] + 32 ) >> 6 );
        }
        break;
    }
    case OP_INTERLEAVE: // a: off_u, off_v, off_dst, n      mc.c:343-353, 367-376
        for( int i = tid; i < a[3]; i += nthreads )
        {
            b[a[2] + 2 * i] = b[a[0] + i];
            b[a[2] + 2 * i + 1] = b[a[1] + i];
        }
        break;
    case OP_DEINTERLEAVE: // a: off_src, off_u, off_v, n      mc.c:355-365
        for( int i = tid; i < a[3]; i += nthreads )
        {
            b[a[1] + i] = b[a[0] + 2 * i];
            b[a[2] + i] = b[a[0] + 2 * i + 1];
        }
        break;
    case OP_HPEL:       // a: off_src (origin of sample (0,0)), pitch, width, height, off_h, off_v (col -2 first), off_c
    {                   // mc.c:144-167.  dsth/dstc: width x height; dstv: (width+5) x height
        const int pitch = a[1], W = a[2], H = a[3], vw = W + 5;
        for( int i = tid; i < vw * H; i += nthreads )
        {
            const int x = i % vw - 2, y = i / vw;
            const uint8_t *s = b + a[0] + y * pitch + x;
#define XS_TAPV( q ) ( (q)[-2 * pitch] + (q)[3 * pitch] - 5 * ( (q)[-pitch] + (q)[2 * pitch] ) + 20 * ( (q)[0] + (q)[pitch] ) )
            b[a[5] + i] = (uint8_t)xd_clip_u8( ( XS_TAPV( s ) + 16 ) >> 5 );
            if( x >= 0 && x < W )
            {
                b[a[4] + y * W + x] = (uint8_t)xd_clip_u8( ( s[-2] + s[3] - 5 * ( s[-1] + s[2] ) + 20 * ( s[0] + s[1] ) + 16 ) >> 5 );
                // centre: six-tap over the unrounded vertical results, held as int16 in the reference's buf[]
                const int v0 = (int16_t)XS_TAPV( s - 2 ), v1 = (int16_t)XS_TAPV( s - 1 ), v2 = (int16_t)XS_TAPV( s );
                const int v3 = (int16_t)XS_TAPV( s + 1 ), v4 = (int16_t)XS_TAPV( s + 2 ), v5 = (int16_t)XS_TAPV( s + 3 );
                b[a[6] + y * W + x] = (uint8_t)xd_clip_u8( ( v0 + v5 - 5 * ( v1 + v4 ) + 20 * ( v2 + v3 ) + 512 ) >> 10 );
            }
#undef XS_TAPV
        }
        break;
    }
    case OP_LOWRES:     // a: off_src, src_pitch, off_dst0, off_dsth, off_dstv, off_dstc, width, height      mc.c:404-428
    {
        const int sp = a[1], W = a[6], H = a[7];
        for( int i = tid; i < W * H; i += nthreads )
        {
            const int x = i % W, y = i / W;
            const uint8_t *s0 = b + a[0] + 2 * y * sp + 2 * x, *s1 = s0 + sp, *s2 = s1 + sp;
#define XS_LR( p, q, r, s ) (uint8_t)( ( ( ( (p) + (q) + 1 ) >> 1 ) + ( ( (r) + (s) + 1 ) >> 1 ) + 1 ) >> 1 )
            b[a[2] + i] = XS_LR( s0[0], s1[0], s0[1], s1[1] );
            b[a[3] + i] = XS_LR( s0[1], s1[1], s0[2], s1[2] );
            b[a[4] + i] = XS_LR( s1[0], s2[0], s1[1], s2[1] );
            b[a[5] + i] = XS_LR( s1[1], s2[1], s1[2], s2[2] );
#undef XS_LR
        }
        break;
    }
    case OP_DB_LUMA:    // a: off_region (pitch 16, first sample = p3 of line 0), dir, intra, alpha, beta, tc0[4] in a[5..8]
    {                   // deblock.c:80-145, 196-259.  dir 0: lines are rows (filter across x); dir 1: lines are columns
        if( tid >= 16 )
            return;
        const int xs = a[1] ? 16 : 1, ys = a[1] ? 1 : 16;
        uint8_t *p = b + a[0] + tid * ys;
        int s[8];
#pragma unroll
        for( int k = 0; k < 8; k++ )
            s[k] = p[k * xs];
        if( a[2] )
            xd_luma_intra_line( s, a[3], a[4] );
        else
        {
            const int tc0 = a[5 + ( tid >> 2 )];
            if( tc0 < 0 )
                return;
            xd_luma_line( s, a[3], a[4], tc0 );
        }
#pragma unroll
        for( int k = 1; k < 7; k++ )
            p[k * xs] = (uint8_t)s[k];
        break;
    }
    case OP_DB_CHROMA:  // a: off_region (pitch 16), dir, intra, alpha, beta, tc0[4] in a[5..8]      deblock.c:147-194, 261-295
    {                   // dir 0: 8 rows x 8 bytes (u v u v | u v u v), line = (row, plane); dir 1: 4 rows x 16 bytes, line = byte column
        if( tid >= 16 )
            return;
        uint8_t *p;
        int xs, tci;
        if( a[1] )
        {
            p = b + a[0] + tid; xs = 16; tci = tid >> 2;
        }
        else
        {
            p = b + a[0] + ( tid >> 1 ) * 16 + ( tid & 1 ); xs = 2; tci = tid >> 2;
        }
        int s[4];
#pragma unroll
        for( int k = 0; k < 4; k++ )
            s[k] = p[k * xs];
        const int tc = a[5 + tci];
        if( !a[2] && tc <= 0 )
            return;
        xd_chroma_line( s, a[3], a[4], tc, a[2] != 0 );
        p[xs] = (uint8_t)s[1];
        p[2 * xs] = (uint8_t)s[2];
        break;
    }
    default:
        break;
    }
}

// deblock_strength_c: served by the batched kernel (deblock.cu)
extern "C" int x264dsp_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const uint8_t *nnz, const int8_t *ref,
                                              const int16_t *mv, uint8_t *bs, void *stream );

// ---------------------------------------------------------------------------------------------
// host side: the process-wide shim context and the staging helpers

static x264dsp_ctx *g_shim_ctx;
static uint8_t *g_h, *g_d;          // host / device alias of the mapped staging buffer
static size_t g_cap;

static void xs_die( const char *what, int rc )
{
    fprintf( stderr, "x264dsp_b200: %s failed (%d): the table shims have no CPU fallback\n", what, rc );
    abort();
}

static void xs_reserve( size_t bytes )
{
    if( bytes <= g_cap )
        return;
    cudaStreamSynchronize( g_shim_ctx->stream );
    if( g_h )
        cudaFreeHost( g_h );
    size_t cap = 1 << 16;
    while( cap < bytes )
        cap <<= 1;
    cudaError_t e = cudaHostAlloc( (void **)&g_h, cap, cudaHostAllocMapped );
    if( e == cudaSuccess )
        e = cudaHostGetDevicePointer( (void **)&g_d, g_h, 0 );
    if( e != cudaSuccess )
        xs_die( "cudaHostAlloc(mapped)", (int)e );
    g_cap = cap;
    g_shim_ctx->shim_host = NULL;   // owned here, not by the context
}

extern "C" struct x264dsp_ctx *x264dsp_tables_context( void )
{
    if( !g_shim_ctx )
    {
        const char *dev = getenv( "X264DSP_DEVICE" );
        x264dsp_ctx *ctx = NULL;
        if( x264dsp_create( dev ? atoi( dev ) : 0, &ctx ) != 0 )
            return NULL;
        g_shim_ctx = ctx;
        xs_reserve( 1 << 16 );
    }
    return g_shim_ctx;
}

static void xs_open( void )
{
    if( !x264dsp_tables_context() )
        xs_die( "opening a CUDA device", X264DSP_E_NOGPU );
}

static void xs_run( const xd_shim_args &A, int work_items )
{
    cudaSetDevice( g_shim_ctx->device );
    int grid = ( work_items + 127 ) / 128;
    if( grid < 1 ) grid = 1;
    if( grid > 1024 ) grid = 1024;
    xd_shim_kernel<<<grid, 128, 0, g_shim_ctx->stream>>>( A, g_d );
    g_shim_ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if( e == cudaSuccess )
        e = cudaStreamSynchronize( g_shim_ctx->stream );
    if( e != cudaSuccess )
        xs_die( "shim kernel", (int)e );
}

// 2-D block <-> staging
static void xs_put( size_t off, const pixel *src, intptr_t stride, int w, int h, int pitch )
{
    for( int y = 0; y < h; y++ )
        memcpy( g_h + off + (size_t)y * pitch, src + y * stride, w );
}
static void xs_get( pixel *dst, intptr_t stride, size_t off, int w, int h, int pitch )
{
    for( int y = 0; y < h; y++ )
        memcpy( dst + y * stride, g_h + off + (size_t)y * pitch, w );
}

static const uint8_t xs_w[8] = { 16, 16, 8, 8, 8, 4, 4, 4 }, xs_h[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

// ---------------------------------------------------------------------------------------------
// x264_pixel_function_t

static void xs_cmp_n( int cmp, int size, pixel *fenc, intptr_t s1, pixel *const *refs, intptr_t s2, int n, int *scores )
{
    const int w = xs_w[size], h = xs_h[size];
    xs_put( 0, fenc, s1, w, h, 16 );
    for( int k = 0; k < n; k++ )
        xs_put( 256 + 256 * k, refs[k], s2, w, h, 16 );
    xd_shim_args A = { OP_CMP, { cmp, w, h, n, 0, 256, 256, 2048 } };
    xs_run( A, 32 );
    memcpy( scores, g_h + 2048, n * sizeof( int ) );
}

#define XS_CMP1( name, cmp, size ) \
    static int name( pixel *p1, intptr_t s1, pixel *p2, intptr_t s2 ) \
    { int r; pixel *refs[1] = { p2 }; xs_cmp_n( cmp, size, p1, s1, refs, s2, 1, &r ); return r; }
#define XS_CMP3( name, cmp, size ) \
    static void name( pixel *fenc, pixel *p0, pixel *p1, pixel *p2, intptr_t s, int scores[3] ) \
    { pixel *refs[3] = { p0, p1, p2 }; xs_cmp_n( cmp, size, fenc, FENC_STRIDE, refs, s, 3, scores ); }
#define XS_CMP4( name, cmp, size ) \
    static void name( pixel *fenc, pixel *p0, pixel *p1, pixel *p2, pixel *p3, intptr_t s, int scores[4] ) \
    { pixel *refs[4] = { p0, p1, p2, p3 }; xs_cmp_n( cmp, size, fenc, FENC_STRIDE, refs, s, 4, scores ); }
#define XS_SEVEN_SIZES( M, prefix, cmp ) \
    M( prefix##_16x16, cmp, 0 ) M( prefix##_16x8, cmp, 1 ) M( prefix##_8x16, cmp, 2 ) M( prefix##_8x8, cmp, 3 ) \
    M( prefix##_8x4, cmp, 4 ) M( prefix##_4x8, cmp, 5 ) M( prefix##_4x4, cmp, 6 )
#define XS_ALL_SIZES( M, prefix, cmp ) XS_SEVEN_SIZES( M, prefix, cmp ) M( prefix##_4x16, cmp, 7 )

XS_ALL_SIZES( XS_CMP1, xs_sad, CMP_SAD )
XS_ALL_SIZES( XS_CMP1, xs_ssd, CMP_SSD )
XS_ALL_SIZES( XS_CMP1, xs_satd, CMP_SATD )
XS_SEVEN_SIZES( XS_CMP3, xs_sad_x3, CMP_SAD )
XS_SEVEN_SIZES( XS_CMP4, xs_sad_x4, CMP_SAD )
XS_SEVEN_SIZES( XS_CMP3, xs_satd_x3, CMP_SATD )
XS_SEVEN_SIZES( XS_CMP4, xs_satd_x4, CMP_SATD )

static uint64_t xs_var( pixel *pix, intptr_t stride, int w, int h )
{
    xs_put( 0, pix, stride, w, h, 16 );
    xd_shim_args A = { OP_VAR, { w, h, 0, 2048 } };
    xs_run( A, 32 );
    uint64_t r;
    memcpy( &r, g_h + 2048, sizeof( r ) );
    return r;
}
static uint64_t xs_var_16x16( pixel *pix, intptr_t stride ) { return xs_var( pix, stride, 16, 16 ); }
static uint64_t xs_var_8x8( pixel *pix, intptr_t stride ) { return xs_var( pix, stride, 8, 8 ); }

static int xs_var2_8x8( pixel *p1, intptr_t s1, pixel *p2, intptr_t s2, int *ssd )
{
    xs_put( 0, p1, s1, 8, 8, 16 );
    xs_put( 256, p2, s2, 8, 8, 16 );
    xd_shim_args A = { OP_VAR2, { 0, 256, 2048 } };
    xs_run( A, 32 );
    const int *r = (const int *)( g_h + 2048 );
    *ssd = r[1];
    return r[0];
}

// fenc block + fdec neighbourhood -> predictions, costs, last prediction left in fdec
static void xs_intra( int cmp, int size, pixel *fenc, pixel *fdec, int n_modes, const int *modes, const int *res_index, int *res )
{
    const size_t F = 1024 + FDEC_STRIDE + 16;          // block origin inside the staged window
    xs_put( 0, fenc, FENC_STRIDE, size, size, FENC_STRIDE );
    // top row with the corner, plus the top-right four for 4x4; left column
    memcpy( g_h + F - FDEC_STRIDE - 1, fdec - FDEC_STRIDE - 1, 1 + ( size == 4 ? 8 : size ) );
    for( int y = 0; y < size; y++ )
        g_h[F + y * FDEC_STRIDE - 1] = fdec[y * FDEC_STRIDE - 1];
    xd_shim_args A = { OP_INTRA, { cmp, size, n_modes, 0, (int)F, 4096 } };
    for( int m = 0; m < n_modes; m++ )
    {
        A.a[6 + m] = modes[m];
        A.a[10 + m] = res_index[m];
    }
    xs_run( A, 32 );
    const int *r = (const int *)( g_h + 4096 );
    for( int m = 0; m < n_modes; m++ )
        res[res_index[m]] = r[res_index[m]];
    xs_get( fdec, FDEC_STRIDE, F, size, size, FDEC_STRIDE );
}

// the neighbourhood of the block at src -> the predicted block, in place (x264_predict_t, predict.h:8).
// The same bytes are read that the reference's predictors may read: the row above with its corner (and the four
// top-right samples for 4x4) and the column to the left -- fdec_buf always has them (common/macroblock.c:242-265).
static void xs_predict_call( int size, int mode, pixel *src )
{
    const size_t F = 1024 + FDEC_STRIDE + 16;
    memcpy( g_h + F - FDEC_STRIDE - 1, src - FDEC_STRIDE - 1, 1 + ( size == 4 ? 8 : size ) );
    for( int y = 0; y < size; y++ )
        g_h[F + y * FDEC_STRIDE - 1] = src[y * FDEC_STRIDE - 1];
    xd_shim_args A = { OP_PREDICT, { size, mode, (int)F } };
    xs_run( A, 32 );
    xs_get( src, FDEC_STRIDE, F, size, size, FDEC_STRIDE );
}
#define XS_PRED( name, size, mode ) static void name( pixel *src ) { xs_predict_call( size, mode, src ); }
XS_PRED( xs_p16_v, 16, PR_V ) XS_PRED( xs_p16_h, 16, PR_H ) XS_PRED( xs_p16_dc, 16, PR_DC ) XS_PRED( xs_p16_p, 16, PR_PLANE )
XS_PRED( xs_p16_dcl, 16, PR_DC_LEFT ) XS_PRED( xs_p16_dct, 16, PR_DC_TOP ) XS_PRED( xs_p16_128, 16, PR_DC_128 )
XS_PRED( xs_p8_v, 8, PR_V ) XS_PRED( xs_p8_h, 8, PR_H ) XS_PRED( xs_p8_dc, 8, PR_DC ) XS_PRED( xs_p8_p, 8, PR_PLANE )
XS_PRED( xs_p8_dcl, 8, PR_DC_LEFT ) XS_PRED( xs_p8_dct, 8, PR_DC_TOP ) XS_PRED( xs_p8_128, 8, PR_DC_128 )
XS_PRED( xs_p4_v, 4, PR_V ) XS_PRED( xs_p4_h, 4, PR_H ) XS_PRED( xs_p4_dc, 4, PR_DC ) XS_PRED( xs_p4_ddl, 4, PR_DDL )
XS_PRED( xs_p4_ddr, 4, PR_DDR ) XS_PRED( xs_p4_vr, 4, PR_VR ) XS_PRED( xs_p4_hd, 4, PR_HD ) XS_PRED( xs_p4_vl, 4, PR_VL )
XS_PRED( xs_p4_hu, 4, PR_HU ) XS_PRED( xs_p4_dcl, 4, PR_DC_LEFT ) XS_PRED( xs_p4_dct, 4, PR_DC_TOP ) XS_PRED( xs_p4_128, 4, PR_DC_128 )

// common/predict.c:474-546; the table order is the reference's enums (predict.h:10-59)
extern "C" void x264_predict_16x16_init( int cpu, x264_predict_t pf[7] )
{
    (void)cpu;
    xs_open();
    pf[0] = xs_p16_v; pf[1] = xs_p16_h; pf[2] = xs_p16_dc; pf[3] = xs_p16_p;
    pf[4] = xs_p16_dcl; pf[5] = xs_p16_dct; pf[6] = xs_p16_128;
}
extern "C" void x264_predict_8x8c_init( int cpu, x264_predict_t pf[7] )
{
    (void)cpu;
    xs_open();
    pf[0] = xs_p8_dc; pf[1] = xs_p8_h; pf[2] = xs_p8_v; pf[3] = xs_p8_p;
    pf[4] = xs_p8_dcl; pf[5] = xs_p8_dct; pf[6] = xs_p8_128;
}
extern "C" void x264_predict_4x4_init( int cpu, x264_predict_t pf[12] )
{
    (void)cpu;
    xs_open();
    pf[0] = xs_p4_v; pf[1] = xs_p4_h; pf[2] = xs_p4_dc; pf[3] = xs_p4_ddl; pf[4] = xs_p4_ddr; pf[5] = xs_p4_vr;
    pf[6] = xs_p4_hd; pf[7] = xs_p4_vl; pf[8] = xs_p4_hu; pf[9] = xs_p4_dcl; pf[10] = xs_p4_dct; pf[11] = xs_p4_128;
}

static const int xs_idx3[3] = { 0, 1, 2 };
#define XS_INTRA3( name, cmp, size, m0, m1, m2 ) \
    static void name( pixel *fenc, pixel *fdec, int res[3] ) \
    { const int modes[3] = { m0, m1, m2 }; xs_intra( cmp, size, fenc, fdec, 3, modes, xs_idx3, res ); }
XS_INTRA3( xs_intra_sad_x3_4x4, CMP_SAD, 4, PR_V, PR_H, PR_DC )
XS_INTRA3( xs_intra_satd_x3_4x4, CMP_SATD, 4, PR_V, PR_H, PR_DC )
XS_INTRA3( xs_intra_sad_x3_8x8c, CMP_SAD, 8, PR_DC, PR_H, PR_V )
XS_INTRA3( xs_intra_satd_x3_8x8c, CMP_SATD, 8, PR_DC, PR_H, PR_V )
XS_INTRA3( xs_intra_sad_x3_16x16, CMP_SAD, 16, PR_V, PR_H, PR_DC )
XS_INTRA3( xs_intra_satd_x3_16x16, CMP_SATD, 16, PR_V, PR_H, PR_DC )

static void xs_intra_satd_x4_4x4_h( pixel *fenc, pixel *fdec, int res[9] )
{
    const int modes[4] = { PR_DDL, PR_DDR, PR_HD, PR_HU }, idx[4] = { 3, 4, 6, 8 };
    xs_intra( CMP_SATD, 4, fenc, fdec, 4, modes, idx, res );
}
static void xs_intra_satd_x4_4x4_v( pixel *fenc, pixel *fdec, int res[9] )
{
    const int modes[4] = { PR_DDL, PR_DDR, PR_VR, PR_VL }, idx[4] = { 3, 4, 5, 7 };
    xs_intra( CMP_SATD, 4, fenc, fdec, 4, modes, idx, res );
}

extern "C" void x264_pixel_init( int cpu, x264_pixel_function_t *pixf )
{
    (void)cpu;
    xs_open();
    memset( pixf, 0, sizeof( *pixf ) );
#define XS_SET8( member, prefix ) \
    pixf->member[0] = prefix##_16x16; pixf->member[1] = prefix##_16x8; pixf->member[2] = prefix##_8x16; \
    pixf->member[3] = prefix##_8x8; pixf->member[4] = prefix##_8x4; pixf->member[5] = prefix##_4x8; \
    pixf->member[6] = prefix##_4x4;
    XS_SET8( sad, xs_sad )          pixf->sad[7] = xs_sad_4x16;
    XS_SET8( sad_aligned, xs_sad )  pixf->sad_aligned[7] = xs_sad_4x16;
    XS_SET8( ssd, xs_ssd )          pixf->ssd[7] = xs_ssd_4x16;
    XS_SET8( satd, xs_satd )        pixf->satd[7] = xs_satd_4x16;
    XS_SET8( sad_x3, xs_sad_x3 )
    XS_SET8( sad_x4, xs_sad_x4 )
    XS_SET8( satd_x3, xs_satd_x3 )
    XS_SET8( satd_x4, xs_satd_x4 )
#undef XS_SET8
    pixf->var[0] = xs_var_16x16;            // PIXEL_16x16
    pixf->var[3] = xs_var_8x8;              // PIXEL_8x8
    pixf->var2[3] = xs_var2_8x8;
    pixf->intra_sad_x3_4x4 = xs_intra_sad_x3_4x4;
    pixf->intra_satd_x3_4x4 = xs_intra_satd_x3_4x4;
    pixf->intra_sad_x3_8x8c = xs_intra_sad_x3_8x8c;
    pixf->intra_satd_x3_8x8c = xs_intra_satd_x3_8x8c;
    pixf->intra_sad_x3_16x16 = xs_intra_sad_x3_16x16;
    pixf->intra_satd_x3_16x16 = xs_intra_satd_x3_16x16;
    pixf->intra_satd_x4_4x4_h = xs_intra_satd_x4_4x4_h;
    pixf->intra_satd_x4_4x4_v = xs_intra_satd_x4_4x4_v;
}

// ---------------------------------------------------------------------------------------------
// x264_dct_function_t / x264_zigzag_function_t
// staging: fenc block at 0 (stride 16), fdec block at 1024 (stride 32), coefficients at 2048

static void xs_sub_dct( dctcoef *dct, pixel *pix1, pixel *pix2, int n )
{
    const int size = n == 1 ? 4 : n == 4 ? 8 : 16;
    xs_put( 0, pix1, FENC_STRIDE, size, size, FENC_STRIDE );
    xs_put( 1024, pix2, FDEC_STRIDE, size, size, FDEC_STRIDE );
    xd_shim_args A = { OP_SUB_DCT, { n, 0, 1024, 2048 } };
    xs_run( A, n );
    memcpy( dct, g_h + 2048, (size_t)n * 16 * sizeof( dctcoef ) );
}
static void xs_sub4x4_dct( dctcoef dct[16], pixel *p1, pixel *p2 ) { xs_sub_dct( dct, p1, p2, 1 ); }
static void xs_sub8x8_dct( dctcoef dct[4][16], pixel *p1, pixel *p2 ) { xs_sub_dct( &dct[0][0], p1, p2, 4 ); }
static void xs_sub16x16_dct( dctcoef dct[16][16], pixel *p1, pixel *p2 ) { xs_sub_dct( &dct[0][0], p1, p2, 16 ); }

static void xs_sub8x8_dct_dc( dctcoef dct[4], pixel *pix1, pixel *pix2 )
{
    xs_put( 0, pix1, FENC_STRIDE, 8, 8, FENC_STRIDE );
    xs_put( 1024, pix2, FDEC_STRIDE, 8, 8, FDEC_STRIDE );
    xd_shim_args A = { OP_SUB_DCT_DC, { 0, 1024, 2048 } };
    xs_run( A, 32 );
    memcpy( dct, g_h + 2048, 4 * sizeof( dctcoef ) );
}

static void xs_add_idct( pixel *dst, const dctcoef *dct, int n, int dc_only )
{
    const int size = dc_only ? ( n == 4 ? 8 : 16 ) : ( n == 1 ? 4 : n == 4 ? 8 : 16 );
    xs_put( 1024, dst, FDEC_STRIDE, size, size, FDEC_STRIDE );
    memcpy( g_h + 2048, dct, (size_t)n * ( dc_only ? 1 : 16 ) * sizeof( dctcoef ) );
    xd_shim_args A = { dc_only ? OP_ADD_IDCT_DC : OP_ADD_IDCT, { n, 1024, 2048 } };
    xs_run( A, n );
    xs_get( dst, FDEC_STRIDE, 1024, size, size, FDEC_STRIDE );
}
static void xs_add4x4_idct( pixel *dst, dctcoef dct[16] ) { xs_add_idct( dst, dct, 1, 0 ); }
static void xs_add8x8_idct( pixel *dst, dctcoef dct[4][16] ) { xs_add_idct( dst, &dct[0][0], 4, 0 ); }
static void xs_add16x16_idct( pixel *dst, dctcoef dct[16][16] ) { xs_add_idct( dst, &dct[0][0], 16, 0 ); }
static void xs_add8x8_idct_dc( pixel *dst, dctcoef dct[4] ) { xs_add_idct( dst, dct, 4, 1 ); }
static void xs_add16x16_idct_dc( pixel *dst, dctcoef dct[16] ) { xs_add_idct( dst, dct, 16, 1 ); }

static void xs_dc4x4( dctcoef d[16], int op )
{
    memcpy( g_h + 2048, d, 16 * sizeof( dctcoef ) );
    xd_shim_args A = { op, { 2048 } };
    xs_run( A, 32 );
    memcpy( d, g_h + 2048, 16 * sizeof( dctcoef ) );
}
static void xs_dct4x4dc( dctcoef d[16] ) { xs_dc4x4( d, OP_DCT4X4DC ); }
static void xs_idct4x4dc( dctcoef d[16] ) { xs_dc4x4( d, OP_IDCT4X4DC ); }

static void xs_scan_4x4( dctcoef level[16], dctcoef dct[16] )
{
    memcpy( g_h + 2048, dct, 16 * sizeof( dctcoef ) );
    xd_shim_args A = { OP_ZIGZAG, { 2048, 2112 } };
    xs_run( A, 32 );
    memcpy( level, g_h + 2112, 16 * sizeof( dctcoef ) );
}

extern "C" void x264_dct_init( int cpu, x264_dct_function_t *dctf )
{
    (void)cpu;
    xs_open();
    dctf->sub4x4_dct = xs_sub4x4_dct;
    dctf->add4x4_idct = xs_add4x4_idct;
    dctf->sub8x8_dct = xs_sub8x8_dct;
    dctf->sub8x8_dct_dc = xs_sub8x8_dct_dc;
    dctf->add8x8_idct = xs_add8x8_idct;
    dctf->add8x8_idct_dc = xs_add8x8_idct_dc;
    dctf->sub16x16_dct = xs_sub16x16_dct;
    dctf->add16x16_idct = xs_add16x16_idct;
    dctf->add16x16_idct_dc = xs_add16x16_idct_dc;
    dctf->dct4x4dc = xs_dct4x4dc;
    dctf->idct4x4dc = xs_idct4x4dc;
}

extern "C" void x264_zigzag_init( int cpu, x264_zigzag_function_t *zigzagf )
{
    (void)cpu;
    xs_open();
    zigzagf->scan_4x4 = xs_scan_4x4;
}

// ---------------------------------------------------------------------------------------------
// x264_quant_function_t.  staging: coefficients at 0, tables at 256 / 512, results at 1024

static int xs_quant_4x4( dctcoef dct[16], udctcoef mf[16], udctcoef bias[16] )
{
    memcpy( g_h, dct, 32 );
    memcpy( g_h + 256, mf, 32 );
    memcpy( g_h + 512, bias, 32 );
    xd_shim_args A = { OP_QUANT, { 16, 0, 256, 512, 1024 } };
    xs_run( A, 32 );
    memcpy( dct, g_h, 32 );
    return *(const int *)( g_h + 1024 );
}
static int xs_quant_dc( dctcoef *dct, int n, int mf, int bias )
{
    memcpy( g_h, dct, n * sizeof( dctcoef ) );
    xd_shim_args A = { OP_QUANT_DC, { n, 0, mf, bias, 1024 } };
    xs_run( A, 32 );
    memcpy( dct, g_h, n * sizeof( dctcoef ) );
    return *(const int *)( g_h + 1024 );
}
static int xs_quant_4x4_dc( dctcoef dct[16], int mf, int bias ) { return xs_quant_dc( dct, 16, mf, bias ); }
static int xs_quant_2x2_dc( dctcoef dct[4], int mf, int bias ) { return xs_quant_dc( dct, 4, mf, bias ); }

static void xs_dequant( dctcoef dct[16], int dequant_mf[6][16], int qp, int op )
{
    memcpy( g_h, dct, 32 );
    memcpy( g_h + 256, dequant_mf[qp % 6], 16 * sizeof( int ) );
    xd_shim_args A = { op, { 0, 256, qp } };
    xs_run( A, 32 );
    memcpy( dct, g_h, 32 );
}
static void xs_dequant_4x4( dctcoef dct[16], int dequant_mf[6][16], int qp ) { xs_dequant( dct, dequant_mf, qp, OP_DEQUANT ); }
static void xs_dequant_4x4_dc( dctcoef dct[16], int dequant_mf[6][16], int qp ) { xs_dequant( dct, dequant_mf, qp, OP_DEQUANT_DC ); }

static int xs_optimize_chroma_2x2_dc( dctcoef dct[4], int dequant_mf )
{
    memcpy( g_h, dct, 8 );
    xd_shim_args A = { OP_OPT_CHROMA_DC, { 0, dequant_mf, 1024 } };
    xs_run( A, 32 );
    memcpy( dct, g_h, 8 );
    return *(const int *)( g_h + 1024 );
}

static void xs_denoise_dct( dctcoef *dct, uint32_t *sum, udctcoef *offset, int size )
{
    const size_t o_sum = 256, o_off = 256 + 64 * 4;
    memcpy( g_h, dct, size * sizeof( dctcoef ) );
    memcpy( g_h + o_sum, sum, size * sizeof( uint32_t ) );
    memcpy( g_h + o_off, offset, size * sizeof( udctcoef ) );
    xd_shim_args A = { OP_DENOISE, { 0, (int)o_sum, (int)o_off, size } };
    xs_run( A, size );
    memcpy( dct, g_h, size * sizeof( dctcoef ) );
    memcpy( sum, g_h + o_sum, size * sizeof( uint32_t ) );
}

static int xs_decimate( dctcoef *dct, int first )
{
    memcpy( g_h, dct, 32 );
    xd_shim_args A = { OP_DECIMATE, { 0, first, 1024 } };
    xs_run( A, 32 );
    return *(const int *)( g_h + 1024 );
}
static int xs_decimate_score15( dctcoef *dct ) { return xs_decimate( dct, 1 ); }
static int xs_decimate_score16( dctcoef *dct ) { return xs_decimate( dct, 0 ); }

static int xs_coeff_last( dctcoef *dct, int n )
{
    memcpy( g_h, dct, n * sizeof( dctcoef ) );
    xd_shim_args A = { OP_COEFF_LAST, { 0, n, 1024 } };
    xs_run( A, 32 );
    return *(const int *)( g_h + 1024 );
}
static int xs_coeff_last4( dctcoef *d ) { return xs_coeff_last( d, 4 ); }
static int xs_coeff_last8( dctcoef *d ) { return xs_coeff_last( d, 8 ); }
static int xs_coeff_last15( dctcoef *d ) { return xs_coeff_last( d, 15 ); }
static int xs_coeff_last16( dctcoef *d ) { return xs_coeff_last( d, 16 ); }
static int xs_coeff_last64( dctcoef *d ) { return xs_coeff_last( d, 64 ); }

// x264_run_level_t (common/bitstream.h:33-38): { int last; int mask; dctcoef level[16]; }
static int xs_level_run( dctcoef *dct, x264_run_level_t *runlevel, int n )
{
    memcpy( g_h, dct, n * sizeof( dctcoef ) );
    xd_shim_args A = { OP_LEVEL_RUN, { 0, n, 256, 1024 } };
    xs_run( A, 32 );
    const int total = *(const int *)( g_h + 1024 );
    memcpy( runlevel, g_h + 256, 2 * sizeof( int ) + total * sizeof( dctcoef ) );
    return total;
}
static int xs_level_run4( dctcoef *d, x264_run_level_t *r ) { return xs_level_run( d, r, 4 ); }
static int xs_level_run8( dctcoef *d, x264_run_level_t *r ) { return xs_level_run( d, r, 8 ); }
static int xs_level_run15( dctcoef *d, x264_run_level_t *r ) { return xs_level_run( d, r, 15 ); }
static int xs_level_run16( dctcoef *d, x264_run_level_t *r ) { return xs_level_run( d, r, 16 ); }

// DCT_* block categories (common/macroblock.h:270-286)
enum { XS_DCT_LUMA_DC = 0, XS_DCT_LUMA_AC = 1, XS_DCT_LUMA_4x4 = 2, XS_DCT_CHROMA_DC = 3, XS_DCT_CHROMA_AC = 4,
       XS_DCT_LUMA_8x8 = 5, XS_DCT_CHROMAU_DC = 6, XS_DCT_CHROMAU_AC = 7, XS_DCT_CHROMAU_4x4 = 8, XS_DCT_CHROMAU_8x8 = 9,
       XS_DCT_CHROMAV_DC = 10, XS_DCT_CHROMAV_AC = 11, XS_DCT_CHROMAV_4x4 = 12, XS_DCT_CHROMAV_8x8 = 13 };

extern "C" void x264_quant_init( x264_t *h, int cpu, x264_quant_function_t *pf )
{
    (void)h; (void)cpu;
    xs_open();
    pf->quant_4x4 = xs_quant_4x4;
    pf->quant_4x4_dc = xs_quant_4x4_dc;
    pf->quant_2x2_dc = xs_quant_2x2_dc;
    pf->dequant_4x4 = xs_dequant_4x4;
    pf->dequant_4x4_dc = xs_dequant_4x4_dc;
    pf->optimize_chroma_2x2_dc = xs_optimize_chroma_2x2_dc;
    pf->denoise_dct = xs_denoise_dct;
    pf->decimate_score15 = xs_decimate_score15;
    pf->decimate_score16 = xs_decimate_score16;
    pf->coeff_last4 = xs_coeff_last4;
    pf->coeff_last8 = xs_coeff_last8;
    pf->coeff_level_run4 = xs_level_run4;
    pf->coeff_level_run8 = xs_level_run8;
    // quant.c:320-334: which categories share which scan length
    static const uint8_t len16[] = { XS_DCT_LUMA_4x4, XS_DCT_LUMA_DC, XS_DCT_CHROMAU_DC, XS_DCT_CHROMAV_DC, XS_DCT_CHROMAU_4x4, XS_DCT_CHROMAV_4x4 };
    static const uint8_t len15[] = { XS_DCT_LUMA_AC, XS_DCT_CHROMA_AC, XS_DCT_CHROMAU_AC, XS_DCT_CHROMAV_AC };
    static const uint8_t len64[] = { XS_DCT_LUMA_8x8, XS_DCT_CHROMAU_8x8, XS_DCT_CHROMAV_8x8 };
    for( unsigned i = 0; i < sizeof( len16 ); i++ )
    {
        pf->coeff_last[len16[i]] = xs_coeff_last16;
        pf->coeff_level_run[len16[i]] = xs_level_run16;
    }
    for( unsigned i = 0; i < sizeof( len15 ); i++ )
    {
        pf->coeff_last[len15[i]] = xs_coeff_last15;
        pf->coeff_level_run[len15[i]] = xs_level_run15;
    }
    for( unsigned i = 0; i < sizeof( len64 ); i++ )
        pf->coeff_last[len64[i]] = xs_coeff_last64;
}

// ---------------------------------------------------------------------------------------------
// x264_mc_functions_t

static const uint8_t xs_hpel_ref0[16] = { 0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1 };   // mc.c:192
static const uint8_t xs_hpel_ref1[16] = { 0, 0, 0, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2 };   // mc.c:193

static void xs_copy2d( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int w, int h )
{
    const size_t n = (size_t)w * h;
    xs_reserve( 2 * n + 64 );
    xs_put( 0, src, i_src, w, h, w );
    xd_shim_args A = { OP_COPY, { 0, (int)n, (int)n } };
    xs_run( A, (int)n );
    xs_get( dst, i_dst, n, w, h, w );
}

static void xs_avg2d( pixel *dst, intptr_t i_dst, pixel *s1, pixel *s2, intptr_t i_src, int w, int h )
{
    const size_t n = (size_t)w * h;
    xs_reserve( 3 * n + 64 );
    xs_put( 0, s1, i_src, w, h, w );
    xs_put( n, s2, i_src, w, h, w );
    xd_shim_args A = { OP_AVG, { 0, (int)n, (int)( 2 * n ), (int)n } };
    xs_run( A, (int)n );
    xs_get( dst, i_dst, 2 * n, w, h, w );
}

static void xs_mc_luma( pixel *dst, intptr_t i_dst, pixel **src, intptr_t i_src, int mvx, int mvy, int w, int h,
                        const x264_weight_t *weight )
{
    (void)weight;                                          // ignored by the reference as well (mc.c:216-239)
    const int qpel = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    const intptr_t offset = ( mvy >> 2 ) * i_src + ( mvx >> 2 );
    pixel *s1 = src[xs_hpel_ref0[qpel]] + offset + ( ( mvy & 3 ) == 3 ) * i_src;
    if( qpel & 5 )
        xs_avg2d( dst, i_dst, s1, src[xs_hpel_ref1[qpel]] + offset + ( ( mvx & 3 ) == 3 ), i_src, w, h );
    else
        xs_copy2d( dst, i_dst, s1, i_src, w, h );
}

static pixel *xs_get_ref( pixel *dst, intptr_t *i_dst, pixel **src, intptr_t i_src, int mvx, int mvy, int w, int h,
                          const x264_weight_t *weight )
{
    (void)weight;
    const int qpel = ( ( mvy & 3 ) << 2 ) + ( mvx & 3 );
    const intptr_t offset = ( mvy >> 2 ) * i_src + ( mvx >> 2 );
    pixel *s1 = src[xs_hpel_ref0[qpel]] + offset + ( ( mvy & 3 ) == 3 ) * i_src;
    if( qpel & 5 )
    {
        xs_avg2d( dst, *i_dst, s1, src[xs_hpel_ref1[qpel]] + offset + ( ( mvx & 3 ) == 3 ), i_src, w, h );
        return dst;
    }
    *i_dst = i_src;                                        // mc.c:259-263: hand back the plane itself
    return s1;
}

static void xs_mc_chroma( pixel *dstu, pixel *dstv, intptr_t i_dst, pixel *src, intptr_t i_src, int mvx, int mvy, int w, int h )
{
    src += ( mvy >> 3 ) * i_src + ( mvx >> 3 ) * 2;
    const int pitch = 2 * w + 2, n = w * h;
    const size_t o_u = (size_t)pitch * ( h + 1 ), o_v = o_u + n;
    xs_reserve( o_v + n + 64 );
    xs_put( 0, src, i_src, pitch, h + 1, pitch );
    xd_shim_args A = { OP_MC_CHROMA, { 0, pitch, (int)o_u, (int)o_v, w, h, mvx & 7, mvy & 7 } };
    xs_run( A, n );
    xs_get( dstu, i_dst, o_u, w, h, w );
    xs_get( dstv, i_dst, o_v, w, h, w );
}

static void xs_copy_w16( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int h ) { xs_copy2d( dst, i_dst, src, i_src, 16, h ); }
static void xs_copy_w8( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int h ) { xs_copy2d( dst, i_dst, src, i_src, 8, h ); }
static void xs_copy_w4( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int h ) { xs_copy2d( dst, i_dst, src, i_src, 4, h ); }
static void xs_plane_copy( pixel *dst, intptr_t i_dst, pixel *src, intptr_t i_src, int w, int h ) { xs_copy2d( dst, i_dst, src, i_src, w, h ); }

static void xs_plane_copy_interleave( pixel *dst, intptr_t i_dst, pixel *srcu, intptr_t i_srcu, pixel *srcv, intptr_t i_srcv, int w, int h )
{
    const size_t n = (size_t)w * h;
    xs_reserve( 4 * n + 64 );
    xs_put( 0, srcu, i_srcu, w, h, w );
    xs_put( n, srcv, i_srcv, w, h, w );
    xd_shim_args A = { OP_INTERLEAVE, { 0, (int)n, (int)( 2 * n ), (int)n } };
    xs_run( A, (int)n );
    xs_get( dst, i_dst, 2 * n, 2 * w, h, 2 * w );
}

static void xs_plane_copy_deinterleave( pixel *dstu, intptr_t i_dstu, pixel *dstv, intptr_t i_dstv, pixel *src, intptr_t i_src, int w, int h )
{
    const size_t n = (size_t)w * h;
    xs_reserve( 4 * n + 64 );
    xs_put( 0, src, i_src, 2 * w, h, 2 * w );
    xd_shim_args A = { OP_DEINTERLEAVE, { 0, (int)( 2 * n ), (int)( 3 * n ), (int)n } };
    xs_run( A, (int)n );
    xs_get( dstu, i_dstu, 2 * n, w, h, w );
    xs_get( dstv, i_dstv, 3 * n, w, h, w );
}

static void xs_store_interleave_chroma( pixel *dst, intptr_t i_dst, pixel *srcu, pixel *srcv, int height )
{
    xs_plane_copy_interleave( dst, i_dst, srcu, FDEC_STRIDE, srcv, FDEC_STRIDE, 8, height );
}
static void xs_load_deinterleave_chroma_fenc( pixel *dst, pixel *src, intptr_t i_src, int height )
{
    xs_plane_copy_deinterleave( dst, FENC_STRIDE, dst + FENC_STRIDE / 2, FENC_STRIDE, src, i_src, 8, height );
}
static void xs_load_deinterleave_chroma_fdec( pixel *dst, pixel *src, intptr_t i_src, int height )
{
    xs_plane_copy_deinterleave( dst, FDEC_STRIDE, dst + FDEC_STRIDE / 2, FDEC_STRIDE, src, i_src, 8, height );
}

static void xs_hpel_filter( pixel *dsth, pixel *dstv, pixel *dstc, pixel *src, intptr_t stride, int width, int height, int16_t *buf )
{
    (void)buf;                       // the int16 line buffer lives in registers on the device
    // needs rows -2..height+2 and columns -4..width+5 of src
    const int pitch = width + 10, rows = height + 5;
    const size_t o_src = 0, o_h = (size_t)pitch * rows, o_v = o_h + (size_t)width * height;
    const size_t o_c = o_v + (size_t)( width + 5 ) * height;
    xs_reserve( o_c + (size_t)width * height + 64 );
    xs_put( o_src, src - 2 * stride - 4, stride, pitch, rows, pitch );
    xd_shim_args A = { OP_HPEL, { (int)( o_src + 2 * pitch + 4 ), pitch, width, height, (int)o_h, (int)o_v, (int)o_c } };
    xs_run( A, ( width + 5 ) * height );
    xs_get( dsth, stride, o_h, width, height, width );
    xs_get( dstv - 2, stride, o_v, width + 5, height, width + 5 );
    xs_get( dstc, stride, o_c, width, height, width );
}

static void xs_frame_init_lowres_core( pixel *src0, pixel *dst0, pixel *dsth, pixel *dstv, pixel *dstc,
                                       intptr_t src_stride, intptr_t dst_stride, int width, int height )
{
    const int sp = 2 * width + 1, rows = 2 * height + 1;
    const size_t n = (size_t)width * height, o0 = (size_t)sp * rows;
    xs_reserve( o0 + 4 * n + 64 );
    xs_put( 0, src0, src_stride, sp, rows, sp );
    xd_shim_args A = { OP_LOWRES, { 0, sp, (int)o0, (int)( o0 + n ), (int)( o0 + 2 * n ), (int)( o0 + 3 * n ), width, height } };
    xs_run( A, (int)n );
    xs_get( dst0, dst_stride, o0, width, height, width );
    xs_get( dsth, dst_stride, o0 + n, width, height, width );
    xs_get( dstv, dst_stride, o0 + 2 * n, width, height, width );
    xs_get( dstc, dst_stride, o0 + 3 * n, width, height, width );
}

static void xs_prefetch_fenc_null( pixel *pix_y, intptr_t stride_y, pixel *pix_uv, intptr_t stride_uv, int mb_x )
{
    (void)pix_y; (void)stride_y; (void)pix_uv; (void)stride_uv; (void)mb_x;
}
static void xs_prefetch_ref_null( pixel *pix, intptr_t stride, int parity ) { (void)pix; (void)stride; (void)parity; }
static void xs_memzero_aligned( void *dst, size_t n ) { memset( dst, 0, n ); }

extern "C" void x264_mc_init( int cpu, x264_mc_functions_t *pf )
{
    (void)cpu;
    xs_open();
    pf->mc_luma = xs_mc_luma;
    pf->get_ref = xs_get_ref;
    pf->mc_chroma = xs_mc_chroma;
    pf->copy[0] = xs_copy_w16;              // PIXEL_16x16
    pf->copy[3] = xs_copy_w8;               // PIXEL_8x8
    pf->copy[6] = xs_copy_w4;               // PIXEL_4x4
    pf->store_interleave_chroma = xs_store_interleave_chroma;
    pf->load_deinterleave_chroma_fenc = xs_load_deinterleave_chroma_fenc;
    pf->load_deinterleave_chroma_fdec = xs_load_deinterleave_chroma_fdec;
    pf->plane_copy = xs_plane_copy;
    pf->plane_copy_interleave = xs_plane_copy_interleave;
    pf->plane_copy_deinterleave = xs_plane_copy_deinterleave;
    pf->hpel_filter = xs_hpel_filter;
    pf->prefetch_fenc_420 = xs_prefetch_fenc_null;
    pf->prefetch_ref = xs_prefetch_ref_null;
    pf->memcpy_aligned = memcpy;            // libc in the reference too (mc.c:487)
    pf->memzero_aligned = xs_memzero_aligned;
    pf->frame_init_lowres_core = xs_frame_init_lowres_core;
}

// ---------------------------------------------------------------------------------------------
// x264_deblock_function_t.  The region around the edge is staged at pitch 16.

static void xs_deblock( pixel *pix, intptr_t stride, int alpha, int beta, const int8_t *tc0, int chroma, int dir )
{
    // dir 0 = vertical edge (filter across x): luma 16 rows x 8 bytes from pix-4, chroma 8 rows x 8 bytes from pix-4
    // dir 1 = horizontal edge (filter across y): luma 8 rows x 16 bytes from pix-4*stride, chroma 4 rows from pix-2*stride
    pixel *top = dir ? pix - ( chroma ? 2 : 4 ) * stride : pix - 4;
    const int w = dir ? 16 : 8, h = dir ? ( chroma ? 4 : 8 ) : ( chroma ? 8 : 16 );
    xs_put( 0, top, stride, w, h, 16 );
    xd_shim_args A = { chroma ? OP_DB_CHROMA : OP_DB_LUMA, { 0, dir, tc0 == NULL, alpha, beta } };
    for( int i = 0; i < 4; i++ )
        A.a[5 + i] = tc0 ? tc0[i] : 0;
    xs_run( A, 32 );
    xs_get( top, stride, 0, w, h, 16 );
}
static void xs_deblock_h_luma( pixel *pix, intptr_t stride, int alpha, int beta, int8_t *tc0 ) { xs_deblock( pix, stride, alpha, beta, tc0, 0, 0 ); }
static void xs_deblock_v_luma( pixel *pix, intptr_t stride, int alpha, int beta, int8_t *tc0 ) { xs_deblock( pix, stride, alpha, beta, tc0, 0, 1 ); }
static void xs_deblock_h_chroma( pixel *pix, intptr_t stride, int alpha, int beta, int8_t *tc0 ) { xs_deblock( pix, stride, alpha, beta, tc0, 1, 0 ); }
static void xs_deblock_v_chroma( pixel *pix, intptr_t stride, int alpha, int beta, int8_t *tc0 ) { xs_deblock( pix, stride, alpha, beta, tc0, 1, 1 ); }
static void xs_deblock_h_luma_intra( pixel *pix, intptr_t stride, int alpha, int beta ) { xs_deblock( pix, stride, alpha, beta, NULL, 0, 0 ); }
static void xs_deblock_v_luma_intra( pixel *pix, intptr_t stride, int alpha, int beta ) { xs_deblock( pix, stride, alpha, beta, NULL, 0, 1 ); }
static void xs_deblock_h_chroma_intra( pixel *pix, intptr_t stride, int alpha, int beta ) { xs_deblock( pix, stride, alpha, beta, NULL, 1, 0 ); }
static void xs_deblock_v_chroma_intra( pixel *pix, intptr_t stride, int alpha, int beta ) { xs_deblock( pix, stride, alpha, beta, NULL, 1, 1 ); }

static void xs_deblock_strength( uint8_t nnz[120], int8_t ref[2][40], int16_t mv[2][40][2], uint8_t bs[2][8][4] )
{
    // operands at fixed offsets of the staging buffer; the batched kernel does the work (n = 1)
    memcpy( g_h, nnz, 120 );
    memcpy( g_h + 128, ref, 80 );
    memcpy( g_h + 256, mv, 320 );
    cudaSetDevice( g_shim_ctx->device );
    int rc = x264dsp_deblock_strength_dev( g_shim_ctx, 1, g_d, (const int8_t *)( g_d + 128 ), (const int16_t *)( g_d + 256 ),
                                           g_d + 1024, NULL );
    if( rc == 0 )
        rc = (int)cudaStreamSynchronize( g_shim_ctx->stream );
    if( rc )
        xs_die( "deblock_strength", rc );
    // only bs[dir][0..3] are written by the reference (deblock.c:297-323)
    for( int dir = 0; dir < 2; dir++ )
        memcpy( bs[dir], g_h + 1024 + dir * 32, 16 );
}

extern "C" void x264_deblock_init( int cpu, x264_deblock_function_t *pf )
{
    (void)cpu;
    xs_open();
    pf->deblock_luma[0] = xs_deblock_h_luma;
    pf->deblock_luma[1] = xs_deblock_v_luma;
    pf->deblock_chroma[0] = xs_deblock_h_chroma;
    pf->deblock_chroma[1] = xs_deblock_v_chroma;
    pf->deblock_luma_intra[0] = xs_deblock_h_luma_intra;
    pf->deblock_luma_intra[1] = xs_deblock_v_luma_intra;
    pf->deblock_chroma_intra[0] = xs_deblock_h_chroma_intra;
    pf->deblock_chroma_intra[1] = xs_deblock_v_chroma_intra;
    pf->deblock_strength = xs_deblock_strength;
}
