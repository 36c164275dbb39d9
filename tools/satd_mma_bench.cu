// satd_mma_bench.cu -- "tensor cores only if ncu shows that an integer-MMA Hadamard beats the ALU path" (north_star),
// decided by measurement: batched SATD 8x8 (x264_pixel_satd_8x8, common/pixel.c:294-337) three ways on the same inputs,
// every result compared with a scalar CPU restatement.
//
//   alu   the form the kernels used before round 2: unpack the bytes, subtract, two butterfly stages, abs -- all on the
//         ALU pipe (~170 instructions per 4x4)
//   idp   the product's form (leaf.cuh xd_satd4x4): horizontal stage as u8 x s8 byte dot products with the +-1 rows of H4
//         (IDP.4A, FMA pipe), vertical stage a 32-bit butterfly whose last level is folded into max(|x|,|y|)
//   mma   the whole 2-D transform of a 4x4 DIFFERENCE as one K = 32 matrix product on the tensor cores:
//         row = [16 source pixels | 16 reference pixels] (u8), B = [ H4 (x) H4 ; -(H4 (x) H4) ] (s8, 32 x 16), i.e. two
//         mma.sync.m16n8k32.s32.u8.s8.s32 per sixteen 4x4 blocks, fed from registers: a quad of lanes holds one 8x4
//         tile, lane t its row t (so no data moves between lanes before the MMA); what is left for the ALU is
//         |c| of the four accumulators per MMA and the sum over the quad
//
// Build / run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/satd_mma_bench
//                         tools/satd_mma_bench.cu && tools/_build/satd_mma_bench [blocks] [reps]   -> one JSON object
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK( x ) do { cudaError_t e_ = ( x ); if( e_ != cudaSuccess ) { fprintf( stderr, "%s: %s\n", #x, cudaGetErrorString( e_ ) ); exit( 1 ); } } while( 0 )

// ---------------------------------------------------------------- scalar reference
static int satd4x4_cpu( const uint8_t *a, const uint8_t *b, int stride )
{
    int d[4][4], t[4][4], s = 0;
    for( int r = 0; r < 4; r++ )
        for( int c = 0; c < 4; c++ )
            d[r][c] = a[r * stride + c] - b[r * stride + c];
    for( int r = 0; r < 4; r++ )
    {
        const int s01 = d[r][0] + d[r][1], m01 = d[r][0] - d[r][1], s23 = d[r][2] + d[r][3], m23 = d[r][2] - d[r][3];
        t[r][0] = s01 + s23; t[r][1] = s01 - s23; t[r][2] = m01 + m23; t[r][3] = m01 - m23;
    }
    for( int c = 0; c < 4; c++ )
    {
        const int s01 = t[0][c] + t[1][c], m01 = t[0][c] - t[1][c], s23 = t[2][c] + t[3][c], m23 = t[2][c] - t[3][c];
        s += abs( s01 + s23 ) + abs( s01 - s23 ) + abs( m01 + m23 ) + abs( m01 - m23 );
    }
    return s;
}
static int satd8x8_cpu( const uint8_t *a, const uint8_t *b )
{
    int s = 0;
    for( int y = 0; y < 8; y += 4 )       // PIXEL_SATD_C( 8, 8, x264_pixel_satd_8x4 ): two 8x4 units, each halved
        s += ( satd4x4_cpu( a + y * 8, b + y * 8, 8 ) + satd4x4_cpu( a + y * 8 + 4, b + y * 8 + 4, 8 ) ) >> 1;
    return s;
}

// ---------------------------------------------------------------- device: ALU form
__device__ __forceinline__ int had_abs_alu( const uint32_t a[4], const uint32_t b[4] )
{
    int t[4][4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        const int d0 = (int)( a[r] & 255 ) - (int)( b[r] & 255 );
        const int d1 = (int)( ( a[r] >> 8 ) & 255 ) - (int)( ( b[r] >> 8 ) & 255 );
        const int d2 = (int)( ( a[r] >> 16 ) & 255 ) - (int)( ( b[r] >> 16 ) & 255 );
        const int d3 = (int)( a[r] >> 24 ) - (int)( b[r] >> 24 );
        const int s01 = d0 + d1, m01 = d0 - d1, s23 = d2 + d3, m23 = d2 - d3;
        t[r][0] = s01 + s23; t[r][1] = s01 - s23; t[r][2] = m01 + m23; t[r][3] = m01 - m23;
    }
    int acc = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        const int s01 = t[0][c] + t[1][c], m01 = t[0][c] - t[1][c], s23 = t[2][c] + t[3][c], m23 = t[2][c] - t[3][c];
        acc += abs( s01 + s23 ) + abs( s01 - s23 ) + abs( m01 + m23 ) + abs( m01 - m23 );
    }
    return acc;
}

// ---------------------------------------------------------------- device: IDP form (the product's)
__device__ __forceinline__ int dp4a_us( uint32_t a, uint32_t b, int c )
{
    int d;
    asm( "dp4a.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}
__device__ __forceinline__ int satd4x4_idp( const uint32_t a[4], const uint32_t b[4] )
{
    constexpr uint32_t P0 = 0x01010101u, P1 = 0xFFFF0101u, P2 = 0x01FFFF01u, P3 = 0xFF01FF01u;
    constexpr uint32_t N0 = 0xFFFFFFFFu, N1 = 0x0101FFFFu, N2 = 0xFF0101FFu, N3 = 0x01FF01FFu;
    int t[4][4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        t[r][0] = dp4a_us( a[r], P0, dp4a_us( b[r], N0, 0 ) );
        t[r][1] = dp4a_us( a[r], P1, dp4a_us( b[r], N1, 0 ) );
        t[r][2] = dp4a_us( a[r], P2, dp4a_us( b[r], N2, 0 ) );
        t[r][3] = dp4a_us( a[r], P3, dp4a_us( b[r], N3, 0 ) );
    }
    int acc = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        const int s01 = t[0][c] + t[1][c], m01 = t[0][c] - t[1][c], s23 = t[2][c] + t[3][c], m23 = t[2][c] - t[3][c];
        acc += max( abs( s01 ), abs( s23 ) ) + max( abs( m01 ), abs( m23 ) );
    }
    return acc;                                   // = sum |coef| / 2
}

// thread per 8x4 tile (two 4x4s side by side, four rows in registers), two threads per 8x8 block
template<bool IDP>
__global__ void __launch_bounds__( 256 ) satd_thread_kernel( const uint2 *__restrict__ fenc, const uint2 *__restrict__ ref, int n,
                                                             int *__restrict__ out )
{
    const int stride = gridDim.x * blockDim.x;
    for( int tile = blockIdx.x * blockDim.x + threadIdx.x; tile < 2 * n; tile += stride )
    {
        uint32_t fl[4], fr[4], pl[4], pr[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint2 f = __ldg( fenc + (size_t)tile * 4 + r ), p = __ldg( ref + (size_t)tile * 4 + r );
            fl[r] = f.x; fr[r] = f.y; pl[r] = p.x; pr[r] = p.y;
        }
        int s = IDP ? satd4x4_idp( fl, pl ) + satd4x4_idp( fr, pr ) : ( had_abs_alu( fl, pl ) + had_abs_alu( fr, pr ) ) >> 1;
        s += __shfl_xor_sync( 0xffffffffu, s, 1 );
        if( !( threadIdx.x & 1 ) )
            out[tile >> 1] = s;
    }
}

// ---------------------------------------------------------------- device: tensor-core form
__device__ __forceinline__ void mma_u8s8( int ( &c )[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1 )
{
    asm volatile( "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                  : "=r"( c[0] ), "=r"( c[1] ), "=r"( c[2] ), "=r"( c[3] )
                  : "r"( a0 ), "r"( a1 ), "r"( a2 ), "r"( a3 ), "r"( b0 ), "r"( b1 ), "r"( 0 ), "r"( 0 ), "r"( 0 ), "r"( 0 ) );
}

// quad of lanes per 8x4 tile, lane t = row t: MMA row g is the tile's left 4x4, row g+8 its right 4x4
__global__ void __launch_bounds__( 256 ) satd_mma_kernel( const uint2 *__restrict__ fenc, const uint2 *__restrict__ ref, int n,
                                                          int *__restrict__ out )
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // B = (H4 (x) H4) for the source half of K, its negation for the reference half; column n = 4u + v of slice s
    uint32_t bw[2];
#pragma unroll
    for( int s = 0; s < 2; s++ )
    {
        const int nn = 8 * s + g, u = nn >> 2, v = nn & 3;
        uint32_t w = 0;
#pragma unroll
        for( int c = 0; c < 4; c++ )
        {
            // h[i][j]: rows of H4 = {++++, ++--, +--+, +-+-}
            const int hu = ( u == 0 ) ? 1 : ( u == 1 ) ? ( t < 2 ? 1 : -1 ) : ( u == 2 ) ? ( ( t == 0 || t == 3 ) ? 1 : -1 ) : ( ( t & 1 ) ? -1 : 1 );
            const int hv = ( v == 0 ) ? 1 : ( v == 1 ) ? ( c < 2 ? 1 : -1 ) : ( v == 2 ) ? ( ( c == 0 || c == 3 ) ? 1 : -1 ) : ( ( c & 1 ) ? -1 : 1 );
            w |= (uint32_t)( ( hu * hv ) & 0xFF ) << ( 8 * c );
        }
        bw[s] = w;
    }
    const uint32_t nb0 = __vneg4( bw[0] ), nb1 = __vneg4( bw[1] );       // +-1 bytes: plain per-byte negation
    const int warps = gridDim.x * ( blockDim.x >> 5 );
    const int tiles = 2 * n;
    for( int base = ( blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 ) ) * 8; base < tiles; base += warps * 8 )
    {
        const int tile = base + g;                   // tiles is a multiple of 8 (n is a multiple of 4)
        const uint2 f = __ldg( fenc + (size_t)tile * 4 + t ), p = __ldg( ref + (size_t)tile * 4 + t );
        int c0[4], c1[4];
        mma_u8s8( c0, f.x, f.y, p.x, p.y, bw[0], nb0 );
        mma_u8s8( c1, f.x, f.y, p.x, p.y, bw[1], nb1 );
        int s = abs( c0[0] ) + abs( c0[1] ) + abs( c1[0] ) + abs( c1[1] );          // left 4x4
        int r = abs( c0[2] ) + abs( c0[3] ) + abs( c1[2] ) + abs( c1[3] );          // right 4x4
        s += r;
        s += __shfl_xor_sync( 0xffffffffu, s, 1 );
        s += __shfl_xor_sync( 0xffffffffu, s, 2 );
        s >>= 1;                                     // the 8x4 unit of pixel.c:294-314
        s += __shfl_xor_sync( 0xffffffffu, s, 4 );   // the block's other tile
        if( !( lane & 7 ) )
            out[tile >> 1] = s;
    }
}

int main( int argc, char **argv )
{
    const int n = argc > 1 ? atoi( argv[1] ) & ~3 : 1 << 21;
    const int reps = argc > 2 ? atoi( argv[2] ) : 20;
    const size_t bytes = (size_t)n * 64;
    uint8_t *h_f = (uint8_t *)malloc( bytes ), *h_p = (uint8_t *)malloc( bytes );
    uint32_t x = 0x9E3779B9u;
    for( size_t i = 0; i < bytes; i++ )
    {
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        h_f[i] = (uint8_t)( x >> 8 );
        // every fourth block adversarial (0 / 255 extremes), the rest a noisy copy of the source
        h_p[i] = ( ( i >> 6 ) & 3 ) == 3 ? ( ( x >> 20 ) & 1 ? 255 : 0 ) : (uint8_t)( h_f[i] + ( ( x >> 16 ) % 23 ) - 11 );
        if( ( ( i >> 6 ) & 3 ) == 3 )
            h_f[i] = ( x >> 21 ) & 1 ? 0 : 255;
    }
    uint2 *d_f, *d_p;
    int *d_o;
    CK( cudaMalloc( &d_f, bytes ) ); CK( cudaMalloc( &d_p, bytes ) ); CK( cudaMalloc( &d_o, (size_t)n * 4 ) );
    CK( cudaMemcpy( d_f, h_f, bytes, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_p, h_p, bytes, cudaMemcpyHostToDevice ) );
    int *h_o = (int *)malloc( (size_t)n * 4 );
    const int check = n < 200000 ? n : 200000;
    int *want = (int *)malloc( (size_t)check * 4 );
    for( int i = 0; i < check; i++ )
        want[i] = satd8x8_cpu( h_f + (size_t)i * 64, h_p + (size_t)i * 64 );

    int sms = 0;
    CK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, 0 ) );
    const int grid = sms * 8;
    cudaEvent_t e0, e1;
    CK( cudaEventCreate( &e0 ) ); CK( cudaEventCreate( &e1 ) );
    const char *names[3] = { "alu", "idp", "mma" };
    double ms[3];
    int ok[3];
    for( int k = 0; k < 3; k++ )
    {
        CK( cudaMemset( d_o, 0xff, (size_t)n * 4 ) );
        for( int r = -3; r < reps; r++ )
        {
            if( r == 0 )
                CK( cudaEventRecord( e0 ) );
            if( k == 0 ) satd_thread_kernel<false><<<grid, 256>>>( d_f, d_p, n, d_o );
            else if( k == 1 ) satd_thread_kernel<true><<<grid, 256>>>( d_f, d_p, n, d_o );
            else satd_mma_kernel<<<grid, 256>>>( d_f, d_p, n, d_o );
        }
        CK( cudaEventRecord( e1 ) );
        CK( cudaDeviceSynchronize() );
        float t;
        CK( cudaEventElapsedTime( &t, e0, e1 ) );
        ms[k] = t / reps;
        CK( cudaMemcpy( h_o, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost ) );
        ok[k] = 1;
        for( int i = 0; i < check; i++ )
            if( h_o[i] != want[i] )
            {
                fprintf( stderr, "%s: block %d: %d, want %d\n", names[k], i, h_o[i], want[i] );
                ok[k] = 0;
                break;
            }
    }
    printf( "{\"blocks\": %d, \"bytes_read\": %zu, \"reps\": %d, \"checked_blocks\": %d", n, 2 * bytes, reps, check );
    for( int k = 0; k < 3; k++ )
        printf( ", \"%s\": {\"ms\": %.5f, \"gpix_cmp_per_s\": %.1f, \"read_GBs\": %.1f, \"bit_exact\": %s}", names[k], ms[k],
                n * 64.0 / ( ms[k] * 1e6 ), 2.0 * bytes / ( ms[k] * 1e6 ), ok[k] ? "true" : "false" );
    printf( "}\n" );
    return ok[0] && ok[1] && ok[2] ? 0 : 1;
}
