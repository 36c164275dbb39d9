// frame.cu -- whole-frame streaming kernels (HBM-bound): picture staging, border expansion,
// half-resolution planes, half-pel planes.  sm_100a.
//
// Semantics (byte-exact, padding included):
//   x264_frame_copy_picture + x264_frame_expand_border_mod16   common/frame.c:198-232, 423-450
//   plane_expand_border / x264_frame_expand_border             common/frame.c:363-396
//   x264_frame_init_lowres + frame_init_lowres_core            common/mc.c:404-456
//   x264_frame_expand_border_lowres                            common/frame.c:415-421
//   hpel_filter + x264_frame_filter                            common/mc.c:144-167, 506-535
//   x264_frame_expand_border_filtered                          common/frame.c:398-413
//
// Every kernel is batched over frame slots (blockIdx.z) so that one launch covers a whole clip.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// I420 -> padded luma plane N + NV12 chroma plane.  One thread = 16 luma bytes or 8 UV pairs.
__global__ void __launch_bounds__( 256 )
xd_load_i420_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ i420, uint8_t *__restrict__ slots )
{
    const int frame = blockIdx.z;
    const size_t pic_bytes = (size_t)g.width * g.height * 3 / 2;
    const uint8_t *sy = i420 + frame * pic_bytes;
    const uint8_t *su = sy + (size_t)g.width * g.height;
    const uint8_t *sv = su + (size_t)( g.width >> 1 ) * ( g.height >> 1 );
    uint8_t *slot = slots + frame * (size_t)g.slot_bytes;
    const int units = g.luma_w >> 4;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;                      // [0, luma_h) luma rows, then chroma rows
    if( t >= units )
        return;
    if( row < g.luma_h )
    {
        const int sr = min( row, g.height - 1 );
        const uint8_t *src = sy + (size_t)sr * g.width;
        uint8_t *dst = slot + g.luma_origin + (size_t)row * g.luma_stride + t * 16;
        uint32_t w[4];
        const int x0 = t * 16;
        if( x0 + 16 <= g.width && ( ( (uintptr_t)( src + x0 ) ) & 15 ) == 0 )
        {
            uint4 v = __ldg( (const uint4 *)( src + x0 ) );
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
        else
        {
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                uint32_t acc = 0;
#pragma unroll
                for( int b = 0; b < 4; b++ )
                    acc |= (uint32_t)__ldg( src + min( x0 + 4 * k + b, g.width - 1 ) ) << ( 8 * b );
                w[k] = acc;
            }
        }
        *(uint4 *)dst = make_uint4( w[0], w[1], w[2], w[3] );
    }
    else
    {
        const int crow = row - g.luma_h;
        const int cw = g.width >> 1, ch = g.height >> 1;
        const int sr = min( crow, ch - 1 );
        const uint8_t *pu = su + (size_t)sr * cw, *pv = sv + (size_t)sr * cw;
        uint8_t *dst = slot + g.slot_chroma_off + g.chroma_origin + (size_t)crow * g.chroma_stride + t * 16;
        uint32_t w[4];
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const int x = min( t * 8 + 2 * k, cw - 1 ), x1 = min( t * 8 + 2 * k + 1, cw - 1 );
            w[k] = (uint32_t)__ldg( pu + x ) | ( (uint32_t)__ldg( pv + x ) << 8 )
                 | ( (uint32_t)__ldg( pu + x1 ) << 16 ) | ( (uint32_t)__ldg( pv + x1 ) << 24 );
        }
        *(uint4 *)dst = make_uint4( w[0], w[1], w[2], w[3] );
    }
}

// luma-only variant for the lookahead path: pictures are width*height bytes each
__global__ void __launch_bounds__( 256 )
xd_load_luma_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ luma, uint8_t *__restrict__ slots )
{
    const int frame = blockIdx.z;
    const uint8_t *sy = luma + frame * (size_t)g.width * g.height;
    uint8_t *slot = slots + frame * (size_t)g.slot_bytes;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if( t >= ( g.luma_w >> 4 ) )
        return;
    const int sr = min( row, g.height - 1 );
    const uint8_t *src = sy + (size_t)sr * g.width;
    uint8_t *dst = slot + g.luma_origin + (size_t)row * g.luma_stride + t * 16;
    const int x0 = t * 16;
    uint32_t w[4];
    if( x0 + 16 <= g.width && ( ( (uintptr_t)( src + x0 ) ) & 15 ) == 0 )
    {
        uint4 v = __ldg( (const uint4 *)( src + x0 ) );
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    else
    {
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t acc = 0;
#pragma unroll
            for( int b = 0; b < 4; b++ )
                acc |= (uint32_t)__ldg( src + min( x0 + 4 * k + b, g.width - 1 ) ) << ( 8 * b );
            w[k] = acc;
        }
    }
    *(uint4 *)dst = make_uint4( w[0], w[1], w[2], w[3] );
}

// ---------------------------------------------------------------------------------------------
// Generic replicate-padding of a plane: final state of plane_expand_border applied to the whole
// plane with top and bottom padding.  `unit` = 1 (luma) or 2 (interleaved UV pairs).
// Work items: the two 32-byte side bands of every picture row, and every 16-byte chunk of the
// rows above / below.  Values are derived from picture samples only, never from padding another
// thread may be writing.
struct xd_border_job
{
    int64_t plane_off;          // offset of sample (0,0) inside a slot
    int32_t stride, w, h, padh, padv, unit;
};

__device__ __forceinline__ uint2 xd_border_chunk( const uint8_t *row, int x0, int w, int unit )
{
    // 8 bytes starting at column x0 (multiple of 8) of a row whose picture part is [0,w)
    if( x0 >= 0 && x0 + 8 <= w )
        return *(const uint2 *)( row + x0 );
    uint32_t a, b;
    if( x0 < 0 )
    {
        a = row[0];
        b = unit == 2 ? row[1] : a;
    }
    else
    {
        a = row[w - unit];
        b = row[w - 1];
    }
    const uint32_t v = ( a | ( b << 8 ) ) * 0x00010001u;
    return make_uint2( v, v );
}

// 8-byte granularity: lowres widths are multiples of 8, not of 16
__global__ void __launch_bounds__( 256 )
xd_expand_border_kernel( xd_border_job job, uint8_t *__restrict__ slots, int64_t slot_bytes, int n_planes,
                         int64_t plane_pitch )
{
    uint8_t *base = slots + ( blockIdx.z / n_planes ) * slot_bytes + ( blockIdx.z % n_planes ) * plane_pitch
                  + job.plane_off;
    const int side_chunks = job.padh >> 3;                          // per side
    const int row_chunks = ( job.w + 2 * job.padh ) >> 3;
    const int n_side = job.h * 2 * side_chunks;
    const int n_total = n_side + 2 * job.padv * row_chunks;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n_total )
        return;
    int y, x0;
    if( i < n_side )
    {
        y = i / ( 2 * side_chunks );
        const int c = i % ( 2 * side_chunks );
        x0 = c < side_chunks ? -job.padh + 8 * c : job.w + 8 * ( c - side_chunks );
    }
    else
    {
        const int j = i - n_side;
        const int r = j / row_chunks;
        y = r < job.padv ? r - job.padv : job.h + ( r - job.padv );
        x0 = -job.padh + 8 * ( j % row_chunks );
    }
    const int ys = min( max( y, 0 ), job.h - 1 );
    const uint2 v = xd_border_chunk( base + (int64_t)ys * job.stride, x0, job.w, job.unit );
    *(uint2 *)( base + (int64_t)y * job.stride + x0 ) = v;
}

// ---------------------------------------------------------------------------------------------
// Half-resolution planes.  One thread = 8 output samples of all four planes on one output row:
// three source rows of 16(+1) bytes in, 4 x 8 bytes out.  Reads are clamped to the picture, which
// is what the reference obtains by first duplicating the last column and row into the source
// plane (mc.c:412-415); that side effect on the source plane is reproduced too.
__device__ __forceinline__ void xd_lowres_line( const uint4 r, uint32_t e, uint2 &o0, uint2 &oh )
{
    // r: 16 vertically-averaged bytes, e: the 17th.  o0 = avg(even, odd), oh = avg(odd, next even)
    const uint32_t ev_lo = __byte_perm( r.x, r.y, 0x6420 ), ev_hi = __byte_perm( r.z, r.w, 0x6420 );
    const uint32_t od_lo = __byte_perm( r.x, r.y, 0x7531 ), od_hi = __byte_perm( r.z, r.w, 0x7531 );
    const uint32_t nx_lo = __byte_perm( ev_lo, ev_hi, 0x4321 ), nx_hi = __byte_perm( ev_hi, e, 0x4321 );
    o0 = make_uint2( xd_avg4( ev_lo, od_lo ), xd_avg4( ev_hi, od_hi ) );
    oh = make_uint2( xd_avg4( od_lo, nx_lo ), xd_avg4( od_hi, nx_hi ) );
}

__global__ void __launch_bounds__( 128 )
xd_lowres_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots )
{
    uint8_t *slot = slots + blockIdx.z * (size_t)g.slot_bytes;
    uint8_t *src = slot + g.luma_origin;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if( t >= ( g.lowres_w >> 3 ) )
        return;
    const int ls = g.luma_stride;
    const int r0 = 2 * y, r1 = 2 * y + 1, r2 = min( 2 * y + 2, g.luma_h - 1 );
    const int xe = min( 16 * t + 16, g.luma_w - 1 );
    const uint4 a = *(const uint4 *)( src + (size_t)r0 * ls + 16 * t );
    const uint4 b = *(const uint4 *)( src + (size_t)r1 * ls + 16 * t );
    const uint4 c = *(const uint4 *)( src + (size_t)r2 * ls + 16 * t );
    const uint32_t ea = src[(size_t)r0 * ls + xe], eb = src[(size_t)r1 * ls + xe], ec = src[(size_t)r2 * ls + xe];

    const uint4 ab = make_uint4( xd_avg4( a.x, b.x ), xd_avg4( a.y, b.y ), xd_avg4( a.z, b.z ), xd_avg4( a.w, b.w ) );
    const uint4 bc = make_uint4( xd_avg4( b.x, c.x ), xd_avg4( b.y, c.y ), xd_avg4( b.z, c.z ), xd_avg4( b.w, c.w ) );
    const uint32_t eab = ( ea + eb + 1 ) >> 1, ebc = ( eb + ec + 1 ) >> 1;

    uint2 o0, oh, ov, oc;
    xd_lowres_line( ab, eab, o0, oh );
    xd_lowres_line( bc, ebc, ov, oc );

    uint8_t *dst = slot + g.slot_lowres_off + g.lowres_origin + (size_t)y * g.lowres_stride + 8 * t;
    *(uint2 *)( dst ) = o0;
    *(uint2 *)( dst + (size_t)g.lowres_plane_size ) = oh;
    *(uint2 *)( dst + 2 * (size_t)g.lowres_plane_size ) = ov;
    *(uint2 *)( dst + 3 * (size_t)g.lowres_plane_size ) = oc;

    // side effect of x264_frame_init_lowres on the source plane: column luma_w of every row and
    // row luma_h (luma_w + 1 bytes) duplicate their neighbours
    if( 16 * t + 16 == g.luma_w )
    {
        src[(size_t)r0 * ls + g.luma_w] = (uint8_t)ea;
        src[(size_t)r1 * ls + g.luma_w] = (uint8_t)eb;
    }
    if( y == g.lowres_h - 1 )
    {
        // b is the last picture row here (r1 == luma_h - 1)
        *(uint4 *)( src + (size_t)g.luma_h * ls + 16 * t ) = b;
        if( 16 * t + 16 == g.luma_w )
            src[(size_t)g.luma_h * ls + g.luma_w] = (uint8_t)eb;
    }
}

// ---------------------------------------------------------------------------------------------
// Half-pel planes.  Tile = 64 x 16 output samples per CTA of 256 threads:
//   1. stage the (64+8) x (16+5) source window in shared memory,
//   2. vertical six-tap -> 16-bit intermediate for 69 columns x 16 rows (shared memory),
//   3. every thread emits 4 samples of H, V and HV on one row.
// Computed region: rows [-8, luma_h+8), columns [0, luma_w+8) -- the rest of the padded plane is
// filled by xd_filtered_border_kernel from these values (the reference overwrites columns < 0).
#define HP_TW 64
#define HP_TH 16
#define HP_SW ( HP_TW + 8 )      // staged source columns: x0-4 .. x0+67
#define HP_SH ( HP_TH + 5 )      // staged source rows:    y0-2 .. y0+18

__device__ __forceinline__ int xd_tap6( int a, int b, int c, int d, int e, int f )
{
    return a + f - 5 * ( b + e ) + 20 * ( c + d );
}

__global__ void __launch_bounds__( 256 )
xd_hpel_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots )
{
    __shared__ __align__( 16 ) uint8_t s_src[HP_SH][HP_SW];
    __shared__ __align__( 16 ) int16_t s_mid[HP_TH][HP_SW];

    uint8_t *slot = slots + blockIdx.z * (size_t)g.slot_bytes;
    const uint8_t *pn = slot + g.luma_origin;
    const int ls = g.luma_stride;
    const int x0 = blockIdx.x * HP_TW;
    const int y0 = blockIdx.y * HP_TH - 8;
    const int tid = threadIdx.x;

    // 1. source window, 4 bytes per access (x0-4 is 4-byte aligned)
    for( int i = tid; i < HP_SH * ( HP_SW / 4 ); i += 256 )
    {
        const int r = i / ( HP_SW / 4 ), c = i % ( HP_SW / 4 );
        *(uint32_t *)&s_src[r][4 * c] = *(const uint32_t *)( pn + (int64_t)( y0 - 2 + r ) * ls + x0 - 4 + 4 * c );
    }
    __syncthreads();

    // 2. vertical filter for staged columns 2..70 (x0-2 .. x0+66)
    for( int i = tid; i < HP_TH * HP_SW; i += 256 )
    {
        const int r = i / HP_SW, c = i % HP_SW;
        s_mid[r][c] = (int16_t)xd_tap6( s_src[r][c], s_src[r + 1][c], s_src[r + 2][c],
                                       s_src[r + 3][c], s_src[r + 4][c], s_src[r + 5][c] );
    }
    __syncthreads();

    // 3. outputs: thread -> row tid/16, columns 4*(tid%16) .. +3
    const int r = tid >> 4, cx = ( tid & 15 ) * 4;
    const int y = y0 + r, x = x0 + cx;
    if( y >= g.luma_h + 8 || x >= g.luma_w + 8 )
        return;
    uint32_t oh = 0, ov = 0, oc = 0;
#pragma unroll
    for( int k = 0; k < 4; k++ )
    {
        const int c = cx + k + 4;                    // staged column of output sample
        const uint8_t *s = &s_src[r + 2][c];
        const int16_t *m = &s_mid[r][c];
        const int hv = xd_tap6( s[-2], s[-1], s[0], s[1], s[2], s[3] );
        const int cv = xd_tap6( m[-2], m[-1], m[0], m[1], m[2], m[3] );
        oh |= (uint32_t)xd_clip_u8( ( hv + 16 ) >> 5 ) << ( 8 * k );
        ov |= (uint32_t)xd_clip_u8( ( m[0] + 16 ) >> 5 ) << ( 8 * k );
        oc |= (uint32_t)xd_clip_u8( ( cv + 512 ) >> 10 ) << ( 8 * k );
    }
    const int64_t o = (int64_t)y * ls + x;
    *(uint32_t *)( slot + (size_t)g.luma_plane_size + g.luma_origin + o ) = oh;
    *(uint32_t *)( slot + 2 * (size_t)g.luma_plane_size + g.luma_origin + o ) = ov;
    *(uint32_t *)( slot + 3 * (size_t)g.luma_plane_size + g.luma_origin + o ) = oc;
}

// Padding of the three filtered planes = the final state of x264_frame_expand_border_filtered
// called per MB row.  With F(x,y) the filtered value (rows [-8,h+8), columns [0,w+8)):
//   value(x,y) = F( x<0 ? 0 : min(x,w+7), clamp(y,-8,h+7) )        for x in [-32, w+40)
// The reference writes w+72 bytes per row; when the stride is w+64 the 8 extra bytes of a row land
// on the first 8 bytes of the next row and the later writer wins:
//   - rows -8 .. h+31: the row's own left band wins (nothing special),
//   - rows -32 .. -8 : columns [-32,-24) end up holding F(0,-7) (row copies made upwards from row -8
//                      carry row -7's left band along),
//   - the 8 bytes after the last row of a plane receive F(0,h+7).
// When the stride is larger the extra bytes fall into the unused gap and simply keep F(w+7,.).
__global__ void __launch_bounds__( 256 )
xd_filtered_border_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots )
{
    const int plane = 1 + blockIdx.z % 3;
    uint8_t *slot = slots + ( blockIdx.z / 3 ) * (size_t)g.slot_bytes;
    uint8_t *base = slot + (size_t)plane * g.luma_plane_size + g.luma_origin;
    const int ls = g.luma_stride, w = g.luma_w, h = g.luma_h;
    const bool tight = ls == w + 64;                 // row tails alias the next row's head
    const int body = h + 16;                         // rows -8 .. h+7
    const int n_side = body * 5;                     // 2 left chunks + 3 right chunks (w+8 .. w+40 = 32 + 8)
    const int row_chunks = ( w + 64 ) >> 4;          // full rows above / below: [-32, w+32)
    const int n_total = n_side + 48 * ( row_chunks + 1 );
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n_total )
        return;
    int y, x0, len = 16;
    if( i < n_side )
    {
        y = i / 5 - 8;
        const int c = i % 5;
        x0 = c < 2 ? -32 + 16 * c : w + 8 + 16 * ( c - 2 );
        if( c == 2 ) { x0 = w + 8; len = 8; }        // [w+8, w+16)
        else if( c >= 3 ) x0 = w + 16 * ( c - 2 );   // [w+16,w+32), [w+32,w+48) -> second clipped below
        if( c == 4 ) len = 8;                        // [w+32, w+40): the row tail
    }
    else
    {
        const int j = i - n_side;
        const int r = j / ( row_chunks + 1 );
        y = r < 24 ? r - 32 : h + 8 + ( r - 24 );
        const int c = j % ( row_chunks + 1 );
        x0 = -32 + 16 * c;
        if( c == row_chunks ) len = 8;               // tail [w+32, w+40)
    }
    const int ys = min( max( y, -8 ), h + 7 );
    const uint8_t *frow = base + (int64_t)ys * ls;
    uint8_t *drow = base + (int64_t)y * ls;
    const bool is_tail = x0 == w + 32;
    if( is_tail )
    {
        if( tight )
        {
            if( y != h + 31 )
                return;                              // overwritten by the next row's head
            const uint32_t v = base[(int64_t)( h + 7 ) * ls] * 0x01010101u;
            *(uint2 *)( drow + x0 ) = make_uint2( v, v );
            return;
        }
        const uint32_t v = frow[w + 7] * 0x01010101u;
        *(uint2 *)( drow + x0 ) = make_uint2( v, v );
        return;
    }
    if( x0 < 0 )
    {
        uint32_t v = frow[0] * 0x01010101u;
        uint32_t v0 = v;
        if( tight && y <= -8 && x0 == -32 )
            v0 = base[(int64_t)( -7 ) * ls] * 0x01010101u;
        // no alignment gap between the planes: the previous plane's last row tail (written last by
        // the reference) lands on this plane's very first 8 bytes
        if( tight && plane >= 2 && y == -32 && x0 == -32 && g.luma_plane_size == ls * ( h + 64 ) )
            v0 = ( base - g.luma_plane_size )[(int64_t)( h + 7 ) * ls] * 0x01010101u;
        *(uint4 *)( drow + x0 ) = make_uint4( v0, v0, v, v );
    }
    else if( x0 >= w + 8 )
    {
        const uint32_t v = frow[w + 7] * 0x01010101u;
        if( len == 8 )
            *(uint2 *)( drow + x0 ) = make_uint2( v, v );
        else
            *(uint4 *)( drow + x0 ) = make_uint4( v, v, v, v );
    }
    else if( x0 + 16 <= w + 8 )
        *(uint4 *)( drow + x0 ) = *(const uint4 *)( frow + x0 );         // rows above / below only
    else
    {
        // chunk [w, w+16) of a row above / below: 8 filtered bytes then 8 replicated
        const uint2 f = *(const uint2 *)( frow + x0 );
        const uint32_t v = frow[w + 7] * 0x01010101u;
        *(uint4 *)( drow + x0 ) = make_uint4( f.x, f.y, v, v );
    }
}

// ---------------------------------------------------------------------------------------------
// host entry points

extern "C" int x264dsp_frame_load_i420_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *i420,
                                             uint8_t *slots, int n_frames, void *stream )
{
    if( !ctx || !g || !i420 || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    dim3 grid( ( ( g->luma_w >> 4 ) + 255 ) / 256, g->luma_h + g->chroma_h, n_frames );
    cudaStream_t s = xd_stream( ctx, stream );
    const int pslot = xd_prof_begin( ctx, XD_PROF_LOAD, s );
    xd_load_i420_kernel<<<grid, 256, 0, s>>>( *g, i420, slots );
    xd_prof_end( ctx, XD_PROF_LOAD, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

int xd_frame_load_luma( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma, uint8_t *slots,
                        int n_frames, cudaStream_t s )
{
    dim3 grid( ( ( g->luma_w >> 4 ) + 255 ) / 256, g->luma_h, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_LOAD, s );
    xd_load_luma_kernel<<<grid, 256, 0, s>>>( *g, luma, slots );
    xd_prof_end( ctx, XD_PROF_LOAD, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_frame_load_luma_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                             uint8_t *slots, int n_frames, void *stream )
{
    if( !ctx || !g || !luma || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    return xd_frame_load_luma( ctx, g, luma, slots, n_frames, xd_stream( ctx, stream ) );
}

static int xd_launch_border( x264dsp_ctx_t *ctx, const xd_border_job &job, uint8_t *slots, int64_t slot_bytes,
                             int n_frames, int n_planes, int64_t plane_pitch, cudaStream_t s )
{
    const int side_chunks = job.padh >> 3, row_chunks = ( job.w + 2 * job.padh ) >> 3;
    const int n_total = job.h * 2 * side_chunks + 2 * job.padv * row_chunks;
    dim3 grid( ( n_total + 255 ) / 256, 1, n_frames * n_planes );
    const int pslot = xd_prof_begin( ctx, XD_PROF_BORDER, s );
    xd_expand_border_kernel<<<grid, 256, 0, s>>>( job, slots, slot_bytes, n_planes, plane_pitch );
    xd_prof_end( ctx, XD_PROF_BORDER, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_frame_expand_border_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                                 int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    xd_border_job luma = { g->luma_origin, g->luma_stride, g->luma_w, g->luma_h, X264DSP_PADH, X264DSP_PADV, 1 };
    xd_border_job chroma = { (int64_t)g->slot_chroma_off + g->chroma_origin, g->chroma_stride, g->luma_w,
                             g->chroma_h, X264DSP_PADH, X264DSP_PADV / 2, 2 };
    int rc = xd_launch_border( ctx, luma, slots, g->slot_bytes, n_frames, 1, 0, s );
    if( rc )
        return rc;
    return xd_launch_border( ctx, chroma, slots, g->slot_bytes, n_frames, 1, 0, s );
}

extern "C" int x264dsp_frame_init_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                               int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    dim3 grid( ( ( g->lowres_w >> 3 ) + 127 ) / 128, g->lowres_h, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_LOWRES, s );
    xd_lowres_kernel<<<grid, 128, 0, s>>>( *g, slots );
    xd_prof_end( ctx, XD_PROF_LOWRES, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    xd_border_job lowres = { (int64_t)g->slot_lowres_off + g->lowres_origin, g->lowres_stride, g->lowres_w,
                             g->lowres_h, X264DSP_PADH, X264DSP_PADV, 1 };
    return xd_launch_border( ctx, lowres, slots, g->slot_bytes, n_frames, 4, g->lowres_plane_size, s );
}

extern "C" int x264dsp_frame_filter_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                          int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    dim3 grid( ( g->luma_w + 8 + HP_TW - 1 ) / HP_TW, ( g->luma_h + 16 + HP_TH - 1 ) / HP_TH, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_HPEL, s );
    xd_hpel_kernel<<<grid, 256, 0, s>>>( *g, slots );
    xd_prof_end( ctx, XD_PROF_HPEL, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    const int n_total = ( g->luma_h + 16 ) * 5 + 48 * ( ( ( g->luma_w + 64 ) >> 4 ) + 1 );
    dim3 bgrid( ( n_total + 255 ) / 256, 1, 3 * n_frames );
    xd_filtered_border_kernel<<<bgrid, 256, 0, s>>>( *g, slots );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
