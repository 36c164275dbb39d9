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
#include <cuda.h>
#include <cstdlib>
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

// 16 source bytes of luma row `row` starting at column x0 (a multiple of 16), plus the byte after them.
// PLANE: from the slot's padded luma plane (reads clamped to the picture, see above).
// RAW:   straight from the caller's width x height picture; the mod-16 padding (frame.c:435-448) and the
//        duplicated column / row are the same clamp.
template<bool RAW>
__device__ __forceinline__ void xd_lowres_src( const uint8_t *base, int pitch, int w, int h, int row, int x0,
                                               uint4 &v, uint32_t &e )
{
    const uint8_t *p = base + (size_t)min( row, h - 1 ) * pitch;
    if( !RAW || ( x0 + 16 <= w && ( ( (uintptr_t)( p + x0 ) ) & 15 ) == 0 ) )
        v = *(const uint4 *)( p + x0 );
    else
    {
        uint32_t q[4];
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t acc = 0;
#pragma unroll
            for( int b = 0; b < 4; b++ )
                acc |= (uint32_t)p[min( x0 + 4 * k + b, w - 1 )] << ( 8 * b );
            q[k] = acc;
        }
        v = make_uint4( q[0], q[1], q[2], q[3] );
    }
    e = p[min( x0 + 16, w - 1 )];
}

// LR_ROWS consecutive rows of one tile (LR_ROWS * 8 contiguous bytes: one 32-byte sector for 4 rows)
#define LR_ROWS 4
__device__ __forceinline__ void xd_tile_store( uint8_t *tile_rows, const uint2 rows[LR_ROWS] )
{
    uint4 *d = (uint4 *)tile_rows;
#pragma unroll
    for( int i = 0; i < LR_ROWS / 2; i++ )
        d[i] = make_uint4( rows[2 * i].x, rows[2 * i].y, rows[2 * i + 1].x, rows[2 * i + 1].y );
}

// The four half-resolution planes are kept in the slot in TILED form only (the layout the lookahead
// searches in; x264dsp_frame_export_lowres_dev produces the reference's row-major planes on request):
// writing both forms costs 2.5 MB more per 1080p frame, and this kernel is bound by HBM writes.
// Thread = one 8-sample column chunk of LR_ROWS lowres rows, i.e. half an 8x8 tile (one 32-byte sector)
// of each of the four planes: the thread walks its 2*LR_ROWS+1 source rows once (16-byte loads) and
// stores the finished half tiles as whole sectors.  Edge threads also write the padding tiles
// (x264_frame_expand_border_lowres, frame.c:415-421: replicate 32 samples on every side).
// RAW = false: x264_frame_init_lowres on a slot whose luma plane is loaded.
// RAW = true : picture staging fused in -- the thread also writes the luma rows it has read into the
//              slot's plane N, which saves re-reading 2 MB per 1080p frame.
// KEEP = false (RAW only): the picture is read as the reference reads frame->plane[0] and NOTHING but the
//              lowres planes is written -- for callers that only want the lookahead of these pictures.
template<bool RAW, bool KEEP>
__global__ void __launch_bounds__( 128 )
xd_lowres_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots, const uint8_t *__restrict__ raw )
{
    uint8_t *slot = slots + blockIdx.z * (size_t)g.slot_bytes;
    uint8_t *plane = slot + g.luma_origin;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int hb = blockIdx.y;                       // group of LR_ROWS lowres rows
    const int chunks = g.lowres_w >> 3;
    if( t >= chunks )
        return;
    const int ls = g.luma_stride;
    const uint8_t *src = RAW ? raw + blockIdx.z * (size_t)g.width * g.height : plane;
    const int pitch = RAW ? g.width : ls, sw = RAW ? g.width : g.luma_w, sh = RAW ? g.height : g.luma_h;
    uint8_t *tl = slot + g.slot_tiled_off;
    const size_t tps = (size_t)g.tiled_plane_size;
    const bool first = t == 0, last = t == chunks - 1;
    const int y0 = LR_ROWS * hb;

    uint2 T[4][LR_ROWS];                             // this thread's rows of the four tiles
    uint4 a, b, c;
    uint32_t ea, eb, ec;
    xd_lowres_src<RAW>( src, pitch, sw, sh, 2 * y0, 16 * t, a, ea );
#pragma unroll
    for( int r = 0; r < LR_ROWS; r++ )
    {
        const int y = y0 + r;
        const int r0 = 2 * y, r1 = 2 * y + 1;
        xd_lowres_src<RAW>( src, pitch, sw, sh, r1, 16 * t, b, eb );
        xd_lowres_src<RAW>( src, pitch, sw, sh, r1 + 1, 16 * t, c, ec );
        const uint4 ab = make_uint4( xd_avg4( a.x, b.x ), xd_avg4( a.y, b.y ), xd_avg4( a.z, b.z ), xd_avg4( a.w, b.w ) );
        const uint4 bc = make_uint4( xd_avg4( b.x, c.x ), xd_avg4( b.y, c.y ), xd_avg4( b.z, c.z ), xd_avg4( b.w, c.w ) );
        const uint32_t eab = ( ea + eb + 1 ) >> 1, ebc = ( eb + ec + 1 ) >> 1;
        xd_lowres_line( ab, eab, T[0][r], T[1][r] );
        xd_lowres_line( bc, ebc, T[2][r], T[3][r] );

        if( RAW && KEEP )
        {
            *(uint4 *)( plane + (size_t)r0 * ls + 16 * t ) = a;
            *(uint4 *)( plane + (size_t)r1 * ls + 16 * t ) = b;
        }
        // side effect of x264_frame_init_lowres on the source plane: column luma_w of every row and
        // row luma_h (luma_w + 1 bytes) duplicate their neighbours
        if( KEEP && last )
        {
            plane[(size_t)r0 * ls + g.luma_w] = (uint8_t)ea;
            plane[(size_t)r1 * ls + g.luma_w] = (uint8_t)eb;
        }
        if( KEEP && y == g.lowres_h - 1 )
        {
            // b is the last picture row here (r1 == luma_h - 1)
            *(uint4 *)( plane + (size_t)g.luma_h * ls + 16 * t ) = b;
            if( last )
                plane[(size_t)g.luma_h * ls + g.luma_w] = (uint8_t)eb;
        }
        a = c;
        ea = ec;
    }

    // ---- the tiles: rows y0 .. y0+LR_ROWS-1 of tile (tx, ty) = (t + 4, y0/8 + 4) of each padded plane
    const int tx = t + X264DSP_PADH / 8, ty = ( y0 >> 3 ) + X264DSP_PADV / 8;
    const int roff = ( y0 & 7 ) * 8;                 // byte offset of the thread's rows inside a tile
    const bool top = y0 == 0, bottom = y0 + LR_ROWS == g.lowres_h;
#pragma unroll
    for( int k = 0; k < 4; k++ )
    {
        uint8_t *tp = tl + k * tps;
        xd_tile_store( tp + ( (size_t)ty * g.tile_w + tx ) * 64 + roff, T[k] );
        uint2 L[LR_ROWS], R[LR_ROWS];                // side padding: the first / last sample of every row
#pragma unroll
        for( int r = 0; r < LR_ROWS; r++ )
        {
            const uint32_t vl = ( T[k][r].x & 255u ) * 0x01010101u, vr = ( T[k][r].y >> 24 ) * 0x01010101u;
            L[r] = make_uint2( vl, vl );
            R[r] = make_uint2( vr, vr );
        }
        if( first )
            for( int cc = 1; cc <= X264DSP_PADH / 8; cc++ )
                xd_tile_store( tp + ( (size_t)ty * g.tile_w + tx - cc ) * 64 + roff, L );
        if( last )
            for( int cc = 1; cc <= X264DSP_PADH / 8; cc++ )
                xd_tile_store( tp + ( (size_t)ty * g.tile_w + tx + cc ) * 64 + roff, R );
        // top / bottom padding: four tile rows that repeat the picture's first / last row
#pragma unroll
        for( int side = 0; side < 2; side++ )
        {
            if( side == 0 ? !top : !bottom )
                continue;
            uint2 E[LR_ROWS], EL[LR_ROWS], ER[LR_ROWS];
#pragma unroll
            for( int r = 0; r < LR_ROWS; r++ )
            {
                E[r] = side == 0 ? T[k][0] : T[k][LR_ROWS - 1];
                EL[r] = side == 0 ? L[0] : L[LR_ROWS - 1];
                ER[r] = side == 0 ? R[0] : R[LR_ROWS - 1];
            }
            for( int rr = 1; rr <= X264DSP_PADV / 8; rr++ )
            {
                const int tyy = side == 0 ? ty - rr : ty + rr;
                for( int half = 0; half < 8 / LR_ROWS; half++ )
                {
                    const int ho = half * LR_ROWS * 8;
                    xd_tile_store( tp + ( (size_t)tyy * g.tile_w + tx ) * 64 + ho, E );
                    if( first )
                        for( int cc = 1; cc <= X264DSP_PADH / 8; cc++ )
                            xd_tile_store( tp + ( (size_t)tyy * g.tile_w + tx - cc ) * 64 + ho, EL );
                    if( last )
                        for( int cc = 1; cc <= X264DSP_PADH / 8; cc++ )
                            xd_tile_store( tp + ( (size_t)tyy * g.tile_w + tx + cc ) * 64 + ho, ER );
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Half-pel planes (hpel_filter, mc.c:144-167).  Computed region: rows [-8, luma_h+8), columns
// [0, luma_w+8) -- the rest of the padded plane is filled by xd_filtered_border_kernel from these
// values (the reference overwrites columns < 0).
//
// Mapping: a lane owns 8 pixels of a row (one "unit", one 8-byte load and three 8-byte stores per row) and walks
// HP_ROWS rows downwards; GW adjacent lanes form a strip whose first and last lane are halo (they hold the
// neighbouring units so that every horizontal neighbour is one shuffle away).  Full strips use GW = 32 (30 units =
// 240 pixels per warp); the units left over at the right edge (one for 1080p and 4K: 1928 = 8 x 240 + 8) are served
// by narrower strips (GW = 4, 8, 16) with 32/GW strips per warp stacked over different row segments, so that the
// remainder does not cost a whole warp column.
//
// Pipes (tools/int_pipe_peak.cu): PRMT / SHF / LOP3 / VIMNMX / I2IP issue on the ALU pipe, IMAD / VIADD.16x2 /
// IDP.4A / IDP.2A on the FMA pipe, each at 2 warp instructions per clock per SM, so the taps are written as integer
// dot products (FMA pipe) and only alignment, rounding and packing stay on the ALU pipe:
//   V   six rows of the lane's pixels as 16-bit pairs in registers; (a+f) - 5 (b+e) + 20 (c+d) two pixels per
//       instruction with each half biased by 2^15 + 16 (the sums fit 16 bits; the +16 is the rounding term)
//   H   IDP.4A over byte windows at even offsets: pixel 2k   = W[2k-2].(1,-5,20,20) + W[2k+2].(-5,1,0,0)
//                                                 pixel 2k+1 = W[2k-2].(0,1,-5,20) + W[2k+2].(20,-5,1,0)
//   HV  IDP.2A (u16 x s8) over the UNCLIPPED biased V pairs (the reference's int16 buf): the bias contributes
//       32 * 2^15 and the +16 contribute exactly the +512 rounding term, so the accumulator starts at -2^20
//   clip + pack: cvt.pack.sat.u8.s32 (I2IP), two pixels per instruction
#define HP_ROWS 48
#define HP_UNITS 30
#define HP_RING 6                       // ring slots per lane (the unroll of the row loop)
#define HP_AHEAD 5                      // source rows in flight per lane

// one 8-byte cp.async as its own group
__device__ __forceinline__ void xd_cp_async8( void *smem, const void *gmem )
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared( smem );
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;" :: "r"( s ), "l"( gmem ) : "memory" );
}
template<int N>
__device__ __forceinline__ void xd_cp_async_wait()
{
    asm volatile( "cp.async.wait_group %0;" :: "n"( N ) : "memory" );
}

// (a+f) - 5 (b+e) + 20 (c+d) + (32768 + 16) in both halves; inputs are 0..255 per half
__device__ __forceinline__ uint32_t xd_hp_tap6_packed( uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f )
{
    const uint32_t pos = ( c + d ) * 20u + ( a + f + 0x80108010u );
    return pos - ( b + e ) * 5u;
}

__device__ __forceinline__ int xd_dp4a_us( uint32_t a, uint32_t b, int c )
{
    int d;
    asm( "dp4a.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}
__device__ __forceinline__ int xd_dp2a_lo_us( uint32_t a, uint32_t b, int c )
{
    int d;
    asm( "dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}
__device__ __forceinline__ int xd_dp2a_hi_us( uint32_t a, uint32_t b, int c )
{
    int d;
    asm( "dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}

// ---- TMA variant of the row source (full strips only): one elected lane fetches HP_TROWS rows x 256 bytes -- the strip with
// its two halo units -- per cp.async.bulk.tensor into a three-stage ring of the warp, completion on an mbarrier per stage; a
// lane then reads its eight bytes of a row from the tile.  The tensor is (byte in row, row of plane N, slot), so rows past
// the plane come back as zeros (they only feed rows that are not stored).
// A box has to START on a 16-byte boundary of the innermost dimension (tools/tma_probe.cu: the same load at byte 24 of a row
// raises "illegal instruction", at 16 or 32 it lands) -- TMA does not fetch byte windows at arbitrary positions.  A strip's 32
// units (30 + two halo units) start at byte 24 + 240 k of a row, 8 bytes off.  So a stage takes TWO boxes: 256 bytes from the
// strip's first real unit (aligned) -- lanes 1 .. 31 -- and 16 bytes ending with the left halo unit -- lane 0.  (The first
// version kept one box and shrank the strips to 29 units: 1080p's 241 units then leave a tail of 9 units in half-warp strips
// instead of 1 unit in a quarter-warp strip, 24 % more instructions, 2.75 against 2.43 us per frame.)
#define HP_TROWS 6                      // = the unroll of the row loop: tile and row-in-tile are compile-time there
#define HP_TSTAGES 3
#define HP_TMAIN ( HP_TROWS * 256 )
#define HP_THALO ( HP_TROWS * 16 )
#define HP_TILE_BYTES ( HP_TMAIN + 128 ) // the halo box sits behind the main one, on a 128-byte boundary

struct xd_hp_tma
{
    const void *tmap, *tmap_halo;   // CUtensorMaps in kernel parameter space: boxes of 256 and of 16 bytes per row
    uint8_t *tiles;                 // this warp's HP_TSTAGES stages
    uint64_t *bars;                 // this warp's HP_TSTAGES mbarriers
    int x, y, z;                    // tensor coordinates of the segment's first source row (x: the strip's first real unit)
};

__device__ __forceinline__ void xd_hp_tma_issue( const xd_hp_tma &T, int tile )
{
    const int st = tile % HP_TSTAGES;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared( T.bars + st );
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared( T.tiles + st * HP_TILE_BYTES );
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( bar ), "r"( HP_TMAIN + HP_THALO ) : "memory" );
    asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  :: "r"( dst ), "l"( (uint64_t)T.tmap ), "r"( bar ), "r"( T.x ), "r"( T.y + tile * HP_TROWS ), "r"( T.z ) : "memory" );
    asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  :: "r"( dst + HP_TMAIN ), "l"( (uint64_t)T.tmap_halo ), "r"( bar ), "r"( T.x - 16 ), "r"( T.y + tile * HP_TROWS ), "r"( T.z )
                  : "memory" );
}
__device__ __forceinline__ void xd_hp_tma_wait( const xd_hp_tma &T, int stage, uint32_t parity )
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared( T.bars + stage );
    uint32_t ok = 0, spins = 0;
    while( !ok )
    {
        asm volatile( "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                      : "=r"( ok ) : "r"( bar ), "r"( parity ) : "memory" );
        if( ++spins > ( 1u << 26 ) )
            __trap();                                            // a tile that never lands is a bug, not something to wait for
    }
}

template<int GW, bool TMA = false>
__device__ __forceinline__ void xd_hpel_body( const x264dsp_geom_t &g, uint8_t *__restrict__ slot, int unit0, int strip_units, int wseg,
                                              uint2 *s_ring, xd_hp_tma T = xd_hp_tma() )
{
    const int lane = threadIdx.x & 31;
    const int gl = lane & ( GW - 1 );
    const int seg = wseg * ( 32 / GW ) + lane / GW;              // row segment of this lane's strip
    const int ls = g.luma_stride;
    const int y_end = g.luma_h + 8;
    const bool seg_ok = seg * HP_ROWS - 8 < y_end;
    if( !__any_sync( 0xffffffffu, seg_ok ) )
        return;
    // strips of a partly filled warp that fall below the plane redo the first segment without storing; a segment
    // reads at most 31 rows past the padded plane N, i.e. into plane H of the same slot
    const int y0 = seg_ok ? seg * HP_ROWS - 8 : -8;              // first output row
    const int n_units = ( g.luma_w + 8 ) >> 3;
    // halo lanes (and idle lanes past the strip) still read inside the allocation: units -1 .. n_units, pad is 32
    const int unit = min( unit0 + gl - 1, n_units );
    const bool writer = seg_ok && gl >= 1 && gl <= GW - 2 && gl - 1 < strip_units;
    const uint8_t *ps = slot + g.luma_origin + 8 * unit + (int64_t)( y0 - 2 ) * ls;      // next source row to load
    uint8_t *dh = slot + (size_t)g.luma_plane_size + g.luma_origin + 8 * unit + (int64_t)y0 * ls;
    uint8_t *dv = dh + (size_t)g.luma_plane_size, *dc = dv + (size_t)g.luma_plane_size;

    uint32_t A[6][4];                                            // rows y-2 .. y+3 as (p0,p1) (p2,p3) (p4,p5) (p6,p7)
    uint2 R[3];                                                  // rows y .. y+2 as loaded
    // TMA: tile 0 = source rows -1 .. 4 of the segment (row -1 is fetched only so that the tiles line up with the row loop:
    // its iteration i reads tile i + 1, row u of the tile in step u), tiles 1 .. 8 the 48 rows the loop consumes
    constexpr int n_tiles = 1 + HP_ROWS / 6;
    const uint8_t *trow = nullptr;                               // this lane's eight bytes in row 0 of stage 0
    int trs = 256;                                               // ... and the distance to its next row
    int st = 0;                                                  // stage / parity of the tile being read
    uint32_t ph = 0;
    if( TMA )
    {
        T.y += y0 - 3;
        trow = lane ? T.tiles + ( lane - 1 ) * 8 : T.tiles + HP_TMAIN + 8;         // lane 0: the left halo unit, in the 16-byte box
        trs = lane ? 256 : 16;
        if( lane == 0 )
        {
            for( int k = 0; k < HP_TSTAGES; k++ )
                asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"( (uint32_t)__cvta_generic_to_shared( T.bars + k ) ) : "memory" );
            asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
            asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
            for( int k = 0; k < HP_TSTAGES && k < n_tiles; k++ )
                xd_hp_tma_issue( T, k );
        }
        __syncwarp();
        xd_hp_tma_wait( T, 0, 0 );
    }
#pragma unroll
    for( int k = 0; k < 5; k++ )
    {
        uint2 w;
        if( TMA )
            w = *(const uint2 *)( trow + ( k + 1 ) * trs );
        else
            w = *(const uint2 *)ps;
        ps += ls;
        A[k][0] = __byte_perm( w.x, 0u, 0x4140 );
        A[k][1] = __byte_perm( w.x, 0u, 0x4342 );
        A[k][2] = __byte_perm( w.y, 0u, 0x4140 );
        A[k][3] = __byte_perm( w.y, 0u, 0x4342 );
        if( k >= 2 )
            R[k - 2] = w;
    }
    // HP_AHEAD source rows are in flight per lane through cp.async into a private ring in shared memory.  Plain
    // prefetch loads into registers did not get past two rows: the loads of a warp share scoreboards, waiting for the
    // oldest one waits for all of them (2, 3 and 6 rows ahead all measured 2.65 us per frame, 40 % of the stall samples
    // on the first use of the row); cp.async groups complete in order and are waited for by count.
    int tile = 1;
    if( TMA )
    {
        __syncwarp();                                            // every lane has read tile 0
        if( lane == 0 && HP_TSTAGES < n_tiles )
            xd_hp_tma_issue( T, HP_TSTAGES );
        st = 1;
    }
    uint2 *ring = s_ring + threadIdx.x;                          // slot k of this thread: ring[k * 128]
    if( !TMA )
    {
#pragma unroll
        for( int k = 0; k < HP_AHEAD; k++ )
        {
            xd_cp_async8( ring + k * 128, ps );
            ps += ls;
        }
    }
    const uint32_t TA0 = 0x1414FB01u, TB0 = 0x000001FBu;         // (1,-5,20,20) (-5,1,0,0)
    const uint32_t TA1 = 0x14FB0100u, TB1 = 0x0001FB14u;         // (0,1,-5,20) (20,-5,1,0)
    const uint32_t T1 = 0x1414FB01u, T2 = 0x010001FBu;           // lo (1,-5) hi (20,20); lo (-5,1) hi (0,1)
    const uint32_t T3 = 0xFB1414FBu, T4 = 0x00000001u;           // lo (-5,20) hi (20,-5); lo (1,0)
    for( int yb = y0; yb < y0 + HP_ROWS; yb += 6 )
    {
#pragma unroll
        for( int u = 0; u < 6; u++ )
        {
            const int y = yb + u;
            const int i0 = u % 6, i1 = ( u + 1 ) % 6, i2 = ( u + 2 ) % 6, i3 = ( u + 3 ) % 6, i4 = ( u + 4 ) % 6, i5 = ( u + 5 ) % 6;
            const uint2 cur = R[u % 3];
            {
                uint2 w;
                if( TMA )
                {
                    if( u == 0 )
                        xd_hp_tma_wait( T, st, ph );
                    w = *(const uint2 *)( trow + st * HP_TILE_BYTES + u * trs );
                    if( u == 5 )
                    {
                        __syncwarp();                            // every lane has read the tile's last row
                        if( lane == 0 && tile + HP_TSTAGES < n_tiles )
                            xd_hp_tma_issue( T, tile + HP_TSTAGES );
                        tile++;
                        if( ++st == HP_TSTAGES )
                        {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
                else
                {
                    xd_cp_async_wait<HP_AHEAD - 1>();             // the oldest row in flight has landed
                    w = ring[( u % HP_RING ) * 128];
                    xd_cp_async8( ring + ( ( u + HP_AHEAD ) % HP_RING ) * 128, ps );    // the slot read one row ago
                    ps += ls;
                }
                A[i5][0] = __byte_perm( w.x, 0u, 0x4140 );
                A[i5][1] = __byte_perm( w.x, 0u, 0x4342 );
                A[i5][2] = __byte_perm( w.y, 0u, 0x4140 );
                A[i5][3] = __byte_perm( w.y, 0u, 0x4342 );
                R[u % 3] = w;
            }
            // ---- V: biased 16-bit intermediates of this lane's eight columns
            uint32_t P[4];
#pragma unroll
            for( int j = 0; j < 4; j++ )
                P[j] = xd_hp_tap6_packed( A[i0][j], A[i1][j], A[i2][j], A[i3][j], A[i4][j], A[i5][j] );
            uint2 ov;
            {
                uint32_t f[4];
#pragma unroll
                for( int j = 0; j < 4; j++ )
                    f[j] = __vminu2( __vmaxu2( ( P[j] >> 5 ) & 0x07FF07FFu, 0x04000400u ), 0x04FF04FFu );   // clip((v+16)>>5) + 1024
                ov.x = __byte_perm( f[0], f[1], 0x6420 );
                ov.y = __byte_perm( f[2], f[3], 0x6420 );
            }

            // ---- H: byte windows at offsets -2, 0, 2, 4, 6, 8 of [prev | cur.x | cur.y | next]
            uint2 oh;
            {
                const uint32_t cP = __shfl_up_sync( 0xffffffffu, cur.y, 1, GW ), cN = __shfl_down_sync( 0xffffffffu, cur.x, 1, GW );
                const uint32_t wm2 = __funnelshift_r( cP, cur.x, 16 ), w2 = __funnelshift_r( cur.x, cur.y, 16 ), w6 = __funnelshift_r( cur.y, cN, 16 );
                const int h0 = xd_dp4a_us( wm2, TA0, xd_dp4a_us( w2, TB0, 16 ) ), h1 = xd_dp4a_us( wm2, TA1, xd_dp4a_us( w2, TB1, 16 ) );
                const int h2 = xd_dp4a_us( cur.x, TA0, xd_dp4a_us( cur.y, TB0, 16 ) ), h3 = xd_dp4a_us( cur.x, TA1, xd_dp4a_us( cur.y, TB1, 16 ) );
                const int h4 = xd_dp4a_us( w2, TA0, xd_dp4a_us( w6, TB0, 16 ) ), h5 = xd_dp4a_us( w2, TA1, xd_dp4a_us( w6, TB1, 16 ) );
                const int h6 = xd_dp4a_us( cur.y, TA0, xd_dp4a_us( cN, TB0, 16 ) ), h7 = xd_dp4a_us( cur.y, TA1, xd_dp4a_us( cN, TB1, 16 ) );
                oh.x = xd_pack_sat4( h0 >> 5, h1 >> 5, h2 >> 5, h3 >> 5 );
                oh.y = xd_pack_sat4( h4 >> 5, h5 >> 5, h6 >> 5, h7 >> 5 );
            }

            // ---- HV: six-tap over the unclipped V pairs of columns -2 .. 11
            uint2 oc;
            {
                const uint32_t Pm = __shfl_up_sync( 0xffffffffu, P[3], 1, GW );
                const uint32_t P4 = __shfl_down_sync( 0xffffffffu, P[0], 1, GW ), P5 = __shfl_down_sync( 0xffffffffu, P[1], 1, GW );
                const int kb = -( 1 << 20 );
                const int c0 = xd_dp2a_lo_us( Pm, T1, xd_dp2a_hi_us( P[0], T1, xd_dp2a_lo_us( P[1], T2, kb ) ) );
                const int c1 = xd_dp2a_hi_us( Pm, T2, xd_dp2a_lo_us( P[0], T3, xd_dp2a_hi_us( P[1], T3, xd_dp2a_lo_us( P[2], T4, kb ) ) ) );
                const int c2 = xd_dp2a_lo_us( P[0], T1, xd_dp2a_hi_us( P[1], T1, xd_dp2a_lo_us( P[2], T2, kb ) ) );
                const int c3 = xd_dp2a_hi_us( P[0], T2, xd_dp2a_lo_us( P[1], T3, xd_dp2a_hi_us( P[2], T3, xd_dp2a_lo_us( P[3], T4, kb ) ) ) );
                const int c4 = xd_dp2a_lo_us( P[1], T1, xd_dp2a_hi_us( P[2], T1, xd_dp2a_lo_us( P[3], T2, kb ) ) );
                const int c5 = xd_dp2a_hi_us( P[1], T2, xd_dp2a_lo_us( P[2], T3, xd_dp2a_hi_us( P[3], T3, xd_dp2a_lo_us( P4, T4, kb ) ) ) );
                const int c6 = xd_dp2a_lo_us( P[2], T1, xd_dp2a_hi_us( P[3], T1, xd_dp2a_lo_us( P4, T2, kb ) ) );
                const int c7 = xd_dp2a_hi_us( P[2], T2, xd_dp2a_lo_us( P[3], T3, xd_dp2a_hi_us( P4, T3, xd_dp2a_lo_us( P5, T4, kb ) ) ) );
                oc.x = xd_pack_sat4( c0 >> 10, c1 >> 10, c2 >> 10, c3 >> 10 );
                oc.y = xd_pack_sat4( c4 >> 10, c5 >> 10, c6 >> 10, c7 >> 10 );
            }

            if( writer && y < y_end )
            {
                *(uint2 *)dh = oh;
                *(uint2 *)dv = ov;
                *(uint2 *)dc = oc;
            }
            dh += ls;
            dv += ls;
            dc += ls;
        }
    }
    if( !TMA )
        xd_cp_async_wait<0>();                                   // nothing of this thread's may still be landing at exit
}

// The four warps of a block take four ADJACENT strips of the same row segment (they start together and stay close, so
// the block touches ~1 KB of consecutive bytes per row and plane at a time; with the warps on four row segments of one
// strip instead, every access of the kernel was an isolated 240-byte piece).  Strip n_full is the tail: the tail_units
// units left over, in strips of tail_gw lanes.
__global__ void __launch_bounds__( 128, 6 )
xd_hpel_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots, int n_full, int tail_units, int tail_gw )
{
    __shared__ uint2 s_ring[HP_RING * 128];
    uint8_t *slot = slots + blockIdx.z * (size_t)g.slot_bytes;
    const int strip = blockIdx.x * 4 + ( threadIdx.x >> 5 );
    const int wseg = blockIdx.y;
    const int unit0 = strip * HP_UNITS;
    if( strip > n_full || ( strip == n_full && !tail_units ) )
        return;
    if( strip < n_full || tail_gw == 32 )
        xd_hpel_body<32>( g, slot, unit0, strip < n_full ? HP_UNITS : tail_units, wseg, s_ring );
    else if( tail_gw == 4 )
        xd_hpel_body<4>( g, slot, unit0, tail_units, wseg, s_ring );
    else if( tail_gw == 8 )
        xd_hpel_body<8>( g, slot, unit0, tail_units, wseg, s_ring );
    else
        xd_hpel_body<16>( g, slot, unit0, tail_units, wseg, s_ring );
}

// The same with the full strips' source rows fetched by TMA (the tail strips keep the cp.async ring).
__global__ void __launch_bounds__( 128, 6 )
xd_hpel_tma_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots, int n_full, int tail_units, int tail_gw,
                    const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_halo )
{
    __shared__ uint2 s_ring[HP_RING * 128];
    __shared__ __align__( 128 ) uint8_t s_tiles[4 * HP_TSTAGES * HP_TILE_BYTES];
    __shared__ __align__( 8 ) uint64_t s_bars[4 * HP_TSTAGES];
    uint8_t *slot = slots + blockIdx.z * (size_t)g.slot_bytes;
    const int warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * 4 + warp;
    const int wseg = blockIdx.y;
    const int unit0 = strip * HP_UNITS;
    if( strip > n_full || ( strip == n_full && !tail_units ) )
        return;
    if( strip < n_full )
    {
        xd_hp_tma T;
        T.tmap = &tmap;
        T.tmap_halo = &tmap_halo;
        T.tiles = s_tiles + warp * HP_TSTAGES * HP_TILE_BYTES;
        T.bars = s_bars + warp * HP_TSTAGES;
        T.x = g.luma_origin % g.luma_stride + 8 * unit0;         // the strip's first real unit: a multiple of 16 (launcher checks)
        T.y = g.luma_origin / g.luma_stride;                     // + the segment's first source row, added in the body
        T.z = blockIdx.z;
        xd_hpel_body<32, true>( g, slot, unit0, HP_UNITS, wseg, s_ring, T );
    }
    else if( tail_gw == 32 )
        xd_hpel_body<32>( g, slot, unit0, tail_units, wseg, s_ring );
    else if( tail_gw == 4 )
        xd_hpel_body<4>( g, slot, unit0, tail_units, wseg, s_ring );
    else if( tail_gw == 8 )
        xd_hpel_body<8>( g, slot, unit0, tail_units, wseg, s_ring );
    else
        xd_hpel_body<16>( g, slot, unit0, tail_units, wseg, s_ring );
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

// row-major padded lowres planes -> tiled copies (for slots whose lowres planes the caller wrote).
// Thread = one 16-byte row chunk, i.e. one row of two neighbouring tiles; eight consecutive threads fill
// the tile pair (128 contiguous bytes out), a warp covers four pairs (64 contiguous bytes per row in).
__global__ void __launch_bounds__( 256 )
xd_retile_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots )
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int tp = t >> 3, row = t & 7;
    const int pairs_per_row = ( g.tile_w + 1 ) >> 1;         // tile_w = mb_w + 8 may be odd: the last pair is half
    if( tp >= pairs_per_row * g.tile_h )
        return;
    uint8_t *slot = slots + blockIdx.y * (size_t)g.slot_bytes;
    const int pl = blockIdx.z;
    const int ty = tp / pairs_per_row, tx = ( tp - ty * pairs_per_row ) * 2;
    const uint8_t *src = slot + g.slot_lowres_off + (size_t)pl * g.lowres_plane_size + g.lowres_origin
                       + (int64_t)( ty * 8 + row - X264DSP_PADV ) * g.lowres_stride + tx * 8 - X264DSP_PADH;
    uint8_t *dst = slot + g.slot_tiled_off + (size_t)pl * g.tiled_plane_size + ( (size_t)ty * g.tile_w + tx ) * 64 + row * 8;
    const uint2 lo = *(const uint2 *)src;
    *(uint2 *)dst = lo;
    if( tx + 1 < g.tile_w )
        *(uint2 *)( dst + 64 ) = *(const uint2 *)( src + 8 );
}

// tiled copies -> the reference's row-major padded lowres planes (lowres[0..3] incl. their 32-sample
// padding) in the slot's lowres region.  Thread = one 16-byte row-major chunk = one row of two tiles.
__global__ void __launch_bounds__( 256 )
xd_export_lowres_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slots )
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int tp = t >> 3, row = t & 7;
    const int pairs_per_row = ( g.tile_w + 1 ) >> 1;
    if( tp >= pairs_per_row * g.tile_h )
        return;
    uint8_t *slot = slots + blockIdx.y * (size_t)g.slot_bytes;
    const int pl = blockIdx.z;
    const int ty = tp / pairs_per_row, tx = ( tp - ty * pairs_per_row ) * 2;
    uint8_t *dst = slot + g.slot_lowres_off + (size_t)pl * g.lowres_plane_size + g.lowres_origin
                 + (int64_t)( ty * 8 + row - X264DSP_PADV ) * g.lowres_stride + tx * 8 - X264DSP_PADH;
    const uint8_t *src = slot + g.slot_tiled_off + (size_t)pl * g.tiled_plane_size + ( (size_t)ty * g.tile_w + tx ) * 64 + row * 8;
    *(uint2 *)dst = *(const uint2 *)src;
    if( tx + 1 < g.tile_w )
        *(uint2 *)( dst + 8 ) = *(const uint2 *)( src + 64 );
}

extern "C" int x264dsp_frame_export_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                                 int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const int chunks = ( ( g->tile_w + 1 ) / 2 ) * g->tile_h * 8;
    xd_export_lowres_kernel<<<dim3( ( chunks + 255 ) / 256, n_frames, 4 ), 256, 0, s>>>( *g, slots );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_frame_retile_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                                 int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const int chunks = ( ( g->tile_w + 1 ) / 2 ) * g->tile_h * 8;
    xd_retile_kernel<<<dim3( ( chunks + 255 ) / 256, n_frames, 4 ), 256, 0, s>>>( *g, slots );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

static int xd_launch_lowres( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots, const uint8_t *raw,
                             int n_frames, cudaStream_t s, bool keep_luma = true )
{
    dim3 grid( ( ( g->lowres_w >> 3 ) + 127 ) / 128, g->lowres_h / LR_ROWS, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_LOWRES, s );
    if( raw && keep_luma )
        xd_lowres_kernel<true, true><<<grid, 128, 0, s>>>( *g, slots, raw );
    else if( raw )
        xd_lowres_kernel<true, false><<<grid, 128, 0, s>>>( *g, slots, raw );
    else
        xd_lowres_kernel<false, true><<<grid, 128, 0, s>>>( *g, slots, NULL );
    xd_prof_end( ctx, XD_PROF_LOWRES, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_frame_init_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                               int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    return xd_launch_lowres( ctx, g, slots, NULL, n_frames, xd_stream( ctx, stream ) );
}

// x264dsp_frame_load_luma_dev followed by x264dsp_frame_init_lowres_dev, in one pass over the picture
int xd_frame_load_luma_lowres( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma, uint8_t *slots,
                               int n_frames, cudaStream_t s )
{
    return xd_launch_lowres( ctx, g, slots, luma, n_frames, s );
}

// lowres planes of the slots straight from the pictures; the slots' luma planes are not touched
int xd_frame_lowres_from_luma( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma, uint8_t *slots,
                               int n_frames, cudaStream_t s )
{
    return xd_launch_lowres( ctx, g, slots, luma, n_frames, s, false );
}

extern "C" int x264dsp_frame_lowres_from_luma_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                                    uint8_t *slots, int n_frames, void *stream )
{
    if( !ctx || !g || !luma || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    return xd_launch_lowres( ctx, g, slots, luma, n_frames, xd_stream( ctx, stream ), false );
}

extern "C" int x264dsp_frame_load_luma_lowres_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma,
                                                    uint8_t *slots, int n_frames, void *stream )
{
    if( !ctx || !g || !luma || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    return xd_launch_lowres( ctx, g, slots, luma, n_frames, xd_stream( ctx, stream ) );
}

// plane N of n_frames consecutive slots as a (byte in row, row, slot) tensor of bytes; box = box_bytes x HP_TROWS rows
#ifndef XD_HPEL_TMA_DEFAULT
#define XD_HPEL_TMA_DEFAULT 0
#endif
typedef CUresult ( *xd_encode_tiled_fn )( CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
static int xd_hpel_tensor_map( const x264dsp_geom_t *g, uint8_t *slots, int n_frames, int box_bytes, CUtensorMap *out )
{
    static xd_encode_tiled_fn encode = nullptr;
    static bool tried = false;
    if( !tried )
    {
        tried = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q ) == cudaSuccess && q == cudaDriverEntryPointSuccess )
            encode = (xd_encode_tiled_fn)fn;
    }
    if( !encode || ( (uintptr_t)slots & 15 ) || ( g->luma_stride & 15 ) || ( g->slot_bytes & 15 ) )
        return 1;
    const cuuint64_t dims[3] = { (cuuint64_t)g->luma_stride, (cuuint64_t)( g->luma_plane_size / g->luma_stride ), (cuuint64_t)n_frames };
    const cuuint64_t strides[2] = { (cuuint64_t)g->luma_stride, (cuuint64_t)g->slot_bytes };
    const cuuint32_t box[3] = { (cuuint32_t)box_bytes, HP_TROWS, 1 }, estr[3] = { 1, 1, 1 };
    return encode( out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, slots, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE ) == CUDA_SUCCESS ? 0 : 1;
}

extern "C" int x264dsp_frame_filter_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots,
                                          int n_frames, void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const int n_units = ( g->luma_w + 8 ) >> 3, n_segs = ( g->luma_h + 16 + HP_ROWS - 1 ) / HP_ROWS;
    const int n_full = n_units / HP_UNITS, tail_units = n_units % HP_UNITS;
    const int tail_gw = tail_units <= 2 ? 4 : tail_units <= 6 ? 8 : tail_units <= 14 ? 16 : 32;
    dim3 grid( ( n_full + ( tail_units ? 1 : 0 ) + 3 ) / 4, n_segs, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_HPEL, s );
    const char *tma_env = getenv( "X264DSP_HPEL_TMA" );          // measurement / test switch for the TMA variant
    const int use_tma = tma_env ? atoi( tma_env ) : XD_HPEL_TMA_DEFAULT;
    CUtensorMap tmap, tmap_halo;
    // strips start at origin + 240 k bytes: aligned boxes need the plane's first sample on a 16-byte boundary of its row
    if( use_tma && n_full > 0 && ( g->luma_origin % g->luma_stride ) % 16 == 0 && xd_hpel_tensor_map( g, slots, n_frames, 256, &tmap ) == 0
        && xd_hpel_tensor_map( g, slots, n_frames, 16, &tmap_halo ) == 0 )
        xd_hpel_tma_kernel<<<grid, 128, 0, s>>>( *g, slots, n_full, tail_units, tail_gw, tmap, tmap_halo );
    else
        xd_hpel_kernel<<<grid, 128, 0, s>>>( *g, slots, n_full, tail_units, tail_gw );
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
