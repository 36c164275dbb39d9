// lookahead.cu -- lowres lookahead frame-cost pass, sm_100a.
//
// Reference: x264_slicetype_mb_cost / x264_slicetype_frame_cost (encoder/slicetype.c:48-322) with
// the search it drives, x264_me_search_ref + refine_subpel at DIA / subme 2 (encoder/me.c:129-587),
// on 8x8 blocks of the half-resolution planes, lambda = 1 (QP 12), do_edges = 0.
//
// Dependency structure: blocks are visited in reverse raster order and block (x,y)
// takes its MV predictors from (x+1,y), (x,y+1), (x-1,y+1), (x+1,y+1).  Rows therefore pipeline
// with a lag of two blocks.  Mapping:
//   * one WARP owns one block row of one frame pair and walks it right to left;
//   * rows are handed out through an atomic ticket in dependency order (bottom rows of all pairs
//     first, then the next row of all pairs, ...), so the row a warp waits on is always held by a
//     warp that is already running -- no co-residency assumption, no deadlock -- and the wavefronts
//     of all pairs of a launch advance together;
//   * a finished block publishes {mv, epoch} as ONE 64-bit word; the row above spins on that word
//     (volatile load, issued a block ahead of its use).  The payload is the flag, so no fence sits
//     on the critical path;
//   * inside a block the warp evaluates up to four candidate positions at once: lane = 8*cand+row,
//     8 pixels per lane (VABSDIFF4.U8.ACC x2), xor-shuffle reduction over the 8 rows, then a
//     packed (cost,order) min over the candidates that reproduces the reference's strict-'<',
//     first-wins tie breaking;
//   * the intra estimate has no dependencies and runs first as a plain thread-per-block kernel.
// Many frame pairs per launch fill the machine; a single pair is latency bound by construction.
#include <stdlib.h>
#include "common.cuh"

#define LA_COST_MAX ( 1 << 28 )
#define LA_WARPS 4

struct xd_la_args
{
    x264dsp_geom_t g;
    const uint8_t *slots;
    const int32_t *b, *p0;
    const uint8_t *want_intra;
    int n_pairs, pair0;
    int16_t *mvs;
    int32_t *costs, *sums, *row_satds;
    const uint16_t *cost_mv;            // centre of the lambda=1 table
    unsigned long long *sync;
    int32_t *icost;
    int32_t *ticket;
    uint32_t epoch;
    int me_range;
    int slack;                          // multi-row kernel: extra blocks a unit stays behind the one below
    unsigned long long *timing;         // optional phase-cycle counters (x264dsp_debug_lookahead_timing)
};

// phase-cycle accounting of the inter kernel, off unless A.timing is set: lane 0 of every warp adds the
// clock64() deltas of its phases at the end of its row
enum { LA_T_WAIT = 0, LA_T_SETUP, LA_T_ZERO, LA_T_CAND, LA_T_DIA, LA_T_SUBPEL, LA_T_TAIL, LA_T_BLOCKS, LA_T_FIRST_WAIT, LA_T_ROWS, LA_T_KINDS };
#define LA_TICK( kind ) do { if( TIMED ) { const unsigned now_ = (unsigned)clock(); tacc[kind] += now_ - tlast; tlast = now_; } } while( 0 )

// ---------------------------------------------------------------------------------------------
// 4-point Hadamard butterfly; output 0 is the plain sum
__device__ __forceinline__ void xd_had4( int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3 )
{
    const int s01 = a + b, d01 = a - b, s23 = c + d, d23 = c - d;
    o0 = s01 + s23; o1 = s01 - s23; o2 = d01 + d23; o3 = d01 - d23;
}

// sum |H4 D H4^T| of a 4x4 block given as four packed rows of differences' operands
__device__ __forceinline__ int xd_satd4x4_words( const uint32_t a[4], const uint32_t b[4] )
{
    int t[4][4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        const int d0 = (int)( a[r] & 255 ) - (int)( b[r] & 255 );
        const int d1 = (int)( ( a[r] >> 8 ) & 255 ) - (int)( ( b[r] >> 8 ) & 255 );
        const int d2 = (int)( ( a[r] >> 16 ) & 255 ) - (int)( ( b[r] >> 16 ) & 255 );
        const int d3 = (int)( a[r] >> 24 ) - (int)( b[r] >> 24 );
        xd_had4( d0, d1, d2, d3, t[r][0], t[r][1], t[r][2], t[r][3] );
    }
    int acc = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        int o0, o1, o2, o3;
        xd_had4( t[0][c], t[1][c], t[2][c], t[3][c], o0, o1, o2, o3 );
        acc += abs( o0 ) + abs( o1 ) + abs( o2 ) + abs( o3 );
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// The lookahead reads the lowres planes through the slot's 8x8-TILED copies (x264dsp_geom_t: tile
// (tx,ty) of the padded plane = 64 contiguous bytes, tiles of a tile row back to back).  A lane reads a
// pixel row of a block; in a row-major plane the rows of a block are one cache line each and the L1 data
// pipe saturates (ncu: 91 % of peak wavefronts at 4 % DRAM), tiled they share one or two lines.
//
// 8 pixels at padded coordinates (X,Y) = (x+32, y+32) of one tiled plane: the row chunk of the tile
// holding X and of its right-hand neighbour (two aligned 8-byte loads, 64 bytes apart), then a funnel
// shift by the position inside the chunk (shf.r.wrap takes the shift modulo 32)
__device__ __forceinline__ uint2 xd_tl8( const uint8_t *tplane, int tw, int X, int Y )
{
    const uint2 *w = (const uint2 *)( tplane + ( ( ( Y >> 3 ) * tw + ( X >> 3 ) ) << 6 ) + ( ( Y & 7 ) << 3 ) );
    const uint2 lo = __ldg( w ), hi = __ldg( w + 8 );
    const uint32_t sh = (uint32_t)X << 3;
    const bool up = ( X & 4 ) != 0;
    const uint32_t w0 = up ? lo.y : lo.x, w1 = up ? hi.x : lo.y, w2 = up ? hi.y : hi.x;
    return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
}

// ---------------------------------------------------------------------------------------------
// intra estimate: x264_intra_satd_x3_8x8c on the SOURCE lowres plane (slicetype.c:145-180).
// Hadamard is linear, so satd(pred - src) is evaluated from ONE transform of the source 4x4 and
// the (sparse) transforms of the three predictions: DC touches coefficient (0,0), V the first row
// of coefficients, H the first column.
__global__ void __launch_bounds__( 128 )
xd_la_intra_kernel( xd_la_args A )
{
    const int pair = A.pair0 + blockIdx.y;
    if( !A.want_intra[pair] )
        return;
    const x264dsp_geom_t &g = A.g;
    const int W = g.mb_w, H = g.mb_h;
    const int inner = ( W - 2 ) * ( H - 2 );
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int icost = 0, by = 0;
    if( i < inner )
    {
        const int bx = 1 + i % ( W - 2 );
        by = 1 + i / ( W - 2 );
        // block (bx,by) is tile (bx+4, by+4) of the tiled plane N; its left neighbours are the last
        // column of the tile before it, its top neighbours the last row of the tile above
        const uint8_t *src = A.slots + (size_t)A.b[pair] * g.slot_bytes + g.slot_tiled_off
                           + ( (size_t)( by + 4 ) * g.tile_w + bx + 4 ) * 64;
        uint2 row[8];
        uint32_t left[8];
#pragma unroll
        for( int r = 0; r < 8; r++ )
        {
            row[r] = __ldg( (const uint2 *)( src + r * 8 ) );
            left[r] = __ldg( src - 64 + r * 8 + 7 );
        }
        const uint2 top = __ldg( (const uint2 *)( src - (size_t)g.tile_w * 64 + 56 ) );
        int tp[8];
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            tp[k] = ( top.x >> ( 8 * k ) ) & 255;
            tp[4 + k] = ( top.y >> ( 8 * k ) ) & 255;
        }
        // quadrant DC values (predict.c:224-262)
        const int s0 = tp[0] + tp[1] + tp[2] + tp[3], s1 = tp[4] + tp[5] + tp[6] + tp[7];
        const int s2 = left[0] + left[1] + left[2] + left[3], s3 = left[4] + left[5] + left[6] + left[7];
        const int dcq[4] = { ( s0 + s2 + 4 ) >> 3, ( s1 + 2 ) >> 2, ( s3 + 2 ) >> 2, ( s1 + s3 + 4 ) >> 3 };

        int sat_dc[4], sat_h[4], sat_v[4];
#pragma unroll
        for( int q = 0; q < 4; q++ )
        {
            const int qx = q & 1, qy = q >> 1;
            int t[4][4];
#pragma unroll
            for( int r = 0; r < 4; r++ )
            {
                const uint32_t w = qx ? row[4 * qy + r].y : row[4 * qy + r].x;
                xd_had4( w & 255, ( w >> 8 ) & 255, ( w >> 16 ) & 255, w >> 24, t[r][0], t[r][1], t[r][2], t[r][3] );
            }
            int f[4][4];                       // f[u][v]: u vertical, v horizontal frequency
#pragma unroll
            for( int c = 0; c < 4; c++ )
                xd_had4( t[0][c], t[1][c], t[2][c], t[3][c], f[0][c], f[1][c], f[2][c], f[3][c] );
            int all = 0, row0 = 0, col0 = 0;
#pragma unroll
            for( int u = 0; u < 4; u++ )
#pragma unroll
                for( int v = 0; v < 4; v++ )
                    all += abs( f[u][v] );
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                row0 += abs( f[0][k] );
                col0 += abs( f[k][0] );
            }
            int tv[4], th[4];
            xd_had4( tp[4 * qx], tp[4 * qx + 1], tp[4 * qx + 2], tp[4 * qx + 3], tv[0], tv[1], tv[2], tv[3] );
            xd_had4( left[4 * qy], left[4 * qy + 1], left[4 * qy + 2], left[4 * qy + 3], th[0], th[1], th[2], th[3] );
            int v_part = 0, h_part = 0;
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                v_part += abs( f[0][k] - 4 * tv[k] );
                h_part += abs( f[k][0] - 4 * th[k] );
            }
            sat_dc[q] = all - abs( f[0][0] ) + abs( f[0][0] - 16 * dcq[q] );
            sat_v[q] = all - row0 + v_part;
            sat_h[q] = all - col0 + h_part;
        }
        // pixel.c:294-335: an 8x8 is two 8x4 halves, each halved once
        const int c_dc = ( ( sat_dc[0] + sat_dc[1] ) >> 1 ) + ( ( sat_dc[2] + sat_dc[3] ) >> 1 );
        const int c_h = ( ( sat_h[0] + sat_h[1] ) >> 1 ) + ( ( sat_h[2] + sat_h[3] ) >> 1 );
        const int c_v = ( ( sat_v[0] + sat_v[1] ) >> 1 ) + ( ( sat_v[2] + sat_v[3] ) >> 1 );
        icost = min( c_dc, min( c_h, c_v ) ) + 5 + 4;         // intra_penalty + lowres_penalty
        A.icost[(size_t)pair * g.mb_count + by * W + bx] = icost;
        if( A.row_satds )
            atomicAdd( &A.row_satds[( (size_t)pair * 2 + 1 ) * H + by], icost );
    }
    // block-level sum of the frame's intra cost (integer adds: order independent)
    int sum = icost, cnt = i < inner ? 1 : 0;
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
    {
        sum += __shfl_xor_sync( 0xffffffffu, sum, o );
        cnt += __shfl_xor_sync( 0xffffffffu, cnt, o );
    }
    if( ( threadIdx.x & 31 ) == 0 && cnt )
    {
        int32_t *s = A.sums + (size_t)pair * X264DSP_LA_SUMS;
        atomicAdd( &s[X264DSP_LA_COST_INTRA], sum );
        atomicAdd( &s[X264DSP_LA_SATD_EVALS], 3 * cnt );
        if( A.p0[pair] < 0 )
            atomicAdd( &s[X264DSP_LA_INTRA_MBS], cnt );   // intra-only analysis: every block is intra
    }
}

// ---------------------------------------------------------------------------------------------
// inter search: warp-wide helpers.  All scalar state is replicated in every lane.

struct xd_la_block
{
    const uint8_t *tref;      // tiled planes N, H, V, HV of the reference frame
    int tplane, tw;           // bytes per tiled plane, tiles per tile row
    int X0, Y0;               // padded coordinates of the block origin
    uint2 fenc;               // this lane's source row (row = lane & 7)
    uint32_t fq[4];           // rows of the 4x4 quadrant (lane & 3) of the source block
    int mvpx, mvpy;
    const uint16_t *cost_mv;
    int sad_evals, satd_evals;
};

__device__ __forceinline__ int xd_la_bits( const xd_la_block &B, int qx, int qy )
{
    return __ldg( B.cost_mv + ( qx - B.mvpx ) ) + __ldg( B.cost_mv + ( qy - B.mvpy ) );
}

// 8 pixels of row `row` of the prediction at quarter-pel position (qx,qy): get_ref / mc_luma
__device__ __forceinline__ uint2 xd_la_fetch( const xd_la_block &B, int qx, int qy, int row )
{
    const int fx = qx & 3, fy = qy & 3, phase = fy * 4 + fx;
    const int X = B.X0 + ( qx >> 2 ), Y = B.Y0 + ( qy >> 2 ) + row;
    uint2 a = xd_tl8( B.tref + xd_qpel_plane_a( phase ) * B.tplane, B.tw, X, Y + ( fy == 3 ? 1 : 0 ) );
    if( phase & 5 )
    {
        const uint2 b = xd_tl8( B.tref + xd_qpel_plane_b( phase ) * B.tplane, B.tw, X + ( fx == 3 ? 1 : 0 ), Y );
        a.x = xd_avg4( a.x, b.x );
        a.y = xd_avg4( a.y, b.y );
    }
    return a;
}

// Costs of the four candidate groups in one step.  Lane = 8*cand + row holds the SAD of its 8
// pixels; the lane of row 0 also adds the candidate's mv bits (or 0xFFFF when the candidate does not
// compete).  Each candidate owns a 16-bit field (an 8x8 SAD plus its mv bits stays below 2^15 at
// lambda = 1), so two whole-warp REDUX.SUM instructions replace the 3 + 2 dependent shuffles of a
// tree reduction; afterwards EVERY lane holds all four costs and picks the winner locally.
struct xd_cost4
{
    int c[4];
};
__device__ __forceinline__ xd_cost4 xd_la_reduce4( int partial, int lane )
{
    const unsigned v = (unsigned)partial << ( ( lane & 8 ) ? 16 : 0 );
    const unsigned lo = __reduce_add_sync( 0xffffffffu, ( lane & 16 ) ? 0u : v );   // candidates 0, 1
    const unsigned hi = __reduce_add_sync( 0xffffffffu, ( lane & 16 ) ? v : 0u );   // candidates 2, 3
    xd_cost4 r;
    r.c[0] = (int)( lo & 0xFFFF ); r.c[1] = (int)( lo >> 16 );
    r.c[2] = (int)( hi & 0xFFFF ); r.c[3] = (int)( hi >> 16 );
    return r;
}

// this lane's share of the cost of candidate (lane>>3) at quarter-pel (qx,qy):
// SAD of its row, plus (row 0 only) the mv bits, or the "does not compete" marker
__device__ __forceinline__ int xd_la_partial( const xd_la_block &B, int qx, int qy, int lane, bool ok, bool with_bits )
{
    const uint2 p = xd_la_fetch( B, qx, qy, lane & 7 );
    int s = (int)( __vsadu4( p.x, B.fenc.x ) + __vsadu4( p.y, B.fenc.y ) );
    if( !ok )
        s = 0;
    if( ( lane & 7 ) == 0 )
        s += !ok ? 0xFFFF : with_bits ? xd_la_bits( B, qx, qy ) : 0;
    return s;
}

// 4 pixels of the prediction at (qx,qy), block-relative position (x,y)
__device__ __forceinline__ uint32_t xd_la_fetch4( const xd_la_block &B, int qx, int qy, int x, int y )
{
    const int fx = qx & 3, fy = qy & 3, phase = fy * 4 + fx;
    const int X = B.X0 + ( qx >> 2 ) + x, Y = B.Y0 + ( qy >> 2 ) + y;
    uint32_t a = xd_tl8( B.tref + xd_qpel_plane_a( phase ) * B.tplane, B.tw, X, Y + ( fy == 3 ? 1 : 0 ) ).x;
    if( phase & 5 )
        a = xd_avg4( a, xd_tl8( B.tref + xd_qpel_plane_b( phase ) * B.tplane, B.tw, X + ( fx == 3 ? 1 : 0 ), Y ).x );
    return a;
}

// SATD 8x8 of the prediction at (qx,qy) against the source block (pixel.c:294-335).
// Lane & 3 selects the 4x4 quadrant; every lane loads its quadrant directly (no redistribution).
__device__ __forceinline__ int xd_la_satd( const xd_la_block &B, int qx, int qy, int lane )
{
    const int q = lane & 3, x = ( q & 1 ) * 4, y = ( q >> 1 ) * 4;
    uint32_t pr[4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
        pr[r] = xd_la_fetch4( B, qx, qy, x, y + r );
    const unsigned s = (unsigned)xd_satd4x4_words( B.fq, pr );
    // upper 8x4 = quadrants 0+1, lower 8x4 = quadrants 2+3, each halved once
    const unsigned up = __reduce_add_sync( 0xffffffffu, lane < 2 ? s : 0u );
    const unsigned dn = __reduce_add_sync( 0xffffffffu, ( lane & ~1 ) == 2 ? s : 0u );
    return (int)( ( up >> 1 ) + ( dn >> 1 ) );
}

// CHECK_MVRANGE (me.c:155-160)
__device__ __forceinline__ bool xd_la_in_range( int mx, int my, int minx, int miny, int maxx, int maxy )
{
    const uint32_t lo = ( (uint32_t)( -minx ) << 16 ) | ( (uint32_t)( -miny ) & 0x7FFF );
    const uint32_t hi = ( (uint32_t)maxx << 16 ) | ( (uint32_t)maxy & 0x7FFF ) | 0x8000;
    const uint32_t v = ( (uint32_t)mx << 16 ) | ( (uint32_t)my & 0x7FFF );
    return !( ( ( v + lo ) | ( hi - v ) ) & 0x80004000u );
}

__device__ __forceinline__ int xd_median3( int a, int b, int c )
{
    return max( min( a, b ), min( max( a, b ), c ) );
}

__device__ __forceinline__ unsigned long long xd_ld_sync( const unsigned long long *p )
{
    unsigned long long v;
    asm volatile( "ld.volatile.global.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );
    return v;
}

__device__ __forceinline__ void xd_st_sync( unsigned long long *p, unsigned long long v )
{
    asm volatile( "st.volatile.global.u64 [%0], %1;" :: "l"( p ), "l"( v ) : "memory" );
}

// wait until the word carries this launch's epoch; returns the packed mv
// `ns0` / `ns_max`: first and longest pause between polls.  Inside a row the expected wait is a
// fraction of a block time, so polls are dense; the FIRST wait of a row (its lower neighbour has to
// get two blocks ahead, which for the upper rows of a frame means waiting for most of the pipeline
// fill) backs off much further so that parked warps leave the issue slots to the working ones.
__device__ __forceinline__ uint32_t xd_la_await( const unsigned long long *p, unsigned long long seen, uint32_t epoch,
                                                 unsigned ns0 = 20, unsigned ns_max = 200 )
{
    unsigned ns = ns0;
    while( (uint32_t)( seen >> 32 ) != epoch )
    {
        __nanosleep( ns );
        if( ns < ns_max )
            ns += ns0;
        seen = xd_ld_sync( p );
    }
    return (uint32_t)seen;
}

#define MVX( m ) ( (int)(int16_t)( ( m ) & 0xFFFF ) )
#define MVY( m ) ( (int)(int16_t)( ( m ) >> 16 ) )

template<bool TIMED>
__global__ void __launch_bounds__( LA_WARPS * 32, 8 )
xd_la_inter_kernel( xd_la_args A, int n_inter, const int32_t *inter_pairs )
{
    const x264dsp_geom_t &g = A.g;
    const int lane = threadIdx.x & 31;
    const int W = g.mb_w, H = g.mb_h, ls = g.lowres_stride;
    const int rows = H - 2;
    const int total = n_inter * rows;
    const int cand = lane >> 3;

    for( ;; )
    {
        int ticket = 0;
        if( lane == 0 )
            ticket = atomicAdd( A.ticket, 1 );
        ticket = __shfl_sync( 0xffffffffu, ticket, 0 );
        if( ticket >= total )
            return;
        // row-major over the pairs of the launch: all bottom rows first, then all second rows, ...
        // Row (pair, r) waits on (pair, r-1), whose ticket is n_inter smaller and therefore already
        // held by a running warp; and when the launch has more rows than the machine has warp slots,
        // the rows handed out next are exactly the ones whose lower neighbours are furthest along,
        // so warps spend their residency searching instead of waiting for a pipeline to fill.
        const int row_idx = ticket / n_inter;
        const int pair = inter_pairs[ticket - row_idx * n_inter];
        const int by = H - 2 - row_idx;
        // the source block (bx,by) is tile (bx+4, by+4) of the current frame's tiled plane N
        const uint8_t *cur = A.slots + (size_t)A.b[pair] * g.slot_bytes + g.slot_tiled_off
                           + ( (size_t)( by + 4 ) * g.tile_w + 4 ) * 64;
        unsigned long long *sync_row = A.sync + (size_t)pair * g.mb_count + (size_t)by * W;
        const unsigned long long *sync_below = sync_row + W;
        const bool has_below = by < H - 2;
        const bool want_intra = A.want_intra[pair] != 0;

        // MV limits (slicetype.c:79-89); the y limits are per row
        const int miny = -( by << 3 ) - 4, maxy = ( ( H - by - 1 ) << 3 ) + 4;
        const int sminy = ( miny - 8 ) << 2, smaxy = ( maxy + 8 ) << 2;

        xd_la_block B;
        B.tref = A.slots + (size_t)A.p0[pair] * g.slot_bytes + g.slot_tiled_off;
        B.tplane = g.tiled_plane_size;
        B.tw = g.tile_w;
        B.X0 = B.Y0 = 32;
        B.cost_mv = A.cost_mv;
        B.sad_evals = 0;
        B.satd_evals = 0;

        unsigned tacc[LA_T_KINDS] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, tlast = TIMED ? (unsigned)clock() : 0u;
        uint32_t mv_right = 0;                           // (x+1, y): border block
        uint32_t mv_b = 0, mv_br = 0, mv_bl = 0;         // (x, y+1), (x+1, y+1), (x-1, y+1)
        unsigned long long pending = 0;
        if( has_below )
        {
            mv_b = xd_la_await( sync_below + ( W - 2 ), xd_ld_sync( sync_below + ( W - 2 ) ), A.epoch, 250, 4000 );
            pending = xd_ld_sync( sync_below + ( W - 3 ) );
        }
        int row_sum = 0, row_intra = 0;
        LA_TICK( LA_T_FIRST_WAIT );

        // The source block, its SATD quadrant rows and the block's intra cost do not depend on any
        // search result: they are fetched one block ahead so that their L2 / HBM latency overlaps
        // the previous block's search instead of opening every block with a stall.
        const int fq_off = ( ( ( lane & 3 ) >> 1 ) * 4 ) * 8 + ( lane & 1 ) * 4;      // inside the 64-byte tile
        const int32_t *icost_row = A.icost + (size_t)pair * g.mb_count + (size_t)by * W;
        uint2 nx_fenc;
        uint32_t nx_fq[4];
        int nx_ic = 0;
        {
            const size_t pel0 = (size_t)( W - 2 ) * 64;
            nx_fenc = __ldg( (const uint2 *)( cur + pel0 + ( lane & 7 ) * 8 ) );
#pragma unroll
            for( int r = 0; r < 4; r++ )
                nx_fq[r] = __ldg( (const uint32_t *)( cur + pel0 + fq_off + r * 8 ) );
            if( want_intra )
                nx_ic = icost_row[W - 2];
        }

        for( int bx = W - 2; bx >= 1; bx-- )
        {
            // (x-1, y+1): needed now; for x == 1 that is a border block (zero)
            if( has_below && bx >= 2 )
                mv_bl = xd_la_await( sync_below + ( bx - 1 ), pending, A.epoch );
            else
                mv_bl = 0;
            // start fetching the word the NEXT block will need
            if( has_below && bx >= 3 )
                pending = xd_ld_sync( sync_below + ( bx - 2 ) );
            LA_TICK( LA_T_WAIT );

            const size_t pel = (size_t)bx * 64;
            B.X0 = 8 * bx + 32;
            B.Y0 = 8 * by + 32;
            B.fenc = nx_fenc;
#pragma unroll
            for( int r = 0; r < 4; r++ )
                B.fq[r] = nx_fq[r];
            const int ic = nx_ic;
            if( bx > 1 )
            {
                nx_fenc = __ldg( (const uint2 *)( cur + pel - 64 + ( lane & 7 ) * 8 ) );
#pragma unroll
                for( int r = 0; r < 4; r++ )
                    nx_fq[r] = __ldg( (const uint32_t *)( cur + pel - 64 + fq_off + r * 8 ) );
                if( want_intra )
                    nx_ic = icost_row[bx - 1];
            }
            const int minx = -( bx << 3 ) - 4, maxx = ( ( W - bx - 1 ) << 3 ) + 4;
            const int sminx = ( minx - 8 ) << 2, smaxx = ( maxx + 8 ) << 2;

            // predictors (slicetype.c:105-113): right, below, below-left, below-right
            B.mvpx = xd_median3( MVX( mv_right ), MVX( mv_b ), MVX( mv_bl ) );
            B.mvpy = xd_median3( MVY( mv_right ), MVY( mv_b ), MVY( mv_bl ) );

            int mvx = 0, mvy = 0, cost = -1;
            LA_TICK( LA_T_SETUP );
            if( !( B.mvpx | B.mvpy ) )                     // slicetype.c:117-125
            {
                const int c0 = xd_la_satd( B, 0, 0, lane );
                B.satd_evals++;
                if( c0 < 64 )
                    cost = c0;
            }
            LA_TICK( LA_T_ZERO );
            if( cost < 0 )
            {
                // ---- x264_me_search_ref, subme < 3 branch (me.c:194-233)
                int bmx = xd_clip3( B.mvpx, minx * 4, maxx * 4 ), bmy = xd_clip3( B.mvpy, miny * 4, maxy * 4 );
                const int pmx = ( bmx + 2 ) >> 2, pmy = ( bmy + 2 ) >> 2;
                const uint32_t pmv = ( (uint32_t)pmx & 0xFFFF ) | ( (uint32_t)pmy << 16 );
                int bcost;
                {
                    // candidates in evaluation order: 0 = rounded MVP (no mv cost), 1..4 = mvc, 5 = (0,0);
                    // pass 0 carries 0..3 in the four lane groups, pass 1 carries 4 and 5
                    int best = 0x7FFFFFFF;
                    int n_evals = 0;
#pragma unroll
                    for( int pass = 0; pass < 2; pass++ )
                    {
                        const int k = pass * 4 + cand;
                        int cx = 0, cy = 0;
                        bool ok = false;
                        if( k == 0 )
                        {
                            cx = pmx; cy = pmy; ok = true;
                        }
                        else if( k <= 4 )
                        {
                            const uint32_t m = k == 1 ? mv_right : k == 2 ? mv_b : k == 3 ? mv_bl : mv_br;
                            cx = xd_clip3( ( MVX( m ) + 2 ) >> 2, minx, maxx );
                            cy = xd_clip3( ( MVY( m ) + 2 ) >> 2, miny, maxy );
                            const uint32_t v = ( (uint32_t)cx & 0xFFFF ) | ( (uint32_t)cy << 16 );
                            ok = v != 0 && v != pmv;
                        }
                        else if( k == 5 )
                            ok = pmv != 0;
                        const xd_cost4 r = xd_la_reduce4( xd_la_partial( B, cx << 2, cy << 2, lane, ok, k != 0 ), lane );
#pragma unroll
                        for( int c = 0; c < 4; c++ )
                            if( r.c[c] < 0xFFFF )
                            {
                                best = min( best, ( r.c[c] << 3 ) | ( pass * 4 + c ) );
                                n_evals++;
                            }
                    }
                    bcost = best >> 3;
                    const int win = best & 7;
                    if( win == 0 ) { bmx = pmx; bmy = pmy; }
                    else if( win == 5 ) { bmx = 0; bmy = 0; }
                    else
                    {
                        const uint32_t m = win == 1 ? mv_right : win == 2 ? mv_b : win == 3 ? mv_bl : mv_br;
                        bmx = xd_clip3( ( MVX( m ) + 2 ) >> 2, minx, maxx );
                        bmy = xd_clip3( ( MVY( m ) + 2 ) >> 2, miny, maxy );
                    }
                    B.sad_evals += n_evals;          // MVP + competing candidates + (0,0), as the reference issues
                }

                LA_TICK( LA_T_CAND );
                // ---- diamond search (me.c:237-274): up, down, left, right
                {
                    const int dx = cand == 2 ? -1 : cand == 3 ? 1 : 0;
                    const int dy = cand == 0 ? -1 : cand == 1 ? 1 : 0;
                    int left = A.me_range;
                    do
                    {
                        const xd_cost4 r = xd_la_reduce4( xd_la_partial( B, ( bmx + dx ) << 2, ( bmy + dy ) << 2, lane, true, true ), lane );
                        B.sad_evals += 4;
                        const int key = min( min( ( r.c[0] << 2 ) | 0, ( r.c[1] << 2 ) | 1 ), min( ( r.c[2] << 2 ) | 2, ( r.c[3] << 2 ) | 3 ) );
                        if( ( key >> 2 ) >= bcost )
                            break;
                        bcost = key >> 2;
                        const int w = key & 3;
                        bmx += w == 2 ? -1 : w == 3 ? 1 : 0;
                        bmy += w == 0 ? -1 : w == 1 ? 1 : 0;
                    } while( --left && xd_la_in_range( bmx, bmy, minx, miny, maxx, maxy ) );
                }

                LA_TICK( LA_T_DIA );
                // ---- me.c:397-414
                int qx = bmx << 2, qy = bmy << 2;
                if( bmx == pmx && bmy == pmy )
                    bcost += xd_la_bits( B, qx, qy );

                // ---- refine_subpel( hpel_iters = 1, qpel_iters = 0 ) (me.c:466-587)
                {
                    const int px = xd_clip3( B.mvpx, sminx + 2, smaxx - 2 ), py = xd_clip3( B.mvpy, sminy + 2, smaxy - 2 );
                    if( px != qx || py != qy )                       // me.c:483-490
                    {
                        // a single candidate: let group 0 carry it
                        const xd_cost4 r = xd_la_reduce4( xd_la_partial( B, px, py, lane, cand == 0, true ), lane );
                        B.sad_evals++;
                        if( r.c[0] < bcost ) { bcost = r.c[0]; qx = px; qy = py; }
                    }
                    // half-pel diamond (me.c:492-517)
                    const int hx = qx + ( cand == 2 ? -2 : cand == 3 ? 2 : 0 );
                    const int hy = qy + ( cand == 0 ? -2 : cand == 1 ? 2 : 0 );
                    const xd_cost4 r = xd_la_reduce4( xd_la_partial( B, hx, hy, lane, true, true ), lane );
                    B.sad_evals += 4;
                    const int key = min( min( ( r.c[0] << 2 ) | 0, ( r.c[1] << 2 ) | 1 ), min( ( r.c[2] << 2 ) | 2, ( r.c[3] << 2 ) | 3 ) );
                    if( ( key >> 2 ) < bcost )
                    {
                        const int w = key & 3;
                        qx += w == 2 ? -2 : w == 3 ? 2 : 0;
                        qy += w == 0 ? -2 : w == 1 ? 2 : 0;
                    }
                    // me.c:519-524: the winner is re-costed with SATD
                    bcost = xd_la_satd( B, qx, qy, lane ) + xd_la_bits( B, qx, qy );
                    B.satd_evals++;
                }
                LA_TICK( LA_T_SUBPEL );
                mvx = qx; mvy = qy;
                cost = bcost - 1;                                    // slicetype.c:128-130
                if( mvx | mvy )
                    cost += 5;
            }

            // publish, then account (slicetype.c:132-196)
            const uint32_t mv_packed = ( (uint32_t)mvx & 0xFFFF ) | ( (uint32_t)mvy << 16 );
            const int xy = by * W + bx;
            if( lane == 0 )
            {
                xd_st_sync( sync_row + bx, ( (unsigned long long)A.epoch << 32 ) | mv_packed );
                *(uint32_t *)( A.mvs + ( (size_t)pair * g.mb_count + xy ) * 2 ) = mv_packed;
                A.costs[(size_t)pair * g.mb_count + xy] = cost;
            }
            int bcost_blk = cost + 4;
            if( want_intra )
            {
                if( ic < bcost_blk )
                {
                    bcost_blk = ic;
                    row_intra++;
                }
            }
            row_sum += bcost_blk;

            mv_right = mv_packed;
            mv_br = mv_b;
            mv_b = mv_bl;
            LA_TICK( LA_T_TAIL );
        }

        if( lane == 0 )
        {
            int32_t *s = A.sums + (size_t)pair * X264DSP_LA_SUMS;
            atomicAdd( &s[X264DSP_LA_COST_INTER], row_sum );
            atomicAdd( &s[X264DSP_LA_INTRA_MBS], row_intra );
            atomicAdd( &s[X264DSP_LA_SAD_EVALS], B.sad_evals );
            atomicAdd( &s[X264DSP_LA_SATD_EVALS], B.satd_evals );
            if( A.row_satds )
                A.row_satds[(size_t)pair * 2 * H + by] = row_sum;
            if( TIMED )
            {
                tacc[LA_T_BLOCKS] = W - 2;
                tacc[LA_T_ROWS] = 1;
                for( int k = 0; k < LA_T_KINDS; k++ )
                    atomicAdd( A.timing + k, (unsigned long long)tacc[k] );
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Multi-row inter kernel: the throughput mapping.
//
// One warp owns 32/LPB consecutive block rows of one frame pair; a group of LPB lanes walks one block
// row and each lane holds 8/LPB pixel rows of the group's current 8x8 block.  The groups run in lock
// step with the reference's own lag of two blocks per row (group sub is at column W-2-t+2*sub in step
// t), so inside a warp the row-to-row dependency is a register shuffle; only the bottom group polls the
// unit below (one {mv, epoch} word per step) and only the top group publishes.  Every lane evaluates
// ALL candidates of a search step for its pixel rows (VABSDIFF4 x2 per row and candidate); the
// per-candidate sums are packed two to a register and reduced over the lanes of a group with
// xor-shuffles, which serves 32/LPB blocks per instruction instead of one.
// SATD: each lane transforms its rows horizontally (x264's packed two-4x4s-per-word form,
// pixel.c:243-266), the vertical 4-point transform runs inside the lane and across lanes.
//
// The reference plane N is read through an 8x8-TILED copy (tile (tx,ty) = 64 contiguous bytes, tiles
// of a tile row back to back, built per launch by xd_la_tile_kernel).  With lane = pixel row(s), the
// lanes of a block read different rows: in the row-major plane that is one cache line per row and the
// L1 data pipe saturates (ncu: 91 % of peak wavefronts, DRAM 4 %); tiled, the same rows sit in one or
// two lines.
#define LQ_FULL 0xffffffffu

// 8 consecutive pixels at an arbitrary byte address of a row-major plane: two aligned 8-byte loads
__device__ __forceinline__ uint2 xd_lq_load8( const uint8_t *p )
{
    const uintptr_t a = (uintptr_t)p;
    const uint2 *w = (const uint2 *)( a & ~(uintptr_t)7 );
    const uint2 lo = __ldg( w ), hi = __ldg( w + 1 );
    const uint32_t sh = (uint32_t)a << 3;
    const bool up = ( (uint32_t)a & 4u ) != 0;
    const uint32_t w0 = up ? lo.y : lo.x, w1 = up ? hi.x : lo.y, w2 = up ? hi.y : hi.x;
    return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
}

// SAD of this lane's 8 pixels against its source row, accumulated onto acc (two VABSDIFF4.U8.ACC)
__device__ __forceinline__ uint32_t xd_lq_sad8( uint2 p, uint2 f, uint32_t acc )
{
    uint32_t t, d;
    asm( "vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"( t ) : "r"( p.x ), "r"( f.x ), "r"( acc ) );
    asm( "vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"( d ) : "r"( p.y ), "r"( f.y ), "r"( t ) );
    return d;
}

// two 16-bit sums in one word
__device__ __forceinline__ uint32_t xd_lq_pack( uint32_t lo, uint32_t hi )
{
    return __byte_perm( lo, hi, 0x5410 );
}

// p[k] | p[k+4] << 16 for k = 0..3
__device__ __forceinline__ void xd_lq_unpack( uint2 p, uint32_t w[4] )
{
    w[0] = __byte_perm( p.x, p.y, 0x7470 ) & 0x00FF00FFu;
    w[1] = __byte_perm( p.x, p.y, 0x7571 ) & 0x00FF00FFu;
    w[2] = __byte_perm( p.x, p.y, 0x7672 ) & 0x00FF00FFu;
    w[3] = __byte_perm( p.x, p.y, 0x7773 ) & 0x00FF00FFu;
}

// abs of both 16-bit halves of a word that holds hi * 65536 + lo with a signed lo (pixel.c:236-241)
__device__ __forceinline__ uint32_t xd_lq_abs2( uint32_t a )
{
    const uint32_t s = ( ( a >> 15 ) & 0x10001u ) * 0xffffu;
    return ( a + s ) ^ s;
}

// ---------------------------------------------------------------------------------------------
// LPB lanes per block (8: one pixel row per lane, four block rows per warp; 4: two pixel rows per
// lane, EIGHT block rows per warp).  Everything that is uniform inside a group --
// candidate coordinates, clips, medians, winner selection, the step bookkeeping -- is executed once per
// warp instruction whatever the group size, so halving the lanes per block halves that share of the
// instruction count per block; the per-pixel work (loads, VABSDIFF4, Hadamard) stays the same per block.
template<int RPL>
struct xd_lm_block
{
    const uint8_t *tref;      // tiled planes N, H, V, HV of the reference frame (in its slot)
    int tplane;               // bytes per tiled plane
    int tw;                   // tiles per tile row
    int X0, Y0;               // padded coordinates (x+32, y+32) of this lane's FIRST row of the block origin
    uint2 fenc[RPL];          // this lane's source rows
    uint32_t fw[RPL][4];      // the same rows as fw[k] = p[k] | p[k+4] << 16
    int mvpx, mvpy;
    const uint16_t *cost_mv;
};

template<int LPB>
__device__ __forceinline__ uint32_t xd_lm_reduce( uint32_t v )
{
#pragma unroll
    for( int o = 1; o < LPB; o <<= 1 )
        v += __shfl_xor_sync( LQ_FULL, v, o );
    return v;
}

template<int RPL>
__device__ __forceinline__ int xd_lm_bits( const xd_lm_block<RPL> &B, int qx, int qy )
{
    return __ldg( B.cost_mv + ( qx - B.mvpx ) ) + __ldg( B.cost_mv + ( qy - B.mvpy ) );
}

__device__ __forceinline__ uint2 xd_lm_assemble( uint2 lo, uint2 hi, uint32_t sh, bool up )
{
    const uint32_t w0 = up ? lo.y : lo.x, w1 = up ? hi.x : lo.y, w2 = up ? hi.y : hi.x;
    return make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
}

// LA_WORD_LOADS (a measured alternative, off): the 8 (or 12) bytes a lane needs from a tiled row start at an arbitrary byte,
// i.e. they are three (four) consecutive 32-bit words of the ROW -- which in the tiled plane sit at byte offsets {0,4,64,68}
// or {4,64,68,128} from the row chunk of the tile holding X, depending on which half of the chunk X falls in.  Loading those
// words one by one (offsets = two multiply-adds on the FMA pipe per fetch) needs no selects: SEL 229 -> 79 in the SASS of
// xd_la_multi_kernel<4>, LDG 127 -> 174.  Measured on B200 (bench.py --no-me, 1792 pairs per launch): 4.87 ms per launch
// against 4.23 ms with the 8-byte loads + three SELs per row (4.94 ms with the registers held at 80) -- the extra L1
// wavefronts cost more than the selects saved; the ALU pipe is not the only thing this kernel is short of.
#ifndef LA_WORD_LOADS
#define LA_WORD_LOADS 0
#endif

// the lane's RPL rows, 8 pixels each, at padded coordinates (X,Y) of the tiled plane at byte offset poff
template<int RPL>
__device__ __forceinline__ void xd_lm_tile( const xd_lm_block<RPL> &B, int poff, int X, int Y, uint2 out[RPL] )
{
    const int off = poff + ( ( ( Y >> 3 ) * B.tw + ( X >> 3 ) ) << 6 ) + ( ( Y & 7 ) << 3 );
    const uint32_t sh = (uint32_t)X << 3;
#if LA_WORD_LOADS
    const int up = ( X >> 2 ) & 1;
    const int o0 = up * 4, o1 = 4 + up * 60, o2 = 64 + up * 4;
    const uint8_t *w = B.tref + off;
    const uint32_t a0 = __ldg( (const uint32_t *)( w + o0 ) ), a1 = __ldg( (const uint32_t *)( w + o1 ) ),
                   a2 = __ldg( (const uint32_t *)( w + o2 ) );
    if( RPL == 2 )
    {
        const uint8_t *w1 = w + ( ( Y & 7 ) == 7 ? ( B.tw << 6 ) - 56 : 8 );
        const uint32_t b0 = __ldg( (const uint32_t *)( w1 + o0 ) ), b1 = __ldg( (const uint32_t *)( w1 + o1 ) ),
                       b2 = __ldg( (const uint32_t *)( w1 + o2 ) );
        out[RPL - 1] = make_uint2( __funnelshift_r( b0, b1, sh ), __funnelshift_r( b1, b2, sh ) );
    }
    out[0] = make_uint2( __funnelshift_r( a0, a1, sh ), __funnelshift_r( a1, a2, sh ) );
#else
    const bool up = ( X & 4 ) != 0;
    const uint2 *w = (const uint2 *)( B.tref + off );
    const uint2 lo = __ldg( w ), hi = __ldg( w + 8 );
    if( RPL == 2 )
    {
        // the next row: 8 bytes further in the same tile, or the first row of the tile below
        const uint2 *w1 = (const uint2 *)( B.tref + ( off + ( ( Y & 7 ) == 7 ? ( B.tw << 6 ) - 56 : 8 ) ) );
        const uint2 lo1 = __ldg( w1 ), hi1 = __ldg( w1 + 8 );
        out[RPL - 1] = xd_lm_assemble( lo1, hi1, sh, up );
    }
    out[0] = xd_lm_assemble( lo, hi, sh, up );
#endif
}

// the lane's rows of the prediction at quarter-pel (qx,qy): get_ref / mc_luma (mc.c:192-264)
template<int RPL>
__device__ __forceinline__ void xd_lm_fetch( const xd_lm_block<RPL> &B, int qx, int qy, uint2 out[RPL] )
{
    const int fx = qx & 3, fy = qy & 3, phase = fy * 4 + fx;
    const int X = B.X0 + ( qx >> 2 ), Y = B.Y0 + ( qy >> 2 );
    xd_lm_tile<RPL>( B, xd_qpel_plane_a( phase ) * B.tplane, X, Y + ( fy == 3 ? 1 : 0 ), out );
    if( phase & 5 )
    {
        uint2 b[RPL];
        xd_lm_tile<RPL>( B, xd_qpel_plane_b( phase ) * B.tplane, X + ( fx == 3 ? 1 : 0 ), Y, b );
#pragma unroll
        for( int j = 0; j < RPL; j++ )
        {
            out[j].x = xd_avg4( out[j].x, b[j].x );
            out[j].y = xd_avg4( out[j].y, b[j].y );
        }
    }
}

template<int RPL>
__device__ __forceinline__ uint32_t xd_lm_sad( const xd_lm_block<RPL> &B, const uint2 p[RPL] )
{
    uint32_t acc = 0;
#pragma unroll
    for( int j = 0; j < RPL; j++ )
        acc = xd_lq_sad8( p[j], B.fenc[j], acc );
    return acc;
}

template<int RPL>
__device__ __forceinline__ uint32_t xd_lm_sad_fpel( const xd_lm_block<RPL> &B, int mx, int my )
{
    uint2 p[RPL];
    xd_lm_tile<RPL>( B, 0, B.X0 + mx, B.Y0 + my, p );
    return xd_lm_sad<RPL>( B, p );
}

// The four neighbours of a diamond step (me.c:237-274) from ONE window of the reference plane: the lane's two rows
// moved up / down / left / right by one pixel all lie inside rows Y-1 .. Y+2, columns X-1 .. X+8.  Four separate
// fetches cost four address computations and sixteen 8-byte loads per lane; the window costs one address computation
// and eight loads (a third tile column only when the window starts on a tile's last byte), and the candidates are
// funnel shifts of the window's words by constant amounts.
template<int RPL>
__device__ __forceinline__ void xd_lm_sad_dia( const xd_lm_block<RPL> &B, int mx, int my,
                                               uint32_t &up, uint32_t &dn, uint32_t &lf, uint32_t &rt )
{
    if constexpr( RPL != 2 )
    {
        up = xd_lm_sad_fpel<RPL>( B, mx, my - 1 );
        dn = xd_lm_sad_fpel<RPL>( B, mx, my + 1 );
        lf = xd_lm_sad_fpel<RPL>( B, mx - 1, my );
        rt = xd_lm_sad_fpel<RPL>( B, mx + 1, my );
    }
    else
    {
        const int XW = B.X0 + mx - 1, R0 = B.Y0 + my - 1;        // window origin: column X-1, row Y-1
        int off = ( ( ( R0 >> 3 ) * B.tw + ( XW >> 3 ) ) << 6 ) + ( ( R0 & 7 ) << 3 );
        const int o = XW & 7;
        const uint32_t sh = (uint32_t)o << 3;                      // the funnel shift takes it modulo 32
        const bool hi4 = ( o & 4 ) != 0, third = o == 7;
        const int jump = ( B.tw << 6 ) - 56;                       // from row 7 of a tile to row 0 of the tile below
        uint32_t V[4][3];                                          // bytes X-1 .. X+10 of the four rows
#if LA_WORD_LOADS
        const int hw = hi4 ? 1 : 0;
        const int o0 = hw * 4, o1 = 4 + hw * 60, o2 = 64 + hw * 4, o3 = 68 + hw * 60;
        const bool fourth = ( o & 3 ) == 3;                        // only then do bytes of a fourth word reach columns X-1 .. X+8
        (void)third;
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint8_t *w = B.tref + off;
            const uint32_t w0 = __ldg( (const uint32_t *)( w + o0 ) ), w1 = __ldg( (const uint32_t *)( w + o1 ) ),
                           w2 = __ldg( (const uint32_t *)( w + o2 ) );
            uint32_t w3 = 0;
            if( fourth )
                w3 = __ldg( (const uint32_t *)( w + o3 ) );
            V[r][0] = __funnelshift_r( w0, w1, sh );
            V[r][1] = __funnelshift_r( w1, w2, sh );
            V[r][2] = __funnelshift_r( w2, w3, sh );
            off += ( ( R0 + r ) & 7 ) == 7 ? jump : 8;
        }
#else
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint2 *w = (const uint2 *)( B.tref + off );
            const uint2 lo = __ldg( w ), hi = __ldg( w + 8 );
            uint32_t ex = 0;
            if( third )
                ex = __ldg( (const uint32_t *)( w + 16 ) );
            const uint32_t w0 = hi4 ? lo.y : lo.x, w1 = hi4 ? hi.x : lo.y, w2 = hi4 ? hi.y : hi.x, w3 = hi4 ? ex : hi.y;
            V[r][0] = __funnelshift_r( w0, w1, sh );
            V[r][1] = __funnelshift_r( w1, w2, sh );
            V[r][2] = __funnelshift_r( w2, w3, sh );
            off += ( ( R0 + r ) & 7 ) == 7 ? jump : 8;
        }
#endif
        uint2 c[4];                                                // the rows at column X
#pragma unroll
        for( int r = 0; r < 4; r++ )
            c[r] = make_uint2( __funnelshift_r( V[r][0], V[r][1], 8 ), __funnelshift_r( V[r][1], V[r][2], 8 ) );
        up = xd_lq_sad8( c[1], B.fenc[1], xd_lq_sad8( c[0], B.fenc[0], 0u ) );
        dn = xd_lq_sad8( c[3], B.fenc[1], xd_lq_sad8( c[2], B.fenc[0], 0u ) );
        lf = xd_lq_sad8( make_uint2( V[2][0], V[2][1] ), B.fenc[1], xd_lq_sad8( make_uint2( V[1][0], V[1][1] ), B.fenc[0], 0u ) );
        const uint2 r1 = make_uint2( __funnelshift_r( V[1][0], V[1][1], 16 ), __funnelshift_r( V[1][1], V[1][2], 16 ) );
        const uint2 r2 = make_uint2( __funnelshift_r( V[2][0], V[2][1], 16 ), __funnelshift_r( V[2][1], V[2][2], 16 ) );
        rt = xd_lq_sad8( r2, B.fenc[1], xd_lq_sad8( r1, B.fenc[0], 0u ) );
    }
}

// The four neighbours of the half-pel diamond (me.c:492-517) around (qx,qy).  In the lookahead every position is on the
// half-pel grid -- the search is full-pel, the refinement half-pel only (subme 2), the predictors are medians of such
// vectors -- so each candidate is one plane, no averaging: up / down are three consecutive rows of the plane that has
// the other vertical phase, left / right two rows and nine columns of the plane with the other horizontal phase.
// Anything else (never seen here, but the reference allows it) takes the general path.
template<int RPL>
__device__ __forceinline__ void xd_lm_sad_hpel_dia( const xd_lm_block<RPL> &B, int qx, int qy,
                                                    uint32_t &up, uint32_t &dn, uint32_t &lf, uint32_t &rt );

template<int RPL>
__device__ __forceinline__ uint32_t xd_lm_sad_qpel( const xd_lm_block<RPL> &B, int qx, int qy )
{
    uint2 p[RPL];
    xd_lm_fetch<RPL>( B, qx, qy, p );
    return xd_lm_sad<RPL>( B, p );
}

template<int RPL>
__device__ __forceinline__ void xd_lm_sad_hpel_dia( const xd_lm_block<RPL> &B, int qx, int qy,
                                                    uint32_t &up, uint32_t &dn, uint32_t &lf, uint32_t &rt )
{
    if( RPL != 2 || ( ( qx | qy ) & 1 ) )
    {
        up = xd_lm_sad_qpel<RPL>( B, qx, qy - 2 );
        dn = xd_lm_sad_qpel<RPL>( B, qx, qy + 2 );
        lf = xd_lm_sad_qpel<RPL>( B, qx - 2, qy );
        rt = xd_lm_sad_qpel<RPL>( B, qx + 2, qy );
        return;
    }
    if constexpr( RPL == 2 )
    {
        const int hx = ( qx >> 1 ) & 1, hy = ( qy >> 1 ) & 1;        // half-pel phase bits: plane = hx | hy << 1
        const int jump = ( B.tw << 6 ) - 56;
        {
            // up / down: plane with the other vertical phase, rows R0 .. R0+2 at column X
            const int X = B.X0 + ( qx >> 2 ), R0 = B.Y0 + ( ( qy - 2 ) >> 2 );
            int off = ( hx | ( ( hy ^ 1 ) << 1 ) ) * B.tplane + ( ( ( R0 >> 3 ) * B.tw + ( X >> 3 ) ) << 6 ) + ( ( R0 & 7 ) << 3 );
            const uint32_t sh = (uint32_t)X << 3;
            const bool hi4 = ( X & 4 ) != 0;
            uint2 c[3];
#if LA_WORD_LOADS
            const int hw = hi4 ? 1 : 0;
            const int o0 = hw * 4, o1 = 4 + hw * 60, o2 = 64 + hw * 4;
#pragma unroll
            for( int r = 0; r < 3; r++ )
            {
                const uint8_t *w = B.tref + off;
                const uint32_t w0 = __ldg( (const uint32_t *)( w + o0 ) ), w1 = __ldg( (const uint32_t *)( w + o1 ) ),
                               w2 = __ldg( (const uint32_t *)( w + o2 ) );
                c[r] = make_uint2( __funnelshift_r( w0, w1, sh ), __funnelshift_r( w1, w2, sh ) );
                off += ( ( R0 + r ) & 7 ) == 7 ? jump : 8;
            }
#else
#pragma unroll
            for( int r = 0; r < 3; r++ )
            {
                const uint2 *w = (const uint2 *)( B.tref + off );
                c[r] = xd_lm_assemble( __ldg( w ), __ldg( w + 8 ), sh, hi4 );
                off += ( ( R0 + r ) & 7 ) == 7 ? jump : 8;
            }
#endif
            up = xd_lq_sad8( c[1], B.fenc[1], xd_lq_sad8( c[0], B.fenc[0], 0u ) );
            dn = xd_lq_sad8( c[2], B.fenc[1], xd_lq_sad8( c[1], B.fenc[0], 0u ) );
        }
        {
            // left / right: plane with the other horizontal phase, rows Y, Y+1, columns XL .. XL+8
            const int XL = B.X0 + ( ( qx - 2 ) >> 2 ), Y = B.Y0 + ( qy >> 2 );
            int off = ( ( hx ^ 1 ) | ( hy << 1 ) ) * B.tplane + ( ( ( Y >> 3 ) * B.tw + ( XL >> 3 ) ) << 6 ) + ( ( Y & 7 ) << 3 );
            const uint32_t sh = (uint32_t)XL << 3;
            const bool hi4 = ( XL & 4 ) != 0;
            uint32_t V[2][3];
#if LA_WORD_LOADS
            // columns XL .. XL+8: nine bytes from byte (XL & 3) of the first word, so a fourth word never contributes
            const int hw = hi4 ? 1 : 0;
            const int o0 = hw * 4, o1 = 4 + hw * 60, o2 = 64 + hw * 4;
#pragma unroll
            for( int r = 0; r < 2; r++ )
            {
                const uint8_t *w = B.tref + off;
                const uint32_t w0 = __ldg( (const uint32_t *)( w + o0 ) ), w1 = __ldg( (const uint32_t *)( w + o1 ) ),
                               w2 = __ldg( (const uint32_t *)( w + o2 ) );
                V[r][0] = __funnelshift_r( w0, w1, sh );
                V[r][1] = __funnelshift_r( w1, w2, sh );
                V[r][2] = __funnelshift_r( w2, 0u, sh );
                off += ( Y & 7 ) == 7 ? jump : 8;
            }
#else
#pragma unroll
            for( int r = 0; r < 2; r++ )
            {
                const uint2 *w = (const uint2 *)( B.tref + off );
                const uint2 lo = __ldg( w ), hi = __ldg( w + 8 );
                const uint32_t w0 = hi4 ? lo.y : lo.x, w1 = hi4 ? hi.x : lo.y, w2 = hi4 ? hi.y : hi.x, w3 = hi4 ? 0u : hi.y;
                V[r][0] = __funnelshift_r( w0, w1, sh );
                V[r][1] = __funnelshift_r( w1, w2, sh );
                V[r][2] = __funnelshift_r( w2, w3, sh );
                off += ( Y & 7 ) == 7 ? jump : 8;
            }
#endif
            lf = xd_lq_sad8( make_uint2( V[1][0], V[1][1] ), B.fenc[1], xd_lq_sad8( make_uint2( V[0][0], V[0][1] ), B.fenc[0], 0u ) );
            const uint2 r0 = make_uint2( __funnelshift_r( V[0][0], V[0][1], 8 ), __funnelshift_r( V[0][1], V[0][2], 8 ) );
            const uint2 r1 = make_uint2( __funnelshift_r( V[1][0], V[1][1], 8 ), __funnelshift_r( V[1][1], V[1][2], 8 ) );
            rt = xd_lq_sad8( r1, B.fenc[1], xd_lq_sad8( r0, B.fenc[0], 0u ) );
        }
    }
}

// SATD 8x8 (pixel.c:294-335); whole warp executes, `on` gates the loads, every lane of a group gets the cost.
// q = lane inside the group; the lane holds pixel rows q*RPL .. q*RPL+RPL-1.
template<int LPB>
__device__ __forceinline__ int xd_lm_satd( const xd_lm_block<8 / LPB> &B, int qx, int qy, int q, bool on )
{
    constexpr int RPL = 8 / LPB;
    uint2 p[RPL];
#pragma unroll
    for( int j = 0; j < RPL; j++ )
        p[j] = make_uint2( 0u, 0u );
    if( on )
        xd_lm_fetch<RPL>( B, qx, qy, p );
    uint32_t h[RPL][4];
#pragma unroll
    for( int j = 0; j < RPL; j++ )
    {
        uint32_t w[4];
        xd_lq_unpack( p[j], w );
        const uint32_t a0 = B.fw[j][0] - w[0], a1 = B.fw[j][1] - w[1], a2 = B.fw[j][2] - w[2], a3 = B.fw[j][3] - w[3];
        const uint32_t t0 = a0 + a1, t1 = a0 - a1, t2 = a2 + a3, t3 = a2 - a3;
        h[j][0] = t0 + t2; h[j][1] = t1 + t3; h[j][2] = t0 - t2; h[j][3] = t1 - t3;
    }
    const uint32_t n1 = ( q & 1 ) ? ~0u : 0u, c1 = q & 1;
    uint32_t sum = 0;
    if( RPL == 1 )
    {
        const uint32_t n2 = ( q & 2 ) ? ~0u : 0u, c2 = ( q >> 1 ) & 1;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t v = h[0][k];
            v = __shfl_xor_sync( LQ_FULL, v, 1 ) + ( v ^ n1 ) + c1;
            v = __shfl_xor_sync( LQ_FULL, v, 2 ) + ( v ^ n2 ) + c2;
            sum += xd_lq_abs2( v );
        }
        sum += __shfl_xor_sync( LQ_FULL, sum, 1 );
        sum += __shfl_xor_sync( LQ_FULL, sum, 2 );
        const uint32_t half = ( ( sum & 0xFFFFu ) + ( sum >> 16 ) ) >> 1;
        return (int)( half + __shfl_xor_sync( LQ_FULL, half, 4 ) );
    }
    else
    {
        // rows (2q, 2q+1) of an 8x4 live in this lane, the other two in lane q^1
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t sk = h[0][k] + h[RPL - 1][k], dk = h[0][k] - h[RPL - 1][k];
            sk = __shfl_xor_sync( LQ_FULL, sk, 1 ) + ( sk ^ n1 ) + c1;
            dk = __shfl_xor_sync( LQ_FULL, dk, 1 ) + ( dk ^ n1 ) + c1;
            sum += xd_lq_abs2( sk ) + xd_lq_abs2( dk );
        }
        sum += __shfl_xor_sync( LQ_FULL, sum, 1 );
        const uint32_t half = ( ( sum & 0xFFFFu ) + ( sum >> 16 ) ) >> 1;
        return (int)( half + __shfl_xor_sync( LQ_FULL, half, 2 ) );
    }
}

template<int LPB>
__global__ void __launch_bounds__( LA_WARPS * 32 )
xd_la_multi_kernel( xd_la_args A, int n_inter, const int32_t *inter_pairs )
{
    constexpr int RPL = 8 / LPB, GROUPS = 32 / LPB;
    const x264dsp_geom_t &g = A.g;
    const int lane = threadIdx.x & 31, sub = lane / LPB, q = lane % LPB;
    const int W = g.mb_w, H = g.mb_h, ls = g.lowres_stride;
    const int rows = H - 2, units = ( rows + GROUPS - 1 ) / GROUPS;
    const int total = n_inter * units;

    for( ;; )
    {
        int ticket = 0;
        if( lane == 0 )
            ticket = atomicAdd( A.ticket, 1 );
        ticket = __shfl_sync( LQ_FULL, ticket, 0 );
        if( ticket >= total )
            return;
        // unit-major over the pairs of the launch: the unit a warp waits on always holds a smaller
        // ticket, i.e. is running or finished
        const int unit = ticket / n_inter;
        const int pair = inter_pairs[ticket - unit * n_inter];
        const int by0 = H - 2 - GROUPS * unit;            // bottom row of the unit
        const int nrows = min( GROUPS, by0 );             // rows by0, by0-1, ... down to row 1
        const int by = by0 - sub;
        const bool row_ok = sub < nrows;
        const bool has_below = unit > 0;
        const bool is_top = sub == nrows - 1;
        // the source block (bx,by) is exactly tile (bx+4, by+4) of the current frame's tiled plane N:
        // 64 contiguous bytes, this lane's rows at q*RPL*8
        const uint8_t *cur = A.slots + (size_t)A.b[pair] * g.slot_bytes + g.slot_tiled_off
                           + ( (size_t)( by + 4 ) * g.tile_w + 4 ) * 64 + q * RPL * 8;
        unsigned long long *sync_row = A.sync + (size_t)pair * g.mb_count + (size_t)by * W;
        const unsigned long long *sync_below = A.sync + (size_t)pair * g.mb_count + (size_t)( by0 + 1 ) * W;
        const bool want_intra = A.want_intra[pair] != 0;
        const int32_t *icost_row = A.icost + (size_t)pair * g.mb_count + (size_t)by * W;

        // MV limits (slicetype.c:79-89); the y limits are per row
        const int miny = -( by << 3 ) - 4, maxy = ( ( H - by - 1 ) << 3 ) + 4;
        const int sminy = ( miny - 8 ) << 2, smaxy = ( maxy + 8 ) << 2;

        xd_lm_block<RPL> B;
        B.tw = g.tile_w;
        B.tplane = g.tiled_plane_size;
        B.tref = A.slots + (size_t)A.p0[pair] * g.slot_bytes + g.slot_tiled_off;
        B.X0 = B.Y0 = 32;
        B.cost_mv = A.cost_mv;
        B.mvpx = B.mvpy = 0;
#pragma unroll
        for( int j = 0; j < RPL; j++ )
        {
            B.fenc[j] = make_uint2( 0u, 0u );
            B.fw[j][0] = B.fw[j][1] = B.fw[j][2] = B.fw[j][3] = 0u;
        }
        int sad_evals = 0, satd_evals = 0, row_sum = 0, row_intra = 0;

        uint32_t mv_right = 0, mv_b = 0, mv_br = 0, mv_bl = 0;
        uint32_t last_mv = 0;                              // this group's previous result (0 when it had none)
        unsigned long long pending = 0;
        // step t = -1 only shifts the neighbour pipeline: it brings (W-2, below) in
        if( has_below )
        {
            const int far = max( W - 2 - A.slack, 1 );
            xd_la_await( sync_below + far, xd_ld_sync( sync_below + far ), A.epoch, 500, 8000 );
            pending = xd_ld_sync( sync_below + ( W - 2 ) );
        }

        // the first block's source rows, fetched ahead like every later one
        uint2 nx_fenc[RPL];
        int nx_ic = 0;
#pragma unroll
        for( int j = 0; j < RPL; j++ )
            nx_fenc[j] = make_uint2( 0u, 0u );
        if( row_ok )
        {
#pragma unroll
            for( int j = 0; j < RPL; j++ )
                nx_fenc[j] = __ldg( (const uint2 *)( cur + (size_t)( W - 2 ) * 64 + j * 8 ) );
            if( want_intra )
                nx_ic = icost_row[W - 2];
        }

        const int n_steps = ( W - 2 ) + 2 * ( nrows - 1 );
        for( int t = -1; t < n_steps; t++ )
        {
            const int bx = W - 2 - t + 2 * sub;
            const bool act = row_ok && bx >= 1 && bx <= W - 2;

            // ---- neighbour of the row below at column bx-1: the group underneath produced it in the
            //      previous step; the bottom group reads it from the unit below
            uint32_t incoming = __shfl_up_sync( LQ_FULL, last_mv, LPB );
            {
                const int bx0 = W - 2 - t;                 // the bottom group's column
                uint32_t polled = 0;
                if( has_below && bx0 - 1 >= 1 && bx0 - 1 <= W - 2 )
                {
                    polled = xd_la_await( sync_below + ( bx0 - 1 ), pending, A.epoch, t < 0 ? 500 : 100, t < 0 ? 8000 : 1000 );
                    if( bx0 - 2 >= 1 )
                        pending = xd_ld_sync( sync_below + ( bx0 - 2 ) );
                }
                if( sub == 0 )
                    incoming = polled;
            }
            mv_bl = incoming;

            int mvx = 0, mvy = 0, cost = 0;
            int ic = 0;
            bool search = false;
            int minx = 0, maxx = 0;
            if( act )
            {
                B.X0 = 8 * bx + 32;
                B.Y0 = 8 * by + q * RPL + 32;
                ic = nx_ic;
#pragma unroll
                for( int j = 0; j < RPL; j++ )
                    B.fenc[j] = nx_fenc[j];
                if( bx > 1 )
                {
#pragma unroll
                    for( int j = 0; j < RPL; j++ )
                        nx_fenc[j] = __ldg( (const uint2 *)( cur + (size_t)( bx - 1 ) * 64 + j * 8 ) );
                    if( want_intra )
                        nx_ic = icost_row[bx - 1];
                }
#pragma unroll
                for( int j = 0; j < RPL; j++ )
                    xd_lq_unpack( B.fenc[j], B.fw[j] );
                minx = -( bx << 3 ) - 4;
                maxx = ( ( W - bx - 1 ) << 3 ) + 4;
                // predictors (slicetype.c:105-113): right, below, below-left (below-right is a candidate only)
                B.mvpx = xd_median3( MVX( mv_right ), MVX( mv_b ), MVX( mv_bl ) );
                B.mvpy = xd_median3( MVY( mv_right ), MVY( mv_b ), MVY( mv_bl ) );
                search = true;
            }

            // ---- slicetype.c:117-125: zero predictor -> try the zero vector with SATD first
            const bool needz = act && !( B.mvpx | B.mvpy );
            if( __any_sync( LQ_FULL, needz ) )
            {
                const int c0 = xd_lm_satd<LPB>( B, 0, 0, q, needz );
                if( needz )
                {
                    satd_evals++;
                    if( c0 < 64 )
                    {
                        cost = c0;
                        search = false;
                    }
                }
            }

            int bmx = 0, bmy = 0, pmx = 0, pmy = 0, bcost = 0;
            // ---- x264_me_search_ref, subme < 3 branch (me.c:194-233): rounded MVP (no mv cost),
            //      the four neighbour MVs, then (0,0)
            if( __any_sync( LQ_FULL, search ) )
            {
                int ccx[6], ccy[6];
                bool cok[6];
                uint32_t pmv = 0;
                if( search )
                {
                    bmx = xd_clip3( B.mvpx, minx * 4, maxx * 4 );
                    bmy = xd_clip3( B.mvpy, miny * 4, maxy * 4 );
                    pmx = ( bmx + 2 ) >> 2;
                    pmy = ( bmy + 2 ) >> 2;
                    pmv = ( (uint32_t)pmx & 0xFFFF ) | ( (uint32_t)pmy << 16 );
                }
                ccx[0] = pmx; ccy[0] = pmy; cok[0] = search;
#pragma unroll
                for( int k = 1; k <= 4; k++ )
                {
                    const uint32_t m = k == 1 ? mv_right : k == 2 ? mv_b : k == 3 ? mv_bl : mv_br;
                    ccx[k] = xd_clip3( ( MVX( m ) + 2 ) >> 2, minx, maxx );
                    ccy[k] = xd_clip3( ( MVY( m ) + 2 ) >> 2, miny, maxy );
                    const uint32_t v = ( (uint32_t)ccx[k] & 0xFFFF ) | ( (uint32_t)ccy[k] << 16 );
                    cok[k] = search && v != 0 && v != pmv;
                }
                ccx[5] = 0; ccy[5] = 0; cok[5] = search && pmv != 0;

                uint32_t w[3] = { 0u, 0u, 0u };
#pragma unroll
                for( int k = 0; k < 6; k++ )
                {
                    uint32_t s = 0;
                    if( cok[k] )
                        s = xd_lm_sad_fpel<RPL>( B, ccx[k], ccy[k] );
                    w[k >> 1] += s << ( 16 * ( k & 1 ) );
                }
                // lane q adds the mv bits of candidates q, q+LPB, ... or the "does not compete" marker
#pragma unroll
                for( int rep = 0; rep < ( 6 + LPB - 1 ) / LPB; rep++ )
                {
                    const int kk = q + rep * LPB;
                    if( search && kk < 6 )
                    {
                        int kx = ccx[0], ky = ccy[0];
                        bool ok = cok[0];
#pragma unroll
                        for( int k = 1; k < 6; k++ )
                            if( kk == k ) { kx = ccx[k]; ky = ccy[k]; ok = cok[k]; }
                        const uint32_t add = !ok ? 0xFFFFu : kk ? (uint32_t)xd_lm_bits<RPL>( B, kx << 2, ky << 2 ) : 0u;
                        const uint32_t sh = add << ( 16 * ( kk & 1 ) );
                        if( ( kk >> 1 ) == 0 ) w[0] += sh;
                        else if( ( kk >> 1 ) == 1 ) w[1] += sh;
                        else w[2] += sh;
                    }
                }
                w[0] = xd_lm_reduce<LPB>( w[0] );
                w[1] = xd_lm_reduce<LPB>( w[1] );
                w[2] = xd_lm_reduce<LPB>( w[2] );
                if( search )
                {
                    int best = 0x7FFFFFFF;
#pragma unroll
                    for( int k = 0; k < 6; k++ )
                    {
                        const int c = (int)( ( w[k >> 1] >> ( 16 * ( k & 1 ) ) ) & 0xFFFFu );
                        if( c < 0xFFFF )
                        {
                            best = min( best, ( c << 3 ) | k );
                            sad_evals++;
                        }
                    }
                    bcost = best >> 3;
                    const int win = best & 7;
#pragma unroll
                    for( int k = 0; k < 6; k++ )
                        if( win == k ) { bmx = ccx[k]; bmy = ccy[k]; }
                }
            }

            // ---- diamond search (me.c:237-274): up, down, left, right
            {
                bool dia = search;
                int left = A.me_range;
                while( __any_sync( LQ_FULL, dia ) )
                {
                    uint32_t w0 = 0, w1 = 0;
                    if( dia )
                    {
                        uint32_t su, sd, sl, sr;
                        xd_lm_sad_dia<RPL>( B, bmx, bmy, su, sd, sl, sr );
                        w0 = xd_lq_pack( su, sd );
                        w1 = xd_lq_pack( sl, sr );
                        if( q < 4 )
                        {
                            const int dx = q == 2 ? -1 : q == 3 ? 1 : 0, dy = q == 0 ? -1 : q == 1 ? 1 : 0;
                            const uint32_t sh = (uint32_t)xd_lm_bits<RPL>( B, ( bmx + dx ) << 2, ( bmy + dy ) << 2 ) << ( 16 * ( q & 1 ) );
                            if( q < 2 ) w0 += sh; else w1 += sh;
                        }
                    }
                    w0 = xd_lm_reduce<LPB>( w0 );
                    w1 = xd_lm_reduce<LPB>( w1 );
                    if( dia )
                    {
                        sad_evals += 4;
                        const int k0 = (int)( ( w0 & 0xFFFFu ) << 2 ), k1 = (int)( ( w0 >> 16 ) << 2 ) | 1;
                        const int k2 = (int)( ( w1 & 0xFFFFu ) << 2 ) | 2, k3 = (int)( ( w1 >> 16 ) << 2 ) | 3;
                        const int key = min( min( k0, k1 ), min( k2, k3 ) );
                        if( ( key >> 2 ) >= bcost )
                            dia = false;
                        else
                        {
                            bcost = key >> 2;
                            const int wn = key & 3;
                            bmx += wn == 2 ? -1 : wn == 3 ? 1 : 0;
                            bmy += wn == 0 ? -1 : wn == 1 ? 1 : 0;
                            if( !--left || !xd_la_in_range( bmx, bmy, minx, miny, maxx, maxy ) )
                                dia = false;
                        }
                    }
                }
            }

            // ---- me.c:397-414, then refine_subpel( hpel_iters = 1, qpel_iters = 0 ) (me.c:466-587)
            if( __any_sync( LQ_FULL, search ) )
            {
                int qx = bmx << 2, qy = bmy << 2, px = 0, py = 0;
                bool single = false;
                if( search )
                {
                    if( bmx == pmx && bmy == pmy )
                        bcost += xd_lm_bits<RPL>( B, qx, qy );
                    const int sminx = ( minx - 8 ) << 2, smaxx = ( maxx + 8 ) << 2;
                    px = xd_clip3( B.mvpx, sminx + 2, smaxx - 2 );
                    py = xd_clip3( B.mvpy, sminy + 2, smaxy - 2 );
                    single = px != qx || py != qy;                   // me.c:483-490
                }
                if( __any_sync( LQ_FULL, single ) )
                {
                    uint32_t s = 0;
                    if( single )
                    {
                        s = xd_lm_sad_qpel<RPL>( B, px, py );
                        if( q == 0 )
                            s += (uint32_t)xd_lm_bits<RPL>( B, px, py );
                    }
                    s = xd_lm_reduce<LPB>( s );
                    if( single )
                    {
                        sad_evals++;
                        if( (int)s < bcost ) { bcost = (int)s; qx = px; qy = py; }
                    }
                }
                // half-pel diamond (me.c:492-517)
                uint32_t w0 = 0, w1 = 0;
                if( search )
                {
                    uint32_t su, sd, sl, sr;
                    xd_lm_sad_hpel_dia<RPL>( B, qx, qy, su, sd, sl, sr );
                    w0 = xd_lq_pack( su, sd );
                    w1 = xd_lq_pack( sl, sr );
                    if( q < 4 )
                    {
                        const int dx = q == 2 ? -2 : q == 3 ? 2 : 0, dy = q == 0 ? -2 : q == 1 ? 2 : 0;
                        const uint32_t sh = (uint32_t)xd_lm_bits<RPL>( B, qx + dx, qy + dy ) << ( 16 * ( q & 1 ) );
                        if( q < 2 ) w0 += sh; else w1 += sh;
                    }
                }
                w0 = xd_lm_reduce<LPB>( w0 );
                w1 = xd_lm_reduce<LPB>( w1 );
                if( search )
                {
                    sad_evals += 4;
                    const int k0 = (int)( ( w0 & 0xFFFFu ) << 2 ), k1 = (int)( ( w0 >> 16 ) << 2 ) | 1;
                    const int k2 = (int)( ( w1 & 0xFFFFu ) << 2 ) | 2, k3 = (int)( ( w1 >> 16 ) << 2 ) | 3;
                    const int key = min( min( k0, k1 ), min( k2, k3 ) );
                    if( ( key >> 2 ) < bcost )
                    {
                        const int wn = key & 3;
                        qx += wn == 2 ? -2 : wn == 3 ? 2 : 0;
                        qy += wn == 0 ? -2 : wn == 1 ? 2 : 0;
                    }
                }
                // me.c:519-524: the winner is re-costed with SATD
                const int c = xd_lm_satd<LPB>( B, qx, qy, q, search );
                if( search )
                {
                    satd_evals++;
                    mvx = qx; mvy = qy;
                    cost = c + xd_lm_bits<RPL>( B, qx, qy ) - 1;      // slicetype.c:128-130
                    if( mvx | mvy )
                        cost += 5;
                }
            }

            // ---- publish, then account (slicetype.c:132-196)
            const uint32_t mv_packed = ( (uint32_t)mvx & 0xFFFF ) | ( (uint32_t)mvy << 16 );
            if( act )
            {
                if( q == 0 )
                {
                    const int xy = by * W + bx;
                    if( is_top )
                        xd_st_sync( sync_row + bx, ( (unsigned long long)A.epoch << 32 ) | mv_packed );
                    *(uint32_t *)( A.mvs + ( (size_t)pair * g.mb_count + xy ) * 2 ) = mv_packed;
                    A.costs[(size_t)pair * g.mb_count + xy] = cost;
                }
                int bcost_blk = cost + 4;
                if( want_intra && ic < bcost_blk )
                {
                    bcost_blk = ic;
                    row_intra++;
                }
                row_sum += bcost_blk;
                mv_right = mv_packed;
                last_mv = mv_packed;
            }
            else
                last_mv = 0;
            mv_br = mv_b;
            mv_b = mv_bl;
        }

        if( row_ok && q == 0 )
        {
            int32_t *s = A.sums + (size_t)pair * X264DSP_LA_SUMS;
            atomicAdd( &s[X264DSP_LA_COST_INTER], row_sum );
            atomicAdd( &s[X264DSP_LA_INTRA_MBS], row_intra );
            atomicAdd( &s[X264DSP_LA_SAD_EVALS], sad_evals );
            atomicAdd( &s[X264DSP_LA_SATD_EVALS], satd_evals );
            if( A.row_satds )
                A.row_satds[(size_t)pair * 2 * H + by] = row_sum;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side

// scratch for `n_pairs` pairs; called once per batch before any group is launched
static int xd_la_prepare( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, int n_pairs, cudaStream_t s )
{
    const size_t blocks = (size_t)n_pairs * g->mb_count;
    if( ctx->la_sync_cap < blocks * sizeof( unsigned long long ) )
    {
        XD_CHECK( cudaDeviceSynchronize() );
        int rc = xd_reserve_dev( (void **)&ctx->la_sync, &ctx->la_sync_cap, blocks * sizeof( unsigned long long ) );
        if( rc )
            return rc;
        XD_CHECK( cudaMemsetAsync( ctx->la_sync, 0, ctx->la_sync_cap, s ) );
        XD_CHECK( cudaStreamSynchronize( s ) );
        ctx->la_epoch = 0;
    }
    if( ctx->la_icost_cap < blocks * sizeof( int32_t ) )
    {
        XD_CHECK( cudaDeviceSynchronize() );
        int rc = xd_reserve_dev( (void **)&ctx->la_icost, &ctx->la_icost_cap, blocks * sizeof( int32_t ) );
        if( rc )
            return rc;
    }
    if( ++ctx->la_epoch == 0 )
    {
        XD_CHECK( cudaMemsetAsync( ctx->la_sync, 0, ctx->la_sync_cap, s ) );
        XD_CHECK( cudaStreamSynchronize( s ) );
        ctx->la_epoch = 1;
    }
    return 0;
}

// analyses pairs [pair0, pair0+count) on stream s; `inter_list` holds the global indices of the
// pairs of that range that have a reference frame; `slot` selects one of the ticket counters
static int xd_la_launch( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *slots,
                         int pair0, int count, const int32_t *b_dev, const int32_t *p0_dev, const uint8_t *wi_dev,
                         const int32_t *inter_list, int n_inter, int ticket_slot,
                         int16_t *mvs, int32_t *costs, int32_t *sums, int32_t *row_satds, cudaStream_t s )
{
    const size_t mbc = g->mb_count;
    XD_CHECK( cudaMemsetAsync( mvs + (size_t)pair0 * mbc * 2, 0, (size_t)count * mbc * 2 * sizeof( int16_t ), s ) );
    XD_CHECK( cudaMemsetAsync( costs + (size_t)pair0 * mbc, 0, (size_t)count * mbc * sizeof( int32_t ), s ) );
    XD_CHECK( cudaMemsetAsync( sums + (size_t)pair0 * X264DSP_LA_SUMS, 0, (size_t)count * X264DSP_LA_SUMS * sizeof( int32_t ), s ) );
    if( row_satds )
        XD_CHECK( cudaMemsetAsync( row_satds + (size_t)pair0 * 2 * g->mb_h, 0, (size_t)count * 2 * g->mb_h * sizeof( int32_t ), s ) );
    XD_CHECK( cudaMemsetAsync( ctx->la_ticket + ticket_slot, 0, sizeof( int32_t ), s ) );

    xd_la_args A;
    A.g = *g;
    A.slots = slots;
    A.b = b_dev; A.p0 = p0_dev; A.want_intra = wi_dev;
    A.n_pairs = count;
    A.pair0 = pair0;
    A.mvs = mvs; A.costs = costs; A.sums = sums; A.row_satds = row_satds;
    A.cost_mv = ctx->cost_mv_dev[X264DSP_LOOKAHEAD_QP] + 4096;
    A.sync = ctx->la_sync;
    A.icost = ctx->la_icost;
    A.ticket = ctx->la_ticket + ticket_slot;
    A.epoch = ctx->la_epoch;
    A.me_range = 16;                                 // x264_param_default: analyse.i_me_range
    {
        static int slack = -1;
        if( slack < 0 )
        {
            const char *e = getenv( "X264DSP_LA_SLACK" );            // tuning knob
            slack = e ? atoi( e ) : 8;
            if( slack < 0 ) slack = 0;
        }
        A.slack = slack;
    }
    A.timing = ctx->la_timing;

    const int inner = ( g->mb_w - 2 ) * ( g->mb_h - 2 );
    dim3 igrid( ( inner + 127 ) / 128, count );
    int pslot = xd_prof_begin( ctx, XD_PROF_LA_INTRA, s );
    xd_la_intra_kernel<<<igrid, 128, 0, s>>>( A );
    xd_prof_end( ctx, XD_PROF_LA_INTRA, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    if( n_inter > 0 )
    {
        // Two mappings of the same search (both bit-exact): the multi-row kernel executes a quarter of the
        // instructions per block and wins once the batch fills the machine; with few pairs in flight the
        // launch is bound by the per-pair dependency chain and the warp-per-row kernel, which spreads a
        // block over 32 lanes and a frame over four times as many warps, has the shorter chain.
        // Measured crossover on B200: a few dozen pairs.  X264DSP_LA_ROW=0/1 forces one of them.
        static int env_row = -2;
        if( env_row == -2 )
        {
            const char *e = getenv( "X264DSP_LA_ROW" );
            env_row = e ? ( atoi( e ) > 0 ) : -1;
        }
        const int timed = ctx->la_timing != NULL;
        const int force_row = ctx->la_kernel ? ( ctx->la_kernel == 1 ) : env_row;
        const int row_kernel = timed || ( force_row >= 0 ? force_row : n_inter < 48 );
        // 2 = 8 lanes per block (four block rows per warp), 3 = 4 lanes per block (eight block rows per warp)
        const int lanes_mode = ctx->la_kernel >= 2 ? ctx->la_kernel : 3;
        const int rows = g->mb_h - 2;
        const int rows_per_warp = lanes_mode == 3 ? 8 : 4;
        const int total_warps = row_kernel ? n_inter * rows : n_inter * ( ( rows + rows_per_warp - 1 ) / rows_per_warp );
        int ctas = ( total_warps + LA_WARPS - 1 ) / LA_WARPS;
        // persistent launch: never more CTAs than fit on the machine at once (work beyond that is
        // picked up through the ticket as warps finish; the ticket order keeps that deadlock-free)
        static int per_sm[4] = { 0, 0, 0, 0 };
        const int which = timed ? 1 : row_kernel ? 0 : lanes_mode;
        if( !per_sm[which] )
        {
            if( which == 1 )
                XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[1], xd_la_inter_kernel<true>, LA_WARPS * 32, 0 ) );
            else if( which == 0 )
                XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[0], xd_la_inter_kernel<false>, LA_WARPS * 32, 0 ) );
            else if( which == 2 )
                XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[2], xd_la_multi_kernel<8>, LA_WARPS * 32, 0 ) );
            else
                XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[3], xd_la_multi_kernel<4>, LA_WARPS * 32, 0 ) );
            if( per_sm[which] < 1 )
                per_sm[which] = 1;
            const char *e = getenv( "X264DSP_LA_CTAS_PER_SM" );      // tuning knob (tools/la_phase_timing.py)
            if( e && atoi( e ) > 0 && atoi( e ) < per_sm[which] )
                per_sm[which] = atoi( e );
        }
        const int cap = ctx->sm_count * per_sm[which];
        if( ctas > cap )
            ctas = cap;
        pslot = xd_prof_begin( ctx, XD_PROF_LA_INTER, s );
        if( which == 1 )
            xd_la_inter_kernel<true><<<ctas, LA_WARPS * 32, 0, s>>>( A, n_inter, inter_list );
        else if( which == 0 )
            xd_la_inter_kernel<false><<<ctas, LA_WARPS * 32, 0, s>>>( A, n_inter, inter_list );
        else if( which == 2 )
            xd_la_multi_kernel<8><<<ctas, LA_WARPS * 32, 0, s>>>( A, n_inter, inter_list );
        else
            xd_la_multi_kernel<4><<<ctas, LA_WARPS * 32, 0, s>>>( A, n_inter, inter_list );
        xd_prof_end( ctx, XD_PROF_LA_INTER, pslot, s );
        ctx->launches++;
        XD_CHECK( cudaGetLastError() );
    }
    return 0;
}

// pair descriptors live in a small device array owned by the context:
// [b | p0 | inter list | want_intra bytes].  inter_pos[i] = number of inter pairs before pair i.
static int xd_la_upload_desc( x264dsp_ctx_t *ctx, int n_pairs, const int32_t *b, const int32_t *p0,
                              const uint8_t *want_intra, int *n_inter_out, cudaStream_t s,
                              const int32_t **b_dev, const int32_t **p0_dev, const int32_t **list_dev,
                              const uint8_t **wi_dev )
{
    const size_t words = (size_t)n_pairs * 3;
    const size_t bytes = words * sizeof( int32_t ) + n_pairs;
    if( ctx->clip_desc_cap < bytes )
        XD_CHECK( cudaDeviceSynchronize() );
    int rc = xd_reserve_dev( (void **)&ctx->clip_desc, &ctx->clip_desc_cap, bytes );
    if( rc )
        return rc;
    int32_t *host = (int32_t *)malloc( bytes );
    if( !host )
        return X264DSP_E_NOMEM;
    int n_inter = 0;
    for( int i = 0; i < n_pairs; i++ )
    {
        host[i] = b[i];
        host[n_pairs + i] = p0[i];
        if( p0[i] >= 0 )
            host[2 * n_pairs + n_inter++] = i;
    }
    memcpy( host + words, want_intra, n_pairs );
    // the same batch shape is usually analysed over and over: keep the last description and skip
    // the upload (and its synchronisation) when nothing changed
    if( ctx->desc_cache && ctx->desc_cache_bytes == bytes && !memcmp( ctx->desc_cache, host, bytes ) )
        free( host );
    else
    {
        cudaError_t e = cudaDeviceSynchronize();         // earlier launches may still read the old one
        if( e == cudaSuccess )
            e = cudaMemcpyAsync( ctx->clip_desc, host, bytes, cudaMemcpyHostToDevice, s );
        if( e == cudaSuccess )
            e = cudaStreamSynchronize( s );              // host staging buffer is pageable
        free( ctx->desc_cache );
        ctx->desc_cache = host;
        ctx->desc_cache_bytes = bytes;
        if( e != cudaSuccess )
        {
            ctx->desc_cache_bytes = 0;
            return (int)e;
        }
    }
    *n_inter_out = n_inter;
    *b_dev = ctx->clip_desc;
    *p0_dev = ctx->clip_desc + n_pairs;
    *list_dev = ctx->clip_desc + 2 * n_pairs;
    *wi_dev = (const uint8_t *)( ctx->clip_desc + words );
    return 0;
}

// b / p0 / want_intra are HOST arrays (a handful of integers describing the batch); everything
// pixel- or result-sized is device memory.
extern "C" int x264dsp_lookahead_frame_cost_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *slots,
                                                  int n_pairs, const int32_t *b, const int32_t *p0,
                                                  const uint8_t *want_intra,
                                                  int16_t *mvs, int32_t *costs, int32_t *sums, int32_t *row_satds,
                                                  void *stream )
{
    if( !ctx || !g || !slots || n_pairs <= 0 || !b || !p0 || !want_intra || !mvs || !costs || !sums )
        return X264DSP_E_ARG;
    if( g->mb_w < 3 || g->mb_h < 3 )
        return X264DSP_E_ARG;                        // do_edges would be forced on (slicetype.c:285)
    cudaStream_t s = xd_stream( ctx, stream );
    int n_inter = 0;
    const int32_t *b_dev, *p0_dev, *list_dev;
    const uint8_t *wi_dev;
    // the sync words, the ticket and the batch description are the context's: a call on another stream queues behind
    // the previous one instead of resetting its counters
    int rc = xd_scratch_acquire( ctx, XD_SCRATCH_LOOKAHEAD, s );
    if( rc )
        return rc;
    rc = xd_la_upload_desc( ctx, n_pairs, b, p0, want_intra, &n_inter, s, &b_dev, &p0_dev, &list_dev, &wi_dev );
    if( rc )
        return rc;
    if( ( rc = xd_la_prepare( ctx, g, n_pairs, s ) ) )
        return rc;
    rc = xd_la_launch( ctx, g, slots, 0, n_pairs, b_dev, p0_dev, wi_dev, list_dev, n_inter, 0,
                       mvs, costs, sums, row_satds, s );
    if( rc )
        return rc;
    return xd_scratch_release( ctx, XD_SCRATCH_LOOKAHEAD, s );
}

int xd_frame_lowres_from_luma( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *luma, uint8_t *slots,
                               int n_frames, cudaStream_t s );

static bool xd_is_pinned( const void *p )
{
    cudaPointerAttributes attr;
    if( cudaPointerGetAttributes( &attr, p ) != cudaSuccess )
    {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Lookahead pass of n_clips independent clips of clip_len frames each, from HOST memory.
// The batch is cut into up to XD_AUX_STREAMS groups of whole clips; each group runs on its own stream
// (H2D copy -> staging kernel -> lowres planes -> intra + inter cost kernels -> D2H), so that one
// group's copies overlap another group's kernels.  Pinned caller buffers (x264dsp_host_alloc) are
// copied from / to directly; ordinary memory goes through the context's pinned staging area.
extern "C" int x264dsp_lookahead_clips_host( x264dsp_ctx_t *ctx, int width, int height, int n_clips, int clip_len,
                                              const uint8_t *luma, int16_t *mvs, int32_t *costs, int32_t *sums )
{
    if( !ctx || !luma || !mvs || !costs || !sums || n_clips <= 0 || clip_len <= 0 )
        return X264DSP_E_ARG;
    x264dsp_geom_t g;
    int rc = x264dsp_geometry( width, height, &g );
    if( rc )
        return rc;
    if( g.mb_w < 3 || g.mb_h < 3 )
        return X264DSP_E_ARG;
    const int n_frames = n_clips * clip_len;
    const size_t pic = (size_t)width * height;
    const size_t mbc = g.mb_count;
    const size_t blocks = (size_t)n_frames * mbc;
    const size_t mv_bytes = blocks * 2 * sizeof( int16_t ), cost_bytes = blocks * sizeof( int32_t );
    const size_t sum_bytes = (size_t)n_frames * X264DSP_LA_SUMS * sizeof( int32_t );
    const size_t out_bytes = mv_bytes + cost_bytes + sum_bytes;
    const bool pinned_in = xd_is_pinned( luma );
    const bool pinned_out = xd_is_pinned( mvs ) && xd_is_pinned( costs ) && xd_is_pinned( sums );

    if( ctx->stage_dev_cap < pic * n_frames || ctx->clip_out_cap < out_bytes
        || ctx->clip_slots_cap < (size_t)g.slot_bytes * n_frames )
        XD_CHECK( cudaDeviceSynchronize() );
    if( !pinned_in && ( rc = xd_reserve_pinned( (void **)&ctx->stage_host, &ctx->stage_host_cap, pic * n_frames ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->stage_dev, &ctx->stage_dev_cap, pic * n_frames ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->clip_out, &ctx->clip_out_cap, out_bytes ) ) ) return rc;
    if( !pinned_out && ( rc = xd_reserve_pinned( (void **)&ctx->clip_out_host, &ctx->clip_out_host_cap, out_bytes ) ) ) return rc;
    if( ctx->clip_slots_cap < (size_t)g.slot_bytes * n_frames )
    {
        if( ( rc = xd_reserve_dev( (void **)&ctx->clip_slots, &ctx->clip_slots_cap, (size_t)g.slot_bytes * n_frames ) ) )
            return rc;
        XD_CHECK( cudaMemsetAsync( ctx->clip_slots, 0, ctx->clip_slots_cap, ctx->stream ) );
        XD_CHECK( cudaStreamSynchronize( ctx->stream ) );
    }

    // batch description: frame i is analysed against frame i-1 unless it opens a clip
    int32_t *b = (int32_t *)malloc( (size_t)n_frames * ( 2 * sizeof( int32_t ) + 1 ) );
    if( !b )
        return X264DSP_E_NOMEM;
    int32_t *p0 = b + n_frames;
    uint8_t *wi = (uint8_t *)( p0 + n_frames );
    for( int i = 0; i < n_frames; i++ )
    {
        b[i] = i;
        p0[i] = ( i % clip_len ) ? i - 1 : -1;
        wi[i] = 1;
    }
    int n_inter = 0;
    const int32_t *b_dev, *p0_dev, *list_dev;
    const uint8_t *wi_dev;
    rc = xd_la_upload_desc( ctx, n_frames, b, p0, wi, &n_inter, ctx->stream, &b_dev, &p0_dev, &list_dev, &wi_dev );
    free( b );
    if( rc )
        return rc;
    if( ( rc = xd_la_prepare( ctx, &g, n_frames, ctx->stream ) ) )
        return rc;

    int16_t *d_mvs = (int16_t *)ctx->clip_out;
    int32_t *d_costs = (int32_t *)( ctx->clip_out + mv_bytes );
    int32_t *d_sums = (int32_t *)( ctx->clip_out + mv_bytes + cost_bytes );
    uint8_t *h_out = ctx->clip_out_host;

    // enough groups that the last group's kernels (the only work not hidden behind a copy) are a small
    // share of the batch, but at least four clips per group so that every launch still fills its pipeline
    int groups = n_clips / 4;
    if( groups < 1 ) groups = 1;
    if( groups > XD_AUX_STREAMS ) groups = XD_AUX_STREAMS;
    if( n_clips <= 4 ) groups = n_clips < 2 ? 1 : 2;
    // a failure in a later group must not return while earlier groups' copies are still writing into the caller's
    // buffers: leave the loop, drain every stream that was used, then report
#define XD_GROUP_CHECK( call ) { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { rc = (int)e_; break; } }
    int used = 0;
    for( int gi = 0; gi < groups && !rc; gi++ )
    {
        cudaStream_t s = ctx->aux[gi];
        const int c0 = (int)( (int64_t)n_clips * gi / groups ), c1 = (int)( (int64_t)n_clips * ( gi + 1 ) / groups );
        const int f0 = c0 * clip_len, nf = ( c1 - c0 ) * clip_len;
        if( nf <= 0 )
            continue;
        used = gi + 1;
        // a lookahead call enqueued earlier on some other stream still owns the context's sync words
        if( ( rc = xd_scratch_acquire( ctx, XD_SCRATCH_LOOKAHEAD, s ) ) )
            break;
        const uint8_t *src = luma + (size_t)f0 * pic;
        if( !pinned_in )
        {
            memcpy( ctx->stage_host + (size_t)f0 * pic, src, pic * nf );
            src = ctx->stage_host + (size_t)f0 * pic;
        }
        XD_GROUP_CHECK( cudaMemcpyAsync( ctx->stage_dev + (size_t)f0 * pic, src, pic * nf, cudaMemcpyHostToDevice, s ) );
        uint8_t *slots = ctx->clip_slots + (size_t)f0 * g.slot_bytes;
        if( !ctx->copies_only && ( rc = xd_frame_lowres_from_luma( ctx, &g, ctx->stage_dev + (size_t)f0 * pic, slots, nf, s ) ) )
            break;
        // every clip of the group contributes clip_len-1 inter pairs, in frame order
        const int inter0 = c0 * ( clip_len - 1 ), ninter = ( c1 - c0 ) * ( clip_len - 1 );
        if( !ctx->copies_only && ( rc = xd_la_launch( ctx, &g, ctx->clip_slots, f0, nf, b_dev, p0_dev, wi_dev, list_dev + inter0,
                                                      ninter, gi, d_mvs, d_costs, d_sums, NULL, s ) ) )
            break;
        uint8_t *o_mv = pinned_out ? (uint8_t *)mvs : h_out;
        uint8_t *o_cost = pinned_out ? (uint8_t *)costs : h_out + mv_bytes;
        uint8_t *o_sum = pinned_out ? (uint8_t *)sums : h_out + mv_bytes + cost_bytes;
        const size_t mo = (size_t)f0 * mbc * 2 * sizeof( int16_t ), co = (size_t)f0 * mbc * sizeof( int32_t );
        const size_t so = (size_t)f0 * X264DSP_LA_SUMS * sizeof( int32_t );
        XD_GROUP_CHECK( cudaMemcpyAsync( o_mv + mo, (uint8_t *)d_mvs + mo, (size_t)nf * mbc * 2 * sizeof( int16_t ), cudaMemcpyDeviceToHost, s ) );
        XD_GROUP_CHECK( cudaMemcpyAsync( o_cost + co, (uint8_t *)d_costs + co, (size_t)nf * mbc * sizeof( int32_t ), cudaMemcpyDeviceToHost, s ) );
        XD_GROUP_CHECK( cudaMemcpyAsync( o_sum + so, (uint8_t *)d_sums + so, (size_t)nf * X264DSP_LA_SUMS * sizeof( int32_t ), cudaMemcpyDeviceToHost, s ) );
    }
#undef XD_GROUP_CHECK
    for( int gi = 0; gi < used; gi++ )
    {
        const cudaError_t e = cudaStreamSynchronize( ctx->aux[gi] );
        if( e != cudaSuccess && !rc )
            rc = (int)e;
    }
    ctx->scratch_busy[XD_SCRATCH_LOOKAHEAD] = 0;         // everything that used the scratch has finished
    if( rc )
        return rc;
    if( !pinned_out )
    {
        memcpy( mvs, h_out, mv_bytes );
        memcpy( costs, h_out + mv_bytes, cost_bytes );
        memcpy( sums, h_out + mv_bytes + cost_bytes, sum_bytes );
    }
    return 0;
}

extern "C" int x264dsp_lookahead_clip_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames,
                                             const uint8_t *luma, int16_t *mvs, int32_t *costs, int32_t *sums )
{
    return x264dsp_lookahead_clips_host( ctx, width, height, 1, n_frames, luma, mvs, costs, sums );
}

// Measurement aid (bench.py's copy-only probe): with on != 0 x264dsp_lookahead_clips_host issues exactly the copies it
// always issues -- same buffers, same streams, same order -- and none of its kernels, so that the time of the call is
// what the host-to-device / device-to-host DMA alone costs on this box.  The outputs are then meaningless.
extern "C" int x264dsp_debug_copies_only( x264dsp_ctx_t *ctx, int on )
{
    if( !ctx )
        return X264DSP_E_ARG;
    ctx->copies_only = on != 0;
    return 0;
}

// mode 0 = pick by batch size, 1 = warp per block row, 2 = four block rows per warp, 3 = eight (identical results)
extern "C" int x264dsp_lookahead_select_kernel( x264dsp_ctx_t *ctx, int mode )
{
    if( !ctx || mode < 0 || mode > 3 )
        return X264DSP_E_ARG;
    ctx->la_kernel = mode;
    return 0;
}

// Debug aid: cycle counters of the inter kernel's phases, summed over all warps since the last call
// with reset != 0.  out[0..9] = wait, setup, zero-mv SATD, candidates, diamond, sub-pel, tail, blocks,
// first wait of a row, rows.
extern "C" int x264dsp_debug_lookahead_timing( x264dsp_ctx_t *ctx, int enable, int reset, uint64_t out[10] )
{
    if( !ctx )
        return X264DSP_E_ARG;
    XD_CHECK( cudaDeviceSynchronize() );
    if( enable && !ctx->la_timing )
    {
        XD_CHECK( cudaMalloc( (void **)&ctx->la_timing, LA_T_KINDS * sizeof( unsigned long long ) ) );
        XD_CHECK( cudaMemset( ctx->la_timing, 0, LA_T_KINDS * sizeof( unsigned long long ) ) );
    }
    if( out && ctx->la_timing )
        XD_CHECK( cudaMemcpy( out, ctx->la_timing, LA_T_KINDS * sizeof( unsigned long long ), cudaMemcpyDeviceToHost ) );
    if( reset && ctx->la_timing )
        XD_CHECK( cudaMemset( ctx->la_timing, 0, LA_T_KINDS * sizeof( unsigned long long ) ) );
    if( !enable && ctx->la_timing )
    {
        XD_CHECK( cudaFree( ctx->la_timing ) );
        ctx->la_timing = NULL;
    }
    return 0;
}
