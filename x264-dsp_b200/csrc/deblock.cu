// deblock.cu -- in-loop deblocking of a whole frame and boundary-strength derivation, sm_100a.
//
// Reference: deblock_{v,h}_{luma,chroma}[_intra]_c (common/deblock.c:80-295), deblock_edge and
// x264_frame_deblock_row (deblock.c:325-427, slice-QP rule), deblock_strength_c (deblock.c:297-323).
//
// The reference filters macroblocks in raster order; MB (x,y) reads pixels its left, top and
// top-right neighbours have already modified, a 2-step wavefront.  Mapping:
//   * one WARP per macroblock row, walking left to right; rows of ALL frames of a launch are taken
//     through an atomic ticket (frame-interleaved, top rows first), and row y proceeds to MB x once
//     row y-1 has finished MB x+1 (a progress counter per row, release/acquire through L2);
//   * the warp keeps the macroblock with its 4-sample left and top aprons (20x20 luma, 10 rows x 20
//     bytes NV12) in a shared-memory tile: the vertical-edge pass has the 32 lanes on the 16 luma
//     lines plus the 8 x {U,V} chroma lines, the horizontal-edge pass on the 16 luma columns plus the
//     16 chroma byte columns; a lane keeps its line in registers across the four edges, exactly the
//     sequential dependence the reference has, while the lines run in parallel;
//   * columns 12..15 stay in the tile and become the next macroblock's left apron, so within a row
//     pixels never round-trip through memory; only the rows shared with the macroblock row above /
//     below cross warps, through L2 (ld.cg / st.cg) with a fence before the row counter advances;
//   * the macroblock's own samples and its side information (bS, types, cbp) are fetched one
//     macroblock ahead.
#include "common.cuh"
#include "leaf.cuh"

// what to do on one edge: mode 0 = nothing, 1 = bS<4 filter, 2 = bS=4 filter
struct xd_edge
{
    int mode;
    uint32_t bs;        // four packed bS values
};

__device__ __forceinline__ xd_edge xd_edge_inter( uint32_t bs, int alpha, int beta )
{
    xd_edge e;
    e.bs = bs;
    e.mode = ( bs != 0 && alpha != 0 && beta != 0 ) ? 1 : 0;      // deblock.c:330-331
    return e;
}

#define DB_WARPS 4
#define DB_PITCH 24                     // tile row pitch: columns -4 .. 19 of the macroblock
#define DB_LROWS 20                     // luma rows -4 .. 15
#define DB_CROWS 10                     // chroma rows -2 .. 7

struct xd_db_args
{
    x264dsp_geom_t g;
    uint8_t *slots;
    int n_frames;
    const int8_t *mb_type;
    const uint8_t *partition;
    const int16_t *cbp;
    const uint8_t *bs;
    xd_db_params P;
    int32_t *progress;                  // [n_frames][mb_h] macroblocks finished per row
    int32_t *ticket;
    int latency;                        // few units in flight: announce progress at once and poll without backing off
};

// one 32-bit item of the next macroblock's side information per lane, fetched a macroblock ahead:
// lanes 0..15 the sixteen bS words ([dir][edge] of 4 bS), 16 mb_type, 17 mb_type of the MB above,
// 18 partition, 19 cbp, 20 mb_type of the MB to the left
__device__ __forceinline__ uint32_t xd_db_meta( const xd_db_args &A, size_t mb0, int xy, int mb_x, int mb_y, int lane )
{
    const int W = A.g.mb_w;
    if( lane < 16 )
        return __ldg( (const uint32_t *)( A.bs + ( mb0 + xy ) * 64 ) + lane );
    if( lane == 16 ) return (uint32_t)(int)__ldg( A.mb_type + mb0 + xy );
    if( lane == 17 ) return mb_y > 0 ? (uint32_t)(int)__ldg( A.mb_type + mb0 + xy - W ) : 127u;
    if( lane == 18 ) return __ldg( A.partition + mb0 + xy );
    if( lane == 19 ) return (uint32_t)(int)__ldg( A.cbp + mb0 + xy );
    if( lane == 20 ) return mb_x > 0 ? (uint32_t)(int)__ldg( A.mb_type + mb0 + xy - 1 ) : 127u;
    return 0;
}

// The macroblock's own samples (nobody has modified them yet): lanes 0..15 one luma row each,
// lanes 16..23 one NV12 chroma row each, as 16 bytes.
__device__ __forceinline__ uint4 xd_db_own( const uint8_t *py, const uint8_t *pc, int ls, int cs, int lane )
{
    if( lane < 16 )
        return __ldcg( (const uint4 *)( py + (int64_t)lane * ls ) );
    if( lane < 24 )
        return __ldcg( (const uint4 *)( pc + (int64_t)( lane - 16 ) * cs ) );
    return make_uint4( 0, 0, 0, 0 );
}

__device__ __forceinline__ void xd_db_put16( uint8_t *row, uint4 v )        // row + 4 is 4-byte aligned
{
    uint32_t *w = (uint32_t *)( row + 4 );
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}

// 8 CTAs per SM (58 registers): a row's warp spends most of its time in dependent-issue latency or waiting for the
// row above, so throughput follows the number of resident rows -- 18.7 -> 14.9 us per 1080p frame in a 192-frame batch
// against 6 CTAs (84 registers); 10 CTAs (48 registers, spills) measured no better
__global__ void __launch_bounds__( DB_WARPS * 32, 8 )
xd_deblock_kernel( xd_db_args A )
{
    __shared__ __align__( 16 ) uint8_t s_luma[DB_WARPS][DB_LROWS * DB_PITCH];
    __shared__ __align__( 16 ) uint8_t s_chroma[DB_WARPS][DB_CROWS * DB_PITCH];
    const int lane = threadIdx.x & 31;
    uint8_t *L = s_luma[threadIdx.x >> 5], *Cc = s_chroma[threadIdx.x >> 5];
    const x264dsp_geom_t &g = A.g;
    const xd_db_params &P = A.P;
    const int W = g.mb_w, H = g.mb_h, ls = g.luma_stride, cs = g.chroma_stride;
    const int total = A.n_frames * H;
    for( ;; )
    {
        int t = 0;
        if( lane == 0 )
            t = atomicAdd( A.ticket, 1 );
        t = __shfl_sync( 0xffffffffu, t, 0 );
        if( t >= total )
            return;
        // rows are dealt frame-interleaved, top rows first: row (f, y) waits on (f, y-1), whose
        // ticket is n_frames smaller and therefore held by a warp that is already running
        const int mb_y = t / A.n_frames, f = t - mb_y * A.n_frames;
        uint8_t *slot = A.slots + (size_t)f * g.slot_bytes;
        const size_t mb0 = (size_t)f * g.mb_count;
        volatile int32_t *mine = A.progress + (size_t)f * H + mb_y;
        volatile int32_t *above = mine - 1;
        uint8_t *row_y = slot + g.luma_origin + (int64_t)( mb_y << 4 ) * ls;
        uint8_t *row_c = slot + g.slot_chroma_off + g.chroma_origin + (int64_t)( mb_y << 3 ) * cs;

        uint4 own = xd_db_own( row_y, row_c, ls, cs, lane );
        uint32_t meta = xd_db_meta( A, mb0, mb_y * W, 0, mb_y, lane );
        int seen = 0;
        for( int mb_x = 0; mb_x < W; mb_x++ )
        {
            uint8_t *py = row_y + ( mb_x << 4 ), *pc = row_c + ( mb_x << 4 );
            // ---- this macroblock's side information (fetched during the previous macroblock)
            uint32_t bs[16];
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                bs[k] = __shfl_sync( 0xffffffffu, meta, k );
                bs[8 + k] = __shfl_sync( 0xffffffffu, meta, 8 + k );
            }
            const int type_cur = (int)__shfl_sync( 0xffffffffu, meta, 16 ), type_top = (int)__shfl_sync( 0xffffffffu, meta, 17 );
            const int part = (int)__shfl_sync( 0xffffffffu, meta, 18 ), cbp_cur = (int)(int16_t)__shfl_sync( 0xffffffffu, meta, 19 );
            const int type_left = (int)__shfl_sync( 0xffffffffu, meta, 20 );
            const bool intra_cur = type_cur < 4;
            const bool first_only = part == 16 && cbp_cur == 0 && !intra_cur;

            // ---- commit the macroblock's own samples to the tile, start fetching the next one's
            if( lane < 16 )
                xd_db_put16( L + ( lane + 4 ) * DB_PITCH, own );
            else if( lane < 24 )
                xd_db_put16( Cc + ( lane - 16 + 2 ) * DB_PITCH, own );
            if( mb_x + 1 < W )
            {
                own = xd_db_own( py + 16, pc + 16, ls, cs, lane );
                meta = xd_db_meta( A, mb0, mb_y * W + mb_x + 1, mb_x + 1, mb_y, lane );
            }

            // ---- the row above must have finished macroblock x+1; then its bottom rows are final
            if( mb_y > 0 )
            {
                const int need = min( mb_x + 2, W );
                unsigned ns = 20;
                while( seen < need )
                {
                    seen = *above;
                    if( seen < need )
                    {
                        __nanosleep( ns );
                        if( ns < 400 )
                            ns += 20;
                    }
                }
                __threadfence();
                if( lane >= 24 && lane < 28 )
                    xd_db_put16( L + ( lane - 24 ) * DB_PITCH, __ldcg( (const uint4 *)( py + (int64_t)( lane - 28 ) * ls ) ) );
                else if( lane >= 28 && lane < 30 )
                    xd_db_put16( Cc + ( lane - 28 ) * DB_PITCH, __ldcg( (const uint4 *)( pc + (int64_t)( lane - 30 ) * cs ) ) );
            }
            __syncwarp();

            // ======================= vertical edges (filter across x) =======================
            {
                xd_edge le[4], ce[2];
                const bool left_intra = mb_x > 0 && ( intra_cur || type_left < 4 );
                le[0].mode = mb_x == 0 ? 0 : left_intra ? 2 : xd_edge_inter( bs[0], P.alpha, P.beta ).mode;
                le[0].bs = bs[0];
                ce[0].mode = mb_x == 0 ? 0 : left_intra ? 2 : xd_edge_inter( bs[0], P.alphac, P.betac ).mode;
                ce[0].bs = bs[0];
#pragma unroll
                for( int e = 1; e < 4; e++ )
                {
                    le[e] = xd_edge_inter( bs[e], P.alpha, P.beta );
                    if( first_only )
                        le[e].mode = 0;
                }
                ce[1] = xd_edge_inter( bs[2], P.alphac, P.betac );
                if( first_only )
                    ce[1].mode = 0;

                if( lane < 16 )
                {
                    // luma line `lane`: pixels x-4 .. x+15 in registers
                    uint32_t *row = (uint32_t *)( L + ( lane + 4 ) * DB_PITCH );
                    uint32_t w[5];
#pragma unroll
                    for( int k = 0; k < 5; k++ )
                        w[k] = row[k];
                    const int grp = lane >> 2;
#pragma unroll
                    for( int e = 0; e < 4; e++ )
                    {
                        if( le[e].mode == 0 )
                            continue;
                        int s[8];
#pragma unroll
                        for( int k = 0; k < 4; k++ )
                        {
                            s[k] = ( w[e] >> ( 8 * k ) ) & 255;
                            s[4 + k] = ( w[e + 1] >> ( 8 * k ) ) & 255;
                        }
                        if( le[e].mode == 2 )
                            xd_luma_intra_line( s, P.alpha, P.beta );
                        else
                        {
                            const int tc0 = xd_tc0( P.ia, ( le[e].bs >> ( 8 * grp ) ) & 255 );
                            if( tc0 >= 0 )
                                xd_luma_line( s, P.alpha, P.beta, tc0 );
                        }
                        w[e] = s[0] | ( s[1] << 8 ) | ( s[2] << 16 ) | ( s[3] << 24 );
                        w[e + 1] = s[4] | ( s[5] << 8 ) | ( s[6] << 16 ) | ( s[7] << 24 );
                    }
#pragma unroll
                    for( int k = 0; k < 5; k++ )
                        row[k] = w[k];
                }
                else
                {
                    // chroma row r, component c: edges at pair 0 (byte 0) and pair 4 (byte 8)
                    const int r = ( lane - 16 ) >> 1, c = ( lane - 16 ) & 1;
                    uint8_t *row = Cc + ( r + 2 ) * DB_PITCH + 4 + c;
                    const int grp = r >> 1;
#pragma unroll
                    for( int e = 0; e < 2; e++ )
                    {
                        if( ce[e].mode == 0 )
                            continue;
                        uint8_t *q = row + 8 * e;
                        int s[4] = { q[-4], q[-2], q[0], q[2] };
                        if( ce[e].mode == 2 )
                            xd_chroma_line( s, P.alphac, P.betac, 0, true );
                        else
                        {
                            const int tc = xd_tc0( P.iac, ( ce[e].bs >> ( 8 * grp ) ) & 255 ) + 1;
                            if( tc > 0 )
                                xd_chroma_line( s, P.alphac, P.betac, tc, false );
                        }
                        q[-2] = (uint8_t)s[1];
                        q[0] = (uint8_t)s[2];
                    }
                }
            }
            __syncwarp();

            // ======================= horizontal edges (filter across y) =======================
            {
                xd_edge le[4], ce[2];
                const bool top_intra = mb_y > 0 && ( intra_cur || type_top < 4 );
                le[0].mode = mb_y == 0 ? 0 : top_intra ? 2 : xd_edge_inter( bs[8], P.alpha, P.beta ).mode;
                le[0].bs = bs[8];
                ce[0].mode = mb_y == 0 ? 0 : top_intra ? 2 : xd_edge_inter( bs[8], P.alphac, P.betac ).mode;
                ce[0].bs = bs[8];
#pragma unroll
                for( int e = 1; e < 4; e++ )
                {
                    le[e] = xd_edge_inter( bs[8 + e], P.alpha, P.beta );
                    if( first_only )
                        le[e].mode = 0;
                }
                ce[1] = xd_edge_inter( bs[10], P.alphac, P.betac );
                if( first_only )
                    ce[1].mode = 0;

                if( lane < 16 )
                {
                    // luma column `lane`: rows -4 .. 15
                    uint8_t *col = L + 4 + lane;
                    const bool any = ( le[0].mode | le[1].mode | le[2].mode | le[3].mode ) != 0;
                    if( any )
                    {
                        int v[20];
#pragma unroll
                        for( int k = 0; k < 20; k++ )
                            v[k] = col[k * DB_PITCH];
                        const int grp = lane >> 2;
#pragma unroll
                        for( int e = 0; e < 4; e++ )
                        {
                            if( le[e].mode == 0 )
                                continue;
                            int s[8];
#pragma unroll
                            for( int k = 0; k < 8; k++ )
                                s[k] = v[4 * e + k];
                            if( le[e].mode == 2 )
                                xd_luma_intra_line( s, P.alpha, P.beta );
                            else
                            {
                                const int tc0 = xd_tc0( P.ia, ( le[e].bs >> ( 8 * grp ) ) & 255 );
                                if( tc0 >= 0 )
                                    xd_luma_line( s, P.alpha, P.beta, tc0 );
                            }
#pragma unroll
                            for( int k = 0; k < 8; k++ )
                                v[4 * e + k] = s[k];
                        }
#pragma unroll
                        for( int k = 1; k < 19; k++ )
                            col[k * DB_PITCH] = (uint8_t)v[k];
                    }
                }
                else
                {
                    // chroma byte b of the 16-byte row (pair b>>1, component b&1): edges at chroma rows 0 and 4
                    const int b = lane - 16;
                    uint8_t *col = Cc + 4 + b;
                    const int grp = b >> 2;
#pragma unroll
                    for( int e = 0; e < 2; e++ )
                    {
                        if( ce[e].mode == 0 )
                            continue;
                        uint8_t *q = col + ( 2 + 4 * e ) * DB_PITCH;
                        int s[4] = { q[-2 * DB_PITCH], q[-DB_PITCH], q[0], q[DB_PITCH] };
                        if( ce[e].mode == 2 )
                            xd_chroma_line( s, P.alphac, P.betac, 0, true );
                        else
                        {
                            const int tc = xd_tc0( P.iac, ( ce[e].bs >> ( 8 * grp ) ) & 255 ) + 1;
                            if( tc > 0 )
                                xd_chroma_line( s, P.alphac, P.betac, tc, false );
                        }
                        q[-DB_PITCH] = (uint8_t)s[1];
                        q[0] = (uint8_t)s[2];
                    }
                }
            }
            __syncwarp();

            // ---- write back what is final now: columns -4 .. 11 of the macroblock's rows (columns
            // 12..15 still face the next macroblock's left edge and travel on in the tile), plus the
            // rows of the macroblock above that the top edge touched
            {
                const bool last = mb_x == W - 1;
                if( lane < 24 )
                {
                    const uint32_t *src = (const uint32_t *)( lane < 16 ? L + ( lane + 4 ) * DB_PITCH : Cc + ( lane - 16 + 2 ) * DB_PITCH );
                    uint32_t *dst = (uint32_t *)( lane < 16 ? py + (int64_t)lane * ls - 4 : pc + (int64_t)( lane - 16 ) * cs - 4 );
                    if( mb_x > 0 )
                        __stcg( dst, src[0] );
                    __stcg( dst + 1, src[1] );
                    __stcg( dst + 2, src[2] );
                    __stcg( dst + 3, src[3] );
                    if( last )
                        __stcg( dst + 4, src[4] );
                }
                if( mb_y > 0 && lane < 16 )
                {
                    if( lane < 12 )
                    {
                        const int r = lane >> 2, w = lane & 3;          // luma rows -3 .. -1
                        __stcg( (uint32_t *)( py + (int64_t)( r - 3 ) * ls ) + w, ( (const uint32_t *)( L + ( r + 1 ) * DB_PITCH + 4 ) )[w] );
                    }
                    else
                        __stcg( (uint32_t *)( pc - cs ) + ( lane - 12 ), ( (const uint32_t *)( Cc + DB_PITCH + 4 ) )[lane - 12] );
                }
                // carry columns 12..15 into the next macroblock's columns -4..-1
                if( lane < 16 )
                    *(uint32_t *)( L + ( lane + 4 ) * DB_PITCH ) = *(const uint32_t *)( L + ( lane + 4 ) * DB_PITCH + 16 );
                else if( lane < 24 )
                    *(uint32_t *)( Cc + ( lane - 16 + 2 ) * DB_PITCH ) = *(const uint32_t *)( Cc + ( lane - 16 + 2 ) * DB_PITCH + 16 );
            }
            __threadfence();            // every lane publishes its own stores device-wide ...
            __syncwarp();               // ... before lane 0 advances the row counter
            if( lane == 0 )
                *mine = mb_x + 1;
        }
    }
}

// =================================================================================================
// Second mapping of the same wavefront: a warp takes a PAIR of macroblock rows (2p, 2p+1) of one frame, lanes 0-15 on
// the upper row at macroblock i, lanes 16-31 on the lower row two macroblocks behind -- the lag the dependency needs,
// so the lower row never waits: what it reads from the row above was stored by its own warp an iteration earlier.
// Only the first row of a pair polls another warp and only the second row publishes progress (one st.release by one
// lane after a __syncwarp, instead of a __threadfence by every lane).  Inside a half-warp a lane is one sample line:
// luma row / chroma (row, component) for the vertical edges, luma column / chroma byte column for the horizontal ones,
// with the line's samples held as integers in registers across the four edges and the branch-free filters of
// dbfilter.cuh, so that both rows and all lines share one instruction stream.  The macroblock's own row and the four
// columns carried over from its left neighbour stay in registers; shared memory is used for the row <-> column
// transposition only.  Per macroblock this issues about a third of the warp instructions of the mapping above.
#include "dbfilter.cuh"

#define DB2_WARPS 4

__device__ __forceinline__ int xd_ld_acquire( const int32_t *p )
{
    int v;
    asm volatile( "ld.acquire.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ int xd_ld_relaxed( const int32_t *p )
{
    int v;
    asm volatile( "ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_st_release( int32_t *p, int v )
{
    asm volatile( "st.release.gpu.global.s32 [%0], %1;" :: "l"( p ), "r"( v ) : "memory" );
}

__device__ __forceinline__ int xd_byte( uint32_t w, int k ) { return (int)__byte_perm( w, 0u, 0x4440u | (unsigned)k ); }   // one PRMT
__device__ __forceinline__ uint32_t xd_pack4( int a, int b, int c, int d )
{
    return (uint32_t)a | ( (uint32_t)b << 8 ) | ( (uint32_t)c << 16 ) | ( (uint32_t)d << 24 );
}
// byte `g` of each of the four words of v, packed: the lane's own bS of the four edges of one direction
__device__ __forceinline__ uint32_t xd_bs_of_group( uint4 v, int g )
{
    const unsigned sel = (unsigned)g | ( (unsigned)( 4 + g ) << 4 );          // byte g of the first, byte g of the second word
    return __byte_perm( __byte_perm( v.x, v.y, sel ), __byte_perm( v.z, v.w, sel ), 0x5410u );
}

struct xd_db2_meta
{
    uint4 bsv, bsh;                 // bs[0][0..3], bs[1][0..3] of the macroblock
    int type, type_top, part, cbp;
};

__device__ __forceinline__ void xd_db2_fetch( const xd_db_args &A, size_t mb0, int xy, bool has_top, xd_db2_meta &M )
{
    const uint4 *b = (const uint4 *)( A.bs + ( mb0 + xy ) * 64 );
    M.bsv = __ldg( b );
    M.bsh = __ldg( b + 2 );
    M.type = __ldg( A.mb_type + mb0 + xy );
    M.type_top = has_top ? (int)__ldg( A.mb_type + mb0 + xy - A.g.mb_w ) : 127;
    M.part = __ldg( A.partition + mb0 + xy );
    M.cbp = __ldg( A.cbp + mb0 + xy );
}

// shared memory through 32-bit shared-window addresses: one register per tile pointer, and the compiler has no
// generic pointers to rebuild from the thread index inside the loop
__device__ __forceinline__ int xd_lds_u8( uint32_t a )
{
    int v;
    asm volatile( "ld.shared.u8 %0, [%1];" : "=r"( v ) : "r"( a ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_sts_u8( uint32_t a, int v )
{
    asm volatile( "st.shared.u8 [%0], %1;" :: "r"( a ), "r"( v ) : "memory" );
}
__device__ __forceinline__ uint32_t xd_lds_u32( uint32_t a )
{
    uint32_t v;
    asm volatile( "ld.shared.u32 %0, [%1];" : "=r"( v ) : "r"( a ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_sts_u32( uint32_t a, uint32_t v )
{
    asm volatile( "st.shared.u32 [%0], %1;" :: "r"( a ), "r"( v ) : "memory" );
}
__device__ __forceinline__ uint2 xd_lds_v2( uint32_t a )
{
    uint2 v;
    asm volatile( "ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"( v.x ), "=r"( v.y ) : "r"( a ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_sts_v2( uint32_t a, uint32_t x, uint32_t y )
{
    asm volatile( "st.shared.v2.u32 [%0], {%1, %2};" :: "r"( a ), "r"( x ), "r"( y ) : "memory" );
}

// per-lane state that lives across the macroblocks of a row
struct xd_db2_state
{
    uint4 own_y, own_c;             // the next macroblock's own samples (row l / chroma row l), fetched one ahead
    xd_db2_meta nxt;                // ... and its side information
    uint32_t carry_y, carry_c;      // columns 12..15 of the previous macroblock's row, final but for this left edge
    bool intra_prev;
    int seen;                       // last progress value read from the row over the pair
    int publish;                    // second row's macroblocks stored but not yet announced (lane 16)
    uint8_t *py, *pc;               // this lane's luma row l / chroma row (l & 7) at the CURRENT macroblock
};

// One macroblock step of both rows of a pair.  EDGE = false is the interior of the row, where both halves have a left
// neighbour and a successor (0 < x < W-1): the per-lane activity is the row's alone and the first / last column
// special cases disappear.
template<bool EDGE>
__device__ __forceinline__ void xd_db2_step( const xd_db_args &A, xd_db2_state &S, const int i, const int lane, const int h,
                                             const int l, const int grp, const uint32_t sL, const uint32_t sC,
                                             const bool row_ok, const int mb_y, const int y0, const size_t mb0,
                                             const int32_t *above, int32_t *mine, const bool luma_on, const bool chroma_on )
{
    const x264dsp_geom_t &g = A.g;
    const xd_db_params &P = A.P;
    const int W = g.mb_w, ls = g.luma_stride, cs = g.chroma_stride;
    const int x = i - 2 * h;
    const bool act = EDGE ? ( row_ok && x >= 0 && x < W ) : row_ok;
    const bool has_left = EDGE ? x > 0 : true;
    // ---- this macroblock (fetched during the previous iteration), then start fetching the next one
    const uint4 cur_y = S.own_y, cur_c = S.own_c;
    const xd_db2_meta cur = S.nxt;
    if( EDGE ? ( row_ok && x + 1 >= 0 && x + 1 < W ) : row_ok )
    {
        S.own_y = __ldcg( (const uint4 *)( S.py + 16 ) );
        if( l < 8 )
            S.own_c = __ldcg( (const uint4 *)( S.pc + 16 ) );
        xd_db2_fetch( A, mb0, mb_y * W + x + 1, mb_y > 0, S.nxt );
    }
    const bool intra_cur = (unsigned)cur.type < 4u;
    const bool first_only = cur.part == 16 && cur.cbp == 0 && !intra_cur;
    const bool left_intra = has_left && ( intra_cur || S.intra_prev );
    const bool top_intra = mb_y > 0 && ( intra_cur || (unsigned)cur.type_top < 4u );
    const uint32_t bsv = xd_bs_of_group( cur.bsv, grp ), bsh = xd_bs_of_group( cur.bsh, grp );
    S.intra_prev = intra_cur;

    // ---- the row over the pair must have finished macroblock i+1 before the upper row's top edge
    if( y0 > 0 && ( !EDGE || i < W ) )
    {
        const int need = min( i + 2, W );
        if( S.seen < need )
        {
            if( lane == 0 )
            {
                // relaxed polls (an acquire load invalidates L1 every time), backing off while the counter stands
                // still -- a pair far down the staircase waits for many macroblock times before its first macroblock
                // and must not eat the issue slots and L2 bandwidth of the rows that work -- then ONE acquire load
                unsigned ns = 32;
                int last = -1;
                for( ;; )
                {
                    const int v = xd_ld_relaxed( above );
                    if( v >= need )
                        break;
                    ns = ( v != last || A.latency ) ? 32 : min( ns * 2, 2048u );
                    last = v;
                    __nanosleep( ns );
                }
                S.seen = xd_ld_acquire( above );
            }
            S.seen = __shfl_sync( 0xffffffffu, S.seen, 0 );
        }
    }
    __syncwarp();
    // the four luma / two chroma rows above the macroblock: final now (upper row: the acquire above; lower row: stored
    // by this warp's other half during the previous iteration).  Needed from the horizontal pass on.
    uint4 top = make_uint4( 0, 0, 0, 0 );
    if( act && mb_y > 0 )
    {
        if( l < 4 )
            top = __ldcg( (const uint4 *)( S.py + (int64_t)( -4 ) * ls ) );          // py is row l: rows -4 .. -1 for l = 0 .. 3
        else if( l < 6 )
            top = __ldcg( (const uint4 *)( S.pc + (int64_t)( -6 ) * cs ) );          // pc is chroma row l: rows -2, -1 for l = 4, 5
    }

    // ======================= vertical edges, luma: lane = row l, columns -4 .. 15 =======================
    {
        int s[20];
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            s[k] = xd_byte( S.carry_y, k );
            s[4 + k] = xd_byte( cur_y.x, k );
            s[8 + k] = xd_byte( cur_y.y, k );
            s[12 + k] = xd_byte( cur_y.z, k );
            s[16 + k] = xd_byte( cur_y.w, k );
        }
#pragma unroll
        for( int e = 0; e < 4; e++ )
        {
            const int bs_e = xd_byte( bsv, e );
            const bool normal = act && luma_on && bs_e > 0 && ( e == 0 ? ( has_left && !left_intra ) : !first_only );
            if( __any_sync( 0xffffffffu, normal ) )
            {
                const int tc0 = xd_byte( P.tc_luma, ( bs_e - 1 ) & 3 );
                xdf_luma_normal( s[4 * e + 1], s[4 * e + 2], s[4 * e + 3], s[4 * e + 4], s[4 * e + 5], s[4 * e + 6],
                                 P.alpha, P.beta, tc0, normal );
            }
            if( e == 0 )
            {
                const bool strong = act && left_intra;
                if( __any_sync( 0xffffffffu, strong ) )
                    xdf_luma_intra( s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], P.alpha, P.beta, strong );
            }
        }
        if( act )
        {
            const uint32_t row = sL + ( l + 4 ) * DB_PITCH;
            xd_sts_v2( row, xd_pack4( s[0], s[1], s[2], s[3] ), xd_pack4( s[4], s[5], s[6], s[7] ) );
            xd_sts_v2( row + 8, xd_pack4( s[8], s[9], s[10], s[11] ), xd_pack4( s[12], s[13], s[14], s[15] ) );
            xd_sts_u32( row + 16, xd_pack4( s[16], s[17], s[18], s[19] ) );
            if( l < 8 )
            {
                const uint32_t crow = sC + ( l + 2 ) * DB_PITCH;
                xd_sts_v2( crow, S.carry_c, cur_c.x );
                xd_sts_v2( crow + 8, cur_c.y, cur_c.z );
                xd_sts_u32( crow + 16, cur_c.w );
            }
        }
    }
    __syncwarp();
    // The previous iteration's stores are announced HERE, a vertical pass later: by now they have been acknowledged and
    // the release's fence returns at once instead of stalling the warp right after its stores.
    if( lane == 16 && S.publish )
    {
        xd_st_release( mine, S.publish );
        S.publish = 0;
    }
    // ======================= vertical edges, chroma: lane = (row l>>1, component l&1) =======================
    {
        const uint32_t q = sC + ( ( l >> 1 ) + 2 ) * DB_PITCH + 4 + ( l & 1 );
        const int b0 = xd_byte( bsv, 0 ), b2 = xd_byte( bsv, 2 );
        const bool strong = act && left_intra;
        const bool n0 = act && chroma_on && has_left && !left_intra && b0 > 0;
        const bool n1 = act && chroma_on && !first_only && b2 > 0;
        if( __any_sync( 0xffffffffu, strong || n0 || n1 ) )
        {
            int a1 = xd_lds_u8( q - 4 ), a0 = xd_lds_u8( q - 2 ), c0 = xd_lds_u8( q ), c1 = xd_lds_u8( q + 2 );
            int d1 = xd_lds_u8( q + 4 ), d0 = xd_lds_u8( q + 6 ), e0 = xd_lds_u8( q + 8 ), e1 = xd_lds_u8( q + 10 );
            xdf_chroma( a1, a0, c0, c1, P.alphac, P.betac, xd_byte( P.tc_chroma, ( b0 - 1 ) & 3 ) + 1, strong, strong || n0 );
            xdf_chroma( d1, d0, e0, e1, P.alphac, P.betac, xd_byte( P.tc_chroma, ( b2 - 1 ) & 3 ) + 1, false, n1 );
            if( strong || n0 )
            {
                xd_sts_u8( q - 2, a0 );
                xd_sts_u8( q, c0 );
            }
            if( n1 )
            {
                xd_sts_u8( q + 6, d0 );
                xd_sts_u8( q + 8, e0 );
            }
        }
        // the rows above the macroblock join the tile for the horizontal pass
        if( act && mb_y > 0 )
        {
            // tile row l <-> luma row l - 4; column 0 sits 4 bytes into the row (4-byte aligned only)
            const uint32_t t0 = l < 4 ? sL + l * DB_PITCH + 4 : sC + ( l - 4 ) * DB_PITCH + 4;
            if( l < 6 )
            {
                xd_sts_u32( t0, top.x );
                xd_sts_u32( t0 + 4, top.y );
                xd_sts_u32( t0 + 8, top.z );
                xd_sts_u32( t0 + 12, top.w );
            }
        }
    }
    __syncwarp();
    // ======================= horizontal edges, luma: lane = column l, rows -4 .. 15 =======================
    {
        const uint32_t col = sL + 4 + l;
        bool normal[4], any = false;
#pragma unroll
        for( int e = 0; e < 4; e++ )
        {
            normal[e] = act && luma_on && xd_byte( bsh, e ) > 0 && ( e == 0 ? ( mb_y > 0 && !top_intra ) : !first_only );
            any = any || normal[e];
        }
        const bool strong = act && top_intra;
        if( __any_sync( 0xffffffffu, any || strong ) )
        {
            int v[20];
#pragma unroll
            for( int k = 0; k < 20; k++ )
                v[k] = xd_lds_u8( col + k * DB_PITCH );
#pragma unroll
            for( int e = 0; e < 4; e++ )
            {
                if( __any_sync( 0xffffffffu, normal[e] ) )
                {
                    const int tc0 = xd_byte( P.tc_luma, ( xd_byte( bsh, e ) - 1 ) & 3 );
                    xdf_luma_normal( v[4 * e + 1], v[4 * e + 2], v[4 * e + 3], v[4 * e + 4], v[4 * e + 5], v[4 * e + 6],
                                     P.alpha, P.beta, tc0, normal[e] );
                }
                if( e == 0 && __any_sync( 0xffffffffu, strong ) )
                    xdf_luma_intra( v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], P.alpha, P.beta, strong );
            }
            if( any || strong )
            {
#pragma unroll
                for( int k = 1; k < 19; k++ )
                    xd_sts_u8( col + k * DB_PITCH, v[k] );
            }
        }
    }
    // ======================= horizontal edges, chroma: lane = byte column l, rows -2 .. 5 =======================
    {
        const uint32_t col = sC + 4 + l;
        const int b0 = xd_byte( bsh, 0 ), b2 = xd_byte( bsh, 2 );
        const bool strong = act && top_intra;
        const bool n0 = act && chroma_on && mb_y > 0 && !top_intra && b0 > 0;
        const bool n1 = act && chroma_on && !first_only && b2 > 0;
        if( __any_sync( 0xffffffffu, strong || n0 || n1 ) )
        {
            int a1 = xd_lds_u8( col ), a0 = xd_lds_u8( col + DB_PITCH ), c0 = xd_lds_u8( col + 2 * DB_PITCH ), c1 = xd_lds_u8( col + 3 * DB_PITCH );
            int d1 = xd_lds_u8( col + 4 * DB_PITCH ), d0 = xd_lds_u8( col + 5 * DB_PITCH ), e0 = xd_lds_u8( col + 6 * DB_PITCH ), e1 = xd_lds_u8( col + 7 * DB_PITCH );
            xdf_chroma( a1, a0, c0, c1, P.alphac, P.betac, xd_byte( P.tc_chroma, ( b0 - 1 ) & 3 ) + 1, strong, strong || n0 );
            xdf_chroma( d1, d0, e0, e1, P.alphac, P.betac, xd_byte( P.tc_chroma, ( b2 - 1 ) & 3 ) + 1, false, n1 );
            if( strong || n0 )
            {
                xd_sts_u8( col + DB_PITCH, a0 );
                xd_sts_u8( col + 2 * DB_PITCH, c0 );
            }
            if( n1 )
            {
                xd_sts_u8( col + 5 * DB_PITCH, d0 );
                xd_sts_u8( col + 6 * DB_PITCH, e0 );
            }
        }
    }
    __syncwarp();
    // ---- write back what is final now: columns -4 .. 11 of the macroblock's rows (columns 12..15 still face the next
    // macroblock's left edge and travel on in registers), plus the rows above that the top edge touched
    if( act )
    {
        const bool last = EDGE ? x == W - 1 : false;
        {
            const uint32_t src = sL + ( l + 4 ) * DB_PITCH;
            const uint2 w01 = xd_lds_v2( src ), w23 = xd_lds_v2( src + 8 );
            const uint32_t w4 = xd_lds_u32( src + 16 );
            uint32_t *dst = (uint32_t *)( S.py - 4 );
            if( has_left )
                __stcg( dst, w01.x );
            __stcg( dst + 1, w01.y );
            __stcg( dst + 2, w23.x );
            __stcg( dst + 3, w23.y );
            if( last )
                __stcg( dst + 4, w4 );
            S.carry_y = w4;
        }
        if( l < 8 )
        {
            const uint32_t src = sC + ( l + 2 ) * DB_PITCH;
            const uint2 w01 = xd_lds_v2( src ), w23 = xd_lds_v2( src + 8 );
            const uint32_t w4 = xd_lds_u32( src + 16 );
            uint32_t *dst = (uint32_t *)( S.pc - 4 );
            if( has_left )
                __stcg( dst, w01.x );
            __stcg( dst + 1, w01.y );
            __stcg( dst + 2, w23.x );
            __stcg( dst + 3, w23.y );
            if( last )
                __stcg( dst + 4, w4 );
            S.carry_c = w4;
        }
        if( mb_y > 0 )
        {
            // luma rows -3 .. -1 (lanes 0..11: row l>>2, word l&3) and chroma row -1 (lanes 12..15), relative to row 0
            if( l < 12 )
                __stcg( (uint32_t *)( S.py + (int64_t)( ( l >> 2 ) - 3 - l ) * ls ) + ( l & 3 ),
                        xd_lds_u32( sL + ( ( l >> 2 ) + 1 ) * DB_PITCH + 4 + 4 * ( l & 3 ) ) );
            else
                __stcg( (uint32_t *)( S.pc + (int64_t)( -1 - ( l & 7 ) ) * cs ) + ( l - 12 ), xd_lds_u32( sC + DB_PITCH + 4 + 4 * ( l - 12 ) ) );
        }
    }
    __syncwarp();               // every lane's stores are ordered before the release that announces them
    if( lane == 16 && act )
    {
        if( A.latency )
            xd_st_release( mine, x + 1 );             // one frame alone: the row below is waiting for exactly this
        else
            S.publish = x + 1;                        // a batch: announced a vertical pass later, when the fence is free
    }
    if( EDGE ? x >= -1 : true )         // the lower half stands at macroblock -1 during its two idle steps
    {
        S.py += 16;
        S.pc += 16;
    }
}

template<int MINB>
__global__ void __launch_bounds__( DB2_WARPS * 32, MINB )
xd_deblock2_kernel( xd_db_args A )
{
    __shared__ __align__( 16 ) uint8_t s_luma[DB2_WARPS][2][DB_LROWS * DB_PITCH];
    __shared__ __align__( 16 ) uint8_t s_chroma[DB2_WARPS][2][DB_CROWS * DB_PITCH];
    const int lane = threadIdx.x & 31, h = lane >> 4, l = lane & 15, grp = l >> 2;
    const uint32_t sL = (uint32_t)__cvta_generic_to_shared( s_luma[threadIdx.x >> 5][h] );
    const uint32_t sC = (uint32_t)__cvta_generic_to_shared( s_chroma[threadIdx.x >> 5][h] );
    const x264dsp_geom_t &g = A.g;
    const xd_db_params &P = A.P;
    const int W = g.mb_w, H = g.mb_h, ls = g.luma_stride, cs = g.chroma_stride;
    const int pairs = ( H + 1 ) >> 1, total = A.n_frames * pairs;
    const bool luma_on = P.alpha != 0 && P.beta != 0, chroma_on = P.alphac != 0 && P.betac != 0;     // deblock.c:330
    for( ;; )
    {
        int t = 0;
        if( lane == 0 )
            t = atomicAdd( A.ticket, 1 );
        t = __shfl_sync( 0xffffffffu, t, 0 );
        if( t >= total )
            return;
        // pairs are dealt frame-interleaved, top pairs first: pair (f, p) waits on (f, p-1), whose ticket is n_frames
        // smaller and therefore held by a warp that is already running
        const int pr = t / A.n_frames, f = t - pr * A.n_frames;
        const int y0 = pr * 2, mb_y = y0 + h;
        const bool row_ok = mb_y < H;
        uint8_t *slot = A.slots + (size_t)f * g.slot_bytes;
        const size_t mb0 = (size_t)f * g.mb_count;
        const int32_t *above = A.progress + (size_t)f * H + y0 - 1;        // the row over the pair (y0 > 0)
        int32_t *mine = A.progress + (size_t)f * H + y0 + 1;               // the pair's second row

        xd_db2_state S;
        // this lane's rows at macroblock 0 of its row; the lower half starts two macroblocks later, at x = -2 .. -1 its
        // pointers stand still (see the end of xd_db2_step)
        S.py = slot + g.luma_origin + (int64_t)( ( mb_y << 4 ) + l ) * ls;
        S.pc = slot + g.slot_chroma_off + g.chroma_origin + (int64_t)( ( mb_y << 3 ) + ( l & 7 ) ) * cs;
        S.own_y = S.own_c = make_uint4( 0, 0, 0, 0 );
        S.nxt.bsv = S.nxt.bsh = make_uint4( 0, 0, 0, 0 );
        S.nxt.type = S.nxt.type_top = 127; S.nxt.part = 0; S.nxt.cbp = 0;
        S.carry_y = S.carry_c = 0;
        S.intra_prev = false;
        S.seen = 0;
        S.publish = 0;
        if( h == 0 )
        {
            S.own_y = __ldcg( (const uint4 *)S.py );
            if( l < 8 )
                S.own_c = __ldcg( (const uint4 *)S.pc );
            xd_db2_fetch( A, mb0, mb_y * W, mb_y > 0, S.nxt );
        }
        // py / pc stand AT the current macroblock of a step (the step's prefetch reads py + 16).  The lower half idles at
        // x = -2 and x = -1: it waits at macroblock -1, fetches macroblock 0 from there during x = -1, then moves on.
        const int lo = min( 3, W + 2 ), hi = max( lo, W - 1 );
        if( h == 1 )
        {
            S.py -= 16;
            S.pc -= 16;
        }
        for( int i = 0; i < lo; i++ )
            xd_db2_step<true>( A, S, i, lane, h, l, grp, sL, sC, row_ok, mb_y, y0, mb0, above, mine, luma_on, chroma_on );
        for( int i = lo; i < hi; i++ )
            xd_db2_step<false>( A, S, i, lane, h, l, grp, sL, sC, row_ok, mb_y, y0, mb0, above, mine, luma_on, chroma_on );
        for( int i = hi; i < W + 2; i++ )
            xd_db2_step<true>( A, S, i, lane, h, l, grp, sL, sC, row_ok, mb_y, y0, mb0, above, mine, luma_on, chroma_on );
        if( lane == 16 && S.publish )
            xd_st_release( mine, S.publish );         // cumulative: covers the other lanes' stores ordered by the __syncwarp
    }
}

// deblock_strength_c for n macroblocks, one thread per (mb, dir, edge, i), eight macroblocks per block.
// STAGED: the block's 8 x (120 + 80 + 320) input bytes come in as 16-byte loads through shared memory (the per-thread
// byte / short gathers of the direct version ran at 10 % of the HBM peak: 7 us per 1080p frame); needs 16-byte
// aligned arrays, which the host checks.
template<bool STAGED>
__global__ void __launch_bounds__( 256 )
xd_deblock_strength_kernel( int n, const int8_t *__restrict__ mb_type, const uint8_t *__restrict__ nnz,
                            const int8_t *__restrict__ ref, const int16_t *__restrict__ mv, uint8_t *__restrict__ bs )
{
    __shared__ __align__( 16 ) uint8_t s_nnz[8 * 120];
    __shared__ __align__( 16 ) int8_t s_ref[8 * 80];
    __shared__ __align__( 16 ) int16_t s_mv[8 * 160];
    const int m0 = blockIdx.x * 8;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = t >> 5, k = t & 31;
    const uint8_t *z;
    const int8_t *r;
    const int16_t *v;
    if( STAGED )
    {
        const int nm = min( 8, n - m0 );                        // macroblocks of this block
        const int q_nnz = nm * 120 / 16 + ( ( nm * 120 ) % 16 ? 1 : 0 ), q_ref = nm * 80 / 16, q_mv = nm * 320 / 16;
        // nm * 120 is a multiple of 16 only for even nm; an odd tail reads one quad past its last macroblock's nnz,
        // which stays inside the array unless this is the very last macroblock: that quad is fetched bytewise
        for( int i = threadIdx.x; i < q_nnz + q_ref + q_mv; i += blockDim.x )
        {
            if( i < q_nnz )
            {
                if( ( i + 1 ) * 16 <= nm * 120 )
                    ( (uint4 *)s_nnz )[i] = __ldg( (const uint4 *)( nnz + (size_t)m0 * 120 ) + i );
                else
                    for( int b = i * 16; b < nm * 120; b++ )
                        s_nnz[b] = __ldg( nnz + (size_t)m0 * 120 + b );
            }
            else if( i < q_nnz + q_ref )
                ( (uint4 *)s_ref )[i - q_nnz] = __ldg( (const uint4 *)( ref + (size_t)m0 * 80 ) + ( i - q_nnz ) );
            else
                ( (uint4 *)s_mv )[i - q_nnz - q_ref] = __ldg( (const uint4 *)( mv + (size_t)m0 * 160 ) + ( i - q_nnz - q_ref ) );
        }
        __syncthreads();
        if( m >= n )
            return;
        z = s_nnz + ( m - m0 ) * 120;
        r = s_ref + ( m - m0 ) * 80;
        v = s_mv + ( m - m0 ) * 160;
    }
    else
    {
        if( m >= n )
            return;
        z = nnz + (size_t)m * 120;
        r = ref + (size_t)m * 80;
        v = mv + (size_t)m * 160;
    }
    const int dir = k >> 4, edge = ( k >> 2 ) & 3, i = k & 3;
    if( mb_type )
    {
        // x264_macroblock_deblock_strength (common/macroblock.c:677-691): an intra macroblock sets its three inner
        // edges to 3 and leaves bs[dir][0] alone (the outer edges of an intra macroblock are filtered with bS 4 by
        // x264_frame_deblock_row whatever is stored there)
        const int type = __ldg( mb_type + m );
        if( type >= 0 && type < 4 )                          // IS_INTRA: I_4x4, I_8x8, I_16x16, I_PCM
        {
            if( edge )
                bs[(size_t)m * 64 + dir * 32 + edge * 4 + i] = 3;
            return;
        }
    }
    const int along = dir ? 1 : 8, across = dir ? 8 : 1;
    const int cur = 12 + edge * across + i * along, nb = cur - across;
    int s;
    if( z[cur] || z[nb] )
        s = 2;
    else if( r[cur] != r[nb] || abs( v[2 * cur] - v[2 * nb] ) >= 4 || abs( v[2 * cur + 1] - v[2 * nb + 1] ) >= 4 )
        s = 1;
    else
        s = 0;
    bs[(size_t)m * 64 + dir * 32 + edge * 4 + i] = (uint8_t)s;
}

extern "C" int x264dsp_deblock_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slots, int n_frames,
                                            const int8_t *mb_type, const uint8_t *partition, const int16_t *cbp,
                                            const uint8_t *bs, int qp, int alpha_c0_offset, int beta_offset,
                                            void *stream )
{
    if( !ctx || !g || !slots || n_frames <= 0 || !mb_type || !partition || !cbp || !bs || qp < 0 || qp > 51 )
        return X264DSP_E_ARG;
    static const uint8_t alpha_h[52] =
    {
        0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 4,4,5,6,7,8,9,10,12,13,15,17,20,22,
        25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255
    };
    static const uint8_t beta_h[52] =
    {
        0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 2,2,2,3,3,3,3,4,4,4,6,6,7,7,
        8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18
    };
    // the reference's tables carry guard entries: below 0 everything is 0, above 51 the last value
    auto clamp_idx = []( int i ) { return i > 51 ? 51 : i; };
    const int qpc = x264dsp_chroma_qp( qp );
    const int ia = qp + alpha_c0_offset, ib = qp + beta_offset, iac = qpc + alpha_c0_offset, ibc = qpc + beta_offset;
    xd_db_args A;
    A.g = *g;
    A.slots = slots;
    A.n_frames = n_frames;
    A.mb_type = mb_type; A.partition = partition; A.cbp = cbp; A.bs = bs;
    A.P.alpha = ia < 0 ? 0 : alpha_h[clamp_idx( ia )];
    A.P.beta = ib < 0 ? 0 : beta_h[clamp_idx( ib )];
    A.P.alphac = iac < 0 ? 0 : alpha_h[clamp_idx( iac )];
    A.P.betac = ibc < 0 ? 0 : beta_h[clamp_idx( ibc )];
    A.P.ia = ia < 0 ? -1 : clamp_idx( ia );
    A.P.iac = iac < 0 ? -1 : clamp_idx( iac );
    A.P.tc_luma = xd_tc0_packed( A.P.ia );
    A.P.tc_chroma = xd_tc0_packed( A.P.iac );

    cudaStream_t s = xd_stream( ctx, stream );
    // progress counters: one per MB row of every frame, plus the ticket
    const size_t rows = (size_t)n_frames * g->mb_h;
    const size_t need = ( rows + 1 ) * sizeof( int32_t );
    if( ctx->db_progress_cap < need )
        XD_CHECK( cudaDeviceSynchronize() );
    int rc = xd_reserve_dev( (void **)&ctx->db_progress, &ctx->db_progress_cap, need );
    if( rc )
        return rc;
    // the progress counters are the context's: a call on another stream queues behind the previous one
    if( ( rc = xd_scratch_acquire( ctx, XD_SCRATCH_DEBLOCK, s ) ) )
        return rc;
    XD_CHECK( cudaMemsetAsync( ctx->db_progress, 0, need, s ) );
    A.progress = ctx->db_progress;
    A.ticket = ctx->db_progress + rows;
    // two mappings of the same wavefront with identical results (see the kernels): a warp per row pair (default) and the
    // older warp per row (X264DSP_DEBLOCK_V1=1, or when the bS array is not 16-byte aligned)
    static int use_v1 = -1;
    if( use_v1 < 0 )
    {
        const char *e = getenv( "X264DSP_DEBLOCK_V1" );
        use_v1 = e && atoi( e ) > 0;
    }
    const bool v1 = use_v1 || ( (uintptr_t)bs & 15 ) != 0;
    static int db2_minb = 0;                             // tuning knob: resident CTAs per SM the kernel is compiled for
    if( !db2_minb )
    {
        const char *e = getenv( "X264DSP_DB2_MINB" );
        db2_minb = e && atoi( e ) == 6 ? 6 : 5;
    }
    static int per_sm[2] = { 0, 0 };
    if( !per_sm[v1] )
    {
        if( v1 )
            XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[1], xd_deblock_kernel, DB_WARPS * 32, 0 ) );
        else if( db2_minb == 6 )
            XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[0], xd_deblock2_kernel<6>, DB2_WARPS * 32, 0 ) );
        else
            XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm[0], xd_deblock2_kernel<5>, DB2_WARPS * 32, 0 ) );
        if( per_sm[v1] < 1 )
            per_sm[v1] = 1;
    }
    const int64_t units = v1 ? (int64_t)rows : (int64_t)n_frames * ( ( g->mb_h + 1 ) / 2 );
    A.latency = units * 2 <= (int64_t)ctx->sm_count * per_sm[v1] * DB2_WARPS;       // the machine is mostly empty
    int64_t ctas = ( units + DB_WARPS - 1 ) / DB_WARPS;
    if( ctas > (int64_t)ctx->sm_count * per_sm[v1] )
        ctas = (int64_t)ctx->sm_count * per_sm[v1];      // persistent: the ticket hands out the remaining units
    const int pslot = xd_prof_begin( ctx, XD_PROF_DEBLOCK, s );
    if( v1 )
        xd_deblock_kernel<<<(int)ctas, DB_WARPS * 32, 0, s>>>( A );
    else if( db2_minb == 6 )
        xd_deblock2_kernel<6><<<(int)ctas, DB2_WARPS * 32, 0, s>>>( A );
    else
        xd_deblock2_kernel<5><<<(int)ctas, DB2_WARPS * 32, 0, s>>>( A );
    xd_prof_end( ctx, XD_PROF_DEBLOCK, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return xd_scratch_release( ctx, XD_SCRATCH_DEBLOCK, s );
}

extern "C" int x264dsp_deblock_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slot,
                                           const int8_t *mb_type, const uint8_t *partition, const int16_t *cbp,
                                           const uint8_t *bs, int qp, int alpha_c0_offset, int beta_offset,
                                           void *stream )
{
    return x264dsp_deblock_frames_dev( ctx, g, slot, 1, mb_type, partition, cbp, bs, qp, alpha_c0_offset, beta_offset, stream );
}

extern "C" int x264dsp_macroblock_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const int8_t *mb_type, const uint8_t *nnz,
                                                         const int8_t *ref, const int16_t *mv, uint8_t *bs, void *stream )
{
    if( !ctx || n < 0 )
        return X264DSP_E_ARG;
    if( n == 0 )
        return 0;
    if( !nnz || !ref || !mv || !bs )
        return X264DSP_E_ARG;
    const int64_t threads = (int64_t)n * 32;
    const bool aligned = ( ( (uintptr_t)nnz | (uintptr_t)ref | (uintptr_t)mv ) & 15 ) == 0;
    if( aligned )
        xd_deblock_strength_kernel<true><<<(int)( ( threads + 255 ) / 256 ), 256, 0, xd_stream( ctx, stream )>>>( n, mb_type, nnz, ref, mv, bs );
    else
        xd_deblock_strength_kernel<false><<<(int)( ( threads + 255 ) / 256 ), 256, 0, xd_stream( ctx, stream )>>>( n, mb_type, nnz, ref, mv, bs );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const uint8_t *nnz, const int8_t *ref,
                                              const int16_t *mv, uint8_t *bs, void *stream )
{
    return x264dsp_macroblock_deblock_strength_dev( ctx, n, nullptr, nnz, ref, mv, bs, stream );
}
