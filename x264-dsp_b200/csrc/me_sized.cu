// me_sized.cu -- motion search of block lists of ONE partition size: G lanes per block, sm_100a.
//
// Reference: x264_me_search_ref (encoder/me.c:129-423), refine_subpel (me.c:466-587),
// x264_me_refine_qpel (me.c:426-435), get_ref plane selection (common/mc.c:192-264).
//
// me.cu gives every block a whole warp (four candidates at a time), which suits single large
// blocks but leaves most lanes idle on the small partitions that dominate a full-frame search
// (a 1080p frame has 130 560 4x4 blocks).  Here the block size is a template parameter and the
// block is cut into its Hadamard base tiles (8x4, or 4x4 for the 4-wide sizes -- the reference's
// own SATD units, pixel.c:267-337): G = tiles-per-block lanes cooperate on one block, each lane
// owning one tile whose source pixels stay in 8 registers for the whole search.  So a 16x16 block
// takes 8 lanes, 8x8 two, 8x4 and 4x4 a single thread, and a warp carries 4 .. 32 blocks.
// Candidates are evaluated one after the other in the reference's order with its strict '<'
// updates, so tie-breaking is the reference's by construction (evaluating the three or four candidates of a step
// together and driving the steps from a state machine with one inlined cost site was tried in round 2: bit-exact, 18 %
// fewer instructions, 165-177 us against 143 us per frame for all sizes -- the rolled loops below stay); the G lanes of a block add their
// partial costs with xor-shuffles under the group's own lane mask, which lets the blocks of a
// warp diverge freely (different search lengths) without any warp-wide synchronisation.
#include "common.cuh"
#include "leaf.cuh"

int xd_me_params_ok( const x264dsp_me_params_t *p );   // me.cu

#define MES_COST_MAX ( 1 << 28 )
#define MES_THREADS 128

template<int W, int H>
struct xd_mes_cfg
{
    static constexpr int TW = W >= 8 ? 8 : 4;                 // tile width
    static constexpr int TILES = ( W / TW ) * ( H / 4 );
    static constexpr int G = W == 4 ? ( H >= 16 ? 2 : 1 ) : TILES;      // lanes per block
    static constexpr int NT = TILES / G;                      // tiles per lane
    static constexpr int WORDS = TW / 4;                      // 32-bit words per tile row
};

struct xd_mes_blk
{
    const uint8_t *ref;             // reference plane N at the block position
    size_t plane_size;
    int stride;
    const uint16_t *cost_mv;        // centre
    int mvpx, mvpy;
    int minx, miny, maxx, maxy;     // full-pel limits
    int sminx, sminy, smaxx, smaxy; // sub-pel limits
    bool fpel_satd;                 // fpelcmp == satd: me=TESA with subme >= 2 (encoder/encoder.c:412-432)
};

__device__ __forceinline__ int xd_mes_bits( const xd_mes_blk &B, int qx, int qy )
{
    return __ldg( B.cost_mv + ( qx - B.mvpx ) ) + __ldg( B.cost_mv + ( qy - B.mvpy ) );
}

// sum over the G lanes of a block (G is a power of two, groups are G-aligned)
template<int G>
__device__ __forceinline__ int xd_mes_group_sum( int v, unsigned gmask )
{
#pragma unroll
    for( int o = 1; o < G; o <<= 1 )
        v += __shfl_xor_sync( gmask, v, o );
    return v;
}

// cost of the block at quarter-pel (qx,qy): SAD or SATD of this lane's tile(s), summed over the group.
// f[] holds the lane's source pixels: tile t, row r, word w at f[( t * 4 + r ) * WORDS + w].
template<int W, int H>
__device__ __forceinline__ int xd_mes_cost( const xd_mes_blk &B, const uint32_t *f, int tile0, int qx, int qy,
                                            bool satd, unsigned gmask )
{
    typedef xd_mes_cfg<W, H> C;
    const int fx = qx & 3, fy = qy & 3, phase = fy * 4 + fx;
    const int64_t base = (int64_t)( qy >> 2 ) * B.stride + ( qx >> 2 );
    const uint8_t *pa = B.ref + (size_t)xd_qpel_plane_a( phase ) * B.plane_size + base + ( fy == 3 ? B.stride : 0 );
    const bool two = ( phase & 5 ) != 0;
    const uint8_t *pb = B.ref + (size_t)xd_qpel_plane_b( phase ) * B.plane_size + base + ( fx == 3 ? 1 : 0 );
    int acc = 0;
#pragma unroll
    for( int t = 0; t < C::NT; t++ )
    {
        const int tile = tile0 + t;
        const int tx = ( tile % ( W / C::TW ) ) * C::TW, ty = ( tile / ( W / C::TW ) ) * 4;
        uint32_t p[4][C::WORDS];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const int64_t o = (int64_t)( ty + r ) * B.stride + tx;
            if( C::WORDS == 2 )
            {
                uint2 a = xd_load8_unaligned( pa + o );
                if( two )
                {
                    const uint2 b = xd_load8_unaligned( pb + o );
                    a.x = xd_avg4( a.x, b.x );
                    a.y = xd_avg4( a.y, b.y );
                }
                p[r][0] = a.x;
                p[r][C::WORDS - 1] = a.y;
            }
            else
            {
                uint32_t a = xd_load4_unaligned( pa + o );
                if( two )
                    a = xd_avg4( a, xd_load4_unaligned( pb + o ) );
                p[r][0] = a;
            }
        }
        const uint32_t *ft = f + t * 4 * C::WORDS;
        if( satd )
        {
            int s = 0;
#pragma unroll
            for( int w = 0; w < C::WORDS; w++ )
            {
                const uint32_t a[4] = { ft[w], ft[C::WORDS + w], ft[2 * C::WORDS + w], ft[3 * C::WORDS + w] };
                const uint32_t b[4] = { p[0][w], p[1][w], p[2][w], p[3][w] };
                s += xd_satd4x4( a, b );
            }
            acc += s;
        }
        else
        {
#pragma unroll
            for( int r = 0; r < 4; r++ )
#pragma unroll
                for( int w = 0; w < C::WORDS; w++ )
                    acc += __vsadu4( p[r][w], ft[r * C::WORDS + w] );
        }
    }
    return xd_mes_group_sum<C::G>( acc, gmask );
}

// CHECK_MVRANGE (me.c:155-160)
__device__ __forceinline__ bool xd_mes_in_range( const xd_mes_blk &B, int mx, int my )
{
    const uint32_t lo = ( (uint32_t)( -B.minx ) << 16 ) | ( (uint32_t)( -B.miny ) & 0x7FFF );
    const uint32_t hi = ( (uint32_t)B.maxx << 16 ) | ( (uint32_t)B.maxy & 0x7FFF ) | 0x8000;
    const uint32_t v = ( (uint32_t)mx << 16 ) | ( (uint32_t)my & 0x7FFF );
    return !( ( ( v + lo ) | ( hi - v ) ) & 0x80004000u );
}

__device__ __forceinline__ uint32_t xd_mes_pack( int x, int y )
{
    return ( (uint32_t)x & 0xFFFF ) | ( (uint32_t)y << 16 );
}

// { refine_hpel, refine_qpel, me_hpel, me_qpel } (me.c:18-32), subme 0..5
__constant__ uint8_t xd_mes_iters[6][4] = { {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2} };

// the four diamond neighbours in the reference's order: up, down, left, right
__constant__ int8_t xd_mes_dia[4][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } };
// hex2[] of me.c:47 and square1[1..8] of me.c:48-51
__constant__ int8_t xd_mes_hex2[8][2] = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
__constant__ int8_t xd_mes_square[8][2] = { {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };

struct xd_mes_state
{
    int mvx, mvy, cost, cost_mv;
};

#define MES_SAD( qx, qy ) ( xd_mes_cost<W, H>( B, f, tile0, ( qx ), ( qy ), FS, gmask ) + xd_mes_bits( B, ( qx ), ( qy ) ) )
#define MES_SATD( qx, qy ) ( xd_mes_cost<W, H>( B, f, tile0, ( qx ), ( qy ), true, gmask ) + xd_mes_bits( B, ( qx ), ( qy ) ) )

// refine_subpel (me.c:466-587) with p_halfpel_thresh == NULL
template<int W, int H, bool FS>
__device__ void xd_mes_refine( const xd_mes_blk &B, const uint32_t *f, int tile0, unsigned gmask, xd_mes_state &S,
                               int subme, int hpel_iters, int qpel_iters, bool final_refine )
{
    int bmx = S.mvx, bmy = S.mvy, bcost = S.cost;
    if( hpel_iters && subme < 3 )                                        // me.c:483-490
    {
        const int px = xd_clip3( B.mvpx, B.sminx + 2, B.smaxx - 2 ), py = xd_clip3( B.mvpy, B.sminy + 2, B.smaxy - 2 );
        if( px != bmx || py != bmy )
        {
            const int c = MES_SAD( px, py );
            if( c < bcost ) { bcost = c; bmx = px; bmy = py; }
        }
    }
    for( int i = hpel_iters; i > 0; i-- )                                // me.c:492-517
    {
        const int omx = bmx, omy = bmy;
#pragma unroll 1
        for( int d = 0; d < 4; d++ )
        {
            const int qx = omx + 2 * xd_mes_dia[d][0], qy = omy + 2 * xd_mes_dia[d][1];
            const int c = MES_SAD( qx, qy );
            if( c < bcost ) { bcost = c; bmx = qx; bmy = qy; }
        }
        if( bmx == omx && bmy == omy )
            break;
    }
    if( !final_refine && !FS )                                           // me.c:519-524 (mbcmp_unaligned != fpelcmp)
        bcost = MES_SATD( bmx, bmy );

    if( subme != 1 )
    {
        int bdir = -1;                                                   // me.c:541-564
        for( int i = qpel_iters; i > 0; i-- )
        {
            if( bmy <= B.sminy || bmy >= B.smaxy || bmx <= B.sminx || bmx >= B.smaxx )
                break;
            const int odir = bdir, omx = bmx, omy = bmy;
#pragma unroll 1
            for( int d = 0; d < 4; d++ )
            {
                if( !final_refine && ( d ^ 1 ) == odir )
                    continue;
                const int qx = omx + xd_mes_dia[d][0], qy = omy + xd_mes_dia[d][1];
                const int c = MES_SATD( qx, qy );
                if( c < bcost ) { bcost = c; bmx = qx; bmy = qy; bdir = d; }
            }
            if( bmx == omx && bmy == omy )
                break;
        }
    }
    else if( bmy > B.sminy && bmy < B.smaxy && bmx > B.sminx && bmx < B.smaxx )   // me.c:565-581
    {
        const int omx = bmx, omy = bmy;
#pragma unroll 1
        for( int d = 0; d < 4; d++ )
        {
            const int qx = omx + xd_mes_dia[d][0], qy = omy + xd_mes_dia[d][1];
            const int c = MES_SAD( qx, qy );
            if( c < bcost ) { bcost = c; bmx = qx; bmy = qy; }
        }
    }
    S.cost = bcost;
    S.mvx = bmx;
    S.mvy = bmy;
    S.cost_mv = xd_mes_bits( B, bmx, bmy );
}

// FS: the full-pel metric is SATD (me = TESA with subme >= 2); a compile-time property so that the ordinary search
// carries none of it
template<int W, int H, bool FS>
__global__ void __launch_bounds__( MES_THREADS )
xd_me_sized_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fenc_slot, const uint8_t *__restrict__ fref_slot,
                    x264dsp_me_params_t P, const uint16_t *__restrict__ cost_mv, int n,
                    const x264dsp_me_block_t *__restrict__ blocks, x264dsp_me_result_t *__restrict__ results )
{
    typedef xd_mes_cfg<W, H> C;
    const int tid = blockIdx.x * MES_THREADS + threadIdx.x;
    const int blk = tid / C::G, sub = tid % C::G;
    if( blk >= n )
        return;                     // whole groups leave together: n is counted in blocks, G divides the CTA
    // blockIdx.y = frame pair of a batch: consecutive slots, n blocks / results per frame
    fenc_slot += blockIdx.y * (size_t)g.slot_bytes;
    fref_slot += blockIdx.y * (size_t)g.slot_bytes;
    blocks += blockIdx.y * (size_t)n;
    results += blockIdx.y * (size_t)n;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = C::G == 32 ? 0xffffffffu : ( ( 1u << C::G ) - 1u ) << ( lane & ~( C::G - 1 ) );
    const x264dsp_me_block_t *in = blocks + blk;

    xd_mes_blk B;
    B.stride = g.luma_stride;
    B.plane_size = (size_t)g.luma_plane_size;
    const int64_t pos = (int64_t)in->by * g.luma_stride + in->bx;
    B.ref = fref_slot + g.luma_origin + pos;
    B.cost_mv = cost_mv;
    B.mvpx = in->mvp[0];
    B.mvpy = in->mvp[1];
    B.minx = in->mv_min_fpel[0]; B.miny = in->mv_min_fpel[1];
    B.maxx = in->mv_max_fpel[0]; B.maxy = in->mv_max_fpel[1];
    B.sminx = in->mv_min_spel[0]; B.sminy = in->mv_min_spel[1];
    B.smaxx = in->mv_max_spel[0]; B.smaxy = in->mv_max_spel[1];
    const int n_mvc = min( max( in->i_mvc, 0 ), 16 );
    const int subme = P.subpel_refine;
    B.fpel_satd = FS;

    // this lane's source tile(s) stay in registers for the whole search
    const int tile0 = sub * C::NT;
    uint32_t f[C::NT * 4 * C::WORDS];
    {
        const uint8_t *fenc = fenc_slot + g.luma_origin + pos;
#pragma unroll
        for( int t = 0; t < C::NT; t++ )
        {
            const int tile = tile0 + t;
            const int tx = ( tile % ( W / C::TW ) ) * C::TW, ty = ( tile / ( W / C::TW ) ) * 4;
#pragma unroll
            for( int r = 0; r < 4; r++ )
            {
                const uint8_t *p = fenc + (int64_t)( ty + r ) * g.luma_stride + tx;
                if( C::WORDS == 2 )
                {
                    const uint2 v = xd_load8_unaligned( p );
                    f[( t * 4 + r ) * C::WORDS] = v.x;
                    f[( t * 4 + r ) * C::WORDS + C::WORDS - 1] = v.y;
                }
                else
                    f[( t * 4 + r ) * C::WORDS] = xd_load4_unaligned( p );
            }
        }
    }

    int bmx = xd_clip3( B.mvpx, B.minx * 4, B.maxx * 4 ), bmy = xd_clip3( B.mvpy, B.miny * 4, B.maxy * 4 );
    const int pmx = ( bmx + 2 ) >> 2, pmy = ( bmy + 2 ) >> 2;
    int bcost = MES_COST_MAX;
    int pred_mx = 0, pred_my = 0, pred_cost = MES_COST_MAX;
    uint32_t pmv;

    if( subme >= 3 )
    {
        // me.c:176-193: sub-pel predictors, evaluated at quarter-pel positions
        pmv = xd_mes_pack( bmx, bmy );
        if( n_mvc > 0 )
        {
            const int c = MES_SAD( bmx, bmy );
            if( c < pred_cost ) { pred_cost = c; pred_mx = bmx; pred_my = bmy; }
        }
        for( int i = 0; i < n_mvc; i++ )
        {
            const int cx = in->mvc[i][0], cy = in->mvc[i][1];
            const uint32_t raw = xd_mes_pack( cx, cy );
            if( raw != 0 && raw != pmv )
            {
                const int qx = xd_clip3( cx, B.minx * 4, B.maxx * 4 ), qy = xd_clip3( cy, B.miny * 4, B.maxy * 4 );
                const int c = MES_SAD( qx, qy );
                if( c < pred_cost ) { pred_cost = c; pred_mx = qx; pred_my = qy; }
            }
        }
        bmx = ( pred_mx + 2 ) >> 2;
        bmy = ( pred_my + 2 ) >> 2;
        bcost = MES_SAD( bmx << 2, bmy << 2 );
    }
    else
    {
        // me.c:194-229: rounded MVP without mv cost, then the rounded / clipped candidates
        bmx = pmx;
        bmy = pmy;
        bcost = xd_mes_cost<W, H>( B, f, tile0, pmx << 2, pmy << 2, FS, gmask );
        pmv = xd_mes_pack( pmx, pmy );
        int sel_x = pmx, sel_y = pmy;
        for( int i = 0; i < n_mvc; i++ )
        {
            const int fx = xd_clip3( ( in->mvc[i][0] + 2 ) >> 2, B.minx, B.maxx );
            const int fy = xd_clip3( ( in->mvc[i][1] + 2 ) >> 2, B.miny, B.maxy );
            const uint32_t v = xd_mes_pack( fx, fy );
            if( v != 0 && v != pmv )
            {
                const int c = MES_SAD( fx << 2, fy << 2 );
                if( c < bcost ) { bcost = c; sel_x = fx; sel_y = fy; }
            }
        }
        bmx = sel_x;
        bmy = sel_y;
    }
    if( pmv )                                                            // me.c:231-233
    {
        const int c = MES_SAD( 0, 0 );
        if( c < bcost ) { bcost = c; bmx = 0; bmy = 0; }
    }

    if( P.me_method == X264DSP_ME_DIA )
    {
        // me.c:237-274
        int left = P.me_range;
        do
        {
            const int omx = bmx, omy = bmy;
            int best = -1;
#pragma unroll 1
            for( int d = 0; d < 4; d++ )
            {
                const int c = MES_SAD( ( omx + xd_mes_dia[d][0] ) << 2, ( omy + xd_mes_dia[d][1] ) << 2 );
                if( c < bcost ) { bcost = c; best = d; }
            }
            if( best < 0 )
                break;
            bmx = omx + xd_mes_dia[best][0];
            bmy = omy + xd_mes_dia[best][1];
        } while( --left && xd_mes_in_range( B, bmx, bmy ) );
    }
    else if( P.me_method == X264DSP_ME_HEX )
    {
        // me.c:276-388: hexagon, then square refinement
        int dir = -1;
        {
            const int omx = bmx, omy = bmy;
#pragma unroll 1
            for( int k = 0; k < 6; k++ )                                 // A B C D E F = hex2[1..6]
            {
                const int c = MES_SAD( ( omx + xd_mes_hex2[k + 1][0] ) << 2, ( omy + xd_mes_hex2[k + 1][1] ) << 2 );
                if( c < bcost ) { bcost = c; dir = k; }
            }
        }
        if( dir >= 0 )
        {
            bmx += xd_mes_hex2[dir + 1][0];
            bmy += xd_mes_hex2[dir + 1][1];
            for( int left = ( P.me_range >> 1 ) - 1; left > 0 && xd_mes_in_range( B, bmx, bmy ); left-- )
            {
                int step = -1;
#pragma unroll 1
                for( int k = 0; k < 3; k++ )                             // hex2[dir+0], [dir+1], [dir+2]
                {
                    const int c = MES_SAD( ( bmx + xd_mes_hex2[dir + k][0] ) << 2, ( bmy + xd_mes_hex2[dir + k][1] ) << 2 );
                    if( c < bcost ) { bcost = c; step = k; }
                }
                if( step < 0 )
                    break;
                dir += step - 1;
                dir = dir < 0 ? dir + 6 : dir >= 6 ? dir - 6 : dir;      // mod6m1
                bmx += xd_mes_hex2[dir + 1][0];
                bmy += xd_mes_hex2[dir + 1][1];
            }
        }
        // square refine (me.c:361-386)
        int sq = -1;
#pragma unroll 1
        for( int k = 0; k < 8; k++ )
        {
            const int c = MES_SAD( ( bmx + xd_mes_square[k][0] ) << 2, ( bmy + xd_mes_square[k][1] ) << 2 );
            if( c < bcost ) { bcost = c; sq = k; }
        }
        if( sq >= 0 )
        {
            bmx += xd_mes_square[sq][0];
            bmy += xd_mes_square[sq][1];
        }
    }
    // UMH / ESA / TESA: no case in the reference's switch (me.c:389-394) -- predictors and sub-pel refinement only

    // me.c:397-414
    xd_mes_state S;
    if( pred_cost < bcost )
    {
        S.mvx = pred_mx; S.mvy = pred_my; S.cost = pred_cost;
    }
    else
    {
        S.mvx = bmx << 2; S.mvy = bmy << 2; S.cost = bcost;
    }
    S.cost_mv = xd_mes_bits( B, S.mvx, S.mvy );
    if( bmx == pmx && bmy == pmy && subme < 3 )
        S.cost += S.cost_mv;

    if( subme >= 2 )
        xd_mes_refine<W, H, FS>( B, f, tile0, gmask, S, subme, xd_mes_iters[subme][2], xd_mes_iters[subme][3], false );
    if( P.refine_qpel )                                                  // me.c:426-435, i_ref_cost = 0
        xd_mes_refine<W, H, FS>( B, f, tile0, gmask, S, subme, xd_mes_iters[subme][0], xd_mes_iters[subme][1], true );

    if( sub == 0 )
    {
        x264dsp_me_result_t r;
        r.mv[0] = (int16_t)S.mvx;
        r.mv[1] = (int16_t)S.mvy;
        r.cost = S.cost;
        r.cost_mv = S.cost_mv;
        results[blk] = r;
    }
}

template<int W, int H>
static int xd_mes_launch( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                          const x264dsp_me_params_t *params, int n_frames, int n, const x264dsp_me_block_t *blocks,
                          x264dsp_me_result_t *results, cudaStream_t s )
{
    const int64_t threads = (int64_t)n * xd_mes_cfg<W, H>::G;
    const dim3 grid( (unsigned)( ( threads + MES_THREADS - 1 ) / MES_THREADS ), n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_ME, s );
    if( params->me_method == X264DSP_ME_TESA && params->subpel_refine >= 2 )
        xd_me_sized_kernel<W, H, true><<<grid, MES_THREADS, 0, s>>>( *g, fenc_slot, fref_slot, *params,
                                                                     ctx->cost_mv_dev[params->qp] + 4096, n, blocks, results );
    else
        xd_me_sized_kernel<W, H, false><<<grid, MES_THREADS, 0, s>>>( *g, fenc_slot, fref_slot, *params,
                                                                      ctx->cost_mv_dev[params->qp] + 4096, n, blocks, results );
    xd_prof_end( ctx, XD_PROF_ME, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_me_search_sized_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                                    const uint8_t *fenc_slot, const uint8_t *fref_slot, int n_frames,
                                                    const x264dsp_me_params_t *params, int i_pixel, int n,
                                                    const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                                    void *stream )
{
    if( !ctx || !g || !fenc_slot || !fref_slot || !params || n < 0 || i_pixel < 0 || i_pixel > 7 || n_frames <= 0
        || n_frames > 65535 )
        return X264DSP_E_ARG;
    if( !xd_me_params_ok( params ) )
        return X264DSP_E_ARG;                        // subme 0 does not exist in the reference
    if( n == 0 )
        return 0;
    if( !blocks || !results )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    switch( i_pixel )
    {
    case X264DSP_PIXEL_16x16: return xd_mes_launch<16, 16>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_16x8:  return xd_mes_launch<16, 8>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_8x16:  return xd_mes_launch<8, 16>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_8x8:   return xd_mes_launch<8, 8>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_8x4:   return xd_mes_launch<8, 4>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_4x8:   return xd_mes_launch<4, 8>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    case X264DSP_PIXEL_4x4:   return xd_mes_launch<4, 4>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    default:                  return xd_mes_launch<4, 16>( ctx, g, fenc_slot, fref_slot, params, n_frames, n, blocks, results, s );
    }
}

extern "C" int x264dsp_me_search_sized_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                             const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                             const x264dsp_me_params_t *params, int i_pixel, int n,
                                             const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                             void *stream )
{
    return x264dsp_me_search_sized_frames_dev( ctx, g, fenc_slot, fref_slot, 1, params, i_pixel, n, blocks, results, stream );
}
