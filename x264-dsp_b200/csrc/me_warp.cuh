// me_warp.cuh -- x264_me_search_ref / refine_subpel / x264_me_refine_qpel for ONE block by one warp (device routine).
//
// Reference: x264_me_search_ref (encoder/me.c:129-423), refine_subpel (me.c:466-587),
// x264_me_refine_qpel (me.c:426-435), get_ref / mc_luma plane selection (common/mc.c:192-264).
//
// Shared by the batched search kernel (me.cu: one warp per list entry) and the P-slice wavefront (pframe.cu: the warp that
// walks a macroblock row searches each macroblock as it gets to it).  The mapping is described in me.cu.
#pragma once
#include "common.cuh"
#include "leaf.cuh"

#define ME_COST_MAX ( 1 << 28 )
// XD_ME_CALL / XD_ME_REFINE_CALL: __noinline__ makes the cost routines / the sub-pel refinement real calls (small code),
// __forceinline__ / nothing inlines them at every site (see the note at xd_me_sad_call)
#ifndef XD_ME_CALL
#define XD_ME_CALL __forceinline__
#endif
#ifndef XD_ME_SAD16_UNROLL
#define XD_ME_SAD16_UNROLL 1
#endif
constexpr int xd_me_sad16_unroll = XD_ME_SAD16_UNROLL;
#ifndef XD_ME_REFINE_CALL
#define XD_ME_REFINE_CALL
#endif

struct xd_me_blk
{
    const uint8_t *fenc;            // source block, luma plane of the source slot
    const uint8_t *ref;             // reference plane N at the block position
    size_t plane_size;
    int stride;
    int w, h;
    const uint16_t *cost_mv;        // centre
    int mvpx, mvpy;
    int minx, miny, maxx, maxy;     // full-pel limits
    int sminx, sminy, smaxx, smaxy; // sub-pel limits
    bool fpel_satd;                 // h->pixf.fpelcmp == satd: me=TESA with subme >= 2 (encoder/encoder.c:412-432)
    bool fixed16;                   // compile-time knowledge that the block is 16x16 (xd_me_search_warp<true>): unrolled cost loops
    bool src_aligned;               // ... that the source block sits at a multiple of min(width, 8) (xd_me_search_warp<.., true>)
};

__device__ __forceinline__ int xd_me_bits( const xd_me_blk &B, int qx, int qy )
{
    return __ldg( B.cost_mv + ( qx - B.mvpx ) ) + __ldg( B.cost_mv + ( qy - B.mvpy ) );
}

// plane addresses for a quarter-pel position (common/mc.c:222-234)
struct xd_qpel_src
{
    const uint8_t *a, *b;           // b == nullptr: no averaging
};
__device__ __forceinline__ xd_qpel_src xd_me_src( const xd_me_blk &B, int qx, int qy )
{
    const int fx = qx & 3, fy = qy & 3, phase = fy * 4 + fx;
    const int64_t base = (int64_t)( qy >> 2 ) * B.stride + ( qx >> 2 );
    xd_qpel_src s;
    s.a = B.ref + (size_t)xd_qpel_plane_a( phase ) * B.plane_size + base + ( fy == 3 ? B.stride : 0 );
    s.b = ( phase & 5 ) ? B.ref + (size_t)xd_qpel_plane_b( phase ) * B.plane_size + base + ( fx == 3 ? 1 : 0 ) : nullptr;
    return s;
}

__device__ __forceinline__ uint2 xd_me_pred8( const xd_qpel_src &s, int stride, int x, int y )
{
    uint2 p = xd_load8_unaligned( s.a + (int64_t)y * stride + x );
    if( s.b )
    {
        const uint2 q = xd_load8_unaligned( s.b + (int64_t)y * stride + x );
        p.x = xd_avg4( p.x, q.x );
        p.y = xd_avg4( p.y, q.y );
    }
    return p;
}
__device__ __forceinline__ uint32_t xd_me_pred4( const xd_qpel_src &s, int stride, int x, int y )
{
    uint32_t p = xd_load4_unaligned( s.a + (int64_t)y * stride + x );
    if( s.b )
        p = xd_avg4( p, xd_load4_unaligned( s.b + (int64_t)y * stride + x ) );
    return p;
}

// A segment of the SOURCE block.  A caller that builds its own block lists (the P-slice wavefront: macroblocks and their
// partitions) knows that a block sits at a multiple of its width in a plane whose rows are 8-byte aligned, so a segment is ONE
// aligned load; blocks handed in from outside may sit anywhere and take the three-words-and-two-funnel-shifts path.
__device__ __forceinline__ uint2 xd_me_src8( const uint8_t *p, bool aligned )
{
    return aligned ? __ldg( (const uint2 *)p ) : xd_load8_unaligned( p );
}
__device__ __forceinline__ uint32_t xd_me_src4( const uint8_t *p, bool aligned )
{
    return aligned ? __ldg( (const uint32_t *)p ) : xd_load4_unaligned( p );
}

// SAD of the block at quarter-pel (qx,qy); lanes of one candidate group cooperate (sub = lane & 7)
// The two cost routines take the block description as scalars so that they can be compiled as real calls
// (-DXD_ME_CALL=__noinline__): the search has some thirty call sites, a kernel that inlines them all is 25 000
// instructions, and ncu shows the warps of the P-slice wavefront waiting for instruction fetches more than for anything
// else (stall no_instruction 10.4 per issue).  Measured (round 2, 96 1080p P frames per launch): as calls the kernel is
// 7 600 instructions and SLOWER -- 158 against 139 us per frame (DIA, subme 1), 292 against 284 (HEX, subme 5); with the
// refinement a call as well 179 / 308 -- so everything stays inlined.
static __device__ XD_ME_CALL int xd_me_sad_call( const uint8_t *fenc, const uint8_t *ref, size_t plane_size, int stride,
                                                   int wh, int qx, int qy, int sub, bool fixed16 = false, bool al = false )
{
    xd_me_blk B;
    B.fenc = fenc; B.ref = ref; B.plane_size = plane_size; B.stride = stride; B.w = wh & 255; B.h = wh >> 8;
    const xd_qpel_src s = xd_me_src( B, qx, qy );
    int acc = 0;
    if( fixed16 )
    {
        // 32 segments of 8 pixels, four per lane.  XD_ME_SAD16_UNROLL = 4 puts all loads of a lane in flight together and
        // wins 2 % at DIA / subme 1, but the kernel grows from 14 k to 17 k instructions and HEX / subme 5 -- which walks far
        // more of the code per macroblock -- loses 11 % (177 -> 197 us per 1080p frame): instruction fetch, again.
#pragma unroll xd_me_sad16_unroll
        for( int k = 0; k < 4; k++ )
        {
            const int i = sub + 8 * k, y = i >> 1, x = ( i & 1 ) * 8;
            const uint2 p = xd_me_pred8( s, B.stride, x, y );
            const uint2 f = xd_me_src8( B.fenc + (int64_t)y * B.stride + x, al );
            acc += __vsadu4( p.x, f.x ) + __vsadu4( p.y, f.y );
        }
    }
    else if( B.w >= 8 )
    {
        // one or two 8-pixel segments per row: the split of i into (row, segment) is a shift and a mask, not a division.
        // Rolled on purpose: the partition kernel is instruction-fetch bound (28 k instructions), every unrolled copy costs
        const int wide = B.w >> 4, nseg = B.h << wide;
#pragma unroll 1
        for( int i = sub; i < nseg; i += 8 )
        {
            const int y = i >> wide, x = ( i & wide ) * 8;
            const uint2 p = xd_me_pred8( s, B.stride, x, y );
            const uint2 f = xd_me_src8( B.fenc + (int64_t)y * B.stride + x, al );
            acc += __vsadu4( p.x, f.x ) + __vsadu4( p.y, f.y );
        }
    }
    else
    {
#pragma unroll 1
        for( int y = sub; y < B.h; y += 8 )
            acc += __vsadu4( xd_me_pred4( s, B.stride, 0, y ), xd_me_src4( B.fenc + (int64_t)y * B.stride, al ) );
    }
    acc += __shfl_xor_sync( 0xffffffffu, acc, 1 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 2 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 4 );
    return acc;
}

// SATD (common/pixel.c:267-337) of the block at quarter-pel (qx,qy)
static __device__ XD_ME_CALL int xd_me_satd_call( const uint8_t *fenc, const uint8_t *ref, size_t plane_size, int stride,
                                                    int wh, int qx, int qy, int sub, bool fixed16 = false, bool al = false )
{
    xd_me_blk B;
    B.fenc = fenc; B.ref = ref; B.plane_size = plane_size; B.stride = stride; B.w = wh & 255; B.h = wh >> 8;
    const xd_qpel_src s = xd_me_src( B, qx, qy );
    int acc = 0;
    if( fixed16 )
    {
        // eight 8x4 units, exactly one per lane of the group: no loop at all
        const int y = ( sub >> 1 ) * 4, x = ( sub & 1 ) * 8;
        uint32_t pa[4], pb[4], fa[4], fb[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint2 p = xd_me_pred8( s, B.stride, x, y + r );
            const uint2 f = xd_me_src8( B.fenc + (int64_t)( y + r ) * B.stride + x, al );
            pa[r] = p.x; pb[r] = p.y; fa[r] = f.x; fb[r] = f.y;
        }
        acc = xd_satd4x4( fa, pa ) + xd_satd4x4( fb, pb );
        acc += __shfl_xor_sync( 0xffffffffu, acc, 1 );
        acc += __shfl_xor_sync( 0xffffffffu, acc, 2 );
        acc += __shfl_xor_sync( 0xffffffffu, acc, 4 );
        return acc;
    }
    const int wide = B.w >> 4;                          // 16-wide blocks have two 8x4 units per row of units
    const int units = ( B.h >> 2 ) << wide;
#pragma unroll 1
    for( int u = sub; u < units; u += 8 )
    {
        const int y = ( u >> wide ) * 4, x = ( u & wide ) * 8;
        uint32_t pa[4], pb[4], fa[4], fb[4];
        if( B.w >= 8 )
        {
#pragma unroll
            for( int r = 0; r < 4; r++ )
            {
                const uint2 p = xd_me_pred8( s, B.stride, x, y + r );
                const uint2 f = xd_me_src8( B.fenc + (int64_t)( y + r ) * B.stride + x, al );
                pa[r] = p.x; pb[r] = p.y; fa[r] = f.x; fb[r] = f.y;
            }
            acc += xd_satd4x4( fa, pa ) + xd_satd4x4( fb, pb );
        }
        else
        {
#pragma unroll
            for( int r = 0; r < 4; r++ )
            {
                pa[r] = xd_me_pred4( s, B.stride, 0, y + r );
                fa[r] = xd_me_src4( B.fenc + (int64_t)( y + r ) * B.stride, al );
            }
            acc += xd_satd4x4( fa, pa );
        }
    }
    acc += __shfl_xor_sync( 0xffffffffu, acc, 1 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 2 );
    acc += __shfl_xor_sync( 0xffffffffu, acc, 4 );
    return acc;
}

__device__ __forceinline__ int xd_me_sad_raw( const xd_me_blk &B, int qx, int qy, int sub )
{
    return xd_me_sad_call( B.fenc, B.ref, B.plane_size, B.stride, B.w | ( B.h << 8 ), qx, qy, sub, B.fixed16, B.src_aligned );
}
__device__ __forceinline__ int xd_me_satd( const xd_me_blk &B, int qx, int qy, int sub )
{
    return xd_me_satd_call( B.fenc, B.ref, B.plane_size, B.stride, B.w | ( B.h << 8 ), qx, qy, sub, B.fixed16, B.src_aligned );
}

// h->pixf.fpelcmp[i_pixel]: SAD, except under TESA where mbcmp_init (encoder.c:429-432) makes it SATD
__device__ __forceinline__ int xd_me_sad( const xd_me_blk &B, int qx, int qy, int sub )
{
    return B.fpel_satd ? xd_me_satd( B, qx, qy, sub ) : xd_me_sad_raw( B, qx, qy, sub );
}

__device__ __forceinline__ int xd_min_groups( int key )
{
    key = min( key, __shfl_xor_sync( 0xffffffffu, key, 8 ) );
    key = min( key, __shfl_xor_sync( 0xffffffffu, key, 16 ) );
    return key;
}

// CHECK_MVRANGE (me.c:155-160)
__device__ __forceinline__ bool xd_me_in_range( const xd_me_blk &B, int mx, int my )
{
    const uint32_t lo = ( (uint32_t)( -B.minx ) << 16 ) | ( (uint32_t)( -B.miny ) & 0x7FFF );
    const uint32_t hi = ( (uint32_t)B.maxx << 16 ) | ( (uint32_t)B.maxy & 0x7FFF ) | 0x8000;
    const uint32_t v = ( (uint32_t)mx << 16 ) | ( (uint32_t)my & 0x7FFF );
    return !( ( ( v + lo ) | ( hi - v ) ) & 0x80004000u );
}

__device__ __forceinline__ uint32_t xd_pack_mv( int x, int y )
{
    return ( (uint32_t)x & 0xFFFF ) | ( (uint32_t)y << 16 );
}

// { refine_hpel, refine_qpel, me_hpel, me_qpel } (me.c:18-32), subme 0..5
static __constant__ uint8_t xd_subpel_iters[6][4] = { {0,0,0,0}, {1,1,0,0}, {0,1,1,0}, {0,2,1,0}, {0,2,1,1}, {0,2,1,2} };

struct xd_me_state
{
    int mvx, mvy, cost, cost_mv;
};

// refine_subpel (me.c:466-587).  thresh = *p_halfpel_thresh (nullptr: the caller passed NULL); when the early
// termination of me.c:527-536 fires, mv and cost are stored and cost_mv keeps its previous value, as there.
// (a function of its own where the compiler keeps it one: FIXED16 repeats the caller's compile-time knowledge inside it)
template<bool FIXED16 = false, bool ALIGNED_SRC = false>
static __device__ XD_ME_REFINE_CALL void xd_me_refine( const xd_me_blk &Bin, xd_me_state &S, int subme, int hpel_iters, int qpel_iters,
                              bool final_refine, int lane, int *thresh = nullptr )
{
    xd_me_blk B = Bin;
    if( FIXED16 )
    {
        B.w = B.h = 16;
        B.fixed16 = true;
    }
    else
        B.fixed16 = false;
    B.src_aligned = ALIGNED_SRC;
    const int cand = lane >> 3, sub = lane & 7;
    const int dx = cand == 2 ? -1 : cand == 3 ? 1 : 0;
    const int dy = cand == 0 ? -1 : cand == 1 ? 1 : 0;
    int bmx = S.mvx, bmy = S.mvy, bcost = S.cost;

    if( hpel_iters && subme < 3 )                                        // me.c:483-490
    {
        const int px = xd_clip3( B.mvpx, B.sminx + 2, B.smaxx - 2 ), py = xd_clip3( B.mvpy, B.sminy + 2, B.smaxy - 2 );
        if( px != bmx || py != bmy )
        {
            const int c = xd_me_sad( B, px, py, sub ) + xd_me_bits( B, px, py );
            if( c < bcost ) { bcost = c; bmx = px; bmy = py; }
        }
    }
    for( int i = hpel_iters; i > 0; i-- )                                // me.c:492-517
    {
        const int qx = bmx + 2 * dx, qy = bmy + 2 * dy;
        const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
        const int key = xd_min_groups( ( c << 2 ) | cand );
        if( ( key >> 2 ) >= bcost )
            break;
        bcost = key >> 2;
        const int w = key & 3;
        bmx += w == 2 ? -2 : w == 3 ? 2 : 0;
        bmy += w == 0 ? -2 : w == 1 ? 2 : 0;
    }
    if( !final_refine && !B.fpel_satd )                                  // me.c:519-524 (mbcmp_unaligned != fpelcmp)
        bcost = xd_me_satd( B, bmx, bmy, sub ) + xd_me_bits( B, bmx, bmy );

    if( thresh )                                                         // me.c:526-539
    {
        if( ( bcost * 7 ) >> 3 > *thresh )
        {
            S.cost = bcost;
            S.mvx = bmx;
            S.mvy = bmy;
            return;
        }
        else if( bcost < *thresh )
            *thresh = bcost;
    }

    if( subme != 1 )
    {
        int bdir = -1;                                                   // me.c:541-564
        for( int i = qpel_iters; i > 0; i-- )
        {
            if( bmy <= B.sminy || bmy >= B.smaxy || bmx <= B.sminx || bmx >= B.smaxx )
                break;
            const int odir = bdir;
            const int qx = bmx + dx, qy = bmy + dy;
            const bool skip = !final_refine && ( cand ^ 1 ) == odir;
            int key = 0x7FFFFFFF;
            // all groups fetch (uniform control flow); a skipped direction simply does not compete
            const int c = xd_me_satd( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
            if( !skip )
                key = ( c << 2 ) | cand;
            key = xd_min_groups( key );
            if( ( key >> 2 ) >= bcost )
                break;
            bcost = key >> 2;
            bdir = key & 3;
            bmx += bdir == 2 ? -1 : bdir == 3 ? 1 : 0;
            bmy += bdir == 0 ? -1 : bdir == 1 ? 1 : 0;
        }
    }
    else if( bmy > B.sminy && bmy < B.smaxy && bmx > B.sminx && bmx < B.smaxx )   // me.c:565-581
    {
        const int qx = bmx + dx, qy = bmy + dy;
        const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
        const int key = xd_min_groups( ( c << 2 ) | cand );
        if( ( key >> 2 ) < bcost )
        {
            bcost = key >> 2;
            const int w = key & 3;
            bmx += w == 2 ? -1 : w == 3 ? 1 : 0;
            bmy += w == 0 ? -1 : w == 1 ? 1 : 0;
        }
    }
    S.cost = bcost;
    S.mvx = bmx;
    S.mvy = bmy;
    S.cost_mv = xd_me_bits( B, bmx, bmy );
}


// The search of one block.  `in` may live in global or shared memory (plain loads); mode and R as in
// x264dsp_me_search_batch_ex_dev: X264DSP_ME_MODE_SEARCH fills R, the two refine modes start from R.  thresh = pointer to
// this block's *p_halfpel_thresh (nullptr: the reference's NULL).  Every lane returns the same R.
// FIXED16: the caller only ever searches 16x16 blocks (the P-slice wavefront without partitions) -- width and height become
// constants, the cost loops unroll and the four segments' loads of a lane are in flight together instead of one after the other.
template<bool FIXED16 = false, bool ALIGNED_SRC = false>
__device__ __forceinline__ void xd_me_search_warp( const x264dsp_geom_t &g, const uint8_t *__restrict__ fenc_slot,
                                                   const uint8_t *__restrict__ fref_slot, const x264dsp_me_params_t &P,
                                                   const uint16_t *__restrict__ cost_mv, const x264dsp_me_block_t *in,
                                                   xd_me_state &R, int mode, int *thresh, int lane )
{
    const int cand = lane >> 3, sub = lane & 7;
    static const uint8_t bw[8] = { 16, 16, 8, 8, 8, 4, 4, 4 }, bh[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

    xd_me_blk B;
    const int size = FIXED16 ? 0 : in->i_pixel & 7;
    B.w = FIXED16 ? 16 : bw[size];
    B.h = FIXED16 ? 16 : bh[size];
    B.fixed16 = FIXED16;
    B.src_aligned = ALIGNED_SRC;
    B.stride = g.luma_stride;
    B.plane_size = (size_t)g.luma_plane_size;
    const int64_t pos = (int64_t)in->by * g.luma_stride + in->bx;
    B.fenc = fenc_slot + g.luma_origin + pos;
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
    B.fpel_satd = P.me_method == X264DSP_ME_TESA && subme >= 2;

    if( mode != X264DSP_ME_MODE_SEARCH )
    {
        // x264_me_refine_qpel_refdupe (me.c:437-440) / x264_me_refine_qpel alone (me.c:426-435; the caller has
        // already taken i_ref_cost off the cost): m->mv, m->cost, m->cost_mv come in through results[blk]
        xd_me_state S = R;
        if( mode == X264DSP_ME_MODE_REFDUPE )
            xd_me_refine<FIXED16, ALIGNED_SRC>( B, S, subme, 0, min( 2, (int)xd_subpel_iters[subme][3] ), false, lane, thresh );
        else
            xd_me_refine<FIXED16, ALIGNED_SRC>( B, S, subme, xd_subpel_iters[subme][0], xd_subpel_iters[subme][1], true, lane );
        R = S;
        return;
    }

    int bmx = xd_clip3( B.mvpx, B.minx * 4, B.maxx * 4 ), bmy = xd_clip3( B.mvpy, B.miny * 4, B.maxy * 4 );
    const int pmx = ( bmx + 2 ) >> 2, pmy = ( bmy + 2 ) >> 2;
    int bcost = ME_COST_MAX;
    int pred_mx = 0, pred_my = 0, pred_cost = ME_COST_MAX;
    uint32_t pmv;

    if( subme >= 3 )
    {
        // me.c:176-193: sub-pel predictors; order 0 = clipped MVP, 1.. = candidates
        pmv = xd_pack_mv( bmx, bmy );
        int best = 0x7FFFFFFF;
        for( int base = 0; base < n_mvc + 1; base += 4 )
        {
            const int k = base + cand;
            int qx = 0, qy = 0;
            bool ok = false;
            if( k == 0 )
            {
                qx = bmx; qy = bmy; ok = n_mvc > 0;
            }
            else if( k <= n_mvc )
            {
                const int cx = in->mvc[k - 1][0], cy = in->mvc[k - 1][1];
                const uint32_t raw = xd_pack_mv( cx, cy );
                ok = raw != 0 && raw != pmv;
                qx = xd_clip3( cx, B.minx * 4, B.maxx * 4 );
                qy = xd_clip3( cy, B.miny * 4, B.maxy * 4 );
            }
            const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
            if( ok )
                best = min( best, ( c << 5 ) | k );
        }
        best = xd_min_groups( best );
        if( best != 0x7FFFFFFF )
        {
            const int k = best & 31;
            pred_cost = best >> 5;
            if( k == 0 ) { pred_mx = bmx; pred_my = bmy; }
            else
            {
                pred_mx = xd_clip3( in->mvc[k - 1][0], B.minx * 4, B.maxx * 4 );
                pred_my = xd_clip3( in->mvc[k - 1][1], B.miny * 4, B.maxy * 4 );
            }
        }
        bmx = ( pred_mx + 2 ) >> 2;
        bmy = ( pred_my + 2 ) >> 2;
        bcost = xd_me_sad( B, bmx << 2, bmy << 2, sub ) + xd_me_bits( B, bmx << 2, bmy << 2 );
        if( pmv )                                                        // me.c:231-233
        {
            const int c = xd_me_sad( B, 0, 0, sub ) + xd_me_bits( B, 0, 0 );
            if( c < bcost ) { bcost = c; bmx = 0; bmy = 0; }
        }
    }
    else
    {
        // me.c:194-233: order 0 = rounded MVP without mv cost, 1..n = rounded/clipped candidates, n+1 = (0,0)
        pmv = xd_pack_mv( pmx, pmy );
        int best = 0x7FFFFFFF;
        for( int base = 0; base < n_mvc + 2; base += 4 )
        {
            const int k = base + cand;
            int fx = 0, fy = 0;
            bool ok = false;
            if( k == 0 )
            {
                fx = pmx; fy = pmy; ok = true;
            }
            else if( k <= n_mvc )
            {
                fx = xd_clip3( ( in->mvc[k - 1][0] + 2 ) >> 2, B.minx, B.maxx );
                fy = xd_clip3( ( in->mvc[k - 1][1] + 2 ) >> 2, B.miny, B.maxy );
                const uint32_t v = xd_pack_mv( fx, fy );
                ok = v != 0 && v != pmv;
            }
            else if( k == n_mvc + 1 )
                ok = pmv != 0;
            int c = xd_me_sad( B, fx << 2, fy << 2, sub );
            if( k != 0 )
                c += xd_me_bits( B, fx << 2, fy << 2 );
            if( ok )
                best = min( best, ( c << 5 ) | k );
        }
        best = xd_min_groups( best );
        bcost = best >> 5;
        const int k = best & 31;
        if( k == 0 ) { bmx = pmx; bmy = pmy; }
        else if( k == n_mvc + 1 ) { bmx = 0; bmy = 0; }
        else
        {
            bmx = xd_clip3( ( in->mvc[k - 1][0] + 2 ) >> 2, B.minx, B.maxx );
            bmy = xd_clip3( ( in->mvc[k - 1][1] + 2 ) >> 2, B.miny, B.maxy );
        }
    }

    if( P.me_method == X264DSP_ME_DIA )
    {
        // me.c:237-274
        const int dx = cand == 2 ? -1 : cand == 3 ? 1 : 0;
        const int dy = cand == 0 ? -1 : cand == 1 ? 1 : 0;
        int left = P.me_range;
        do
        {
            const int qx = ( bmx + dx ) << 2, qy = ( bmy + dy ) << 2;
            const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
            const int key = xd_min_groups( ( c << 2 ) | cand );
            if( ( key >> 2 ) >= bcost )
                break;
            bcost = key >> 2;
            const int w = key & 3;
            bmx += w == 2 ? -1 : w == 3 ? 1 : 0;
            bmy += w == 0 ? -1 : w == 1 ? 1 : 0;
        } while( --left && xd_me_in_range( B, bmx, bmy ) );
    }
    else if( P.me_method == X264DSP_ME_HEX )
    {
        // me.c:276-388.  hexagon points in the reference's order; index i <-> hex2[i+1]
        const int hx[6] = { -2, -1, 1, 2, 1, -1 }, hy[6] = { 0, 2, 2, 0, -2, -2 };
        int best = 0x7FFFFFFF;
#pragma unroll
        for( int pass = 0; pass < 2; pass++ )
        {
            const int k = pass * 4 + cand;
            const int kk = k < 6 ? k : 0;
            const int qx = ( bmx + hx[kk] ) << 2, qy = ( bmy + hy[kk] ) << 2;
            const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
            if( k < 6 )
                best = min( best, ( c << 3 ) | k );
        }
        best = xd_min_groups( best );
        if( ( best >> 3 ) < bcost )
        {
            bcost = best >> 3;
            int dir = best & 7;
            bmx += hx[dir];
            bmy += hy[dir];
            for( int left = ( P.me_range >> 1 ) - 1; left > 0 && xd_me_in_range( B, bmx, bmy ); left-- )
            {
                // three new points: directions dir-1, dir, dir+1 (mod 6)
                const int k = cand < 3 ? ( dir + cand + 5 ) % 6 : dir;
                const int qx = ( bmx + hx[k] ) << 2, qy = ( bmy + hy[k] ) << 2;
                const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
                int key = cand < 3 ? ( c << 2 ) | cand : 0x7FFFFFFF;
                key = xd_min_groups( key );
                if( ( key >> 2 ) >= bcost )
                    break;
                bcost = key >> 2;
                dir = ( dir + ( key & 3 ) + 5 ) % 6;
                bmx += hx[dir];
                bmy += hy[dir];
            }
        }
        // square refine (me.c:361-386)
        const int sx[8] = { 0, 0, -1, 1, -1, -1, 1, 1 }, sy[8] = { -1, 1, 0, 0, -1, 1, -1, 1 };
        best = 0x7FFFFFFF;
#pragma unroll
        for( int pass = 0; pass < 2; pass++ )
        {
            const int k = pass * 4 + cand;
            const int qx = ( bmx + sx[k] ) << 2, qy = ( bmy + sy[k] ) << 2;
            const int c = xd_me_sad( B, qx, qy, sub ) + xd_me_bits( B, qx, qy );
            best = min( best, ( c << 3 ) | k );
        }
        best = xd_min_groups( best );
        if( ( best >> 3 ) < bcost )
        {
            bcost = best >> 3;
            bmx += sx[best & 7];
            bmy += sy[best & 7];
        }
    }

    // me.c:397-414
    xd_me_state S;
    if( pred_cost < bcost )
    {
        S.mvx = pred_mx; S.mvy = pred_my; S.cost = pred_cost;
    }
    else
    {
        S.mvx = bmx << 2; S.mvy = bmy << 2; S.cost = bcost;
    }
    S.cost_mv = xd_me_bits( B, S.mvx, S.mvy );
    if( bmx == pmx && bmy == pmy && subme < 3 )
        S.cost += S.cost_mv;

    if( subme >= 2 )
        xd_me_refine<FIXED16, ALIGNED_SRC>( B, S, subme, xd_subpel_iters[subme][2], xd_subpel_iters[subme][3], false, lane, thresh );
    if( P.refine_qpel )                                                  // me.c:426-435, i_ref_cost = 0
        xd_me_refine<FIXED16, ALIGNED_SRC>( B, S, subme, xd_subpel_iters[subme][0], xd_subpel_iters[subme][1], true, lane );

    R = S;
}
