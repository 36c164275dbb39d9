// pframe.cu -- the P-slice macroblock loop as a wavefront: analysis + coding of every macroblock of a P frame on the device
// (SURVEY 8(f) N2), sm_100a.
//
// Reference: x264_macroblock_analyse for P slices (encoder/analyse.c:1059-1232) = x264_mb_analyse_init (327-420: MV limits),
// the neighbour-derived predictions of x264_macroblock_cache_load (x264_mb_predict_mv_pskip, common/mvpred.c:139-155), the
// fast P_SKIP probe (analyse.c:1093-1105), x264_mb_analyse_inter_p16x16 (787-860: x264_mb_predict_mv_16x16,
// x264_mb_predict_mv_ref16x16, x264_me_search_ref, the early P_SKIP exit), x264_me_refine_qpel (1187-1191); then
// x264_macroblock_encode (encoder/macroblock.c:310-485): x264_mb_mc, the residual coder, the forced-P_SKIP rule.
// One reference frame, analyse.inter == 0 (P16x16 only, the reference's default); the reference's P-slice analysis has no
// intra candidates (analyse.c:1206-1210 is compiled out), so every macroblock ends up P_L0 16x16 or P_SKIP.
//
// Mapping.  What one macroblock needs from others is small -- type, final vector and 16x16 search vector of its left, top,
// top-left and top-right neighbours -- but it needs them FINAL, including the outcome of the residual coder (a macroblock
// whose residual quantises to nothing at the P_SKIP vector becomes P_SKIP, and a skipped neighbour switches the fast probe
// on).  So the whole per-macroblock chain runs inside the wavefront: ONE WARP WALKS ONE MACROBLOCK ROW left to right and
// does, per macroblock, prediction -> [probe] -> search -> refine -> motion compensation -> residual coding with the same
// warp-level device routines the frame kernels use (me_warp.cuh, residual_warp.cuh, mvpred.cuh); the left neighbour stays
// in registers, the three upper neighbours are read after an acquire on the progress counter of the row above (row y may
// start macroblock x once row y-1 has published macroblock x+1).  Rows are handed out through a ticket, frame-interleaved
// and top rows first, so the row a warp waits for always holds a smaller ticket (no co-residency assumption).  One frame
// alone is latency bound (mb_w + 2 mb_h macroblock times); throughput comes from many independent frames (closed GOPs /
// streams) per launch, exactly as for the lookahead and the deblocking wavefronts.
#include "me_warp.cuh"
#include "residual_warp.cuh"
#include "mvpred.cuh"

#ifndef PF_WARPS
#define PF_WARPS 2
#endif
#ifndef PF_MINB
#define PF_MINB 12                  // resident CTAs per SM the register allocation aims at (80 registers, 24 warps per SM)
#endif
#ifndef PF_PART_MINB
#define PF_PART_MINB 8              // ... of the partition kernel (128 registers; 96 / 80 spill and measure the same)
#endif

int xd_me_params_ok( const x264dsp_me_params_t *p );   // me.cu

struct xd_pf_args
{
    x264dsp_geom_t g;
    const uint8_t *fenc, *fref;
    uint8_t *recon;
    int n_frames;
    x264dsp_pframe_params_t P;
    x264dsp_me_params_t MP;
    xd_res_tables T;
    const uint16_t *cost_mv;
    int lambda;
    const int16_t *lowres_mv, *l0_mv16;
    int8_t *mb_type;
    uint8_t *partition;             // sub-16x16 kernel only (mv / mvd are then [mb][4][2])
    int16_t *mv, *mvr, *mvd, *levels;
    uint8_t *nnz;
    int16_t *cbp;
    int32_t *progress, *ticket;
};

__device__ __forceinline__ int xd_pf_ld_acquire( const int32_t *p )
{
    int v;
    asm volatile( "ld.acquire.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ int xd_pf_ld_relaxed( const int32_t *p )
{
    int v;
    asm volatile( "ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_pf_st_release( int32_t *p, int v )
{
    asm volatile( "st.release.gpu.global.s32 [%0], %1;" :: "l"( p ), "r"( v ) : "memory" );
}

// x264_macroblock_probe_pskip: the prediction at the (clipped) P_SKIP vector goes into the reconstruction slot -- where it
// stays if the macroblock is skipped -- then the "would anything be coded" test (residual_warp.cuh, PROBE)
__device__ __forceinline__ int xd_pf_probe( const xd_pf_args &A, const uint8_t *fenc, const uint8_t *fref, uint8_t *recon, int mb,
                                            uint32_t pskip, int lane )
{
    const int16_t mvs[2] = { (int16_t)( pskip & 0xFFFF ), (int16_t)( pskip >> 16 ) };
    xd_mc_mb32<1>( A.g, fref, mvs, recon, mb % A.g.mb_w, mb / A.g.mb_w, lane );
    __syncwarp();
    const int ok = xd_residual_mb<false, true>( A.g, fenc, recon, A.T, nullptr, nullptr, nullptr, nullptr, nullptr, mb, lane );
    __syncwarp();
    return ok;
}

__global__ void __launch_bounds__( PF_WARPS * 32, PF_MINB )
xd_pframe_kernel( xd_pf_args A )
{
    __shared__ x264dsp_me_block_t s_blk[PF_WARPS];
    const x264dsp_geom_t &g = A.g;
    const int lane = threadIdx.x & 31;
    x264dsp_me_block_t *blk = &s_blk[threadIdx.x >> 5];
    const int W = g.mb_w, H = g.mb_h;
    const int total = A.n_frames * H;
    const int fmv_range = A.P.mv_range << 2, border = 6;
    const int subme = A.P.subpel_refine;
    for( ;; )
    {
        int t = 0;
        if( lane == 0 )
            t = atomicAdd( A.ticket, 1 );
        t = __shfl_sync( 0xffffffffu, t, 0 );
        if( t >= total )
            return;
        const int mb_y = t / A.n_frames, f = t - mb_y * A.n_frames;
        const uint8_t *fenc = A.fenc + (size_t)f * g.slot_bytes, *fref = A.fref + (size_t)f * g.slot_bytes;
        uint8_t *recon = A.recon + (size_t)f * g.slot_bytes;
        const size_t mb0 = (size_t)f * g.mb_count;
        int8_t *types = A.mb_type + mb0;
        uint32_t *mvs = (uint32_t *)A.mv + mb0, *mvrs = (uint32_t *)A.mvr + mb0;
        uint32_t *mvds = A.mvd ? (uint32_t *)A.mvd + mb0 : nullptr;
        int16_t *levels = A.levels + mb0 * X264DSP_RES_LEVELS_PER_MB;
        uint8_t *nnz = A.nnz + mb0 * X264DSP_RES_NNZ_PER_MB;
        int16_t *cbp = A.cbp + mb0;
        const uint32_t *lowres = A.lowres_mv ? (const uint32_t *)A.lowres_mv + mb0 : nullptr;
        const uint32_t *l0 = A.l0_mv16 ? (const uint32_t *)A.l0_mv16 + mb0 : nullptr;
        const bool have_lowres = lowres && ( __ldg( lowres ) & 0xFFFFu ) != 0x7fffu;
        int32_t *mine = A.progress + (size_t)f * H + mb_y;
        const int32_t *above = mine - 1;

        // x264_mb_analyse_init (analyse.c:373-397): the vertical limits are set at the start of a row
        const int min_y = ( -( mb_y << 4 ) - 24 ) << 2, max_y = ( ( ( H - mb_y - 1 ) << 4 ) + 24 ) << 2;
        const int smin_y = xd_clip3( min_y, -fmv_range, fmv_range ), smax_y = xd_clip3( max_y, -fmv_range, fmv_range - 1 );

        int left_type = -1;
        uint32_t left_mv = 0, left_mvr = 0;
        int seen = 0;
        for( int mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            // ---- the row above must have published macroblock x+1
            if( mb_y > 0 )
            {
                const int need = min( mb_x + 2, W );
                if( seen < need )
                {
                    if( lane == 0 )
                    {
                        unsigned ns = 40;
                        while( xd_pf_ld_relaxed( above ) < need )
                        {
                            __nanosleep( ns );
                            if( ns < 1000 )
                                ns += 40;
                        }
                    }
                    __syncwarp();
                    seen = xd_pf_ld_acquire( above );
                }
            }
            // ---- neighbours A (left), B (top), C (top-right), D (top-left): reference 0 inside the frame, -2 outside
            x264dsp_mv_neighbours_t nb;
            int ntype[4] = { left_type, -1, -1, -1 };
            uint32_t nmv[4] = { left_mv, 0, 0, 0 }, nmvr[4] = { left_mvr, 0, 0, 0 };
            bool have[4] = { mb_x > 0, mb_y > 0, mb_y > 0 && mb_x < W - 1, mb_x > 0 && mb_y > 0 };
            const int nxy[4] = { xy - 1, xy - W, xy - W + 1, xy - W - 1 };
#pragma unroll
            for( int k = 1; k < 4; k++ )
                if( have[k] )
                {
                    ntype[k] = __ldcg( types + nxy[k] );
                    nmv[k] = __ldcg( mvs + nxy[k] );
                    nmvr[k] = __ldcg( mvrs + nxy[k] );
                }
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                nb.ref[k] = have[k] ? 0 : -2;
                nb.mv[k][0] = have[k] ? (int16_t)( nmv[k] & 0xFFFF ) : 0;
                nb.mv[k][1] = have[k] ? (int16_t)( nmv[k] >> 16 ) : 0;
            }
            const uint32_t pskip = xd_predict_mv_pskip( nb );
            const int pskip_x = (int16_t)( pskip & 0xFFFF ), pskip_y = (int16_t)( pskip >> 16 );

            const int min_x = ( -( mb_x << 4 ) - 24 ) << 2, max_x = ( ( ( W - mb_x - 1 ) << 4 ) + 24 ) << 2;
            const int smin_x = xd_clip3( min_x, -fmv_range, fmv_range - 1 ), smax_x = xd_clip3( max_x, -fmv_range, fmv_range - 1 );

            int type = X264DSP_MB_P_L0, out_cbp = 0;
            uint32_t out_mv = 0, out_mvr = 0;
            bool done = false;
            // The probe has two callers -- the fast P_SKIP detection before the search (analyse.c:1093-1105) and the early
            // termination after it (analyse.c:839-849) -- and ONE site here (the kernel's code has to stay small: see
            // me_warp.cuh): the loop body runs probe?, search, probe? in that order.
            bool try_probe = A.P.fast_pskip && subme < 3
                             && ( ntype[0] == X264DSP_MB_P_SKIP || ntype[1] == X264DSP_MB_P_SKIP || ntype[2] == X264DSP_MB_P_SKIP
                                  || ntype[3] == X264DSP_MB_P_SKIP );
            bool searched = false;
            xd_me_state R;
            R.mvx = R.mvy = R.cost = R.cost_mv = 0;
#pragma unroll 1
            for( ;; )
            {
                if( try_probe && xd_pf_probe( A, fenc, fref, recon, xy, pskip, lane ) )
                {
                    type = X264DSP_MB_P_SKIP;
                    out_mv = pskip;
                    if( !searched )
                        out_mvr = 0;                           // analyse.c:1109-1117: later macroblocks see a zero 16x16 vector
                    done = true;
                }
                if( done || searched )
                    break;
                // ---- x264_mb_analyse_inter_p16x16: MVP, candidate list (mvpred.c:167-219), search
                const uint32_t mvp = xd_predict_mv_16x16( nb, 0 );
                if( lane == 0 )
                {
                    int n = 0;
                    if( have_lowres )
                    {
                        const uint32_t m = __ldg( lowres + xy );
                        blk->mvc[n][0] = (int16_t)( (int16_t)( m & 0xFFFF ) * 2 );
                        blk->mvc[n][1] = (int16_t)( (int16_t)( m >> 16 ) * 2 );
                        n++;
                    }
                    const int order[4] = { 0, 1, 3, 2 };                          // left, top, top-left, top-right
#pragma unroll
                    for( int k = 0; k < 4; k++, n++ )
                    {
                        const uint32_t m = have[order[k]] ? nmvr[order[k]] : 0u;
                        blk->mvc[n][0] = (int16_t)( m & 0xFFFF );
                        blk->mvc[n][1] = (int16_t)( m >> 16 );
                    }
                    if( l0 )
                    {
                        const int tq[3] = { xy, mb_x < W - 1 ? xy + 1 : -1, mb_y < H - 1 ? xy + W : -1 };
#pragma unroll
                        for( int k = 0; k < 3; k++ )
                            if( tq[k] >= 0 )
                            {
                                const uint32_t m = __ldg( l0 + tq[k] );
                                blk->mvc[n][0] = (int16_t)( ( (int16_t)( m & 0xFFFF ) * A.P.mvc_scale + 128 ) >> 8 );
                                blk->mvc[n][1] = (int16_t)( ( (int16_t)( m >> 16 ) * A.P.mvc_scale + 128 ) >> 8 );
                                n++;
                            }
                    }
                    blk->i_pixel = X264DSP_PIXEL_16x16;
                    blk->bx = mb_x << 4;
                    blk->by = mb_y << 4;
                    blk->mvp[0] = (int16_t)( mvp & 0xFFFF );
                    blk->mvp[1] = (int16_t)( mvp >> 16 );
                    blk->i_mvc = n;
                    blk->mv_min_spel[0] = smin_x; blk->mv_max_spel[0] = smax_x;
                    blk->mv_min_spel[1] = smin_y; blk->mv_max_spel[1] = smax_y;
                    blk->mv_min_fpel[0] = ( smin_x >> 2 ) + border; blk->mv_max_fpel[0] = ( smax_x >> 2 ) - border;
                    blk->mv_min_fpel[1] = ( smin_y >> 2 ) + border; blk->mv_max_fpel[1] = ( smax_y >> 2 ) - border;
                }
                __syncwarp();
                xd_me_search_warp<true, true>( g, fenc, fref, A.MP, A.cost_mv, blk, R, X264DSP_ME_MODE_SEARCH, nullptr, lane );
                searched = true;
                out_mvr = xd_pack_mv( R.mvx, R.mvy );                                  // analyse.c:825
                try_probe = A.P.fast_pskip && subme >= 3 && R.cost - R.cost_mv < 300 * A.lambda
                            && abs( R.mvx - pskip_x ) + abs( R.mvy - pskip_y ) <= 1;
                if( !try_probe )
                    break;
            }
            if( !done )
            {
                // ---- x264_me_refine_qpel (analyse.c:1187-1191; one reference: i_ref_cost = 0)
                xd_me_search_warp<true, true>( g, fenc, fref, A.MP, A.cost_mv, blk, R, X264DSP_ME_MODE_REFINE_QPEL, nullptr, lane );
                out_mv = xd_pack_mv( R.mvx, R.mvy );
                __syncwarp();                                   // the block description is free for the next macroblock
                // ---- x264_macroblock_encode, inter branch: x264_mb_mc, residual, forced P_SKIP (macroblock.c:379-485)
                const int16_t v[2] = { (int16_t)( out_mv & 0xFFFF ), (int16_t)( out_mv >> 16 ) };
                xd_mc_mb32<1>( g, fref, v, recon, mb_x, mb_y, lane );
                __syncwarp();
                out_cbp = xd_residual_mb<false, false>( g, fenc, recon, A.T, levels, nnz, cbp, nullptr, nullptr, xy, lane );
                if( !( out_cbp & 0x3f ) && out_mv == pskip )
                    type = X264DSP_MB_P_SKIP;
            }
            // ---- publish
            if( lane == 0 )
            {
                types[xy] = (int8_t)type;
                mvs[xy] = out_mv;
                mvrs[xy] = out_mvr;
                if( mvds )
                {
                    // encoder/cabac.c:284-287: the 16x16 partition's prediction is x264_mb_predict_mv_16x16's
                    const uint32_t mvp = xd_predict_mv_16x16( nb, 0 );
                    mvds[xy] = type == X264DSP_MB_P_SKIP ? 0u
                             : xd_pack_mv( (int16_t)( out_mv & 0xFFFF ) - (int16_t)( mvp & 0xFFFF ), (int16_t)( out_mv >> 16 ) - (int16_t)( mvp >> 16 ) );
                }
                if( done )
                    cbp[xy] = 0;
            }
            __syncwarp();
            if( lane == 0 )
                xd_pf_st_release( mine, mb_x + 1 );
            left_type = type;
            left_mv = out_mv;
            left_mvr = out_mvr;
        }
    }
}


// ================================================================================================================
// The same wavefront with analyse.inter = X264_ANALYSE_PSUB16x16: after the 16x16 search the reference analyses P8x8
// (analyse.c:864-921), then -- when the 8x8 cost says they might pay (1153-1170) -- P16x8 (923-990) and P8x16 (992-1054), each
// with its early exit, picks the cheapest (1133-1172), refines every partition of the winner (1176-1203) and compensates
// per partition (common/macroblock.c:28-48).  Up to nine searches and four refinements per macroblock; the kernel has ONE
// call site for each (a loop over "jobs": 0 = 16x16, 1..4 = the 8x8 blocks, 5 6 = upper / lower 16x8, 7 8 = left / right
// 8x16), for the code-size reason given in me_warp.cuh.
//
// x264_mb_predict_mv reads a partition's neighbours A / B / C / D from h->mb.cache.mv / .ref, which holds the surrounding
// macroblocks' vectors and, inside the macroblock, whatever the analysis has written so far.  The warp keeps the same
// picture in shared memory at 4x4-cell granularity: cells (-1..4, -1..3) around the macroblock, reference 0 where a vector
// is known and -2 elsewhere (outside the frame, to the right, not yet analysed).  Vectors are published per 8x8 block
// (mv8 [mb][4], raster order): with no sub-8x8 partitions that is all later macroblocks can see.
struct xd_pp_warp
{
    x264dsp_me_block_t blk;
    uint32_t cell_mv[5][6];         // [cy + 1][cx + 1]
    int8_t cell_ref[5][8];
    uint32_t job_mv[9], job_mvp[9];
    int job_cost[9], job_cost_mv[9];
    uint32_t final_mv[4];           // the macroblock's vectors per 8x8 block: what x264_mb_mc works from
    uint32_t mvd[4];
};

// job -> partition: top-left cell (x, y), width and height in cells, x264_mb_predict_mv's rule, pixel size
__device__ __forceinline__ void xd_pp_job( int j, int &x, int &y, int &w, int &hc, int &shape, int &pix )
{
    if( j == 0 )      { x = 0; y = 0; w = 4; hc = 4; shape = 0; pix = X264DSP_PIXEL_16x16; }
    else if( j <= 4 ) { x = 2 * ( ( j - 1 ) & 1 ); y = 2 * ( ( j - 1 ) >> 1 ); w = 2; hc = 2; shape = 0; pix = X264DSP_PIXEL_8x8; }
    else if( j <= 6 ) { x = 0; y = 2 * ( j - 5 ); w = 4; hc = 2; shape = 1 + ( j - 5 ); pix = X264DSP_PIXEL_16x8; }
    else              { x = 2 * ( j - 7 ); y = 0; w = 2; hc = 4; shape = 3 + ( j - 7 ); pix = X264DSP_PIXEL_8x16; }
}

__device__ __forceinline__ x264dsp_mv_neighbours_t xd_pp_neighbours( const xd_pp_warp *S, int x, int y, int w )
{
    // A (x-1, y), B (x, y-1), C (x+w, y-1), D (x-1, y-1) -- mvpred.c:24-33
    const int cx[4] = { x - 1, x, x + w, x - 1 }, cy[4] = { y, y - 1, y - 1, y - 1 };
    x264dsp_mv_neighbours_t nb;
#pragma unroll
    for( int k = 0; k < 4; k++ )
    {
        const uint32_t m = S->cell_mv[cy[k] + 1][cx[k] + 1];
        nb.ref[k] = S->cell_ref[cy[k] + 1][cx[k] + 1];
        nb.mv[k][0] = (int16_t)( m & 0xFFFF );
        nb.mv[k][1] = (int16_t)( m >> 16 );
    }
    return nb;
}

// x264_macroblock_cache_mv_ptr( x, y, w, hc ) + cache_ref 0
__device__ __forceinline__ void xd_pp_write_cells( xd_pp_warp *S, int x, int y, int w, int hc, uint32_t mv, int lane )
{
    if( lane < w * hc )
    {
        const int cx = x + lane % w, cy = y + lane / w;
        S->cell_mv[cy + 1][cx + 1] = mv;
        S->cell_ref[cy + 1][cx + 1] = 0;
    }
    __syncwarp();
}

__global__ void __launch_bounds__( PF_WARPS * 32, PF_PART_MINB )
xd_pframe_part_kernel( xd_pf_args A )
{
    __shared__ xd_pp_warp s_warp[PF_WARPS];
    const x264dsp_geom_t &g = A.g;
    const int lane = threadIdx.x & 31;
    xd_pp_warp *S = &s_warp[threadIdx.x >> 5];
    x264dsp_me_block_t *blk = &S->blk;
    const int W = g.mb_w, H = g.mb_h;
    const int total = A.n_frames * H;
    const int fmv_range = A.P.mv_range << 2, border = 6;
    const int subme = A.P.subpel_refine;
    const bool psub = A.P.analyse_inter != 0;
    for( ;; )
    {
        int t = 0;
        if( lane == 0 )
            t = atomicAdd( A.ticket, 1 );
        t = __shfl_sync( 0xffffffffu, t, 0 );
        if( t >= total )
            return;
        const int mb_y = t / A.n_frames, f = t - mb_y * A.n_frames;
        const uint8_t *fenc = A.fenc + (size_t)f * g.slot_bytes, *fref = A.fref + (size_t)f * g.slot_bytes;
        uint8_t *recon = A.recon + (size_t)f * g.slot_bytes;
        const size_t mb0 = (size_t)f * g.mb_count;
        int8_t *types = A.mb_type + mb0;
        uint8_t *parts = A.partition + mb0;
        uint32_t *mvs = (uint32_t *)A.mv + mb0 * 4, *mvrs = (uint32_t *)A.mvr + mb0;
        uint32_t *mvds = A.mvd ? (uint32_t *)A.mvd + mb0 * 4 : nullptr;
        int16_t *levels = A.levels + mb0 * X264DSP_RES_LEVELS_PER_MB;
        uint8_t *nnz = A.nnz + mb0 * X264DSP_RES_NNZ_PER_MB;
        int16_t *cbp = A.cbp + mb0;
        const uint32_t *lowres = A.lowres_mv ? (const uint32_t *)A.lowres_mv + mb0 : nullptr;
        const uint32_t *l0 = A.l0_mv16 ? (const uint32_t *)A.l0_mv16 + mb0 : nullptr;
        const bool have_lowres = lowres && ( __ldg( lowres ) & 0xFFFFu ) != 0x7fffu;
        int32_t *mine = A.progress + (size_t)f * H + mb_y;
        const int32_t *above = mine - 1;

        const int min_y = ( -( mb_y << 4 ) - 24 ) << 2, max_y = ( ( ( H - mb_y - 1 ) << 4 ) + 24 ) << 2;
        const int smin_y = xd_clip3( min_y, -fmv_range, fmv_range ), smax_y = xd_clip3( max_y, -fmv_range, fmv_range - 1 );

        int left_type = -1;
        uint32_t left_mv1 = 0, left_mv3 = 0, left_mvr = 0;
        int seen = 0;
        for( int mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            if( mb_y > 0 )
            {
                const int need = min( mb_x + 2, W );
                if( seen < need )
                {
                    if( lane == 0 )
                    {
                        unsigned ns = 40;
                        while( xd_pf_ld_relaxed( above ) < need )
                        {
                            __nanosleep( ns );
                            if( ns < 1000 )
                                ns += 40;
                        }
                    }
                    __syncwarp();
                    seen = xd_pf_ld_acquire( above );
                }
            }
            // ---- neighbouring macroblocks: types, 16x16 search vectors, and the 8x8 blocks that border this macroblock
            const bool have[4] = { mb_x > 0, mb_y > 0, mb_y > 0 && mb_x < W - 1, mb_x > 0 && mb_y > 0 };
            const int nxy[4] = { xy - 1, xy - W, xy - W + 1, xy - W - 1 };
            int ntype[4] = { left_type, -1, -1, -1 };
            uint32_t nmvr[4] = { left_mvr, 0, 0, 0 };
            uint32_t top2 = 0, top3 = 0, tr2 = 0, tl3 = 0;
#pragma unroll
            for( int k = 1; k < 4; k++ )
                if( have[k] )
                {
                    ntype[k] = __ldcg( types + nxy[k] );
                    nmvr[k] = __ldcg( mvrs + nxy[k] );
                }
            if( have[1] )
            {
                const uint2 v = __ldcg( (const uint2 *)( mvs + 4 * (size_t)nxy[1] + 2 ) );
                top2 = v.x;
                top3 = v.y;
            }
            if( have[2] )
                tr2 = __ldcg( mvs + 4 * (size_t)nxy[2] + 2 );
            if( have[3] )
                tl3 = __ldcg( mvs + 4 * (size_t)nxy[3] + 3 );
            __syncwarp();
            if( lane < 30 )
            {
                const int cy = lane / 6 - 1, cx = lane % 6 - 1;
                uint32_t m = 0;
                bool ok = false;
                if( cy < 0 )
                {
                    if( cx < 0 )       { m = tl3; ok = have[3]; }
                    else if( cx < 2 )  { m = top2; ok = have[1]; }
                    else if( cx < 4 )  { m = top3; ok = have[1]; }
                    else               { m = tr2; ok = have[2]; }
                }
                else if( cx < 0 )
                {
                    m = cy < 2 ? left_mv1 : left_mv3;
                    ok = have[0];
                }
                S->cell_mv[cy + 1][cx + 1] = ok ? m : 0u;
                S->cell_ref[cy + 1][cx + 1] = ok ? 0 : -2;
            }
            __syncwarp();
            const x264dsp_mv_neighbours_t nb16 = xd_pp_neighbours( S, 0, 0, 4 );
            const uint32_t pskip = xd_predict_mv_pskip( nb16 );
            const int pskip_x = (int16_t)( pskip & 0xFFFF ), pskip_y = (int16_t)( pskip >> 16 );

            const int min_x = ( -( mb_x << 4 ) - 24 ) << 2, max_x = ( ( ( W - mb_x - 1 ) << 4 ) + 24 ) << 2;
            const int smin_x = xd_clip3( min_x, -fmv_range, fmv_range - 1 ), smax_x = xd_clip3( max_x, -fmv_range, fmv_range - 1 );
            if( lane == 0 )
            {
                blk->mv_min_spel[0] = smin_x; blk->mv_max_spel[0] = smax_x;
                blk->mv_min_spel[1] = smin_y; blk->mv_max_spel[1] = smax_y;
                blk->mv_min_fpel[0] = ( smin_x >> 2 ) + border; blk->mv_max_fpel[0] = ( smax_x >> 2 ) - border;
                blk->mv_min_fpel[1] = ( smin_y >> 2 ) + border; blk->mv_max_fpel[1] = ( smax_y >> 2 ) - border;
            }

            int type = X264DSP_MB_P_L0, part = 16, out_cbp = 0;
            uint32_t out_mvr = 0;
            bool skipped = false, searched = false;
            bool try_probe = A.P.fast_pskip && subme < 3
                             && ( ntype[0] == X264DSP_MB_P_SKIP || ntype[1] == X264DSP_MB_P_SKIP || ntype[2] == X264DSP_MB_P_SKIP
                                  || ntype[3] == X264DSP_MB_P_SKIP );
            int j = 0, i_cost = 0, est168 = 0, est816 = 0;
            xd_me_state R;
            R.mvx = R.mvy = R.cost = R.cost_mv = 0;
#pragma unroll 1
            for( ;; )
            {
                // one probe site for its two callers, as in the 16x16 kernel
                if( try_probe )
                {
                    try_probe = false;
                    if( xd_pf_probe( A, fenc, fref, recon, xy, pskip, lane ) )
                    {
                        skipped = true;
                        if( !searched )
                            out_mvr = 0;
                        break;
                    }
                }
                if( j < 0 )
                    break;
                int px, py, pw, ph, shape, pix;
                xd_pp_job( j, px, py, pw, ph, shape, pix );
                const uint32_t mvp = xd_predict_mv_part( xd_pp_neighbours( S, px, py, pw ), 0, shape, false );
                if( lane == 0 )
                {
                    int n = 0;
                    if( j == 0 )
                    {
                        // x264_mb_predict_mv_ref16x16 (mvpred.c:167-219)
                        if( have_lowres )
                        {
                            const uint32_t m = __ldg( lowres + xy );
                            blk->mvc[n][0] = (int16_t)( (int16_t)( m & 0xFFFF ) * 2 );
                            blk->mvc[n][1] = (int16_t)( (int16_t)( m >> 16 ) * 2 );
                            n++;
                        }
                        const int order[4] = { 0, 1, 3, 2 };
#pragma unroll
                        for( int k = 0; k < 4; k++, n++ )
                        {
                            const uint32_t m = have[order[k]] ? nmvr[order[k]] : 0u;
                            blk->mvc[n][0] = (int16_t)( m & 0xFFFF );
                            blk->mvc[n][1] = (int16_t)( m >> 16 );
                        }
                        if( l0 )
                        {
                            const int tq[3] = { xy, mb_x < W - 1 ? xy + 1 : -1, mb_y < H - 1 ? xy + W : -1 };
#pragma unroll
                            for( int k = 0; k < 3; k++ )
                                if( tq[k] >= 0 )
                                {
                                    const uint32_t m = __ldg( l0 + tq[k] );
                                    blk->mvc[n][0] = (int16_t)( ( (int16_t)( m & 0xFFFF ) * A.P.mvc_scale + 128 ) >> 8 );
                                    blk->mvc[n][1] = (int16_t)( ( (int16_t)( m >> 16 ) * A.P.mvc_scale + 128 ) >> 8 );
                                    n++;
                                }
                        }
                    }
                    else
                    {
                        // analyse.c:880-881 / 947-950 / 1016-1018: the 16x16 vector, then 8x8 vectors
                        int first, count;
                        if( j <= 4 )      { first = 1; count = j - 1; }                   // the 8x8 blocks searched so far
                        else if( j <= 6 ) { first = 1 + 2 * ( j - 5 ); count = 2; }       // the two 8x8 blocks of this half
                        else              { first = 1 + ( j - 7 ); count = -2; }          // ... of this column: i, i + 2
                        uint32_t m = S->job_mv[0];
                        blk->mvc[0][0] = (int16_t)( m & 0xFFFF ); blk->mvc[0][1] = (int16_t)( m >> 16 );
                        n = 1;
                        const int step = count < 0 ? 2 : 1;
                        for( int k = 0; k < abs( count ); k++, n++ )
                        {
                            m = S->job_mv[first + k * step];
                            blk->mvc[n][0] = (int16_t)( m & 0xFFFF ); blk->mvc[n][1] = (int16_t)( m >> 16 );
                        }
                    }
                    blk->i_pixel = pix;
                    blk->bx = ( mb_x << 4 ) + 4 * px;
                    blk->by = ( mb_y << 4 ) + 4 * py;
                    blk->mvp[0] = (int16_t)( mvp & 0xFFFF );
                    blk->mvp[1] = (int16_t)( mvp >> 16 );
                    blk->i_mvc = n;
                }
                __syncwarp();
                xd_me_search_warp<false, true>( g, fenc, fref, A.MP, A.cost_mv, blk, R, X264DSP_ME_MODE_SEARCH, nullptr, lane );
                const uint32_t found = xd_pack_mv( R.mvx, R.mvy );
                if( lane == 0 )
                {
                    S->job_mv[j] = found;
                    S->job_mvp[j] = mvp;
                    S->job_cost[j] = R.cost;
                    S->job_cost_mv[j] = R.cost_mv;
                }
                __syncwarp();
                if( j == 0 )
                {
                    searched = true;
                    out_mvr = found;                                                   // analyse.c:825
                    i_cost = R.cost;
                    try_probe = A.P.fast_pskip && subme >= 3 && R.cost - R.cost_mv < 300 * A.lambda
                                && abs( R.mvx - pskip_x ) + abs( R.mvy - pskip_y ) <= 1;
                    j = psub ? 1 : -1;
                }
                else if( j < 4 )
                {
                    xd_pp_write_cells( S, px, py, 2, 2, found, lane );
                    j++;
                }
                else if( j == 4 )
                {
                    xd_pp_write_cells( S, px, py, 2, 2, found, lane );
                    const int c1 = S->job_cost[1], c2 = S->job_cost[2], c3 = S->job_cost[3], c4 = S->job_cost[4];
                    const int m2 = S->job_cost_mv[2], m3 = S->job_cost_mv[3], m4 = S->job_cost_mv[4];
                    const int cost8 = c1 + c2 + c3 + c4, cost16 = i_cost;
                    if( cost8 < cost16 )                                               // analyse.c:1150-1156
                    {
                        type = X264DSP_MB_P_8x8;
                        part = 13;
                        i_cost = cost8;
                    }
                    // analyse.c:1165-1176: job k searched 8x8 block k - 1
                    est168 = ( c3 - m3 ) + ( c4 - m4 ) + ( ( m3 + m4 + 1 ) >> 1 );
                    est816 = ( c2 - m2 ) + ( c4 - m4 ) + ( ( m2 + m4 + 1 ) >> 1 );
                    j = cost8 < cost16 + m2 + m3 ? 5 : -1;
                }
                else if( j == 5 || j == 7 )
                {
                    // analyse.c:975-979 / 1042-1046: the first half alone already too expensive
                    if( R.cost + ( j == 5 ? est168 : est816 ) > i_cost )
                        j = j == 5 ? 7 : -1;
                    else
                    {
                        xd_pp_write_cells( S, px, py, pw, ph, found, lane );
                        j++;
                    }
                }
                else
                {
                    const int c = S->job_cost[j - 1] + R.cost;
                    if( c < i_cost )
                    {
                        i_cost = c;
                        type = X264DSP_MB_P_L0;
                        part = j == 6 ? 14 : 15;
                    }
                    j = j == 6 ? 7 : -1;
                }
            }
            if( skipped )
            {
                type = X264DSP_MB_P_SKIP;
                part = 16;
                if( lane < 4 )
                {
                    S->final_mv[lane] = pskip;
                    S->mvd[lane] = 0;
                }
                __syncwarp();
            }
            else
            {
                // ---- x264_me_refine_qpel on every partition of the winner (analyse.c:1176-1203)
                const int first = part == 16 ? 0 : part == 13 ? 1 : part == 14 ? 5 : 7;
                const int count = part == 16 ? 1 : part == 13 ? 4 : 2;
#pragma unroll 1
                for( int k = 0; k < count; k++ )
                {
                    int px, py, pw, ph, shape, pix;
                    xd_pp_job( first + k, px, py, pw, ph, shape, pix );
                    if( lane == 0 )
                    {
                        const uint32_t mvp = S->job_mvp[first + k];
                        blk->i_pixel = pix;
                        blk->bx = ( mb_x << 4 ) + 4 * px;
                        blk->by = ( mb_y << 4 ) + 4 * py;
                        blk->mvp[0] = (int16_t)( mvp & 0xFFFF );
                        blk->mvp[1] = (int16_t)( mvp >> 16 );
                        blk->i_mvc = 0;
                    }
                    __syncwarp();
                    const uint32_t m = S->job_mv[first + k];
                    R.mvx = (int16_t)( m & 0xFFFF );
                    R.mvy = (int16_t)( m >> 16 );
                    R.cost = S->job_cost[first + k];
                    R.cost_mv = S->job_cost_mv[first + k];
                    xd_me_search_warp<false, true>( g, fenc, fref, A.MP, A.cost_mv, blk, R, X264DSP_ME_MODE_REFINE_QPEL, nullptr, lane );
                    __syncwarp();
                    // the 8x8 blocks this partition covers
                    if( lane < ( pw >> 1 ) * ( ph >> 1 ) )
                        S->final_mv[( ( py >> 1 ) + lane / ( pw >> 1 ) ) * 2 + ( px >> 1 ) + lane % ( pw >> 1 )] = xd_pack_mv( R.mvx, R.mvy );
                }
                __syncwarp();
                // ---- x264_macroblock_encode, inter branch: x264_mb_mc per partition, residual, forced P_SKIP
                xd_mc_mb32<4>( g, fref, (const int16_t *)S->final_mv, recon, mb_x, mb_y, lane );
                __syncwarp();
                out_cbp = xd_residual_mb<false, false>( g, fenc, recon, A.T, levels, nnz, cbp, nullptr, nullptr, xy, lane );
                if( type == X264DSP_MB_P_L0 && part == 16 && !( out_cbp & 0x3f ) && S->final_mv[0] == pskip )
                    type = X264DSP_MB_P_SKIP;
                // ---- the vector differences the entropy coder writes (encoder/cabac.c:278-300, 352-412): partitions in coding
                //      order, each predicted from the FINAL vectors of the ones before it
                if( mvds )
                {
                    if( lane < 16 )
                        S->cell_ref[lane / 4 + 1][lane % 4 + 1] = -2;
                    __syncwarp();
#pragma unroll 1
                    for( int k = 0; k < count; k++ )
                    {
                        int px, py, pw, ph, shape, pix;
                        xd_pp_job( first + k, px, py, pw, ph, shape, pix );
                        const uint32_t mvp = xd_predict_mv_part( xd_pp_neighbours( S, px, py, pw ), 0, shape, false );
                        const uint32_t m = S->final_mv[( py >> 1 ) * 2 + ( px >> 1 )];
                        const uint32_t d = type == X264DSP_MB_P_SKIP ? 0u
                                         : xd_pack_mv( (int16_t)( m & 0xFFFF ) - (int16_t)( mvp & 0xFFFF ), (int16_t)( m >> 16 ) - (int16_t)( mvp >> 16 ) );
                        __syncwarp();
                        if( lane < ( pw >> 1 ) * ( ph >> 1 ) )
                            S->mvd[( ( py >> 1 ) + lane / ( pw >> 1 ) ) * 2 + ( px >> 1 ) + lane % ( pw >> 1 )] = d;
                        xd_pp_write_cells( S, px, py, pw, ph, m, lane );
                    }
                }
            }
            // ---- publish
            if( lane == 0 )
            {
                types[xy] = (int8_t)type;
                parts[xy] = (uint8_t)part;
                *(uint4 *)( mvs + 4 * (size_t)xy ) = make_uint4( S->final_mv[0], S->final_mv[1], S->final_mv[2], S->final_mv[3] );
                mvrs[xy] = out_mvr;
                if( mvds )
                    *(uint4 *)( mvds + 4 * (size_t)xy ) = make_uint4( S->mvd[0], S->mvd[1], S->mvd[2], S->mvd[3] );
                if( skipped )
                    cbp[xy] = 0;
            }
            left_type = type;
            left_mv1 = S->final_mv[1];
            left_mv3 = S->final_mv[3];
            left_mvr = out_mvr;
            __syncwarp();
            if( lane == 0 )
                xd_pf_st_release( mine, mb_x + 1 );
        }
    }
}

static int xd_p_frames_launch( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                               const uint8_t *fref_slots, uint8_t *recon_slots, int n_frames,
                               const x264dsp_pframe_params_t *params, const int16_t *lowres_mv, const int16_t *l0_mv16,
                               int8_t *mb_type, uint8_t *partition, int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *levels,
                               uint8_t *nnz, int16_t *cbp, void *stream )
{
    const bool by_part = partition != nullptr;
    if( !ctx || !g || !fenc_slots || !fref_slots || !recon_slots || !params || !mb_type || !mv || !mvr || !levels || !nnz
        || !cbp || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    x264dsp_me_params_t mp = { params->me_method, params->subpel_refine, params->me_range, params->qp, 0 };
    if( !xd_me_params_ok( &mp ) || params->mv_range < 1 )
        return X264DSP_E_ARG;
    // the kernels read source segments and write vectors with aligned 8 / 16-byte accesses: slots on 16-byte boundaries (slot_bytes is
    // a multiple of 256), per-8x8 vector arrays on 16-byte boundaries
    if( ( (uintptr_t)fenc_slots | (uintptr_t)fref_slots | (uintptr_t)recon_slots ) & 15 )
        return X264DSP_E_ARG;
    if( by_part && ( ( (uintptr_t)mv | (uintptr_t)mvd ) & 15 ) )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    xd_pf_args A;
    memset( &A, 0, sizeof( A ) );
    A.g = *g;
    A.fenc = fenc_slots;
    A.fref = fref_slots;
    A.recon = recon_slots;
    A.n_frames = n_frames;
    A.P = *params;
    A.MP = mp;
    xd_residual_tables( params->qp, &A.T );
    A.cost_mv = ctx->cost_mv_dev[params->qp] + 4096;
    A.lambda = x264dsp_lambda( params->qp );
    A.lowres_mv = lowres_mv;
    A.l0_mv16 = l0_mv16;
    A.mb_type = mb_type;
    A.partition = partition;
    A.mv = mv;
    A.mvr = mvr;
    A.mvd = mvd;
    A.levels = levels;
    A.nnz = nnz;
    A.cbp = cbp;
    // progress counters (one per macroblock row of every frame) + the ticket: the deblocking wavefront's scratch
    const size_t rows = (size_t)n_frames * g->mb_h, need = ( rows + 1 ) * sizeof( int32_t );
    if( ctx->db_progress_cap < need )
        XD_CHECK( cudaDeviceSynchronize() );
    int rc = xd_reserve_dev( (void **)&ctx->db_progress, &ctx->db_progress_cap, need );
    if( rc )
        return rc;
    if( ( rc = xd_scratch_acquire( ctx, XD_SCRATCH_DEBLOCK, s ) ) )
        return rc;
    XD_CHECK( cudaMemsetAsync( ctx->db_progress, 0, need, s ) );
    A.progress = ctx->db_progress;
    A.ticket = ctx->db_progress + rows;
    // skipped macroblocks code nothing
    const size_t nmb = (size_t)n_frames * g->mb_count;
    XD_CHECK( cudaMemsetAsync( levels, 0, nmb * X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t ), s ) );
    XD_CHECK( cudaMemsetAsync( nnz, 0, nmb * X264DSP_RES_NNZ_PER_MB, s ) );
    int per_sm = 0;
    if( by_part )
        XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, xd_pframe_part_kernel, PF_WARPS * 32, 0 ) );
    else
        XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, xd_pframe_kernel, PF_WARPS * 32, 0 ) );
    if( per_sm < 1 )
        per_sm = 1;
    int64_t ctas = (int64_t)ctx->sm_count * per_sm;
    const int64_t wanted = ( (int64_t)rows + PF_WARPS - 1 ) / PF_WARPS;
    if( ctas > wanted )
        ctas = wanted;
    if( by_part )
        xd_pframe_part_kernel<<<(unsigned)ctas, PF_WARPS * 32, 0, s>>>( A );
    else
        xd_pframe_kernel<<<(unsigned)ctas, PF_WARPS * 32, 0, s>>>( A );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return xd_scratch_release( ctx, XD_SCRATCH_DEBLOCK, s );
}

extern "C" int x264dsp_p_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                                     const uint8_t *fref_slots, uint8_t *recon_slots, int n_frames,
                                     const x264dsp_pframe_params_t *params, const int16_t *lowres_mv, const int16_t *l0_mv16,
                                     int8_t *mb_type, int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *levels, uint8_t *nnz,
                                     int16_t *cbp, void *stream )
{
    if( params && params->analyse_inter )
        return X264DSP_E_ARG;                          // one vector per macroblock cannot hold the outcome: _part_dev
    return xd_p_frames_launch( ctx, g, fenc_slots, fref_slots, recon_slots, n_frames, params, lowres_mv, l0_mv16, mb_type, nullptr,
                               mv, mvr, mvd, levels, nnz, cbp, stream );
}

extern "C" int x264dsp_p_frames_part_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots,
                                          const uint8_t *fref_slots, uint8_t *recon_slots, int n_frames,
                                          const x264dsp_pframe_params_t *params, const int16_t *lowres_mv, const int16_t *l0_mv16,
                                          int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8,
                                          int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream )
{
    if( !partition )
        return X264DSP_E_ARG;
    return xd_p_frames_launch( ctx, g, fenc_slots, fref_slots, recon_slots, n_frames, params, lowres_mv, l0_mv16, mb_type, partition,
                               mv8, mvr, mvd8, levels, nnz, cbp, stream );
}
