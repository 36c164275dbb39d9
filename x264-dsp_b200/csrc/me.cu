// me.cu -- batched motion search on full-resolution planes, all partition sizes, sm_100a.
//
// Reference: x264_me_search_ref (encoder/me.c:129-423), refine_subpel (me.c:466-587),
// x264_me_refine_qpel (me.c:426-435), get_ref / mc_luma plane selection (common/mc.c:192-264).
//
// One WARP per block; blocks are independent (the caller supplies mvp / mvc / limits), so the grid
// is simply n_blocks / warps-per-CTA.  Inside a block the warp evaluates up to four candidate
// positions per pass: lane = 8*candidate + sub; the 8 lanes of a candidate share the block's
// pixels in 8-pixel (4 for 4-wide blocks) row segments -- SAD with VABSDIFF4.U8.ACC -- or in 8x4 /
// 4x4 Hadamard tiles -- SATD, tiles being the reference's own base blocks so the >>1 lands in the
// same place -- and three xor-shuffles finish the sum.  Candidate selection packs (cost, order)
// into one integer and takes the min, which reproduces the reference's strict '<' / first-wins
// ordering for every search stage.  All scalar search state is replicated in the 32 lanes.
#include "me_warp.cuh"

#define ME_WARPS 4

__global__ void __launch_bounds__( ME_WARPS * 32 )
xd_me_search_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fenc_slot, const uint8_t *__restrict__ fref_slot,
                     x264dsp_me_params_t P, const uint16_t *__restrict__ cost_mv, int n,
                     const x264dsp_me_block_t *__restrict__ blocks, x264dsp_me_result_t *__restrict__ results,
                     int mode, int32_t *__restrict__ halfpel_thresh )
{
    const int lane = threadIdx.x & 31;
    const int blk = blockIdx.x * ME_WARPS + ( threadIdx.x >> 5 );
    if( blk >= n )
        return;
    int thresh_val = halfpel_thresh ? halfpel_thresh[blk] : 0;
    xd_me_state R;
    R.mvx = R.mvy = R.cost = R.cost_mv = 0;
    if( mode != X264DSP_ME_MODE_SEARCH )
    {
        R.mvx = results[blk].mv[0];
        R.mvy = results[blk].mv[1];
        R.cost = results[blk].cost;
        R.cost_mv = results[blk].cost_mv;
        __syncwarp();
    }
    xd_me_search_warp( g, fenc_slot, fref_slot, P, cost_mv, blocks + blk, R, mode, halfpel_thresh ? &thresh_val : nullptr, lane );
    if( lane == 0 )
    {
        x264dsp_me_result_t r;
        r.mv[0] = (int16_t)R.mvx;
        r.mv[1] = (int16_t)R.mvy;
        r.cost = R.cost;
        r.cost_mv = R.cost_mv;
        results[blk] = r;
        if( halfpel_thresh )
            halfpel_thresh[blk] = thresh_val;
    }
}

// me_method: DIA and HEX are the reference's two pattern searches; UMH / ESA / TESA pass its parameter check
// (encoder/encoder.c:251-259) but have no case in the switch of x264_me_search_ref (me.c:235-394), so the
// search is "predictors, then sub-pel refinement" -- and TESA with subme >= 2 additionally turns fpelcmp into SATD.
int xd_me_params_ok( const x264dsp_me_params_t *p )
{
    return p->subpel_refine >= 1 && p->subpel_refine <= 5 && p->qp >= 0 && p->qp <= 51
        && p->me_method >= X264DSP_ME_DIA && p->me_method <= X264DSP_ME_TESA && p->me_range >= 1;
}

extern "C" int x264dsp_me_search_batch_ex_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                                const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                                const x264dsp_me_params_t *params, int n,
                                                const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                                int mode, int32_t *halfpel_thresh, void *stream )
{
    if( !ctx || !g || !fenc_slot || !fref_slot || !params || n < 0 )
        return X264DSP_E_ARG;
    if( !xd_me_params_ok( params ) || mode < X264DSP_ME_MODE_SEARCH || mode > X264DSP_ME_MODE_REFINE_QPEL )
        return X264DSP_E_ARG;                        // subme 0 does not exist in the reference
    if( n == 0 )
        return 0;
    if( !blocks || !results )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const int grid = ( n + ME_WARPS - 1 ) / ME_WARPS;
    const int pslot = xd_prof_begin( ctx, XD_PROF_ME, s );
    xd_me_search_kernel<<<grid, ME_WARPS * 32, 0, s>>>( *g, fenc_slot, fref_slot, *params,
                                                        ctx->cost_mv_dev[params->qp] + 4096, n, blocks, results,
                                                        mode, halfpel_thresh );
    xd_prof_end( ctx, XD_PROF_ME, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_me_search_batch_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                             const uint8_t *fenc_slot, const uint8_t *fref_slot,
                                             const x264dsp_me_params_t *params, int n,
                                             const x264dsp_me_block_t *blocks, x264dsp_me_result_t *results,
                                             void *stream )
{
    return x264dsp_me_search_batch_ex_dev( ctx, g, fenc_slot, fref_slot, params, n, blocks, results,
                                           X264DSP_ME_MODE_SEARCH, nullptr, stream );
}
