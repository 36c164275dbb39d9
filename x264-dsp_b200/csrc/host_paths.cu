// host_paths.cu -- the full-resolution paths from HOST memory: motion search of every partition size (SURVEY 8(d)
// config 3) and motion compensation + residual coding + in-loop filter (config 4), pictures in, results out.
//
// These are the end-to-end doors bench.py times next to the device-resident launches: host -> device copies of the
// pictures and of the per-block / per-macroblock side information and device -> host copies of the results are part
// of the call.  Work is cut into groups that run on the context's auxiliary streams, so that one group's copies
// overlap another group's kernels; every group ends with its own device -> host copies and the call returns when all
// groups have drained.  Pinned caller buffers (x264dsp_host_alloc) are copied from / to directly; ordinary memory is
// staged by the driver.
#include <string.h>
#include <cstdlib>
#include "common.cuh"

extern "C" int x264dsp_levels_pack_dev( x264dsp_ctx_t *ctx, int n_frames, int mb_count, const int16_t *levels, const uint8_t *nnz,
                                        int16_t *packed, int64_t packed_stride, int32_t *mb_offset, int32_t *frame_total,
                                        void *stream );

#define XH_CHECK( call ) do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { rc = (int)e_; goto drain; } } while( 0 )
#define XH_RC( call ) do { rc = ( call ); if( rc ) goto drain; } while( 0 )

// ---------------------------------------------------------------------------------------------
// slot -> planar I420: the inverse of xd_load_i420_kernel for the picture area (width x height), so that a
// reconstructed frame leaves the device in the format it came in.  One thread = 16 luma bytes or 8 U + 8 V bytes.
__global__ void __launch_bounds__( 256 )
xd_store_i420_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ slots, uint8_t *__restrict__ i420 )
{
    const int frame = blockIdx.z;
    const size_t pic_bytes = (size_t)g.width * g.height * 3 / 2;
    uint8_t *dy = i420 + frame * pic_bytes;
    uint8_t *du = dy + (size_t)g.width * g.height;
    uint8_t *dv = du + (size_t)( g.width >> 1 ) * ( g.height >> 1 );
    const uint8_t *slot = slots + frame * (size_t)g.slot_bytes;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;                      // [0, height) luma rows, then height/2 chroma rows
    const int x0 = t * 16;
    if( x0 >= g.width )
        return;
    if( row < g.height )
    {
        const uint4 v = *(const uint4 *)( slot + g.luma_origin + (size_t)row * g.luma_stride + x0 );
        uint8_t *dst = dy + (size_t)row * g.width + x0;
        if( x0 + 16 <= g.width && ( (uintptr_t)dst & 15 ) == 0 )
            *(uint4 *)dst = v;
        else
        {
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
            for( int k = 0; k < 16 && x0 + k < g.width; k++ )
                dst[k] = (uint8_t)( w[k >> 2] >> ( 8 * ( k & 3 ) ) );
        }
    }
    else
    {
        const int crow = row - g.height;
        if( crow >= ( g.height >> 1 ) )
            return;
        const int cw = g.width >> 1;
        const uint4 v = *(const uint4 *)( slot + g.slot_chroma_off + g.chroma_origin + (size_t)crow * g.chroma_stride + x0 );
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        uint8_t *pu = du + (size_t)crow * cw + ( x0 >> 1 ), *pv = dv + (size_t)crow * cw + ( x0 >> 1 );
        for( int k = 0; k < 8 && ( x0 >> 1 ) + k < cw; k++ )
        {
            const uint32_t pair = w[k >> 1] >> ( 16 * ( k & 1 ) );
            pu[k] = (uint8_t)pair;
            pv[k] = (uint8_t)( pair >> 8 );
        }
    }
}

extern "C" int x264dsp_frame_store_i420_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *slots,
                                              uint8_t *i420, int n_frames, void *stream )
{
    if( !ctx || !g || !slots || !i420 || n_frames <= 0 )
        return X264DSP_E_ARG;
    dim3 grid( ( ( ( g->width + 15 ) >> 4 ) + 255 ) / 256, g->height + ( g->height >> 1 ), n_frames );
    xd_store_i420_kernel<<<grid, 256, 0, xd_stream( ctx, stream )>>>( *g, slots, i420 );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

// ---------------------------------------------------------------------------------------------
// config 3 from host memory
extern "C" int x264dsp_me_search_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_pairs, const uint8_t *luma,
                                               const x264dsp_me_params_t *params, int n_sizes, const int32_t *i_pixel,
                                               const int32_t *n_blocks, const x264dsp_me_block_t *const *blocks,
                                               x264dsp_me_result_t *const *results )
{
    if( !ctx || !luma || !params || n_pairs <= 0 || n_sizes <= 0 || n_sizes > 8 || !i_pixel || !n_blocks || !blocks || !results )
        return X264DSP_E_ARG;
    x264dsp_geom_t g;
    int rc = x264dsp_geometry( width, height, &g );
    if( rc )
        return rc;
    const int nf = n_pairs + 1;
    const size_t pic = (size_t)width * height;
    size_t blk_bytes = 0, res_bytes = 0, off_blk[8], off_res[8];
    for( int s = 0; s < n_sizes; s++ )
    {
        if( n_blocks[s] < 0 || i_pixel[s] < 0 || i_pixel[s] > 7 || ( n_blocks[s] && ( !blocks[s] || !results[s] ) ) )
            return X264DSP_E_ARG;
        off_blk[s] = blk_bytes;
        off_res[s] = res_bytes;
        blk_bytes += ( (size_t)n_pairs * n_blocks[s] * sizeof( x264dsp_me_block_t ) + 255 ) & ~(size_t)255;
        res_bytes += ( (size_t)n_pairs * n_blocks[s] * sizeof( x264dsp_me_result_t ) + 255 ) & ~(size_t)255;
    }
    XD_CHECK( cudaSetDevice( ctx->device ) );
    if( ctx->stage_dev_cap < pic * nf || ctx->clip_slots_cap < (size_t)nf * g.slot_bytes || ctx->me_blocks_cap < blk_bytes
        || ctx->me_results_cap < res_bytes )
        XD_CHECK( cudaDeviceSynchronize() );             // the buffers below may still be in use by an earlier call
    if( ( rc = xd_reserve_dev( (void **)&ctx->stage_dev, &ctx->stage_dev_cap, pic * nf ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->clip_slots, &ctx->clip_slots_cap, (size_t)nf * g.slot_bytes ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->me_blocks, &ctx->me_blocks_cap, blk_bytes ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->me_results, &ctx->me_results_cap, res_bytes ) ) ) return rc;
    if( !ctx->host_ev )
        XD_CHECK( cudaEventCreateWithFlags( &ctx->host_ev, cudaEventDisableTiming ) );

    // planes of every frame on the context's stream; the searches of the partition sizes fan out over the auxiliary
    // streams behind an event (different sizes read the same planes and write different results)
    int used = 0;
    {
        cudaStream_t s0 = ctx->stream;
        XH_CHECK( cudaMemcpyAsync( ctx->stage_dev, luma, pic * nf, cudaMemcpyHostToDevice, s0 ) );
        XH_RC( x264dsp_frame_load_luma_dev( ctx, &g, ctx->stage_dev, ctx->clip_slots, nf, s0 ) );
        XH_RC( x264dsp_frame_expand_border_dev( ctx, &g, ctx->clip_slots, nf, s0 ) );
        XH_RC( x264dsp_frame_filter_dev( ctx, &g, ctx->clip_slots, nf, s0 ) );
        XH_CHECK( cudaEventRecord( ctx->host_ev, s0 ) );
        for( int s = 0; s < n_sizes; s++ )
        {
            if( !n_blocks[s] )
                continue;
            cudaStream_t st = ctx->aux[s % XD_AUX_STREAMS];
            used = used > s % XD_AUX_STREAMS + 1 ? used : s % XD_AUX_STREAMS + 1;
            const size_t nb = (size_t)n_pairs * n_blocks[s];
            x264dsp_me_block_t *d_blk = (x264dsp_me_block_t *)( ctx->me_blocks + off_blk[s] );
            x264dsp_me_result_t *d_res = (x264dsp_me_result_t *)( ctx->me_results + off_res[s] );
            XH_CHECK( cudaMemcpyAsync( d_blk, blocks[s], nb * sizeof( x264dsp_me_block_t ), cudaMemcpyHostToDevice, st ) );
            XH_CHECK( cudaStreamWaitEvent( st, ctx->host_ev, 0 ) );
            XH_RC( x264dsp_me_search_sized_frames_dev( ctx, &g, ctx->clip_slots + g.slot_bytes, ctx->clip_slots, n_pairs, params,
                                                       i_pixel[s], n_blocks[s], d_blk, d_res, st ) );
            XH_CHECK( cudaMemcpyAsync( results[s], d_res, nb * sizeof( x264dsp_me_result_t ), cudaMemcpyDeviceToHost, st ) );
        }
    }
drain:
    {
        cudaError_t e = cudaStreamSynchronize( ctx->stream );
        if( e != cudaSuccess && !rc )
            rc = (int)e;
        for( int i = 0; i < used; i++ )
        {
            e = cudaStreamSynchronize( ctx->aux[i] );
            if( e != cudaSuccess && !rc )
                rc = (int)e;
        }
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// config 4 from host memory
extern "C" int x264dsp_recon_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                           const int16_t *mv16, int qp, const int8_t *mb_type, const uint8_t *partition,
                                           const uint8_t *bs, int alpha_c0_offset, int beta_offset,
                                           int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 )
{
    if( !ctx || !i420 || !mv16 || !mb_type || !partition || !bs || !levels || !nnz || !cbp || !recon_i420 || n_frames <= 0
        || qp < 0 || qp > 51 )
        return X264DSP_E_ARG;
    x264dsp_geom_t g;
    int rc = x264dsp_geometry( width, height, &g );
    if( rc )
        return rc;
    const size_t pic = (size_t)width * height * 3 / 2, nmb = g.mb_count;
    // groups of frames, each on its own stream; group k needs pictures [f0, f1] (f0 is the reference of its first frame)
    int groups = n_frames / 4;
    if( groups < 1 ) groups = 1;
    if( groups > XD_AUX_STREAMS ) groups = XD_AUX_STREAMS;
    // device memory per group: its pictures, their slots (reference side: luma N/H/V/HV + chroma), one prediction /
    // reconstruction slot per frame, side information, outputs
    const size_t per_mb_in = 2 * sizeof( int16_t ) + 1 + 1 + 64, per_mb_out = X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t )
                           + X264DSP_RES_NNZ_PER_MB + sizeof( int16_t );
    const size_t side_bytes = ( (size_t)n_frames * nmb * ( per_mb_in + per_mb_out ) + 4096 ) & ~(size_t)255;
    const size_t need_slots = (size_t)( 2 * n_frames + groups ) * g.slot_bytes;
    const size_t need_pics = (size_t)( n_frames + groups ) * pic + (size_t)n_frames * pic;
    XD_CHECK( cudaSetDevice( ctx->device ) );
    if( ctx->stage_dev_cap < need_pics || ctx->clip_slots_cap < need_slots || ctx->me_blocks_cap < side_bytes )
        XD_CHECK( cudaDeviceSynchronize() );
    if( ( rc = xd_reserve_dev( (void **)&ctx->stage_dev, &ctx->stage_dev_cap, need_pics ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->clip_slots, &ctx->clip_slots_cap, need_slots ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->me_blocks, &ctx->me_blocks_cap, side_bytes ) ) ) return rc;

    uint8_t *d_side = ctx->me_blocks;
    int16_t *d_mv = (int16_t *)d_side;                            d_side += ( (size_t)n_frames * nmb * 4 + 15 ) & ~(size_t)15;
    int16_t *d_lv = (int16_t *)d_side;                            d_side += (size_t)n_frames * nmb * X264DSP_RES_LEVELS_PER_MB * 2;
    int16_t *d_cbp = (int16_t *)d_side;                           d_side += ( (size_t)n_frames * nmb * 2 + 15 ) & ~(size_t)15;
    uint8_t *d_bs = d_side;                                       d_side += (size_t)n_frames * nmb * 64;
    uint8_t *d_nz = d_side;                                       d_side += ( (size_t)n_frames * nmb * X264DSP_RES_NNZ_PER_MB + 15 ) & ~(size_t)15;
    int8_t *d_type = (int8_t *)d_side;                            d_side += ( (size_t)n_frames * nmb + 15 ) & ~(size_t)15;
    uint8_t *d_part = d_side;

    int used = 0;
    size_t slot_cursor = 0, pic_cursor = 0;
    uint8_t *d_out_pics = ctx->stage_dev + (size_t)( n_frames + groups ) * pic;
    for( int gi = 0; gi < groups && !rc; gi++ )
    {
        const int f0 = (int)( (int64_t)n_frames * gi / groups ), f1 = (int)( (int64_t)n_frames * ( gi + 1 ) / groups );
        const int nf = f1 - f0;
        if( nf <= 0 )
            continue;
        cudaStream_t st = ctx->aux[gi];
        used = gi + 1;
        uint8_t *d_pics = ctx->stage_dev + pic_cursor;           pic_cursor += (size_t)( nf + 1 ) * pic;
        uint8_t *d_src = ctx->clip_slots + slot_cursor;           slot_cursor += (size_t)( nf + 1 ) * g.slot_bytes;
        uint8_t *d_pred = ctx->clip_slots + slot_cursor;          slot_cursor += (size_t)nf * g.slot_bytes;
        const size_t m0 = (size_t)f0 * nmb, mn = (size_t)nf * nmb;
        XH_CHECK( cudaMemcpyAsync( d_pics, i420 + (size_t)f0 * pic, (size_t)( nf + 1 ) * pic, cudaMemcpyHostToDevice, st ) );
        XH_CHECK( cudaMemcpyAsync( d_mv + m0 * 2, mv16 + m0 * 2, mn * 4, cudaMemcpyHostToDevice, st ) );
        XH_CHECK( cudaMemcpyAsync( d_bs + m0 * 64, bs + m0 * 64, mn * 64, cudaMemcpyHostToDevice, st ) );
        XH_CHECK( cudaMemcpyAsync( d_type + m0, mb_type + m0, mn, cudaMemcpyHostToDevice, st ) );
        XH_CHECK( cudaMemcpyAsync( d_part + m0, partition + m0, mn, cudaMemcpyHostToDevice, st ) );
        XH_RC( x264dsp_frame_load_i420_dev( ctx, &g, d_pics, d_src, nf + 1, st ) );
        XH_RC( x264dsp_frame_expand_border_dev( ctx, &g, d_src, nf + 1, st ) );
        XH_RC( x264dsp_frame_filter_dev( ctx, &g, d_src, nf, st ) );                 // reference frames only
        XH_RC( x264dsp_mc_frames_dev( ctx, &g, d_src, nf, d_mv + m0 * 2, d_pred, st ) );
        XH_RC( x264dsp_residual_frames_dev( ctx, &g, d_src + g.slot_bytes, d_pred, nf, qp, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB,
                                            d_nz + m0 * X264DSP_RES_NNZ_PER_MB, d_cbp + m0, st ) );
        XH_RC( x264dsp_deblock_frames_dev( ctx, &g, d_pred, nf, d_type + m0, d_part + m0, d_cbp + m0, d_bs + m0 * 64, qp,
                                           alpha_c0_offset, beta_offset, st ) );
        XH_RC( x264dsp_frame_store_i420_dev( ctx, &g, d_pred, d_out_pics + (size_t)f0 * pic, nf, st ) );
        XH_CHECK( cudaMemcpyAsync( levels + m0 * X264DSP_RES_LEVELS_PER_MB, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB,
                                   mn * X264DSP_RES_LEVELS_PER_MB * 2, cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( nnz + m0 * X264DSP_RES_NNZ_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB, mn * X264DSP_RES_NNZ_PER_MB,
                                   cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( cbp + m0, d_cbp + m0, mn * 2, cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( recon_i420 + (size_t)f0 * pic, d_out_pics + (size_t)f0 * pic, (size_t)nf * pic,
                                   cudaMemcpyDeviceToHost, st ) );
    }
drain:
    for( int i = 0; i < used; i++ )
    {
        const cudaError_t e = cudaStreamSynchronize( ctx->aux[i] );
        if( e != cudaSuccess && !rc )
            rc = (int)e;
    }
    ctx->scratch_busy[XD_SCRATCH_DEBLOCK] = 0;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// the P-slice macroblock loop from host memory (SURVEY 8(f) N2 as a door): n_frames independent P frames, frame f + 1 coded
// against picture f.  Everything an encoder would have on the device is built there -- reference planes (border, half-pel),
// half-resolution planes, the lookahead's vectors of each pair (the search's first candidate) -- then x264dsp_p_frames_dev;
// types, vectors, mvr, mvd, levels, nnz, cbp and the reconstruction (planar I420) come back.  All copies are inside the call;
// groups of frames run on separate streams (their wavefronts queue behind each other, their copies overlap).
// partition == NULL: x264dsp_p_frames_dev, one vector per macroblock; else x264dsp_p_frames_part_dev, four (one per 8x8)
static int xh_p_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                             const x264dsp_pframe_params_t *params, int8_t *mb_type, uint8_t *partition, int16_t *mv, int16_t *mvr,
                             int16_t *mvd, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420,
                             int16_t *packed = nullptr, int64_t packed_cap = 0, int64_t *frame_offset = nullptr,
                             int32_t *mb_offset = nullptr )
{
    // packed != NULL: the levels leave as the compact stream of x264dsp_levels_pack_dev (`levels` may then be NULL);
    // recon_i420 == NULL: the reconstruction stays on the device (an encoder only needs it there, as the next reference)
    const size_t nv = partition ? 4 : 1;
    if( !ctx || !i420 || !params || !mb_type || !mv || !mvr || ( !levels && !packed ) || !nnz || !cbp || n_frames <= 0
        || ( packed && ( !frame_offset || !mb_offset || packed_cap <= 0 ) ) )
        return X264DSP_E_ARG;
    x264dsp_geom_t g;
    int rc = x264dsp_geometry( width, height, &g );
    if( rc )
        return rc;
    if( g.mb_w < 3 || g.mb_h < 3 )
        return X264DSP_E_ARG;
    const size_t pic = (size_t)width * height * 3 / 2, nmb = g.mb_count;
    // Groups: the unit that moves through the three-stage pipeline below (upload | kernels | download).  The wavefront wants many
    // frames per launch (a frame alone is latency bound), the pipeline wants several groups so that copies hide behind kernels.
    // Measured on 384 1080p frames (tools/bench_pframe_host.py, DIA / subme 1): dense levels (download bound, 9.8 MB per frame)
    // 2 / 4 / 8 groups = 95 / 84 / 91 ms; compact levels without the reconstruction (1.1 MB per frame) 2 / 3 / 4 / 6 / 8 groups =
    // 54 / 56 / 60 / 74 / 89 ms.
    int groups = packed ? ( n_frames + 96 ) / 192 : ( n_frames + 48 ) / 96;
    if( const char *e = getenv( "X264DSP_PF_HOST_GROUPS" ) )     // measurement knob (tools/bench_pframe_host.py, tests)
        groups = atoi( e );
    if( groups < 1 ) groups = 1;
    if( groups > XD_AUX_STREAMS ) groups = XD_AUX_STREAMS;
    const size_t per_mb = 2 + ( 2 + 2 * nv ) * 2 * sizeof( int16_t ) + X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t ) + X264DSP_RES_NNZ_PER_MB
                        + sizeof( int16_t ) + 4 + X264DSP_LA_SUMS
                        + ( packed ? X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t ) + sizeof( int32_t ) : 0 );
    const size_t side_bytes = ( (size_t)n_frames * nmb * per_mb + (size_t)n_frames * 64 + 8192 ) & ~(size_t)255;
    const size_t need_slots = (size_t)( 2 * n_frames + groups ) * g.slot_bytes;
    const size_t need_pics = (size_t)( n_frames + groups ) * pic + (size_t)n_frames * pic;
    XD_CHECK( cudaSetDevice( ctx->device ) );
    if( ctx->stage_dev_cap < need_pics || ctx->clip_slots_cap < need_slots || ctx->me_blocks_cap < side_bytes )
        XD_CHECK( cudaDeviceSynchronize() );
    if( ( rc = xd_reserve_dev( (void **)&ctx->stage_dev, &ctx->stage_dev_cap, need_pics ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->clip_slots, &ctx->clip_slots_cap, need_slots ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->me_blocks, &ctx->me_blocks_cap, side_bytes ) ) ) return rc;
    const size_t N = (size_t)n_frames * nmb;
    uint8_t *d_side = ctx->me_blocks;
    int16_t *d_lv = (int16_t *)d_side;      d_side += N * X264DSP_RES_LEVELS_PER_MB * 2;
    int16_t *d_mv = (int16_t *)d_side;      d_side += N * 4 * nv;
    int16_t *d_mvd = (int16_t *)d_side;     d_side += N * 4 * nv;
    int16_t *d_mvr = (int16_t *)d_side;     d_side += N * 4;
    int16_t *d_lmv = (int16_t *)d_side;     d_side += N * 4;
    int32_t *d_lc = (int32_t *)d_side;      d_side += N * 4;
    int32_t *d_ls = (int32_t *)d_side;      d_side += ( (size_t)n_frames * X264DSP_LA_SUMS * 4 + 15 ) & ~(size_t)15;
    int16_t *d_cbp = (int16_t *)d_side;     d_side += ( N * 2 + 15 ) & ~(size_t)15;
    uint8_t *d_nz = d_side;                 d_side += ( N * X264DSP_RES_NNZ_PER_MB + 15 ) & ~(size_t)15;
    int8_t *d_type = (int8_t *)d_side;      d_side += ( N + 15 ) & ~(size_t)15;
    uint8_t *d_part = d_side;               d_side += ( N + 15 ) & ~(size_t)15;
    int32_t *d_ftot = (int32_t *)d_side;    d_side += ( (size_t)n_frames * 4 + 15 ) & ~(size_t)15;
    int32_t *d_mboff = (int32_t *)d_side;   d_side += packed ? N * 4 : 0;
    int16_t *d_packed = (int16_t *)d_side;
    const size_t packed_stride = nmb * X264DSP_RES_LEVELS_PER_MB;     // per frame on the device: the dense size is the worst case
    int32_t *h_ftot = nullptr;
    // Three streams, one pipeline: pictures go up on `sh`, every kernel of every group runs on `sc` in group order, results
    // come down on `sd`; events hand a group from one stage to the next.  (A stream per group let the persistent kernels of
    // different groups -- the lookahead of one, the wavefront of another -- share the SMs, and a call then took anything
    // between 52 and 282 ms for the same 384 frames; in order, the same call takes 50.)
    cudaStream_t sh = ctx->aux[0], sc = ctx->aux[1], sd = ctx->aux[2];
    cudaEvent_t tot_ready[XD_AUX_STREAMS], ev_in[XD_AUX_STREAMS], ev_out[XD_AUX_STREAMS];
    int n_events = 0, n_in = 0, n_out = 0;
    int group_f0[XD_AUX_STREAMS + 1];

    int used = 0;
    size_t slot_cursor = 0, pic_cursor = 0;
    uint8_t *d_out_pics = ctx->stage_dev + (size_t)( n_frames + groups ) * pic;
    int32_t *idx = (int32_t *)malloc( ( (size_t)n_frames + 1 ) * ( 2 * sizeof( int32_t ) + 1 ) );
    if( !idx )
        return X264DSP_E_NOMEM;
    if( packed )
    {
        // the context's pinned result buffer (kept between calls: page-locking memory is not something to do per call)
        if( ( rc = xd_reserve_pinned( (void **)&ctx->clip_out_host, &ctx->clip_out_host_cap, (size_t)n_frames * sizeof( int32_t ) ) ) )
        {
            free( idx );
            return rc;
        }
        h_ftot = (int32_t *)ctx->clip_out_host;
    }
    for( int gi = 0; gi < groups && !rc; gi++ )
    {
        const int f0 = (int)( (int64_t)n_frames * gi / groups ), f1 = (int)( (int64_t)n_frames * ( gi + 1 ) / groups );
        const int nf = f1 - f0;
        if( nf <= 0 )
            continue;
        cudaStream_t st = sc;
        used = 3;
        uint8_t *d_pics = ctx->stage_dev + pic_cursor;           pic_cursor += (size_t)( nf + 1 ) * pic;
        uint8_t *d_src = ctx->clip_slots + slot_cursor;           slot_cursor += (size_t)( nf + 1 ) * g.slot_bytes;
        uint8_t *d_rec = ctx->clip_slots + slot_cursor;           slot_cursor += (size_t)nf * g.slot_bytes;
        const size_t m0 = (size_t)f0 * nmb, mn = (size_t)nf * nmb;
        XH_CHECK( cudaMemcpyAsync( d_pics, i420 + (size_t)f0 * pic, (size_t)( nf + 1 ) * pic, cudaMemcpyHostToDevice, sh ) );
        XH_CHECK( cudaEventCreateWithFlags( &ev_in[n_in], cudaEventDisableTiming ) );
        n_in++;
        XH_CHECK( cudaEventRecord( ev_in[n_in - 1], sh ) );
        XH_CHECK( cudaStreamWaitEvent( sc, ev_in[n_in - 1], 0 ) );
        XH_RC( x264dsp_frame_load_i420_dev( ctx, &g, d_pics, d_src, nf + 1, st ) );
        XH_RC( x264dsp_frame_expand_border_dev( ctx, &g, d_src, nf + 1, st ) );
        XH_RC( x264dsp_frame_filter_dev( ctx, &g, d_src, nf, st ) );                 // reference frames only
        XH_RC( x264dsp_frame_init_lowres_dev( ctx, &g, d_src, nf + 1, st ) );
        {
            // the lookahead of every pair of the group: frame k + 1 against frame k, no intra estimate
            int32_t *b = idx, *p0 = idx + nf;
            uint8_t *wi = (uint8_t *)( p0 + nf );
            for( int k = 0; k < nf; k++ )
            {
                b[k] = k + 1;
                p0[k] = k;
                wi[k] = 0;
            }
            XH_RC( x264dsp_lookahead_frame_cost_dev( ctx, &g, d_src, nf, b, p0, wi, d_lmv + m0 * 2, d_lc + m0,
                                                     d_ls + (size_t)f0 * X264DSP_LA_SUMS, NULL, st ) );
        }
        if( partition )
            XH_RC( x264dsp_p_frames_part_dev( ctx, &g, d_src + g.slot_bytes, d_src, d_rec, nf, params, d_lmv + m0 * 2, NULL, d_type + m0,
                                              d_part + m0, d_mv + m0 * 8, d_mvr + m0 * 2, d_mvd + m0 * 8,
                                              d_lv + m0 * X264DSP_RES_LEVELS_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB, d_cbp + m0, st ) );
        else
            XH_RC( x264dsp_p_frames_dev( ctx, &g, d_src + g.slot_bytes, d_src, d_rec, nf, params, d_lmv + m0 * 2, NULL, d_type + m0,
                                         d_mv + m0 * 2, d_mvr + m0 * 2, d_mvd + m0 * 2, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB,
                                         d_nz + m0 * X264DSP_RES_NNZ_PER_MB, d_cbp + m0, st ) );
        group_f0[gi] = f0;
        group_f0[gi + 1] = f1;
        if( packed )
        {
            // the compact stream and, first of all copies, its per-frame lengths: the host needs them to place the frames
            XH_RC( x264dsp_levels_pack_dev( ctx, nf, (int)nmb, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB,
                                            d_packed + (size_t)f0 * packed_stride, (int64_t)packed_stride, d_mboff + m0, d_ftot + f0, st ) );
        }
        if( recon_i420 )
            XH_RC( x264dsp_frame_store_i420_dev( ctx, &g, d_rec, d_out_pics + (size_t)f0 * pic, nf, st ) );
        // ---- the group's results are complete: everything below is copies on the download stream
        XH_CHECK( cudaEventCreateWithFlags( &ev_out[n_out], cudaEventDisableTiming ) );
        n_out++;
        XH_CHECK( cudaEventRecord( ev_out[n_out - 1], sc ) );
        XH_CHECK( cudaStreamWaitEvent( sd, ev_out[n_out - 1], 0 ) );
        st = sd;
        if( packed )
        {
            XH_CHECK( cudaMemcpyAsync( h_ftot + f0, d_ftot + f0, (size_t)nf * sizeof( int32_t ), cudaMemcpyDeviceToHost, st ) );
            XH_CHECK( cudaEventCreateWithFlags( &tot_ready[n_events], cudaEventDisableTiming ) );
            n_events++;
            XH_CHECK( cudaEventRecord( tot_ready[n_events - 1], st ) );
            XH_CHECK( cudaMemcpyAsync( mb_offset + m0, d_mboff + m0, mn * sizeof( int32_t ), cudaMemcpyDeviceToHost, st ) );
        }
        XH_CHECK( cudaMemcpyAsync( mb_type + m0, d_type + m0, mn, cudaMemcpyDeviceToHost, st ) );
        if( partition )
            XH_CHECK( cudaMemcpyAsync( partition + m0, d_part + m0, mn, cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( mv + m0 * 2 * nv, d_mv + m0 * 2 * nv, mn * 4 * nv, cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( mvr + m0 * 2, d_mvr + m0 * 2, mn * 4, cudaMemcpyDeviceToHost, st ) );
        if( mvd )
            XH_CHECK( cudaMemcpyAsync( mvd + m0 * 2 * nv, d_mvd + m0 * 2 * nv, mn * 4 * nv, cudaMemcpyDeviceToHost, st ) );
        if( levels )
            XH_CHECK( cudaMemcpyAsync( levels + m0 * X264DSP_RES_LEVELS_PER_MB, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB,
                                       mn * X264DSP_RES_LEVELS_PER_MB * 2, cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( nnz + m0 * X264DSP_RES_NNZ_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB, mn * X264DSP_RES_NNZ_PER_MB,
                                   cudaMemcpyDeviceToHost, st ) );
        XH_CHECK( cudaMemcpyAsync( cbp + m0, d_cbp + m0, mn * 2, cudaMemcpyDeviceToHost, st ) );
        if( recon_i420 )
            XH_CHECK( cudaMemcpyAsync( recon_i420 + (size_t)f0 * pic, d_out_pics + (size_t)f0 * pic, (size_t)nf * pic,
                                       cudaMemcpyDeviceToHost, st ) );
    }
    if( packed )
    {
        // group by group, as their lengths arrive (the later groups' kernels run meanwhile): frames back to back on the host
        int64_t at = 0;
        frame_offset[0] = 0;
        for( int gi = 0; gi < n_events; gi++ )
        {
            XH_CHECK( cudaEventSynchronize( tot_ready[gi] ) );
            for( int f = group_f0[gi]; f < group_f0[gi + 1]; f++ )
            {
                const int64_t len = h_ftot[f];
                if( at + len > packed_cap )
                {
                    rc = X264DSP_E_ARG;                          // the caller's buffer is too small for this content
                    goto drain;
                }
                if( len )
                    XH_CHECK( cudaMemcpyAsync( packed + at, d_packed + (size_t)f * packed_stride, (size_t)len * sizeof( int16_t ),
                                               cudaMemcpyDeviceToHost, sd ) );
                at += len;
                frame_offset[f + 1] = at;
            }
        }
    }
drain:
    for( int i = 0; i < used; i++ )
    {
        const cudaError_t e = cudaStreamSynchronize( ctx->aux[i] );
        if( e != cudaSuccess && !rc )
            rc = (int)e;
    }
    for( int i = 0; i < n_events; i++ )
        cudaEventDestroy( tot_ready[i] );
    for( int i = 0; i < n_in; i++ )
        cudaEventDestroy( ev_in[i] );
    for( int i = 0; i < n_out; i++ )
        cudaEventDestroy( ev_out[i] );
    free( idx );
    ctx->scratch_busy[XD_SCRATCH_DEBLOCK] = 0;
    ctx->scratch_busy[XD_SCRATCH_LOOKAHEAD] = 0;
    return rc;
}

extern "C" int x264dsp_p_frames_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                       const x264dsp_pframe_params_t *params, int8_t *mb_type, int16_t *mv, int16_t *mvr,
                                       int16_t *mvd, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *recon_i420 )
{
    return xh_p_frames_host( ctx, width, height, n_frames, i420, params, mb_type, nullptr, mv, mvr, mvd, levels, nnz, cbp, recon_i420 );
}

extern "C" int x264dsp_p_frames_part_host( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                            const x264dsp_pframe_params_t *params, int8_t *mb_type, uint8_t *partition,
                                            int16_t *mv8, int16_t *mvr, int16_t *mvd8, int16_t *levels, uint8_t *nnz, int16_t *cbp,
                                            uint8_t *recon_i420 )
{
    if( !partition )
        return X264DSP_E_ARG;
    return xh_p_frames_host( ctx, width, height, n_frames, i420, params, mb_type, partition, mv8, mvr, mvd8, levels, nnz, cbp, recon_i420 );
}

// ... with the levels as the compact stream the entropy coder reads (x264dsp_levels_pack_dev) and, optionally, without the
// reconstruction: what a transcoder moves over PCIe per coded frame drops from 9.8 MB to the picture's own content
extern "C" int x264dsp_p_frames_host_packed( x264dsp_ctx_t *ctx, int width, int height, int n_frames, const uint8_t *i420,
                                              const x264dsp_pframe_params_t *params, int8_t *mb_type, uint8_t *partition,
                                              int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *packed_levels, int64_t packed_capacity,
                                              int64_t *frame_offset, int32_t *mb_offset, uint8_t *nnz, int16_t *cbp,
                                              uint8_t *recon_i420 )
{
    if( !packed_levels )
        return X264DSP_E_ARG;
    return xh_p_frames_host( ctx, width, height, n_frames, i420, params, mb_type, partition, mv, mvr, mvd, nullptr, nnz, cbp, recon_i420,
                             packed_levels, packed_capacity, frame_offset, mb_offset );
}

// ---------------------------------------------------------------------------------------------
// Closed GOPs from host memory (x264dsp_gops_encode_dev behind a door): pictures in, what the entropy coder needs out.  The
// unit of the pipeline is a GOP POSITION: while position t of all GOPs is coded (every launch at the full batch size), position
// t + 1 is uploaded and position t - 1 downloaded -- three streams, events between the stages.  Device and host arrays are both
// position-major, so every result array of a position is one contiguous copy; only the upload is strided ([gop][t] on the host).
extern "C" int x264dsp_gops_encode_step_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                             int n_gops, int gop_len, int t, const x264dsp_gop_encode_params_t *params,
                                             const int16_t *lowres_mv, int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr,
                                             int16_t *mvd8, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16,
                                             uint8_t *chroma_mode, uint8_t *modes4, int16_t *luma_dc, void *stream );

extern "C" int x264dsp_gops_encode_host( x264dsp_ctx_t *ctx, int width, int height, int n_gops, int gop_len, const uint8_t *i420,
                                          const x264dsp_gop_encode_params_t *params, int8_t *mb_type, uint8_t *partition,
                                          int16_t *mv8, int16_t *mvr, int16_t *mvd8, uint8_t *nnz, int16_t *cbp, uint8_t *mode16,
                                          uint8_t *chroma_mode, uint8_t *modes4, int16_t *luma_dc, int16_t *packed_levels,
                                          int64_t packed_capacity, int64_t *frame_offset, int32_t *frame_size, int32_t *mb_offset )
{
    if( !ctx || !i420 || !params || !mb_type || !partition || !mv8 || !mvr || !nnz || !cbp || !mode16 || !chroma_mode || !modes4
        || !luma_dc || !packed_levels || !frame_offset || !frame_size || !mb_offset || n_gops <= 0 || gop_len <= 0 || packed_capacity <= 0 )
        return X264DSP_E_ARG;
    x264dsp_geom_t g;
    int rc = x264dsp_geometry( width, height, &g );
    if( rc )
        return rc;
    if( g.mb_w < 3 || g.mb_h < 3 )
        return X264DSP_E_ARG;
    const size_t pic = (size_t)width * height * 3 / 2, nmb = g.mb_count;
    const size_t n_frames = (size_t)n_gops * gop_len, N = n_frames * nmb, PP = (size_t)n_gops * nmb;       // PP: macroblocks per position
    const size_t per_mb = 1 + 1 + 16 + 4 + 16 + X264DSP_RES_NNZ_PER_MB + 2 + 4 /* lowres mv */ + 4 /* lookahead cost */
                        + 2 * X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t ) + 4 /* mb_offset */ + 1 + 1 + 16 + 32 /* I-frame side */;
    const size_t side_bytes = ( N * per_mb + n_frames * ( 4 + X264DSP_LA_SUMS * 4 ) + 65536 ) & ~(size_t)255;
    XD_CHECK( cudaSetDevice( ctx->device ) );
    if( ctx->stage_dev_cap < n_frames * pic || ctx->clip_slots_cap < 2 * n_frames * g.slot_bytes || ctx->me_blocks_cap < side_bytes )
        XD_CHECK( cudaDeviceSynchronize() );
    if( ( rc = xd_reserve_dev( (void **)&ctx->stage_dev, &ctx->stage_dev_cap, n_frames * pic ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->clip_slots, &ctx->clip_slots_cap, 2 * n_frames * g.slot_bytes ) ) ) return rc;
    if( ( rc = xd_reserve_dev( (void **)&ctx->me_blocks, &ctx->me_blocks_cap, side_bytes ) ) ) return rc;
    if( ( rc = xd_reserve_pinned( (void **)&ctx->clip_out_host, &ctx->clip_out_host_cap, n_frames * sizeof( int32_t ) ) ) ) return rc;
    int32_t *h_ftot = (int32_t *)ctx->clip_out_host;
    uint8_t *d = ctx->me_blocks;
#define XG_TAKE( type, name, count ) type *name = (type *)d; d += ( (size_t)( count ) * sizeof( type ) + 255 ) & ~(size_t)255
    XG_TAKE( int16_t, d_lv, N * X264DSP_RES_LEVELS_PER_MB );
    XG_TAKE( int16_t, d_packed, N * X264DSP_RES_LEVELS_PER_MB );
    XG_TAKE( int16_t, d_mv8, N * 8 );
    XG_TAKE( int16_t, d_mvd8, N * 8 );
    XG_TAKE( int16_t, d_mvr, N * 2 );
    XG_TAKE( int16_t, d_lmv, N * 2 );
    XG_TAKE( int32_t, d_lc, N );
    XG_TAKE( int32_t, d_ls, n_frames * X264DSP_LA_SUMS );
    XG_TAKE( int32_t, d_mboff, N );
    XG_TAKE( int32_t, d_ftot, n_frames );
    XG_TAKE( int16_t, d_cbp, N );
    XG_TAKE( int16_t, d_dc, PP * 16 );
    XG_TAKE( uint8_t, d_nz, N * X264DSP_RES_NNZ_PER_MB );
    XG_TAKE( int8_t, d_type, N );
    XG_TAKE( uint8_t, d_part, N );
    XG_TAKE( uint8_t, d_m16, PP );
    XG_TAKE( uint8_t, d_cm, PP );
    XG_TAKE( uint8_t, d_m4, PP * 16 );
#undef XG_TAKE
    const size_t packed_stride = nmb * X264DSP_RES_LEVELS_PER_MB;
    uint8_t *d_pics = ctx->stage_dev, *d_src = ctx->clip_slots, *d_rec = ctx->clip_slots + n_frames * g.slot_bytes;
    cudaStream_t sh = ctx->aux[0], sc = ctx->aux[1], sd = ctx->aux[2];
    cudaEvent_t *ev = (cudaEvent_t *)calloc( (size_t)gop_len * 3, sizeof( cudaEvent_t ) );
    int32_t *idx = (int32_t *)malloc( ( (size_t)n_gops + 1 ) * ( 2 * sizeof( int32_t ) + 1 ) );
    int n_ev = 0, used = 0, positions = 0;
    if( !ev || !idx )
    {
        free( ev );
        free( idx );
        return X264DSP_E_NOMEM;
    }
    for( int t = 0; t < gop_len; t++ )
    {
        const size_t f0 = (size_t)t * n_gops, m0 = f0 * nmb;
        cudaEvent_t *e_in = &ev[3 * t], *e_out = &ev[3 * t + 1], *e_tot = &ev[3 * t + 2];
        used = 3;
        for( int k = 0; k < 3; k++ )
        {
            XH_CHECK( cudaEventCreateWithFlags( &ev[3 * t + k], cudaEventDisableTiming ) );
            n_ev = 3 * t + k + 1;
        }
        // ---- upload position t: the host holds [gop][t] pictures, one strided copy brings frame t of every GOP
        XH_CHECK( cudaMemcpy2DAsync( d_pics + f0 * pic, pic, i420 + (size_t)t * pic, (size_t)gop_len * pic, pic, n_gops,
                                     cudaMemcpyHostToDevice, sh ) );
        XH_CHECK( cudaEventRecord( *e_in, sh ) );
        XH_CHECK( cudaStreamWaitEvent( sc, *e_in, 0 ) );
        // ---- kernels: staging and half-resolution planes of the new frames, the lookahead of every (t - 1, t) pair, the chain step
        XH_RC( x264dsp_frame_load_i420_dev( ctx, &g, d_pics + f0 * pic, d_src + f0 * g.slot_bytes, n_gops, sc ) );
        XH_RC( x264dsp_frame_expand_border_dev( ctx, &g, d_src + f0 * g.slot_bytes, n_gops, sc ) );
        XH_RC( x264dsp_frame_init_lowres_dev( ctx, &g, d_src + f0 * g.slot_bytes, n_gops, sc ) );
        if( t > 0 )
        {
            int32_t *b = idx, *p0 = idx + n_gops;
            uint8_t *wi = (uint8_t *)( p0 + n_gops );
            for( int k = 0; k < n_gops; k++ )
            {
                b[k] = (int32_t)f0 + k;
                p0[k] = (int32_t)( f0 - n_gops ) + k;
                wi[k] = 0;
            }
            XH_RC( x264dsp_lookahead_frame_cost_dev( ctx, &g, d_src, n_gops, b, p0, wi, d_lmv + m0 * 2, d_lc + m0, d_ls + f0 * X264DSP_LA_SUMS,
                                                     NULL, sc ) );
        }
        XH_RC( x264dsp_gops_encode_step_dev( ctx, &g, d_src, d_rec, n_gops, gop_len, t, params, d_lmv, d_type, d_part, d_mv8, d_mvr, d_mvd8,
                                             d_lv, d_nz, d_cbp, d_m16, d_cm, d_m4, d_dc, sc ) );
        XH_RC( x264dsp_levels_pack_dev( ctx, n_gops, (int)nmb, d_lv + m0 * X264DSP_RES_LEVELS_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB,
                                        d_packed + f0 * packed_stride, (int64_t)packed_stride, d_mboff + m0, d_ftot + f0, sc ) );
        XH_CHECK( cudaEventRecord( *e_out, sc ) );
        XH_CHECK( cudaStreamWaitEvent( sd, *e_out, 0 ) );
        // ---- download position t
        XH_CHECK( cudaMemcpyAsync( h_ftot + f0, d_ftot + f0, (size_t)n_gops * sizeof( int32_t ), cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaEventRecord( *e_tot, sd ) );
        positions = t + 1;
        XH_CHECK( cudaMemcpyAsync( mb_type + m0, d_type + m0, PP, cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( partition + m0, d_part + m0, PP, cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( mv8 + m0 * 8, d_mv8 + m0 * 8, PP * 16, cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( mvr + m0 * 2, d_mvr + m0 * 2, PP * 4, cudaMemcpyDeviceToHost, sd ) );
        if( mvd8 )
            XH_CHECK( cudaMemcpyAsync( mvd8 + m0 * 8, d_mvd8 + m0 * 8, PP * 16, cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( nnz + m0 * X264DSP_RES_NNZ_PER_MB, d_nz + m0 * X264DSP_RES_NNZ_PER_MB, PP * X264DSP_RES_NNZ_PER_MB,
                                   cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( cbp + m0, d_cbp + m0, PP * 2, cudaMemcpyDeviceToHost, sd ) );
        XH_CHECK( cudaMemcpyAsync( mb_offset + m0, d_mboff + m0, PP * 4, cudaMemcpyDeviceToHost, sd ) );
        if( t == 0 )
        {
            XH_CHECK( cudaMemcpyAsync( mode16, d_m16, PP, cudaMemcpyDeviceToHost, sd ) );
            XH_CHECK( cudaMemcpyAsync( chroma_mode, d_cm, PP, cudaMemcpyDeviceToHost, sd ) );
            XH_CHECK( cudaMemcpyAsync( modes4, d_m4, PP * 16, cudaMemcpyDeviceToHost, sd ) );
            XH_CHECK( cudaMemcpyAsync( luma_dc, d_dc, PP * 32, cudaMemcpyDeviceToHost, sd ) );
        }
    }
    {
        // the compact levels, position by position as their lengths arrive (later positions are being coded meanwhile)
        int64_t at = 0;
        for( int t = 0; t < positions; t++ )
        {
            XH_CHECK( cudaEventSynchronize( ev[3 * t + 2] ) );
            for( int k = 0; k < n_gops; k++ )
            {
                const size_t f = (size_t)t * n_gops + k;
                const int64_t len = h_ftot[f];
                if( at + len > packed_capacity )
                {
                    rc = X264DSP_E_ARG;
                    goto drain;
                }
                if( len )
                    XH_CHECK( cudaMemcpyAsync( packed_levels + at, d_packed + f * packed_stride, (size_t)len * sizeof( int16_t ),
                                               cudaMemcpyDeviceToHost, sd ) );
                frame_offset[f] = at;
                frame_size[f] = (int32_t)len;
                at += len;
            }
        }
    }
drain:
    for( int i = 0; i < used; i++ )
    {
        const cudaError_t e = cudaStreamSynchronize( ctx->aux[i] );
        if( e != cudaSuccess && !rc )
            rc = (int)e;
    }
    for( int i = 0; i < n_ev; i++ )
        cudaEventDestroy( ev[i] );
    free( ev );
    free( idx );
    ctx->scratch_busy[XD_SCRATCH_DEBLOCK] = 0;
    ctx->scratch_busy[XD_SCRATCH_LOOKAHEAD] = 0;
    return rc;
}
