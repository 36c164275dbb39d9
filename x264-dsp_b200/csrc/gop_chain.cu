// gop_chain.cu -- closed GOPs coded on the device from the first macroblock to the last reference plane, sm_100a.
//
// The slice kernels (iframe.cu, pframe.cu) code one frame against a finished reference.  An encoder's frames follow each
// other: frame t's reconstruction -- deblocked (x264_macroblock_deblock_strength + x264_frame_deblock_row, common/macroblock.c:
// 677-691, common/deblock.c:341-427), border-expanded (x264_frame_expand_border, common/frame.c:386) and half-pel filtered
// (x264_frame_filter, common/mc.c:506) as encoder/encoder.c:1359-1385 does row by row -- is frame t+1's reference.  Here that
// whole chain stays on the device: per GOP position t ONE launch of each stage over the frames of ALL GOPs at that position
// (closed GOPs are independent: that is the batch dimension that fills the machine), then the next position.
// The host sees none of it; what comes back is what the entropy coder needs.
//
// New here: the boundary strengths from the slice kernels' own outputs (the glue took them from the reference's host code):
// bS of an edge segment = 2 if either 4x4 block next to it has a coded coefficient, else 1 if their vectors differ by a
// full sample or more in x or y, else 0 (deblock_strength_c, common/deblock.c:297-323, one reference frame); an intra
// macroblock has 3 on its inner edges and its outer edges are filtered with the intra filter whatever is stored.
#include "common.cuh"

extern "C" int x264dsp_i_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                     int n_frames, int qp, int8_t *mb_type, uint8_t *mode16, uint8_t *chroma_mode, uint8_t *modes4,
                                     int16_t *levels, int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, void *stream );

// coding index (block_idx: the order of nnz[0..15]) of the luma 4x4 at raster (x, y)
__device__ __forceinline__ int xd_gc_block_index( int x, int y )
{
    return ( x & 1 ) + 2 * ( y & 1 ) + 4 * ( x >> 1 ) + 8 * ( y >> 1 );
}

// one thread = one edge of one macroblock (four segments = one 32-bit word of bs [mb][2][8][4])
__global__ void __launch_bounds__( 256 )
xd_bs_frames_kernel( int mb_w, int mb_count, int n_frames, const int8_t *__restrict__ mb_type, const uint8_t *__restrict__ nnz,
                     const uint32_t *__restrict__ mv8, uint8_t *__restrict__ bs )
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= (size_t)n_frames * mb_count * 8 )
        return;
    const int e = (int)( t & 3 ), dir = (int)( ( t >> 2 ) & 1 );
    const size_t rec = t >> 3;                                   // frame * mb_count + mb
    const int mb = (int)( rec % mb_count ), mb_x = mb % mb_w;
    uint32_t word = 0;
    if( mb_type[rec] <= 3 )                                      // IS_INTRA (common/macroblock.h:41-52)
        word = e ? 0x03030303u : 0u;
    else
    {
        // the neighbour across edge 0: the macroblock to the left (dir 0) or above (dir 1), if there is one
        const bool outer = e == 0;
        const bool have = !outer || ( dir == 0 ? mb_x > 0 : mb >= mb_w );
        if( have )
        {
            const size_t nrec = outer ? ( dir == 0 ? rec - 1 : rec - mb_w ) : rec;
#pragma unroll
            for( int i = 0; i < 4; i++ )
            {
                const int qx = dir == 0 ? e : i, qy = dir == 0 ? i : e;
                const int px = dir == 0 ? ( e + 3 ) & 3 : i, py = dir == 0 ? i : ( e + 3 ) & 3;
                const int nzq = nnz[rec * X264DSP_RES_NNZ_PER_MB + xd_gc_block_index( qx, qy )];
                const int nzp = nnz[nrec * X264DSP_RES_NNZ_PER_MB + xd_gc_block_index( px, py )];
                int s = 2;
                if( !( nzq | nzp ) )
                {
                    const uint32_t a = mv8[rec * 4 + ( qy >> 1 ) * 2 + ( qx >> 1 )], b = mv8[nrec * 4 + ( py >> 1 ) * 2 + ( px >> 1 )];
                    const int dx = abs( (int16_t)( a & 0xFFFF ) - (int16_t)( b & 0xFFFF ) ), dy = abs( (int16_t)( a >> 16 ) - (int16_t)( b >> 16 ) );
                    s = ( dx >= 4 || dy >= 4 ) ? 1 : 0;
                }
                word |= (uint32_t)s << ( 8 * i );
            }
        }
    }
    ( (uint32_t *)bs )[( rec * 2 + dir ) * 8 + e] = word;
}

// one vector per macroblock -> one per 8x8 block
__global__ void __launch_bounds__( 256 )
xd_mv_spread_kernel( size_t n, const uint32_t *__restrict__ mv, uint32_t *__restrict__ mv8 )
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if( t < n )
        ( (uint4 *)mv8 )[t] = make_uint4( mv[t], mv[t], mv[t], mv[t] );
}

extern "C" int x264dsp_boundary_strength_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, int n_frames, const int8_t *mb_type,
                                                      const uint8_t *nnz, const int16_t *mv8, uint8_t *bs, void *stream )
{
    if( !ctx || !g || !mb_type || !nnz || !bs || n_frames <= 0 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const size_t n = (size_t)n_frames * g->mb_count * 8;
    xd_bs_frames_kernel<<<(unsigned)( ( n + 255 ) / 256 ), 256, 0, s>>>( g->mb_w, g->mb_count, n_frames, mb_type, nnz, (const uint32_t *)mv8, bs );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

// one GOP position: the slice kernel over the frames of all GOPs at position t, then the in-loop filter stages
static int xd_gops_step( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots, int n_gops, int t,
                         const x264dsp_gop_encode_params_t *P, const int16_t *lowres_mv, int8_t *mb_type, uint8_t *partition,
                         int16_t *mv8, int16_t *mvr, int16_t *mvd8, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16,
                         uint8_t *chroma_mode, uint8_t *modes4, int16_t *luma_dc, cudaStream_t s )
{
    const size_t nmb = g->mb_count, per_pos = (size_t)n_gops * nmb;
    int rc;
    // scratch: boundary strengths of one position, and the 16x16 kernel's one-vector-per-macroblock outputs
    const size_t need = per_pos * 64 + ( P->analyse_inter ? 0 : per_pos * 8 );
    if( ctx->gc_scratch_cap < need )
        XD_CHECK( cudaDeviceSynchronize() );
    if( ( rc = xd_reserve_dev( (void **)&ctx->gc_scratch, &ctx->gc_scratch_cap, need ) ) )
        return rc;
    uint8_t *d_bs = ctx->gc_scratch;
    int16_t *d_mv1 = (int16_t *)( ctx->gc_scratch + per_pos * 64 ), *d_mvd1 = d_mv1 + per_pos * 2;
    const size_t o = (size_t)t * per_pos;
    const uint8_t *fenc = fenc_slots + (size_t)t * n_gops * g->slot_bytes;
    uint8_t *recon = recon_slots + (size_t)t * n_gops * g->slot_bytes;
    const int qp = t == 0 ? P->qp_i : P->qp_p;
    XD_CHECK( cudaMemsetAsync( partition + o, 16, per_pos, s ) );                  // D_16x16 unless the partition kernel says otherwise
    if( t == 0 )
    {
        if( ( rc = x264dsp_i_frames_dev( ctx, g, fenc, recon, n_gops, qp, mb_type, mode16, chroma_mode, modes4, levels, luma_dc, nnz,
                                         cbp, s ) ) )
            return rc;
        XD_CHECK( cudaMemsetAsync( mv8, 0, per_pos * 16, s ) );
        XD_CHECK( cudaMemsetAsync( mvr, 0, per_pos * 4, s ) );
        if( mvd8 )
            XD_CHECK( cudaMemsetAsync( mvd8, 0, per_pos * 16, s ) );
    }
    else
    {
        x264dsp_pframe_params_t pp;
        pp.me_method = P->me_method;
        pp.subpel_refine = P->subpel_refine;
        pp.me_range = P->me_range;
        pp.qp = qp;
        pp.mv_range = P->mv_range;
        pp.fast_pskip = P->fast_pskip;
        // the previous frame's 16x16 vectors are temporal candidates once that frame is a P frame; consecutive frames,
        // no B frames: (curpoc - refpoc) * inv_ref_poc = 2 * 128 (common/mvpred.c:203-218, encoder/encoder.c:1138-1150)
        pp.mvc_scale = t > 1 ? 256 : 0;
        pp.analyse_inter = P->analyse_inter;
        const uint8_t *fref = recon_slots + (size_t)( t - 1 ) * n_gops * g->slot_bytes;
        const int16_t *lmv = lowres_mv ? lowres_mv + o * 2 : NULL;
        const int16_t *l0 = t > 1 ? mvr + ( o - per_pos ) * 2 : NULL;
        if( P->analyse_inter )
            rc = x264dsp_p_frames_part_dev( ctx, g, fenc, fref, recon, n_gops, &pp, lmv, l0, mb_type + o, partition + o, mv8 + o * 8,
                                            mvr + o * 2, mvd8 ? mvd8 + o * 8 : NULL, levels + o * X264DSP_RES_LEVELS_PER_MB,
                                            nnz + o * X264DSP_RES_NNZ_PER_MB, cbp + o, s );
        else
        {
            rc = x264dsp_p_frames_dev( ctx, g, fenc, fref, recon, n_gops, &pp, lmv, l0, mb_type + o, d_mv1, mvr + o * 2,
                                       mvd8 ? d_mvd1 : NULL, levels + o * X264DSP_RES_LEVELS_PER_MB, nnz + o * X264DSP_RES_NNZ_PER_MB,
                                       cbp + o, s );
            if( !rc )
            {
                const unsigned blocks = (unsigned)( ( per_pos + 255 ) / 256 );
                xd_mv_spread_kernel<<<blocks, 256, 0, s>>>( per_pos, (const uint32_t *)d_mv1, (uint32_t *)( mv8 + o * 8 ) );
                if( mvd8 )
                    xd_mv_spread_kernel<<<blocks, 256, 0, s>>>( per_pos, (const uint32_t *)d_mvd1, (uint32_t *)( mvd8 + o * 8 ) );
                ctx->launches += mvd8 ? 2 : 1;
            }
        }
        if( rc )
            return rc;
    }
    // ---- the in-loop filter of position t: this reconstruction is position t + 1's reference
    if( P->deblock )
    {
        if( ( rc = x264dsp_boundary_strength_frames_dev( ctx, g, n_gops, mb_type + o, nnz + o * X264DSP_RES_NNZ_PER_MB, mv8 + o * 8, d_bs, s ) ) )
            return rc;
        if( ( rc = x264dsp_deblock_frames_dev( ctx, g, recon, n_gops, mb_type + o, partition + o, cbp + o, d_bs, qp,
                                               P->alpha_c0_offset, P->beta_offset, s ) ) )
            return rc;
    }
    if( ( rc = x264dsp_frame_expand_border_dev( ctx, g, recon, n_gops, s ) ) )
        return rc;
    if( ( rc = x264dsp_frame_filter_dev( ctx, g, recon, n_gops, s ) ) )
        return rc;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

static bool xd_gops_args_ok( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots, int n_gops,
                             int gop_len, const x264dsp_gop_encode_params_t *P, int8_t *mb_type, uint8_t *partition, int16_t *mv8,
                             int16_t *mvr, int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode,
                             uint8_t *modes4, int16_t *luma_dc )
{
    return ctx && g && fenc_slots && recon_slots && P && mb_type && partition && mv8 && mvr && levels && nnz && cbp && mode16
           && chroma_mode && modes4 && luma_dc && n_gops > 0 && n_gops <= 65535 && gop_len > 0;
}

extern "C" int x264dsp_gops_encode_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                        int n_gops, int gop_len, const x264dsp_gop_encode_params_t *P, const int16_t *lowres_mv,
                                        int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8,
                                        int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode,
                                        uint8_t *modes4, int16_t *luma_dc, void *stream )
{
    if( !xd_gops_args_ok( ctx, g, fenc_slots, recon_slots, n_gops, gop_len, P, mb_type, partition, mv8, mvr, levels, nnz, cbp, mode16,
                          chroma_mode, modes4, luma_dc ) )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    for( int t = 0; t < gop_len; t++ )
    {
        const int rc = xd_gops_step( ctx, g, fenc_slots, recon_slots, n_gops, t, P, lowres_mv, mb_type, partition, mv8, mvr, mvd8, levels,
                                     nnz, cbp, mode16, chroma_mode, modes4, luma_dc, s );
        if( rc )
            return rc;
    }
    return 0;
}

// One position of the same: for a caller that feeds the positions as they arrive (x264dsp_gops_encode_host uploads position
// t + 1 and downloads position t - 1 while position t is coded).  Same arrays and layout as x264dsp_gops_encode_dev; positions
// must be coded in order, position t reads the reconstruction and the 16x16 vectors of position t - 1.
extern "C" int x264dsp_gops_encode_step_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                             int n_gops, int gop_len, int t, const x264dsp_gop_encode_params_t *P, const int16_t *lowres_mv,
                                             int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8,
                                             int16_t *levels, uint8_t *nnz, int16_t *cbp, uint8_t *mode16, uint8_t *chroma_mode,
                                             uint8_t *modes4, int16_t *luma_dc, void *stream )
{
    if( !xd_gops_args_ok( ctx, g, fenc_slots, recon_slots, n_gops, gop_len, P, mb_type, partition, mv8, mvr, levels, nnz, cbp, mode16,
                          chroma_mode, modes4, luma_dc ) || t < 0 || t >= gop_len )
        return X264DSP_E_ARG;
    return xd_gops_step( ctx, g, fenc_slots, recon_slots, n_gops, t, P, lowres_mv, mb_type, partition, mv8, mvr, mvd8, levels, nnz, cbp,
                         mode16, chroma_mode, modes4, luma_dc, xd_stream( ctx, stream ) );
}
