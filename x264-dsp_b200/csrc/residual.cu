// residual.cu -- frame-wide motion compensation and residual coding of inter macroblocks, sm_100a.
//
// Reference:
//   x264_mb_mc / x264_mb_mc_xywh (16x16 case)          common/macroblock.c:8-28
//   mc_luma, mc_chroma                                  common/mc.c:216-239, 290-323
//   x264_macroblock_encode, inter branch + cbp packing  encoder/macroblock.c:379-471
//   x264_mb_encode_chroma (b_inter = 1)                 encoder/macroblock.c:175-305
//   sub4x4_dct / add4x4_idct / add*_idct_dc             common/dct.c:115-150, 197-284
//   quant_4x4 / quant_2x2_dc / dequant_4x4              common/quant.c:29-81
//   optimize_chroma_2x2_dc, decimate_score15/16         common/quant.c:133-261
// with h->mb.b_dct_decimate = 1 (P slice), no noise reduction, CABAC cbp layout.
//
// Mapping of the residual kernel: one WARP per macroblock.  Lanes 0..15 own the sixteen luma 4x4
// blocks (coding order), lanes 16..19 the U blocks and 20..23 the V blocks; a 4x4 block lives
// entirely in one lane's registers from the pixel loads to the reconstructed store (DCT, quant,
// zig-zag, dequant, decimate score, IDCT).  The decisions that couple blocks (8x8 / MB
// decimation, chroma DC 2x2 transform, variance early-out, cbp) are taken with shuffles and
// ballots.  HBM traffic per MB: 2 x 384 B in, 384 B recon + 784 B levels + 29 B flags out.
#include "common.cuh"
#include "leaf.cuh"
#include <string.h>

#include "residual_warp.cuh"

template<int NMV>
__global__ void __launch_bounds__( 256 )
xd_mc_frame_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fref, const int16_t *__restrict__ mv,
                    uint8_t *__restrict__ pred )
{
    // a warp per macroblock, eight macroblocks of a macroblock row per CTA (blockIdx.y = the row: no division anywhere)
    const int mb_x = blockIdx.x * 8 + ( threadIdx.x >> 5 ), mb_y = blockIdx.y;
    if( mb_x >= g.mb_w )
        return;
    // blockIdx.z = frame of a batch: consecutive slots, mb_count MVs per frame
    fref += blockIdx.z * (size_t)g.slot_bytes;
    pred += blockIdx.z * (size_t)g.slot_bytes;
    mv += ( blockIdx.z * (size_t)g.mb_count + mb_y * g.mb_w + mb_x ) * 2 * NMV;
    int16_t mvs[2 * NMV];
    if( NMV == 4 )
    {
        const uint32_t *mw = (const uint32_t *)mv;              // (x, y) pairs: 4-byte aligned, nothing more is promised
        const uint32_t w[4] = { __ldg( mw ), __ldg( mw + 1 ), __ldg( mw + 2 ), __ldg( mw + 3 ) };
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            mvs[2 * k] = (int16_t)( w[k] & 0xFFFF );
            mvs[2 * k + 1] = (int16_t)( w[k] >> 16 );
        }
    }
    else
    {
        const uint32_t v = __ldg( (const uint32_t *)mv );
        mvs[0] = (int16_t)( v & 0xFFFF );
        mvs[1] = (int16_t)( v >> 16 );
    }
    xd_mc_mb32<NMV>( g, fref, mvs, pred, mb_x, mb_y, threadIdx.x & 31 );
}

template<bool TYPED, bool PROBE = false>
__global__ void __launch_bounds__( 128 )
xd_residual_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fenc, uint8_t *__restrict__ pred,
                    xd_res_tables T, int16_t *__restrict__ levels, uint8_t *__restrict__ nnz_out,
                    int16_t *__restrict__ cbp_out, const uint8_t *__restrict__ mb_kind, int16_t *__restrict__ luma_dc )
{
    const int lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + ( threadIdx.x >> 5 );
    if( mb >= g.mb_count )
        return;
    // blockIdx.y = frame of a batch: consecutive slots, per-frame output arrays back to back
    fenc += blockIdx.y * (size_t)g.slot_bytes;
    pred += blockIdx.y * (size_t)g.slot_bytes;
    if( !PROBE )
    {
        levels += blockIdx.y * (size_t)g.mb_count * X264DSP_RES_LEVELS_PER_MB;
        nnz_out += blockIdx.y * (size_t)g.mb_count * X264DSP_RES_NNZ_PER_MB;
        cbp_out += blockIdx.y * (size_t)g.mb_count;
    }
    else
        nnz_out += blockIdx.y * (size_t)g.mb_count;
    if( TYPED && mb_kind )
        mb_kind += blockIdx.y * (size_t)g.mb_count;
    if( TYPED && luma_dc )
        luma_dc += blockIdx.y * (size_t)g.mb_count * 16;
    xd_residual_mb<TYPED, PROBE>( g, fenc, pred, T, levels, nnz_out, cbp_out, mb_kind, luma_dc, mb, lane );
}

// ---------------------------------------------------------------------------------------------
// Luma of I4x4 macroblocks (encoder/macroblock.c:355-377, encoder/macroblock.h:37-61): sixteen blocks in coding order,
// each predicted from the reconstruction of the blocks before it (and of the neighbouring macroblocks, which must be
// final in pred: the caller's launch order guarantees it), transformed, quantised with the CQM_4IY tables and
// reconstructed before the next one starts.  One warp per macroblock, the macroblock and its neighbourhood (row -1 from
// column -1 to 19, column -1) in a shared-memory tile; lane = pixel for the prediction, the 4x4 transform pipeline runs
// on every lane alike (the chain is serial by nature; an I slice comes once per keyint).  Runs after
// xd_residual_kernel<true>, which has done the chroma and written the chroma bits of cbp.
#define I4_PITCH 32
__global__ void __launch_bounds__( 128 )
xd_intra4_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fenc, uint8_t *__restrict__ pred, xd_res_tables T,
                  const uint8_t *__restrict__ mb_kind, const uint8_t *__restrict__ i4_modes,
                  int16_t *__restrict__ levels, uint8_t *__restrict__ nnz_out, int16_t *__restrict__ cbp_out )
{
    __shared__ __align__( 16 ) uint8_t s_tile[4][17 * I4_PITCH];
    const int lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + ( threadIdx.x >> 5 );
    if( mb >= g.mb_count )
        return;
    const size_t fmb = blockIdx.y * (size_t)g.mb_count + mb;
    const int kind = mb_kind[fmb];
    if( ( kind & 3 ) != 2 )
        return;
    const bool replicate5 = ( kind & 4 ) != 0;
    fenc += blockIdx.y * (size_t)g.slot_bytes;
    pred += blockIdx.y * (size_t)g.slot_bytes;
    const int mb_x = mb % g.mb_w, mb_y = mb / g.mb_w;
    const int ls = g.luma_stride;
    const int64_t org = g.luma_origin + (int64_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
    uint8_t *tile = s_tile[threadIdx.x >> 5] + I4_PITCH + 8;          // macroblock origin: tile row 1, byte 8
    if( lane < 21 )
        tile[-I4_PITCH - 1 + lane] = pred[org - ls - 1 + lane];
    if( lane < 16 )
        tile[lane * I4_PITCH - 1] = pred[org + (int64_t)lane * ls - 1];
    __syncwarp();
    const uint8_t *modes = i4_modes + fmb * 16;
    int16_t *mb_levels = levels + fmb * X264DSP_RES_LEVELS_PER_MB;
    int cbp_luma = 0;
    uint32_t nz_bits = 0;
    for( int idx = 0; idx < 16; idx++ )
    {
        const int x = ( ( idx & 1 ) + ( ( idx >> 2 ) & 1 ) * 2 ) * 4, y = ( ( ( idx >> 1 ) & 1 ) + ( ( idx >> 3 ) & 1 ) * 2 ) * 4;
        uint8_t *dst = tile + y * I4_PITCH + x;
        // missing top-right samples: the block's top-right block is coded later (or lies in a macroblock that is not there)
        if( idx == 3 || idx == 7 || idx == 11 || idx == 13 || idx == 15 || ( idx == 5 && replicate5 ) )
        {
            const uint8_t v = dst[3 - I4_PITCH];
            __syncwarp();
            if( lane < 4 )
                dst[4 - I4_PITCH + lane] = v;
            __syncwarp();
        }
        int e[13];
#pragma unroll
        for( int k = 0; k < 4; k++ )
            e[3 - k] = dst[k * I4_PITCH - 1];
        e[4] = dst[-I4_PITCH - 1];
#pragma unroll
        for( int k = 0; k < 8; k++ )
            e[5 + k] = dst[-I4_PITCH + k];
        const int mode = modes[idx];
        __syncwarp();
        if( lane < 16 )
            dst[( lane >> 2 ) * I4_PITCH + ( lane & 3 )] = (uint8_t)xd_pred4x4_px( mode, lane & 3, lane >> 2, e );
        __syncwarp();
        uint32_t f[4], p[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            f[r] = __ldg( (const uint32_t *)( fenc + org + (int64_t)( y + r ) * ls + x ) );
            p[r] = *(const uint32_t *)( dst + r * I4_PITCH );
        }
        int dct[16], lv[16];
        xd_sub4x4_dct( dct, f, p );
        const int nz = xd_quant_4x4( dct, T.luma_i );
        xd_zigzag( lv, dct );
        if( nz )
        {
            xd_dequant_4x4( dct, T.luma_i );
            xd_add4x4_idct( p, dct );
            cbp_luma |= 1 << ( idx >> 2 );
            nz_bits |= 1u << idx;
        }
        __syncwarp();
        if( lane == 0 )
        {
            xd_store_levels( mb_levels + idx * 16, lv );
#pragma unroll
            for( int r = 0; r < 4; r++ )
                *(uint32_t *)( dst + r * I4_PITCH ) = p[r];
        }
        __syncwarp();
    }
    // the reconstructed macroblock, its flags, the luma bits of cbp
    if( lane < 16 )
    {
        const uint2 r0 = *(const uint2 *)( tile + lane * I4_PITCH ), r1 = *(const uint2 *)( tile + lane * I4_PITCH + 8 );
        *(uint4 *)( pred + org + (int64_t)lane * ls ) = make_uint4( r0.x, r0.y, r1.x, r1.y );
        nnz_out[fmb * X264DSP_RES_NNZ_PER_MB + lane] = (uint8_t)( ( nz_bits >> lane ) & 1u );
    }
    if( lane == 0 )
        cbp_out[fmb] = (int16_t)( ( cbp_out[fmb] & ~0x10F ) | cbp_luma );
}

// ---------------------------------------------------------------------------------------------

static const int xd_lambda2_tab[52] =
{
        14,     18,     22,     28,     36,     45,     57,     72,     91,    115,    145,    182,    230,
       290,    365,    460,    580,    731,    921,   1161,   1462,   1843,   2322,   2925,   3686,   4644,
      5851,   7372,   9289,  11703,  14745,  18578,  23407,  29491,  37156,  46814,  58982,  74313,  93628,
    117964, 148626, 187257, 235929, 297252, 374514, 471859, 594505, 749029, 943718,1189010,1498059,1887436
};

static void xd_fill_qparams( xd_qparams *q, int qp, int b_inter = 1 )
{
    uint16_t mf[16], bias[16];
    int dq[6][16];
    x264dsp_quant_tables( b_inter, qp, mf, bias );
    x264dsp_dequant_table( dq );
    static const int rep[3] = { 0, 1, 5 };          // a position of each class
    for( int c = 0; c < 3; c++ )
    {
        q->mf[c] = mf[rep[c]];
        q->bias[c] = bias[rep[c]];
        q->dmf[c] = dq[qp % 6][rep[c]];
    }
    q->qbits = qp / 6 - 4;
}

// every scalar the residual coder and the P_SKIP probe take from the quantiser tables for one slice QP
void xd_residual_tables( int qp, xd_res_tables *out )
{
    xd_res_tables T;
    memset( &T, 0, sizeof( T ) );
    const int qpc = x264dsp_chroma_qp( qp );
    xd_fill_qparams( &T.luma, qp );
    xd_fill_qparams( &T.chroma, qpc );
    xd_fill_qparams( &T.luma_i, qp, 0 );
    xd_fill_qparams( &T.chroma_i, qpc, 0 );
    {
        uint16_t mf[16], bias[16];
        int dq[6][16];
        x264dsp_quant_tables( 1, qpc, mf, bias );
        x264dsp_dequant_table( dq );
        T.chroma_dc_mf = mf[0] >> 1;
        T.chroma_dc_bias = bias[0] << 1;
        T.chroma_dmf_full = dq[qpc % 6][0] << ( qpc / 6 );
        x264dsp_quant_tables( 0, qpc, mf, bias );
        T.chroma_dc_bias_i = bias[0] << 1;
        x264dsp_quant_tables( 0, qp, mf, bias );
        T.luma_dc_mf = mf[0] >> 1;
        T.luma_dc_bias = bias[0] << 1;
        T.luma_dc_dmf = dq[qp % 6][0];
        T.luma_dc_qbits = qp / 6 - 6;
    }
    T.qpc = qpc;
    T.thresh = ( xd_lambda2_tab[qpc] + 32 ) >> 6;
    *out = T;
}

static int xd_residual_launch( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot,
                               int n_frames, int qp, const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels,
                               int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, bool typed, void *stream )
{
    if( !ctx || !g || !fenc_slot || !pred_slot || !levels || !nnz || !cbp || qp < 0 || qp > 51 || n_frames <= 0
        || n_frames > 65535 )
        return X264DSP_E_ARG;
    xd_res_tables T;
    xd_residual_tables( qp, &T );
    cudaStream_t s = xd_stream( ctx, stream );
    const dim3 grid( ( g->mb_count + 3 ) / 4, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_RESIDUAL, s );
    if( typed )
    {
        xd_residual_kernel<true><<<grid, 128, 0, s>>>( *g, fenc_slot, pred_slot, T, levels, nnz, cbp, mb_kind, luma_dc );
        if( mb_kind && i4_modes )
        {
            xd_intra4_kernel<<<grid, 128, 0, s>>>( *g, fenc_slot, pred_slot, T, mb_kind, i4_modes, levels, nnz, cbp );
            ctx->launches++;
        }
    }
    else
        xd_residual_kernel<false><<<grid, 128, 0, s>>>( *g, fenc_slot, pred_slot, T, levels, nnz, cbp, NULL, NULL );
    xd_prof_end( ctx, XD_PROF_RESIDUAL, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_residual_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                             const uint8_t *fenc_slot, uint8_t *pred_slot, int n_frames, int qp,
                                             int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream )
{
    return xd_residual_launch( ctx, g, fenc_slot, pred_slot, n_frames, qp, NULL, NULL, levels, NULL, nnz, cbp, false, stream );
}

extern "C" int x264dsp_residual_frames_typed_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                                   const uint8_t *fenc_slot, uint8_t *pred_slot, int n_frames, int qp,
                                                   const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels,
                                                   int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, void *stream )
{
    return xd_residual_launch( ctx, g, fenc_slot, pred_slot, n_frames, qp, mb_kind, i4_modes, levels, luma_dc, nnz, cbp, true, stream );
}

extern "C" int x264dsp_probe_pskip_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slot,
                                               const uint8_t *pred_slot, int n_frames, int qp, uint8_t *skip, void *stream )
{
    if( !ctx || !g || !fenc_slot || !pred_slot || !skip || qp < 0 || qp > 51 || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    xd_res_tables T;
    memset( &T, 0, sizeof( T ) );
    const int qpc = x264dsp_chroma_qp( qp );
    xd_fill_qparams( &T.luma, qp );
    xd_fill_qparams( &T.chroma, qpc );
    {
        uint16_t mf[16], bias[16];
        x264dsp_quant_tables( 1, qpc, mf, bias );
        T.chroma_dc_mf = mf[0] >> 1;
        T.chroma_dc_bias = bias[0] << 1;
    }
    T.qpc = qpc;
    T.thresh = ( xd_lambda2_tab[qpc] + 32 ) >> 6;
    cudaStream_t s = xd_stream( ctx, stream );
    const dim3 grid( ( g->mb_count + 3 ) / 4, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_RESIDUAL, s );
    // the kernel only reads pred in this mode
    xd_residual_kernel<false, true><<<grid, 128, 0, s>>>( *g, fenc_slot, const_cast<uint8_t *>( pred_slot ), T, NULL, skip, NULL, NULL, NULL );
    xd_prof_end( ctx, XD_PROF_RESIDUAL, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_residual_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g,
                                            const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                                            int16_t *levels, uint8_t *nnz, int16_t *cbp, void *stream )
{
    return x264dsp_residual_frames_dev( ctx, g, fenc_slot, pred_slot, 1, qp, levels, nnz, cbp, stream );
}

extern "C" int x264dsp_mc_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slot,
                                       int n_frames, const int16_t *mv, uint8_t *pred_slot, void *stream )
{
    if( !ctx || !g || !fref_slot || !mv || !pred_slot || fref_slot == pred_slot || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const dim3 grid( ( g->mb_w + 7 ) / 8, g->mb_h, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_MC, s );
    xd_mc_frame_kernel<1><<<grid, 256, 0, s>>>( *g, fref_slot, mv, pred_slot );
    xd_prof_end( ctx, XD_PROF_MC, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_mc_frames_part_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slot,
                                            int n_frames, const int16_t *mv8x8, uint8_t *pred_slot, void *stream )
{
    if( !ctx || !g || !fref_slot || !mv8x8 || !pred_slot || fref_slot == pred_slot || n_frames <= 0 || n_frames > 65535 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    const dim3 grid( ( g->mb_w + 7 ) / 8, g->mb_h, n_frames );
    const int pslot = xd_prof_begin( ctx, XD_PROF_MC, s );
    xd_mc_frame_kernel<4><<<grid, 256, 0, s>>>( *g, fref_slot, mv8x8, pred_slot );
    xd_prof_end( ctx, XD_PROF_MC, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_mc_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fref_slot,
                                      const int16_t *mv, uint8_t *pred_slot, void *stream )
{
    return x264dsp_mc_frames_dev( ctx, g, fref_slot, 1, mv, pred_slot, stream );
}
