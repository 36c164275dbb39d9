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

// ---------------------------------------------------------------------------------------------
// motion compensation: 64 threads per macroblock (4 luma pixels + 1 UV pair each), 4 MBs per CTA.
// NMV = 1: one MV per macroblock (x264_mb_mc, D_16x16); NMV = 4: one MV per 8x8 in raster order, which covers every
// partition the reference analyses -- x264_mb_mc's 16x8 / 8x16 / 8x8 cases are 8x8-wise the same samples, mc_luma and
// mc_chroma being per-pixel rules (common/macroblock.c:8-48)
template<int NMV>
__global__ void __launch_bounds__( 256 )
xd_mc_frame_kernel( x264dsp_geom_t g, const uint8_t *__restrict__ fref, const int16_t *__restrict__ mv,
                    uint8_t *__restrict__ pred )
{
    const int mb = blockIdx.x * 4 + ( threadIdx.x >> 6 );
    if( mb >= g.mb_count )
        return;
    // blockIdx.y = frame of a batch: consecutive slots, mb_count MVs per frame
    fref += blockIdx.y * (size_t)g.slot_bytes;
    pred += blockIdx.y * (size_t)g.slot_bytes;
    mv += blockIdx.y * (size_t)g.mb_count * 2 * NMV;
    const int t = threadIdx.x & 63;
    const int mb_x = mb % g.mb_w, mb_y = mb / g.mb_w;
    const int ls = g.luma_stride, cs = g.chroma_stride;
    // analyse.c:378-390: mv_min / mv_max of the macroblock (quarter-pel)
    const int lo_x = ( -( mb_x << 4 ) - 24 ) << 2, hi_x = ( ( ( g.mb_w - mb_x - 1 ) << 4 ) + 24 ) << 2;
    const int lo_y = ( -( mb_y << 4 ) - 24 ) << 2, hi_y = ( ( ( g.mb_h - mb_y - 1 ) << 4 ) + 24 ) << 2;
    {
        // luma: row t/4, pixels 4*(t%4) .. +3
        const int y = t >> 2, x = ( t & 3 ) * 4;
        const int part = NMV == 4 ? ( y >> 3 ) * 2 + ( x >> 3 ) : 0;
        const int mvx = xd_clip3( mv[2 * ( mb * NMV + part )], lo_x, hi_x ), mvy = xd_clip3( mv[2 * ( mb * NMV + part ) + 1], lo_y, hi_y );
        const int fx = mvx & 3, fy = mvy & 3, phase = fy * 4 + fx;
        const int64_t pos = (int64_t)( ( mb_y << 4 ) + y + ( mvy >> 2 ) ) * ls + ( mb_x << 4 ) + x + ( mvx >> 2 );
        const uint8_t *base = fref + g.luma_origin;
        uint32_t a = xd_load4_unaligned( base + (size_t)xd_qpel_plane_a( phase ) * g.luma_plane_size + pos + ( fy == 3 ? ls : 0 ) );
        if( phase & 5 )
            a = xd_avg4( a, xd_load4_unaligned( base + (size_t)xd_qpel_plane_b( phase ) * g.luma_plane_size + pos + ( fx == 3 ? 1 : 0 ) ) );
        *(uint32_t *)( pred + g.luma_origin + (int64_t)( ( mb_y << 4 ) + y ) * ls + ( mb_x << 4 ) + x ) = a;
    }
    {
        // chroma: row t/8, pair t%8; eighth-pel bilinear on NV12 (mc.c:290-323)
        const int y = t >> 3, x = t & 7;
        const int part = NMV == 4 ? ( y >> 2 ) * 2 + ( x >> 2 ) : 0;
        const int mvx = xd_clip3( mv[2 * ( mb * NMV + part )], lo_x, hi_x ), mvy = xd_clip3( mv[2 * ( mb * NMV + part ) + 1], lo_y, hi_y );
        const int dx = mvx & 7, dy = mvy & 7;
        const int cA = ( 8 - dx ) * ( 8 - dy ), cB = dx * ( 8 - dy ), cC = ( 8 - dx ) * dy, cD = dx * dy;
        const uint8_t *s0 = fref + g.slot_chroma_off + g.chroma_origin
                          + (int64_t)( ( mb_y << 3 ) + y + ( mvy >> 3 ) ) * cs + ( mb_x << 4 ) + 2 * ( x + ( mvx >> 3 ) );
        // the four bytes U0 V0 U1 V1 of each of the two rows as one (unaligned) word instead of eight byte loads
        const uint32_t r0 = xd_load4_unaligned( s0 ), r1 = xd_load4_unaligned( s0 + cs );
        const int u = ( cA * (int)( r0 & 255 ) + cB * (int)( ( r0 >> 16 ) & 255 ) + cC * (int)( r1 & 255 ) + cD * (int)( ( r1 >> 16 ) & 255 ) + 32 ) >> 6;
        const int v = ( cA * (int)( ( r0 >> 8 ) & 255 ) + cB * (int)( r0 >> 24 ) + cC * (int)( ( r1 >> 8 ) & 255 ) + cD * (int)( r1 >> 24 ) + 32 ) >> 6;
        *(uint16_t *)( pred + g.slot_chroma_off + g.chroma_origin + (int64_t)( ( mb_y << 3 ) + y ) * cs + ( mb_x << 4 ) + 2 * x )
            = (uint16_t)( u | ( v << 8 ) );
    }
}

struct xd_res_tables
{
    xd_qparams luma, chroma;
    int chroma_dc_mf, chroma_dc_bias, chroma_dmf_full;   // mf[0]>>1, bias[0]<<1, dequant_mf[qpc%6][0] << qpc/6
    int qpc, thresh;                                      // chroma qp, (lambda2[qpc]+32)>>6
    // I16x16 macroblocks of I slices (typed entry point): the intra quant tables (CQM_4IY / CQM_4IC differ from the
    // inter ones in the rounding bias only) and the luma DC block's scalars (macroblock.c:123, quant.c:83-101)
    xd_qparams luma_i, chroma_i;
    int chroma_dc_bias_i;
    int luma_dc_mf, luma_dc_bias, luma_dc_dmf, luma_dc_qbits;   // mf[0]>>1, bias[0]<<1, dequant_mf[qp%6][0], qp/6 - 6
};

// block_idx_xy_1d (common/macroblock.h): coding index of a luma 4x4 -> raster index x + 4 y
__device__ __forceinline__ int xd_blk_raster( int i )
{
    return ( ( i & 1 ) + ( ( i >> 2 ) & 1 ) * 2 ) + 4 * ( ( ( i >> 1 ) & 1 ) + ( ( i >> 3 ) & 1 ) * 2 );
}

// dct4x4dc / idct4x4dc (dct.c:36-100): 4x4 Hadamard of the sixteen luma DC terms, every lane on the same values
__device__ __forceinline__ void xd_hadamard_dc( int d[16], bool halve )
{
    int t[16];
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        const int s01 = d[4 * i] + d[4 * i + 1], d01 = d[4 * i] - d[4 * i + 1];
        const int s23 = d[4 * i + 2] + d[4 * i + 3], d23 = d[4 * i + 2] - d[4 * i + 3];
        t[i] = (int16_t)( s01 + s23 ); t[4 + i] = (int16_t)( s01 - s23 );
        t[8 + i] = (int16_t)( d01 - d23 ); t[12 + i] = (int16_t)( d01 + d23 );
    }
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        const int s01 = t[4 * i] + t[4 * i + 1], d01 = t[4 * i] - t[4 * i + 1];
        const int s23 = t[4 * i + 2] + t[4 * i + 3], d23 = t[4 * i + 2] - t[4 * i + 3];
        const int r = halve ? 1 : 0;
        d[4 * i] = (int16_t)( ( s01 + s23 + r ) >> r ); d[4 * i + 1] = (int16_t)( ( s01 - s23 + r ) >> r );
        d[4 * i + 2] = (int16_t)( ( d01 - d23 + r ) >> r ); d[4 * i + 3] = (int16_t)( ( d01 + d23 + r ) >> r );
    }
}

// TYPED: mb_kind[mb] != 0 marks an I16x16 macroblock of an I slice (x264_mb_encode_i16x16, macroblock.c:72-162, and
// x264_mb_encode_chroma with b_inter = 0, no decimation); the inter-only instantiation carries none of that code
// PROBE (third instantiation): x264_macroblock_probe_pskip (macroblock.c:492-604) on the P_SKIP prediction in `pred` --
// same transforms and quantisers, but nothing is stored except one flag per macroblock (nnz_out[mb] = 1: skippable);
// the early exits of the reference are sums here (scores only grow, so "ever >= 6" is "total >= 6")
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
    bool intra = false, i16 = false, i4 = false;
    if( TYPED )
    {
        if( mb_kind )
        {
            const int kind = mb_kind[blockIdx.y * (size_t)g.mb_count + mb] & 3;
            intra = kind != 0;
            i16 = kind == 1;
            i4 = kind == 2;         // chroma here (intra rules); luma, its levels / nnz / cbp bits in xd_intra4_kernel
        }
        if( luma_dc )
        {
            luma_dc += ( blockIdx.y * (size_t)g.mb_count + mb ) * 16;
            if( !i16 && lane < 2 )
                ( (uint4 *)luma_dc )[lane] = make_uint4( 0u, 0u, 0u, 0u );
        }
    }
    const int mb_x = mb % g.mb_w, mb_y = mb / g.mb_w;
    const bool is_luma = lane < 16, is_chroma = lane >= 16 && lane < 24;
    const int ch = ( lane - 16 ) >> 2, ci = ( lane - 16 ) & 3;

    // ---- load the lane's 4x4 source and prediction rows
    uint32_t f[4] = { 0, 0, 0, 0 }, p[4] = { 0, 0, 0, 0 };
    int64_t luma_off = 0, chroma_off = 0;
    if( is_luma )
    {
        const int bx = ( ( lane & 1 ) + ( ( lane >> 2 ) & 1 ) * 2 ) * 4, by = ( ( ( lane >> 1 ) & 1 ) + ( ( lane >> 3 ) & 1 ) * 2 ) * 4;
        luma_off = g.luma_origin + (int64_t)( ( mb_y << 4 ) + by ) * g.luma_stride + ( mb_x << 4 ) + bx;
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            f[r] = __ldg( (const uint32_t *)( fenc + luma_off + (int64_t)r * g.luma_stride ) );
            p[r] = *(const uint32_t *)( pred + luma_off + (int64_t)r * g.luma_stride );
        }
    }
    else if( is_chroma )
    {
        const int bx = ( ci & 1 ) * 4, by = ( ci >> 1 ) * 4;
        chroma_off = g.slot_chroma_off + g.chroma_origin + (int64_t)( ( mb_y << 3 ) + by ) * g.chroma_stride + ( mb_x << 4 ) + 2 * bx;
        const uint32_t sel = ch ? 0x7531 : 0x6420;
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint2 a = __ldg( (const uint2 *)( fenc + chroma_off + (int64_t)r * g.chroma_stride ) );
            const uint2 b = *(const uint2 *)( pred + chroma_off + (int64_t)r * g.chroma_stride );
            f[r] = __byte_perm( a.x, a.y, sel );
            p[r] = __byte_perm( b.x, b.y, sel );
        }
    }

    int dct[16], lv[16];
    xd_sub4x4_dct( dct, f, p );

    // ---- chroma: variance early-out statistics (macroblock.c:188-196, pixel.c:209-231)
    int diff_sum = dct[0];                         // sum of the 16 differences == DC coefficient
    int diff_sqr = 0;
    if( is_chroma )
    {
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint32_t d = __vabsdiffu4( f[r], p[r] );
            diff_sqr += __dp4a( d, d, 0u );
        }
    }
    int psum = diff_sum, psqr = diff_sqr;          // totals over the 4 lanes of a plane
    psum += __shfl_xor_sync( 0xffffffffu, psum, 1 ); psqr += __shfl_xor_sync( 0xffffffffu, psqr, 1 );
    psum += __shfl_xor_sync( 0xffffffffu, psum, 2 ); psqr += __shfl_xor_sync( 0xffffffffu, psqr, 2 );
    const int sum_u = __shfl_sync( 0xffffffffu, psum, 16 ), sqr_u = __shfl_sync( 0xffffffffu, psqr, 16 );
    const int sum_v = __shfl_sync( 0xffffffffu, psum, 20 ), sqr_v = __shfl_sync( 0xffffffffu, psqr, 20 );
    bool early = false;
    if( T.qpc >= 18 && !intra && !PROBE )
    {
        const unsigned au = (unsigned)abs( sum_u ), av = (unsigned)abs( sum_v );
        const int var_u = (int)( (unsigned)sqr_u - (unsigned)( ( (unsigned long long)au * au ) >> 6 ) );
        const int var_v = (int)( (unsigned)sqr_v - (unsigned)( ( (unsigned long long)av * av ) >> 6 ) );
        early = var_u < ( T.thresh << 2 ) && var_u + var_v < ( T.thresh << 2 );
    }

    // ---- chroma DC of the plane: 2x2 transform of the four DC terms, on every lane of the plane
    const int base4 = lane & ~3;
    const int c0 = __shfl_sync( 0xffffffffu, dct[0], base4 ), c1 = __shfl_sync( 0xffffffffu, dct[0], base4 + 1 );
    const int c2 = __shfl_sync( 0xffffffffu, dct[0], base4 + 2 ), c3 = __shfl_sync( 0xffffffffu, dct[0], base4 + 3 );
    int dc[4] = { (int16_t)( c0 + c1 + c2 + c3 ), (int16_t)( c0 + c1 - c2 - c3 ), (int16_t)( c0 - c1 + c2 - c3 ), (int16_t)( c0 - c1 - c2 + c3 ) };
    // note the reference's ordering d[1] = (c0+c1)-(c2+c3), d[2] = (c0-c1)+(c2-c3)

    const xd_qparams &Q = intra ? ( is_luma ? T.luma_i : T.chroma_i ) : ( is_luma ? T.luma : T.chroma );
    const int luma_dcv = dct[0];                   // I16x16: the DC terms leave for their own 4x4 block (macroblock.c:92-93)
    if( is_chroma || intra )
        dct[0] = 0;                                // dct2x2dc clears the DC terms (macroblock.c:55-58)
    int nz = 0, score = 0;
    if( is_luma || ( is_chroma && !early ) )
    {
        nz = xd_quant_4x4( dct, Q );
        xd_zigzag( lv, dct );
        if( nz )
        {
            xd_dequant_4x4( dct, Q );
            score = xd_decimate( lv, is_luma ? 0 : 1 );
        }
    }
    else
    {
#pragma unroll
        for( int i = 0; i < 16; i++ )
            lv[i] = 0;
    }
    if( PROBE )
    {
        // luma: the decimate scores of all coded 4x4s; chroma per plane: SSD below thresh passes, a coded DC fails,
        // SSD below 4 thresh passes, else the AC decimate scores decide (macroblock.c:510-600)
        int luma_sum = is_luma ? score : 0, plane_sum = is_chroma ? score : 0;
#pragma unroll
        for( int o = 1; o < 16; o <<= 1 )
            luma_sum += __shfl_xor_sync( 0xffffffffu, luma_sum, o );
        luma_sum = __shfl_sync( 0xffffffffu, luma_sum, 0 );
        plane_sum += __shfl_xor_sync( 0xffffffffu, plane_sum, 1 );
        plane_sum += __shfl_xor_sync( 0xffffffffu, plane_sum, 2 );
        int nz_dc = 0;
#pragma unroll
        for( int i = 0; i < 4; i++ )
            nz_dc |= xd_quant1( dc[i], T.chroma_dc_mf, T.chroma_dc_bias );
        const int ssd = ch ? sqr_v : sqr_u;
        const bool plane_fail = is_chroma && ssd >= T.thresh && ( nz_dc != 0 || ( ssd >= ( T.thresh << 2 ) && plane_sum >= 7 ) );
        const bool fail = luma_sum >= 6 || __any_sync( 0xffffffffu, plane_fail );
        if( lane == 0 )
            nnz_out[mb] = fail ? 0 : 1;
        return;
    }
    int16_t *mb_levels = levels + (size_t)mb * X264DSP_RES_LEVELS_PER_MB;
    if( is_luma && !i4 )
        xd_store_levels( mb_levels + lane * 16, lv );
    else if( is_chroma )
        xd_store_levels( mb_levels + 264 + ( lane - 16 ) * 16, lv );

    // ---- luma decimation (macroblock.c:394-452): scores add up in coding order while < 6
    const int s0 = __shfl_sync( 0xffffffffu, score, base4 ), s1 = __shfl_sync( 0xffffffffu, score, base4 + 1 );
    const int s2 = __shfl_sync( 0xffffffffu, score, base4 + 2 ), s3 = __shfl_sync( 0xffffffffu, score, base4 + 3 );
    int score8 = s0;
    if( score8 < 6 ) score8 += s1;
    if( score8 < 6 ) score8 += s2;
    if( score8 < 6 ) score8 += s3;
    // (a block that quantised to zero has score 0, which is what "skipped" adds)
    const int mb_score = __shfl_sync( 0xffffffffu, score8, 0 ) + __shfl_sync( 0xffffffffu, score8, 4 )
                       + __shfl_sync( 0xffffffffu, score8, 8 ) + __shfl_sync( 0xffffffffu, score8, 12 );
    bool keep8 = score8 >= 4 && mb_score >= 6;
    int nnz_flag = 0;
    int nz_luma_dc = 0;
    if( TYPED && i16 )
    {
        // no decimation in an I slice: all sixteen blocks are coded as soon as one of them has a coefficient
        const bool any_ac = ( __ballot_sync( 0xffffffffu, is_luma && nz ) ) != 0;
        keep8 = any_ac;
        int d[16];
#pragma unroll
        for( int i = 0; i < 16; i++ )
            d[xd_blk_raster( i )] = __shfl_sync( 0xffffffffu, luma_dcv, i );
        xd_hadamard_dc( d, true );
#pragma unroll
        for( int i = 0; i < 16; i++ )
        {
            d[i] = xd_quant1( d[i], T.luma_dc_mf, T.luma_dc_bias );
            nz_luma_dc |= d[i];
        }
        nz_luma_dc = nz_luma_dc != 0;
        int dlv[16];
        xd_zigzag( dlv, d );
        if( !nz_luma_dc )
        {
#pragma unroll
            for( int i = 0; i < 16; i++ )
                dlv[i] = 0;
        }
        if( lane == 0 && luma_dc )
            xd_store_levels( luma_dc, dlv );
        int my_dc = 0;
        if( nz_luma_dc )
        {
            xd_hadamard_dc( d, false );
            if( T.luma_dc_qbits >= 0 )
            {
#pragma unroll
                for( int i = 0; i < 16; i++ )
                    d[i] = (int16_t)( d[i] * ( T.luma_dc_dmf << T.luma_dc_qbits ) );
            }
            else
            {
                const int f = 1 << ( -T.luma_dc_qbits - 1 );
#pragma unroll
                for( int i = 0; i < 16; i++ )
                    d[i] = (int16_t)( ( d[i] * T.luma_dc_dmf + f ) >> ( -T.luma_dc_qbits ) );
            }
            const int mine = xd_blk_raster( lane & 15 );
#pragma unroll
            for( int i = 0; i < 16; i++ )
                if( mine == i )
                    my_dc = d[i];
        }
        if( is_luma )
        {
            nnz_flag = nz;
            if( any_ac )
            {
                dct[0] = my_dc;
                xd_add4x4_idct( p, dct );
            }
            else if( nz_luma_dc )
                xd_add4x4_dc( p, my_dc );
        }
    }
    else if( is_luma && !i4 )
    {
        nnz_flag = keep8 ? nz : 0;
        if( keep8 )
            xd_add4x4_idct( p, dct );
    }
    const unsigned keep_mask = __ballot_sync( 0xffffffffu, is_luma && keep8 && !i4 );
    const int cbp_luma = ( ( keep_mask >> 0 ) & 1 ) | ( ( ( keep_mask >> 4 ) & 1 ) << 1 )
                       | ( ( ( keep_mask >> 8 ) & 1 ) << 2 ) | ( ( ( keep_mask >> 12 ) & 1 ) << 3 );

    // ---- chroma (macroblock.c:175-305)
    int nz_dc_final = 0, plane_cbp = 0;
    int dc_levels[4] = { 0, 0, 0, 0 };
    if( is_chroma )
    {
        const int dmf = T.chroma_dmf_full;
        const int psc = s0 + s1 + s2 + s3;                          // decimate score of the plane
        const unsigned nzmask = __ballot_sync( 0x00ff0000u, nz != 0 );
        const bool nz_ac = ( ( nzmask >> base4 ) & 15 ) != 0;
        const int ssd = ch ? sqr_v : sqr_u;
        bool do_dc = false, ac_coded = false;
        if( early )
            do_dc = ssd > T.thresh;
        else
        {
            ac_coded = intra ? nz_ac : !( psc < 7 || !nz_ac );
            do_dc = true;
        }
        int nz_dc = 0;
        if( do_dc )
        {
            nz_dc = 0;
#pragma unroll
            for( int i = 0; i < 4; i++ )
            {
                dc[i] = xd_quant1( dc[i], T.chroma_dc_mf, intra ? T.chroma_dc_bias_i : T.chroma_dc_bias );
                nz_dc |= dc[i];
            }
            nz_dc = nz_dc != 0;
        }
        nz_dc_final = nz_dc;
        if( nz_dc && !ac_coded && T.qpc <= 22 )
        {
            // every lane of the plane runs the same scalar optimiser on the same values
            if( !xd_optimize_chroma_dc( dc, dmf ) )
                nz_dc_final = 0;
        }
        if( ac_coded )
        {
            plane_cbp = 1;
            if( nz_dc )
            {
                // idct_dequant_2x2_dc (macroblock.c:17-29): this lane's DC term
                const int a = dc[0] + dc[1], b = dc[2] + dc[3], c = dc[0] - dc[1], d = dc[2] - dc[3];
                const int rec = ci == 0 ? a + b : ci == 1 ? a - b : ci == 2 ? c + d : c - d;
                dct[0] = (int16_t)( rec * ( dmf >> 5 ) );
            }
            xd_add4x4_idct( p, dct );
            nnz_flag = nz;
        }
        else
        {
            nnz_flag = 0;
            if( nz_dc_final )
            {
                const int a = dc[0] + dc[1], b = dc[2] + dc[3], c = dc[0] - dc[1], d = dc[2] - dc[3];
                const int rec = ci == 0 ? a + b : ci == 1 ? a - b : ci == 2 ? c + d : c - d;
                xd_add4x4_dc( p, (int16_t)( rec * ( dmf >> 5 ) ) );
                if( early )
                    plane_cbp = 1;
            }
        }
        if( nz_dc_final )
        {
            dc_levels[0] = dc[0]; dc_levels[1] = dc[2]; dc_levels[2] = dc[1]; dc_levels[3] = dc[3];
        }
        if( ci == 0 )
            *(uint2 *)( mb_levels + 256 + 4 * ch ) = make_uint2( ( dc_levels[0] & 0xFFFF ) | ( dc_levels[1] << 16 ),
                                                                 ( dc_levels[2] & 0xFFFF ) | ( dc_levels[3] << 16 ) );
    }

    // ---- stores: reconstruction, flags
    if( is_luma && !i4 )
    {
#pragma unroll
        for( int r = 0; r < 4; r++ )
            *(uint32_t *)( pred + luma_off + (int64_t)r * g.luma_stride ) = p[r];
    }
    // interleave U (lanes 16..19) with V (lanes 20..23)
    uint32_t other[4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
        other[r] = __shfl_sync( 0xffffffffu, p[r], is_chroma ? lane ^ 4 : lane );
    if( is_chroma && ch == 0 )
    {
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            const uint32_t lo = __byte_perm( p[r], other[r], 0x5140 ), hi = __byte_perm( p[r], other[r], 0x7362 );
            *(uint2 *)( pred + chroma_off + (int64_t)r * g.chroma_stride ) = make_uint2( lo, hi );
        }
    }

    const int dc_u = __shfl_sync( 0xffffffffu, nz_dc_final, 16 ), dc_v = __shfl_sync( 0xffffffffu, nz_dc_final, 20 );
    const int pc_u = __shfl_sync( 0xffffffffu, plane_cbp, 16 ), pc_v = __shfl_sync( 0xffffffffu, plane_cbp, 20 );
    int cbp_chroma = pc_u | pc_v;
    if( !early )
        cbp_chroma += dc_u | dc_v | cbp_chroma;                          // macroblock.c:303-304
    uint8_t *mb_nnz = nnz_out + (size_t)mb * X264DSP_RES_NNZ_PER_MB;
    if( lane < 24 && !( i4 && lane < 16 ) )
        mb_nnz[lane] = (uint8_t)nnz_flag;
    if( lane == 24 )
        mb_nnz[24] = (uint8_t)nz_luma_dc;                                // luma DC: I16x16 only
    if( lane == 25 )
        mb_nnz[25] = (uint8_t)dc_u;
    if( lane == 26 )
        mb_nnz[26] = (uint8_t)dc_v;
    if( lane == 0 )
        cbp_out[mb] = (int16_t)( ( cbp_chroma << 4 ) | cbp_luma | ( nz_luma_dc << 8 ) | ( dc_u << 9 ) | ( dc_v << 10 ) );
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

static int xd_residual_launch( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot,
                               int n_frames, int qp, const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels,
                               int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, bool typed, void *stream )
{
    if( !ctx || !g || !fenc_slot || !pred_slot || !levels || !nnz || !cbp || qp < 0 || qp > 51 || n_frames <= 0
        || n_frames > 65535 )
        return X264DSP_E_ARG;
    xd_res_tables T;
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
    const dim3 grid( ( g->mb_count + 3 ) / 4, n_frames );
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
    const dim3 grid( ( g->mb_count + 3 ) / 4, n_frames );
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
