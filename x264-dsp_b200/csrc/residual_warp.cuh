// residual_warp.cuh -- one macroblock's motion compensation (64 threads) and residual coding / P_SKIP probe (one warp) as
// device routines, shared by the frame kernels of residual.cu and the P-slice wavefront of pframe.cu.
#pragma once
#include "common.cuh"
#include "leaf.cuh"

// ---------------------------------------------------------------------------------------------
// motion compensation: 64 threads per macroblock (4 luma pixels + 1 UV pair each), 4 MBs per CTA.
// NMV = 1: one MV per macroblock (x264_mb_mc, D_16x16); NMV = 4: one MV per 8x8 in raster order, which covers every
// partition the reference analyses -- x264_mb_mc's 16x8 / 8x16 / 8x8 cases are 8x8-wise the same samples, mc_luma and
// mc_chroma being per-pixel rules (common/macroblock.c:8-48)
// one macroblock's prediction, thread t of 64: mvs = the macroblock's NMV vectors (quarter-pel, unclipped)
template<int NMV>
__device__ __forceinline__ void xd_mc_mb( const x264dsp_geom_t &g, const uint8_t *__restrict__ fref, const int16_t *mvs,
                                          uint8_t *pred, int mb, int t )
{
    const int mb_x = mb % g.mb_w, mb_y = mb / g.mb_w;
    const int ls = g.luma_stride, cs = g.chroma_stride;
    // analyse.c:378-390: mv_min / mv_max of the macroblock (quarter-pel)
    const int lo_x = ( -( mb_x << 4 ) - 24 ) << 2, hi_x = ( ( ( g.mb_w - mb_x - 1 ) << 4 ) + 24 ) << 2;
    const int lo_y = ( -( mb_y << 4 ) - 24 ) << 2, hi_y = ( ( ( g.mb_h - mb_y - 1 ) << 4 ) + 24 ) << 2;
    {
        // luma: row t/4, pixels 4*(t%4) .. +3
        const int y = t >> 2, x = ( t & 3 ) * 4;
        const int part = NMV == 4 ? ( y >> 3 ) * 2 + ( x >> 3 ) : 0;
        const int mvx = xd_clip3( mvs[2 * part], lo_x, hi_x ), mvy = xd_clip3( mvs[2 * part + 1], lo_y, hi_y );
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
        const int mvx = xd_clip3( mvs[2 * part], lo_x, hi_x ), mvy = xd_clip3( mvs[2 * part + 1], lo_y, hi_y );
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

// The same prediction by ONE WARP (the mapping every caller uses now): lane = half a luma row (8 pixels: one unaligned
// 8-byte fetch per plane instead of two 4-byte ones, the vector clip and the address arithmetic once per 8 pixels) and two
// adjacent UV pairs of a chroma row, whose four taps are one byte dot product per sample: the weights (8-dx)(8-dy), dx(8-dy),
// (8-dx)dy, dx dy are at most 64 and go into one word, the taps U(x,y) U(x+1,y) U(x,y+1) U(x+1,y+1) are picked out of the
// two rows' words with PRMT (IDP.4A on the FMA pipe; the first version spent sixteen integer instructions per sample).
// Measured with tools/bench_paths.py on 1080p: see DESIGN.md section 3.
template<int NMV>
__device__ __forceinline__ void xd_mc_mb32( const x264dsp_geom_t &g, const uint8_t *__restrict__ fref, const int16_t *mvs,
                                            uint8_t *pred, int mb_x, int mb_y, int lane )
{
    const int ls = g.luma_stride, cs = g.chroma_stride;
    const int lo_x = ( -( mb_x << 4 ) - 24 ) << 2, hi_x = ( ( ( g.mb_w - mb_x - 1 ) << 4 ) + 24 ) << 2;
    const int lo_y = ( -( mb_y << 4 ) - 24 ) << 2, hi_y = ( ( ( g.mb_h - mb_y - 1 ) << 4 ) + 24 ) << 2;
    {
        // luma: row lane/2, pixels 8*(lane%2) .. +7
        const int y = lane >> 1, x = ( lane & 1 ) * 8;
        const int part = NMV == 4 ? ( y >> 3 ) * 2 + ( lane & 1 ) : 0;
        const int mvx = xd_clip3( mvs[2 * part], lo_x, hi_x ), mvy = xd_clip3( mvs[2 * part + 1], lo_y, hi_y );
        const int fx = mvx & 3, fy = mvy & 3, phase = fy * 4 + fx;
        const int64_t pos = (int64_t)( ( mb_y << 4 ) + y + ( mvy >> 2 ) ) * ls + ( mb_x << 4 ) + x + ( mvx >> 2 );
        const uint8_t *base = fref + g.luma_origin;
        uint2 a = xd_load8_unaligned( base + (size_t)xd_qpel_plane_a( phase ) * g.luma_plane_size + pos + ( fy == 3 ? ls : 0 ) );
        if( phase & 5 )
        {
            const uint2 b = xd_load8_unaligned( base + (size_t)xd_qpel_plane_b( phase ) * g.luma_plane_size + pos + ( fx == 3 ? 1 : 0 ) );
            a.x = xd_avg4( a.x, b.x );
            a.y = xd_avg4( a.y, b.y );
        }
        *(uint2 *)( pred + g.luma_origin + (int64_t)( ( mb_y << 4 ) + y ) * ls + ( mb_x << 4 ) + x ) = a;
    }
    {
        // chroma: row lane/4, UV pairs 2*(lane%4) and 2*(lane%4)+1; eighth-pel bilinear on NV12 (mc.c:290-323)
        const int y = lane >> 2, x = ( lane & 3 ) * 2;
        const int part = NMV == 4 ? ( y >> 2 ) * 2 + ( x >> 2 ) : 0;
        const int mvx = xd_clip3( mvs[2 * part], lo_x, hi_x ), mvy = xd_clip3( mvs[2 * part + 1], lo_y, hi_y );
        const int dx = mvx & 7, dy = mvy & 7;
        const uint32_t coef = (uint32_t)( ( 8 - dx ) * ( 8 - dy ) ) | ( (uint32_t)( dx * ( 8 - dy ) ) << 8 )
                            | ( (uint32_t)( ( 8 - dx ) * dy ) << 16 ) | ( (uint32_t)( dx * dy ) << 24 );
        const uint8_t *s0 = fref + g.slot_chroma_off + g.chroma_origin
                          + (int64_t)( ( mb_y << 3 ) + y + ( mvy >> 3 ) ) * cs + ( mb_x << 4 ) + 2 * ( x + ( mvx >> 3 ) );
        // bytes U0 V0 U1 V1 U2 V2 of the two rows (the 8-byte fetch brings two more)
        const uint2 r0 = xd_load8_unaligned( s0 ), r1 = xd_load8_unaligned( s0 + cs );
        const uint32_t m0 = __funnelshift_r( r0.x, r0.y, 16 ), m1 = __funnelshift_r( r1.x, r1.y, 16 );     // U1 V1 U2 V2
        const uint32_t u0 = ( __dp4a( __byte_perm( r0.x, r1.x, 0x6420 ), coef, 32u ) ) >> 6;
        const uint32_t v0 = ( __dp4a( __byte_perm( r0.x, r1.x, 0x7531 ), coef, 32u ) ) >> 6;
        const uint32_t u1 = ( __dp4a( __byte_perm( m0, m1, 0x6420 ), coef, 32u ) ) >> 6;
        const uint32_t v1 = ( __dp4a( __byte_perm( m0, m1, 0x7531 ), coef, 32u ) ) >> 6;
        *(uint32_t *)( pred + g.slot_chroma_off + g.chroma_origin + (int64_t)( ( mb_y << 3 ) + y ) * cs + ( mb_x << 4 ) + 2 * x )
            = u0 | ( v0 << 8 ) | ( u1 << 16 ) | ( v1 << 24 );
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

void xd_residual_tables( int qp, xd_res_tables *out );      // residual.cu (host)

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
// One macroblock by one warp; fenc / pred / the output arrays are the FRAME's (indexed by mb).  Returns h->mb.cbp of the
// macroblock on every lane (PROBE: 1 when x264_macroblock_probe_pskip would return 1, else 0).
template<bool TYPED, bool PROBE>
__device__ __forceinline__ int xd_residual_mb( const x264dsp_geom_t &g, const uint8_t *__restrict__ fenc, uint8_t *pred,
                                               const xd_res_tables &T, int16_t *levels, uint8_t *nnz_out, int16_t *cbp_out,
                                               const uint8_t *mb_kind, int16_t *luma_dc, int mb, int lane )
{
    bool intra = false, i16 = false, i4 = false;
    if( TYPED )
    {
        if( mb_kind )
        {
            const int kind = mb_kind[mb] & 3;
            intra = kind != 0;
            i16 = kind == 1;
            i4 = kind == 2;         // chroma here (intra rules); luma, its levels / nnz / cbp bits in xd_intra4_kernel
        }
        if( luma_dc )
        {
            luma_dc += (size_t)mb * 16;
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
        uint32_t codes;
        nz = xd_quant_4x4_codes( dct, Q, codes );
        xd_zigzag( lv, dct );
        if( nz )
        {
            xd_dequant_4x4( dct, Q );
            score = xd_decimate_codes( codes, is_luma ? 0 : 1 );
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
        if( lane == 0 && nnz_out )
            nnz_out[mb] = fail ? 0 : 1;
        return fail ? 0 : 1;
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
    const int cbp_all = ( cbp_chroma << 4 ) | cbp_luma | ( nz_luma_dc << 8 ) | ( dc_u << 9 ) | ( dc_v << 10 );
    if( lane == 0 )
        cbp_out[mb] = (int16_t)cbp_all;
    return cbp_all;
}

