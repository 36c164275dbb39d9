// leaf.cuh -- register-level device routines shared by the batched kernels and the per-call
// table shims: 4x4 transform / quant pipeline (common/dct.c, common/quant.c), Hadamard cost
// (common/pixel.c:243-314), deblocking line filters (common/deblock.c:80-295).
#pragma once
#include "common.cuh"

static __constant__ uint8_t xd_blk_w[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static __constant__ uint8_t xd_blk_h[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

// ---------------------------------------------------------------------------------------------
// Hadamard cost of one 4x4: sum |H4 (a-b) H4^T|, rows given as packed pixels (not yet halved)
//
// The horizontal stage runs on the FMA pipe as byte dot products with the four +-1 rows of H4 (IDP.4A, u8 x s8):
// coefficient k of row r of the DIFFERENCE is dp4a( a_r, h_k ) + dp4a( b_r, -h_k ), so the pixels are never unpacked
// and never subtracted one by one.  The vertical stage is a 32-bit butterfly whose last level is folded into the
// absolute sum: |x + y| + |x - y| = 2 max( |x|, |y| ).  32 IDP + ~48 ALU instructions per 4x4 against ~170 for the
// unpack / subtract / butterfly form this replaces (the order of the coefficients does not matter, only their sum).
__device__ __forceinline__ int xd_dp4a_u8s8( uint32_t a, uint32_t b, int c )
{
    int d;
    asm( "dp4a.u32.s32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( c ) );
    return d;
}

// half of the Hadamard cost of one 4x4 -- exactly x264_pixel_satd_4x4's return value (pixel.c:267-291)
__device__ __forceinline__ int xd_satd4x4( const uint32_t a[4], const uint32_t b[4] )
{
    // rows of H4 as s8x4, and their negations
    constexpr uint32_t P0 = 0x01010101u, P1 = 0xFFFF0101u, P2 = 0x01FFFF01u, P3 = 0xFF01FF01u;
    constexpr uint32_t N0 = 0xFFFFFFFFu, N1 = 0x0101FFFFu, N2 = 0xFF0101FFu, N3 = 0x01FF01FFu;
    int t[4][4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        t[r][0] = xd_dp4a_u8s8( a[r], P0, xd_dp4a_u8s8( b[r], N0, 0 ) );
        t[r][1] = xd_dp4a_u8s8( a[r], P1, xd_dp4a_u8s8( b[r], N1, 0 ) );
        t[r][2] = xd_dp4a_u8s8( a[r], P2, xd_dp4a_u8s8( b[r], N2, 0 ) );
        t[r][3] = xd_dp4a_u8s8( a[r], P3, xd_dp4a_u8s8( b[r], N3, 0 ) );
    }
    int acc = 0;
#pragma unroll
    for( int c = 0; c < 4; c++ )
    {
        const int s01 = t[0][c] + t[1][c], m01 = t[0][c] - t[1][c];
        const int s23 = t[2][c] + t[3][c], m23 = t[2][c] - t[3][c];
        acc += max( abs( s01 ), abs( s23 ) ) + max( abs( m01 ), abs( m23 ) );
    }
    return acc;
}

__device__ __forceinline__ int xd_had_abs4x4( const uint32_t a[4], const uint32_t b[4] )
{
    return 2 * xd_satd4x4( a, b );
}

__device__ __forceinline__ uint32_t xd_sq4( uint32_t a, uint32_t b )
{
    // sum of squared differences of four packed pixels
    const uint32_t d = __vabsdiffu4( a, b );
    return __dp4a( d, d, 0u );
}

// ---------------------------------------------------------------------------------------------
// 4x4 pipeline in registers

__device__ __forceinline__ void xd_fwd4( int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3 )
{
    const int s_ad = a + d, s_bc = b + c, d_ad = a - d, d_bc = b - c;
    o0 = s_ad + s_bc; o1 = 2 * d_ad + d_bc; o2 = s_ad - s_bc; o3 = d_ad - 2 * d_bc;
}

__device__ __forceinline__ void xd_inv4( int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3 )
{
    const int e = a + c, f = a - c, gg = b + ( d >> 1 ), hh = ( b >> 1 ) - d;
    o0 = e + gg; o1 = f + hh; o2 = f - hh; o3 = e - gg;
}

// residual 4x4 DCT: f, p = four packed rows each (dct.c:115-150).  The horizontal stage of a row is four dot products of
// its differences with (1,1,1,1) (2,1,-1,-2) (1,-1,-1,1) (1,-2,2,-1); a difference never has to exist as a number: the source
// bytes take the weights, the prediction bytes their negatives (IDP.4A u8 x s8, FMA pipe).  8 IDP per row against 20 ALU
// instructions for unpack / subtract / butterfly.
__device__ __forceinline__ void xd_sub4x4_dct( int dct[16], const uint32_t f[4], const uint32_t p[4] )
{
    int t[16];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        t[r]      = xd_dp4a_u8s8( f[r], 0x01010101u, xd_dp4a_u8s8( p[r], 0xFFFFFFFFu, 0 ) );
        t[4 + r]  = xd_dp4a_u8s8( f[r], 0xFEFF0102u, xd_dp4a_u8s8( p[r], 0x0201FFFEu, 0 ) );
        t[8 + r]  = xd_dp4a_u8s8( f[r], 0x01FFFF01u, xd_dp4a_u8s8( p[r], 0xFF0101FFu, 0 ) );
        t[12 + r] = xd_dp4a_u8s8( f[r], 0xFF02FE01u, xd_dp4a_u8s8( p[r], 0x01FE02FFu, 0 ) );
    }
#pragma unroll
    for( int i = 0; i < 4; i++ )
        xd_fwd4( t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3], dct[4 * i], dct[4 * i + 1], dct[4 * i + 2], dct[4 * i + 3] );
}

// inverse transform + add to the prediction rows, clipped (dct.c:197-235); coefficients are
// truncated to 16 bits between the stages exactly as the reference's dctcoef stores do
__device__ __forceinline__ void xd_add4x4_idct( uint32_t p[4], const int dct[16] )
{
    int t[16], r[16];
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int o0, o1, o2, o3;
        xd_inv4( dct[i], dct[4 + i], dct[8 + i], dct[12 + i], o0, o1, o2, o3 );
        t[4 * i] = (int16_t)o0; t[4 * i + 1] = (int16_t)o1; t[4 * i + 2] = (int16_t)o2; t[4 * i + 3] = (int16_t)o3;
    }
#pragma unroll
    for( int i = 0; i < 4; i++ )
    {
        int o0, o1, o2, o3;
        xd_inv4( t[i], t[4 + i], t[8 + i], t[12 + i], o0, o1, o2, o3 );
        r[i] = (int16_t)( ( o0 + 32 ) >> 6 ); r[4 + i] = (int16_t)( ( o1 + 32 ) >> 6 );
        r[8 + i] = (int16_t)( ( o2 + 32 ) >> 6 ); r[12 + i] = (int16_t)( ( o3 + 32 ) >> 6 );
    }
    // prediction byte + residual = one dot product of the packed row with a unit weight, accumulated onto the residual;
    // clip and pack two samples per I2IP
#pragma unroll
    for( int y = 0; y < 4; y++ )
        p[y] = xd_pack_sat4( xd_dp4a_u8s8( p[y], 0x00000001u, r[4 * y] ), xd_dp4a_u8s8( p[y], 0x00000100u, r[4 * y + 1] ),
                             xd_dp4a_u8s8( p[y], 0x00010000u, r[4 * y + 2] ), xd_dp4a_u8s8( p[y], 0x01000000u, r[4 * y + 3] ) );
}

__device__ __forceinline__ void xd_add4x4_dc( uint32_t p[4], int dc )
{
    dc = (int16_t)( ( dc + 32 ) >> 6 );
#pragma unroll
    for( int y = 0; y < 4; y++ )
        p[y] = xd_pack_sat4( xd_dp4a_u8s8( p[y], 0x00000001u, dc ), xd_dp4a_u8s8( p[y], 0x00000100u, dc ),
                             xd_dp4a_u8s8( p[y], 0x00010000u, dc ), xd_dp4a_u8s8( p[y], 0x01000000u, dc ) );
}

// quant.c:29-36: coef > 0 ? (bias + coef) * mf >> 16 : -((bias - coef) * mf >> 16), i.e. (bias + |coef|) * mf >> 16 with the
// sign put back.  |coef| < 2^15, mf <= 26214 and bias * mf <= 2^15 for every table the reference builds, so nothing wraps
// and the dctcoef store truncates nothing; bias * mf is the same for all coefficients of a class (one multiply-add each).
__device__ __forceinline__ uint32_t xd_quant1_abs( int c, int mf, int bias )
{
    return ( (uint32_t)abs( c ) * (uint32_t)mf + (uint32_t)bias * (uint32_t)mf ) >> 16;
}
__device__ __forceinline__ int xd_quant1( int c, int mf, int bias )
{
    const int s = c >> 31;
    return ( (int)xd_quant1_abs( c, mf, bias ) ^ s ) - s;
}

// position class of coefficient i for the flat quant matrices: 0 (even,even) 1 (mixed) 2 (odd,odd)
#define XD_POS_CLASS( i ) ( ( ( i ) & 1 ) + ( ( ( i ) >> 2 ) & 1 ) )

struct xd_qparams
{
    int mf[3], bias[3], dmf[3];      // per position class
    int qbits;                       // qp/6 - 4
};

// position of raster coefficient i in zig-zag order (the inverse of xd_zigzag below)
#define XD_ZZ_POS( i ) ( ( 0xFDC6EB75A8419320ull >> ( 4 * ( i ) ) ) & 15 )

// quant + what the decimation score needs: min(|level|, 2) of every coefficient as a 2-bit code at its ZIG-ZAG position
// (the absolute level exists on the way anyway); xd_decimate_codes then scores a block with a dozen bit operations instead
// of walking sixteen levels.  Returns nz like xd_quant_4x4.
__device__ __forceinline__ int xd_quant_4x4_codes( int dct[16], const xd_qparams &Q, uint32_t &codes )
{
    uint32_t w = 0;
#pragma unroll
    for( int i = 0; i < 16; i++ )
    {
        const int s = dct[i] >> 31;
        const uint32_t q = xd_quant1_abs( dct[i], Q.mf[XD_POS_CLASS( i )], Q.bias[XD_POS_CLASS( i )] );
        dct[i] = ( (int)q ^ s ) - s;
        w += min( q, 2u ) << ( 2 * (int)XD_ZZ_POS( i ) );
    }
    codes = w;
    return w != 0;
}

__device__ __forceinline__ int xd_quant_4x4( int dct[16], const xd_qparams &Q )
{
    int nz = 0;
#pragma unroll
    for( int i = 0; i < 16; i++ )
    {
        dct[i] = xd_quant1( dct[i], Q.mf[XD_POS_CLASS( i )], Q.bias[XD_POS_CLASS( i )] );
        nz |= dct[i];
    }
    return nz != 0;
}

// quant.c:64-81
__device__ __forceinline__ void xd_dequant_4x4( int dct[16], const xd_qparams &Q )
{
    if( Q.qbits >= 0 )
    {
        const int m[3] = { Q.dmf[0] << Q.qbits, Q.dmf[1] << Q.qbits, Q.dmf[2] << Q.qbits };      // (a * b) << s == a * (b << s) mod 2^32
#pragma unroll
        for( int i = 0; i < 16; i++ )
            dct[i] = (int16_t)( dct[i] * m[XD_POS_CLASS( i )] );
    }
    else
    {
        const int f = 1 << ( -Q.qbits - 1 );
#pragma unroll
        for( int i = 0; i < 16; i++ )
            dct[i] = (int16_t)( ( dct[i] * Q.dmf[XD_POS_CLASS( i )] + f ) >> ( -Q.qbits ) );
    }
}

// zig-zag order (dct.c:329-347)
__device__ __forceinline__ void xd_zigzag( int lv[16], const int q[16] )
{
    lv[0] = q[0];   lv[1] = q[4];   lv[2] = q[1];   lv[3] = q[2];
    lv[4] = q[5];   lv[5] = q[8];   lv[6] = q[12];  lv[7] = q[9];
    lv[8] = q[6];   lv[9] = q[3];   lv[10] = q[7];  lv[11] = q[10];
    lv[12] = q[13]; lv[13] = q[14]; lv[14] = q[11]; lv[15] = q[15];
}

// x264_decimate_score_internal (quant.c:226-252) over lv[first..15]
__device__ __forceinline__ int xd_decimate( const int lv[16], int first )
{
    int score = 0, run = 0;
    bool seen = false, big = false;
#pragma unroll
    for( int i = 15; i >= 0; i-- )
    {
        if( i < first )
            continue;
        const int v = lv[i];
        if( v != 0 )
        {
            big |= v > 1 || v < -1;
            if( seen )
                score += run == 0 ? 3 : run <= 2 ? 2 : run <= 5 ? 1 : 0;
            seen = true;
            run = 0;
        }
        else if( seen )
            run++;
    }
    if( seen )
        score += run == 0 ? 3 : run <= 2 ? 2 : run <= 5 ? 1 : 0;
    return big ? 9 : score;
}

// x264_decimate_score_internal on the 2-bit codes of xd_quant_4x4_codes.  A level above 1 anywhere scores 9.  Otherwise every
// nonzero level adds table[run] with run = the zeros between it and the next lower nonzero level (table = 3 2 2 1 1 1 0 ...):
// "run <= k" is "one of the k+1 positions below is set", so the three thresholds are three AND + POPC; the lowest nonzero
// level has nothing below it and counts its zeros down to `first` (1 for chroma AC: position 0 is the DC's) separately.
__device__ __forceinline__ int xd_decimate_codes( uint32_t codes, int first )
{
    if( first )
        codes &= ~3u;
    if( codes & 0xAAAAAAAAu )
        return 9;
    if( !codes )
        return 0;
    const uint32_t n = codes;
    const uint32_t b = ( n << 2 ) | ( n << 4 ) | ( n << 6 ), c = b | ( n << 8 ) | ( n << 10 ) | ( n << 12 );
    const int run0 = ( ( __ffs( (int)n ) - 1 ) >> 1 ) - first;
    return __popc( n & ( n << 2 ) ) + __popc( n & b ) + __popc( n & c ) + ( run0 == 0 ? 3 : run0 <= 2 ? 2 : run0 <= 5 ? 1 : 0 );
}

__device__ __forceinline__ void xd_store_levels( int16_t *dst, const int lv[16] )
{
    uint4 a, b;
    a.x = ( lv[0] & 0xFFFF ) | ( lv[1] << 16 );   a.y = ( lv[2] & 0xFFFF ) | ( lv[3] << 16 );
    a.z = ( lv[4] & 0xFFFF ) | ( lv[5] << 16 );   a.w = ( lv[6] & 0xFFFF ) | ( lv[7] << 16 );
    b.x = ( lv[8] & 0xFFFF ) | ( lv[9] << 16 );   b.y = ( lv[10] & 0xFFFF ) | ( lv[11] << 16 );
    b.z = ( lv[12] & 0xFFFF ) | ( lv[13] << 16 ); b.w = ( lv[14] & 0xFFFF ) | ( lv[15] << 16 );
    ( (uint4 *)dst )[0] = a;
    ( (uint4 *)dst )[1] = b;
}

// quant.c:133-192 on one lane
__device__ __forceinline__ void xd_chroma_dc_recon( int out[4], const int dc[4], int dmf )
{
    const int a = dc[0] + dc[1], b = dc[2] + dc[3], c = dc[0] - dc[1], d = dc[2] - dc[3];
    out[0] = (int16_t)( ( ( a + b ) * dmf >> 5 ) + 32 );
    out[1] = (int16_t)( ( ( a - b ) * dmf >> 5 ) + 32 );
    out[2] = (int16_t)( ( ( c + d ) * dmf >> 5 ) + 32 );
    out[3] = (int16_t)( ( ( c - d ) * dmf >> 5 ) + 32 );
}

__device__ __noinline__ static int xd_optimize_chroma_dc( int dc[4], int dmf )
{
    int want[4], got[4];
    xd_chroma_dc_recon( want, dc, dmf );
    if( !( ( want[0] | want[1] | want[2] | want[3] ) >> 6 ) )
        return 0;
    int nz = 0;
    for( int k = 3; k >= 0; k-- )
    {
        int level = dc[k];
        const int step = level < 0 ? -1 : 1;
        while( level )
        {
            dc[k] = (int16_t)( level - step );
            xd_chroma_dc_recon( got, dc, dmf );
            const int diff = ( want[0] ^ got[0] ) | ( want[1] ^ got[1] ) | ( want[2] ^ got[2] ) | ( want[3] ^ got[3] );
            if( diff >> 6 )
            {
                nz = 1;
                dc[k] = (int16_t)level;
                break;
            }
            level -= step;
        }
    }
    return nz;
}

// ---------------------------------------------------------------------------------------------
// deblocking line filters
#define XD_TC0_ROWS \
{ \
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0}, \
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,1},{0,0,1},{0,0,1}, \
    {0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2}, \
    {1,1,2},{1,2,3},{1,2,3},{2,2,3},{2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6}, \
    {4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},{10,13,20}, \
    {11,15,23},{13,17,25} \
}
static __constant__ int8_t xd_tc0_tab[52][3] = XD_TC0_ROWS;
static const int8_t xd_tc0_host[52][3] = XD_TC0_ROWS;
// tc0 of bS = 1, 2, 3 at a (clamped) indexA, one byte each; indexA < 0: zeros
static inline uint32_t xd_tc0_packed( int ia )
{
    if( ia < 0 )
        return 0;
    return (uint32_t)xd_tc0_host[ia][0] | ( (uint32_t)xd_tc0_host[ia][1] << 8 ) | ( (uint32_t)xd_tc0_host[ia][2] << 16 );
}


struct xd_db_params
{
    int alpha, beta, alphac, betac;     // luma / chroma thresholds from the slice QP
    int ia, iac;                        // clamped indexA (luma, chroma); < 0 : tc0 = 0
    uint32_t tc_luma, tc_chroma;        // tc0 for bS = 1, 2, 3 at ia / iac, one byte each (xd_tc0_packed)
};

__device__ __forceinline__ int xd_tc0( int index_a, int bs )
{
    if( bs == 0 )
        return -1;
    return index_a < 0 ? 0 : xd_tc0_tab[index_a][bs - 1];
}

// bS < 4 luma line: s[0..7] = p3 p2 p1 p0 q0 q1 q2 q3 (deblock.c:80-120)
__device__ __forceinline__ void xd_luma_line( int s[8], int alpha, int beta, int tc0 )
{
    const int p2 = s[1], p1 = s[2], p0 = s[3], q0 = s[4], q1 = s[5], q2 = s[6];
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    int tc = tc0;
    if( abs( p2 - p0 ) < beta )
    {
        if( tc0 )
            s[2] = p1 + xd_clip3( ( ( p2 + ( ( p0 + q0 + 1 ) >> 1 ) ) >> 1 ) - p1, -tc0, tc0 );
        tc++;
    }
    if( abs( q2 - q0 ) < beta )
    {
        if( tc0 )
            s[5] = q1 + xd_clip3( ( ( q2 + ( ( p0 + q0 + 1 ) >> 1 ) ) >> 1 ) - q1, -tc0, tc0 );
        tc++;
    }
    const int delta = xd_clip3( ( ( ( q0 - p0 ) << 2 ) + ( p1 - q1 ) + 4 ) >> 3, -tc, tc );
    s[3] = xd_clip_u8( p0 + delta );
    s[4] = xd_clip_u8( q0 - delta );
}

// bS = 4 luma line (deblock.c:196-243)
__device__ __forceinline__ void xd_luma_intra_line( int s[8], int alpha, int beta )
{
    const int p3 = s[0], p2 = s[1], p1 = s[2], p0 = s[3], q0 = s[4], q1 = s[5], q2 = s[6], q3 = s[7];
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    if( abs( p0 - q0 ) < ( ( alpha >> 2 ) + 2 ) )
    {
        if( abs( p2 - p0 ) < beta )
        {
            s[3] = ( p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4 ) >> 3;
            s[2] = ( p2 + p1 + p0 + q0 + 2 ) >> 2;
            s[1] = ( 2 * p3 + 3 * p2 + p1 + p0 + q0 + 4 ) >> 3;
        }
        else
            s[3] = ( 2 * p1 + p0 + q1 + 2 ) >> 2;
        if( abs( q2 - q0 ) < beta )
        {
            s[4] = ( p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4 ) >> 3;
            s[5] = ( p0 + q0 + q1 + q2 + 2 ) >> 2;
            s[6] = ( 2 * q3 + 3 * q2 + q1 + q0 + p0 + 4 ) >> 3;
        }
        else
            s[4] = ( 2 * q1 + q0 + p1 + 2 ) >> 2;
    }
    else
    {
        s[3] = ( 2 * p1 + p0 + q1 + 2 ) >> 2;
        s[4] = ( 2 * q1 + q0 + p1 + 2 ) >> 2;
    }
}

// chroma line: s = p1 p0 q0 q1 (deblock.c:147-167, 261-278)
__device__ __forceinline__ void xd_chroma_line( int s[4], int alpha, int beta, int tc, bool intra )
{
    const int p1 = s[0], p0 = s[1], q0 = s[2], q1 = s[3];
    if( abs( p0 - q0 ) >= alpha || abs( p1 - p0 ) >= beta || abs( q1 - q0 ) >= beta )
        return;
    if( intra )
    {
        s[1] = ( 2 * p1 + p0 + q1 + 2 ) >> 2;
        s[2] = ( 2 * q1 + q0 + p1 + 2 ) >> 2;
    }
    else
    {
        const int delta = xd_clip3( ( ( ( q0 - p0 ) << 2 ) + ( p1 - q1 ) + 4 ) >> 3, -tc, tc );
        s[1] = xd_clip_u8( p0 + delta );
        s[2] = xd_clip_u8( q0 - delta );
    }
}

// ---------------------------------------------------------------------------------------------
// intra prediction modes, numbered like the reference's I_PRED_4x4_* (common/predict.h:44-59)
enum { PR_V = 0, PR_H, PR_DC, PR_DDL, PR_DDR, PR_VR, PR_HD, PR_VL, PR_HU, PR_DC_LEFT, PR_DC_TOP, PR_DC_128, PR_PLANE };

#define XS_F1( a, b ) ( ( ( a ) + ( b ) + 1 ) >> 1 )
#define XS_F2( a, b, c ) ( ( ( a ) + 2 * ( b ) + ( c ) + 2 ) >> 2 )

// one predicted pixel of a 4x4 block (common/predict.c:330-470).  e[0..3] = l3..l0, e[4] = lt,
// e[5..12] = t0..t7.
__device__ __forceinline__ int xd_pred4x4_px( int mode, int x, int y, const int e[13] )
{
    const int *t = e + 5;
#define L( k ) e[3 - ( k )]
    switch( mode )
    {
    case PR_V: return t[x];
    case PR_H: return L( y );
    case PR_DC: return ( L( 0 ) + L( 1 ) + L( 2 ) + L( 3 ) + t[0] + t[1] + t[2] + t[3] + 4 ) >> 3;
    case PR_DC_LEFT: return ( L( 0 ) + L( 1 ) + L( 2 ) + L( 3 ) + 2 ) >> 2;          // predict.c:334-343
    case PR_DC_TOP: return ( t[0] + t[1] + t[2] + t[3] + 2 ) >> 2;
    case PR_DC_128: return 128;
    case PR_DDL:
        return ( x == 3 && y == 3 ) ? XS_F2( t[6], t[7], t[7] ) : XS_F2( t[x + y], t[x + y + 1], t[x + y + 2] );
    case PR_DDR:
    {
        const int i = 4 + x - y;
        return XS_F2( e[i - 1], e[i], e[i + 1] );
    }
    case PR_VR:
    {
        const int z = 2 * x - y, i = 4 + x - ( y >> 1 );
        if( z >= 0 )
            return ( z & 1 ) ? XS_F2( e[i - 1], e[i], e[i + 1] ) : XS_F1( e[i], e[i + 1] );
        if( z == -1 )
            return XS_F2( e[3], e[4], e[5] );
        return XS_F2( e[4 - y], e[5 - y], e[6 - y] );
    }
    case PR_HD:
    {
        const int z = 2 * y - x, k = y - ( x >> 1 );
        if( z >= -1 )
            return ( z & 1 ) ? XS_F2( e[5 - k], e[4 - k], e[3 - k] ) : XS_F1( e[4 - k], e[3 - k] );
        return XS_F2( e[4 + x], e[3 + x], e[2 + x] );
    }
    case PR_VL:
    {
        const int i = x + ( y >> 1 );
        return ( y & 1 ) ? XS_F2( t[i], t[i + 1], t[i + 2] ) : XS_F1( t[i], t[i + 1] );
    }
    default: // PR_HU
    {
        const int z = x + 2 * y, k = y + ( x >> 1 );
        if( z > 5 )
            return L( 3 );
        if( z == 5 )
            return XS_F2( L( 2 ), L( 3 ), L( 3 ) );
        return ( z & 1 ) ? XS_F2( L( k ), L( k + 1 ), L( k + 2 ) ) : XS_F1( L( k ), L( k + 1 ) );
    }
    }
#undef L
}

