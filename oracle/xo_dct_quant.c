/* xo_dct_quant.c -- oracle: 4x4 integer transform, quantisation, decimation.
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/dct.c:36-100 (DC Hadamards), 115-195 (forward), 197-284 (inverse), 329-347 (zig-zag)
 * common/quant.c:29-101 (quant / dequant), 133-192 (chroma DC optimiser), 221-276 (decimate, last)
 *
 * All stores to coefficient arrays truncate to int16_t exactly where the reference's dctcoef
 * stores do (including the intermediate tmp[] arrays of the transforms).
 */
#include <string.h>
#include "xo.h"

#define FENC XO_FENC_STRIDE
#define FDEC XO_FDEC_STRIDE

/* forward core transform of one row/column: (a b c d) -> (a+d+b+c, 2(a-d)+(b-c), a+d-b-c, (a-d)-2(b-c)) */
static void fwd4( int a, int b, int c, int d, coef_t *o0, coef_t *o1, coef_t *o2, coef_t *o3 )
{
    int s_ad = a + d, s_bc = b + c, d_ad = a - d, d_bc = b - c;
    *o0 = (coef_t)( s_ad + s_bc );
    *o1 = (coef_t)( 2*d_ad + d_bc );
    *o2 = (coef_t)( s_ad - s_bc );
    *o3 = (coef_t)( d_ad - 2*d_bc );
}

/* dct.c:115-150 */
void xo_sub4x4_dct( coef_t dct[16], const pixel_t *fenc, const pixel_t *fdec )
{
    coef_t r[16], t[16];
    int i, j;
    for( i = 0; i < 4; i++ )
        for( j = 0; j < 4; j++ )
            r[4*i + j] = (coef_t)( fenc[i*FENC + j] - fdec[i*FDEC + j] );
    for( i = 0; i < 4; i++ )      /* transform each row, store transposed */
        fwd4( r[4*i], r[4*i+1], r[4*i+2], r[4*i+3], &t[i], &t[4+i], &t[8+i], &t[12+i] );
    for( i = 0; i < 4; i++ )
        fwd4( t[4*i], t[4*i+1], t[4*i+2], t[4*i+3], &dct[4*i], &dct[4*i+1], &dct[4*i+2], &dct[4*i+3] );
}

void xo_sub8x8_dct( coef_t dct[4][16], const pixel_t *fenc, const pixel_t *fdec )
{
    int k;
    for( k = 0; k < 4; k++ )
        xo_sub4x4_dct( dct[k], fenc + (k >> 1)*4*FENC + (k & 1)*4, fdec + (k >> 1)*4*FDEC + (k & 1)*4 );
}

void xo_sub16x16_dct( coef_t dct[16][16], const pixel_t *fenc, const pixel_t *fdec )
{
    int k;
    for( k = 0; k < 4; k++ )
        xo_sub8x8_dct( &dct[4*k], fenc + (k >> 1)*8*FENC + (k & 1)*8, fdec + (k >> 1)*8*FDEC + (k & 1)*8 );
}

/* dct.c:168-195 */
void xo_sub8x8_dct_dc( coef_t dc[4], const pixel_t *fenc, const pixel_t *fdec )
{
    int k, x, y, s[4];
    for( k = 0; k < 4; k++ )
    {
        const pixel_t *a = fenc + (k >> 1)*4*FENC + (k & 1)*4, *b = fdec + (k >> 1)*4*FDEC + (k & 1)*4;
        int acc = 0;
        for( y = 0; y < 4; y++ )
            for( x = 0; x < 4; x++ )
                acc += a[y*FENC + x] - b[y*FDEC + x];
        s[k] = (coef_t)acc;
    }
    dc[0] = (coef_t)( s[0] + s[1] + s[2] + s[3] );
    dc[1] = (coef_t)( s[0] + s[1] - s[2] - s[3] );
    dc[2] = (coef_t)( s[0] - s[1] + s[2] - s[3] );
    dc[3] = (coef_t)( s[0] - s[1] - s[2] + s[3] );
}

/* inverse core transform of one line */
static void inv4( int a, int b, int c, int d, int *o0, int *o1, int *o2, int *o3 )
{
    int e = a + c, f = a - c, gg = b + (d >> 1), hh = (b >> 1) - d;
    *o0 = e + gg; *o1 = f + hh; *o2 = f - hh; *o3 = e - gg;
}

/* dct.c:197-235 */
void xo_add4x4_idct( pixel_t *fdec, const coef_t dct[16] )
{
    coef_t t[16], r[16];
    int i, x, y, o0, o1, o2, o3;
    for( i = 0; i < 4; i++ )
    {
        inv4( dct[i], dct[4+i], dct[8+i], dct[12+i], &o0, &o1, &o2, &o3 );
        t[4*i] = (coef_t)o0; t[4*i+1] = (coef_t)o1; t[4*i+2] = (coef_t)o2; t[4*i+3] = (coef_t)o3;
    }
    for( i = 0; i < 4; i++ )
    {
        inv4( t[i], t[4+i], t[8+i], t[12+i], &o0, &o1, &o2, &o3 );
        r[i]    = (coef_t)( (o0 + 32) >> 6 );
        r[4+i]  = (coef_t)( (o1 + 32) >> 6 );
        r[8+i]  = (coef_t)( (o2 + 32) >> 6 );
        r[12+i] = (coef_t)( (o3 + 32) >> 6 );
    }
    for( y = 0; y < 4; y++ )
        for( x = 0; x < 4; x++ )
        {
            int v = fdec[y*FDEC + x] + r[4*y + x];
            fdec[y*FDEC + x] = v < 0 ? 0 : v > 255 ? 255 : (pixel_t)v;
        }
}

void xo_add8x8_idct( pixel_t *fdec, coef_t dct[4][16] )
{
    int k;
    for( k = 0; k < 4; k++ )
        xo_add4x4_idct( fdec + (k >> 1)*4*FDEC + (k & 1)*4, dct[k] );
}

void xo_add16x16_idct( pixel_t *fdec, coef_t dct[16][16] )
{
    int k;
    for( k = 0; k < 4; k++ )
        xo_add8x8_idct( fdec + (k >> 1)*8*FDEC + (k & 1)*8, &dct[4*k] );
}

/* dct.c:253-264 */
static void add_dc_4x4( pixel_t *fdec, int dc )
{
    int x, y;
    dc = (coef_t)( (dc + 32) >> 6 );
    for( y = 0; y < 4; y++ )
        for( x = 0; x < 4; x++ )
        {
            int v = fdec[y*FDEC + x] + dc;
            fdec[y*FDEC + x] = v < 0 ? 0 : v > 255 ? 255 : (pixel_t)v;
        }
}

void xo_add8x8_idct_dc( pixel_t *fdec, const coef_t dc[4] )
{
    int k;
    for( k = 0; k < 4; k++ )
        add_dc_4x4( fdec + (k >> 1)*4*FDEC + (k & 1)*4, dc[k] );
}

/* dct.c:274-284: raster order of 4x4 blocks, NOT coding order */
void xo_add16x16_idct_dc( pixel_t *fdec, const coef_t dc[16] )
{
    int k;
    for( k = 0; k < 16; k++ )
        add_dc_4x4( fdec + (k >> 2)*4*FDEC + (k & 3)*4, dc[k] );
}

/* 4-point Hadamard butterfly in the reference's output order (dct.c:45-53) */
static void had4( int a, int b, int c, int d, int *o0, int *o1, int *o2, int *o3 )
{
    int s01 = a + b, d01 = a - b, s23 = c + d, d23 = c - d;
    *o0 = s01 + s23; *o1 = s01 - s23; *o2 = d01 - d23; *o3 = d01 + d23;
}

static void hadamard_dc( coef_t d[16], int round_half )
{
    coef_t t[16];
    int i, o0, o1, o2, o3;
    for( i = 0; i < 4; i++ )
    {
        had4( d[4*i], d[4*i+1], d[4*i+2], d[4*i+3], &o0, &o1, &o2, &o3 );
        t[i] = (coef_t)o0; t[4+i] = (coef_t)o1; t[8+i] = (coef_t)o2; t[12+i] = (coef_t)o3;
    }
    for( i = 0; i < 4; i++ )
    {
        had4( t[4*i], t[4*i+1], t[4*i+2], t[4*i+3], &o0, &o1, &o2, &o3 );
        if( round_half )
        {
            o0 = (o0 + 1) >> 1; o1 = (o1 + 1) >> 1; o2 = (o2 + 1) >> 1; o3 = (o3 + 1) >> 1;
        }
        d[4*i] = (coef_t)o0; d[4*i+1] = (coef_t)o1; d[4*i+2] = (coef_t)o2; d[4*i+3] = (coef_t)o3;
    }
}

void xo_dct4x4dc( coef_t d[16] )  { hadamard_dc( d, 1 ); }     /* dct.c:36-68 */
void xo_idct4x4dc( coef_t d[16] ) { hadamard_dc( d, 0 ); }     /* dct.c:70-100 */

/* dct.c:329-347 */
void xo_zigzag_4x4( coef_t level[16], const coef_t dct[16] )
{
    static const uint8_t order[16] = { 0, 4, 1, 2, 5, 8, 12, 9, 6, 3, 7, 10, 13, 14, 11, 15 };
    int i;
    for( i = 0; i < 16; i++ )
        level[i] = dct[order[i]];
}

/* quant.c:29-36: sign * ((bias + |c|) * mf >> 16) with int arithmetic */
static int quant_one( coef_t *c, int mf, int bias )
{
    int v = *c;
    if( v > 0 )
        v = (bias + v) * mf >> 16;
    else
        v = -( (bias - v) * mf >> 16 );
    *c = (coef_t)v;
    return *c;
}

int xo_quant_4x4( coef_t dct[16], const uint16_t mf[16], const uint16_t bias[16] )
{
    int i, nz = 0;
    for( i = 0; i < 16; i++ )
        nz |= quant_one( &dct[i], mf[i], bias[i] );
    return nz != 0;
}

int xo_quant_4x4_dc( coef_t dct[16], int mf, int bias )
{
    int i, nz = 0;
    for( i = 0; i < 16; i++ )
        nz |= quant_one( &dct[i], mf, bias );
    return nz != 0;
}

int xo_quant_2x2_dc( coef_t dct[4], int mf, int bias )
{
    int i, nz = 0;
    for( i = 0; i < 4; i++ )
        nz |= quant_one( &dct[i], mf, bias );
    return nz != 0;
}

/* quant.c:64-81 */
void xo_dequant_4x4( coef_t dct[16], int dequant_mf[6][16], int qp )
{
    const int *mf = dequant_mf[qp % 6];
    int bits = qp / 6 - 4, i;
    if( bits >= 0 )
        for( i = 0; i < 16; i++ )
            dct[i] = (coef_t)( ( dct[i] * mf[i] ) << bits );
    else
        for( i = 0; i < 16; i++ )
            dct[i] = (coef_t)( ( dct[i] * mf[i] + (1 << (-bits - 1)) ) >> -bits );
}

/* quant.c:83-101 */
void xo_dequant_4x4_dc( coef_t dct[16], int dequant_mf[6][16], int qp )
{
    int bits = qp / 6 - 6, i;
    if( bits >= 0 )
    {
        int dmf = dequant_mf[qp % 6][0] << bits;
        for( i = 0; i < 16; i++ )
            dct[i] = (coef_t)( dct[i] * dmf );
    }
    else
    {
        int dmf = dequant_mf[qp % 6][0], f = 1 << (-bits - 1);
        for( i = 0; i < 16; i++ )
            dct[i] = (coef_t)( ( dct[i] * dmf + f ) >> -bits );
    }
}

/* quant.c:133-143: 2x2 inverse DC transform + dequant, results kept as int16 and biased by 32 */
static void chroma_dc_recon( coef_t out[4], const coef_t dct[4], int dmf )
{
    int a = dct[0] + dct[1], b = dct[2] + dct[3], c = dct[0] - dct[1], d = dct[2] - dct[3];
    out[0] = (coef_t)( ((a + b) * dmf >> 5) + 32 );
    out[1] = (coef_t)( ((a - b) * dmf >> 5) + 32 );
    out[2] = (coef_t)( ((c + d) * dmf >> 5) + 32 );
    out[3] = (coef_t)( ((c - d) * dmf >> 5) + 32 );
}

/* quant.c:157-192: greedily pull each level towards zero while the reconstructed DCs
 * (after the >>6 of the idct) stay what they were */
int xo_optimize_chroma_2x2_dc( coef_t dct[4], int dmf )
{
    coef_t want[4], got[4];
    int any = 0, i, k, nz = 0;
    chroma_dc_recon( want, dct, dmf );
    for( i = 0; i < 4; i++ )
        any |= want[i];
    if( !(any >> 6) )
        return 0;
    for( k = 3; k >= 0; k-- )
    {
        int level = dct[k];
        int step = level < 0 ? -1 : 1;
        while( level )
        {
            int diff = 0;
            dct[k] = (coef_t)( level - step );
            chroma_dc_recon( got, dct, dmf );
            for( i = 0; i < 4; i++ )
                diff |= want[i] ^ got[i];
            if( diff >> 6 )
            {
                nz = 1;
                dct[k] = (coef_t)level;
                break;
            }
            level -= step;
        }
    }
    return nz;
}

/* quant.c:221-261 */
static int decimate_score( const coef_t *level, int n )
{
    static const uint8_t run_score[16] = { 3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    int i = n - 1, score = 0;
    while( i >= 0 && level[i] == 0 )
        i--;
    while( i >= 0 )
    {
        int run = 0;
        if( level[i] > 1 || level[i] < -1 )
            return 9;
        i--;
        while( i >= 0 && level[i] == 0 )
        {
            i--;
            run++;
        }
        score += run_score[run];
    }
    return score;
}

int xo_decimate_score15( const coef_t *level ) { return decimate_score( level + 1, 15 ); }
int xo_decimate_score16( const coef_t *level ) { return decimate_score( level, 16 ); }

/* quant.c:263-276 */
int xo_coeff_last( const coef_t *level, int n )
{
    int i = n - 1;
    while( i >= 0 && level[i] == 0 )
        i--;
    return i;
}
