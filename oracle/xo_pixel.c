/* xo_pixel.c -- oracle: block cost metrics.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/pixel.c:44-70 (SAD), 76-102 (SSD), 185-231 (var, var2), 243-337 (SATD),
 * 489-505 (intra x3), common/predict.c:224-288 (8x8 chroma-style predictors).
 *
 * SATD here is a straight 32-bit 4x4 Hadamard.  The reference packs two 16-bit lanes in one
 * 32-bit word and halves once per 4x4 (satd_4x4) or once per 8x4 (satd_8x4); for 8-bit input
 * the lanes never overflow, so  sum|H d H^T| >> 1  per base block is the same integer.  The
 * two base blocks differ in WHERE the halving happens, which is kept: 8-wide sizes halve the
 * sum of two 4x4 transforms, 4-wide sizes halve each 4x4 on its own.
 */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

static const uint8_t blk_w[8] = { 16, 16, 8, 8, 8, 4, 4, 4 };
static const uint8_t blk_h[8] = { 16, 8, 16, 8, 4, 8, 4, 16 };

int xo_block_w( int size ) { return blk_w[size]; }
int xo_block_h( int size ) { return blk_h[size]; }

static __thread int64_t work[4];

void xo_work_counters( int64_t counts[4], int reset )
{
    if( counts )
        memcpy( counts, work, sizeof(work) );
    if( reset )
        memset( work, 0, sizeof(work) );
}

int xo_sad( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb )
{
    int w = blk_w[size], h = blk_h[size], x, y, acc = 0;
    for( y = 0; y < h; y++, a += sa, b += sb )
        for( x = 0; x < w; x++ )
            acc += abs( a[x] - b[x] );
    work[0] += w * h;
    work[2]++;
    return acc;
}

int xo_ssd( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb )
{
    int w = blk_w[size], h = blk_h[size], x, y, acc = 0;
    for( y = 0; y < h; y++, a += sa, b += sb )
        for( x = 0; x < w; x++ )
        {
            int d = a[x] - b[x];
            acc += d * d;
        }
    return acc;
}

/* sum of |H4 * D * H4^T| over one 4x4 difference block, NOT yet halved */
static int hadamard4x4_abs( const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb )
{
    int d[4][4], t[4][4], i, j, acc = 0;
    for( i = 0; i < 4; i++ )
        for( j = 0; j < 4; j++ )
            d[i][j] = a[i*sa + j] - b[i*sb + j];
    for( i = 0; i < 4; i++ )          /* rows */
    {
        int s01 = d[i][0] + d[i][1], d01 = d[i][0] - d[i][1];
        int s23 = d[i][2] + d[i][3], d23 = d[i][2] - d[i][3];
        t[i][0] = s01 + s23; t[i][1] = s01 - s23; t[i][2] = d01 + d23; t[i][3] = d01 - d23;
    }
    for( j = 0; j < 4; j++ )          /* columns */
    {
        int s01 = t[0][j] + t[1][j], d01 = t[0][j] - t[1][j];
        int s23 = t[2][j] + t[3][j], d23 = t[2][j] - t[3][j];
        acc += abs( s01 + s23 ) + abs( s01 - s23 ) + abs( d01 + d23 ) + abs( d01 - d23 );
    }
    return acc;
}

int xo_satd( int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb )
{
    int w = blk_w[size], h = blk_h[size], x, y, acc = 0;
    if( w == 4 )
    {
        /* pixel.c:267-291, 336-337: satd_4x4 halves its own sum; 4x8 / 4x16 add halved parts */
        for( y = 0; y < h; y += 4 )
            acc += hadamard4x4_abs( a + y*sa, sa, b + y*sb, sb ) >> 1;
    }
    else
    {
        /* pixel.c:294-335: satd_8x4 halves the sum of two side-by-side 4x4 transforms */
        for( y = 0; y < h; y += 4 )
            for( x = 0; x < w; x += 8 )
                acc += ( hadamard4x4_abs( a + y*sa + x, sa, b + y*sb + x, sb )
                       + hadamard4x4_abs( a + y*sa + x + 4, sa, b + y*sb + x + 4, sb ) ) >> 1;
    }
    work[1] += w * h;
    work[3]++;
    return acc;
}

int xo_cmp( int cmp, int size, const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb )
{
    switch( cmp )
    {
        case X264DSP_CMP_SAD:  return xo_sad( size, a, sa, b, sb );
        case X264DSP_CMP_SSD:  return xo_ssd( size, a, sa, b, sb );
        default:               return xo_satd( size, a, sa, b, sb );
    }
}

void xo_cost_batch( int cmp, int n, const pixel_t *pix1, const int64_t *off1, int stride1,
                    const pixel_t *pix2, const int64_t *off2, int stride2,
                    const uint8_t *size, int32_t *out )
{
    int i;
    for( i = 0; i < n; i++ )
        out[i] = xo_cmp( cmp, size[i], pix1 + off1[i], stride1, pix2 + off2[i], stride2 );
}

/* pixel.c:185-203: sum in the low 32 bits, sum of squares in the high 32 bits */
uint64_t xo_var( int size, const pixel_t *p, intptr_t stride )
{
    int w = blk_w[size], h = blk_h[size], x, y;
    uint32_t s = 0, q = 0;
    for( y = 0; y < h; y++, p += stride )
        for( x = 0; x < w; x++ )
        {
            s += p[x];
            q += p[x] * p[x];
        }
    return s + ((uint64_t)q << 32);
}

/* pixel.c:209-231 */
int xo_var2_8x8( const pixel_t *a, intptr_t sa, const pixel_t *b, intptr_t sb, int *ssd )
{
    int x, y, s = 0;
    uint32_t q = 0;
    for( y = 0; y < 8; y++, a += sa, b += sb )
        for( x = 0; x < 8; x++ )
        {
            int d = a[x] - b[x];
            s += d;
            q += d * d;
        }
    s = abs( s );
    *ssd = (int)q;
    return (int)( q - (uint32_t)( ((uint64_t)s * s) >> 6 ) );
}

/* predict.c:224-288.  src points at the block's top-left inside an FDEC-stride buffer whose row
 * above (8 samples) and column to the left (8 samples) hold the neighbours. */
void xo_predict_8x8c( int mode, pixel_t *src )
{
    const int S = XO_FDEC_STRIDE;
    int x, y;
    if( mode == 1 )            /* H */
    {
        for( y = 0; y < 8; y++ )
            memset( src + y*S, src[y*S - 1], 8 );
    }
    else if( mode == 2 )       /* V */
    {
        for( y = 0; y < 8; y++ )
            memcpy( src + y*S, src - S, 8 );
    }
    else                       /* DC: four 4x4 quadrants with their own means */
    {
        int top_l = 0, top_r = 0, left_u = 0, left_d = 0, dc[2][2];
        for( x = 0; x < 4; x++ )
        {
            top_l += src[x - S];
            top_r += src[x + 4 - S];
            left_u += src[x*S - 1];
            left_d += src[(x + 4)*S - 1];
        }
        dc[0][0] = (top_l + left_u + 4) >> 3;
        dc[0][1] = (top_r + 2) >> 2;
        dc[1][0] = (left_d + 2) >> 2;
        dc[1][1] = (top_r + left_d + 4) >> 3;
        for( y = 0; y < 8; y++ )
            for( x = 0; x < 8; x++ )
                src[y*S + x] = (pixel_t)dc[y >> 2][x >> 2];
    }
}

/* pixel.c:489-505: predict into fdec, then cost(fdec, fenc); order DC, H, V */
void xo_intra_x3_8x8c( int use_satd, const pixel_t *fenc, pixel_t *fdec, int res[3] )
{
    int m;
    for( m = 0; m < 3; m++ )
    {
        xo_predict_8x8c( m, fdec );
        res[m] = use_satd ? xo_satd( X264DSP_PIXEL_8x8, fdec, XO_FDEC_STRIDE, fenc, XO_FENC_STRIDE )
                          : xo_sad( X264DSP_PIXEL_8x8, fdec, XO_FDEC_STRIDE, fenc, XO_FENC_STRIDE );
    }
}
