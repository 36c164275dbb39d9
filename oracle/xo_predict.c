/* xo_predict.c -- oracle: the twelve 4x4 intra predictors and the residual coding of I4x4 macroblocks.
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/predict.c:320-470 (x264_predict_4x4_*_c), encoder/macroblock.h:37-61 (x264_mb_encode_i4x4),
 * encoder/macroblock.c:355-377 (I_4x4 branch of x264_macroblock_encode, missing top-right samples).
 *
 * The predictors are restated as ONE per-pixel rule over the thirteen edge samples
 *   e[0..3] = l3 l2 l1 l0 (left column, bottom to top), e[4] = lt (corner), e[5..12] = t0 .. t7 (row above)
 * so that the diagonal modes read as filters along that line (the reference spells out every pixel). */
#include <string.h>
#include "xo.h"

#define FENC XO_FENC_STRIDE
#define FDEC XO_FDEC_STRIDE
#define F1( a, b ) ( ( (a) + (b) + 1 ) >> 1 )
#define F2( a, b, c ) ( ( (a) + 2*(b) + (c) + 2 ) >> 2 )

/* mode numbering = the reference's I_PRED_4x4_* (predict.h:44-59) */
static int pred4x4_px( int mode, int x, int y, const int e[13] )
{
    const int *t = e + 5;
#define L( k ) e[3 - (k)]
    switch( mode )
    {
    case 0:  return t[x];                                                        /* V */
    case 1:  return L( y );                                                      /* H */
    case 2:  return ( L(0) + L(1) + L(2) + L(3) + t[0] + t[1] + t[2] + t[3] + 4 ) >> 3;     /* DC */
    case 9:  return ( L(0) + L(1) + L(2) + L(3) + 2 ) >> 2;                      /* DC_LEFT */
    case 10: return ( t[0] + t[1] + t[2] + t[3] + 2 ) >> 2;                      /* DC_TOP */
    case 11: return 128;                                                         /* DC_128 */
    case 3:                                                                      /* DDL: along the row above */
        return ( x == 3 && y == 3 ) ? F2( t[6], t[7], t[7] ) : F2( t[x+y], t[x+y+1], t[x+y+2] );
    case 4:                                                                      /* DDR: along left-corner-top */
    {
        int i = 4 + x - y;
        return F2( e[i-1], e[i], e[i+1] );
    }
    case 5:                                                                      /* VR */
    {
        int z = 2*x - y, i = 4 + x - (y >> 1);
        if( z >= 0 )
            return (z & 1) ? F2( e[i-1], e[i], e[i+1] ) : F1( e[i], e[i+1] );
        if( z == -1 )
            return F2( e[3], e[4], e[5] );
        return F2( e[4-y], e[5-y], e[6-y] );
    }
    case 6:                                                                      /* HD */
    {
        int z = 2*y - x, k = y - (x >> 1);
        if( z >= -1 )
            return (z & 1) ? F2( e[5-k], e[4-k], e[3-k] ) : F1( e[4-k], e[3-k] );
        return F2( e[4+x], e[3+x], e[2+x] );
    }
    case 7:                                                                      /* VL */
    {
        int i = x + (y >> 1);
        return (y & 1) ? F2( t[i], t[i+1], t[i+2] ) : F1( t[i], t[i+1] );
    }
    default:                                                                     /* 8: HU */
    {
        int z = x + 2*y, k = y + (x >> 1);
        if( z > 5 )
            return L( 3 );
        if( z == 5 )
            return F2( L(2), L(3), L(3) );
        return (z & 1) ? F2( L(k), L(k+1), L(k+2) ) : F1( L(k), L(k+1) );
    }
    }
#undef L
}

/* predict the 4x4 block at src (FDEC stride) in place from its neighbours */
void xo_predict_4x4( int mode, pixel_t *src )
{
    int e[13], k, x, y;
    for( k = 0; k < 4; k++ )
        e[3-k] = src[k*FDEC - 1];
    e[4] = src[-FDEC - 1];
    for( k = 0; k < 8; k++ )
        e[5+k] = src[-FDEC + k];
    for( y = 0; y < 4; y++ )
        for( x = 0; x < 4; x++ )
            src[y*FDEC + x] = (pixel_t)pred4x4_px( mode, x, y, e );
}

/* The luma part of an I4x4 macroblock (macroblock.c:355-377 + macroblock.h:37-61), I slice.
 * fdec points at the macroblock origin inside an FDEC-stride buffer that holds the reconstructed neighbours: row -1
 * from column -1 to column 19, column -1 of rows 0..15 (the reference's fdec_buf).  modes[16] are I_PRED_4x4_* values
 * in coding order; replicate5 says block 5 has a row above but no top-right macroblock (blocks 3, 7, 11, 13, 15
 * always replicate: their top-right block is coded later).  Returns i_cbp_luma. */
int xo_encode_luma_i4x4( const pixel_t *fenc, pixel_t *fdec, int qp, const uint8_t modes[16], int replicate5,
                         int16_t *levels /*[16][16]*/, uint8_t *nnz /*[16]*/ )
{
    uint16_t mf[16], bias[16];
    int dequant[6][16];
    int idx, cbp = 0;
    xo_quant_tables( 0, qp, mf, bias );
    xo_dequant_table( dequant );
    for( idx = 0; idx < 16; idx++ )
    {
        const int x = ((idx & 1) + ((idx >> 2) & 1) * 2) * 4, y = (((idx >> 1) & 1) + ((idx >> 3) & 1) * 2) * 4;
        pixel_t *dst = fdec + y*FDEC + x;
        coef_t dct[16];
        int nz;
        if( idx == 3 || idx == 7 || idx == 11 || idx == 13 || idx == 15 || ( idx == 5 && replicate5 ) )
            memset( dst + 4 - FDEC, dst[3 - FDEC], 4 );
        xo_predict_4x4( modes[idx], dst );
        xo_sub4x4_dct( dct, fenc + y*FENC + x, dst );
        nz = xo_quant_4x4( dct, mf, bias );
        nnz[idx] = (uint8_t)nz;
        xo_zigzag_4x4( levels + idx*16, dct );
        if( nz )
        {
            cbp |= 1 << (idx >> 2);
            xo_dequant_4x4( dct, dequant, qp );
            xo_add4x4_idct( dst, dct );
        }
    }
    return cbp;
}
