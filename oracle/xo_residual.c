/* xo_residual.c -- oracle: residual coding of inter macroblocks (P slices) and I16x16 macroblocks (I slices).
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * encoder/macroblock.c:9-59 (2x2 chroma DC helpers), 175-305 (x264_mb_encode_chroma),
 * 379-471 (inter branch of x264_macroblock_encode + cbp packing), with
 * h->mb.b_dct_decimate = 1 (P slices, common/macroblock.c:238-239), no noise reduction, CABAC.
 *
 * Output convention (include/x264dsp_b200.h): levels are the zig-zagged quantised coefficients of
 * every 4x4 (all zero when the block quantised to zero); chroma DC levels are reported only when
 * their final nnz flag is set.  Decimation changes nnz / cbp / reconstruction, never the levels,
 * exactly as in the reference (macroblock.c:69-70).
 */
#include <string.h>
#include "xo.h"

#define FENC XO_FENC_STRIDE
#define FDEC XO_FDEC_STRIDE

/* byte offset of 4x4 block idx (coding order) inside a 16x16 at the given stride */
static int blk_off( int idx, int stride )
{
    int x = ((idx & 1) + ((idx >> 2) & 1) * 2) * 4;
    int y = (((idx >> 1) & 1) + ((idx >> 3) & 1) * 2) * 4;
    return y * stride + x;
}

typedef struct
{
    int16_t *luma;        /* [16][16] */
    int16_t *chroma_dc;   /* [2][4]   */
    int16_t *chroma_ac;   /* [2][4][16] */
    uint8_t *nnz;         /* [27]: 16 luma, 4 U, 4 V, luma DC, U DC, V DC */
} mb_out_t;

/* (dequant_mf[qp%6][0] << qp/6) >> 5 scaling of the 2x2 inverse DC transform (macroblock.c:17-43) */
static void chroma_dc_inverse( coef_t out[4], const coef_t dc[4], int dmf )
{
    int a = dc[0] + dc[1], b = dc[2] + dc[3], c = dc[0] - dc[1], d = dc[2] - dc[3];
    out[0] = (coef_t)( (a + b) * dmf );
    out[1] = (coef_t)( (a - b) * dmf );
    out[2] = (coef_t)( (c + d) * dmf );
    out[3] = (coef_t)( (c - d) * dmf );
}

static void store_chroma_dc_levels( int16_t *dst, const coef_t dc[4] )
{
    dst[0] = dc[0]; dst[1] = dc[2]; dst[2] = dc[1]; dst[3] = dc[3];     /* macroblock.c:9-15 */
}

/* x264_mb_encode_chroma (macroblock.c:175-305); returns i_cbp_chroma.  b_inter selects the quant tables
 * (CQM_4IC + b_inter) and stands for h->mb.b_dct_decimate as well: inter macroblocks live in P slices
 * (decimation and the variance early-out on), intra macroblocks in I slices (both off). */
static int encode_chroma( const pixel_t *fenc_u, const pixel_t *fenc_v, pixel_t *fdec_u, pixel_t *fdec_v,
                          int qpc, mb_out_t *o, int b_inter )
{
    const int b_decimate = b_inter;
    uint16_t mf[16], bias[16];
    int dequant[6][16];
    const pixel_t *src[2] = { fenc_u, fenc_v };
    pixel_t *dst[2] = { fdec_u, fdec_v };
    int cbp_chroma = 0, ch, i;
    const int dmf_full = 0;
    (void)dmf_full;

    xo_quant_tables( b_inter, qpc, mf, bias );
    xo_dequant_table( dequant );
    {
        const int dmf = dequant[qpc % 6][0] << (qpc / 6);
        const int dc_mf = mf[0] >> 1, dc_bias = bias[0] << 1;

        if( b_decimate && qpc >= 18 )                         /* macroblock.c:188-232 */
        {
            int thresh = (xo_lambda2( qpc ) + 32) >> 6;
            int ssd[2] = { 0, 0 };
            int score = xo_var2_8x8( fenc_u, FENC, fdec_u, FDEC, &ssd[0] );
            if( score < (thresh << 2) )
                score += xo_var2_8x8( fenc_v, FENC, fdec_v, FDEC, &ssd[1] );
            if( score < (thresh << 2) )
            {
                for( ch = 0; ch < 2; ch++ )
                {
                    coef_t dc[4], rec[4];
                    if( ssd[ch] <= thresh )
                        continue;
                    xo_sub8x8_dct_dc( dc, src[ch], dst[ch] );
                    if( !xo_quant_2x2_dc( dc, dc_mf, dc_bias ) )
                        continue;
                    if( qpc <= 22 && !xo_optimize_chroma_2x2_dc( dc, dmf ) )
                        continue;
                    o->nnz[25 + ch] = 1;
                    store_chroma_dc_levels( o->chroma_dc + 4*ch, dc );
                    chroma_dc_inverse( rec, dc, dmf >> 5 );
                    xo_add8x8_idct_dc( dst[ch], rec );
                    cbp_chroma = 1;
                }
                return cbp_chroma;
            }
        }

        for( ch = 0; ch < 2; ch++ )                           /* macroblock.c:234-301 */
        {
            coef_t dct[4][16], dc[4];
            int score = 0, nz_ac = 0, nz_dc;
            xo_sub8x8_dct( dct, src[ch], dst[ch] );
            {
                /* dct2x2dc, macroblock.c:45-59 */
                int a = dct[0][0] + dct[1][0], b = dct[2][0] + dct[3][0];
                int c = dct[0][0] - dct[1][0], d = dct[2][0] - dct[3][0];
                dc[0] = (coef_t)( a + b ); dc[2] = (coef_t)( c + d );
                dc[1] = (coef_t)( a - b ); dc[3] = (coef_t)( c - d );
                dct[0][0] = dct[1][0] = dct[2][0] = dct[3][0] = 0;
            }
            for( i = 0; i < 4; i++ )
            {
                int nz = xo_quant_4x4( dct[i], mf, bias );
                int16_t *lv = o->chroma_ac + (ch*4 + i) * 16;
                o->nnz[16 + ch*4 + i] = (uint8_t)nz;
                xo_zigzag_4x4( lv, dct[i] );
                if( nz )
                {
                    nz_ac = 1;
                    xo_dequant_4x4( dct[i], dequant, qpc );
                    if( b_decimate )
                        score += xo_decimate_score15( lv );
                }
            }
            nz_dc = xo_quant_2x2_dc( dc, dc_mf, dc_bias );
            o->nnz[25 + ch] = (uint8_t)nz_dc;

            if( ( b_decimate && score < 7 ) || !nz_ac )
            {
                coef_t rec[4];
                memset( o->nnz + 16 + ch*4, 0, 4 );
                if( !nz_dc )
                    continue;
                if( qpc <= 22 && !xo_optimize_chroma_2x2_dc( dc, dmf ) )
                {
                    o->nnz[25 + ch] = 0;
                    continue;
                }
                store_chroma_dc_levels( o->chroma_dc + 4*ch, dc );
                chroma_dc_inverse( rec, dc, dmf >> 5 );
                xo_add8x8_idct_dc( dst[ch], rec );
            }
            else
            {
                cbp_chroma = 1;
                if( nz_dc )
                {
                    coef_t rec[4];
                    store_chroma_dc_levels( o->chroma_dc + 4*ch, dc );
                    chroma_dc_inverse( rec, dc, dmf >> 5 );
                    for( i = 0; i < 4; i++ )
                        dct[i][0] = rec[i];
                }
                xo_add8x8_idct( dst[ch], dct );
            }
        }
        /* macroblock.c:303-304: 0 none, 1 DC only, 2 DC+AC */
        cbp_chroma += o->nnz[25] | o->nnz[26] | cbp_chroma;
    }
    return cbp_chroma;
}

/* inter branch of x264_macroblock_encode (macroblock.c:379-454); returns i_cbp_luma */
static int encode_luma_inter( const pixel_t *fenc, pixel_t *fdec, int qp, mb_out_t *o )
{
    uint16_t mf[16], bias[16];
    int dequant[6][16];
    coef_t dct[16][16];
    int cbp = 0, mb_score = 0, i8, i4;
    xo_quant_tables( 1, qp, mf, bias );
    xo_dequant_table( dequant );
    xo_sub16x16_dct( dct, fenc, fdec );
    for( i8 = 0; i8 < 4; i8++ )
    {
        int score8 = 0;
        for( i4 = 0; i4 < 4; i4++ )
        {
            int idx = i8*4 + i4;
            int nz = xo_quant_4x4( dct[idx], mf, bias );
            o->nnz[idx] = (uint8_t)nz;
            xo_zigzag_4x4( o->luma + idx*16, dct[idx] );
            if( nz )
            {
                xo_dequant_4x4( dct[idx], dequant, qp );
                if( score8 < 6 )
                    score8 += xo_decimate_score16( o->luma + idx*16 );
            }
        }
        mb_score += score8;
        if( score8 < 4 )
            memset( o->nnz + i8*4, 0, 4 );
        else
            cbp |= 1 << i8;
    }
    if( mb_score < 6 )
    {
        cbp = 0;
        memset( o->nnz, 0, 16 );
    }
    else
        for( i8 = 0; i8 < 4; i8++ )
            if( cbp & (1 << i8) )
                xo_add8x8_idct( fdec + ((i8 & 1) + (i8 >> 1) * FDEC) * 8, &dct[i8*4] );
    return cbp;
}

/* one macroblock on FENC/FDEC-strided buffers, the reference's fenc_buf / fdec_buf shapes
 * (common/macroblock.c:260-265): luma 16x16; chroma U at +0 and V at +8 (fenc) / +16 (fdec) */
static int encode_inter_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                            int qp, mb_out_t *o )
{
    int cbp_luma, cbp_chroma;
    memset( o->luma, 0, 16*16*sizeof(int16_t) );
    memset( o->chroma_dc, 0, 8*sizeof(int16_t) );
    memset( o->chroma_ac, 0, 8*16*sizeof(int16_t) );
    memset( o->nnz, 0, X264DSP_RES_NNZ_PER_MB );
    cbp_luma = encode_luma_inter( fenc_y, fdec_y, qp, o );
    cbp_chroma = encode_chroma( fenc_c, fenc_c + 8, fdec_c, fdec_c + 16, xo_chroma_qp( qp ), o, 1 );
    /* macroblock.c:465-471, CABAC */
    return (cbp_chroma << 4) | cbp_luma | (o->nnz[24] << 8) | (o->nnz[25] << 9) | (o->nnz[26] << 10);
}

/* x264_macroblock_probe_pskip (encoder/macroblock.c:492-604) after its motion compensation: fdec holds the P_SKIP
 * prediction (mc_luma / mc_chroma at the clipped pskip mv).  Returns 1 when the macroblock may be skipped: the luma
 * decimate scores of all coded 4x4s stay below 6 and each chroma plane passes its SSD / DC / AC-decimate ladder. */
static int probe_pskip_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, const pixel_t *fdec_y, const pixel_t *fdec_c, int qp )
{
    uint16_t mf[16], bias[16];
    int i8, i4, ch, score = 0;
    const int qpc = xo_chroma_qp( qp );
    xo_quant_tables( 1, qp, mf, bias );
    for( i8 = 0; i8 < 4; i8++ )                                               /* macroblock.c:510-534 */
    {
        coef_t dct[4][16], scan[16];
        xo_sub8x8_dct( dct, fenc_y + ((i8 & 1) + (i8 >> 1) * FENC) * 8, fdec_y + ((i8 & 1) + (i8 >> 1) * FDEC) * 8 );
        for( i4 = 0; i4 < 4; i4++ )
        {
            if( !xo_quant_4x4( dct[i4], mf, bias ) )
                continue;
            xo_zigzag_4x4( scan, dct[i4] );
            score += xo_decimate_score16( scan );
            if( score >= 6 )
                return 0;
        }
    }
    {
        const int thresh = ( xo_lambda2( qpc ) + 32 ) >> 6;                   /* macroblock.c:536-600 */
        xo_quant_tables( 1, qpc, mf, bias );
        for( ch = 0; ch < 2; ch++ )
        {
            const pixel_t *src = fenc_c + 8*ch, *dst = fdec_c + 16*ch;
            coef_t dc[4], dct[4][16], scan[16];
            int ssd = xo_ssd( 3 /* PIXEL_8x8 */, dst, FDEC, src, FENC );
            if( ssd < thresh )
                continue;
            xo_sub8x8_dct_dc( dc, src, dst );
            if( xo_quant_2x2_dc( dc, mf[0] >> 1, bias[0] << 1 ) )
                return 0;
            if( ssd < (thresh << 2) )
                continue;
            xo_sub8x8_dct( dct, src, dst );
            for( i4 = 0, score = 0; i4 < 4; i4++ )
            {
                dct[i4][0] = 0;
                if( !xo_quant_4x4( dct[i4], mf, bias ) )
                    continue;
                xo_zigzag_4x4( scan, dct[i4] );
                score += xo_decimate_score15( scan );
                if( score >= 7 )
                    return 0;
            }
        }
    }
    return 1;
}

int xo_probe_pskip_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, const pixel_t *fdec_y, const pixel_t *fdec_c, int qp )
{
    return probe_pskip_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp );
}

/* every macroblock of a frame: pred_slot holds the P_SKIP prediction of each macroblock (x264dsp_mc_frame_dev /
 * xo_mc_frame at the pskip MVs); skip[xy] = 1 when x264_macroblock_probe_pskip would return 1 */
void xo_probe_pskip_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *pred_slot, int qp, uint8_t *skip )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, x, y;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
        for( mb_x = 0; mb_x < g->mb_w; mb_x++ )
        {
            pixel_t fenc_y[16*FENC], fenc_c[8*FENC], fdec_y[16*FDEC], fdec_c[8*FDEC];
            const pixel_t *sy = fenc_slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            const pixel_t *sc = fenc_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            const pixel_t *py = pred_slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            const pixel_t *pc = pred_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            memset( fdec_y, 0, sizeof(fdec_y) );
            memset( fdec_c, 0, sizeof(fdec_c) );
            for( y = 0; y < 16; y++ )
            {
                memcpy( fenc_y + y*FENC, sy + (ptrdiff_t)y*ls, 16 );
                memcpy( fdec_y + y*FDEC, py + (ptrdiff_t)y*ls, 16 );
            }
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    fenc_c[y*FENC + x]      = sc[(ptrdiff_t)y*cs + 2*x];
                    fenc_c[y*FENC + 8 + x]  = sc[(ptrdiff_t)y*cs + 2*x + 1];
                    fdec_c[y*FDEC + x]      = pc[(ptrdiff_t)y*cs + 2*x];
                    fdec_c[y*FDEC + 16 + x] = pc[(ptrdiff_t)y*cs + 2*x + 1];
                }
            skip[mb_y * g->mb_w + mb_x] = (uint8_t)probe_pskip_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp );
        }
}

/* block_idx_xy_1d (common/macroblock.h): coding index of a luma 4x4 -> raster index x + 4 y */
static const uint8_t blk_raster[16] = { 0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15 };

/* x264_mb_encode_i16x16 (macroblock.c:72-162) with the prediction already in fdec and
 * h->mb.b_dct_decimate = 0 (I slice: decimate_score starts at 9, nothing is dropped); returns i_cbp_luma.
 * luma_dc[16] receives the zig-zagged levels of the DC block (h->dct.luma16x16_dc). */
static int encode_luma_i16x16( const pixel_t *fenc, pixel_t *fdec, int qp, mb_out_t *o, int16_t *luma_dc )
{
    uint16_t mf[16], bias[16];
    int dequant[6][16];
    coef_t dct[16][16], dc[16];
    int i, nz, block_cbp = 0;
    xo_quant_tables( 0, qp, mf, bias );
    xo_dequant_table( dequant );
    xo_sub16x16_dct( dct, fenc, fdec );
    for( i = 0; i < 16; i++ )
    {
        dc[blk_raster[i]] = dct[i][0];
        dct[i][0] = 0;
        nz = xo_quant_4x4( dct[i], mf, bias );
        o->nnz[i] = (uint8_t)nz;
        xo_zigzag_4x4( o->luma + i*16, dct[i] );
        if( nz )
        {
            xo_dequant_4x4( dct[i], dequant, qp );
            block_cbp = 0xf;
        }
    }
    xo_dct4x4dc( dc );
    nz = xo_quant_4x4_dc( dc, mf[0] >> 1, bias[0] << 1 );
    o->nnz[24] = (uint8_t)nz;
    memset( luma_dc, 0, 16*sizeof(int16_t) );
    if( nz )
    {
        xo_zigzag_4x4( luma_dc, dc );
        xo_idct4x4dc( dc );
        xo_dequant_4x4_dc( dc, dequant, qp );
        if( block_cbp )
            for( i = 0; i < 16; i++ )
                dct[i][0] = dc[blk_raster[i]];
    }
    if( block_cbp )
        xo_add16x16_idct( fdec, dct );
    else if( nz )
        xo_add16x16_idct_dc( fdec, dc );
    return block_cbp;
}

static int encode_intra16_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                              int qp, mb_out_t *o, int16_t *luma_dc )
{
    int cbp_luma, cbp_chroma;
    memset( o->luma, 0, 16*16*sizeof(int16_t) );
    memset( o->chroma_dc, 0, 8*sizeof(int16_t) );
    memset( o->chroma_ac, 0, 8*16*sizeof(int16_t) );
    memset( o->nnz, 0, X264DSP_RES_NNZ_PER_MB );
    cbp_luma = encode_luma_i16x16( fenc_y, fdec_y, qp, o, luma_dc );
    cbp_chroma = encode_chroma( fenc_c, fenc_c + 8, fdec_c, fdec_c + 16, xo_chroma_qp( qp ), o, 0 );
    return (cbp_chroma << 4) | cbp_luma | (o->nnz[24] << 8) | (o->nnz[25] << 9) | (o->nnz[26] << 10);
}

/* I4x4 macroblock of an I slice: luma through xo_encode_luma_i4x4 (fdec_y with its reconstructed neighbours, see
 * xo_predict.c), chroma as for I16x16 on the caller's chroma prediction */
static int encode_intra4_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                             int qp, const uint8_t modes[16], int replicate5, mb_out_t *o )
{
    int cbp_luma, cbp_chroma;
    memset( o->luma, 0, 16*16*sizeof(int16_t) );
    memset( o->chroma_dc, 0, 8*sizeof(int16_t) );
    memset( o->chroma_ac, 0, 8*16*sizeof(int16_t) );
    memset( o->nnz, 0, X264DSP_RES_NNZ_PER_MB );
    cbp_luma = xo_encode_luma_i4x4( fenc_y, fdec_y, qp, modes, replicate5, o->luma, o->nnz );
    cbp_chroma = encode_chroma( fenc_c, fenc_c + 8, fdec_c, fdec_c + 16, xo_chroma_qp( qp ), o, 0 );
    return (cbp_chroma << 4) | cbp_luma | (o->nnz[25] << 9) | (o->nnz[26] << 10);
}

int xo_encode_intra4_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                         int qp, const uint8_t modes[16], int replicate5, int16_t *levels, uint8_t *nnz )
{
    mb_out_t o = { levels, levels + 256, levels + 264, nnz };
    return encode_intra4_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp, modes, replicate5, &o );
}

int xo_encode_intra16_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                          int qp, int16_t *levels, int16_t *luma_dc, uint8_t *nnz )
{
    mb_out_t o = { levels, levels + 256, levels + 264, nnz };
    return encode_intra16_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp, &o, luma_dc );
}

void xo_residual_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                        int16_t *levels, uint8_t *nnz, int16_t *cbp )
{
    xo_residual_frame_typed( g, fenc_slot, pred_slot, qp, NULL, NULL, levels, NULL, nnz, cbp );
}

/* mb_kind[xy]: 0 = inter macroblock of a P slice, 1 = I16x16 macroblock of an I slice, 2 = I4x4 macroblock of an
 * I slice with modes i4_modes[xy][16] (+4: it has a row above but no top-right macroblock); NULL: all inter.
 * Macroblocks are coded in raster order, so an I4x4 macroblock sees its neighbours' reconstruction in pred_slot.
 * luma_dc[xy][16] is written for kind 1 (zeroed otherwise) when not NULL */
void xo_residual_frame_typed( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *pred_slot, int qp,
                              const uint8_t *mb_kind, const uint8_t *i4_modes, int16_t *levels, int16_t *luma_dc,
                              uint8_t *nnz, int16_t *cbp )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, x, y;
    (void)blk_off;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
        for( mb_x = 0; mb_x < g->mb_w; mb_x++ )
        {
            int xy = mb_y * g->mb_w + mb_x;
            pixel_t fenc_y[16*FENC], fenc_c[8*FENC], fdec_y[16*FDEC], fdec_c[8*FDEC];
            const pixel_t *sy = fenc_slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            const pixel_t *sc = fenc_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            pixel_t *py = pred_slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            pixel_t *pc = pred_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            int16_t *lv = levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB;
            mb_out_t o = { lv, lv + 256, lv + 264, nnz + (size_t)xy * X264DSP_RES_NNZ_PER_MB };
            memset( fdec_y, 0, sizeof(fdec_y) );
            memset( fdec_c, 0, sizeof(fdec_c) );
            for( y = 0; y < 16; y++ )
            {
                memcpy( fenc_y + y*FENC, sy + (ptrdiff_t)y*ls, 16 );
                memcpy( fdec_y + y*FDEC, py + (ptrdiff_t)y*ls, 16 );
            }
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    fenc_c[y*FENC + x]      = sc[(ptrdiff_t)y*cs + 2*x];
                    fenc_c[y*FENC + 8 + x]  = sc[(ptrdiff_t)y*cs + 2*x + 1];
                    fdec_c[y*FDEC + x]      = pc[(ptrdiff_t)y*cs + 2*x];
                    fdec_c[y*FDEC + 16 + x] = pc[(ptrdiff_t)y*cs + 2*x + 1];
                }
            if( mb_kind && ( mb_kind[xy] & 3 ) == 2 )
            {
                /* the macroblock with its neighbourhood: row -1 from column -1 to 19, column -1 */
                pixel_t nb[17*FDEC];
                pixel_t *org = nb + FDEC + 8;
                memset( nb, 0, sizeof(nb) );
                memcpy( org - FDEC - 1, py - ls - 1, 21 );
                for( y = 0; y < 16; y++ )
                    memcpy( org + y*FDEC - 1, py + (ptrdiff_t)y*ls - 1, 17 );
                if( luma_dc )
                    memset( luma_dc + (size_t)xy*16, 0, 16*sizeof(int16_t) );
                cbp[xy] = (int16_t)encode_intra4_mb( fenc_y, fenc_c, org, fdec_c, qp, i4_modes + (size_t)xy*16,
                                                     ( mb_kind[xy] & 4 ) != 0, &o );
                for( y = 0; y < 16; y++ )
                    memcpy( fdec_y + y*FDEC, org + y*FDEC, 16 );
            }
            else if( mb_kind && mb_kind[xy] )
            {
                int16_t dc_tmp[16];
                cbp[xy] = (int16_t)encode_intra16_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp, &o, luma_dc ? luma_dc + (size_t)xy*16 : dc_tmp );
            }
            else
            {
                if( luma_dc )
                    memset( luma_dc + (size_t)xy*16, 0, 16*sizeof(int16_t) );
                cbp[xy] = (int16_t)encode_inter_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp, &o );
            }
            for( y = 0; y < 16; y++ )
                memcpy( py + (ptrdiff_t)y*ls, fdec_y + y*FDEC, 16 );
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    pc[(ptrdiff_t)y*cs + 2*x]     = fdec_c[y*FDEC + x];
                    pc[(ptrdiff_t)y*cs + 2*x + 1] = fdec_c[y*FDEC + 16 + x];
                }
        }
}

/* single-macroblock doorway used by the oracle-vs-reference test (same buffers as above) */
int xo_encode_inter_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c,
                        int qp, int16_t *levels, uint8_t *nnz )
{
    mb_out_t o = { levels, levels + 256, levels + 264, nnz };
    return encode_inter_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp, &o );
}
