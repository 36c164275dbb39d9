/*
 * xo_iframe.c -- CPU oracle: the I-slice macroblock loop.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * Restates, for analyse.intra = I4x4 (no 8x8 transform: the reference's build) and the SATD metric every valid subme uses,
 *   x264_macroblock_analyse, I slices      encoder/analyse.c:1079-1088
 *     x264_mb_analyse_intra                encoder/analyse.c:565-763   (I16x16 modes, I4x4 block by block with the
 *                                                                       shortcuts, the early exit and the in-place coding)
 *     x264_mb_analyse_intra_chroma         encoder/analyse.c:509-563
 *     mode availability / mode prediction  encoder/analyse.c:424-508, common/macroblock.h:373-387, common/macroblock.c:655-676
 *   the 16x16 and 8x8c predictors          common/predict.c:42-318
 *   x264_macroblock_encode, intra branches (through xo_encode_intra16_mb / xo_encode_intra4_mb)
 * Pinned against the running reference encoder: tests/test_oracle_iframe.py captures every I frame of real encodes (types,
 * the 4x4 modes the neighbours see, chroma modes, cbp, reconstruction) and requires this function to reproduce them.
 */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

#define FENC XO_FENC_STRIDE
#define FDEC XO_FDEC_STRIDE
#define COST_MAX ( 1 << 28 )
enum { NB_LEFT = 1, NB_TOP = 2, NB_TOPRIGHT = 4, NB_TOPLEFT = 8 };           /* common/macroblock.h:10-13 */

static int clip_u8( int v ) { return v < 0 ? 0 : v > 255 ? 255 : v; }

/* x264_predict_16x16_* / x264_predict_8x8c_* (common/predict.c:42-318) as one routine over the block's edge:
 * kind 0 = V, 1 = H, 2 = DC, 3 = plane, 4 = DC from the left column, 5 = DC from the row above, 6 = 128.
 * A 16x16 block has one DC; an 8x8 chroma block has one per 4x4 quadrant (predict.c:163-222). */
static void predict_block( int size, int kind, pixel_t *src )
{
    int x, y, i;
    if( kind == 3 )
    {
        const int half = size >> 1;
        int H = 0, V = 0, a, b, c, i00;
        for( i = 0; i < half; i++ )
        {
            H += ( i + 1 ) * ( src[half + i - FDEC] - src[half - 2 - i - FDEC] );
            V += ( i + 1 ) * ( src[-1 + ( half + i ) * FDEC] - src[-1 + ( half - 2 - i ) * FDEC] );
        }
        a = 16 * ( src[-1 + ( size - 1 ) * FDEC] + src[size - 1 - FDEC] );
        b = size == 16 ? ( 5 * H + 32 ) >> 6 : ( 17 * H + 16 ) >> 5;
        c = size == 16 ? ( 5 * V + 32 ) >> 6 : ( 17 * V + 16 ) >> 5;
        i00 = a - ( half - 1 ) * ( b + c ) + 16;
        for( y = 0; y < size; y++ )
            for( x = 0; x < size; x++ )
                src[y * FDEC + x] = (pixel_t)clip_u8( ( i00 + b * x + c * y ) >> 5 );
        return;
    }
    if( kind == 0 )
    {
        for( y = 0; y < size; y++ )
            memcpy( src + y * FDEC, src - FDEC, size );
        return;
    }
    if( kind == 1 )
    {
        for( y = 0; y < size; y++ )
            memset( src + y * FDEC, src[y * FDEC - 1], size );
        return;
    }
    {
        int dc[2][2];
        if( size == 16 )
        {
            int s = 0;
            for( i = 0; i < 16; i++ )
                s += ( kind != 5 ? src[-1 + i * FDEC] : 0 ) + ( kind != 4 ? src[i - FDEC] : 0 );
            dc[0][0] = kind == 6 ? 128 : kind == 2 ? ( s + 16 ) >> 5 : ( s + 8 ) >> 4;
            dc[0][1] = dc[1][0] = dc[1][1] = dc[0][0];
        }
        else
        {
            int t0 = 0, t1 = 0, l0 = 0, l1 = 0;
            for( i = 0; i < 4; i++ )
            {
                t0 += src[i - FDEC]; t1 += src[i + 4 - FDEC];
                l0 += src[i * FDEC - 1]; l1 += src[( i + 4 ) * FDEC - 1];
            }
            if( kind == 2 )
            {
                dc[0][0] = ( t0 + l0 + 4 ) >> 3; dc[0][1] = ( t1 + 2 ) >> 2;
                dc[1][0] = ( l1 + 2 ) >> 2;      dc[1][1] = ( t1 + l1 + 4 ) >> 3;
            }
            else if( kind == 4 )
            {
                dc[0][0] = dc[0][1] = ( l0 + 2 ) >> 2;
                dc[1][0] = dc[1][1] = ( l1 + 2 ) >> 2;
            }
            else if( kind == 5 )
            {
                dc[0][0] = dc[1][0] = ( t0 + 2 ) >> 2;
                dc[0][1] = dc[1][1] = ( t1 + 2 ) >> 2;
            }
            else
                dc[0][0] = dc[0][1] = dc[1][0] = dc[1][1] = 128;
        }
        for( y = 0; y < size; y++ )
            for( x = 0; x < size; x++ )
                src[y * FDEC + x] = (pixel_t)dc[size == 16 ? 0 : y >> 2][size == 16 ? 0 : x >> 2];
    }
}

/* the reference's enums (common/predict.h:8-59) -> kind */
static int kind_16x16( int mode ) { return mode; }                                     /* V H DC P DC_LEFT DC_TOP DC_128 */
static int kind_chroma( int mode ) { static const int k[7] = { 2, 1, 0, 3, 4, 5, 6 }; return k[mode]; }   /* DC H V P ... */

void xo_predict_16x16( int mode, pixel_t *src ) { predict_block( 16, kind_16x16( mode ), src ); }
void xo_predict_chroma( int mode, pixel_t *src ) { predict_block( 8, kind_chroma( mode ), src ); }

static int ue_bits( int v )                       /* bs_size_ue */
{
    int n = 1;
    for( v++; v > 1; v >>= 1 )
        n += 2;
    return n;
}

/* rows of the availability tables (analyse.c:424-508): 0 none, 1 left, 2 top, 3 top + left, 4 top + left + top-left */
static int avail_row( int nb )
{
    const int k = nb & ( NB_TOP | NB_LEFT | NB_TOPLEFT );
    return k == ( NB_TOP | NB_LEFT | NB_TOPLEFT ) ? 4 : k & ( NB_TOP | NB_LEFT );
}
static const int8_t modes_16x16[5][5] = { { 6, -1 }, { 4, 1, -1 }, { 5, 0, -1 }, { 0, 1, 2, -1 }, { 0, 1, 2, 3, -1 } };
static const int8_t modes_chroma[5][5] = { { 6, -1 }, { 4, 1, -1 }, { 5, 2, -1 }, { 2, 1, 0, -1 }, { 2, 1, 0, 3, -1 } };
static const int8_t modes_4x4[5][10] = { { 11, -1 }, { 9, 1, 8, -1 }, { 10, 0, 3, 7, -1 }, { 2, 1, 0, 3, 7, 8, -1 },
                                         { 2, 1, 0, 3, 4, 5, 6, 7, 8, -1 } };
static const int8_t mode4x4_fix[13] = { -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 2, 2, 2 };     /* predict.h:60-68, index = mode + 1 */
static const int8_t chroma_fix[7] = { 0, 1, 2, 3, 0, 0, 0 }, mode16_fix[7] = { 0, 1, 2, 3, 2, 2, 2 };

typedef struct
{
    int type;                   /* 0 = I_4x4, 2 = I_16x16 */
    int mode16, chroma_mode;
    uint8_t modes4[16];         /* coding order, the predictors actually used (DC variants included) */
} intra_decision_t;

/* x264_mb_analyse_intra + the I-slice branch of x264_macroblock_analyse + x264_mb_analyse_intra_chroma for one macroblock.
 * fenc_y / fenc_c: fenc_buf; fdec_y / fdec_c: fdec_buf with the reconstructed neighbours (row -1 from column -1 to 19,
 * column -1); both are scratch afterwards.  nb = available neighbour macroblocks; mode_left[4] / mode_top[4]: the 4x4 modes
 * of the blocks next to the macroblock (-1 = not available, as h->mb.cache.intra4x4_pred_mode holds them). */
static void analyse_intra_mb( const pixel_t *fenc_y, const pixel_t *fenc_c, pixel_t *fdec_y, pixel_t *fdec_c, int nb,
                              const int8_t mode_left[4], const int8_t mode_top[4], int qp, intra_decision_t *D )
{
    const int lambda = xo_lambda( qp );
    const int8_t *list;
    int satd16 = COST_MAX, satd4 = COST_MAX, idx, k;
    int8_t cache[5][5];         /* [1 + by][1 + bx]: modes of the sixteen blocks and their left / top neighbours */
    uint16_t mf[16], bias[16];
    int dequant[6][16];
    xo_quant_tables( 0, qp, mf, bias );
    xo_dequant_table( dequant );

    /* ---- 16x16 (analyse.c:590-627): V, H, DC with 1 / 3 / 3 bits, then the plane mode */
    D->mode16 = 0;
    for( list = modes_16x16[avail_row( nb )]; *list >= 0; list++ )
    {
        int c;
        xo_predict_16x16( *list, fdec_y );
        c = xo_satd( X264DSP_PIXEL_16x16, fdec_y, FDEC, fenc_y, FENC ) + lambda * ue_bits( mode16_fix[*list] );
        if( c < satd16 ) { satd16 = c; D->mode16 = *list; }
    }

    /* ---- 4x4 (analyse.c:629-763) */
    for( k = 0; k < 4; k++ )
    {
        cache[0][1 + k] = mode_top[k];
        cache[1 + k][0] = mode_left[k];
    }
    {
        int cost = lambda * 40;
        const int thresh = satd16;                       /* b_early_terminate, no inter cost in an I slice */
        for( idx = 0; ; idx++ )
        {
            const int bx = ( idx & 1 ) + ( ( idx >> 2 ) & 1 ) * 2, by = ( ( idx >> 1 ) & 1 ) + ( ( idx >> 3 ) & 1 ) * 2;
            const pixel_t *src = fenc_y + by * 4 * FENC + bx * 4;
            pixel_t *dst = fdec_y + by * 4 * FDEC + bx * 4;
            int nb4, row, best = COST_MAX, best_mode = 2, pred, ma, mb_, satd[9];
            const int8_t *extra = NULL;
            static const int8_t shortcut[2][2][5] = {                  /* [all nine available][favor vertical] */
                { { 8, -1 }, { 3, 7, -1 } }, { { 4, 6, 8, -1 }, { 3, 4, 5, 7, -1 } } };
            /* neighbours of the block (macroblock.c:217-226, 655-676) */
            if( idx == 6 || idx == 9 || idx == 12 || idx == 14 )
                nb4 = NB_LEFT | NB_TOP | NB_TOPLEFT | NB_TOPRIGHT;
            else if( idx == 3 || idx == 7 || idx == 11 || idx == 13 || idx == 15 )
                nb4 = NB_LEFT | NB_TOP | NB_TOPLEFT;
            else if( idx == 0 )
                nb4 = ( nb & ( NB_TOP | NB_LEFT | NB_TOPLEFT ) ) | ( ( nb & NB_TOP ) ? NB_TOPRIGHT : 0 );
            else if( idx == 1 || idx == 4 )
                nb4 = NB_LEFT | ( ( nb & NB_TOP ) ? NB_TOP | NB_TOPLEFT | NB_TOPRIGHT : 0 );
            else if( idx == 5 )
                nb4 = NB_LEFT | ( nb & NB_TOPRIGHT ) | ( ( nb & NB_TOP ) ? NB_TOP | NB_TOPLEFT : 0 );
            else                                                        /* 2, 8, 10 */
                nb4 = NB_TOP | NB_TOPRIGHT | ( ( nb & NB_LEFT ) ? NB_LEFT | NB_TOPLEFT : 0 );
            row = avail_row( nb4 );
            /* x264_mb_predict_intra4x4_mode */
            ma = mode4x4_fix[cache[1 + by][bx] + 1];
            mb_ = mode4x4_fix[cache[by][1 + bx] + 1];
            pred = ma < mb_ ? ma : mb_;
            if( pred < 0 )
                pred = 2;
            if( ( nb4 & ( NB_TOPRIGHT | NB_TOP ) ) == NB_TOP )           /* emulate the missing top-right samples */
                memset( dst + 4 - FDEC, dst[3 - FDEC], 4 );
            list = modes_4x4[row];
            if( row >= 3 )
            {
                /* DC / H / V all available: intra_mbcmp_x3_4x4, then the direction-dependent four or a shortcut list */
                int favor_vertical;
                for( k = 0; k < 3; k++ )
                {
                    xo_predict_4x4( k, dst );
                    satd[k] = xo_satd( X264DSP_PIXEL_4x4, dst, FDEC, src, FENC );
                }
                favor_vertical = satd[1] > satd[0];
                if( row == 4 )
                {
                    static const int8_t four[2][4] = { { 3, 4, 6, 8 }, { 3, 4, 5, 7 } };
                    for( k = 0; k < 4; k++ )
                    {
                        xo_predict_4x4( four[favor_vertical][k], dst );
                        satd[four[favor_vertical][k]] = xo_satd( X264DSP_PIXEL_4x4, dst, FDEC, src, FENC );
                    }
                }
                satd[pred] -= 3 * lambda;
                best = satd[2]; best_mode = 2;
                if( satd[1] < best ) { best = satd[1]; best_mode = 1; }
                if( satd[0] < best ) { best = satd[0]; best_mode = 0; }
                if( row == 4 )
                {
                    static const int8_t order[2][4] = { { 3, 4, 6, 8 }, { 3, 4, 5, 7 } };
                    for( k = 0; k < 4; k++ )
                        if( satd[order[favor_vertical][k]] < best ) { best = satd[order[favor_vertical][k]]; best_mode = order[favor_vertical][k]; }
                    list = NULL;
                }
                else
                    extra = shortcut[0][favor_vertical];
                if( row != 4 )
                    list = extra;
            }
            if( list && best > 0 )
                for( ; *list >= 0; list++ )
                {
                    int c;
                    xo_predict_4x4( *list, dst );
                    c = xo_satd( X264DSP_PIXEL_4x4, dst, FDEC, src, FENC );
                    if( pred == mode4x4_fix[*list + 1] )
                    {
                        c -= 3 * lambda;
                        if( c <= 0 )
                        {
                            best = c;
                            best_mode = *list;
                            break;
                        }
                    }
                    if( c < best ) { best = c; best_mode = *list; }
                }
            D->modes4[idx] = (uint8_t)best_mode;
            cost += best + 3 * lambda;
            if( cost > thresh || idx == 15 )
                break;
            /* predict with the chosen mode and code the block now: the next ones predict from its reconstruction */
            cache[1 + by][1 + bx] = (int8_t)best_mode;
            {
                coef_t dct[16];
                xo_predict_4x4( best_mode, dst );
                xo_sub4x4_dct( dct, src, dst );
                if( xo_quant_4x4( dct, mf, bias ) )
                {
                    xo_dequant_4x4( dct, dequant, qp );
                    xo_add4x4_idct( dst, dct );
                }
            }
        }
        if( idx == 15 )
            satd4 = cost;
    }
    D->type = satd4 < satd16 ? 0 : 2;                    /* COPY2_IF_LT( i_cost, i_satd_i4x4, type, I_4x4 ) */

    /* ---- chroma (analyse.c:509-563): compared in the order V, H, DC, plane when all four are available */
    {
        int best = COST_MAX;
        pixel_t *fu = fdec_c, *fv = fdec_c + 16;
        D->chroma_mode = 0;
        for( list = modes_chroma[avail_row( nb )]; *list >= 0; list++ )
        {
            int c;
            xo_predict_chroma( *list, fu );
            xo_predict_chroma( *list, fv );
            c = xo_satd( X264DSP_PIXEL_8x8, fu, FDEC, fenc_c, FENC ) + xo_satd( X264DSP_PIXEL_8x8, fv, FDEC, fenc_c + 8, FENC )
              + lambda * ue_bits( chroma_fix[*list] );
            if( c < best ) { best = c; D->chroma_mode = *list; }
        }
    }
}

/* Every macroblock of an I frame, in raster order: analysis, then x264_macroblock_encode's intra branch.
 *   recon_slot  receives the reconstruction (luma plane N, NV12 chroma)
 *   mb_type     [mb] 0 = I_4x4, 2 = I_16x16 (the reference's enum);  mode16 [mb];  chroma_mode [mb]
 *   modes4      [mb][16] the 4x4 predictors in coding order (I_16x16 macroblocks: all 2 = DC, what the neighbours see)
 *   levels / luma_dc / nnz / cbp  as xo_residual_frame_typed */
void xo_i_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, uint8_t *recon_slot, int qp, int8_t *mb_type,
                 uint8_t *mode16, uint8_t *chroma_mode, uint8_t *modes4, int16_t *levels, int16_t *luma_dc, uint8_t *nnz,
                 int16_t *cbp )
{
    const int W = g->mb_w, H = g->mb_h, ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, x, y, k;
    for( mb_y = 0; mb_y < H; mb_y++ )
        for( mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            const int nb = ( mb_x > 0 ? NB_LEFT : 0 ) | ( mb_y > 0 ? NB_TOP : 0 ) | ( mb_x > 0 && mb_y > 0 ? NB_TOPLEFT : 0 )
                         | ( mb_y > 0 && mb_x < W - 1 ? NB_TOPRIGHT : 0 );
            pixel_t fenc_y[16 * FENC], fenc_c[8 * FENC];
            pixel_t ybuf[2][18 * FDEC + 32], cbuf[2][10 * FDEC + 32];
            pixel_t *fy[2], *fc[2];
            const pixel_t *sy = fenc_slot + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
            const pixel_t *sc = fenc_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
            pixel_t *ry = recon_slot + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
            pixel_t *rc = recon_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
            int8_t mode_left[4], mode_top[4];
            intra_decision_t D;
            int16_t *out_levels = levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB;
            int16_t *out_dc = luma_dc + (size_t)xy * 16;
            uint8_t *out_nnz = nnz + (size_t)xy * X264DSP_RES_NNZ_PER_MB;
            int c;
            for( y = 0; y < 16; y++ )
                memcpy( fenc_y + y * FENC, sy + (ptrdiff_t)y * ls, 16 );
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    fenc_c[y * FENC + x] = sc[(ptrdiff_t)y * cs + 2 * x];
                    fenc_c[y * FENC + 8 + x] = sc[(ptrdiff_t)y * cs + 2 * x + 1];
                }
            /* fdec_buf: the macroblock with its reconstructed neighbourhood, twice (analysis scribbles over its copy) */
            for( k = 0; k < 2; k++ )
            {
                memset( ybuf[k], 0, sizeof(ybuf[k]) );
                memset( cbuf[k], 0, sizeof(cbuf[k]) );
                fy[k] = ybuf[k] + FDEC + 8;
                fc[k] = cbuf[k] + FDEC + 8;                    /* U at +0, V at +16 */
                if( nb & NB_TOP )
                    for( x = -1; x < 20; x++ )
                        fy[k][-FDEC + x] = ry[-(ptrdiff_t)ls + x];
                else if( nb & NB_LEFT )
                    fy[k][-FDEC - 1] = 0;
                if( nb & NB_LEFT )
                    for( y = 0; y < 16; y++ )
                        fy[k][y * FDEC - 1] = ry[(ptrdiff_t)y * ls - 1];
                if( nb & NB_TOP )
                    for( x = -1; x < 8; x++ )
                    {
                        fc[k][-FDEC + x] = rc[-(ptrdiff_t)cs + 2 * x];
                        fc[k][-FDEC + 16 + x] = rc[-(ptrdiff_t)cs + 2 * x + 1];
                    }
                if( nb & NB_LEFT )
                    for( y = 0; y < 8; y++ )
                    {
                        fc[k][y * FDEC - 1] = rc[(ptrdiff_t)y * cs - 2];
                        fc[k][y * FDEC + 15] = rc[(ptrdiff_t)y * cs - 1];
                    }
            }
            /* the 4x4 modes next to the macroblock: the left macroblock's right column (blocks 5, 7, 13, 15), the upper
             * macroblock's bottom row (10, 11, 14, 15); -1 outside the frame (macroblock.c:447-522) */
            {
                static const int8_t right_col[4] = { 5, 7, 13, 15 }, bottom_row[4] = { 10, 11, 14, 15 };
                for( k = 0; k < 4; k++ )
                {
                    mode_left[k] = ( nb & NB_LEFT ) ? (int8_t)modes4[(size_t)( xy - 1 ) * 16 + right_col[k]] : -1;
                    mode_top[k] = ( nb & NB_TOP ) ? (int8_t)modes4[(size_t)( xy - W ) * 16 + bottom_row[k]] : -1;
                }
            }
            analyse_intra_mb( fenc_y, fenc_c, fy[0], fc[0], nb, mode_left, mode_top, qp, &D );
            mb_type[xy] = (int8_t)D.type;
            mode16[xy] = (uint8_t)D.mode16;
            chroma_mode[xy] = (uint8_t)D.chroma_mode;
            /* x264_macroblock_encode on a fresh fdec_buf */
            xo_predict_chroma( D.chroma_mode, fc[1] );
            xo_predict_chroma( D.chroma_mode, fc[1] + 16 );
            memset( out_dc, 0, 16 * sizeof(int16_t) );
            if( D.type == 2 )
            {
                xo_predict_16x16( D.mode16, fy[1] );
                c = xo_encode_intra16_mb( fenc_y, fenc_c, fy[1], fc[1], qp, out_levels, out_dc, out_nnz );
                memset( modes4 + (size_t)xy * 16, 2, 16 );
            }
            else
            {
                const int replicate5 = ( nb & ( NB_TOPRIGHT | NB_TOP ) ) == NB_TOP;
                c = xo_encode_intra4_mb( fenc_y, fenc_c, fy[1], fc[1], qp, D.modes4, replicate5, out_levels, out_nnz );
                memcpy( modes4 + (size_t)xy * 16, D.modes4, 16 );
            }
            cbp[xy] = (int16_t)c;
            for( y = 0; y < 16; y++ )
                memcpy( ry + (ptrdiff_t)y * ls, fy[1] + y * FDEC, 16 );
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    rc[(ptrdiff_t)y * cs + 2 * x] = fc[1][y * FDEC + x];
                    rc[(ptrdiff_t)y * cs + 2 * x + 1] = fc[1][y * FDEC + 16 + x];
                }
        }
}
