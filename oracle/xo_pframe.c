/*
 * xo_pframe.c -- CPU oracle: the P-slice macroblock loop.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * Restates, for one reference frame and analyse.inter == 0 (the reference's default),
 *   x264_macroblock_analyse            encoder/analyse.c:1059-1232
 *     x264_mb_analyse_init             encoder/analyse.c:327-420      (MV limits)
 *     x264_macroblock_cache_load       common/macroblock.c            (neighbours -> mvp, pskip_mv; mvpred.c:101-155)
 *     x264_mb_analyse_inter_p16x16     encoder/analyse.c:787-860
 *     x264_me_refine_qpel              encoder/me.c:426-435
 *   x264_macroblock_encode             encoder/macroblock.c:310-485   (x264_mb_mc, residual, forced P_SKIP)
 * on top of the pieces the oracle already has (xo_predict_mv_*, xo_me_search_batch_ex, xo_mc_*, the macroblock
 * residual coder, the P_SKIP probe).  Pinned against the running reference encoder: tests/test_oracle_pframe.py captures
 * every P frame of real encodes (types, vectors, cbp, reconstruction) and requires this function to reproduce them.
 */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

static int clip3( int v, int lo, int hi ) { return v < lo ? lo : v > hi ? hi : v; }

typedef struct
{
    int mv_min[2], mv_max[2], min_spel[2], max_spel[2], min_fpel[2], max_fpel[2];
} limits_t;

/* prediction of macroblock (mb_x, mb_y) at quarter-pel (mvx, mvy) into the recon slot: x264_mb_mc_xywh( 0, 0, 4, 4 ) */
static void mc_mb( const x264dsp_geom_t *g, const uint8_t *fref_slot, uint8_t *dst_slot, int mb_x, int mb_y,
                   int mvx, int mvy, const limits_t *L )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    const pixel_t *src[4];
    pixel_t *dy = dst_slot + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
    pixel_t *dc = dst_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
    pixel_t u[64], v[64];
    int k, x, y;
    mvx = clip3( mvx, L->mv_min[0], L->mv_max[0] );                 /* common/macroblock.c:12-13 */
    mvy = clip3( mvy, L->mv_min[1], L->mv_max[1] );
    for( k = 0; k < 4; k++ )
        src[k] = fref_slot + (size_t)k * g->luma_plane_size + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
    xo_mc_luma( dy, ls, src, ls, mvx, mvy, 16, 16 );
    xo_mc_chroma( u, v, 8, fref_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 ),
                  cs, mvx, mvy, 8, 8 );
    for( y = 0; y < 8; y++ )
        for( x = 0; x < 8; x++ )
        {
            dc[y * cs + 2 * x] = u[y * 8 + x];
            dc[y * cs + 2 * x + 1] = v[y * 8 + x];
        }
}

/* macroblock (mb_x, mb_y) of a slot as the encoder's fenc_buf / fdec_buf (common/macroblock.c:242-265) */
static void load_mb( const x264dsp_geom_t *g, const uint8_t *slot, int mb_x, int mb_y, pixel_t *y_buf, int y_stride,
                     pixel_t *c_buf, int c_stride, int v_off )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    const pixel_t *sy = slot + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
    const pixel_t *sc = slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
    int x, y;
    for( y = 0; y < 16; y++ )
        memcpy( y_buf + y * y_stride, sy + (ptrdiff_t)y * ls, 16 );
    for( y = 0; y < 8; y++ )
        for( x = 0; x < 8; x++ )
        {
            c_buf[y * c_stride + x] = sc[(ptrdiff_t)y * cs + 2 * x];
            c_buf[y * c_stride + v_off + x] = sc[(ptrdiff_t)y * cs + 2 * x + 1];
        }
}

static void store_mb( const x264dsp_geom_t *g, uint8_t *slot, int mb_x, int mb_y, const pixel_t *fdec_y, const pixel_t *fdec_c )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    pixel_t *dy = slot + g->luma_origin + (ptrdiff_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
    pixel_t *dc = slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( mb_y << 3 ) * cs + ( mb_x << 4 );
    int x, y;
    for( y = 0; y < 16; y++ )
        memcpy( dy + (ptrdiff_t)y * ls, fdec_y + y * XO_FDEC_STRIDE, 16 );
    for( y = 0; y < 8; y++ )
        for( x = 0; x < 8; x++ )
        {
            dc[(ptrdiff_t)y * cs + 2 * x] = fdec_c[y * XO_FDEC_STRIDE + x];
            dc[(ptrdiff_t)y * cs + 2 * x + 1] = fdec_c[y * XO_FDEC_STRIDE + 16 + x];
        }
}

/* x264_macroblock_probe_pskip on the prediction at the (clipped) P_SKIP vector, which is left in the recon slot */
static int probe_pskip( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot, uint8_t *recon_slot,
                        int mb_x, int mb_y, const int16_t pskip_mv[2], const limits_t *L, int qp )
{
    pixel_t fenc_y[16 * XO_FENC_STRIDE], fenc_c[8 * XO_FENC_STRIDE], fdec_y[16 * XO_FDEC_STRIDE], fdec_c[8 * XO_FDEC_STRIDE];
    mc_mb( g, fref_slot, recon_slot, mb_x, mb_y, pskip_mv[0], pskip_mv[1], L );
    memset( fdec_y, 0, sizeof(fdec_y) );
    memset( fdec_c, 0, sizeof(fdec_c) );
    load_mb( g, fenc_slot, mb_x, mb_y, fenc_y, XO_FENC_STRIDE, fenc_c, XO_FENC_STRIDE, 8 );
    load_mb( g, recon_slot, mb_x, mb_y, fdec_y, XO_FDEC_STRIDE, fdec_c, XO_FDEC_STRIDE, 16 );
    return xo_probe_pskip_mb( fenc_y, fenc_c, fdec_y, fdec_c, qp );
}

void xo_p_frame( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot, uint8_t *recon_slot,
                 const x264dsp_pframe_params_t *prm, const int16_t *lowres_mv, const int16_t *l0_mv16,
                 int8_t *mb_type, int16_t *mv, int16_t *mvr, int16_t *mvd, int16_t *levels, uint8_t *nnz, int16_t *cbp )
{
    const int W = g->mb_w, H = g->mb_h;
    const int fmv_range = prm->mv_range << 2, border = 6;
    const int lambda = xo_lambda( prm->qp );
    limits_t L;
    int mb_x, mb_y, k;
    memset( &L, 0, sizeof(L) );
    memset( levels, 0, (size_t)W * H * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t) );
    memset( nnz, 0, (size_t)W * H * X264DSP_RES_NNZ_PER_MB );
    for( mb_y = 0; mb_y < H; mb_y++ )
        for( mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            /* neighbours A (left), B (top), C (top-right), D (top-left): reference 0 when inside the frame (every
             * macroblock of a P slice is inter here), -2 when not; their vectors are final */
            const int nxy[4] = { mb_x > 0 ? xy - 1 : -1, mb_y > 0 ? xy - W : -1,
                                 ( mb_y > 0 && mb_x < W - 1 ) ? xy - W + 1 : -1, ( mb_x > 0 && mb_y > 0 ) ? xy - W - 1 : -1 };
            x264dsp_mv_neighbours_t nb;
            int16_t pskip_mv[2], mvp[2];
            int skip = 0, type;
            int16_t *out_levels = levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB;
            uint8_t *out_nnz = nnz + (size_t)xy * X264DSP_RES_NNZ_PER_MB;
            for( k = 0; k < 4; k++ )
            {
                nb.ref[k] = nxy[k] >= 0 ? 0 : -2;
                nb.mv[k][0] = nxy[k] >= 0 ? mv[2 * nxy[k]] : 0;
                nb.mv[k][1] = nxy[k] >= 0 ? mv[2 * nxy[k] + 1] : 0;
            }
            xo_predict_mv_pskip( &nb, pskip_mv );                      /* x264_macroblock_cache_load, P slices */
            if( mvd )
                mvd[2 * xy] = mvd[2 * xy + 1] = 0;

            /* x264_mb_analyse_init (analyse.c:373-397); the vertical limits change at the start of a row only */
            L.mv_min[0] = ( -( mb_x << 4 ) - 24 ) << 2;
            L.mv_max[0] = ( ( ( W - mb_x - 1 ) << 4 ) + 24 ) << 2;
            L.min_spel[0] = clip3( L.mv_min[0], -fmv_range, fmv_range - 1 );
            L.max_spel[0] = clip3( L.mv_max[0], -fmv_range, fmv_range - 1 );
            L.min_fpel[0] = ( L.min_spel[0] >> 2 ) + border;
            L.max_fpel[0] = ( L.max_spel[0] >> 2 ) - border;
            if( mb_x == 0 )
            {
                L.mv_min[1] = ( -( mb_y << 4 ) - 24 ) << 2;
                L.mv_max[1] = ( ( ( H - mb_y - 1 ) << 4 ) + 24 ) << 2;
                L.min_spel[1] = clip3( L.mv_min[1], -fmv_range, fmv_range );
                L.max_spel[1] = clip3( L.mv_max[1], -fmv_range, fmv_range - 1 );
                L.min_fpel[1] = ( L.min_spel[1] >> 2 ) + border;
                L.max_fpel[1] = ( L.max_spel[1] >> 2 ) - border;
            }

            /* fast P_SKIP detection (analyse.c:1093-1105) */
            if( prm->fast_pskip && prm->subpel_refine < 3 )
            {
                int any = 0;
                for( k = 0; k < 4; k++ )
                    any |= nxy[k] >= 0 && mb_type[nxy[k]] == X264DSP_MB_P_SKIP;
                if( any )
                    skip = probe_pskip( g, fenc_slot, fref_slot, recon_slot, mb_x, mb_y, pskip_mv, &L, prm->qp );
            }
            if( skip )
            {
                /* analyse.c:1109-1117: no search; later macroblocks see a zero 16x16 vector */
                mb_type[xy] = X264DSP_MB_P_SKIP;
                mv[2 * xy] = pskip_mv[0]; mv[2 * xy + 1] = pskip_mv[1];
                mvr[2 * xy] = mvr[2 * xy + 1] = 0;
                cbp[xy] = 0;
                continue;                                              /* the prediction is the reconstruction */
            }
            {
                /* x264_mb_analyse_inter_p16x16 */
                x264dsp_me_block_t blk;
                x264dsp_me_result_t res;
                x264dsp_me_params_t mp = { prm->me_method, prm->subpel_refine, prm->me_range, prm->qp, 0 };
                int16_t mvc[9][2];
                int n_mvc = 0;
                memset( &blk, 0, sizeof(blk) );
                xo_predict_mv_16x16( &nb, 0, mvp );
                /* x264_mb_predict_mv_ref16x16 (mvpred.c:167-219) */
                if( lowres_mv && lowres_mv[0] != 0x7fff )
                {
                    mvc[n_mvc][0] = (int16_t)( lowres_mv[2 * xy] * 2 );
                    mvc[n_mvc][1] = (int16_t)( lowres_mv[2 * xy + 1] * 2 );
                    n_mvc++;
                }
                {
                    const int sp[4] = { nxy[0], nxy[1], nxy[3], nxy[2] };      /* left, top, top-left, top-right */
                    for( k = 0; k < 4; k++, n_mvc++ )
                    {
                        mvc[n_mvc][0] = sp[k] >= 0 ? mvr[2 * sp[k]] : 0;
                        mvc[n_mvc][1] = sp[k] >= 0 ? mvr[2 * sp[k] + 1] : 0;
                    }
                }
                if( l0_mv16 )
                {
                    const int t[3] = { xy, mb_x < W - 1 ? xy + 1 : -1, mb_y < H - 1 ? xy + W : -1 };
                    for( k = 0; k < 3; k++ )
                        if( t[k] >= 0 )
                        {
                            mvc[n_mvc][0] = (int16_t)( ( l0_mv16[2 * t[k]] * prm->mvc_scale + 128 ) >> 8 );
                            mvc[n_mvc][1] = (int16_t)( ( l0_mv16[2 * t[k] + 1] * prm->mvc_scale + 128 ) >> 8 );
                            n_mvc++;
                        }
                }
                blk.i_pixel = X264DSP_PIXEL_16x16;
                blk.bx = mb_x << 4;
                blk.by = mb_y << 4;
                blk.mvp[0] = mvp[0]; blk.mvp[1] = mvp[1];
                blk.i_mvc = n_mvc;
                memcpy( blk.mvc, mvc, sizeof(int16_t) * 2 * n_mvc );
                for( k = 0; k < 2; k++ )
                {
                    blk.mv_min_fpel[k] = L.min_fpel[k]; blk.mv_max_fpel[k] = L.max_fpel[k];
                    blk.mv_min_spel[k] = L.min_spel[k]; blk.mv_max_spel[k] = L.max_spel[k];
                }
                xo_me_search_batch_ex( g, fenc_slot, fref_slot, &mp, 1, &blk, &res, 0, NULL );
                mvr[2 * xy] = res.mv[0]; mvr[2 * xy + 1] = res.mv[1];          /* analyse.c:825 */
                /* early termination (analyse.c:839-849) */
                if( prm->fast_pskip && prm->subpel_refine >= 3 && res.cost - res.cost_mv < 300 * lambda
                    && abs( res.mv[0] - pskip_mv[0] ) + abs( res.mv[1] - pskip_mv[1] ) <= 1
                    && probe_pskip( g, fenc_slot, fref_slot, recon_slot, mb_x, mb_y, pskip_mv, &L, prm->qp ) )
                {
                    mb_type[xy] = X264DSP_MB_P_SKIP;
                    mv[2 * xy] = pskip_mv[0]; mv[2 * xy + 1] = pskip_mv[1];
                    cbp[xy] = 0;
                    continue;
                }
                /* x264_me_refine_qpel (analyse.c:1187-1191; one reference: i_ref_cost = 0) */
                xo_me_search_batch_ex( g, fenc_slot, fref_slot, &mp, 1, &blk, &res, 2, NULL );
                mv[2 * xy] = res.mv[0]; mv[2 * xy + 1] = res.mv[1];
                type = X264DSP_MB_P_L0;
            }
            {
                /* x264_macroblock_encode, inter branch (macroblock.c:379-485) */
                pixel_t fenc_y[16 * XO_FENC_STRIDE], fenc_c[8 * XO_FENC_STRIDE], fdec_y[16 * XO_FDEC_STRIDE], fdec_c[8 * XO_FDEC_STRIDE];
                int c;
                mc_mb( g, fref_slot, recon_slot, mb_x, mb_y, mv[2 * xy], mv[2 * xy + 1], &L );
                memset( fdec_y, 0, sizeof(fdec_y) );
                memset( fdec_c, 0, sizeof(fdec_c) );
                load_mb( g, fenc_slot, mb_x, mb_y, fenc_y, XO_FENC_STRIDE, fenc_c, XO_FENC_STRIDE, 8 );
                load_mb( g, recon_slot, mb_x, mb_y, fdec_y, XO_FDEC_STRIDE, fdec_c, XO_FDEC_STRIDE, 16 );
                c = xo_encode_inter_mb( fenc_y, fenc_c, fdec_y, fdec_c, prm->qp, out_levels, out_nnz );
                store_mb( g, recon_slot, mb_x, mb_y, fdec_y, fdec_c );
                cbp[xy] = (int16_t)c;
                /* macroblock.c:465-485: nothing coded and the vector is the P_SKIP one */
                if( !( c & 0x3f ) && mv[2 * xy] == pskip_mv[0] && mv[2 * xy + 1] == pskip_mv[1] )
                    type = X264DSP_MB_P_SKIP;
                mb_type[xy] = (int8_t)type;
                /* what x264_cabac_mvd writes for the 16x16 partition (encoder/cabac.c:278-300) */
                if( mvd && type == X264DSP_MB_P_L0 )
                {
                    mvd[2 * xy] = (int16_t)( mv[2 * xy] - mvp[0] );
                    mvd[2 * xy + 1] = (int16_t)( mv[2 * xy + 1] - mvp[1] );
                }
            }
        }
}
