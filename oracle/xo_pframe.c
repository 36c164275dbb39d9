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

/* ================================================================================================
 * The same loop with analyse.inter = X264_ANALYSE_PSUB16x16: after the 16x16 search come
 *   x264_mb_analyse_inter_p8x8    encoder/analyse.c:864-921   four 8x8 searches, each predicted from the cache as it fills
 *   x264_mb_analyse_inter_p16x8   encoder/analyse.c:923-990   } only when the 8x8 cost says they might pay
 *   x264_mb_analyse_inter_p8x16   encoder/analyse.c:992-1054  } (analyse.c:1150-1170), each with its early exit
 * the cost comparison (analyse.c:1133-1172), x264_me_refine_qpel on every partition of the winner (1176-1203) and the
 * per-partition x264_mb_mc (common/macroblock.c:28-48).  One reference frame: every i_ref_cost is zero.
 *
 * Vectors are kept per 8x8 block (mv8 [mb][4][2], raster order inside the macroblock) -- with no sub-8x8 partitions that is
 * exactly the granularity of h->mb.cache.mv -- and x264_mb_predict_mv's neighbours A / B / C / D of a partition are read
 * from the cells the reference reads (common/mvpred.c:22-99): other macroblocks' final vectors, or this macroblock's own
 * partitions searched before it.  mvd8 [mb][4][2]: what x264_cabac_mvd writes per partition (encoder/cabac.c:278-300, the
 * prediction taken from the FINAL vectors), replicated over the partition's 8x8 blocks.
 */
typedef struct
{
    int16_t mv[2];
    int ref;                       /* 0, or -2 = not available */
} cell_t;

/* the vector in 4x4 cell (cx, cy) relative to the macroblock's top-left cell (-1 .. 4 in x, -1 .. 3 in y), as the cache holds
 * it at this point of the analysis: cur[] = this macroblock's 8x8 blocks, set[] = which of them have been written */
static cell_t cell_at( int cx, int cy, int mb_x, int mb_y, int W, const int16_t *mv8, const int16_t cur[4][2], const int set[4] )
{
    cell_t c = { { 0, 0 }, -2 };
    if( cx >= 0 && cx < 4 && cy >= 0 )
    {
        const int k = ( cy >> 1 ) * 2 + ( cx >> 1 );
        if( set[k] )
        {
            c.mv[0] = cur[k][0]; c.mv[1] = cur[k][1]; c.ref = 0;
        }
        return c;
    }
    {
        const int nx = mb_x + ( cx < 0 ? -1 : cx > 3 ? 1 : 0 ), ny = mb_y + ( cy < 0 ? -1 : 0 );
        const int lx = cx & 3, ly = cy & 3;
        if( nx < 0 || nx >= W || ny < 0 || ( cx > 3 && cy >= 0 ) )      /* outside the frame, or to the right: not coded yet */
            return c;
        {
            const int16_t *m = mv8 + ( (size_t)( ny * W + nx ) * 4 + ( ly >> 1 ) * 2 + ( lx >> 1 ) ) * 2;
            c.mv[0] = m[0]; c.mv[1] = m[1]; c.ref = 0;
        }
    }
    return c;
}

/* x264_mb_predict_mv( idx at cell (x, y), width w cells ) with the partition rule `shape` of xo_predict_mv_part */
static void predict_part( int x, int y, int w, int shape, int mb_x, int mb_y, int W, const int16_t *mv8,
                          const int16_t cur[4][2], const int set[4], int16_t mvp[2] )
{
    const cell_t n[4] = { cell_at( x - 1, y, mb_x, mb_y, W, mv8, cur, set ), cell_at( x, y - 1, mb_x, mb_y, W, mv8, cur, set ),
                          cell_at( x + w, y - 1, mb_x, mb_y, W, mv8, cur, set ), cell_at( x - 1, y - 1, mb_x, mb_y, W, mv8, cur, set ) };
    x264dsp_mv_neighbours_t nb;
    int k;
    for( k = 0; k < 4; k++ )
    {
        nb.ref[k] = (int8_t)n[k].ref;
        nb.mv[k][0] = n[k].mv[0];
        nb.mv[k][1] = n[k].mv[1];
    }
    xo_predict_mv_part( &nb, 0, shape, 0, mvp );
}

typedef struct
{
    x264dsp_me_block_t blk;
    x264dsp_me_result_t res;
} part_me_t;

static void search_part( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot,
                         const x264dsp_me_params_t *mp, const limits_t *L, int i_pixel, int bx, int by, const int16_t mvp[2],
                         int16_t (*mvc)[2], int n_mvc, part_me_t *m )
{
    int k;
    memset( &m->blk, 0, sizeof(m->blk) );
    m->blk.i_pixel = i_pixel;
    m->blk.bx = bx;
    m->blk.by = by;
    m->blk.mvp[0] = mvp[0]; m->blk.mvp[1] = mvp[1];
    m->blk.i_mvc = n_mvc;
    memcpy( m->blk.mvc, mvc, sizeof(int16_t) * 2 * n_mvc );
    for( k = 0; k < 2; k++ )
    {
        m->blk.mv_min_fpel[k] = L->min_fpel[k]; m->blk.mv_max_fpel[k] = L->max_fpel[k];
        m->blk.mv_min_spel[k] = L->min_spel[k]; m->blk.mv_max_spel[k] = L->max_spel[k];
    }
    xo_me_search_batch_ex( g, fenc_slot, fref_slot, mp, 1, &m->blk, &m->res, 0, NULL );
}

void xo_p_frame_part( const x264dsp_geom_t *g, const uint8_t *fenc_slot, const uint8_t *fref_slot, uint8_t *recon_slot,
                      const x264dsp_pframe_params_t *prm, const int16_t *lowres_mv, const int16_t *l0_mv16,
                      int8_t *mb_type, uint8_t *partition, int16_t *mv8, int16_t *mvr, int16_t *mvd8, int16_t *levels,
                      uint8_t *nnz, int16_t *cbp )
{
    const int W = g->mb_w, H = g->mb_h;
    const int fmv_range = prm->mv_range << 2, border = 6;
    const int lambda = xo_lambda( prm->qp );
    const int psub = prm->analyse_inter != 0;
    const x264dsp_me_params_t mp = { prm->me_method, prm->subpel_refine, prm->me_range, prm->qp, 0 };
    const int none[4] = { 0, 0, 0, 0 };
    limits_t L;
    int mb_x, mb_y, k;
    memset( &L, 0, sizeof(L) );
    memset( levels, 0, (size_t)W * H * X264DSP_RES_LEVELS_PER_MB * sizeof(int16_t) );
    memset( nnz, 0, (size_t)W * H * X264DSP_RES_NNZ_PER_MB );
    memset( mvd8, 0, (size_t)W * H * 8 * sizeof(int16_t) );
    for( mb_y = 0; mb_y < H; mb_y++ )
        for( mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            const int nxy[4] = { mb_x > 0 ? xy - 1 : -1, mb_y > 0 ? xy - W : -1,
                                 ( mb_y > 0 && mb_x < W - 1 ) ? xy - W + 1 : -1, ( mb_x > 0 && mb_y > 0 ) ? xy - W - 1 : -1 };
            int16_t cur[4][2] = { { 0 } }, pskip_mv[2], mvp16[2];
            int set[4] = { 0, 0, 0, 0 }, skip = 0, type = X264DSP_MB_P_L0, part = 16;
            int16_t *out_mv = mv8 + (size_t)xy * 8;
            part_me_t m16, m8[4], m168[2], m816[2];
            x264dsp_mv_neighbours_t nb16;
            {
                const cell_t n[4] = { cell_at( -1, 0, mb_x, mb_y, W, mv8, cur, none ), cell_at( 0, -1, mb_x, mb_y, W, mv8, cur, none ),
                                      cell_at( 4, -1, mb_x, mb_y, W, mv8, cur, none ), cell_at( -1, -1, mb_x, mb_y, W, mv8, cur, none ) };
                for( k = 0; k < 4; k++ )
                {
                    nb16.ref[k] = (int8_t)n[k].ref;
                    nb16.mv[k][0] = n[k].mv[0];
                    nb16.mv[k][1] = n[k].mv[1];
                }
            }
            xo_predict_mv_pskip( &nb16, pskip_mv );
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
                mb_type[xy] = X264DSP_MB_P_SKIP;
                partition[xy] = 16;
                for( k = 0; k < 4; k++ ) { out_mv[2 * k] = pskip_mv[0]; out_mv[2 * k + 1] = pskip_mv[1]; }
                mvr[2 * xy] = mvr[2 * xy + 1] = 0;
                cbp[xy] = 0;
                continue;
            }
            /* ---- P16x16 */
            {
                int16_t mvc[9][2];
                int n_mvc = 0;
                xo_predict_mv_16x16( &nb16, 0, mvp16 );
                if( lowres_mv && lowres_mv[0] != 0x7fff )
                {
                    mvc[n_mvc][0] = (int16_t)( lowres_mv[2 * xy] * 2 );
                    mvc[n_mvc][1] = (int16_t)( lowres_mv[2 * xy + 1] * 2 );
                    n_mvc++;
                }
                {
                    const int sp[4] = { nxy[0], nxy[1], nxy[3], nxy[2] };
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
                search_part( g, fenc_slot, fref_slot, &mp, &L, X264DSP_PIXEL_16x16, mb_x << 4, mb_y << 4, mvp16, mvc, n_mvc, &m16 );
            }
            mvr[2 * xy] = m16.res.mv[0]; mvr[2 * xy + 1] = m16.res.mv[1];
            if( prm->fast_pskip && prm->subpel_refine >= 3 && m16.res.cost - m16.res.cost_mv < 300 * lambda
                && abs( m16.res.mv[0] - pskip_mv[0] ) + abs( m16.res.mv[1] - pskip_mv[1] ) <= 1
                && probe_pskip( g, fenc_slot, fref_slot, recon_slot, mb_x, mb_y, pskip_mv, &L, prm->qp ) )
            {
                mb_type[xy] = X264DSP_MB_P_SKIP;
                partition[xy] = 16;
                for( k = 0; k < 4; k++ ) { out_mv[2 * k] = pskip_mv[0]; out_mv[2 * k + 1] = pskip_mv[1]; }
                cbp[xy] = 0;
                continue;
            }
            if( psub )
            {
                int16_t mvc[5][2], mvp[2];
                int satd8[4], cost8 = 0, i_cost = m16.res.cost, cost168 = 1 << 28, cost816 = 1 << 28, i;
                /* ---- P8x8 (analyse.c:864-921) */
                mvc[0][0] = m16.res.mv[0]; mvc[0][1] = m16.res.mv[1];
                for( i = 0; i < 4; i++ )
                {
                    const int x8 = i & 1, y8 = i >> 1;
                    predict_part( 2 * x8, 2 * y8, 2, 0, mb_x, mb_y, W, mv8, cur, set, mvp );
                    search_part( g, fenc_slot, fref_slot, &mp, &L, X264DSP_PIXEL_8x8, ( mb_x << 4 ) + 8 * x8, ( mb_y << 4 ) + 8 * y8,
                                 mvp, mvc, 1 + i, &m8[i] );
                    cur[i][0] = m8[i].res.mv[0]; cur[i][1] = m8[i].res.mv[1];
                    set[i] = 1;
                    mvc[1 + i][0] = m8[i].res.mv[0]; mvc[1 + i][1] = m8[i].res.mv[1];
                    satd8[i] = m8[i].res.cost - m8[i].res.cost_mv;
                    cost8 += m8[i].res.cost;
                }
                if( cost8 < i_cost )                                      /* analyse.c:1138-1144 (b_early_terminate) */
                {
                    type = X264DSP_MB_P_8x8;
                    part = 13;
                    i_cost = cost8;
                }
                if( cost8 < m16.res.cost + m8[1].res.cost_mv + m8[2].res.cost_mv )         /* analyse.c:1153-1170 */
                {
                    int16_t c3[3][2];
                    const int est168 = satd8[2] + satd8[3] + ( ( m8[2].res.cost_mv + m8[3].res.cost_mv + 1 ) >> 1 );
                    const int est816 = satd8[1] + satd8[3] + ( ( m8[1].res.cost_mv + m8[3].res.cost_mv + 1 ) >> 1 );
                    int16_t c168[4][2], c816[4][2];
                    int s168[4] = { 1, 1, 1, 1 }, s816[4] = { 1, 1, 1, 1 };
                    /* ---- P16x8 (analyse.c:923-990): the cache still holds the four 8x8 vectors */
                    memcpy( c168, cur, sizeof(c168) );
                    for( i = 0; i < 2; i++ )
                    {
                        c3[0][0] = m16.res.mv[0]; c3[0][1] = m16.res.mv[1];
                        memcpy( c3[1], m8[2 * i].res.mv, 4 );
                        memcpy( c3[2], m8[2 * i + 1].res.mv, 4 );
                        predict_part( 0, 2 * i, 4, 1 + i, mb_x, mb_y, W, mv8, c168, s168, mvp );
                        search_part( g, fenc_slot, fref_slot, &mp, &L, X264DSP_PIXEL_16x8, mb_x << 4, ( mb_y << 4 ) + 8 * i, mvp, c3, 3, &m168[i] );
                        if( i == 0 && m168[0].res.cost + est168 > i_cost )
                            break;
                        memcpy( c168[2 * i], m168[i].res.mv, 4 );
                        memcpy( c168[2 * i + 1], m168[i].res.mv, 4 );
                    }
                    if( i == 2 )
                    {
                        cost168 = m168[0].res.cost + m168[1].res.cost;
                        memcpy( cur, c168, sizeof(c168) );             /* what the cache holds when P8x16 starts */
                    }
                    if( cost168 < i_cost )
                    {
                        i_cost = cost168;
                        type = X264DSP_MB_P_L0;
                        part = 14;
                    }
                    /* ---- P8x16 (analyse.c:992-1054) */
                    memcpy( c816, cur, sizeof(c816) );
                    for( i = 0; i < 2; i++ )
                    {
                        c3[0][0] = m16.res.mv[0]; c3[0][1] = m16.res.mv[1];
                        memcpy( c3[1], m8[i].res.mv, 4 );
                        memcpy( c3[2], m8[i + 2].res.mv, 4 );
                        predict_part( 2 * i, 0, 2, 3 + i, mb_x, mb_y, W, mv8, c816, s816, mvp );
                        search_part( g, fenc_slot, fref_slot, &mp, &L, X264DSP_PIXEL_8x16, ( mb_x << 4 ) + 8 * i, mb_y << 4, mvp, c3, 3, &m816[i] );
                        if( i == 0 && m816[0].res.cost + est816 > i_cost )
                            break;
                        memcpy( c816[i], m816[i].res.mv, 4 );
                        memcpy( c816[i + 2], m816[i].res.mv, 4 );
                    }
                    if( i == 2 )
                        cost816 = m816[0].res.cost + m816[1].res.cost;
                    if( cost816 < i_cost )
                    {
                        i_cost = cost816;
                        type = X264DSP_MB_P_L0;
                        part = 15;
                    }
                }
            }
            /* ---- x264_me_refine_qpel on every partition of the winner (analyse.c:1176-1203), then the final vectors */
            {
                part_me_t *win[4] = { &m16, &m16, &m16, &m16 };
                int n_part = 1;
                if( part == 14 ) { win[0] = win[1] = &m168[0]; win[2] = win[3] = &m168[1]; n_part = 2; }
                else if( part == 15 ) { win[0] = win[2] = &m816[0]; win[1] = win[3] = &m816[1]; n_part = 2; }
                else if( part == 13 ) { win[0] = &m8[0]; win[1] = &m8[1]; win[2] = &m8[2]; win[3] = &m8[3]; n_part = 4; }
                {
                    part_me_t *done[4] = { NULL, NULL, NULL, NULL };
                    int nd = 0;
                    for( k = 0; k < 4; k++ )
                    {
                        int j, seen = 0;
                        for( j = 0; j < nd; j++ )
                            seen |= done[j] == win[k];
                        if( !seen )
                        {
                            xo_me_search_batch_ex( g, fenc_slot, fref_slot, &mp, 1, &win[k]->blk, &win[k]->res, 2, NULL );
                            done[nd++] = win[k];
                        }
                    }
                    (void)n_part;
                }
                for( k = 0; k < 4; k++ )
                {
                    out_mv[2 * k] = win[k]->res.mv[0];
                    out_mv[2 * k + 1] = win[k]->res.mv[1];
                }
            }
            {
                /* ---- x264_macroblock_encode: x264_mb_mc per partition (8x8-wise the same samples), residual, forced P_SKIP */
                pixel_t fenc_y[16 * XO_FENC_STRIDE], fenc_c[8 * XO_FENC_STRIDE], fdec_y[16 * XO_FDEC_STRIDE], fdec_c[8 * XO_FDEC_STRIDE];
                int16_t *out_levels = levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB;
                uint8_t *out_nnz = nnz + (size_t)xy * X264DSP_RES_NNZ_PER_MB;
                const int ls = g->luma_stride, cs = g->chroma_stride;
                int c, p, x, y;
                for( p = 0; p < 4; p++ )
                {
                    const int px = ( p & 1 ) * 8, py = ( p >> 1 ) * 8;
                    const int mvx = clip3( out_mv[2 * p], L.mv_min[0], L.mv_max[0] ), mvy = clip3( out_mv[2 * p + 1], L.mv_min[1], L.mv_max[1] );
                    const pixel_t *src[4];
                    pixel_t *dy = recon_slot + g->luma_origin + (ptrdiff_t)( ( mb_y << 4 ) + py ) * ls + ( mb_x << 4 ) + px;
                    pixel_t *dc = recon_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( ( mb_y << 3 ) + py / 2 ) * cs + ( mb_x << 4 ) + px;
                    pixel_t u[16], v[16];
                    for( k = 0; k < 4; k++ )
                        src[k] = fref_slot + (size_t)k * g->luma_plane_size + g->luma_origin + (ptrdiff_t)( ( mb_y << 4 ) + py ) * ls + ( mb_x << 4 ) + px;
                    xo_mc_luma( dy, ls, src, ls, mvx, mvy, 8, 8 );
                    xo_mc_chroma( u, v, 4, fref_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)( ( mb_y << 3 ) + py / 2 ) * cs + ( mb_x << 4 ) + px,
                                  cs, mvx, mvy, 4, 4 );
                    for( y = 0; y < 4; y++ )
                        for( x = 0; x < 4; x++ )
                        {
                            dc[y * cs + 2 * x] = u[y * 4 + x];
                            dc[y * cs + 2 * x + 1] = v[y * 4 + x];
                        }
                }
                memset( fdec_y, 0, sizeof(fdec_y) );
                memset( fdec_c, 0, sizeof(fdec_c) );
                load_mb( g, fenc_slot, mb_x, mb_y, fenc_y, XO_FENC_STRIDE, fenc_c, XO_FENC_STRIDE, 8 );
                load_mb( g, recon_slot, mb_x, mb_y, fdec_y, XO_FDEC_STRIDE, fdec_c, XO_FDEC_STRIDE, 16 );
                c = xo_encode_inter_mb( fenc_y, fenc_c, fdec_y, fdec_c, prm->qp, out_levels, out_nnz );
                store_mb( g, recon_slot, mb_x, mb_y, fdec_y, fdec_c );
                cbp[xy] = (int16_t)c;
                if( type == X264DSP_MB_P_L0 && part == 16 && !( c & 0x3f ) && out_mv[0] == pskip_mv[0] && out_mv[1] == pskip_mv[1] )
                    type = X264DSP_MB_P_SKIP;
                mb_type[xy] = (int8_t)type;
                partition[xy] = (uint8_t)part;
            }
            /* ---- mvd of every partition as the CABAC writer computes it: predictions from the FINAL vectors, partitions in
             *      coding order, each seeing the ones before it (encoder/cabac.c:278-300, 352-412) */
            if( type != X264DSP_MB_P_SKIP )
            {
                int16_t fin[4][2], mvp[2];
                int fset[4] = { 0, 0, 0, 0 };
                int16_t *d = mvd8 + (size_t)xy * 8;
                memcpy( fin, out_mv, sizeof(fin) );
                if( part == 16 )
                {
                    for( k = 0; k < 4; k++ ) { d[2 * k] = (int16_t)( fin[0][0] - mvp16[0] ); d[2 * k + 1] = (int16_t)( fin[0][1] - mvp16[1] ); }
                }
                else if( part == 14 )
                    for( k = 0; k < 2; k++ )
                    {
                        predict_part( 0, 2 * k, 4, 1 + k, mb_x, mb_y, W, mv8, fin, fset, mvp );
                        d[4 * k] = d[4 * k + 2] = (int16_t)( fin[2 * k][0] - mvp[0] );
                        d[4 * k + 1] = d[4 * k + 3] = (int16_t)( fin[2 * k][1] - mvp[1] );
                        fset[2 * k] = fset[2 * k + 1] = 1;
                    }
                else if( part == 15 )
                    for( k = 0; k < 2; k++ )
                    {
                        predict_part( 2 * k, 0, 2, 3 + k, mb_x, mb_y, W, mv8, fin, fset, mvp );
                        d[2 * k] = d[2 * k + 4] = (int16_t)( fin[k][0] - mvp[0] );
                        d[2 * k + 1] = d[2 * k + 5] = (int16_t)( fin[k][1] - mvp[1] );
                        fset[k] = fset[k + 2] = 1;
                    }
                else
                    for( k = 0; k < 4; k++ )
                    {
                        predict_part( 2 * ( k & 1 ), 2 * ( k >> 1 ), 2, 0, mb_x, mb_y, W, mv8, fin, fset, mvp );
                        d[2 * k] = (int16_t)( fin[k][0] - mvp[0] );
                        d[2 * k + 1] = (int16_t)( fin[k][1] - mvp[1] );
                        fset[k] = 1;
                    }
            }
        }
}
