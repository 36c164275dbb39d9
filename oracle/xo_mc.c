/* xo_mc.c -- oracle: quarter-pel luma fetch, chroma bilinear MC, per-frame MC.
 * TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * common/mc.c:74-103 (pixel_avg, mc_copy), 192-264 (hpel_ref tables, mc_luma, get_ref),
 * 290-323 (mc_chroma); common/macroblock.c:8-28 (x264_mb_mc_xywh, 16x16 case).
 */
#include <string.h>
#include "xo.h"

/* Which of the four planes (0 N, 1 H, 2 V, 3 HV) serves quarter-pel phase (fy*4+fx): the first
 * plane always, the second only when the phase is a true quarter position (mc.c:192-193). */
static const uint8_t first_plane[16]  = { 0,1,1,1, 0,1,1,1, 2,3,3,3, 0,1,1,1 };
static const uint8_t second_plane[16] = { 0,0,0,0, 2,2,3,2, 2,2,3,2, 2,2,3,2 };

typedef struct
{
    const pixel_t *a, *b;    /* b == NULL: plain copy of a */
} qpel_src_t;

static qpel_src_t qpel_sources( const pixel_t *const src[4], intptr_t stride, int mvx, int mvy )
{
    qpel_src_t r;
    int fx = mvx & 3, fy = mvy & 3, phase = fy * 4 + fx;
    intptr_t base = (intptr_t)(mvy >> 2) * stride + (mvx >> 2);
    r.a = src[first_plane[phase]] + base + (fy == 3 ? stride : 0);
    r.b = (phase & 5) ? src[second_plane[phase]] + base + (fx == 3 ? 1 : 0) : NULL;
    return r;
}

static void average_block( pixel_t *dst, intptr_t ds, const pixel_t *a, const pixel_t *b,
                           intptr_t ss, int w, int h )
{
    int x, y;
    for( y = 0; y < h; y++, dst += ds, a += ss, b += ss )
        for( x = 0; x < w; x++ )
            dst[x] = (pixel_t)( (a[x] + b[x] + 1) >> 1 );
}

void xo_mc_luma( pixel_t *dst, intptr_t dst_stride, const pixel_t *const src[4], intptr_t src_stride,
                 int mvx, int mvy, int w, int h )
{
    qpel_src_t s = qpel_sources( src, src_stride, mvx, mvy );
    int y;
    if( s.b )
        average_block( dst, dst_stride, s.a, s.b, src_stride, w, h );
    else
        for( y = 0; y < h; y++ )
            memcpy( dst + y*dst_stride, s.a + y*src_stride, w );
}

const pixel_t *xo_get_ref( pixel_t *dst, intptr_t *dst_stride, const pixel_t *const src[4],
                           intptr_t src_stride, int mvx, int mvy, int w, int h )
{
    qpel_src_t s = qpel_sources( src, src_stride, mvx, mvy );
    if( s.b )
    {
        average_block( dst, *dst_stride, s.a, s.b, src_stride, w, h );
        return dst;
    }
    *dst_stride = src_stride;
    return s.a;
}

/* mc.c:290-323: eighth-pel bilinear on interleaved UV, weights (8-dx)(8-dy) .. dx*dy, (+32)>>6 */
void xo_mc_chroma( pixel_t *dstu, pixel_t *dstv, intptr_t dst_stride, const pixel_t *src,
                   intptr_t src_stride, int mvx, int mvy, int w, int h )
{
    int dx = mvx & 7, dy = mvy & 7, x, y;
    int w00 = (8 - dx) * (8 - dy), w01 = dx * (8 - dy), w10 = (8 - dx) * dy, w11 = dx * dy;
    const pixel_t *r0 = src + (intptr_t)(mvy >> 3) * src_stride + (mvx >> 3) * 2;
    for( y = 0; y < h; y++, r0 += src_stride, dstu += dst_stride, dstv += dst_stride )
    {
        const pixel_t *r1 = r0 + src_stride;
        for( x = 0; x < w; x++ )
        {
            dstu[x] = (pixel_t)( (w00*r0[2*x]   + w01*r0[2*x+2] + w10*r1[2*x]   + w11*r1[2*x+2] + 32) >> 6 );
            dstv[x] = (pixel_t)( (w00*r0[2*x+1] + w01*r0[2*x+3] + w10*r1[2*x+1] + w11*r1[2*x+3] + 32) >> 6 );
        }
    }
}

/* Prediction frame for P_L0 16x16 macroblocks: x264_mb_mc_xywh(h,0,0,4,4) per MB
 * (common/macroblock.c:8-28) with the MV clipped to h->mb.mv_min/max as analyse.c:378-390 sets
 * them ((-16*mb_x - 24)*4 .. (16*(mb_w-mb_x-1) + 24)*4, likewise in y).
 * mv: [mb][2] quarter-pel.  Writes luma plane N and the NV12 chroma plane of pred_slot. */
void xo_mc_frame( const x264dsp_geom_t *g, const uint8_t *fref_slot, const int16_t *mv, uint8_t *pred_slot )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, k, x, y;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
        for( mb_x = 0; mb_x < g->mb_w; mb_x++ )
        {
            int i = mb_y * g->mb_w + mb_x;
            int min_x = (-(mb_x << 4) - 24) << 2, max_x = (((g->mb_w - mb_x - 1) << 4) + 24) << 2;
            int min_y = (-(mb_y << 4) - 24) << 2, max_y = (((g->mb_h - mb_y - 1) << 4) + 24) << 2;
            int mvx = mv[2*i], mvy = mv[2*i+1];
            const pixel_t *src[4];
            pixel_t *dy = pred_slot + g->luma_origin + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            pixel_t *dc = pred_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4);
            pixel_t u[8*8], v[8*8];
            mvx = mvx < min_x ? min_x : mvx > max_x ? max_x : mvx;
            mvy = mvy < min_y ? min_y : mvy > max_y ? max_y : mvy;
            for( k = 0; k < 4; k++ )
                src[k] = fref_slot + (size_t)k * g->luma_plane_size + g->luma_origin
                       + (ptrdiff_t)(mb_y << 4) * ls + (mb_x << 4);
            xo_mc_luma( dy, ls, src, ls, mvx, mvy, 16, 16 );
            xo_mc_chroma( u, v, 8, fref_slot + g->slot_chroma_off + g->chroma_origin
                          + (ptrdiff_t)(mb_y << 3) * cs + (mb_x << 4), cs, mvx, mvy, 8, 8 );
            for( y = 0; y < 8; y++ )
                for( x = 0; x < 8; x++ )
                {
                    dc[y*cs + 2*x]     = u[y*8 + x];
                    dc[y*cs + 2*x + 1] = v[y*8 + x];
                }
        }
}

/* x264_mb_mc for every partition the reference analyses (common/macroblock.c:8-48), the macroblock's MVs written out
 * per 8x8 block (mv: [mb][4][2], raster order): x264_mb_mc_xywh( x, y, 2, 2 ) four times -- the 16x16 / 16x8 / 8x16
 * cases produce the same samples 8x8-wise, mc_luma and mc_chroma being per-pixel rules. */
void xo_mc_frame_part( const x264dsp_geom_t *g, const uint8_t *fref_slot, const int16_t *mv, uint8_t *pred_slot )
{
    const int ls = g->luma_stride, cs = g->chroma_stride;
    int mb_x, mb_y, k, x, y, p;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
        for( mb_x = 0; mb_x < g->mb_w; mb_x++ )
        {
            int i = mb_y * g->mb_w + mb_x;
            int min_x = (-(mb_x << 4) - 24) << 2, max_x = (((g->mb_w - mb_x - 1) << 4) + 24) << 2;
            int min_y = (-(mb_y << 4) - 24) << 2, max_y = (((g->mb_h - mb_y - 1) << 4) + 24) << 2;
            for( p = 0; p < 4; p++ )
            {
                const int px = (p & 1) * 8, py = (p >> 1) * 8;
                int mvx = mv[2*(4*i + p)], mvy = mv[2*(4*i + p) + 1];
                const pixel_t *src[4];
                pixel_t *dy = pred_slot + g->luma_origin + (ptrdiff_t)((mb_y << 4) + py) * ls + (mb_x << 4) + px;
                pixel_t *dc = pred_slot + g->slot_chroma_off + g->chroma_origin + (ptrdiff_t)((mb_y << 3) + py/2) * cs + (mb_x << 4) + px;
                pixel_t u[4*4], v[4*4];
                mvx = mvx < min_x ? min_x : mvx > max_x ? max_x : mvx;
                mvy = mvy < min_y ? min_y : mvy > max_y ? max_y : mvy;
                for( k = 0; k < 4; k++ )
                    src[k] = fref_slot + (size_t)k * g->luma_plane_size + g->luma_origin
                           + (ptrdiff_t)((mb_y << 4) + py) * ls + (mb_x << 4) + px;
                xo_mc_luma( dy, ls, src, ls, mvx, mvy, 8, 8 );
                xo_mc_chroma( u, v, 4, fref_slot + g->slot_chroma_off + g->chroma_origin
                              + (ptrdiff_t)((mb_y << 3) + py/2) * cs + (mb_x << 4) + px, cs, mvx, mvy, 4, 4 );
                for( y = 0; y < 4; y++ )
                    for( x = 0; x < 4; x++ )
                    {
                        dc[y*cs + 2*x]     = u[y*4 + x];
                        dc[y*cs + 2*x + 1] = v[y*4 + x];
                    }
            }
        }
}
