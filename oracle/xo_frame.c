/* xo_frame.c -- oracle: frame geometry, picture staging, border expansion, half-pel planes and
 * half-resolution planes.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * Follows common/frame.c:7-153 (layout), 198-232 (copy_picture), 363-450 (borders),
 * common/mc.c:144-167 (hpel_filter), 404-456 (lowres), 506-535 (frame_filter) and the row
 * schedule of encoder/encoder.c:1359-1385.
 */
#include <string.h>
#include <stdlib.h>
#include "xo.h"

/* common/frame.c:7-20 */
static int stride_rule( int x, int align, int disalign )
{
    x = (x + align - 1) & ~(align - 1);
    return (x & (disalign - 1)) ? x : x + align;
}
static int plane_size_rule( int x, int disalign )
{
    return (x & (disalign - 1)) ? x : x + 128;
}

void xo_geometry( int width, int height, x264dsp_geom_t *g )
{
    memset( g, 0, sizeof(*g) );
    g->width = width;
    g->height = height;
    g->mb_w = (width + 15) >> 4;
    g->mb_h = (height + 15) >> 4;
    g->mb_count = g->mb_w * g->mb_h;
    g->luma_w = g->mb_w << 4;
    g->luma_h = g->mb_h << 4;
    /* frame.c:38-53 */
    g->luma_stride = stride_rule( g->luma_w + 2*X264DSP_PADH, 16, 1 << 10 );
    g->luma_plane_size = plane_size_rule( g->luma_stride * (g->luma_h + 2*X264DSP_PADV), 1 << 10 );
    g->luma_origin = g->luma_stride * X264DSP_PADV + X264DSP_PADH;
    g->chroma_stride = g->luma_stride;
    g->chroma_h = g->luma_h >> 1;
    g->chroma_plane_size = g->chroma_stride * (g->chroma_h + X264DSP_PADV);   /* 2 * (PADV/2) rows */
    g->chroma_origin = g->chroma_stride * (X264DSP_PADV >> 1) + X264DSP_PADH;
    g->lowres_w = g->luma_w >> 1;
    g->lowres_h = g->luma_h >> 1;
    g->lowres_stride = stride_rule( g->lowres_w + 2*X264DSP_PADH, 16, 1 << 11 );
    g->lowres_plane_size = plane_size_rule( g->lowres_stride * (g->lowres_h + 2*X264DSP_PADV), 1 << 10 );
    g->lowres_origin = g->lowres_stride * X264DSP_PADV + X264DSP_PADH;
    {
        /* +64: the filtered-plane border writes 8 bytes past the last plane (frame.c:406-412) */
        int64_t off = 4 * (int64_t)g->luma_plane_size + 64;
        off = (off + 255) & ~(int64_t)255;
        g->slot_chroma_off = (int32_t)off;
        off += g->chroma_plane_size;
        off = (off + 255) & ~(int64_t)255;
        g->slot_lowres_off = (int32_t)off;
        off += 4 * (int64_t)g->lowres_plane_size;
        off = (off + 255) & ~(int64_t)255;
        /* 8x8-tiled copies of the padded lowres planes (our own search layout, include/x264dsp_b200.h) */
        g->tile_w = (g->lowres_w + 2*X264DSP_PADH) / 8;
        g->tile_h = (g->lowres_h + 2*X264DSP_PADV) / 8;
        g->tiled_plane_size = g->tile_w * g->tile_h * 64;
        g->slot_tiled_off = (int32_t)off;
        off += 4 * (int64_t)g->tiled_plane_size;
        g->slot_bytes = (off + 255) & ~(int64_t)255;
    }
}

static pixel_t *luma_plane( const x264dsp_geom_t *g, uint8_t *slot, int k )
{
    return slot + (size_t)k * g->luma_plane_size + g->luma_origin;
}
static pixel_t *chroma_plane( const x264dsp_geom_t *g, uint8_t *slot )
{
    return slot + g->slot_chroma_off + g->chroma_origin;
}
static pixel_t *lowres_plane( const x264dsp_geom_t *g, uint8_t *slot, int k )
{
    return slot + g->slot_lowres_off + (size_t)k * g->lowres_plane_size + g->lowres_origin;
}

/* x264_frame_copy_picture (I420 branch, frame.c:225-229) + x264_frame_expand_border_mod16 (423-450) */
void xo_frame_load_i420( const x264dsp_geom_t *g, const uint8_t *i420, uint8_t *slot )
{
    const int w = g->width, h = g->height, cw = w >> 1, ch = h >> 1;
    const uint8_t *sy = i420, *su = i420 + (size_t)w * h, *sv = su + (size_t)cw * ch;
    pixel_t *y = luma_plane( g, slot, 0 ), *c = chroma_plane( g, slot );
    const int ls = g->luma_stride, cs = g->chroma_stride;
    const int padx = g->luma_w - w;
    int r, x;

    for( r = 0; r < h; r++ )
        memcpy( y + (size_t)r * ls, sy + (size_t)r * w, w );
    for( r = 0; r < ch; r++ )
        for( x = 0; x < cw; x++ )
        {
            c[(size_t)r * cs + 2*x]     = su[(size_t)r * cw + x];
            c[(size_t)r * cs + 2*x + 1] = sv[(size_t)r * cw + x];
        }

    if( padx )
    {
        for( r = 0; r < h; r++ )
            memset( y + (size_t)r * ls + w, y[(size_t)r * ls + w - 1], padx );
        for( r = 0; r < ch; r++ )
            for( x = 0; x < padx; x += 2 )
            {
                c[(size_t)r * cs + w + x]     = c[(size_t)r * cs + w - 2];
                c[(size_t)r * cs + w + x + 1] = c[(size_t)r * cs + w - 1];
            }
    }
    for( r = h; r < g->luma_h; r++ )
        memcpy( y + (size_t)r * ls, y + (size_t)(h - 1) * ls, g->luma_w );
    for( r = ch; r < g->chroma_h; r++ )
        memcpy( c + (size_t)r * cs, c + (size_t)(ch - 1) * cs, g->luma_w );
}

/* plane_expand_border (frame.c:363-383).  `unit` = 1 for luma, 2 for an interleaved UV pair.
 * The row copies use memmove: the filtered-plane call passes width+2*padh > stride for the
 * common strides, so consecutive rows overlap by a few bytes, exactly as in the reference. */
static void expand_border( pixel_t *pix, int stride, int width, int height, int padh, int padv,
                           int top, int bottom, int unit )
{
    int y, x;
    for( y = 0; y < height; y++ )
    {
        pixel_t *row = pix + (ptrdiff_t)y * stride;
        pixel_t l0 = row[0], l1 = row[unit - 1];
        pixel_t r0 = row[width - unit], r1 = row[width - 1];
        for( x = 0; x < padh; x += unit )
        {
            row[-padh + x] = l0;
            row[-padh + x + unit - 1] = l1;
        }
        for( x = 0; x < padh; x += unit )
        {
            row[width + x] = r0;
            row[width + x + unit - 1] = r1;
        }
    }
    if( top )
        for( y = 0; y < padv; y++ )
            memmove( pix - padh - (ptrdiff_t)(y + 1) * stride, pix - padh, width + 2*padh );
    if( bottom )
        for( y = 0; y < padv; y++ )
            memmove( pix - padh + (ptrdiff_t)(height + y) * stride,
                     pix - padh + (ptrdiff_t)(height - 1) * stride, width + 2*padh );
}

/* x264_frame_expand_border for mb_y = 0 .. mb_h-1 (frame.c:386-396) */
void xo_frame_expand_border( const x264dsp_geom_t *g, uint8_t *slot )
{
    pixel_t *y = luma_plane( g, slot, 0 ), *c = chroma_plane( g, slot );
    int mb_y;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
    {
        int top = mb_y == 0, bot = mb_y == g->mb_h - 1;
        int rows = bot ? 20 : 16;
        int start = (mb_y << 4) - (top ? 0 : 4);
        expand_border( y + (ptrdiff_t)start * g->luma_stride, g->luma_stride, g->luma_w, rows,
                       X264DSP_PADH, X264DSP_PADV, top, bot, 1 );
        expand_border( c + (((ptrdiff_t)start * g->chroma_stride) >> 1), g->chroma_stride, g->luma_w,
                       rows >> 1, X264DSP_PADH, X264DSP_PADV >> 1, top, bot, 2 );
    }
}

static inline pixel_t clip_u8( int v )
{
    return v < 0 ? 0 : v > 255 ? 255 : (pixel_t)v;
}

/* hpel_filter (mc.c:144-167): six-tap (1,-5,20,20,-5,1).  The vertical intermediate is kept at
 * 16-bit precision for x in [-2, width+3) and the centre plane is the horizontal filter of that
 * intermediate, rounded once with (+512)>>10. */
void xo_hpel_filter( pixel_t *dsth, pixel_t *dstv, pixel_t *dstc, const pixel_t *src,
                     intptr_t stride, int width, int height )
{
    int16_t *mid = malloc( (width + 5) * sizeof(int16_t) );
    int x, y;
    for( y = 0; y < height; y++ )
    {
        const pixel_t *s = src + y * stride;
        for( x = -2; x < width + 3; x++ )
        {
            int v = s[x - 2*stride] + s[x + 3*stride]
                  - 5 * ( s[x - stride] + s[x + 2*stride] )
                  + 20 * ( s[x] + s[x + stride] );
            dstv[y*stride + x] = clip_u8( (v + 16) >> 5 );
            mid[x + 2] = (int16_t)v;
        }
        for( x = 0; x < width; x++ )
        {
            const int16_t *m = mid + x + 2;
            int c = m[-2] + m[3] - 5 * ( m[-1] + m[2] ) + 20 * ( m[0] + m[1] );
            int hsum = s[x-2] + s[x+3] - 5 * ( s[x-1] + s[x+2] ) + 20 * ( s[x] + s[x+1] );
            dstc[y*stride + x] = clip_u8( (c + 512) >> 10 );
            dsth[y*stride + x] = clip_u8( (hsum + 16) >> 5 );
        }
    }
    free( mid );
}

/* x264_frame_filter (mc.c:506-535) + x264_frame_expand_border_filtered (frame.c:398-413), called
 * for every MB row in the order of x264_fdec_filter_row (encoder.c:1359-1385).  Plane N must
 * already have its border expanded. */
void xo_frame_filter( const x264dsp_geom_t *g, uint8_t *slot )
{
    const int stride = g->luma_stride;
    pixel_t *pn = luma_plane( g, slot, 0 );
    pixel_t *ph = luma_plane( g, slot, 1 ), *pv = luma_plane( g, slot, 2 ), *pc = luma_plane( g, slot, 3 );
    int mb_y, k;
    for( mb_y = 0; mb_y < g->mb_h; mb_y++ )
    {
        int end = mb_y == g->mb_h - 1;
        int start = (mb_y << 4) - 8;
        int stop = (end ? g->luma_h : (mb_y << 4)) + 8;
        ptrdiff_t offs = (ptrdiff_t)start * stride - 8;
        xo_hpel_filter( ph + offs, pv + offs, pc + offs, pn + offs, stride, g->luma_w + 16, stop - start );
        {
            int rows = end ? ((g->mb_h - mb_y) << 4) + 16 : 16;
            ptrdiff_t o = (ptrdiff_t)start * stride;
            pixel_t *pl[3] = { ph, pv, pc };
            for( k = 0; k < 3; k++ )
                expand_border( pl[k] + o, stride, g->luma_w + 8, rows, X264DSP_PADH, X264DSP_PADV - 8,
                               mb_y == 0, end, 1 );
        }
    }
}

/* frame_init_lowres_core (mc.c:432-456) */
void xo_lowres_core( const pixel_t *src0, pixel_t *dst0, pixel_t *dsth, pixel_t *dstv, pixel_t *dstc,
                     intptr_t src_stride, intptr_t dst_stride, int width, int height )
{
    int x, y;
#define AVG2( a, b ) ( ((a) + (b) + 1) >> 1 )
    for( y = 0; y < height; y++ )
    {
        const pixel_t *r0 = src0 + 2*y*src_stride, *r1 = r0 + src_stride, *r2 = r1 + src_stride;
        for( x = 0; x < width; x++ )
        {
            int a0 = AVG2( r0[2*x], r1[2*x] ),     a1 = AVG2( r0[2*x+1], r1[2*x+1] ), a2 = AVG2( r0[2*x+2], r1[2*x+2] );
            int b0 = AVG2( r1[2*x], r2[2*x] ),     b1 = AVG2( r1[2*x+1], r2[2*x+1] ), b2 = AVG2( r1[2*x+2], r2[2*x+2] );
            dst0[y*dst_stride + x] = AVG2( a0, a1 );
            dsth[y*dst_stride + x] = AVG2( a1, a2 );
            dstv[y*dst_stride + x] = AVG2( b0, b1 );
            dstc[y*dst_stride + x] = AVG2( b1, b2 );
        }
    }
#undef AVG2
}

/* x264_frame_init_lowres (mc.c:404-419) + x264_frame_expand_border_lowres (frame.c:415-421).
 * Note the side effect on the SOURCE plane: column luma_w and row luma_h are written. */
void xo_frame_init_lowres( const x264dsp_geom_t *g, uint8_t *slot )
{
    pixel_t *src = luma_plane( g, slot, 0 );
    const int ls = g->luma_stride, w = g->luma_w, h = g->luma_h;
    int y, k;
    for( y = 0; y < h; y++ )
        src[w + (ptrdiff_t)y * ls] = src[w - 1 + (ptrdiff_t)y * ls];
    memcpy( src + (ptrdiff_t)ls * h, src + (ptrdiff_t)ls * (h - 1), w + 1 );
    xo_lowres_core( src, lowres_plane( g, slot, 0 ), lowres_plane( g, slot, 1 ), lowres_plane( g, slot, 2 ),
                    lowres_plane( g, slot, 3 ), ls, g->lowres_stride, g->lowres_w, g->lowres_h );
    for( k = 0; k < 4; k++ )
        expand_border( lowres_plane( g, slot, k ), g->lowres_stride, g->lowres_w, g->lowres_h,
                       X264DSP_PADH, X264DSP_PADV, 1, 1, 1 );
    xo_frame_retile_lowres( g, slot );
}

/* the slot's 8x8-tiled copies of the padded lowres planes (layout: include/x264dsp_b200.h; not a
 * reference structure -- the product's search layout, mirrored so that whole slots compare) */
void xo_frame_retile_lowres( const x264dsp_geom_t *g, uint8_t *slot )
{
    int k, X, Y;
    for( k = 0; k < 4; k++ )
    {
        const pixel_t *src = lowres_plane( g, slot, k );
        uint8_t *dst = slot + g->slot_tiled_off + (size_t)k * g->tiled_plane_size;
        for( Y = 0; Y < g->tile_h * 8; Y++ )
            for( X = 0; X < g->tile_w * 8; X++ )
                dst[( (size_t)( Y >> 3 ) * g->tile_w + ( X >> 3 ) ) * 64 + ( Y & 7 ) * 8 + ( X & 7 )] =
                    src[(ptrdiff_t)( Y - X264DSP_PADV ) * g->lowres_stride + X - X264DSP_PADH];
    }
}
