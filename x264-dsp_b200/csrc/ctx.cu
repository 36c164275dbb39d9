// ctx.cu -- context, geometry, constant tables, memory helpers of libx264dsp_b200.so.
//
// Geometry follows common/frame.c:7-57, 77-97, 126-133; the cost tables follow
// encoder/analyse.c:98-111, 171-206, 243-315; the flat-CQM quant tables common/set.c:265-353;
// the chroma QP map common/macroblock.h:251-266.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

extern "C" const char *x264dsp_version( void ) { return "x264dsp_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------ geometry

static int xd_align_stride( int x, int align, int disalign )
{
    x = ( x + align - 1 ) / align * align;
    if( x % disalign == 0 )
        x += align;
    return x;
}

extern "C" int x264dsp_geometry( int width, int height, x264dsp_geom_t *g )
{
    if( !g || width < 16 || height < 16 || ( width & 1 ) || ( height & 1 ) || width > 16384 || height > 16384 )
        return X264DSP_E_ARG;
    memset( g, 0, sizeof( *g ) );
    g->width = width;
    g->height = height;
    g->mb_w = ( width + 15 ) / 16;
    g->mb_h = ( height + 15 ) / 16;
    g->mb_count = g->mb_w * g->mb_h;
    g->luma_w = g->mb_w * 16;
    g->luma_h = g->mb_h * 16;

    g->luma_stride = xd_align_stride( g->luma_w + 2 * X264DSP_PADH, 16, 1024 );
    int64_t lp = (int64_t)g->luma_stride * ( g->luma_h + 2 * X264DSP_PADV );
    if( lp % 1024 == 0 )
        lp += 128;
    g->luma_plane_size = (int32_t)lp;
    g->luma_origin = g->luma_stride * X264DSP_PADV + X264DSP_PADH;

    g->chroma_stride = g->luma_stride;
    g->chroma_h = g->luma_h / 2;
    g->chroma_plane_size = g->chroma_stride * ( g->chroma_h + 2 * ( X264DSP_PADV / 2 ) );
    g->chroma_origin = g->chroma_stride * ( X264DSP_PADV / 2 ) + X264DSP_PADH;

    g->lowres_w = g->luma_w / 2;
    g->lowres_h = g->luma_h / 2;
    g->lowres_stride = xd_align_stride( g->lowres_w + 2 * X264DSP_PADH, 16, 2048 );
    int64_t wp = (int64_t)g->lowres_stride * ( g->lowres_h + 2 * X264DSP_PADV );
    if( wp % 1024 == 0 )
        wp += 128;
    g->lowres_plane_size = (int32_t)wp;
    g->lowres_origin = g->lowres_stride * X264DSP_PADV + X264DSP_PADH;

    // +64: x264_frame_expand_border_filtered writes 8 bytes past the last plane (frame.c:406-412)
    int64_t off = 4 * lp + 64;
    off = ( off + 255 ) / 256 * 256;
    g->slot_chroma_off = (int32_t)off;
    off += g->chroma_plane_size;
    off = ( off + 255 ) / 256 * 256;
    g->slot_lowres_off = (int32_t)off;
    off += 4 * wp;
    off = ( off + 255 ) / 256 * 256;
    g->tile_w = ( g->lowres_w + 2 * X264DSP_PADH ) / 8;
    g->tile_h = ( g->lowres_h + 2 * X264DSP_PADV ) / 8;
    g->tiled_plane_size = g->tile_w * g->tile_h * 64;
    g->slot_tiled_off = (int32_t)off;
    off += 4 * (int64_t)g->tiled_plane_size;
    g->slot_bytes = ( off + 255 ) / 256 * 256;
    return 0;
}

// ------------------------------------------------------------------------------------ sharding

extern "C" int x264dsp_frame_range( int n_frames, int rank, int world, int *first, int *count, int *need_prev )
{
    if( n_frames < 0 || world < 1 || rank < 0 || rank >= world || !first || !count || !need_prev )
        return X264DSP_E_ARG;
    const int lo = (int)( (int64_t)n_frames * rank / world ), hi = (int)( (int64_t)n_frames * ( rank + 1 ) / world );
    *first = lo;
    *count = hi - lo;
    *need_prev = hi > lo && lo > 0;
    return 0;
}

// ------------------------------------------------------------------------------------ tables

static const uint16_t xd_lambda_tab[52] =
{
     1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  4,  4,  4,
     5,  6,  6,  7,  8,  9, 10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57, 64, 72, 81, 91
};

extern "C" int x264dsp_lambda( int qp )
{
    return ( qp < 0 || qp > 51 ) ? X264DSP_E_ARG : xd_lambda_tab[qp];
}

// number of bits charged for an mv delta of |d| quarter-pels: (int)(2*log2(d+1) + 1.718 + .5) for
// d >= 1, expressed as the reference does, by the first |d| of every bit count
static int xd_mv_bits( int d )
{
    static const uint16_t first[24] = { 1, 2, 3, 5, 7, 10, 14, 20, 29, 41, 59, 83, 118, 167, 237, 335,
                                        474, 671, 949, 1342, 1898, 2685, 3797, 4097 };
    int k = 0;
    while( d >= first[k + 1] )
        k++;
    return 4 + k;
}

extern "C" int x264dsp_cost_mv_table( int qp, uint16_t *out )
{
    if( qp < 0 || qp > 51 || !out )
        return X264DSP_E_ARG;
    const int lambda = xd_lambda_tab[qp];
    out[4096] = (uint16_t)lambda;
    for( int d = 1; d <= 4096; d++ )
        out[4096 + d] = out[4096 - d] = (uint16_t)( lambda * xd_mv_bits( d ) );
    return 0;
}

extern "C" int x264dsp_quant_tables( int b_inter, int qp, uint16_t *mf, uint16_t *bias )
{
    static const uint16_t scale[6][3] =
    {
        { 13107, 8066, 5243 }, { 11916, 7490, 4660 }, { 10082, 6554, 4194 },
        {  9362, 5825, 3647 }, {  8192, 5243, 3355 }, {  7282, 4559, 2893 }
    };
    if( qp < 0 || qp > 51 || !mf || !bias )
        return X264DSP_E_ARG;
    const int deadzone = ( b_inter ? 11 : 21 ) << 10;
    const int s = qp / 6 - 1;
    for( int i = 0; i < 16; i++ )
    {
        int v = scale[qp % 6][( i & 1 ) + ( ( i >> 2 ) & 1 )];
        v = s < 0 ? v << 1 : s == 0 ? v : ( v + ( 1 << ( s - 1 ) ) ) >> s;
        mf[i] = (uint16_t)v;
        int near = ( deadzone + ( v >> 1 ) ) / v, cap = 32768 / v;
        bias[i] = (uint16_t)( near < cap ? near : cap );
    }
    return 0;
}

extern "C" int x264dsp_dequant_table( int out[6][16] )
{
    static const uint8_t scale[6][3] =
    {
        { 10, 13, 16 }, { 11, 14, 18 }, { 13, 16, 20 }, { 14, 18, 23 }, { 16, 20, 25 }, { 18, 23, 29 }
    };
    if( !out )
        return X264DSP_E_ARG;
    for( int q = 0; q < 6; q++ )
        for( int i = 0; i < 16; i++ )
            out[q][i] = 16 * scale[q][( i & 1 ) + ( ( i >> 2 ) & 1 )];
    return 0;
}

extern "C" int x264dsp_chroma_qp( int qp )
{
    static const uint8_t tail[22] = { 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };
    if( qp < 0 )
        return 0;
    return qp < 30 ? qp : qp <= 51 ? tail[qp - 30] : 39;
}

// ------------------------------------------------------------------------------------ memory

int xd_reserve_dev( void **p, size_t *cap, size_t bytes )
{
    if( *cap >= bytes )
        return 0;
    if( *p )
        cudaFree( *p );
    *p = NULL;
    *cap = 0;
    size_t want = bytes + bytes / 4;
    XD_CHECK( cudaMalloc( p, want ) );
    *cap = want;
    return 0;
}

int xd_reserve_pinned( void **p, size_t *cap, size_t bytes )
{
    if( *cap >= bytes )
        return 0;
    if( *p )
        cudaFreeHost( *p );
    *p = NULL;
    *cap = 0;
    size_t want = bytes + bytes / 4;
    XD_CHECK( cudaHostAlloc( p, want, cudaHostAllocDefault ) );
    *cap = want;
    return 0;
}

// every failure after the calloc goes through x264dsp_destroy, which copes with a half-built context
#define XD_CREATE_CHECK( call ) \
    do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { free( host ); x264dsp_destroy( ctx ); cudaGetLastError(); return (int)e_; } } while( 0 )

extern "C" int x264dsp_create( int device, x264dsp_ctx_t **out )
{
    if( !out )
        return X264DSP_E_ARG;
    *out = NULL;
    int count = 0;
    if( cudaGetDeviceCount( &count ) != cudaSuccess || count <= 0 || device < 0 || device >= count )
    {
        cudaGetLastError();
        return X264DSP_E_NOGPU;            // no CPU path exists: fail loudly
    }
    XD_CHECK( cudaSetDevice( device ) );
    x264dsp_ctx *ctx = (x264dsp_ctx *)calloc( 1, sizeof( x264dsp_ctx ) );
    if( !ctx )
        return X264DSP_E_NOMEM;
    uint16_t *host = NULL;
    ctx->device = device;
    cudaDeviceProp prop;
    XD_CREATE_CHECK( cudaGetDeviceProperties( &prop, device ) );
    ctx->sm_count = prop.multiProcessorCount;
    XD_CREATE_CHECK( cudaStreamCreateWithFlags( &ctx->stream, cudaStreamNonBlocking ) );
    for( int i = 0; i < XD_AUX_STREAMS; i++ )
        XD_CREATE_CHECK( cudaStreamCreateWithFlags( &ctx->aux[i], cudaStreamNonBlocking ) );

    // one cost table per distinct lambda, shared between the QPs that map to it
    {
        host = (uint16_t *)malloc( 52 * 8193 * sizeof( uint16_t ) );
        if( !host )
        {
            x264dsp_destroy( ctx );
            return X264DSP_E_NOMEM;
        }
        int n_tables = 0;
        int first_qp_of_table[52];
        int table_of_qp[52];
        for( int qp = 0; qp < 52; qp++ )
        {
            if( qp > 0 && xd_lambda_tab[qp] == xd_lambda_tab[qp - 1] )
                table_of_qp[qp] = table_of_qp[qp - 1];
            else
            {
                first_qp_of_table[n_tables] = qp;
                table_of_qp[qp] = n_tables++;
            }
        }
        for( int t = 0; t < n_tables; t++ )
            x264dsp_cost_mv_table( first_qp_of_table[t], host + (size_t)t * 8193 );
        // pad each table to 8200 entries so that every table starts 16-byte aligned
        XD_CREATE_CHECK( cudaMalloc( (void **)&ctx->cost_mv_store, (size_t)n_tables * 8200 * sizeof( uint16_t ) ) );
        for( int t = 0; t < n_tables; t++ )
            XD_CREATE_CHECK( cudaMemcpy( ctx->cost_mv_store + (size_t)t * 8200, host + (size_t)t * 8193,
                                         8193 * sizeof( uint16_t ), cudaMemcpyHostToDevice ) );
        for( int qp = 0; qp < 52; qp++ )
            ctx->cost_mv_dev[qp] = ctx->cost_mv_store + (size_t)table_of_qp[qp] * 8200;
        free( host );
        host = NULL;
    }
    XD_CREATE_CHECK( cudaMalloc( (void **)&ctx->la_ticket, 64 * sizeof( int32_t ) ) );
    XD_CREATE_CHECK( cudaMemset( ctx->la_ticket, 0, 64 * sizeof( int32_t ) ) );
    ctx->la_epoch = 0;
    *out = ctx;
    return 0;
}

// the DEVICE copy of cost_mv[qp] read back into host memory (what the search kernels index), for tests
extern "C" int x264dsp_cost_mv_table_dev( x264dsp_ctx_t *ctx, int qp, uint16_t *out8193 )
{
    if( !ctx || !out8193 || qp < 0 || qp > 51 )
        return X264DSP_E_ARG;
    XD_CHECK( cudaSetDevice( ctx->device ) );
    XD_CHECK( cudaMemcpy( out8193, ctx->cost_mv_dev[qp], 8193 * sizeof( uint16_t ), cudaMemcpyDeviceToHost ) );
    return 0;
}

extern "C" void x264dsp_destroy( x264dsp_ctx_t *ctx )
{
    if( !ctx )
        return;
    cudaSetDevice( ctx->device );
    if( ctx->stream )
        cudaStreamSynchronize( ctx->stream );
    cudaFree( ctx->cost_mv_store );
    cudaFree( ctx->la_sync );
    cudaFree( ctx->la_icost );
    cudaFree( ctx->la_ticket );
    cudaFree( ctx->stage_dev );
    cudaFree( ctx->clip_slots );
    cudaFree( ctx->clip_out );
    cudaFree( ctx->db_progress );
    cudaFree( ctx->me_blocks );
    cudaFree( ctx->me_results );
    if( ctx->host_ev )
        cudaEventDestroy( ctx->host_ev );
    cudaFree( ctx->clip_desc );
    free( ctx->desc_cache );
    cudaFree( ctx->shim_dev );
    if( ctx->stage_host ) cudaFreeHost( ctx->stage_host );
    if( ctx->clip_out_host ) cudaFreeHost( ctx->clip_out_host );
    if( ctx->gc_scratch ) cudaFree( ctx->gc_scratch );
    if( ctx->if_weights ) cudaFree( ctx->if_weights );
    if( ctx->shim_host ) cudaFreeHost( ctx->shim_host );
    for( int k = 0; k < XD_PROF_KINDS; k++ )
        for( int i = 0; i < XD_PROF_MAX; i++ )
            if( ctx->prof_ev[k][i][0] )
            {
                cudaEventDestroy( ctx->prof_ev[k][i][0] );
                cudaEventDestroy( ctx->prof_ev[k][i][1] );
            }
    for( int i = 0; i < 2; i++ )
        if( ctx->scratch_ev[i] )
            cudaEventDestroy( ctx->scratch_ev[i] );
    for( int i = 0; i < XD_AUX_STREAMS; i++ )
        if( ctx->aux[i] )
            cudaStreamDestroy( ctx->aux[i] );
    if( ctx->stream )
        cudaStreamDestroy( ctx->stream );
    free( ctx );
}

extern "C" void *x264dsp_stream( x264dsp_ctx_t *ctx ) { return ctx ? (void *)ctx->stream : NULL; }

extern "C" int x264dsp_sync( x264dsp_ctx_t *ctx )
{
    if( !ctx )
        return X264DSP_E_ARG;
    XD_CHECK( cudaStreamSynchronize( ctx->stream ) );
    return 0;
}

extern "C" int64_t x264dsp_launch_count( const x264dsp_ctx_t *ctx ) { return ctx ? ctx->launches : 0; }

extern "C" int x264dsp_dev_alloc( x264dsp_ctx_t *ctx, size_t bytes, void **dev )
{
    if( !ctx || !dev )
        return X264DSP_E_ARG;
    XD_CHECK( cudaMalloc( dev, bytes ) );
    return 0;
}

extern "C" int x264dsp_dev_free( x264dsp_ctx_t *ctx, void *dev )
{
    if( !ctx )
        return X264DSP_E_ARG;
    XD_CHECK( cudaFree( dev ) );
    return 0;
}

extern "C" int x264dsp_dev_zero( x264dsp_ctx_t *ctx, void *dev, size_t bytes, void *stream )
{
    if( !ctx || !dev )
        return X264DSP_E_ARG;
    XD_CHECK( cudaMemsetAsync( dev, 0, bytes, xd_stream( ctx, stream ) ) );
    return 0;
}

extern "C" int x264dsp_h2d( x264dsp_ctx_t *ctx, void *dev, const void *host, size_t bytes, void *stream )
{
    if( !ctx || !dev || !host )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    XD_CHECK( cudaMemcpyAsync( dev, host, bytes, cudaMemcpyHostToDevice, s ) );
    XD_CHECK( cudaStreamSynchronize( s ) );
    return 0;
}

extern "C" int x264dsp_d2h( x264dsp_ctx_t *ctx, void *host, const void *dev, size_t bytes, void *stream )
{
    if( !ctx || !dev || !host )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    XD_CHECK( cudaMemcpyAsync( host, dev, bytes, cudaMemcpyDeviceToHost, s ) );
    XD_CHECK( cudaStreamSynchronize( s ) );
    return 0;
}

// pinned host memory for callers that want the host entry points to copy straight from / to it
extern "C" int x264dsp_host_alloc( x264dsp_ctx_t *ctx, size_t bytes, void **host )
{
    if( !ctx || !host )
        return X264DSP_E_ARG;
    XD_CHECK( cudaHostAlloc( host, bytes, cudaHostAllocDefault ) );
    return 0;
}

extern "C" int x264dsp_host_free( x264dsp_ctx_t *ctx, void *host )
{
    if( !ctx )
        return X264DSP_E_ARG;
    XD_CHECK( cudaFreeHost( host ) );
    return 0;
}

// page-lock memory the caller already owns (an encoder's frame buffers), so that copies from / to it run at the link's rate
extern "C" int x264dsp_host_register( x264dsp_ctx_t *ctx, void *host, size_t bytes )
{
    if( !ctx || !host || !bytes )
        return X264DSP_E_ARG;
    XD_CHECK( cudaSetDevice( ctx->device ) );
    const cudaError_t e = cudaHostRegister( host, bytes, cudaHostRegisterDefault );
    if( e != cudaSuccess )
    {
        (void)cudaGetLastError();                     // e.g. a page shared with an earlier registration: not sticky
        return (int)e;
    }
    return 0;
}

extern "C" int x264dsp_host_unregister( x264dsp_ctx_t *ctx, void *host )
{
    if( !ctx || !host )
        return X264DSP_E_ARG;
    XD_CHECK( cudaHostUnregister( host ) );
    return 0;
}

// ------------------------------------------------------------------------------------ profiling

extern "C" int x264dsp_profile_enable( x264dsp_ctx_t *ctx, int on )
{
    if( !ctx )
        return X264DSP_E_ARG;
    ctx->prof_on = on != 0;
    for( int k = 0; k < XD_PROF_KINDS; k++ )
        ctx->prof_n[k] = 0;
    return 0;
}

extern "C" int x264dsp_profile_read( x264dsp_ctx_t *ctx, int kind, double *total_ms, int *count )
{
    if( !ctx || kind < 0 || kind >= XD_PROF_KINDS || !total_ms || !count )
        return X264DSP_E_ARG;
    XD_CHECK( cudaDeviceSynchronize() );
    double acc = 0;
    for( int i = 0; i < ctx->prof_n[kind]; i++ )
    {
        float ms = 0;
        XD_CHECK( cudaEventElapsedTime( &ms, ctx->prof_ev[kind][i][0], ctx->prof_ev[kind][i][1] ) );
        acc += ms;
    }
    *total_ms = acc;
    *count = ctx->prof_n[kind];
    return 0;
}
