/* lookahead_host.c -- the hot path from plain C (the reference's own language): a host program that links against
 * libx264dsp_b200.so through include/x264dsp_b200.h and nothing else -- no CUDA, C++ or Python on this side.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/lookahead_host.c -o examples/_build/lookahead_host \
 *       -Lx264-dsp_b200 -l:libx264dsp_b200.so -Wl,-rpath,$PWD/x264-dsp_b200
 *   examples/_build/lookahead_host [width height frames]
 *
 * It runs the lowres lookahead of an n-frame synthetic clip twice:
 *   (1) the way an encoder would call it -- x264dsp_lookahead_clip_host: host pictures in, lowres MVs / MV costs /
 *       frame costs out (what x264_slicetype_frame_cost leaves in fenc->lowres_mvs, lowres_mv_costs, i_cost_est,
 *       encoder/slicetype.c:223-322);
 *   (2) step by step on device memory with the frame-batched entry points -- x264dsp_frame_load_luma_dev,
 *       x264dsp_frame_init_lowres_dev (x264_frame_init_lowres, common/mc.c:404), x264dsp_lookahead_frame_cost_dev --
 * checks that both give the same numbers and prints one JSON line with the per-frame costs (tests/test_gpu_c_host.py
 * compares them with the CPU oracle).  Exit code 0 = consistent. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "x264dsp_b200.h"

#define CHECK( call ) do { int rc_ = ( call ); if( rc_ ) { fprintf( stderr, "%s -> %d\n", #call, rc_ ); return 2; } } while( 0 )

int main( int argc, char **argv )
{
    const int w = argc > 2 ? atoi( argv[1] ) : 352, h = argc > 2 ? atoi( argv[2] ) : 288;
    const int n = argc > 3 ? atoi( argv[3] ) : 6;
    x264dsp_ctx_t *ctx = NULL;
    x264dsp_geom_t g;
    int i, k, same = 1;

    CHECK( x264dsp_create( 0, &ctx ) );
    CHECK( x264dsp_geometry( w, h, &g ) );

    /* synthetic pictures (luma is all the lookahead reads); pinned, so the library copies straight out of them */
    uint8_t *luma, *chroma = malloc( (size_t)w * h / 2 );
    int16_t *mvs;
    int32_t *costs, *sums;
    CHECK( x264dsp_host_alloc( ctx, (size_t)n * w * h, (void **)&luma ) );
    CHECK( x264dsp_host_alloc( ctx, (size_t)n * g.mb_count * 2 * sizeof(int16_t), (void **)&mvs ) );
    CHECK( x264dsp_host_alloc( ctx, (size_t)n * g.mb_count * sizeof(int32_t), (void **)&costs ) );
    CHECK( x264dsp_host_alloc( ctx, (size_t)n * X264DSP_LA_SUMS * sizeof(int32_t), (void **)&sums ) );
    for( i = 0; i < n; i++ )
        CHECK( x264dsp_synth_frame( w, h, i, -1, luma + (size_t)i * w * h, chroma, chroma + (size_t)w * h / 4 ) );

    /* (1) one call, host to host */
    CHECK( x264dsp_lookahead_clip_host( ctx, w, h, n, luma, mvs, costs, sums ) );

    /* (2) the same on device memory, step by step */
    uint8_t *d_luma, *d_slots;
    int16_t *d_mvs;
    int32_t *d_costs, *d_sums;
    int32_t *b = malloc( n * sizeof(int32_t) ), *p0 = malloc( n * sizeof(int32_t) );
    uint8_t *want_intra = malloc( n );
    int16_t *mvs2 = malloc( (size_t)n * g.mb_count * 2 * sizeof(int16_t) );
    int32_t *costs2 = malloc( (size_t)n * g.mb_count * sizeof(int32_t) ), *sums2 = malloc( (size_t)n * X264DSP_LA_SUMS * sizeof(int32_t) );
    CHECK( x264dsp_dev_alloc( ctx, (size_t)n * w * h, (void **)&d_luma ) );
    CHECK( x264dsp_dev_alloc( ctx, (size_t)n * g.slot_bytes, (void **)&d_slots ) );
    CHECK( x264dsp_dev_alloc( ctx, (size_t)n * g.mb_count * 2 * sizeof(int16_t), (void **)&d_mvs ) );
    CHECK( x264dsp_dev_alloc( ctx, (size_t)n * g.mb_count * sizeof(int32_t), (void **)&d_costs ) );
    CHECK( x264dsp_dev_alloc( ctx, (size_t)n * X264DSP_LA_SUMS * sizeof(int32_t), (void **)&d_sums ) );
    CHECK( x264dsp_dev_zero( ctx, d_slots, (size_t)n * g.slot_bytes, NULL ) );
    CHECK( x264dsp_dev_zero( ctx, d_mvs, (size_t)n * g.mb_count * 2 * sizeof(int16_t), NULL ) );
    CHECK( x264dsp_dev_zero( ctx, d_costs, (size_t)n * g.mb_count * sizeof(int32_t), NULL ) );
    CHECK( x264dsp_h2d( ctx, d_luma, luma, (size_t)n * w * h, NULL ) );
    CHECK( x264dsp_frame_load_luma_dev( ctx, &g, d_luma, d_slots, n, NULL ) );      /* x264_frame_copy_picture + mod16 padding */
    CHECK( x264dsp_frame_init_lowres_dev( ctx, &g, d_slots, n, NULL ) );            /* x264_frame_init_lowres */
    for( i = 0; i < n; i++ )
    {
        b[i] = i;
        p0[i] = i ? i - 1 : -1;                                                     /* frame 0: intra only */
        want_intra[i] = 1;
    }
    CHECK( x264dsp_lookahead_frame_cost_dev( ctx, &g, d_slots, n, b, p0, want_intra, d_mvs, d_costs, d_sums, NULL, NULL ) );
    CHECK( x264dsp_d2h( ctx, mvs2, d_mvs, (size_t)n * g.mb_count * 2 * sizeof(int16_t), NULL ) );
    CHECK( x264dsp_d2h( ctx, costs2, d_costs, (size_t)n * g.mb_count * sizeof(int32_t), NULL ) );
    CHECK( x264dsp_d2h( ctx, sums2, d_sums, (size_t)n * X264DSP_LA_SUMS * sizeof(int32_t), NULL ) );
    CHECK( x264dsp_sync( ctx ) );

    for( i = 0; i < n; i++ )
    {
        for( k = 0; k < 3; k++ )
            same &= sums[i * X264DSP_LA_SUMS + k] == sums2[i * X264DSP_LA_SUMS + k] || ( !i && k != X264DSP_LA_COST_INTRA );
        if( i )
        {
            same &= !memcmp( mvs + (size_t)i * g.mb_count * 2, mvs2 + (size_t)i * g.mb_count * 2, g.mb_count * 2 * sizeof(int16_t) );
            same &= !memcmp( costs + (size_t)i * g.mb_count, costs2 + (size_t)i * g.mb_count, g.mb_count * sizeof(int32_t) );
        }
    }

    printf( "{\"version\": \"%s\", \"width\": %d, \"height\": %d, \"frames\": %d, \"consistent\": %s, \"launches\": %lld, \"frames_out\": [",
            x264dsp_version(), w, h, n, same ? "true" : "false", (long long)x264dsp_launch_count( ctx ) );
    for( i = 0; i < n; i++ )
    {
        long long mvsum = 0;
        for( k = 0; k < g.mb_count * 2; k++ )
            mvsum += abs( mvs[(size_t)i * g.mb_count * 2 + k] );
        printf( "%s{\"cost_inter\": %d, \"cost_intra\": %d, \"intra_mbs\": %d, \"mv_abs_sum\": %lld}", i ? ", " : "",
                i ? sums[i * X264DSP_LA_SUMS + X264DSP_LA_COST_INTER] : -1, sums[i * X264DSP_LA_SUMS + X264DSP_LA_COST_INTRA],
                i ? sums[i * X264DSP_LA_SUMS + X264DSP_LA_INTRA_MBS] : -1, i ? mvsum : 0 );
    }
    printf( "]}\n" );

    x264dsp_dev_free( ctx, d_luma ); x264dsp_dev_free( ctx, d_slots ); x264dsp_dev_free( ctx, d_mvs );
    x264dsp_dev_free( ctx, d_costs ); x264dsp_dev_free( ctx, d_sums );
    x264dsp_host_free( ctx, luma ); x264dsp_host_free( ctx, mvs ); x264dsp_host_free( ctx, costs ); x264dsp_host_free( ctx, sums );
    free( chroma ); free( b ); free( p0 ); free( want_intra ); free( mvs2 ); free( costs2 ); free( sums2 );
    x264dsp_destroy( ctx );
    return same ? 0 : 1;
}
