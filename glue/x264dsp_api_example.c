/* x264dsp_api_example.c -- an APPLICATION of the reference's library API (x264.h only), linked like x264ref_gpu with the
 * reference objects, the three glue files and libx264dsp_b200.so: what a user who calls x264_encoder_open /
 * x264_encoder_encode directly gets.  The reference CLI has no switch for x264_param_t.analyse.inter (common/common.c:106
 * fixes it to 0); an API user sets it, and X264_ANALYSE_PSUB16x16 sends every P slice through x264dsp_p_frames_part_dev.
 *
 *   x264api_gpu in.yuv out.264 WIDTH HEIGHT [psub16x16=0] [me=1] [subme=5] [qp=26]
 *
 * X264DSP_GLUE=0 in the environment leaves the doors closed: the same binary is then the reference library alone, which
 * is what tests/test_gpu_glue_cli.py compares the device run with, byte for byte. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "x264.h"

static void quiet( void *p, int level, const char *fmt, va_list ap ) { (void)p; (void)level; (void)fmt; (void)ap; }

int main( int argc, char **argv )
{
    if( argc < 5 )
    {
        fprintf( stderr, "usage: %s in.yuv out.264 WIDTH HEIGHT [psub16x16] [me] [subme] [qp]\n", argv[0] );
        return 2;
    }
    const int w = atoi( argv[3] ), ht = atoi( argv[4] );
    x264_param_t param;
    x264_param_default( &param );
    param.i_width = w;
    param.i_height = ht;
    param.i_csp = X264_CSP_I420;
    param.pf_log = quiet;
    if( argc > 5 && atoi( argv[5] ) )
        param.analyse.inter |= X264_ANALYSE_PSUB16x16;
    if( argc > 6 )
        param.analyse.i_me_method = atoi( argv[6] );
    if( argc > 7 )
        param.analyse.i_subpel_refine = atoi( argv[7] );
    if( argc > 8 )
    {
        param.rc.i_rc_method = X264_RC_CQP;
        param.rc.i_qp_constant = atoi( argv[8] );
    }
    x264_t *h = x264_encoder_open( &param );
    FILE *in = fopen( argv[1], "rb" ), *out = fopen( argv[2], "wb" );
    if( !h || !in || !out )
    {
        fprintf( stderr, "cannot open the encoder or the files\n" );
        return 1;
    }
    const size_t luma = (size_t)w * ht, pic_bytes = luma * 3 / 2;
    uint8_t *buf = malloc( pic_bytes );
    int flushing = 0, i_frame = 0, idle = 0;
    while( idle < 64 )
    {
        x264_picture_t pic, pic_out;
        x264_nal_t *nal;
        int n_nal = 0, size;
        if( !flushing && fread( buf, 1, pic_bytes, in ) != pic_bytes )
            flushing = 1;
        if( !flushing )
        {
            x264_picture_init( &pic );
            pic.img.i_csp = X264_CSP_I420;
            pic.img.i_plane = 3;
            pic.img.plane[0] = buf;
            pic.img.plane[1] = buf + luma;
            pic.img.plane[2] = buf + luma + luma / 4;
            pic.img.i_stride[0] = w;
            pic.img.i_stride[1] = pic.img.i_stride[2] = w / 2;
            pic.i_pts = i_frame++;
            size = x264_encoder_encode( h, &nal, &n_nal, &pic, &pic_out );
        }
        else
            size = x264_encoder_encode( h, &nal, &n_nal, NULL, &pic_out );
        if( size < 0 )
            return 1;
        for( int k = 0; k < n_nal; k++ )
            fwrite( nal[k].p_payload, 1, nal[k].i_payload, out );
        if( flushing && size == 0 )
            break;
        idle = size ? 0 : idle + flushing;
    }
    x264_encoder_close( h );
    fclose( out );
    fclose( in );
    free( buf );
    return 0;
}
