// predict_warp.cuh -- the 16x16, 8x8c and 4x4 intra predictors (common/predict.c:42-470) by one warp on an fdec-style buffer
// (stride 32, neighbours in the row above / the column to the left); shared by the per-call table shims (tables.cu) and the
// I-slice wavefront (iframe.cu).
#pragma once
#include "common.cuh"
#include "leaf.cuh"

#ifndef FDEC_STRIDE
#define FDEC_STRIDE 32          // common/common.h:872
#endif

// predict `mode` into the block at fd (stride 32) from its top row / left column
// (common/predict.c:42-130 16x16, 224-288 8x8c, 330-470 4x4); the warp writes the whole block
static __device__ void xs_predict( uint8_t *fd, int size, int mode, int lane )
{
    if( size == 4 )
    {
        int e[13];
        for( int k = 0; k < 4; k++ )
            e[3 - k] = fd[k * FDEC_STRIDE - 1];
        e[4] = fd[-FDEC_STRIDE - 1];
        for( int k = 0; k < 8; k++ )
            e[5 + k] = fd[-FDEC_STRIDE + k];
        if( lane < 16 )
            fd[( lane >> 2 ) * FDEC_STRIDE + ( lane & 3 )] = (uint8_t)xd_pred4x4_px( mode, lane & 3, lane >> 2, e );
        return;
    }
    const int n = size * size;
    int dcq[4] = { 0, 0, 0, 0 };
    if( mode == PR_PLANE )
    {
        // x264_predict_16x16_p_c (predict.c:125-158), x264_predict_8x8c_p_c (predict.c:290-318)
        const int half = size >> 1;
        int H = 0, V = 0;
        for( int i = 0; i < half; i++ )
        {
            H += ( i + 1 ) * ( fd[half + i - FDEC_STRIDE] - fd[half - 2 - i - FDEC_STRIDE] );
            V += ( i + 1 ) * ( fd[-1 + ( half + i ) * FDEC_STRIDE] - fd[-1 + ( half - 2 - i ) * FDEC_STRIDE] );
        }
        const int a = 16 * ( fd[-1 + ( size - 1 ) * FDEC_STRIDE] + fd[size - 1 - FDEC_STRIDE] );
        const int b = size == 16 ? ( 5 * H + 32 ) >> 6 : ( 17 * H + 16 ) >> 5;
        const int c = size == 16 ? ( 5 * V + 32 ) >> 6 : ( 17 * V + 16 ) >> 5;
        const int i00 = a - ( half - 1 ) * ( b + c ) + 16;
        __syncwarp();
        for( int i = lane; i < n; i += 32 )
        {
            const int x = i % size, y = i / size;
            fd[y * FDEC_STRIDE + x] = (uint8_t)min( max( ( i00 + b * x + c * y ) >> 5, 0 ), 255 );
        }
        return;
    }
    if( mode == PR_DC_LEFT || mode == PR_DC_TOP || mode == PR_DC_128 )
    {
        // predict.c:62-94 (16x16), 163-213 (8x8c): one value for the block, or one per half for 8x8c
        if( mode == PR_DC_128 )
            dcq[0] = dcq[1] = dcq[2] = dcq[3] = 128;
        else if( size == 16 )
        {
            int dc = 0;
            for( int i = 0; i < 16; i++ )
                dc += mode == PR_DC_LEFT ? fd[i * FDEC_STRIDE - 1] : fd[i - FDEC_STRIDE];
            dcq[0] = ( dc + 8 ) >> 4;
        }
        else
        {
            int d0 = 0, d1 = 0;
            for( int i = 0; i < 4; i++ )
            {
                d0 += mode == PR_DC_LEFT ? fd[i * FDEC_STRIDE - 1] : fd[i - FDEC_STRIDE];
                d1 += mode == PR_DC_LEFT ? fd[( i + 4 ) * FDEC_STRIDE - 1] : fd[i + 4 - FDEC_STRIDE];
            }
            d0 = ( d0 + 2 ) >> 2;
            d1 = ( d1 + 2 ) >> 2;
            if( mode == PR_DC_LEFT ) { dcq[0] = dcq[1] = d0; dcq[2] = dcq[3] = d1; }
            else { dcq[0] = dcq[2] = d0; dcq[1] = dcq[3] = d1; }
        }
    }
    if( mode == PR_DC )
    {
        if( size == 16 )
        {
            int dc = 0;
            for( int i = 0; i < 16; i++ )
                dc += fd[i * FDEC_STRIDE - 1] + fd[i - FDEC_STRIDE];
            dcq[0] = ( dc + 16 ) >> 5;
        }
        else
        {
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for( int i = 0; i < 4; i++ )
            {
                s0 += fd[i - FDEC_STRIDE];
                s1 += fd[i + 4 - FDEC_STRIDE];
                s2 += fd[i * FDEC_STRIDE - 1];
                s3 += fd[( i + 4 ) * FDEC_STRIDE - 1];
            }
            dcq[0] = ( s0 + s2 + 4 ) >> 3; dcq[1] = ( s1 + 2 ) >> 2; dcq[2] = ( s3 + 2 ) >> 2; dcq[3] = ( s1 + s3 + 4 ) >> 3;
        }
    }
    for( int i = lane; i < n; i += 32 )
    {
        const int x = i % size, y = i / size;
        int v;
        if( mode == PR_V )
            v = fd[x - FDEC_STRIDE];
        else if( mode == PR_H )
            v = fd[y * FDEC_STRIDE - 1];
        else
            v = size == 16 ? dcq[0] : dcq[( y >> 2 ) * 2 + ( x >> 2 )];
        fd[y * FDEC_STRIDE + x] = (uint8_t)v;
    }
}

