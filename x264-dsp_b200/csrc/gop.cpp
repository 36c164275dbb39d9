// gop.cpp -- slice types and closed GOPs from the lookahead's frame costs, and their distribution over ranks (host code).
//
// Reference: x264_slicetype_analyse + scenecut (encoder/slicetype.c:322-435), the key-frame rules of
// x264_slicetype_decide (slicetype.c:508-537), for the configuration the reference runs: no B frames, closed GOPs.
//
// SURVEY 8(f) N4, "lookahead-first GOP split": every frame cost the decision reads -- the intra estimate of frame k and its
// inter estimate against frame k-1 (x264dsp_lookahead_frame_cost_dev's sums) -- comes from SOURCE frames only, so the
// types of a whole sequence follow from one lookahead pass (itself sharded by frame range, x264dsp_frame_range) in one
// sequential scan; after that every GOP is an independent unit of encoding work (with one reference frame nothing reaches
// across an I frame) and GOPs, not frames, are what the ranks of a box share out.
#include <stdlib.h>
#include "../../include/x264dsp_b200.h"

extern "C" int x264dsp_slicetype_decide( int n_frames, const int32_t *icost, const int32_t *pcost,
                                         const x264dsp_gop_params_t *p, uint8_t *types )
{
    if( n_frames < 0 || !p || p->keyint_max < 1 || p->keyint_min < 1 || p->keyint_min > p->keyint_max
        || ( n_frames > 0 && ( !icost || !pcost || !types ) ) )
        return X264DSP_E_ARG;
    int last_key = -p->keyint_max;                                       // encoder/lookahead.c:35
    for( int k = 0; k < n_frames; k++ )
    {
        int t = 0;                                                        // X264_TYPE_AUTO
        if( k > 0 )
        {
            // x264_slicetype_analyse: one undecided frame at a time, frames[0] = the previous frame
            const int keyint_limit = p->keyint_max - ( k - 1 ) + last_key - 1;
            if( keyint_limit <= 0 )
                t = X264DSP_TYPE_I;
            else
            {
                t = X264DSP_TYPE_P;
                if( p->scenecut_threshold )
                {
                    // scenecut (slicetype.c:322-352): the bias grows with the distance to the last key frame
                    const int gop = k - last_key;
                    const int tmax = p->scenecut_threshold;
                    const int tmin = p->keyint_min == p->keyint_max ? tmax : tmax >> 2;
                    int bias;
                    if( gop <= ( p->keyint_min >> 2 ) )
                        bias = tmin >> 2;
                    else if( gop <= p->keyint_min )
                        bias = tmin * gop / p->keyint_min;
                    else
                        bias = tmin + ( tmax - tmin ) * ( gop - p->keyint_min ) / ( p->keyint_max - p->keyint_min );
                    if( 100 * (int64_t)pcost[k] >= (int64_t)( 100 - bias ) * icost[k] )
                        t = X264DSP_TYPE_I;
                }
            }
        }
        // x264_slicetype_decide (slicetype.c:515-537), closed GOPs
        if( k - last_key >= p->keyint_max && ( t == 0 || t == X264DSP_TYPE_I ) )
            t = X264DSP_TYPE_IDR;
        if( t == X264DSP_TYPE_I && k - last_key >= p->keyint_min )
            t = X264DSP_TYPE_IDR;
        if( t == X264DSP_TYPE_IDR )
            last_key = k;
        types[k] = (uint8_t)( t ? t : X264DSP_TYPE_P );
    }
    return 0;
}

extern "C" int x264dsp_gop_ranges( int n_frames, const uint8_t *types, int32_t *gop_first, int32_t *gop_count, int *n_gops )
{
    if( n_frames < 0 || !n_gops || ( n_frames > 0 && ( !types || !gop_first || !gop_count ) ) )
        return X264DSP_E_ARG;
    int n = 0;
    for( int k = 0; k < n_frames; k++ )
    {
        if( k == 0 || types[k] != X264DSP_TYPE_P )
        {
            gop_first[n] = k;
            gop_count[n] = 0;
            n++;
        }
        gop_count[n - 1]++;
    }
    *n_gops = n;
    return 0;
}

// contiguous runs of GOPs, cut where the running frame count crosses rank * total / world: every rank gets whole GOPs,
// in order, and within one GOP's length of its fair share of frames
extern "C" int x264dsp_gop_shard( int n_gops, const int32_t *gop_count, int rank, int world, int *first_gop, int *count )
{
    if( n_gops < 0 || world < 1 || rank < 0 || rank >= world || !first_gop || !count || ( n_gops > 0 && !gop_count ) )
        return X264DSP_E_ARG;
    int64_t total = 0;
    for( int i = 0; i < n_gops; i++ )
        total += gop_count[i];
    int lo = n_gops, hi = n_gops;
    int64_t before = 0;
    for( int i = 0, r = 0; i < n_gops; i++ )
    {
        // GOP i belongs to the rank whose share its first frame falls into
        while( r < world - 1 && before * world >= total * ( r + 1 ) )
            r++;
        if( r == rank )
        {
            if( lo == n_gops )
                lo = i;
            hi = i + 1;
        }
        else if( r > rank && lo == n_gops )
            lo = hi = i;
        before += gop_count[i];
    }
    if( lo == n_gops )
        hi = n_gops;
    *first_gop = lo;
    *count = hi - lo;
    return 0;
}
