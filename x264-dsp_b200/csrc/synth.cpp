// synth.cpp -- seeded synthetic YUV 4:2:0 input.  Host code, no CUDA.
//
// Luma of frame n = clip( T[(y+2n) mod H][(x+3n) mod W]  (a box-blurred xorshift32 field stretched
// to 0..255, i.e. a texture panning 3,2 px/frame), overlaid with two 48x48 gradient squares moving
// at (3,2) and (-2,1) px/frame, plus uniform noise in [-2,2] ).  U = Y(2x,2y)/2 + 64,
// V = 191 - Y(2x,2y)/2.  From frame `cut_frame` on a different texture seed is used (scene cut).
// Every consumer (tests, bench, CPU baseline) gets its bytes from here, so they are identical.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <mutex>
#include "../../include/x264dsp_b200.h"

namespace {

struct Texture
{
    int w = 0, h = 0;
    uint32_t seed = 0;
    std::vector<uint8_t> pix;
};

inline uint32_t xorshift32( uint32_t &s )
{
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}

void build_texture( Texture &t, int w, int h, uint32_t seed )
{
    t.w = w; t.h = h; t.seed = seed;
    std::vector<int> raw( (size_t)w * h ), tmp( (size_t)w * h );
    uint32_t s = seed;
    for( size_t i = 0; i < raw.size(); i++ )
        raw[i] = xorshift32( s ) & 255;
    const int R = 3;                                   // 7x7 box, wrapping
    for( int y = 0; y < h; y++ )
    {
        const int *row = &raw[(size_t)y * w];
        int acc = 0;
        for( int k = -R; k <= R; k++ )
            acc += row[( k + w ) % w];
        for( int x = 0; x < w; x++ )
        {
            tmp[(size_t)y * w + x] = acc;
            acc += row[( x + R + 1 ) % w] - row[( x - R + w ) % w];
        }
    }
    int lo = 1 << 30, hi = -1;
    for( int x = 0; x < w; x++ )
    {
        int acc = 0;
        for( int k = -R; k <= R; k++ )
            acc += tmp[(size_t)( ( k + h ) % h ) * w + x];
        for( int y = 0; y < h; y++ )
        {
            raw[(size_t)y * w + x] = acc;
            lo = acc < lo ? acc : lo;
            hi = acc > hi ? acc : hi;
            acc += tmp[(size_t)( ( y + R + 1 ) % h ) * w + x] - tmp[(size_t)( ( y - R + h ) % h ) * w + x];
        }
    }
    t.pix.resize( (size_t)w * h );
    const int span = hi > lo ? hi - lo : 1;
    for( size_t i = 0; i < raw.size(); i++ )
        t.pix[i] = (uint8_t)( (int64_t)( raw[i] - lo ) * 255 / span );
}

std::mutex g_lock;
Texture g_tex[2];

} // namespace

extern "C" int x264dsp_synth_frame( int width, int height, int n, int cut_frame,
                                    uint8_t *y, uint8_t *u, uint8_t *v )
{
    if( width < 64 || height < 64 || ( width & 1 ) || ( height & 1 ) || !y || n < 0 )
        return X264DSP_E_ARG;
    const bool after_cut = cut_frame >= 0 && n >= cut_frame;
    const uint32_t seed = after_cut ? 0x9E3779B9u ^ 0x5bd1e995u : 0x9E3779B9u;
    const uint8_t *tex;
    {
        std::lock_guard<std::mutex> guard( g_lock );
        Texture &t = g_tex[after_cut ? 1 : 0];
        if( t.w != width || t.h != height || t.seed != seed )
            build_texture( t, width, height, seed );
        tex = t.pix.data();
    }
    const int sq = 48;
    const int ax = ( width / 4 + 3 * n ) % ( width - sq ), ay = ( height / 4 + 2 * n ) % ( height - sq );
    const int bx = ( ( width / 2 - 2 * n ) % ( width - sq ) + ( width - sq ) ) % ( width - sq );
    const int by = ( height / 2 + n ) % ( height - sq );
    uint32_t s = 0x85EBCA6Bu + (uint32_t)n;
    for( int r = 0; r < height; r++ )
    {
        const uint8_t *trow = tex + (size_t)( ( r + 2 * n ) % height ) * width;
        uint8_t *dst = y + (size_t)r * width;
        for( int c = 0; c < width; c++ )
        {
            int val = trow[( c + 3 * n ) % width];
            if( (unsigned)( c - ax ) < (unsigned)sq && (unsigned)( r - ay ) < (unsigned)sq )
                val = ( ( c - ax ) * 4 + ( r - ay ) * 2 ) & 255;
            else if( (unsigned)( c - bx ) < (unsigned)sq && (unsigned)( r - by ) < (unsigned)sq )
                val = 255 - ( ( ( c - bx ) * 3 + ( r - by ) * 5 ) & 255 );
            val += (int)( xorshift32( s ) % 5u ) - 2;
            dst[c] = (uint8_t)( val < 0 ? 0 : val > 255 ? 255 : val );
        }
    }
    if( u && v )
        for( int r = 0; r < height / 2; r++ )
            for( int c = 0; c < width / 2; c++ )
            {
                int half = y[(size_t)( 2 * r ) * width + 2 * c] >> 1;
                u[(size_t)r * ( width / 2 ) + c] = (uint8_t)( half + 64 );
                v[(size_t)r * ( width / 2 ) + c] = (uint8_t)( 191 - half );
            }
    return 0;
}
