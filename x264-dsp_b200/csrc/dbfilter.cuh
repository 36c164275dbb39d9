// dbfilter.cuh -- branch-free deblocking line filters (common/deblock.c:80-295), one sample line per call.
//
// Every function computes the filtered samples unconditionally and selects at the end, so that the 32 lanes of a warp
// (different lines, different boundary strengths) run one instruction stream.  `act` is the lane's own "this line is
// filtered at all" flag (bS > 0, edge enabled); the sample-dependent conditions of the standard are evaluated inside.
// The file is plain C++ apart from the qualifiers, so tests/host_dbfilter_check.cpp compiles it with g++ and compares it
// with the straightforward forms in leaf.cuh's style.
#pragma once

#if defined( __CUDACC__ )
#define XDF_FN __host__ __device__ __forceinline__
#else
#define XDF_FN static inline
#endif

// primitives: single instructions on the device (VABSDIFF, VIMNMX, VIADDMNMX.RELU), plain C on the host.  All samples
// are 0..255 and every clip range has lo <= hi.
#if defined( __CUDA_ARCH__ )
XDF_FN int xdf_absdiff( int a, int b ) { return (int)__sad( a, b, 0u ); }
XDF_FN int xdf_clip3( int v, int lo, int hi ) { return min( max( v, lo ), hi ); }
XDF_FN int xdf_u8( int v ) { return min( max( v, 0 ), 255 ); }
#else
XDF_FN int xdf_absdiff( int a, int b ) { return a < b ? b - a : a - b; }
XDF_FN int xdf_clip3( int v, int lo, int hi ) { return v < lo ? lo : ( v > hi ? hi : v ); }
XDF_FN int xdf_u8( int v ) { return v < 0 ? 0 : ( v > 255 ? 255 : v ); }
#endif
XDF_FN int xdf_rhadd( int a, int b ) { return ( a + b + 1 ) >> 1; }
XDF_FN int xdf_hadd( int a, int b ) { return ( a + b ) >> 1; }

// bS < 4, luma (deblock_edge_luma_c, deblock.c:80-120).  p2 / q2 are read only.  tc0 >= 0 when act.
// "if( tc0 )" of the reference needs no branch: with tc0 == 0 the clip range is [0, 0] and p1 / q1 keep their value.
XDF_FN void xdf_luma_normal( int p2, int &p1, int &p0, int &q0, int &q1, int q2, int alpha, int beta, int tc0, bool act )
{
    const bool on = act && xdf_absdiff( p0, q0 ) < alpha && xdf_absdiff( p1, p0 ) < beta && xdf_absdiff( q1, q0 ) < beta;
    const bool ap = xdf_absdiff( p2, p0 ) < beta, aq = xdf_absdiff( q2, q0 ) < beta;
    const int avg = xdf_rhadd( p0, q0 );
    const int tc = tc0 + ( ap ? 1 : 0 ) + ( aq ? 1 : 0 );
    const int d = xdf_clip3( ( ( ( q0 - p0 ) << 2 ) + ( p1 - q1 ) + 4 ) >> 3, -tc, tc );
    const int np1 = p1 + xdf_clip3( xdf_hadd( p2, avg ) - p1, -tc0, tc0 );
    const int nq1 = q1 + xdf_clip3( xdf_hadd( q2, avg ) - q1, -tc0, tc0 );
    const int np0 = xdf_u8( p0 + d ), nq0 = xdf_u8( q0 - d );
    p1 = ( on && ap ) ? np1 : p1;
    q1 = ( on && aq ) ? nq1 : q1;
    p0 = on ? np0 : p0;
    q0 = on ? nq0 : q0;
}

// bS == 4, luma (deblock_edge_luma_intra_c, deblock.c:196-243)
XDF_FN void xdf_luma_intra( int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3, int alpha, int beta, bool act )
{
    const int d0 = xdf_absdiff( p0, q0 );
    const bool on = act && d0 < alpha && xdf_absdiff( p1, p0 ) < beta && xdf_absdiff( q1, q0 ) < beta;
    const bool strong = d0 < ( ( alpha >> 2 ) + 2 );
    const bool sp = strong && xdf_absdiff( p2, p0 ) < beta, sq = strong && xdf_absdiff( q2, q0 ) < beta;
    const int wp0 = ( 2 * p1 + p0 + q1 + 2 ) >> 2, wq0 = ( 2 * q1 + q0 + p1 + 2 ) >> 2;
    const int sp0 = ( p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4 ) >> 3;
    const int sp1 = ( p2 + p1 + p0 + q0 + 2 ) >> 2;
    const int sp2 = ( 2 * p3 + 3 * p2 + p1 + p0 + q0 + 4 ) >> 3;
    const int sq0 = ( p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4 ) >> 3;
    const int sq1 = ( p0 + q0 + q1 + q2 + 2 ) >> 2;
    const int sq2 = ( 2 * q3 + 3 * q2 + q1 + q0 + p0 + 4 ) >> 3;
    const int np0 = sp ? sp0 : wp0, nq0 = sq ? sq0 : wq0;
    p0 = on ? np0 : p0;
    q0 = on ? nq0 : q0;
    p1 = ( on && sp ) ? sp1 : p1;
    p2 = ( on && sp ) ? sp2 : p2;
    q1 = ( on && sq ) ? sq1 : q1;
    q2 = ( on && sq ) ? sq2 : q2;
}

// chroma, bS < 4 with tc = tc0 + 1 > 0 (deblock_edge_chroma_c, deblock.c:147-167) or bS == 4 (deblock.c:261-278)
XDF_FN void xdf_chroma( int p1, int &p0, int &q0, int q1, int alpha, int beta, int tc, bool intra, bool act )
{
    const bool on = act && xdf_absdiff( p0, q0 ) < alpha && xdf_absdiff( p1, p0 ) < beta && xdf_absdiff( q1, q0 ) < beta;
    const int d = xdf_clip3( ( ( ( q0 - p0 ) << 2 ) + ( p1 - q1 ) + 4 ) >> 3, -tc, tc );
    const int np0 = intra ? ( 2 * p1 + p0 + q1 + 2 ) >> 2 : xdf_u8( p0 + d );
    const int nq0 = intra ? ( 2 * q1 + q0 + p1 + 2 ) >> 2 : xdf_u8( q0 - d );
    p0 = on ? np0 : p0;
    q0 = on ? nq0 : q0;
}
