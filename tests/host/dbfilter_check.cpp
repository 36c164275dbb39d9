// Host-side door to x264-dsp_b200/csrc/dbfilter.cuh (the branch-free line filters the deblocking kernel runs per lane),
// so that tests/test_host_leaf.py can compare them with the oracle without a GPU.  TEST INFRASTRUCTURE.
#include "../../x264-dsp_b200/csrc/dbfilter.cuh"

extern "C" {
// s = p3 p2 p1 p0 q0 q1 q2 q3, updated in place
void chk_luma_normal( int *s, int alpha, int beta, int tc0, int act )
{
    xdf_luma_normal( s[1], s[2], s[3], s[4], s[5], s[6], alpha, beta, tc0, act != 0 );
}
void chk_luma_intra( int *s, int alpha, int beta, int act )
{
    xdf_luma_intra( s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], alpha, beta, act != 0 );
}
// s = p1 p0 q0 q1
void chk_chroma( int *s, int alpha, int beta, int tc, int intra, int act )
{
    xdf_chroma( s[0], s[1], s[2], s[3], alpha, beta, tc, intra != 0, act != 0 );
}
}
