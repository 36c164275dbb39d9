/* xo_tables.c -- oracle: constant tables of the hot path.  TEST INFRASTRUCTURE ONLY (see xo.h).
 *
 * lambda / mv-bit costs: encoder/analyse.c:98-111, 171-206, 243-315
 * flat-CQM quant / dequant / bias: common/set.c:265-353 (portable branch)
 * luma->chroma QP map: common/macroblock.h:251-266 (chroma_qp_index_offset = 0)
 */
#include <string.h>
#include "xo.h"

/* lambda(qp) ~ 2^(qp/6 - 2), the reference's rounded integer table (analyse.c:98-111), qp 0..51 */
static const uint16_t lambda_tab[52] =
{
     1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,
     2,  2,  2,  2,  3,  3,  3,  4,  4,  4,  5,  6,  6,  7,  8,  9,
    10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57,
    64, 72, 81, 91
};

/* lambda^2 * .9 * 256 (analyse.c:117-131), qp 0..51 */
static const int lambda2_tab[52] =
{
        14,     18,     22,     28,     36,     45,     57,     72,
        91,    115,    145,    182,    230,    290,    365,    460,
       580,    731,    921,   1161,   1462,   1843,   2322,   2925,
      3686,   4644,   5851,   7372,   9289,  11703,  14745,  18578,
     23407,  29491,  37156,  46814,  58982,  74313,  93628, 117964,
    148626, 187257, 235929, 297252, 374514, 471859, 594505, 749029,
    943718,1189010,1498059,1887436
};

int xo_lambda( int qp )  { return lambda_tab[qp]; }
int xo_lambda2( int qp ) { return lambda2_tab[qp]; }

/* mv-bit step function (analyse.c:171-206): |d| in [first, first+count) costs `bits` bits;
 * the last range is cut so that the table ends at |d| = 4096. */
static const uint16_t mv_bits_steps[23][3] =
{
    {  4,    1,    1 }, {  5,    2,    1 }, {  6,    3,    2 }, {  7,    5,    2 },
    {  8,    7,    3 }, {  9,   10,    4 }, { 10,   14,    6 }, { 11,   20,    9 },
    { 12,   29,   12 }, { 13,   41,   18 }, { 14,   59,   24 }, { 15,   83,   35 },
    { 16,  118,   49 }, { 17,  167,   70 }, { 18,  237,   98 }, { 19,  335,  139 },
    { 20,  474,  197 }, { 21,  671,  278 }, { 22,  949,  393 }, { 23, 1342,  556 },
    { 24, 1898,  787 }, { 25, 2685, 1112 }, { 26, 3797,  300 }
};

/* cost_mv[qp] (analyse.c:243-315): entry 0 = lambda, entry +-d = lambda * bits(d) */
void xo_cost_mv_table( int qp, uint16_t out[8193] )
{
    const int lambda = lambda_tab[qp];
    int s, d;
    memset( out, 0, 8193 * sizeof(uint16_t) );
    out[4096] = (uint16_t)lambda;
    for( s = 0; s < 23; s++ )
        for( d = mv_bits_steps[s][1]; d < mv_bits_steps[s][1] + mv_bits_steps[s][2]; d++ )
            out[4096 + d] = out[4096 - d] = (uint16_t)( lambda * mv_bits_steps[s][0] );
}

/* set.c:271-286: per-position scale classes of the 4x4 core transform */
static const uint8_t dequant_scale[6][3] =
{
    { 10, 13, 16 }, { 11, 14, 18 }, { 13, 16, 20 }, { 14, 18, 23 }, { 16, 20, 25 }, { 18, 23, 29 }
};
static const uint16_t quant_scale[6][3] =
{
    { 13107, 8066, 5243 }, { 11916, 7490, 4660 }, { 10082, 6554, 4194 },
    {  9362, 5825, 3647 }, {  8192, 5243, 3355 }, {  7282, 4559, 2893 }
};

static int pos_class( int i ) { return (i & 1) + ((i >> 2) & 1); }

/* set.c:337-350.  b_inter selects the dead zone: intra 32-11 = 21, inter 32-21 = 11 (set.c:291-294) */
void xo_quant_tables( int b_inter, int qp, uint16_t mf[16], uint16_t bias[16] )
{
    const int deadzone = b_inter ? 11 : 21;
    const int shift = qp / 6 - 1;
    int i;
    for( i = 0; i < 16; i++ )
    {
        int base = quant_scale[qp % 6][pos_class( i )];
        int m = shift <= 0 ? base << -shift : (base + (1 << (shift - 1))) >> shift;
        int rounded, cap;
        m &= 0xffff;
        mf[i] = (uint16_t)m;
        rounded = ((deadzone << 10) + (m >> 1)) / m;
        cap = (1 << 15) / m;
        bias[i] = (uint16_t)( rounded < cap ? rounded : cap );
    }
}

/* set.c:325-331 with the flat scaling list (16) */
void xo_dequant_table( int out[6][16] )
{
    int q, i;
    for( q = 0; q < 6; q++ )
        for( i = 0; i < 16; i++ )
            out[q][i] = dequant_scale[q][pos_class( i )] * 16;
}

/* common/macroblock.h:251-266 */
int xo_chroma_qp( int qp )
{
    static const uint8_t above29[22] =
        { 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };
    if( qp < 0 ) return 0;
    if( qp < 30 ) return qp;
    if( qp <= 51 ) return above29[qp - 30];
    return 39;
}
