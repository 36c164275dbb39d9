// iframe.cu -- the I-slice macroblock loop as a wavefront: intra analysis + coding of every macroblock of an I frame on the
// device, sm_100a.
//
// Reference: the I-slice branch of x264_macroblock_analyse (encoder/analyse.c:1079-1088) = x264_mb_analyse_intra
// (analyse.c:565-763: the 16x16 modes; the sixteen 4x4 blocks one after the other, each with the mode list its neighbours
// allow, the V / H / DC costs first, then the direction-dependent four or a shortcut list, the "predicted mode is free"
// bonus with its early break, the running-cost exit against the 16x16 cost, and the block coded at once because the next
// ones predict from its reconstruction), the decision, x264_mb_analyse_intra_chroma (509-563); mode availability and mode
// prediction (analyse.c:424-508, common/macroblock.h:373-387, common/macroblock.c:217-226, 655-676); then
// x264_macroblock_encode's I16x16 / I4x4 branches (encoder/macroblock.c:72-162, 175-305, 355-377).  Metric: SATD, which is
// what mbcmp_init (encoder/encoder.c:412-432) selects for every subme the reference accepts.
//
// Mapping: as in pframe.cu a warp walks one macroblock row; a macroblock needs the RECONSTRUCTION next to it (the row above
// from the top-left to the top-right macroblock, the column to the left) and the 4x4 modes of the blocks next to it, so row
// y may start macroblock x once row y-1 has published x+1.  The macroblock lives in a shared-memory copy of the reference's
// fdec_buf (stride 32, neighbours in row -1 / column -1); the predictors are the warp-level ones of the table shims
// (predict_warp.cuh), a 16x16 / 8x8 cost is sixteen / four 4x4 SATDs on as many lanes, and the candidate modes of a 4x4
// block are evaluated by one lane each -- the reference's sequential choice (with its breaks) is then replayed on the
// cost vector by every lane alike.
#include "residual_warp.cuh"
#include "predict_warp.cuh"

#define IF_WARPS 2
#ifndef IF_MINB
#define IF_MINB 6                   // resident CTAs per SM the register allocation aims at
#endif
#define IF_COST_MAX ( 1 << 28 )
enum { NB_LEFT = 1, NB_TOP = 2, NB_TOPRIGHT = 4, NB_TOPLEFT = 8 };           // common/macroblock.h:10-13

struct xd_if_args
{
    x264dsp_geom_t g;
    const uint8_t *fenc;
    uint8_t *recon;
    int n_frames, qp, lambda;
    xd_res_tables T;
    int8_t *mb_type;
    uint8_t *mode16, *chroma_mode, *modes4, *mb_kind;
    int16_t *levels, *luma_dc;
    uint8_t *nnz;
    int16_t *cbp;
    int32_t *progress, *ticket;
    const uint4 *weights;           // xd_if_weights_kernel's table
};

struct xd_if_smem
{
    uint8_t fenc_y[16 * 16];                    // fenc_buf: stride 16
    uint8_t fenc_c[8 * 16];                     // U at +0, V at +8
    uint8_t y[18 * FDEC_STRIDE + 8];            // fdec_buf luma: macroblock origin at row 1, byte 8
    uint8_t c[10 * FDEC_STRIDE + 8];            // fdec_buf chroma: U at +0, V at +16 of the same rows
};

__device__ __forceinline__ int xd_if_ld_acquire( const int32_t *p )
{
    int v;
    asm volatile( "ld.acquire.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ int xd_if_ld_relaxed( const int32_t *p )
{
    int v;
    asm volatile( "ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ void xd_if_st_release( int32_t *p, int v )
{
    asm volatile( "st.release.gpu.global.s32 [%0], %1;" :: "l"( p ), "r"( v ) : "memory" );
}

__device__ __forceinline__ int xd_if_ue_bits( int v )                          // bs_size_ue
{
    return 2 * ( 31 - __clz( v + 1 ) ) + 1;
}

// rows of the availability tables (analyse.c:424-508): 0 none, 1 left, 2 top, 3 top + left, 4 top + left + top-left
__device__ __forceinline__ int xd_if_row( int nb )
{
    const int k = nb & ( NB_TOP | NB_LEFT | NB_TOPLEFT );
    return k == ( NB_TOP | NB_LEFT | NB_TOPLEFT ) ? 4 : k & ( NB_TOP | NB_LEFT );
}

// the reference's per-size mode numbers (common/predict.h) -> the predictor routine's
static __constant__ int8_t xd_if_pr16[7] = { PR_V, PR_H, PR_DC, PR_PLANE, PR_DC_LEFT, PR_DC_TOP, PR_DC_128 };
static __constant__ int8_t xd_if_prc[7] = { PR_DC, PR_H, PR_V, PR_PLANE, PR_DC_LEFT, PR_DC_TOP, PR_DC_128 };
static __constant__ int8_t xd_if_modes16[5][5] = { { 6, -1 }, { 4, 1, -1 }, { 5, 0, -1 }, { 0, 1, 2, -1 }, { 0, 1, 2, 3, -1 } };
static __constant__ int8_t xd_if_modesc[5][5] = { { 6, -1 }, { 4, 1, -1 }, { 5, 2, -1 }, { 2, 1, 0, -1 }, { 2, 1, 0, 3, -1 } };
static __constant__ int8_t xd_if_modes4[5][10] = { { 11, -1 }, { 9, 1, 8, -1 }, { 10, 0, 3, 7, -1 }, { 2, 1, 0, 3, 7, 8, -1 },
                                                   { 2, 1, 0, 3, 4, 5, 6, 7, 8, -1 } };
static __constant__ int8_t xd_if_fix4[13] = { -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 2, 2, 2 };      // predict.h:60-68, index = mode + 1
static __constant__ int8_t xd_if_fixc[7] = { 0, 1, 2, 3, 0, 0, 0 }, xd_if_fix16[7] = { 0, 1, 2, 3, 2, 2, 2 };
static __constant__ int8_t xd_if_four[2][4] = { { 3, 4, 6, 8 }, { 3, 4, 5, 7 } };            // intra_mbcmp_x4_4x4_h / _v
static __constant__ int8_t xd_if_short[2][3] = { { 8, -1, -1 }, { 3, 7, -1 } };              // intra_analysis_shortcut[0][0][fv]

// SATD of `blocks` 4x4s (lane l < blocks: block l of a row of `per_row` blocks) of the prediction at pd against fenc at pf
__device__ __forceinline__ int xd_if_satd( const uint8_t *pf, int sf, const uint8_t *pd, int per_row, int blocks, int lane )
{
    int s = 0;
    if( lane < blocks )
    {
        const int bx = lane % per_row, by = lane / per_row;
        uint32_t a[4], b[4];
#pragma unroll
        for( int r = 0; r < 4; r++ )
        {
            a[r] = *(const uint32_t *)( pf + ( by * 4 + r ) * sf + bx * 4 );
            b[r] = *(const uint32_t *)( pd + ( by * 4 + r ) * FDEC_STRIDE + bx * 4 );
        }
        s = xd_satd4x4( a, b );
    }
#pragma unroll
    for( int o = 1; o < 16; o <<= 1 )
        s += __shfl_xor_sync( 0xffffffffu, s, o );
    return __shfl_sync( 0xffffffffu, s, 0 );
}

// x264_mb_encode_i4x4 on the block at dst (already predicted): every lane runs the same 4x4 pipeline, lane 0 stores
__device__ __forceinline__ int xd_if_code4x4( const uint8_t *src, uint8_t *dst, const xd_res_tables &T, int16_t *lv_out, int lane )
{
    uint32_t f[4], p[4];
#pragma unroll
    for( int r = 0; r < 4; r++ )
    {
        f[r] = *(const uint32_t *)( src + r * 16 );
        p[r] = *(const uint32_t *)( dst + r * FDEC_STRIDE );
    }
    int dct[16], lv[16];
    xd_sub4x4_dct( dct, f, p );
    const int nz = xd_quant_4x4( dct, T.luma_i );
    xd_zigzag( lv, dct );
    if( nz )
    {
        xd_dequant_4x4( dct, T.luma_i );
        xd_add4x4_idct( p, dct );
    }
    __syncwarp();
    if( lane == 0 )
    {
        xd_store_levels( lv_out, lv );
#pragma unroll
        for( int r = 0; r < 4; r++ )
            *(uint32_t *)( dst + r * FDEC_STRIDE ) = p[r];
    }
    __syncwarp();
    return nz;
}

// ---------------------------------------------------------------------------------------------
// The 4x4 predictors as byte dot products.  Every sample of every I_PRED_4x4 mode is (sum_k w_k e_k + rnd) >> sh over the
// thirteen neighbouring samples e[0..12] = l3 l2 l1 l0 lt t0..t7 (common/predict.c:330-470): the three-tap and two-tap filters
// with weights 1 2 1 and 2 2 over four, a copied sample with weight 4 (all rnd 2, sh 2), DC as eight weights of 1 (rnd 4, sh 3),
// DC_128 as no weights and rnd 514.  With one lane per candidate mode a switch over the mode runs all twelve cases one after
// the other (ncu r02at: 30 % of the kernel's instructions, 40 % of its stall samples); as a table of weights -- one 16-byte entry
// per (mode, sample): the thirteen weights as bytes -- a sample is one shared-memory load and four IDP.4A, the same code on every
// lane.  The table is DERIVED from xd_pred4x4_px (the per-sample restatement of the reference that tests/ pin): the response to
// 4 at neighbour k and 0 elsewhere is w_k for the rnd 2 / sh 2 modes and for DC alike.
#define IF_W_ENTRIES ( 12 * 16 )

__global__ void xd_if_weights_kernel( uint4 *table )
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= IF_W_ENTRIES )
        return;
    const int mode = t >> 4, x = t & 3, y = ( t >> 2 ) & 3;
    uint32_t w[4] = { 0, 0, 0, 0 };
    if( mode != PR_DC_128 )
        for( int k = 0; k < 13; k++ )
        {
            int e[13];
            for( int j = 0; j < 13; j++ )
                e[j] = j == k ? 4 : 0;
            // DC = (l0 + .. + l3 + t0 + .. + t3 + 4) >> 3: a weight of 1 on each of its eight samples
            const int r = mode == PR_DC ? ( ( k < 4 || ( k >= 5 && k < 9 ) ) ? 1 : 0 ) : xd_pred4x4_px( mode, x, y, e );
            w[k >> 2] |= (uint32_t)r << ( 8 * ( k & 3 ) );
        }
    table[t] = make_uint4( w[0], w[1], w[2], w[3] );
}

// the thirteen neighbours of the 4x4 block at dst (shared-memory fdec, stride FDEC_STRIDE) packed four to a word
struct xd_if_edges
{
    uint32_t e0, e1, e2, e3;
};
__device__ __forceinline__ xd_if_edges xd_if_load_edges( const uint8_t *dst )
{
    const uint32_t t0 = *(const uint32_t *)( dst - FDEC_STRIDE ), t1 = *(const uint32_t *)( dst - FDEC_STRIDE + 4 );
    xd_if_edges E;
    E.e0 = (uint32_t)dst[3 * FDEC_STRIDE - 1] | ( (uint32_t)dst[2 * FDEC_STRIDE - 1] << 8 ) | ( (uint32_t)dst[FDEC_STRIDE - 1] << 16 )
         | ( (uint32_t)dst[-1] << 24 );
    E.e1 = (uint32_t)dst[-FDEC_STRIDE - 1] | ( t0 << 8 );
    E.e2 = ( t0 >> 24 ) | ( t1 << 8 );
    E.e3 = t1 >> 24;
    return E;
}
__device__ __forceinline__ int xd_if_pred_px( const xd_if_edges &E, const uint4 w, int mode )
{
    const int rnd = mode == PR_DC ? 4 : mode == PR_DC_128 ? 514 : 2, sh = mode == PR_DC ? 3 : 2;
    return (int)( __dp4a( E.e0, w.x, __dp4a( E.e1, w.y, __dp4a( E.e2, w.z, __dp4a( E.e3, w.w, (uint32_t)rnd ) ) ) ) >> sh );
}

__global__ void __launch_bounds__( IF_WARPS * 32, IF_MINB )
xd_iframe_kernel( xd_if_args A )
{
    __shared__ __align__( 16 ) xd_if_smem s_mb[IF_WARPS];
    __shared__ uint4 s_w[IF_W_ENTRIES];
    for( int i = threadIdx.x; i < IF_W_ENTRIES; i += IF_WARPS * 32 )
        s_w[i] = A.weights[i];
    __syncthreads();
    const x264dsp_geom_t &g = A.g;
    const int lane = threadIdx.x & 31;
    xd_if_smem &S = s_mb[threadIdx.x >> 5];
    uint8_t *fy = S.y + FDEC_STRIDE + 8, *fc = S.c + FDEC_STRIDE + 8;
    const int W = g.mb_w, H = g.mb_h, ls = g.luma_stride, cs = g.chroma_stride;
    const int total = A.n_frames * H, lambda = A.lambda;
    for( ;; )
    {
        int t = 0;
        if( lane == 0 )
            t = atomicAdd( A.ticket, 1 );
        t = __shfl_sync( 0xffffffffu, t, 0 );
        if( t >= total )
            return;
        const int mb_y = t / A.n_frames, f = t - mb_y * A.n_frames;
        const uint8_t *fenc = A.fenc + (size_t)f * g.slot_bytes;
        uint8_t *recon = A.recon + (size_t)f * g.slot_bytes;
        const size_t mb0 = (size_t)f * g.mb_count;
        uint8_t *modes4 = A.modes4 + mb0 * 16;
        int16_t *levels = A.levels + mb0 * X264DSP_RES_LEVELS_PER_MB;
        uint8_t *nnz = A.nnz + mb0 * X264DSP_RES_NNZ_PER_MB;
        int16_t *cbp = A.cbp + mb0;
        int32_t *mine = A.progress + (size_t)f * H + mb_y;
        const int32_t *above = mine - 1;
        int seen = 0;
        for( int mb_x = 0; mb_x < W; mb_x++ )
        {
            const int xy = mb_y * W + mb_x;
            if( mb_y > 0 )
            {
                const int need = min( mb_x + 2, W );
                if( seen < need )
                {
                    if( lane == 0 )
                    {
                        unsigned ns = 40;
                        while( xd_if_ld_relaxed( above ) < need )
                        {
                            __nanosleep( ns );
                            if( ns < 1000 )
                                ns += 40;
                        }
                    }
                    __syncwarp();
                    seen = xd_if_ld_acquire( above );
                }
            }
            const int nb = ( mb_x > 0 ? NB_LEFT : 0 ) | ( mb_y > 0 ? NB_TOP : 0 ) | ( mb_x > 0 && mb_y > 0 ? NB_TOPLEFT : 0 )
                         | ( mb_y > 0 && mb_x < W - 1 ? NB_TOPRIGHT : 0 );
            const int64_t oy = g.luma_origin + (int64_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
            const int64_t oc = g.slot_chroma_off + g.chroma_origin + (int64_t)( mb_y << 3 ) * cs + ( mb_x << 4 );

            // ---- fenc_buf and fdec_buf (common/macroblock.c:242-265): source samples, reconstructed neighbours
            if( lane < 16 )
                *(uint4 *)( S.fenc_y + lane * 16 ) = __ldg( (const uint4 *)( fenc + oy + (int64_t)lane * ls ) );
            else if( lane < 24 )
            {
                const uint4 v = __ldg( (const uint4 *)( fenc + oc + (int64_t)( lane - 16 ) * cs ) );     // U0 V0 U1 V1 ...
                uint8_t *d = S.fenc_c + ( lane - 16 ) * 16;
                *(uint32_t *)d = __byte_perm( v.x, v.y, 0x6420 );
                *(uint32_t *)( d + 4 ) = __byte_perm( v.z, v.w, 0x6420 );
                *(uint32_t *)( d + 8 ) = __byte_perm( v.x, v.y, 0x7531 );
                *(uint32_t *)( d + 12 ) = __byte_perm( v.z, v.w, 0x7531 );
            }
            if( ( nb & NB_TOP ) && lane < 21 )
                fy[-FDEC_STRIDE - 1 + lane] = __ldcg( recon + oy - ls - 1 + lane );
            if( ( nb & NB_LEFT ) && lane < 16 )
                fy[lane * FDEC_STRIDE - 1] = __ldcg( recon + oy + (int64_t)lane * ls - 1 );
            if( ( nb & NB_TOP ) && lane < 18 )
            {
                // row above of U (x = -1 .. 7) and V: NV12 bytes -2 .. 15
                const uint8_t v = __ldcg( recon + oc - cs - 2 + lane );
                fc[-FDEC_STRIDE + ( lane & 1 ? 16 : 0 ) + ( lane >> 1 ) - 1] = v;
            }
            if( ( nb & NB_LEFT ) && lane >= 24 )
            {
                const int r = lane - 24;
                fc[r * FDEC_STRIDE - 1] = __ldcg( recon + oc + (int64_t)r * cs - 2 );
                fc[r * FDEC_STRIDE + 15] = __ldcg( recon + oc + (int64_t)r * cs - 1 );
            }
            // the 4x4 modes next to the macroblock (macroblock.c:447-522): left neighbour's right column, upper one's bottom row
            int mleft = -1, mtop = -1;
            if( lane < 4 )
            {
                const int rc = lane == 0 ? 5 : lane == 1 ? 7 : lane == 2 ? 13 : 15, br = lane == 0 ? 10 : lane == 1 ? 11 : lane == 2 ? 14 : 15;
                if( nb & NB_LEFT )
                    mleft = (int8_t)__ldcg( modes4 + (size_t)( xy - 1 ) * 16 + rc );
                if( nb & NB_TOP )
                    mtop = (int8_t)__ldcg( modes4 + (size_t)( xy - W ) * 16 + br );
            }
            __syncwarp();

            // ---- 16x16 (analyse.c:590-627)
            int satd16 = IF_COST_MAX, mode16 = 0;
            {
                const int8_t *list = xd_if_modes16[xd_if_row( nb )];
                for( int k = 0; k < 4 && list[k] >= 0; k++ )
                {
                    const int m = list[k];
                    xs_predict( fy, 16, xd_if_pr16[m], lane );
                    __syncwarp();
                    const int c = xd_if_satd( S.fenc_y, 16, fy, 4, 16, lane ) + lambda * xd_if_ue_bits( xd_if_fix16[m] );
                    if( c < satd16 ) { satd16 = c; mode16 = m; }
                    __syncwarp();
                }
            }

            // ---- 4x4 (analyse.c:629-763)
            int satd4 = IF_COST_MAX;
            uint32_t nz_bits = 0;
            int cbp_luma = 0;
            uint32_t my_modes = 0;                    // lane l < 16 keeps the mode of block l
            {
                int cost = lambda * 40, idx;
                int mode15 = 2;
                // the modes of the blocks already decided, as a 5x5 cache [1 + by][1 + bx] spread over the lanes: lane 5 * r + c
                int cache = -1;
                {
                    const int r = lane / 5, c = lane - 5 * r;
                    const int top_v = __shfl_sync( 0xffffffffu, mtop, ( lane - 1 ) & 3 );
                    const int left_v = __shfl_sync( 0xffffffffu, mleft, ( r - 1 ) & 3 );
                    if( lane >= 1 && lane <= 4 )
                        cache = top_v;                                                          // row 0, columns 1 .. 4
                    if( lane < 25 && c == 0 && r >= 1 )
                        cache = left_v;                                                         // column 0, rows 1 .. 4
                }
                for( idx = 0; ; idx++ )
                {
                    const int bx = ( idx & 1 ) + ( ( idx >> 2 ) & 1 ) * 2, by = ( ( idx >> 1 ) & 1 ) + ( ( idx >> 3 ) & 1 ) * 2;
                    const uint8_t *src = S.fenc_y + by * 4 * 16 + bx * 4;
                    uint8_t *dst = fy + by * 4 * FDEC_STRIDE + bx * 4;
                    int nb4;
                    if( idx == 6 || idx == 9 || idx == 12 || idx == 14 )
                        nb4 = NB_LEFT | NB_TOP | NB_TOPLEFT | NB_TOPRIGHT;
                    else if( idx == 3 || idx == 7 || idx == 11 || idx == 13 || idx == 15 )
                        nb4 = NB_LEFT | NB_TOP | NB_TOPLEFT;
                    else if( idx == 0 )
                        nb4 = ( nb & ( NB_TOP | NB_LEFT | NB_TOPLEFT ) ) | ( ( nb & NB_TOP ) ? NB_TOPRIGHT : 0 );
                    else if( idx == 1 || idx == 4 )
                        nb4 = NB_LEFT | ( ( nb & NB_TOP ) ? NB_TOP | NB_TOPLEFT | NB_TOPRIGHT : 0 );
                    else if( idx == 5 )
                        nb4 = NB_LEFT | ( nb & NB_TOPRIGHT ) | ( ( nb & NB_TOP ) ? NB_TOP | NB_TOPLEFT : 0 );
                    else
                        nb4 = NB_TOP | NB_TOPRIGHT | ( ( nb & NB_LEFT ) ? NB_LEFT | NB_TOPLEFT : 0 );
                    const int row = xd_if_row( nb4 );
                    // x264_mb_predict_intra4x4_mode
                    const int ma = xd_if_fix4[__shfl_sync( 0xffffffffu, cache, 5 * ( 1 + by ) + bx ) + 1];
                    const int mb_ = xd_if_fix4[__shfl_sync( 0xffffffffu, cache, 5 * by + 1 + bx ) + 1];
                    int pred = min( ma, mb_ );
                    if( pred < 0 )
                        pred = 2;
                    if( ( nb4 & ( NB_TOPRIGHT | NB_TOP ) ) == NB_TOP )     // emulate the missing top-right samples
                    {
                        const uint8_t v = dst[3 - FDEC_STRIDE];
                        __syncwarp();
                        if( lane < 4 )
                            dst[4 - FDEC_STRIDE + lane] = v;
                        __syncwarp();
                    }
                    // every candidate mode's SATD, one lane per mode
                    int satd[12];
                    {
                        const xd_if_edges E = xd_if_load_edges( dst );
                        const int m = lane < 12 ? lane : 0;
                        uint32_t a[4], b[4];
#pragma unroll
                        for( int r = 0; r < 4; r++ )
                        {
                            a[r] = *(const uint32_t *)( src + r * 16 );
                            b[r] = (uint32_t)xd_if_pred_px( E, s_w[m * 16 + 4 * r], m ) | ( (uint32_t)xd_if_pred_px( E, s_w[m * 16 + 4 * r + 1], m ) << 8 )
                                 | ( (uint32_t)xd_if_pred_px( E, s_w[m * 16 + 4 * r + 2], m ) << 16 )
                                 | ( (uint32_t)xd_if_pred_px( E, s_w[m * 16 + 4 * r + 3], m ) << 24 );
                        }
                        const int c = xd_satd4x4( a, b );
#pragma unroll
                        for( int k = 0; k < 12; k++ )
                            satd[k] = __shfl_sync( 0xffffffffu, c, k );
                    }
                    // the reference's sequential choice, replayed on the cost vector
                    int best = IF_COST_MAX, best_mode = 2;
                    const int8_t *list = xd_if_modes4[row];
                    bool walk = true;
                    if( row >= 3 )
                    {
                        const int fv = satd[1] > satd[0];
#pragma unroll
                        for( int k = 0; k < 12; k++ )
                            if( k == pred )
                                satd[k] -= 3 * lambda;       // only modes 0 .. 8 can be predicted; the bonus below is the loop's own
                        best = satd[2]; best_mode = 2;
                        if( satd[1] < best ) { best = satd[1]; best_mode = 1; }
                        if( satd[0] < best ) { best = satd[0]; best_mode = 0; }
                        if( row == 4 )
                        {
#pragma unroll
                            for( int k = 0; k < 4; k++ )
                            {
                                const int m = xd_if_four[fv][k];
                                int c = 0;
#pragma unroll
                                for( int j = 3; j < 9; j++ )
                                    if( j == m )
                                        c = satd[j];
                                if( c < best ) { best = c; best_mode = m; }
                            }
                            walk = false;
                        }
                        else
                        {
                            list = xd_if_short[fv];
                            // the shortcut loop applies its own bonus: undo the one taken above for its candidates
#pragma unroll
                            for( int k = 3; k < 12; k++ )
                                if( k == pred )
                                    satd[k] += 3 * lambda;
                        }
                    }
                    if( walk && best > 0 )
                        for( int k = 0; list[k] >= 0; k++ )
                        {
                            const int m = list[k];
                            int c = 0;
#pragma unroll
                            for( int j = 0; j < 12; j++ )
                                if( j == m )
                                    c = satd[j];
                            if( pred == xd_if_fix4[m + 1] )
                            {
                                c -= 3 * lambda;
                                if( c <= 0 )
                                {
                                    best = c;
                                    best_mode = m;
                                    break;
                                }
                            }
                            if( c < best ) { best = c; best_mode = m; }
                        }
                    if( lane == idx )
                        my_modes = (uint32_t)best_mode;
                    cost += best + 3 * lambda;
                    if( idx == 15 )
                        mode15 = best_mode;
                    if( cost > satd16 || idx == 15 )
                        break;
                    if( lane == 5 * ( 1 + by ) + 1 + bx )
                        cache = best_mode;
                    // predict with the chosen mode and code the block: the next ones predict from its reconstruction
                    {
                        const xd_if_edges E = xd_if_load_edges( dst );
                        __syncwarp();
                        if( lane < 16 )
                            dst[( lane >> 2 ) * FDEC_STRIDE + ( lane & 3 )] = (uint8_t)xd_if_pred_px( E, s_w[best_mode * 16 + lane], best_mode );
                        __syncwarp();
                    }
                    if( xd_if_code4x4( src, dst, A.T, levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB + idx * 16, lane ) )
                    {
                        cbp_luma |= 1 << ( idx >> 2 );
                        nz_bits |= 1u << idx;
                    }
                }
                if( idx == 15 )
                {
                    satd4 = cost;
                    if( satd4 < satd16 )
                    {
                        // I_4x4 it is: block 15 is coded now (x264_macroblock_encode does it, macroblock.c:360-377)
                        uint8_t *dst = fy + 12 * FDEC_STRIDE + 12;
                        const xd_if_edges E = xd_if_load_edges( dst );
                        __syncwarp();
                        if( lane < 16 )
                            dst[( lane >> 2 ) * FDEC_STRIDE + ( lane & 3 )] = (uint8_t)xd_if_pred_px( E, s_w[mode15 * 16 + lane], mode15 );
                        __syncwarp();
                        if( xd_if_code4x4( S.fenc_y + 12 * 16 + 12, dst, A.T, levels + (size_t)xy * X264DSP_RES_LEVELS_PER_MB + 15 * 16, lane ) )
                        {
                            cbp_luma |= 8;
                            nz_bits |= 1u << 15;
                        }
                    }
                }
            }
            const bool i4 = satd4 < satd16;                       // COPY2_IF_LT( i_cost, i_satd_i4x4, type, I_4x4 )

            // ---- chroma (analyse.c:509-563)
            int cmode = 0;
            {
                int best = IF_COST_MAX;
                const int8_t *list = xd_if_modesc[xd_if_row( nb )];
                for( int k = 0; k < 4 && list[k] >= 0; k++ )
                {
                    const int m = list[k];
                    xs_predict( fc, 8, xd_if_prc[m], lane );
                    xs_predict( fc + 16, 8, xd_if_prc[m], lane );
                    __syncwarp();
                    const int c = xd_if_satd( S.fenc_c, 16, fc, 2, 4, lane ) + xd_if_satd( S.fenc_c + 8, 16, fc + 16, 2, 4, lane )
                                + lambda * xd_if_ue_bits( xd_if_fixc[m] );
                    if( c < best ) { best = c; cmode = m; }
                    __syncwarp();
                }
                xs_predict( fc, 8, xd_if_prc[cmode], lane );
                xs_predict( fc + 16, 8, xd_if_prc[cmode], lane );
            }
            if( !i4 )
                xs_predict( fy, 16, xd_if_pr16[mode16], lane );
            __syncwarp();

            // ---- x264_macroblock_encode: the luma prediction (I_16x16) or reconstruction (I_4x4) and the chroma prediction
            //      go to the frame, the typed residual routine does the rest in place
            if( lane < 16 )
                *(uint4 *)( recon + oy + (int64_t)lane * ls ) = make_uint4( *(const uint32_t *)( fy + lane * FDEC_STRIDE ),
                    *(const uint32_t *)( fy + lane * FDEC_STRIDE + 4 ), *(const uint32_t *)( fy + lane * FDEC_STRIDE + 8 ),
                    *(const uint32_t *)( fy + lane * FDEC_STRIDE + 12 ) );
            else if( lane < 24 )
            {
                const int r = lane - 16;
                const uint32_t u0 = *(const uint32_t *)( fc + r * FDEC_STRIDE ), u1 = *(const uint32_t *)( fc + r * FDEC_STRIDE + 4 );
                const uint32_t v0 = *(const uint32_t *)( fc + r * FDEC_STRIDE + 16 ), v1 = *(const uint32_t *)( fc + r * FDEC_STRIDE + 20 );
                *(uint4 *)( recon + oc + (int64_t)r * cs ) = make_uint4( __byte_perm( u0, v0, 0x5140 ), __byte_perm( u0, v0, 0x7362 ),
                                                                         __byte_perm( u1, v1, 0x5140 ), __byte_perm( u1, v1, 0x7362 ) );
            }
            if( lane == 0 )
            {
                A.mb_kind[mb0 + xy] = i4 ? 2 : 1;
                A.mb_type[mb0 + xy] = i4 ? 0 : 2;                 // I_4x4 / I_16x16 (common/macroblock.h:41-50)
                A.mode16[mb0 + xy] = (uint8_t)mode16;
                A.chroma_mode[mb0 + xy] = (uint8_t)cmode;
            }
            if( lane < 16 )
                modes4[(size_t)xy * 16 + lane] = i4 ? (uint8_t)my_modes : 2;     // what the neighbours see (macroblock.c:736-745)
            __syncwarp();
            xd_residual_mb<true, false>( g, fenc, recon, A.T, levels, nnz, cbp, A.mb_kind + mb0, A.luma_dc + mb0 * 16, xy, lane );
            __syncwarp();
            if( i4 )
            {
                if( lane < 16 )
                    nnz[(size_t)xy * X264DSP_RES_NNZ_PER_MB + lane] = (uint8_t)( ( nz_bits >> lane ) & 1u );
                if( lane == 0 )
                    cbp[xy] = (int16_t)( ( cbp[xy] & ~0x10F ) | cbp_luma );
            }
            __syncwarp();
            if( lane == 0 )
                xd_if_st_release( mine, mb_x + 1 );
        }
    }
}

extern "C" int x264dsp_i_frames_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, const uint8_t *fenc_slots, uint8_t *recon_slots,
                                     int n_frames, int qp, int8_t *mb_type, uint8_t *mode16, uint8_t *chroma_mode,
                                     uint8_t *modes4, int16_t *levels, int16_t *luma_dc, uint8_t *nnz, int16_t *cbp, void *stream )
{
    if( !ctx || !g || !fenc_slots || !recon_slots || !mb_type || !mode16 || !chroma_mode || !modes4 || !levels || !luma_dc
        || !nnz || !cbp || n_frames <= 0 || n_frames > 65535 || qp < 0 || qp > 51 )
        return X264DSP_E_ARG;
    cudaStream_t s = xd_stream( ctx, stream );
    xd_if_args A;
    memset( &A, 0, sizeof( A ) );
    A.g = *g;
    A.fenc = fenc_slots;
    A.recon = recon_slots;
    A.n_frames = n_frames;
    A.qp = qp;
    A.lambda = x264dsp_lambda( qp );
    xd_residual_tables( qp, &A.T );
    A.mb_type = mb_type;
    A.mode16 = mode16;
    A.chroma_mode = chroma_mode;
    A.modes4 = modes4;
    A.levels = levels;
    A.luma_dc = luma_dc;
    A.nnz = nnz;
    A.cbp = cbp;
    // scratch: progress counters, the ticket, one kind byte per macroblock (the deblocking wavefront's buffer)
    const size_t rows = (size_t)n_frames * g->mb_h, nmb = (size_t)n_frames * g->mb_count;
    const size_t need = ( rows + 1 ) * sizeof( int32_t ) + nmb;
    if( ctx->db_progress_cap < need )
        XD_CHECK( cudaDeviceSynchronize() );
    int rc = xd_reserve_dev( (void **)&ctx->db_progress, &ctx->db_progress_cap, need );
    if( rc )
        return rc;
    if( ( rc = xd_scratch_acquire( ctx, XD_SCRATCH_DEBLOCK, s ) ) )
        return rc;
    XD_CHECK( cudaMemsetAsync( ctx->db_progress, 0, need, s ) );
    A.progress = ctx->db_progress;
    A.ticket = ctx->db_progress + rows;
    A.mb_kind = (uint8_t *)( ctx->db_progress + rows + 1 );
    if( !ctx->if_weights )
    {
        // the predictors' weight table, once per context
        XD_CHECK( cudaMalloc( &ctx->if_weights, IF_W_ENTRIES * sizeof( uint4 ) ) );
        xd_if_weights_kernel<<<( IF_W_ENTRIES + 63 ) / 64, 64, 0, s>>>( (uint4 *)ctx->if_weights );
        ctx->launches++;
    }
    A.weights = (const uint4 *)ctx->if_weights;
    XD_CHECK( cudaMemsetAsync( levels, 0, nmb * X264DSP_RES_LEVELS_PER_MB * sizeof( int16_t ), s ) );
    int per_sm = 0;
    XD_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, xd_iframe_kernel, IF_WARPS * 32, 0 ) );
    if( per_sm < 1 )
        per_sm = 1;
    int64_t ctas = (int64_t)ctx->sm_count * per_sm;
    const int64_t wanted = ( (int64_t)rows + IF_WARPS - 1 ) / IF_WARPS;
    if( ctas > wanted )
        ctas = wanted;
    xd_iframe_kernel<<<(unsigned)ctas, IF_WARPS * 32, 0, s>>>( A );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return xd_scratch_release( ctx, XD_SCRATCH_DEBLOCK, s );
}
