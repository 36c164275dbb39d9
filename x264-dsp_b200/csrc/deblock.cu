// deblock.cu -- in-loop deblocking of a whole frame and boundary-strength derivation, sm_100a.
//
// Reference: deblock_{v,h}_{luma,chroma}[_intra]_c (common/deblock.c:80-295), deblock_edge and
// x264_frame_deblock_row (deblock.c:325-427, slice-QP rule), deblock_strength_c (deblock.c:297-323).
//
// The reference filters macroblocks in raster order; MB (x,y) reads pixels its left, top and
// top-right neighbours have already modified, a 2-step wavefront.  Mapping:
//   * one WARP per macroblock row, walking left to right; rows are taken through an atomic ticket
//     in top-down order, and row y proceeds to MB x once row y-1 has finished MB x+1 (a progress
//     counter per row, release/acquire through L2);
//   * inside a macroblock the 32 lanes are the 16 luma lines plus the 8 x {U,V} chroma lines of a
//     vertical-edge pass (then the 16 luma columns plus the 16 chroma bytes of a horizontal-edge
//     pass): a lane keeps its line in registers across the four edges, exactly the sequential
//     dependence the reference has, while the lines run in parallel;
//   * pixels cross between lanes and between warps only through memory, with L2-coherent
//     accesses (ld.cg / st.cg), a __syncwarp between the two passes and a fence before the row
//     counter is advanced.
#include "common.cuh"
#include "leaf.cuh"

// what to do on one edge: mode 0 = nothing, 1 = bS<4 filter, 2 = bS=4 filter
struct xd_edge
{
    int mode;
    uint32_t bs;        // four packed bS values
};

__device__ __forceinline__ xd_edge xd_edge_inter( uint32_t bs, int alpha, int beta )
{
    xd_edge e;
    e.bs = bs;
    e.mode = ( bs != 0 && alpha != 0 && beta != 0 ) ? 1 : 0;      // deblock.c:330-331
    return e;
}

__device__ __forceinline__ int xd_ld_u8( const uint8_t *p ) { return __ldcg( p ); }
__device__ __forceinline__ void xd_st_u8( uint8_t *p, int v ) { __stcg( p, (uint8_t)v ); }

__global__ void __launch_bounds__( 128 )
xd_deblock_kernel( x264dsp_geom_t g, uint8_t *__restrict__ slot, const int8_t *__restrict__ mb_type,
                   const uint8_t *__restrict__ partition, const int16_t *__restrict__ cbp,
                   const uint8_t *__restrict__ bs_all, xd_db_params P, int32_t *progress, int32_t *ticket )
{
    const int lane = threadIdx.x & 31;
    const int W = g.mb_w, H = g.mb_h, ls = g.luma_stride, cs = g.chroma_stride;
    for( ;; )
    {
        int mb_y = 0;
        if( lane == 0 )
            mb_y = atomicAdd( ticket, 1 );
        mb_y = __shfl_sync( 0xffffffffu, mb_y, 0 );
        if( mb_y >= H )
            return;
        volatile int32_t *above = progress + mb_y - 1;
        int seen = 0;
        for( int mb_x = 0; mb_x < W; mb_x++ )
        {
            if( mb_y > 0 )
            {
                const int need = min( mb_x + 2, W );
                unsigned ns = 20;
                while( seen < need )
                {
                    seen = *above;
                    if( seen < need )
                    {
                        __nanosleep( ns );
                        if( ns < 400 )
                            ns += 20;
                    }
                }
                __threadfence();
            }
            const int xy = mb_y * W + mb_x;
            const bool intra_cur = mb_type[xy] < 4;
            const bool first_only = partition[xy] == 16 && cbp[xy] == 0 && !intra_cur;
            const uint32_t *bs = (const uint32_t *)( bs_all + (size_t)xy * 64 );     // [2][8] words of 4 bS
            uint8_t *py = slot + g.luma_origin + (int64_t)( mb_y << 4 ) * ls + ( mb_x << 4 );
            uint8_t *pc = slot + g.slot_chroma_off + g.chroma_origin + (int64_t)( mb_y << 3 ) * cs + ( mb_x << 4 );

            // ======================= vertical edges (filter across x) =======================
            {
                xd_edge le[4], ce[2];
                const bool left_intra = mb_x > 0 && ( intra_cur || mb_type[xy - 1] < 4 );
                le[0].mode = mb_x == 0 ? 0 : left_intra ? 2 : xd_edge_inter( bs[0], P.alpha, P.beta ).mode;
                le[0].bs = bs[0];
                ce[0].mode = mb_x == 0 ? 0 : left_intra ? 2 : xd_edge_inter( bs[0], P.alphac, P.betac ).mode;
                ce[0].bs = bs[0];
#pragma unroll
                for( int e = 1; e < 4; e++ )
                {
                    le[e] = xd_edge_inter( bs[e], P.alpha, P.beta );
                    if( first_only )
                        le[e].mode = 0;
                }
                ce[1] = xd_edge_inter( bs[2], P.alphac, P.betac );
                if( first_only )
                    ce[1].mode = 0;

                if( lane < 16 )
                {
                    // luma line `lane`: pixels x-4 .. x+15 in registers
                    uint8_t *row = py + (int64_t)lane * ls;
                    uint32_t w[5];
#pragma unroll
                    for( int k = 0; k < 5; k++ )
                        w[k] = __ldcg( (const uint32_t *)( row - 4 + 4 * k ) );
                    const int grp = lane >> 2;
#pragma unroll
                    for( int e = 0; e < 4; e++ )
                    {
                        if( le[e].mode == 0 )
                            continue;
                        int s[8];
#pragma unroll
                        for( int k = 0; k < 4; k++ )
                        {
                            s[k] = ( w[e] >> ( 8 * k ) ) & 255;
                            s[4 + k] = ( w[e + 1] >> ( 8 * k ) ) & 255;
                        }
                        if( le[e].mode == 2 )
                            xd_luma_intra_line( s, P.alpha, P.beta );
                        else
                        {
                            const int tc0 = xd_tc0( P.ia, ( le[e].bs >> ( 8 * grp ) ) & 255 );
                            if( tc0 >= 0 )
                                xd_luma_line( s, P.alpha, P.beta, tc0 );
                        }
                        w[e] = s[0] | ( s[1] << 8 ) | ( s[2] << 16 ) | ( s[3] << 24 );
                        w[e + 1] = s[4] | ( s[5] << 8 ) | ( s[6] << 16 ) | ( s[7] << 24 );
                    }
#pragma unroll
                    for( int k = 0; k < 5; k++ )
                        if( k > 0 || mb_x > 0 )
                            __stcg( (uint32_t *)( row - 4 + 4 * k ), w[k] );
                }
                else
                {
                    // chroma row r, component c: edges at pair 0 (byte 0) and pair 4 (byte 8)
                    const int r = ( lane - 16 ) >> 1, c = ( lane - 16 ) & 1;
                    uint8_t *row = pc + (int64_t)r * cs + c;
                    const int grp = r >> 1;
#pragma unroll
                    for( int e = 0; e < 2; e++ )
                    {
                        if( ce[e].mode == 0 )
                            continue;
                        uint8_t *q = row + 8 * e;
                        int s[4] = { xd_ld_u8( q - 4 ), xd_ld_u8( q - 2 ), xd_ld_u8( q ), xd_ld_u8( q + 2 ) };
                        const int p0 = s[1], q0 = s[2];
                        if( ce[e].mode == 2 )
                            xd_chroma_line( s, P.alphac, P.betac, 0, true );
                        else
                        {
                            const int tc = xd_tc0( P.iac, ( ce[e].bs >> ( 8 * grp ) ) & 255 ) + 1;
                            if( tc > 0 )
                                xd_chroma_line( s, P.alphac, P.betac, tc, false );
                        }
                        if( s[1] != p0 ) xd_st_u8( q - 2, s[1] );
                        if( s[2] != q0 ) xd_st_u8( q, s[2] );
                    }
                }
            }
            __syncwarp();
            __threadfence_block();

            // ======================= horizontal edges (filter across y) =======================
            {
                xd_edge le[4], ce[2];
                const bool top_intra = mb_y > 0 && ( intra_cur || mb_type[xy - W] < 4 );
                le[0].mode = mb_y == 0 ? 0 : top_intra ? 2 : xd_edge_inter( bs[8], P.alpha, P.beta ).mode;
                le[0].bs = bs[8];
                ce[0].mode = mb_y == 0 ? 0 : top_intra ? 2 : xd_edge_inter( bs[8], P.alphac, P.betac ).mode;
                ce[0].bs = bs[8];
#pragma unroll
                for( int e = 1; e < 4; e++ )
                {
                    le[e] = xd_edge_inter( bs[8 + e], P.alpha, P.beta );
                    if( first_only )
                        le[e].mode = 0;
                }
                ce[1] = xd_edge_inter( bs[10], P.alphac, P.betac );
                if( first_only )
                    ce[1].mode = 0;

                if( lane < 16 )
                {
                    // luma column `lane`: rows -4 .. 15
                    uint8_t *col = py + lane;
                    int v[20];
                    const bool any = ( le[0].mode | le[1].mode | le[2].mode | le[3].mode ) != 0;
                    if( any )
                    {
#pragma unroll
                        for( int k = 0; k < 20; k++ )
                            v[k] = ( k >= 4 || mb_y > 0 ) ? xd_ld_u8( col + (int64_t)( k - 4 ) * ls ) : 0;
                        int orig[20];
#pragma unroll
                        for( int k = 0; k < 20; k++ )
                            orig[k] = v[k];
                        const int grp = lane >> 2;
#pragma unroll
                        for( int e = 0; e < 4; e++ )
                        {
                            if( le[e].mode == 0 )
                                continue;
                            int s[8];
#pragma unroll
                            for( int k = 0; k < 8; k++ )
                                s[k] = v[4 * e + k];
                            if( le[e].mode == 2 )
                                xd_luma_intra_line( s, P.alpha, P.beta );
                            else
                            {
                                const int tc0 = xd_tc0( P.ia, ( le[e].bs >> ( 8 * grp ) ) & 255 );
                                if( tc0 >= 0 )
                                    xd_luma_line( s, P.alpha, P.beta, tc0 );
                            }
#pragma unroll
                            for( int k = 0; k < 8; k++ )
                                v[4 * e + k] = s[k];
                        }
#pragma unroll
                        for( int k = 0; k < 20; k++ )
                            if( v[k] != orig[k] )
                                xd_st_u8( col + (int64_t)( k - 4 ) * ls, v[k] );
                    }
                }
                else
                {
                    // chroma byte b of the 16-byte row (pair b>>1, component b&1): edges at chroma rows 0 and 4
                    const int b = lane - 16;
                    uint8_t *col = pc + b;
                    const int grp = b >> 2;
#pragma unroll
                    for( int e = 0; e < 2; e++ )
                    {
                        if( ce[e].mode == 0 )
                            continue;
                        uint8_t *q = col + (int64_t)( 4 * e ) * cs;
                        int s[4] = { xd_ld_u8( q - 2 * (int64_t)cs ), xd_ld_u8( q - cs ), xd_ld_u8( q ), xd_ld_u8( q + cs ) };
                        const int p0 = s[1], q0 = s[2];
                        if( ce[e].mode == 2 )
                            xd_chroma_line( s, P.alphac, P.betac, 0, true );
                        else
                        {
                            const int tc = xd_tc0( P.iac, ( ce[e].bs >> ( 8 * grp ) ) & 255 ) + 1;
                            if( tc > 0 )
                                xd_chroma_line( s, P.alphac, P.betac, tc, false );
                        }
                        if( s[1] != p0 ) xd_st_u8( q - cs, s[1] );
                        if( s[2] != q0 ) xd_st_u8( q, s[2] );
                    }
                }
            }
            __threadfence();            // every lane publishes its own stores device-wide ...
            __syncwarp();               // ... before lane 0 advances the row counter
            if( lane == 0 )
                *( (volatile int32_t *)( progress + mb_y ) ) = mb_x + 1;
        }
    }
}

// deblock_strength_c for n macroblocks, one thread per (mb, dir, edge, i)
__global__ void __launch_bounds__( 256 )
xd_deblock_strength_kernel( int n, const uint8_t *__restrict__ nnz, const int8_t *__restrict__ ref,
                            const int16_t *__restrict__ mv, uint8_t *__restrict__ bs )
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = t >> 5, k = t & 31;
    if( m >= n )
        return;
    const int dir = k >> 4, edge = ( k >> 2 ) & 3, i = k & 3;
    const int along = dir ? 1 : 8, across = dir ? 8 : 1;
    const int cur = 12 + edge * across + i * along, nb = cur - across;
    const uint8_t *z = nnz + (size_t)m * 120;
    const int8_t *r = ref + (size_t)m * 80;
    const int16_t *v = mv + (size_t)m * 160;
    int s;
    if( z[cur] || z[nb] )
        s = 2;
    else if( r[cur] != r[nb] || abs( v[2 * cur] - v[2 * nb] ) >= 4 || abs( v[2 * cur + 1] - v[2 * nb + 1] ) >= 4 )
        s = 1;
    else
        s = 0;
    bs[(size_t)m * 64 + dir * 32 + edge * 4 + i] = (uint8_t)s;
}

extern "C" int x264dsp_deblock_frame_dev( x264dsp_ctx_t *ctx, const x264dsp_geom_t *g, uint8_t *slot,
                                           const int8_t *mb_type, const uint8_t *partition, const int16_t *cbp,
                                           const uint8_t *bs, int qp, int alpha_c0_offset, int beta_offset,
                                           void *stream )
{
    if( !ctx || !g || !slot || !mb_type || !partition || !cbp || !bs || qp < 0 || qp > 51 )
        return X264DSP_E_ARG;
    static const uint8_t alpha_h[52] =
    {
        0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 4,4,5,6,7,8,9,10,12,13,15,17,20,22,
        25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255
    };
    static const uint8_t beta_h[52] =
    {
        0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0, 2,2,2,3,3,3,3,4,4,4,6,6,7,7,
        8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18
    };
    // the reference's tables carry guard entries: below 0 everything is 0, above 51 the last value
    auto clamp_idx = []( int i ) { return i > 51 ? 51 : i; };
    const int qpc = x264dsp_chroma_qp( qp );
    const int ia = qp + alpha_c0_offset, ib = qp + beta_offset, iac = qpc + alpha_c0_offset, ibc = qpc + beta_offset;
    xd_db_params P;
    P.alpha = ia < 0 ? 0 : alpha_h[clamp_idx( ia )];
    P.beta = ib < 0 ? 0 : beta_h[clamp_idx( ib )];
    P.alphac = iac < 0 ? 0 : alpha_h[clamp_idx( iac )];
    P.betac = ibc < 0 ? 0 : beta_h[clamp_idx( ibc )];
    P.ia = ia < 0 ? -1 : clamp_idx( ia );
    P.iac = iac < 0 ? -1 : clamp_idx( iac );

    cudaStream_t s = xd_stream( ctx, stream );
    // progress counters: one per MB row, plus the ticket, in the lookahead ticket array's tail
    const size_t need = ( (size_t)g->mb_h + 1 ) * sizeof( int32_t );
    if( ctx->db_progress_cap < need )
        XD_CHECK( cudaDeviceSynchronize() );
    int rc = xd_reserve_dev( (void **)&ctx->db_progress, &ctx->db_progress_cap, need );
    if( rc )
        return rc;
    XD_CHECK( cudaMemsetAsync( ctx->db_progress, 0, need, s ) );
    const int ctas = ( g->mb_h + 3 ) / 4;
    const int pslot = xd_prof_begin( ctx, XD_PROF_DEBLOCK, s );
    xd_deblock_kernel<<<ctas, 128, 0, s>>>( *g, slot, mb_type, partition, cbp, bs, P, ctx->db_progress,
                                            ctx->db_progress + g->mb_h );
    xd_prof_end( ctx, XD_PROF_DEBLOCK, pslot, s );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}

extern "C" int x264dsp_deblock_strength_dev( x264dsp_ctx_t *ctx, int n, const uint8_t *nnz, const int8_t *ref,
                                              const int16_t *mv, uint8_t *bs, void *stream )
{
    if( !ctx || n < 0 )
        return X264DSP_E_ARG;
    if( n == 0 )
        return 0;
    if( !nnz || !ref || !mv || !bs )
        return X264DSP_E_ARG;
    const int64_t threads = (int64_t)n * 32;
    xd_deblock_strength_kernel<<<(int)( ( threads + 255 ) / 256 ), 256, 0, xd_stream( ctx, stream )>>>( n, nnz, ref, mv, bs );
    ctx->launches++;
    XD_CHECK( cudaGetLastError() );
    return 0;
}
