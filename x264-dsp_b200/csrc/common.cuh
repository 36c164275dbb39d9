// common.cuh -- shared declarations of the CUDA side of libx264dsp_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/x264dsp_b200.h"

// kernel classes for x264dsp_profile_read
enum { XD_PROF_LOAD = 0, XD_PROF_LOWRES, XD_PROF_LA_INTRA, XD_PROF_LA_INTER, XD_PROF_HPEL, XD_PROF_BORDER,
       XD_PROF_COST, XD_PROF_ME, XD_PROF_MC, XD_PROF_RESIDUAL, XD_PROF_DEBLOCK, XD_PROF_LA_TILE, XD_PROF_KINDS };
#define XD_PROF_MAX 512
#define XD_AUX_STREAMS 16

// ---------------------------------------------------------------------------------------------
// context: one per process/GPU.  Owns the stream, the constant tables in HBM and the scratch
// buffers of the batched entry points.
struct x264dsp_ctx
{
    int device;
    int sm_count;
    cudaStream_t stream;
    int64_t launches;

    // cost_mv tables, one per distinct lambda (encoder/analyse.c:243-315); cost_mv_dev[qp] points at
    // entry 0 of a uint16[8193] whose centre (index 4096) is mv delta 0.
    uint16_t *cost_mv_store;
    const uint16_t *cost_mv_dev[52];

    // lookahead scratch (grown on demand)
    unsigned long long *la_sync;   // [pairs][mb_count] {mv, epoch} words
    int32_t *la_icost;             // [pairs][mb_count] intra cost per block
    int32_t *la_ticket;            // work-queue counter
    size_t la_sync_cap, la_icost_cap;
    uint32_t la_epoch;
    unsigned long long *la_timing; // phase-cycle counters of the inter kernel (debug aid, normally NULL)
    int la_kernel;                 // x264dsp_lookahead_select_kernel: 0 auto, 1 warp-per-row, 2 quad-row
    int copies_only;               // x264dsp_debug_copies_only: the host entry points skip their kernels

    // host-API staging (x264dsp_lookahead_clip_host)
    uint8_t *stage_host;  size_t stage_host_cap;     // pinned
    uint8_t *stage_dev;   size_t stage_dev_cap;
    uint8_t *clip_slots;  size_t clip_slots_cap;
    uint8_t *clip_out;    size_t clip_out_cap;       // device results
    uint8_t *clip_out_host; size_t clip_out_host_cap; // pinned results
    int32_t *clip_desc;   size_t clip_desc_cap;
    void *desc_cache;     size_t desc_cache_bytes;   // host copy of what clip_desc holds

    // host-level full-resolution paths (host_paths.cu): block lists / side information and results on the device
    uint8_t *me_blocks;   size_t me_blocks_cap;
    uint8_t *me_results;  size_t me_results_cap;
    void *if_weights;                                // iframe.cu: the 4x4 predictors' weight table
    uint8_t *gc_scratch;  size_t gc_scratch_cap;     // gop_chain.cu: boundary strengths of one GOP position (+ 16x16 vectors)
    cudaEvent_t host_ev;

    // deblock wavefront: per-row progress counters + ticket
    int32_t *db_progress; size_t db_progress_cap;

    // extra streams so that independent groups of a host-level batch overlap copies and kernels
    cudaStream_t aux[XD_AUX_STREAMS];

    // optional per-kernel timing with CUDA events (x264dsp_profile_*)
    int prof_on;
    int prof_n[XD_PROF_KINDS];
    cudaEvent_t prof_ev[XD_PROF_KINDS][XD_PROF_MAX][2];

    // the wavefront kernels' scratch above (lookahead: sync words / tickets / epoch; deblock: progress counters) is
    // per context, not per call: launches that reuse it are ordered after its previous user even when they arrive
    // on different streams (xd_scratch_acquire / _release)
    cudaEvent_t scratch_ev[2];
    cudaStream_t scratch_last[2];
    int scratch_busy[2];

    // per-call table shims (tables.cu)
    uint8_t *shim_host;   // pinned
    uint8_t *shim_dev;
    size_t shim_cap;
};

#define XD_CHECK( call )                                                        \
    do {                                                                        \
        cudaError_t e_ = ( call );                                              \
        if( e_ != cudaSuccess )                                                 \
            return (int)e_;                                                     \
    } while( 0 )

static inline cudaStream_t xd_stream( x264dsp_ctx *ctx, void *stream )
{
    return stream ? (cudaStream_t)stream : ctx->stream;
}

// event pair around one launch when profiling is on; returns the slot or -1
static inline int xd_prof_begin( x264dsp_ctx *ctx, int kind, cudaStream_t s )
{
    if( !ctx->prof_on || ctx->prof_n[kind] >= XD_PROF_MAX )
        return -1;
    const int i = ctx->prof_n[kind]++;
    if( !ctx->prof_ev[kind][i][0] )
    {
        cudaEventCreate( &ctx->prof_ev[kind][i][0] );
        cudaEventCreate( &ctx->prof_ev[kind][i][1] );
    }
    cudaEventRecord( ctx->prof_ev[kind][i][0], s );
    return i;
}
static inline void xd_prof_end( x264dsp_ctx *ctx, int kind, int slot, cudaStream_t s )
{
    if( slot >= 0 )
        cudaEventRecord( ctx->prof_ev[kind][slot][1], s );
}

enum { XD_SCRATCH_LOOKAHEAD = 0, XD_SCRATCH_DEBLOCK = 1 };
// Before a launch that uses the context's wavefront scratch `which` on stream s: if the previous user ran on another
// stream, make s wait for it (stream-side wait, the host does not block).  After the launch: mark s as the user.
static inline int xd_scratch_acquire( x264dsp_ctx *ctx, int which, cudaStream_t s )
{
    if( ctx->scratch_busy[which] && ctx->scratch_last[which] != s )
        XD_CHECK( cudaStreamWaitEvent( s, ctx->scratch_ev[which], 0 ) );
    return 0;
}
static inline int xd_scratch_release( x264dsp_ctx *ctx, int which, cudaStream_t s )
{
    if( !ctx->scratch_ev[which] )
        XD_CHECK( cudaEventCreateWithFlags( &ctx->scratch_ev[which], cudaEventDisableTiming ) );
    XD_CHECK( cudaEventRecord( ctx->scratch_ev[which], s ) );
    ctx->scratch_last[which] = s;
    ctx->scratch_busy[which] = 1;
    return 0;
}

// grows a device buffer to at least `bytes`
int xd_reserve_dev( void **p, size_t *cap, size_t bytes );
int xd_reserve_pinned( void **p, size_t *cap, size_t bytes );

// host tables (tables_host.cpp)
extern "C" int x264dsp_lambda( int qp );

// ---------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

// per-byte rounded average (a+b+1)>>1 of four packed pixels, the pixel_avg of common/mc.c:74-87
__device__ __forceinline__ uint32_t xd_avg4( uint32_t a, uint32_t b )
{
    return ( a | b ) - ( ( ( a ^ b ) & 0xFEFEFEFEu ) >> 1 );
}

// clip to 0..255 and pack four values, a0 in the lowest byte (two I2IP)
__device__ __forceinline__ uint32_t xd_pack_sat4( int a0, int a1, int a2, int a3 )
{
    uint32_t t, d;
    asm( "cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"( t ) : "r"( a3 ), "r"( a2 ), "r"( 0 ) );
    asm( "cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a1 ), "r"( a0 ), "r"( t ) );
    return d;
}

// four-pixel SAD with accumulate: one VABSDIFF4.U8.ACC
__device__ __forceinline__ uint32_t xd_sad4( uint32_t a, uint32_t b, uint32_t acc )
{
    return __vsadu4( a, b ) + acc;
}

__device__ __forceinline__ int xd_clip3( int v, int lo, int hi )
{
    return min( max( v, lo ), hi );
}

__device__ __forceinline__ int xd_clip_u8( int v )
{
    return min( max( v, 0 ), 255 );
}

// 8 consecutive pixels starting at an arbitrary byte address, through the read-only path:
// three aligned words and two funnel shifts.
__device__ __forceinline__ uint2 xd_load8_unaligned( const uint8_t *p )
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)( a & ~(uintptr_t)3 );
    const uint32_t sh = ( (uint32_t)a & 3u ) * 8u;
    uint32_t w0 = __ldg( w ), w1 = __ldg( w + 1 ), w2 = __ldg( w + 2 );
    uint2 r;
    r.x = __funnelshift_r( w0, w1, sh );
    r.y = __funnelshift_r( w1, w2, sh );
    return r;
}

// 4 consecutive pixels at an arbitrary byte address
__device__ __forceinline__ uint32_t xd_load4_unaligned( const uint8_t *p )
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)( a & ~(uintptr_t)3 );
    const uint32_t sh = ( (uint32_t)a & 3u ) * 8u;
    return __funnelshift_r( __ldg( w ), __ldg( w + 1 ), sh );
}

// Quarter-pel plane selection of mc_luma / get_ref (common/mc.c:192-193, 222-234).
// phase = (mvy&3)*4 + (mvx&3).  Packed as two bits per phase.
__device__ __forceinline__ int xd_qpel_plane_a( int phase )
{
    // {0,1,1,1, 0,1,1,1, 2,3,3,3, 0,1,1,1}
    return (int)( ( 0x54FE5454u >> ( phase * 2 ) ) & 3u );
}
__device__ __forceinline__ int xd_qpel_plane_b( int phase )
{
    // {0,0,0,0, 2,2,3,2, 2,2,3,2, 2,2,3,2}
    return (int)( ( 0xBABABA00u >> ( phase * 2 ) ) & 3u );
}

#endif // __CUDACC__
