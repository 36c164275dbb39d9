// int_pipe_peak.cu -- dependency-free issue-rate microbenchmark of the integer instructions the ME / SATD / lookahead
// kernels are made of ("measure it once on the box ... and record inst/clk/SM").  Stand-alone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/int_pipe_peak tools/int_pipe_peak.cu
//   tools/_build/int_pipe_peak > profiles/int_pipe_peak.json       (on the GPU box)
// Each kernel keeps 8 independent accumulator chains per thread (ILP 8), 1024 threads x 2 CTAs per SM (all 64 warp
// slots), so that neither latency nor occupancy hides the pipe's issue rate.  Rates are thread-instructions per
// clock per SM, once from in-kernel clock64() deltas (independent of the clock the GPU actually ran at) and once from
// the event-timed duration at the nominal maximum SM clock (cudaDevAttrClockRate).
// Result on B200 (profiles/int_pipe_peak.json): every ALU-pipe integer instruction the kernels use issues at
// 64 thread-instructions / clk / SM = 2 warp instructions / clk / SM (one per 2 clk per SMSP), SHFL at half of that;
// the integer-pipe peak is therefore HALF of the 4 warp-instructions / clk / SM issue peak.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define ITERS 4096
#define ILP 8

enum { OP_VABSDIFF4_ACC, OP_VABSDIFF4, OP_IDP4A, OP_VIADD16X2, OP_IADD3, OP_LOP3, OP_PRMT, OP_VIMNMX, OP_IMAD, OP_SHFL,
       OP_SHF, OP_MIX_SAD, OP_KINDS };
static const char *op_name[OP_KINDS] = { "VABSDIFF4.U8.ACC", "VABSDIFF4.U8", "IDP.4A.U8.U8", "VIADD.16x2", "IADD3", "LOP3",
                                         "PRMT", "VIMNMX3", "IMAD", "SHFL.BFLY", "SHF.R (funnel)", "mix: 2 VABSDIFF4.ACC + 1 SHF + 1 IADD3" };

template<int OP>
__device__ __forceinline__ uint32_t step( uint32_t acc, uint32_t a, uint32_t b )
{
    uint32_t r;
    if( OP == OP_VABSDIFF4_ACC )
        asm volatile( "vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"( r ) : "r"( a ), "r"( b ), "r"( acc ) );
    else if( OP == OP_VABSDIFF4 )
        asm volatile( "vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"( r ) : "r"( acc ), "r"( b ), "r"( 0u ) );
    else if( OP == OP_IDP4A )
        asm volatile( "dp4a.u32.u32 %0, %1, %2, %3;" : "=r"( r ) : "r"( a ), "r"( b ), "r"( acc ) );
    else if( OP == OP_VIADD16X2 )
        asm volatile( "add.u16x2 %0, %1, %2;" : "=r"( r ) : "r"( acc ), "r"( a ) );
    else if( OP == OP_IADD3 )
        asm volatile( "{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( b ) );   // one IADD3
    else if( OP == OP_LOP3 )
        asm volatile( "lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( b ) );
    else if( OP == OP_PRMT )
        asm volatile( "prmt.b32 %0, %1, %2, %3;" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( 0x5410u ) );
    else if( OP == OP_VIMNMX )
        asm volatile( "{ .reg .u32 t; min.u32 t, %1, %2; min.u32 %0, t, %3; }" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( b ) );   // one VIMNMX3
    else if( OP == OP_IMAD )
        asm volatile( "mad.lo.u32 %0, %1, %2, %3;" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( b ) );
    else if( OP == OP_SHFL )
        asm volatile( "shfl.sync.bfly.b32 %0, %1, 1, 0x1f, 0xffffffff;" : "=r"( r ) : "r"( acc ) );
    else if( OP == OP_SHF )
        asm volatile( "shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"( r ) : "r"( acc ), "r"( a ), "r"( b ) );
    else
        r = acc;
    return r;
}

template<int OP>
__global__ void __launch_bounds__( 1024, 2 ) pipe_kernel( uint32_t *out, unsigned long long *clk, uint32_t seed )
{
    uint32_t acc[ILP];
    const uint32_t a = seed * ( threadIdx.x + 1 ), b = seed ^ ( blockIdx.x * 0x9E3779B9u + threadIdx.x );
#pragma unroll
    for( int i = 0; i < ILP; i++ )
        acc[i] = a + i;
    __syncthreads();
    const long long t0 = clock64();
    if( OP == OP_MIX_SAD )
    {
        // the inner loop of a SAD row as the lookahead / ME kernels issue it: two accumulating 4-pixel SADs, one funnel
        // shift to align the candidate, one address / counter add
        for( int it = 0; it < ITERS; it++ )
#pragma unroll
            for( int i = 0; i < ILP; i += 4 )
            {
                acc[i] = step<OP_VABSDIFF4_ACC>( acc[i], a, b );
                acc[i + 1] = step<OP_VABSDIFF4_ACC>( acc[i + 1], b, a );
                acc[i + 2] = step<OP_SHF>( acc[i + 2], a, b );
                acc[i + 3] = step<OP_IADD3>( acc[i + 3], a, b );
            }
    }
    else
    {
        for( int it = 0; it < ITERS; it++ )
#pragma unroll
            for( int i = 0; i < ILP; i++ )
                acc[i] = step<OP>( acc[i], a, b );
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for( int i = 0; i < ILP; i++ )
        s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if( threadIdx.x == 0 )
        clk[blockIdx.x] = (unsigned long long)( t1 - t0 );
}

// two different instructions alternating on independent chains: 128 thread-inst / clk / SM means the two issue on
// different pipes (ALU + FMA), 64 means they share one
template<int OPA, int OPB>
__global__ void __launch_bounds__( 1024, 2 ) pair_kernel( uint32_t *out, unsigned long long *clk, uint32_t seed )
{
    uint32_t acc[ILP];
    const uint32_t a = seed * ( threadIdx.x + 1 ), b = seed ^ ( blockIdx.x * 0x9E3779B9u + threadIdx.x );
#pragma unroll
    for( int i = 0; i < ILP; i++ )
        acc[i] = a + i;
    __syncthreads();
    const long long t0 = clock64();
    for( int it = 0; it < ITERS; it++ )
#pragma unroll
        for( int i = 0; i < ILP; i += 2 )
        {
            acc[i] = step<OPA>( acc[i], a, b );
            acc[i + 1] = step<OPB>( acc[i + 1], a, b );
        }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for( int i = 0; i < ILP; i++ )
        s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if( threadIdx.x == 0 )
        clk[blockIdx.x] = (unsigned long long)( t1 - t0 );
}

template<int OPA, int OPB>
static void run_pair( int sms, uint32_t *out, unsigned long long *clk, unsigned long long *clk_h, int last )
{
    const int grid = sms * 2;
    pair_kernel<OPA, OPB><<<grid, 1024>>>( out, clk, 12345u );
    cudaDeviceSynchronize();
    double best_cyc = 1e30;
    for( int rep = 0; rep < 5; rep++ )
    {
        pair_kernel<OPA, OPB><<<grid, 1024>>>( out, clk, 12345u + rep );
        cudaDeviceSynchronize();
        cudaMemcpy( clk_h, clk, grid * sizeof( *clk ), cudaMemcpyDeviceToHost );
        double mx = 0;
        for( int i = 0; i < grid; i++ )
            if( (double)clk_h[i] > mx )
                mx = (double)clk_h[i];
        if( mx < best_cyc )
            best_cyc = mx;
    }
    const double inst_sm = 2.0 * 1024 * ITERS * ILP;
    printf( "  {\"pair\": [\"%s\", \"%s\"], \"thread_inst_per_clk_per_sm\": %.1f, \"warp_inst_per_clk_per_sm\": %.2f}%s\n",
            op_name[OPA], op_name[OPB], inst_sm / best_cyc, inst_sm / best_cyc / 32.0, last ? "" : "," );
}

template<int OP>
static void run( int sms, int khz, uint32_t *out, unsigned long long *clk, unsigned long long *clk_h, int last )
{
    const int grid = sms * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate( &e0 );
    cudaEventCreate( &e1 );
    pipe_kernel<OP><<<grid, 1024>>>( out, clk, 12345u );   // warm-up
    cudaDeviceSynchronize();
    float best_ms = 1e30f;
    double best_cyc = 1e30;
    for( int rep = 0; rep < 5; rep++ )
    {
        cudaEventRecord( e0 );
        pipe_kernel<OP><<<grid, 1024>>>( out, clk, 12345u + rep );
        cudaEventRecord( e1 );
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime( &ms, e0, e1 );
        cudaMemcpy( clk_h, clk, grid * sizeof( *clk ), cudaMemcpyDeviceToHost );
        double mx = 0;
        for( int i = 0; i < grid; i++ )
            if( (double)clk_h[i] > mx )
                mx = (double)clk_h[i];
        if( ms < best_ms )
            best_ms = ms;
        if( mx < best_cyc )
            best_cyc = mx;
    }
    // thread-instructions per SM: 2 CTAs x 1024 threads x ITERS x ILP
    const double inst_sm = 2.0 * 1024 * ITERS * ILP;
    const double per_clk_incl = inst_sm / best_cyc;                              // in-kernel clock64 (clock independent)
    const double per_clk_event = inst_sm / ( best_ms * 1e-3 * khz * 1e3 );      // event time x nominal max clock
    printf( "  {\"op\": \"%s\", \"thread_inst_per_clk_per_sm\": %.1f, \"warp_inst_per_clk_per_sm\": %.2f, "
            "\"by_event_time_at_max_clock\": %.1f, \"ms\": %.4f}%s\n",
            op_name[OP], per_clk_incl, per_clk_incl / 32.0, per_clk_event, best_ms, last ? "" : "," );
}

int main()
{
    cudaDeviceProp p;
    if( cudaGetDeviceProperties( &p, 0 ) != cudaSuccess )
    {
        fprintf( stderr, "no CUDA device\n" );
        return 1;
    }
    int khz = 0;
    cudaDeviceGetAttribute( &khz, cudaDevAttrClockRate, 0 );
    const int sms = p.multiProcessorCount;
    uint32_t *out;
    unsigned long long *clk, *clk_h;
    cudaMalloc( &out, (size_t)sms * 2 * 1024 * 4 );
    cudaMalloc( &clk, (size_t)sms * 2 * 8 );
    clk_h = (unsigned long long *)malloc( (size_t)sms * 2 * 8 );
    printf( "{\"device\": \"%s\", \"sms\": %d, \"max_clock_khz\": %d, \"ilp\": %d, \"threads_per_sm\": 2048, \"ops\": [\n",
            p.name, sms, khz, ILP );
    run<OP_VABSDIFF4_ACC>( sms, khz, out, clk, clk_h, 0 );
    run<OP_VABSDIFF4>( sms, khz, out, clk, clk_h, 0 );
    run<OP_IDP4A>( sms, khz, out, clk, clk_h, 0 );
    run<OP_VIADD16X2>( sms, khz, out, clk, clk_h, 0 );
    run<OP_IADD3>( sms, khz, out, clk, clk_h, 0 );
    run<OP_LOP3>( sms, khz, out, clk, clk_h, 0 );
    run<OP_PRMT>( sms, khz, out, clk, clk_h, 0 );
    run<OP_VIMNMX>( sms, khz, out, clk, clk_h, 0 );
    run<OP_IMAD>( sms, khz, out, clk, clk_h, 0 );
    run<OP_SHFL>( sms, khz, out, clk, clk_h, 0 );
    run<OP_SHF>( sms, khz, out, clk, clk_h, 0 );
    run<OP_MIX_SAD>( sms, khz, out, clk, clk_h, 1 );
    printf( "], \"pairs\": [\n" );
    run_pair<OP_IMAD, OP_IADD3>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IMAD, OP_VABSDIFF4_ACC>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IMAD, OP_VIADD16X2>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IMAD, OP_PRMT>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IMAD, OP_VIMNMX>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IMAD, OP_IDP4A>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IDP4A, OP_LOP3>( sms, out, clk, clk_h, 0 );
    run_pair<OP_IDP4A, OP_SHF>( sms, out, clk, clk_h, 0 );
    run_pair<OP_VABSDIFF4_ACC, OP_IADD3>( sms, out, clk, clk_h, 0 );
    run_pair<OP_VABSDIFF4_ACC, OP_SHF>( sms, out, clk, clk_h, 0 );
    run_pair<OP_SHFL, OP_IADD3>( sms, out, clk, clk_h, 0 );
    run_pair<OP_SHFL, OP_IMAD>( sms, out, clk, clk_h, 1 );
    printf( "]}\n" );
    return 0;
}
