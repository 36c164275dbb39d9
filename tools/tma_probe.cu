// tma_probe.cu -- the smallest program that does what xd_hpel_tma_kernel's row source does: a 3-D byte tensor (byte in row,
// row, slot), one cp.async.bulk.tensor.3d of 256 x 8 bytes into shared memory, completion on an mbarrier.  Prints what the
// driver and the device say.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

typedef CUresult ( *encode_fn )( CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill );

__global__ void probe( const __grid_constant__ CUtensorMap tmap, int x, int y, int z, uint8_t *out, int form, int tx )
{
    __shared__ __align__( 128 ) uint8_t tile[2048];
    __shared__ __align__( 8 ) uint64_t bar;
    const int lane = threadIdx.x;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared( &bar ), d = (uint32_t)__cvta_generic_to_shared( tile );
    if( lane == 0 )
    {
        asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"( b ) : "memory" );
        if( form != 11 )
            asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
        if( form != 12 )
            asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
        if( form >= 10 && form <= 12 )
            ;
        else
        asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( b ), "r"( tx ) : "memory" );
        if( form >= 10 && form != 30 )
            ;
        else if( form == 30 )
            asm volatile( "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                          :: "r"( d ), "l"( (uint64_t)&tmap ), "r"( b ), "r"( x ), "r"( y ) : "memory" );
        else if( form == 0 )
            asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                          :: "r"( d ), "l"( (uint64_t)&tmap ), "r"( b ), "r"( x ), "r"( y ), "r"( z ) : "memory" );
        else
            asm volatile( "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                          :: "r"( d ), "l"( (uint64_t)&tmap ), "r"( b ), "r"( x ), "r"( y ), "r"( z ) : "memory" );
    }
    __syncwarp();
    uint32_t ok = 0, spins = 0;
    while( ( form < 10 || form == 30 ) && !ok && spins < ( 1u << 22 ) )
    {
        asm volatile( "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                      : "=r"( ok ) : "r"( b ), "r"( 0 ) : "memory" );
        spins++;
    }
    for( int i = lane; i < 2048; i += 32 )
        out[i] = ok ? tile[i] : 0xEE;
    if( lane == 0 )
        out[2048] = (uint8_t)ok;
}

int main( int argc, char **argv )
{
    const int only = argc > 1 ? atoi( argv[1] ) : 0;
    const int stride = 416, rows = 360, slots = 3;
    const size_t slot_bytes = ( (size_t)stride * rows * 4 + 255 ) / 256 * 256;
    uint8_t *h = (uint8_t *)malloc( slot_bytes * slots ), *d, *d_out, h_out[2049];
    for( size_t i = 0; i < slot_bytes * slots; i++ )
        h[i] = (uint8_t)( i * 2654435761u >> 13 );
    cudaMalloc( &d, slot_bytes * slots );
    cudaMalloc( &d_out, 2049 );
    cudaMemcpy( d, h, slot_bytes * slots, cudaMemcpyHostToDevice );
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q );
    printf( "entry point: %s, query %d, fn %p\n", cudaGetErrorString( e ), (int)q, fn );
    CUtensorMap tmap;
    const cuuint64_t dims[3] = { (cuuint64_t)stride, (cuuint64_t)rows, (cuuint64_t)slots };
    const cuuint64_t strides[2] = { (cuuint64_t)stride, (cuuint64_t)slot_bytes };
    const int bw = argc > 2 ? atoi( argv[2] ) : 256;
    const cuuint32_t box[3] = { (cuuint32_t)bw, 8, 1 }, estr[3] = { 1, 1, 1 };
    CUresult r = ( (encode_fn)fn )( &tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, only == 30 ? 2 : 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
    printf( "encode: %d\n", (int)r );
    for( int form = only; form <= only; form++ )
    {
        const int x = argc > 3 ? atoi( argv[3] ) : 24, y = 22, z = only == 30 ? 0 : 1;
        probe<<<1, 32>>>( tmap, x, y, z, d_out, form, bw * 8 );
        e = cudaDeviceSynchronize();
        printf( "form %d (%s): %s\n", form, form ? "shared::cta" : "shared::cluster", cudaGetErrorString( e ) );
        if( e != cudaSuccess )
            return 1;
        cudaMemcpy( h_out, d_out, 2049, cudaMemcpyDeviceToHost );
        int bad = 0;
        for( int r8 = 0; r8 < 8; r8++ )
            for( int c = 0; c < bw; c++ )
                bad += h_out[r8 * bw + c] != h[z * slot_bytes + (size_t)( y + r8 ) * stride + x + c];
        printf( "  landed %d, %d bytes differ from the plane\n", h_out[2048], bad );
    }
    return 0;
}
