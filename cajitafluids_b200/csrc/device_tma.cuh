// device_tma.cuh — PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) and the 7-point row of the
// pressure operator, shared by the stencil kernels (kernels_stencil.cu, kernels_fused.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32( const void* p )
{
    return (uint32_t)__cvta_generic_to_shared( p );
}
__device__ __forceinline__ void mbar_init( uint32_t bar, uint32_t count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( bar ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
}
__device__ __forceinline__ void mbar_expect_tx( uint32_t bar, uint32_t bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( bar ), "r"( bytes )
                  : "memory" );
}
__device__ __forceinline__ bool mbar_try_wait( uint32_t bar, uint32_t parity )
{
    uint32_t ok;
    asm volatile( "{\n"
                  ".reg .pred P1;\n"
                  "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
                  "selp.u32 %0, 1, 0, P1;\n"
                  "}"
                  : "=r"( ok )
                  : "r"( bar ), "r"( parity )
                  : "memory" );
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait( uint32_t bar, uint32_t parity )
{
    uint32_t spins = 0;
    while ( !mbar_try_wait( bar, parity ) )
        if ( ++spins > ( 1u << 22 ) )
            __trap();
}
__device__ __forceinline__ void tma_load_3d( uint32_t dst, const CUtensorMap* map, uint32_t bar, int x,
                                             int y, int z )
{
    asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
                  "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"( dst ),
                  "l"( map ), "r"( bar ), "r"( x ), "r"( y ), "r"( z )
                  : "memory" );
}
__device__ __forceinline__ void prefetch_tmap( const CUtensorMap* map )
{
    asm volatile( "prefetch.tensormap [%0];" ::"l"( map ) : "memory" );
}

