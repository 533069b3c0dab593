// device_reduce.cuh — deterministic block reduction + "last block finalises" grid reduction.
//
// Every block reduces its threads' values in a fixed tree (warp shuffles, then one warp over the
// per-warp sums), writes one partial per value, and takes a ticket.  The block that draws the
// last ticket re-reads all partials in index order and reduces them with the same fixed tree, so
// for a given launch configuration the result is bit-reproducible run to run (no floating-point
// atomics anywhere).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double warp_sum( double v )
{
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 )
        v += __shfl_down_sync( 0xffffffffu, v, o );
    return v;
}

// Sum over the block; result valid in thread 0.  NT = blockDim.x (multiple of 32, <= 1024).
template <int NT>
__device__ __forceinline__ double block_sum( double v, double* smem /* >= NT/32 doubles */ )
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum( v );
    __syncthreads(); // smem may still be in use by a previous call
    if ( lane == 0 )
        smem[wid] = v;
    __syncthreads();
    if ( wid == 0 )
    {
        v = lane < NT / 32 ? smem[lane] : 0.0;
        v = warp_sum( v );
    }
    return v;
}

// Block partials -> global partial arrays (value n at partials[n * stride + block]) -> the last
// block sums them.  Returns true in every thread of the last block; vals[] then holds the grid
// totals in thread 0.  The ticket is reset for the next launch.
template <int NT, int NV>
__device__ __forceinline__ bool block_reduce_finalize( double vals[NV], double* partials, int stride,
                                                       unsigned int* ticket )
{
    __shared__ double s_red[NT / 32];
    __shared__ bool s_last;
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = ( blockIdx.z * gridDim.y + blockIdx.y ) * gridDim.x + blockIdx.x;
#pragma unroll
    for ( int n = 0; n < NV; ++n )
    {
        double s = block_sum<NT>( vals[n], s_red );
        if ( threadIdx.x == 0 )
            partials[n * stride + bid] = s;
    }
    if ( threadIdx.x == 0 )
    {
        __threadfence();
        unsigned t = atomicAdd( ticket, 1u );
        s_last = ( t == nblocks - 1 );
    }
    __syncthreads();
    if ( !s_last )
        return false;
    __threadfence();
#pragma unroll
    for ( int n = 0; n < NV; ++n )
    {
        double s = 0.0;
        for ( unsigned b = threadIdx.x; b < nblocks; b += NT )
            s += __ldcg( partials + n * stride + b );
        vals[n] = block_sum<NT>( s, s_red );
    }
    if ( threadIdx.x == 0 )
        *ticket = 0u;
    return true;
}
