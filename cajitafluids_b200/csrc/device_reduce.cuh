// device_reduce.cuh — grid reductions whose result does not depend on the summation order (to ~2^-104).
//
// The CG scalars (r.r, z.r, p.Ap) steer the whole solve; a 1-ulp difference in one of them is
// amplified by ~1e4-1e5 over a few hundred iterations.  The reference leaves their summation order
// to Kokkos::parallel_reduce.  Here every thread accumulates the (double-rounded) products in
// double-double (TwoSum error-free transformation; the low word is a plain sum of the error terms, so
// the pair carries ~106 bits: relative error ~2^-104), warps / blocks / grid combine double-doubles, and
// only the final value is rounded to double.  That is not an exact (Kulisch) accumulator: two orders of
// summation can differ in the ~106th bit and hence, with probability ~2^-50 per sum, in the rounded
// double.  In practice — and in every test, 1 to 8 blocks, every tiling — thread count, tile shape,
// launch configuration and block decomposition do not change a bit of alpha / beta, and the CPU checker,
// which accumulates the same way, agrees bit for bit; the tests assert the stated bar (iterations +-1,
// 1e-10) first and bit-identity second.  Cost: 7 FP64 adds per term, invisible in HBM-bound kernels.
// No floating-point atomics anywhere.
#pragma once
#include <cuda_runtime.h>

struct dd_t
{
    double hi, lo;
};

// a += x   (x a plain double)
__device__ __forceinline__ void dd_acc( dd_t& a, double x )
{
    const double s = a.hi + x;
    const double bb = s - a.hi;
    const double e = ( a.hi - ( s - bb ) ) + ( x - bb );
    a.hi = s;
    a.lo += e;
}

// a + b, renormalised
__device__ __forceinline__ dd_t dd_add( dd_t a, dd_t b )
{
    const double s = a.hi + b.hi;
    const double bb = s - a.hi;
    double e = ( a.hi - ( s - bb ) ) + ( b.hi - bb );
    e += a.lo + b.lo;
    dd_t r;
    r.hi = s + e;
    r.lo = e - ( r.hi - s );
    return r;
}

__device__ __forceinline__ dd_t dd_warp_sum( dd_t v )
{
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 )
    {
        dd_t w;
        w.hi = __shfl_down_sync( 0xffffffffu, v.hi, o );
        w.lo = __shfl_down_sync( 0xffffffffu, v.lo, o );
        v = dd_add( v, w );
    }
    return v;
}

// Sum over the block; result valid in thread 0.  NT = blockDim.x (multiple of 32, <= 1024).
template <int NT>
__device__ __forceinline__ dd_t dd_block_sum( dd_t v, dd_t* smem /* >= NT/32 entries */ )
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = dd_warp_sum( v );
    __syncthreads(); // smem may still be in use by a previous call
    if ( lane == 0 )
        smem[wid] = v;
    __syncthreads();
    if ( wid == 0 )
    {
        if ( lane < NT / 32 )
            v = smem[lane];
        else
            v.hi = v.lo = 0.0;
        v = dd_warp_sum( v );
    }
    return v;
}

// Block partials -> global (value n of block b at partials[(n * stride + b) * 2 + {0,1}]) -> the
// block that draws the last ticket sums them.  Returns true in every thread of that block; vals[]
// then holds the grid totals (double-double) in thread 0.  The ticket is reset for the next launch.
template <int NT, int NV>
__device__ __forceinline__ bool block_reduce_finalize( dd_t vals[NV], double* partials, int stride,
                                                       unsigned int* ticket )
{
    __shared__ dd_t s_red[NT / 32];
    __shared__ bool s_last;
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = ( blockIdx.z * gridDim.y + blockIdx.y ) * gridDim.x + blockIdx.x;
#pragma unroll
    for ( int n = 0; n < NV; ++n )
    {
        dd_t s = dd_block_sum<NT>( vals[n], s_red );
        if ( threadIdx.x == 0 )
        {
            partials[( (size_t)n * stride + bid ) * 2 + 0] = s.hi;
            partials[( (size_t)n * stride + bid ) * 2 + 1] = s.lo;
        }
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
        dd_t s;
        s.hi = s.lo = 0.0;
        for ( unsigned b = threadIdx.x; b < nblocks; b += NT )
        {
            dd_t w;
            w.hi = __ldcg( partials + ( (size_t)n * stride + b ) * 2 + 0 );
            w.lo = __ldcg( partials + ( (size_t)n * stride + b ) * 2 + 1 );
            s = dd_add( s, w );
        }
        vals[n] = dd_block_sum<NT>( s, s_red );
    }
    if ( threadIdx.x == 0 )
        *ticket = 0u;
    return true;
}
