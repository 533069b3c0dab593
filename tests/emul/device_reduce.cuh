// device_reduce.cuh — HOST STAND-IN (tests/emul only) for the product's grid reductions.
//
// Same interface and the same arithmetic (double-double accumulation, one rounding at the end); the
// warp-shuffle / shared-memory tree is replaced by a running sum, which is legitimate because the
// result — the correctly rounded exact sum — does not depend on the order (that property is what the
// GPU parity tests check on the real helper).  Relies on the emulated launch order: inside a block
// thread 0 runs last, blocks run in order.
#pragma once
#include <cuda_runtime.h>

struct dd_t
{
    double hi, lo;
};
inline void dd_acc( dd_t& a, double x )
{
    const double s = a.hi + x;
    const double bb = s - a.hi;
    const double e = ( a.hi - ( s - bb ) ) + ( x - bb );
    a.hi = s;
    a.lo += e;
}
inline dd_t dd_add( dd_t a, dd_t b )
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

// Sum over the block, cooperative launches only (one fiber per CUDA thread, real barriers): between two
// barriers the fibers run one after the other, so a shared running sum is race-free.  Valid in every thread.
template <int NT>
inline dd_t dd_block_sum( dd_t v, dd_t* /*smem*/ )
{
    static thread_local dd_t acc;
    __syncthreads(); // a previous call's result has been read by everybody
    if ( threadIdx.x == 0 )
        acc = dd_t{ 0.0, 0.0 };
    __syncthreads();
    acc = dd_add( acc, v );
    __syncthreads();
    return acc;
}

template <int NT, int NV>
inline bool block_reduce_finalize( dd_t vals[NV], double* partials, int stride, unsigned int* ticket )
{
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = ( blockIdx.z * gridDim.y + blockIdx.y ) * gridDim.x + blockIdx.x;
    if ( cfb_emul::coop_active() )
    {
        // the product's structure: block sums, partials, ticket, the last block adds the partials
        static thread_local bool s_last;
        for ( int n = 0; n < NV; ++n )
        {
            const dd_t s = dd_block_sum<NT>( vals[n], nullptr );
            if ( threadIdx.x == 0 )
            {
                partials[( (size_t)n * stride + bid ) * 2 + 0] = s.hi;
                partials[( (size_t)n * stride + bid ) * 2 + 1] = s.lo;
            }
        }
        if ( threadIdx.x == 0 )
            s_last = atomicAdd( ticket, 1u ) == nblocks - 1;
        __syncthreads();
        if ( !s_last )
            return false;
        if ( threadIdx.x == 0 )
        {
            for ( int n = 0; n < NV; ++n )
            {
                dd_t s{ 0.0, 0.0 };
                for ( unsigned b = 0; b < nblocks; ++b )
                    s = dd_add( s, dd_t{ partials[( (size_t)n * stride + b ) * 2 + 0], partials[( (size_t)n * stride + b ) * 2 + 1] } );
                vals[n] = s;
            }
            *ticket = 0u;
        }
        return true;
    }
    // sequential launches: thread 0 of a block runs last
    static thread_local dd_t acc[NV];
    if ( threadIdx.x == blockDim.x - 1 ) // first thread of the block to run
        for ( int n = 0; n < NV; ++n )
            acc[n] = dd_t{ 0.0, 0.0 };
    for ( int n = 0; n < NV; ++n )
        acc[n] = dd_add( acc[n], vals[n] );
    if ( threadIdx.x != 0 )
        return false;
    for ( int n = 0; n < NV; ++n )
    {
        partials[( (size_t)n * stride + bid ) * 2 + 0] = acc[n].hi;
        partials[( (size_t)n * stride + bid ) * 2 + 1] = acc[n].lo;
    }
    const unsigned t = atomicAdd( ticket, 1u );
    if ( t != nblocks - 1 )
        return false;
    for ( int n = 0; n < NV; ++n )
    {
        dd_t s{ 0.0, 0.0 };
        for ( unsigned b = 0; b < nblocks; ++b )
            s = dd_add( s, dd_t{ partials[( (size_t)n * stride + b ) * 2 + 0], partials[( (size_t)n * stride + b ) * 2 + 1] } );
        vals[n] = s;
    }
    *ticket = 0u;
    return true;
}
