// device_geo.cuh — device-side geometry, boundary test and B-spline sampling helpers.
//
// Cajita pieces restated from the published algorithm (the library is not in the reference
// tree): LocalMesh::coordinates, Spline<1>/Spline<3>, evaluateSpline, G2P::value.
// The expression order below is deliberately identical to the statement order the checker
// (oracle/) uses, and the library is compiled with -fmad=false, so element-wise results are
// bit-comparable.
#pragma once
#include "cfb_internal.h"

__device__ __forceinline__ long long geo_off( const Geo& g, int i, int j, int k )
{
    return g.origin + (long long)k * g.sz + (long long)j * g.sy + i;
}

// LocalMesh::coordinates( entity, local_index, x ); `idx` is the OWNED index, the reference's
// local (ghosted) index is idx + halo.  ent: 0 = Cell, 1 + d = Face<d>.
template <int D>
__device__ __forceinline__ void geo_coordinates( const Geo& g, int ent, const int idx[3], double x[3] )
{
#pragma unroll
    for ( int d = 0; d < D; ++d )
    {
        const double l = (double)( idx[d] + g.h );
        x[d] = ( ent - 1 == d ) ? g.ghost_low[d] + l * g.celld[d] : g.ghost_low[d] + ( l + 0.5 ) * g.celld[d];
    }
}

// InflowSource box test (src/InflowSource.hpp:40-41), z added for D == 3.
template <int D>
__device__ __forceinline__ bool in_box( const InflowConst& s, const double x[3] )
{
    bool in = true;
#pragma unroll
    for ( int d = 0; d < D; ++d )
        in = in && ( x[d] >= s.lo[d] && x[d] < s.hi[d] );
    return in;
}

// BoundaryCondition::operator()( Face<d>, ... )  src/BoundaryConditions.hpp:102-129:
// zero the wall-normal velocity on SOLID walls; `gi` is the global index along d.
// bc.min = 0, bc.max = global cells - 1 (src/Solver.hpp:109-110).
__device__ __forceinline__ double bc_face( const Geo& g, int d, int gi, double val )
{
    if ( gi <= 0 && g.bt[d] == CFB_SOLID )
        val = 0.0;
    if ( gi > g.gn[d] - 1 && g.bt[3 + d] == CFB_SOLID )
        val = 0.0;
    return val;
}

// Number of SOLID walls a cell touches along dim d (BoundaryCondition::build_matrix,
// src/BoundaryConditions.hpp:56-97: `gi <= min` / `gi > max - 1`).
__device__ __forceinline__ int wall_count( const Geo& g, int d, int gi )
{
    int c = 0;
    if ( gi <= 0 && g.bt[d] == CFB_SOLID )
        ++c;
    if ( gi > g.gn[d] - 2 && g.bt[3 + d] == CFB_SOLID )
        ++c;
    return c;
}

// Cajita::Spline<1>: stencil int(xl), int(xl)+1; weights 1-f, f.
__device__ __forceinline__ void spline1( double xl, int& s0, double w[2] )
{
    const int i0 = (int)xl;
    s0 = i0;
    const double xn = xl - (double)i0;
    w[0] = 1.0 - xn;
    w[1] = xn;
}

// Cajita::Spline<3>: stencil int(xl)-1 .. int(xl)+2; cubic B-spline weights from the distance to
// the first knot (xn = f + 1), stepping xn -= 1 per knot.
__device__ __forceinline__ void spline3( double xl, int& s0, double w[4] )
{
    const int i0 = (int)xl;
    s0 = i0 - 1;
    const double one_sixth = 1.0 / 6.0;
    const double two_thirds = one_sixth * 4.0;
    const double four_thirds = 2.0 * two_thirds;
    double xn = xl - (double)i0 + 1.0;
    double xn2 = xn * xn;
    w[0] = -xn * xn2 * one_sixth + xn2 - 2.0 * xn + four_thirds;
    xn -= 1.0;
    xn2 = xn * xn;
    w[1] = 0.5 * xn * xn2 - xn2 + two_thirds;
    xn -= 1.0;
    xn2 = xn * xn;
    w[2] = -0.5 * xn * xn2 - xn2 + two_thirds;
    xn -= 1.0;
    xn2 = xn * xn;
    w[3] = xn * xn2 * one_sixth + xn2 + 2.0 * xn + four_thirds;
}

// Interpolation::interpolateField<D, ORDER, Entity>  src/Interpolation.hpp:30-41.
// Stencil indices are clamped into the ghosted allocation (the reference does not clamp: leaving
// the halo is undefined behaviour there; parity is defined for CFL <= 1 only).
// Fast path (the normal case, CFL <= 1): no stencil index is clamped, so the (ORDER+1)^D samples sit at
// base + a + b * sy + c * sz — one 64-bit offset per row instead of a clamp and a full index computation per
// sample (the kernel is instruction-issue bound: profiles/r2_advect_*).  Same loads, same products, same
// summation order as the clamped path, hence the same bits.
template <int D, int ORDER>
__device__ __forceinline__ double interp_field( const Geo& g, int ent, const double* __restrict__ f,
                                                const double loc[3] )
{
    constexpr int NK = ORDER + 1;
    int s0[3] = { 0, 0, 0 };
    double w[3][NK];
    bool inside = true;
#pragma unroll
    for ( int d = 0; d < D; ++d )
    {
        // position of local entity 0: coordinates( entity, {0,0,0} )
        const double low = ( ent - 1 == d ) ? g.ghost_low[d] + 0.0 * g.celld[d]
                                            : g.ghost_low[d] + ( 0.0 + 0.5 ) * g.celld[d];
        const double xl = ( loc[d] - low ) * g.rdxd[d];
        if ( ORDER == 1 )
            spline1( xl, s0[d], w[d] );
        else
            spline3( xl, s0[d], w[d] );
        const int emax = g.n[d] + 2 * g.h + ( ent - 1 == d ? 1 : 0 ) - 1;
        // 0 <= s0 and s0 + NK - 1 <= emax, written so that no garbage index (the int conversion of a NaN or of a
        // huge coordinate, and s0 = int - 1 wrapping around) can pass: such points take the clamped path
        const int hi = emax - ( NK - 1 );
        inside = inside && hi >= 0 && (unsigned)s0[d] <= (unsigned)hi;
    }
    double value = 0.0;
    if ( inside )
    {
        // local ghosted index -> owned index: - halo.  One pointer per (b, c) row of the stencil box, the samples
        // of a row at immediate offsets from it.
        const double* base = f + geo_off( g, s0[0] - g.h, s0[1] - g.h, D == 3 ? s0[2] - g.h : 0 );
        const int rsy = (int)g.sy, rsz = (int)g.sz; // row / plane strides fit 32 bits (asserted in cfb_create)
        if ( D == 2 )
        {
            const double* row[NK];
#pragma unroll
            for ( int b = 0; b < NK; ++b )
                row[b] = base + (unsigned)( b * rsy );
#pragma unroll
            for ( int a = 0; a < NK; ++a )
#pragma unroll
                for ( int b = 0; b < NK; ++b )
                    value += __ldg( row[b] + a ) * w[0][a] * w[1][b];
        }
        else
        {
            const double* row[NK][NK];
#pragma unroll
            for ( int b = 0; b < NK; ++b )
#pragma unroll
                for ( int c = 0; c < NK; ++c )
                    row[b][c] = base + (unsigned)( c * rsz + b * rsy );
#pragma unroll
            for ( int a = 0; a < NK; ++a )
#pragma unroll
                for ( int b = 0; b < NK; ++b )
#pragma unroll
                    for ( int c = 0; c < NK; ++c )
                        value += __ldg( row[b][c] + a ) * w[0][a] * w[1][b] * w[2][c];
        }
        return value;
    }
    // Stencil indices clamped into the ghosted allocation (the reference does not clamp: leaving the halo is
    // undefined behaviour there; parity is defined for CFL <= 1 only).
    int s[3][NK];
#pragma unroll
    for ( int d = 0; d < D; ++d )
    {
        const int emax = g.n[d] + 2 * g.h + ( ent - 1 == d ? 1 : 0 ) - 1;
#pragma unroll
        for ( int a = 0; a < NK; ++a )
        {
            int si = s0[d] + a;
            si = si < 0 ? 0 : ( si > emax ? emax : si );
            s[d][a] = si - g.h;
        }
    }
    if ( D == 2 )
    {
#pragma unroll
        for ( int a = 0; a < NK; ++a )
#pragma unroll
            for ( int b = 0; b < NK; ++b )
                value += __ldg( f + geo_off( g, s[0][a], s[1][b], 0 ) ) * w[0][a] * w[1][b];
    }
    else
    {
#pragma unroll
        for ( int a = 0; a < NK; ++a )
#pragma unroll
            for ( int b = 0; b < NK; ++b )
#pragma unroll
                for ( int c = 0; c < NK; ++c )
                    value += __ldg( f + geo_off( g, s[0][a], s[1][b], s[2][c] ) ) * w[0][a] * w[1][b] *
                             w[2][c];
    }
    return value;
}

// Interpolation::interpolateVelocity<D, 1>  src/Interpolation.hpp:43-54 (+ w for D == 3).
template <int D>
__device__ __forceinline__ void interp_velocity( const Geo& g, const double* const vel[4],
                                                 const double loc[3], double out[3] )
{
#pragma unroll
    for ( int d = 0; d < D; ++d )
        out[d] = interp_field<D, 1>( g, 1 + d, vel[1 + d], loc );
}

// ---- the 7-point row of A, in the reference's stencil order ---------------------------------
// {0}, {-x}, {+x}, {-y}, {+y}, {-z}, {+z}; one fused multiply-add per term (bit-identical to the
// checker's apply_A).
__device__ __forceinline__ double apply_row( double diag, double ns, double c, double xm, double xp,
                                             double ym, double yp, double zm, double zp )
{
    double a = diag * c;
    a = fma( ns, xm, a );
    a = fma( ns, xp, a );
    a = fma( ns, ym, a );
    a = fma( ns, yp, a );
    a = fma( ns, zm, a );
    a = fma( ns, zp, a );
    return a;
}
