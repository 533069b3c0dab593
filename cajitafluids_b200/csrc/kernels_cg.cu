// kernels_cg.cu — streaming kernels of the Jacobi-preconditioned CG (everything except q = A p).
//
// Restates Cajita::ReferenceConjugateGradient::solve (driven from
// src/VelocityCorrector.hpp:276 with the diagonal preconditioner of :166-179) in three fused
// kernels per iteration instead of four, with all scalars kept on the device:
//
//   cg_axpy     kernel 1 (x += a p, r -= a q, sum r^2) + the reduction of kernel 2 (sum r.M^-1 r)
//   cg_pupdate  convergence test of the iteration + kernel 3 (p = M^-1 r + b p); z is never stored
//   stencil_dot kernel 4 (kernels_stencil.cu)
//
// Matrix and preconditioner arrays do not exist: the diagonal is a function of how many SOLID
// walls a cell touches (src/BoundaryConditions.hpp:56-97), OpConst holds the 7 possible values.
//
// Algorithmic bytes per owned cell: cg_axpy 48 (r x,p,r,q; w x,r), cg_pupdate 24 (r r,p; w p),
// divergence+init 56+8.  All bound by HBM bandwidth.
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"
#include "device_cg1.cuh"

namespace
{

constexpr int NT = 256;

// Decode pair index t -> owned (i, j, k), i even.  npx = ceil(nx / 2).
__device__ __forceinline__ void pair_decode( const Geo& g, unsigned t, unsigned npx, int& i, int& j, int& k )
{
    unsigned row = t / npx;
    i = 2 * (int)( t - row * npx );
    k = (int)( row / (unsigned)g.n[1] );
    j = (int)( row - (unsigned)k * (unsigned)g.n[1] );
}

// SOLID-wall count of the two cells of a pair.
__device__ __forceinline__ void pair_walls( const Geo& g, int i, int j, int k, int& c0, int& c1 )
{
    const int cyz = wall_count( g, 1, j + g.off[1] ) + wall_count( g, 2, k + g.off[2] );
    c0 = cyz + wall_count( g, 0, i + g.off[0] );
    c1 = cyz + wall_count( g, 0, i + 1 + g.off[0] );
}

// Single block: the rounded sums are final.  Multi-GPU: keep the local double-doubles; they are
// all-gathered and combined exactly (in rank order) by cg_combine_kernel, so the global value is
// again the correctly rounded exact sum, whatever the decomposition.
__device__ __forceinline__ void publish_rr_rz( CgState* S, dd_t rr, dd_t rz )
{
    if ( S->world > 1 )
    {
        S->loc[2] = rz.hi;
        S->loc[3] = rz.lo;
        S->loc[4] = rr.hi;
        S->loc[5] = rr.lo;
    }
    else
    {
        S->rr = rr.hi + rr.lo;
        S->rz_new = rz.hi + rz.lo;
    }
}

// which == 0: pAp from gath[rank][0..1] ; which == 1: (rz_new, rr) from gath[rank][0..3] ;
// which == 2 / 3: (r.r, r.u, w.u) of the single-reduction form from gath[rank][0..5], then its scalar step (3: the
// launch that starts a solve)
__global__ void cg_combine_kernel( CgState* S, int which )
{
    const int nv = which == 0 ? 1 : ( which == 1 ? 2 : 3 );
    dd_t acc[3] = { { 0.0, 0.0 }, { 0.0, 0.0 }, { 0.0, 0.0 } };
    for ( int r = 0; r < S->world; ++r )
        for ( int v = 0; v < nv; ++v )
        {
            dd_t w = { S->gath[( r * nv + v ) * 2], S->gath[( r * nv + v ) * 2 + 1] };
            acc[v] = dd_add( acc[v], w );
        }
    if ( which == 0 )
        S->pAp = acc[0].hi + acc[0].lo;
    else if ( which == 1 )
    {
        S->rz_new = acc[0].hi + acc[0].lo;
        S->rr = acc[1].hi + acc[1].lo;
    }
    else if ( !S->done ) // (a stencil launch behind convergence did nothing: no scalar step either)
        cg1_finish( S, acc[0].hi + acc[0].lo, acc[1].hi + acc[1].lo, acc[2].hi + acc[2].lo, which == 3 );
}

// ---------------------------------------------------------------------------------------------
// VelocityCorrector::_buildRHS (src/VelocityCorrector.hpp:204-210) + lhs = 0 (:272)
template <int D>
__global__ void __launch_bounds__( NT )
    divergence_kernel( const __grid_constant__ Geo g, const double* __restrict__ u,
                       const double* __restrict__ v, const double* __restrict__ w,
                       double* __restrict__ rhs, double* __restrict__ lhs )
{
    const long long total = (long long)g.n[0] * g.n[1] * g.n[2];
    const double scale = 1.0 / g.cell;
    for ( long long t = blockIdx.x * (long long)NT + threadIdx.x; t < total; t += (long long)gridDim.x * NT )
    {
        const int i = (int)( t % g.n[0] );
        const int j = (int)( ( t / g.n[0] ) % g.n[1] );
        const int k = (int)( t / ( (long long)g.n[0] * g.n[1] ) );
        const long long o = geo_off( g, i, j, k );
        double div = u[o + 1] - u[o] + v[o + g.sy] - v[o];
        if ( D == 3 )
            div = div + w[o + g.sz] - w[o];
        rhs[o] = -scale * div;
        lhs[o] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Start of solve: x0 = 0 => r0 = b ; z0 = M^-1 r0 ; p0 = z0 ; sum r0^2 ; sum z0.r0
__global__ void __launch_bounds__( NT )
    cg_init_kernel( const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                    const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r,
                    double* __restrict__ p, CgState* S, double* partials, int fixed )
{
    const unsigned npx = (unsigned)( ( g.n[0] + 1 ) >> 1 );
    const unsigned total = npx * (unsigned)g.n[1] * (unsigned)g.n[2];
    dd_t rr = { 0.0, 0.0 }, rz = { 0.0, 0.0 };
    for ( unsigned t = blockIdx.x * NT + threadIdx.x; t < total; t += gridDim.x * NT )
    {
        int i, j, k, c0, c1;
        pair_decode( g, t, npx, i, j, k );
        pair_walls( g, i, j, k, c0, c1 );
        const long long o = geo_off( g, i, j, k );
        const bool two = i + 1 < g.n[0];
        if ( two )
        {
            const double2 bv = *reinterpret_cast<const double2*>( b + o );
            const double z0 = op.minv[c0] * bv.x, z1 = op.minv[c1] * bv.y;
            *reinterpret_cast<double2*>( x + o ) = make_double2( 0.0, 0.0 );
            *reinterpret_cast<double2*>( r + o ) = bv;
            *reinterpret_cast<double2*>( p + o ) = make_double2( z0, z1 );
            dd_acc( rr, bv.x * bv.x );
            dd_acc( rr, bv.y * bv.y );
            dd_acc( rz, z0 * bv.x );
            dd_acc( rz, z1 * bv.y );
        }
        else
        {
            const double bv = b[o];
            const double z0 = op.minv[c0] * bv;
            x[o] = 0.0;
            r[o] = bv;
            p[o] = z0;
            dd_acc( rr, bv * bv );
            dd_acc( rz, z0 * bv );
        }
    }
    dd_t vals[2] = { rr, rz };
    if ( block_reduce_finalize<NT, 2>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
        {
            publish_rr_rz( S, vals[0], vals[1] );
            S->iter = 0;
            S->done = 0;
            S->fixed = fixed;
        }
    }
}

// After the (optional) allreduce of (rz_new, rr): pick the threshold, handle r0 already converged.
__global__ void cg_check0_kernel( CgState* S, double tol, int stop_rel )
{
    const double bnorm = sqrt( S->rr );
    S->bnorm = bnorm;
    S->thresh = stop_rel ? tol * bnorm : tol;
    if ( !S->fixed && bnorm <= S->thresh )
        S->done = 1;
}

// ---------------------------------------------------------------------------------------------
// kernel 1 + reduction of kernel 2
__global__ void __launch_bounds__( NT )
    cg_axpy_kernel( const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                    const double* __restrict__ p, const double* __restrict__ q,
                    double* __restrict__ x, double* __restrict__ r, CgState* S, double* partials )
{
    if ( S->done )
        return;
    const double alpha = S->rz_old / S->pAp;
    const double nalpha = -alpha;
    const unsigned npx = (unsigned)( ( g.n[0] + 1 ) >> 1 );
    const unsigned total = npx * (unsigned)g.n[1] * (unsigned)g.n[2];
    dd_t rr = { 0.0, 0.0 }, rz = { 0.0, 0.0 };
    for ( unsigned t = blockIdx.x * NT + threadIdx.x; t < total; t += gridDim.x * NT )
    {
        int i, j, k, c0, c1;
        pair_decode( g, t, npx, i, j, k );
        pair_walls( g, i, j, k, c0, c1 );
        const long long o = geo_off( g, i, j, k );
        if ( i + 1 < g.n[0] )
        {
            const double2 pv = *reinterpret_cast<const double2*>( p + o );
            const double2 qv = *reinterpret_cast<const double2*>( q + o );
            double2 xv = *reinterpret_cast<double2*>( x + o );
            double2 rv = *reinterpret_cast<double2*>( r + o );
            xv.x = fma( alpha, pv.x, xv.x );
            xv.y = fma( alpha, pv.y, xv.y );
            rv.x = fma( nalpha, qv.x, rv.x );
            rv.y = fma( nalpha, qv.y, rv.y );
            *reinterpret_cast<double2*>( x + o ) = xv;
            *reinterpret_cast<double2*>( r + o ) = rv;
            dd_acc( rr, rv.x * rv.x );
            dd_acc( rr, rv.y * rv.y );
            dd_acc( rz, ( op.minv[c0] * rv.x ) * rv.x );
            dd_acc( rz, ( op.minv[c1] * rv.y ) * rv.y );
        }
        else
        {
            const double xv = fma( alpha, p[o], x[o] );
            const double rv = fma( nalpha, q[o], r[o] );
            x[o] = xv;
            r[o] = rv;
            dd_acc( rr, rv * rv );
            dd_acc( rz, ( op.minv[c0] * rv ) * rv );
        }
    }
    dd_t vals[2] = { rr, rz };
    if ( block_reduce_finalize<NT, 2>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
            publish_rr_rz( S, vals[0], vals[1] );
    }
}

// convergence bookkeeping of the iteration + kernel 3
__global__ void __launch_bounds__( NT )
    cg_pupdate_kernel( const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                       const double* __restrict__ r, double* __restrict__ p, CgState* S )
{
    if ( S->done )
        return;
    const double resid = sqrt( S->rr );
    const bool conv = !S->fixed && resid <= S->thresh;
    if ( blockIdx.x == 0 && threadIdx.x == 0 )
    {
        const int it = S->iter;
        if ( it < CFB_HIST_MAX )
            S->hist[it] = resid;
        S->iter = it + 1;
        if ( conv )
            S->done = 1;
    }
    if ( conv )
        return;
    const double beta = S->rz_new / S->rz_old;
    const unsigned npx = (unsigned)( ( g.n[0] + 1 ) >> 1 );
    const unsigned total = npx * (unsigned)g.n[1] * (unsigned)g.n[2];
    for ( unsigned t = blockIdx.x * NT + threadIdx.x; t < total; t += gridDim.x * NT )
    {
        int i, j, k, c0, c1;
        pair_decode( g, t, npx, i, j, k );
        pair_walls( g, i, j, k, c0, c1 );
        const long long o = geo_off( g, i, j, k );
        if ( i + 1 < g.n[0] )
        {
            const double2 rv = *reinterpret_cast<const double2*>( r + o );
            double2 pv = *reinterpret_cast<double2*>( p + o );
            pv.x = fma( beta, pv.x, op.minv[c0] * rv.x );
            pv.y = fma( beta, pv.y, op.minv[c1] * rv.y );
            *reinterpret_cast<double2*>( p + o ) = pv;
        }
        else
        {
            p[o] = fma( beta, p[o], op.minv[c0] * r[o] );
        }
    }
}

inline int stream_grid( const cfb_ctx* c, long long pairs )
{
    long long b = ( pairs + NT - 1 ) / NT;
    long long cap = (long long)c->sm_count * 8; // 8 x 256 threads = full occupancy, one wave
    if ( cap > CFB_MAX_PARTIALS )
        cap = CFB_MAX_PARTIALS;
    return (int)( b < 1 ? 1 : ( b > cap ? cap : b ) );
}

} // namespace

// Multi-GPU: all-gather the local double-double sums and combine them exactly.
int cg_global_sum( cfb_ctx* c, int which )
{
    const int nd = which == 0 ? 2 : ( which == 1 ? 4 : 6 ); // doubles per rank
    int rc = halo_allgather( c, &c->d_state->loc[which == 1 ? 2 : 0], c->d_state->gath, nd );
    if ( rc )
        return rc;
    cg_combine_kernel<<<1, 1, 0, c->stream>>>( c->d_state, which );
    c->stats.kernel_launches += 1;
    return CFB_OK;
}

int launch_divergence( cfb_ctx* c )
{
    const Geo& g = c->g;
    long long total = (long long)g.n[0] * g.n[1] * g.n[2];
    long long b = ( total + NT - 1 ) / NT;
    long long cap = (long long)c->sm_count * 16;
    int grid = (int)( b > cap ? cap : b );
    const double* u = field_ptr( c, CFB_U, CFB_CURRENT );
    const double* v = field_ptr( c, CFB_V, CFB_CURRENT );
    const double* w = field_ptr( c, CFB_W, CFB_CURRENT );
    if ( g.D == 2 )
        divergence_kernel<2><<<grid, NT, 0, c->stream>>>( g, u, v, w, c->rhs, c->lhs );
    else
        divergence_kernel<3><<<grid, NT, 0, c->stream>>>( g, u, v, w, c->rhs, c->lhs );
    return 1;
}

int launch_cg_init( cfb_ctx* c, int fixed )
{
    const Geo& g = c->g;
    long long pairs = (long long)( ( g.n[0] + 1 ) / 2 ) * g.n[1] * g.n[2];
    int grid = stream_grid( c, pairs );
    cg_init_kernel<<<grid, NT, 0, c->stream>>>( g, c->op, c->rhs, c->lhs, c->cg_r, c->cg_p, c->d_state,
                                                c->d_partials, fixed );
    int launches = 1;
    if ( cg_peer_mode( c ) )
        note_rc( c, peer_exchange( c, 1, false, c->pcur, true ) ); // sums + the faces of p0 (x: staging AND ghost columns)
    else if ( c->cfg.use_nccl )
        note_rc( c, cg_global_sum( c, 1 ) );
    cg_check0_kernel<<<1, 1, 0, c->stream>>>( c->d_state, c->cfg.cg_tolerance,
                                              c->cfg.cg_stop_rule == CFB_STOP_REL );
    return launches + 1;
}

int launch_cg_axpy( cfb_ctx* c )
{
    const Geo& g = c->g;
    long long pairs = (long long)( ( g.n[0] + 1 ) / 2 ) * g.n[1] * g.n[2];
    int grid = stream_grid( c, pairs );
    cg_axpy_kernel<<<grid, NT, 0, c->stream>>>( g, c->op, c->cg_p, c->cg_q, c->lhs, c->cg_r, c->d_state,
                                                c->d_partials );
    if ( c->cfg.use_nccl )
        note_rc( c, cg_global_sum( c, 1 ) );
    return 1;
}

int launch_cg_pupdate( cfb_ctx* c )
{
    const Geo& g = c->g;
    long long pairs = (long long)( ( g.n[0] + 1 ) / 2 ) * g.n[1] * g.n[2];
    int grid = stream_grid( c, pairs );
    cg_pupdate_kernel<<<grid, NT, 0, c->stream>>>( g, c->op, c->cg_r, c->cg_p, c->d_state );
    return 1;
}
