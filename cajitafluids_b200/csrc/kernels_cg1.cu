// kernels_cg1.cu — opt-in single-reduction form of the Jacobi-PCG (cg_variant 3).
//
// Chronopoulos and Gear's rearrangement of the conjugate-gradient recurrences: the same iterates as
// Cajita::ReferenceConjugateGradient::solve (driven from src/VelocityCorrector.hpp:276) in exact arithmetic, but
// ONE point per iteration where global sums are needed instead of the reference's three (two in the default
// two-kernel form), and one ghost exchange (the faces of r):
//
//   update   p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s       u = M^-1 r formed on the fly
//            (72 B/cell: read r, w, p, s, x; write p, s, x, r)                      [this file]
//   stencil  u = M^-1 r ; w = A u ; sum r.r ; gamma = sum r.u ; delta = sum w.u      (16 B/cell: read r, write w)
//            [kernels_stencil.cu MODE 2], then by one thread
//            beta = gamma / gamma_old ; alpha = gamma / ( delta - beta gamma / alpha_old )   [device_cg1.cuh]
//
// 88 B/cell against 72: slower where bandwidth rules, meant for small blocks per GPU where the reduction latency
// does.  Not bit-identical to the other forms (different recurrences): iteration counts within +-1, pressure to
// rounding; bit-identical to the checker's statement of the same algorithm.
// Arrays: p = cg_pbuf[0], s = cg_pbuf[1] (the search direction needs no double buffer here), w = cg_q.
#include "cfb_internal.h"
#include "device_geo.cuh"

#include <cmath>

namespace
{

constexpr int NT = 256;

__global__ void __launch_bounds__( NT )
    cg1_update_kernel( const __grid_constant__ Geo g, const __grid_constant__ OpConst op, const double* __restrict__ w,
                       double* __restrict__ p, double* __restrict__ s, double* __restrict__ x, double* __restrict__ r,
                       const CgState* S )
{
    if ( S->done )
        return;
    const double alpha = S->alpha, nalpha = -alpha, beta = S->beta;
    const bool first = S->iter == 0; // p and s hold nothing yet
    const unsigned npx = (unsigned)( ( g.n[0] + 1 ) >> 1 );
    const unsigned total = npx * (unsigned)g.n[1] * (unsigned)g.n[2];
    for ( unsigned t = blockIdx.x * NT + threadIdx.x; t < total; t += gridDim.x * NT )
    {
        const unsigned row = t / npx;
        const int i = 2 * (int)( t - row * npx );
        const int k = (int)( row / (unsigned)g.n[1] );
        const int j = (int)( row - (unsigned)k * (unsigned)g.n[1] );
        const int cyz = wall_count( g, 1, j + g.off[1] ) + wall_count( g, 2, k + g.off[2] );
        const int c0 = cyz + wall_count( g, 0, i + g.off[0] ), c1 = cyz + wall_count( g, 0, i + 1 + g.off[0] );
        const long long o = geo_off( g, i, j, k );
        if ( i + 1 < g.n[0] )
        {
            double2 rv = *reinterpret_cast<double2*>( r + o );
            const double2 wv = *reinterpret_cast<const double2*>( w + o );
            double2 xv = *reinterpret_cast<double2*>( x + o );
            const double u0 = op.minv[c0] * rv.x, u1 = op.minv[c1] * rv.y;
            double2 pv, sv;
            if ( first )
            {
                pv = make_double2( u0, u1 );
                sv = wv;
            }
            else
            {
                pv = *reinterpret_cast<double2*>( p + o );
                sv = *reinterpret_cast<double2*>( s + o );
                pv.x = fma( beta, pv.x, u0 );
                pv.y = fma( beta, pv.y, u1 );
                sv.x = fma( beta, sv.x, wv.x );
                sv.y = fma( beta, sv.y, wv.y );
            }
            xv.x = fma( alpha, pv.x, xv.x );
            xv.y = fma( alpha, pv.y, xv.y );
            rv.x = fma( nalpha, sv.x, rv.x );
            rv.y = fma( nalpha, sv.y, rv.y );
            *reinterpret_cast<double2*>( p + o ) = pv;
            *reinterpret_cast<double2*>( s + o ) = sv;
            *reinterpret_cast<double2*>( x + o ) = xv;
            *reinterpret_cast<double2*>( r + o ) = rv;
        }
        else
        {
            const double u0 = op.minv[c0] * r[o];
            const double pv = first ? u0 : fma( beta, p[o], u0 );
            const double sv = first ? w[o] : fma( beta, s[o], w[o] );
            p[o] = pv;
            s[o] = sv;
            x[o] = fma( alpha, pv, x[o] );
            r[o] = fma( nalpha, sv, r[o] );
        }
    }
}

} // namespace

int launch_cg1_update( cfb_ctx* c )
{
    const Geo& g = c->g;
    const long long pairs = (long long)( ( g.n[0] + 1 ) / 2 ) * g.n[1] * g.n[2];
    long long b = ( pairs + NT - 1 ) / NT;
    const long long cap = (long long)c->sm_count * 8;
    const int grid = (int)( b < 1 ? 1 : ( b > cap ? cap : b ) );
    cg1_update_kernel<<<grid, NT, 0, c->stream>>>( g, c->op, c->cg_q, c->cg_pbuf[0], c->cg_pbuf[1], c->lhs, c->cg_r,
                                                   c->d_state );
    return 1;
}

// the ghosts of r the stencil needs (several blocks) and, over NCCL, the global sums behind it
static int cg1_stencil_step( cfb_ctx* c, int init )
{
    int n = 0;
    const bool peer = c->cfg.use_nccl && c->peer_ok && c->use_peer && !( c->g.D == 2 && c->flat_2d );
    if ( peer )
    {
        // faces of r straight into the neighbours' ghost layers (side stream, joined at once: the update kernel
        // touches every cell, so there is nothing to run under the transfer), sums through the mailboxes in the
        // stencil kernel's last block
        CFB_CUDA( c, cudaEventRecord( c->ev_phase[0], c->stream ) );
        note_rc( c, peer_faces_async( c, 0, -1, c->ev_phase[0] ) );
        note_rc( c, peer_faces_join( c ) );
        n += launch_cg1_stencil( c, init, true );
    }
    else if ( c->cfg.use_nccl )
    {
        note_rc( c, halo_exchange_cells( c, c->cg_r, 1 ) );
        n += launch_cg1_stencil( c, init, false );
        note_rc( c, cg_global_sum( c, 2 + init ) );
    }
    else
        n += launch_cg1_stencil( c, init, false );
    return n;
}

// Jacobi-PCG from x0 = 0 in the single-reduction form.  Same polling scheme as pcg_solve (cfb_api.cu).
int cg1_pcg_solve( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid )
{
    const int fixed = fixed_iters > 0;
    const int max_it = fixed ? fixed_iters : c->cfg.cg_max_iter;
    const size_t head = offsetof( CgState, hist );
    long long launches = 0;
    c->sticky_rc = 0;
    cg_select_p( c, 0 );
    // x = 0, r = b, sum r^2 -> threshold / already converged (the search direction cg_init leaves in p is not used)
    const int variant = c->cg_variant;
    c->cg_variant = 0; // cg_init: sums over NCCL or through the exchange kernel, no face traffic of p0 wanted here
    launches += launch_cg_init( c, fixed );
    c->cg_variant = variant;
    launches += cg1_stencil_step( c, 1 );
    int enq = 0;
    bool done = false, pending = false;
    int batch = c->poll_every > 0 ? c->poll_every : 8;
    while ( enq < max_it && !done && !c->sticky_rc )
    {
        const int b = std::min( batch, max_it - enq );
        for ( int i = 0; i < b && !c->sticky_rc; ++i )
        {
            // "time_kernels": ms_k_axpy = the update kernel, ms_k_stencil = ghost exchange + stencil + reduction
            cudaEvent_t* e = ( c->time_kernels && c->ktimed < CFB_KTIMED ) ? c->kev[c->ktimed++] : nullptr;
            if ( e )
                cudaEventRecord( e[0], c->stream );
            launches += launch_cg1_update( c );
            if ( e )
            {
                cudaEventRecord( e[4], c->stream );
                cudaEventRecord( e[1], c->stream );
                cudaEventRecord( e[2], c->stream );
            }
            launches += cg1_stencil_step( c, 0 );
            if ( e )
            {
                cudaEventRecord( e[5], c->stream );
                cudaEventRecord( e[3], c->stream );
            }
        }
        enq += b;
        if ( fixed )
            continue;
        if ( pending )
        {
            CFB_CUDA( c, cudaEventSynchronize( c->ev[14] ) );
            done = c->h_state->done != 0;
            pending = false;
        }
        if ( !done )
        {
            CFB_CUDA( c, cudaMemcpyAsync( c->h_state, c->d_state, head, cudaMemcpyDeviceToHost, c->stream ) );
            CFB_CUDA( c, cudaEventRecord( c->ev[14], c->stream ) );
            pending = true;
        }
    }
    note_rc( c, peer_faces_join( c ) );
    CFB_CUDA( c, cudaMemcpyAsync( c->h_state, c->d_state, head, cudaMemcpyDeviceToHost, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    if ( c->sticky_rc )
    {
        const int rc = c->sticky_rc;
        c->sticky_rc = 0;
        return rc;
    }
    cudaError_t e = cudaGetLastError();
    if ( e != cudaSuccess )
        return cfb_fail( c, CFB_ERR_CUDA, std::string( "cg1_pcg_solve: " ) + cudaGetErrorString( e ) );
    c->stats.kernel_launches += launches;
    for ( int i = 0; i < c->ktimed; ++i )
    {
        float a = 0, d = 0;
        cudaEventElapsedTime( &a, c->kev[i][0], c->kev[i][1] );
        cudaEventElapsedTime( &d, c->kev[i][2], c->kev[i][3] );
        c->stats.ms_k_axpy += a;
        c->stats.ms_k_stencil += d;
        c->stats.k_timed_iters++;
    }
    c->ktimed = 0;
    if ( c->h_state->xerror )
        return cfb_fail( c, CFB_ERR_NCCL, "peer-memory exchange timed out: a rank never published its CG sums" );
    c->last_iters = c->h_state->iter;
    c->last_resid = std::sqrt( c->h_state->rr );
    c->stats.cg_iterations += c->last_iters;
    if ( num_iter )
        *num_iter = c->last_iters;
    if ( resid )
        *resid = c->last_resid;
    if ( c->cfg.cg_print_level > 0 && c->cfg.world_rank == 0 )
        std::printf( "Cajita CG Finished in %d iterations, |r|_2 = %g\n", c->last_iters, c->last_resid );
    if ( !fixed && !c->h_state->done )
        return cfb_fail( c, CFB_ERR_NOT_CONVERGED, "Cajita CG solver did not converge" );
    return CFB_OK;
}
