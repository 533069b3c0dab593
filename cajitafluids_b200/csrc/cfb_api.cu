// cfb_api.cu — host side of the C ABI (include/cfb.h): context, geometry, transfers and the
// orchestration that replaces Solver / ProblemManager / VelocityCorrector / TimeIntegrator::step.
// There is no CPU fallback: without an sm_100 device cfb_create fails with CFB_ERR_NO_DEVICE.
#include "cfb_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

std::string g_cfb_error;

int cfb_fail( cfb_ctx* c, int code, const std::string& msg )
{
    if ( c )
        c->err = msg;
    g_cfb_error = msg;
    return code;
}

int ensure_partials( cfb_ctx* c, long long units )
{
    if ( units > CFB_MAX_UNITS )
        return cfb_fail( c, CFB_ERR_INVALID,
                         "tiling yields " + std::to_string( units ) + " units in one launch (limit " +
                             std::to_string( (long long)CFB_MAX_UNITS ) + "): choose larger tiles" );
    const int need = (int)std::max<long long>( units, CFB_MAX_PARTIALS );
    if ( c->d_partials && need <= c->partials_cap )
        return CFB_OK;
    if ( c->stream )
        CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    if ( c->d_partials )
        cudaFree( c->d_partials );
    c->d_partials = nullptr;
    c->partials_cap = 0;
    CFB_CUDA( c, cudaMalloc( &c->d_partials, (size_t)3 * 2 * need * sizeof( double ) ) ); // <= 3 values per block
    c->partials_cap = need;
    return CFB_OK;
}

namespace
{

// Cajita GlobalGrid partition: n/nb cells per block, the first n%nb blocks get one more.
void partition( int n, int nb, int b, int& owned, int& offset )
{
    int base = n / nb, rem = n % nb;
    owned = base + ( b < rem ? 1 : 0 );
    offset = b * base + std::min( b, rem );
}

inline int field_entity( int field ) { return ( field >= CFB_U && field <= CFB_W ) ? field : 0; }

void owned_extent( const Geo& g, int field, int ext[3] )
{
    int ent = field_entity( field );
    for ( int d = 0; d < 3; ++d )
        ext[d] = ( d < g.D ) ? ( ent - 1 == d ? g.nf[d] : g.n[d] ) : 1;
}

// Per-phase device timers.  A phase records start/stop events on the stream; the elapsed time is
// collected lazily (next use of the slot or cfb_get_stats), so timing never serialises host and
// device inside a step.
enum { PH_ADVECT = 0, PH_INPUTS, PH_RHS, PH_PCG, PH_APPLY, PH_COUNT };
enum { EV_BENCH0 = 12, EV_BENCH1 = 13, EV_POLL = 14 };

double* phase_acc( cfb_ctx* c, int slot )
{
    switch ( slot )
    {
    case PH_ADVECT:
        return &c->stats.ms_advect;
    case PH_INPUTS:
        return &c->stats.ms_add_inputs;
    case PH_RHS:
        return &c->stats.ms_build_rhs;
    case PH_PCG:
        return &c->stats.ms_pcg;
    default:
        return &c->stats.ms_apply_pressure;
    }
}
void timer_collect( cfb_ctx* c, int slot )
{
    if ( !c->ev_pending[slot] )
        return;
    cudaEventSynchronize( c->ev[2 * slot + 1] );
    float ms = 0;
    cudaEventElapsedTime( &ms, c->ev[2 * slot], c->ev[2 * slot + 1] );
    *phase_acc( c, slot ) += ms;
    c->ev_pending[slot] = false;
}
void timer_start( cfb_ctx* c, int slot )
{
    timer_collect( c, slot );
    cudaEventRecord( c->ev[2 * slot], c->stream );
}
void timer_stop( cfb_ctx* c, int slot )
{
    cudaEventRecord( c->ev[2 * slot + 1], c->stream );
    c->ev_pending[slot] = true;
}

int check_cuda( cfb_ctx* c, cudaError_t e )
{
    return e == cudaSuccess ? CFB_OK : cfb_fail( c, CFB_ERR_CUDA, cudaGetErrorString( e ) );
}

int check_async( cfb_ctx* c, const char* what )
{
    cudaError_t e = cudaGetLastError();
    if ( e != cudaSuccess )
        return cfb_fail( c, CFB_ERR_CUDA, std::string( what ) + ": " + cudaGetErrorString( e ) );
    return CFB_OK;
}

int copy3d( cfb_ctx* c, int field, int version, int region, double* host, bool to_device )
{
    double* base = field_ptr( c, field, version );
    if ( !base )
        return cfb_fail( c, CFB_ERR_INVALID, "invalid field id" );
    const Geo& g = c->g;
    int ext[3];
    owned_extent( g, field, ext );
    int lo[3] = { 0, 0, 0 };
    if ( region == CFB_GHOSTED )
    {
        // Cajita's Ghost index space: owned cells + 2*halo, +1 along a face normal on EVERY block
        // (tests/tstMesh.cpp:61-68), i.e. local indices [0, n + 2h (+1)).
        const int ent = field_entity( field );
        for ( int d = 0; d < g.D; ++d )
        {
            ext[d] = g.n[d] + 2 * g.h + ( ent - 1 == d ? 1 : 0 );
            lo[d] = -g.h;
        }
    }
    double* dev = base + g.origin + (long long)lo[2] * g.sz + (long long)lo[1] * g.sy + lo[0];
    cudaMemcpy3DParms p{};
    cudaPitchedPtr dptr = make_cudaPitchedPtr( dev, (size_t)g.sy * 8, (size_t)g.sy, (size_t)g.ay );
    cudaPitchedPtr hptr = make_cudaPitchedPtr( host, (size_t)ext[0] * 8, (size_t)ext[0], (size_t)ext[1] );
    p.srcPtr = to_device ? hptr : dptr;
    p.dstPtr = to_device ? dptr : hptr;
    p.extent = make_cudaExtent( (size_t)ext[0] * 8, (size_t)ext[1], (size_t)ext[2] );
    p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    CFB_CUDA( c, cudaMemcpy3DAsync( &p, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    return CFB_OK;
}

// Number of CG iterations enqueued between two looks at the device state.
int poll_batch( const cfb_ctx* c )
{
    if ( c->poll_every > 0 )
        return c->poll_every;
    const Geo& g = c->g;
    double cells = (double)g.n[0] * g.n[1] * g.n[2];
    double t_iter_us = std::max( cells * 88.0 / 6.5e12 * 1e6, 9.0 ); // bandwidth vs launch bound
    int b = (int)std::ceil( 200.0 / t_iter_us );
    return std::min( 32, std::max( 2, b ) );
}

// One CG iteration = 3 kernels.  With the "time_kernels" tuning switch on, the first
// CFB_KTIMED iterations of a solve are bracketed kernel by kernel with CUDA events on the
// launching stream (this is where bench.py's roofline numbers come from).
int enqueue_iteration( cfb_ctx* c )
{
    int n = 0;
    cudaEvent_t* e = nullptr;
    if ( c->time_kernels && c->ktimed < CFB_KTIMED )
        e = c->kev[c->ktimed++];
    if ( c->cg_variant >= 1 )
    {
        // two-kernel forms (kernels_fused.cu): e[0..1] phase A, e[2..3] phase B.  Variant 2 (64 B/cell) does
        // not store q: its phase A' recomputes A p from the search direction, whose ghosts the peer exchange
        // after phase B has delivered (NCCL path: one more exchange, phase B recomputes its ring itself)
        const bool peer = cg_peer_mode( c );
        if ( e )
            cudaEventRecord( e[0], c->stream );
        if ( peer_overlapped( c ) )
        {
            // Overlapped exchange.
            //   main stream: phase A (its last block runs the (r.z, r.r) mailboxes) -> interior units of phase B
            //                -> [boundary units done] -> next phase A ...
            //   side stream: [phase A done] -> faces of r to the neighbours, wait for theirs -> boundary units of phase B
            //                -> faces of the new search direction, wait for theirs (under the interior units and the
            //                next phase A) -> ...
            // The side stream has the higher priority, so the boundary units are scheduled ahead of the interior
            // units still waiting; the two launches of phase B run concurrently and share one ticket counter: the
            // block that draws the last ticket, whichever launch it belongs to, runs the p.Ap mailboxes.
            // Nobody overwrites a ghost layer that may still be read.  A rank stores its r faces after its own
            // phase A, which it enters only behind the p.Ap mailboxes of the previous iteration, i.e. after every
            // rank's boundary units — the readers of the old r ghosts — are done.  It stores the faces of the new
            // search direction after its boundary units, which ran behind the r-face handshake with its
            // neighbours, who raised their flag after their own phase A, i.e. after their boundary units of the
            // previous iteration — the last readers of that buffer's ghosts (the direction is double-buffered).
            // The x staging slots (one for r, one per direction buffer) follow the same argument: the receiver
            // scatters a slot on its side stream before its boundary units.
            n += launch_cg_rupdate_mail( c );
            note_rc( c, check_cuda( c, cudaEventRecord( c->ev_phase[0], c->stream ) ) );
            if ( e )
            {
                cudaEventRecord( e[4], c->stream );
                cudaEventRecord( e[1], c->stream );
                cudaEventRecord( e[2], c->stream );
            }
            note_rc( c, peer_faces_async( c, 0, -1, c->ev_phase[0] ) );
            n += launch_cg_fused_mail( c, 2, true );
            note_rc( c, check_cuda( c, cudaEventRecord( c->ev_bnd, c->comm_stream ) ) );
            note_rc( c, peer_faces_async( c, 1, c->pcur ^ 1, nullptr ) );
            n += launch_cg_fused_mail( c, 1, false );
            note_rc( c, check_cuda( c, cudaStreamWaitEvent( c->stream, c->ev_bnd, 0 ) ) );
            cg_select_p( c, c->pcur ^ 1 );
            if ( e )
            {
                cudaEventRecord( e[5], c->stream );
                cudaEventRecord( e[3], c->stream );
            }
            return n;
        }
        if ( c->cg_variant == 2 )
        {
            if ( c->cfg.use_nccl && !peer )
                note_rc( c, halo_exchange_cells( c, c->cg_p, 1 ) );
            n += launch_stencil_rupdate( c );
        }
        else
            n += launch_cg_rupdate( c );
        if ( e )
            cudaEventRecord( e[4], c->stream );
        if ( peer )
            note_rc( c, peer_exchange( c, 1, true, -1, !peer_xstaged( c ) ) ); // r faces -> neighbours, (rz_new, rr) -> all
        else if ( c->cfg.use_nccl )
            note_rc( c, cg_global_sum( c, 1 ) );
        if ( e )
        {
            cudaEventRecord( e[1], c->stream );
            cudaEventRecord( e[2], c->stream );
        }
        if ( peer )
        {
            n += launch_cg_fused( c, 0 );
            if ( e )
                cudaEventRecord( e[5], c->stream );
            note_rc( c, peer_exchange( c, 0, false, c->pcur ^ 1, !peer_xstaged( c ) ) ); // new p faces, pAp -> all
        }
        else if ( c->cfg.use_nccl )
        {
            double* fl[2] = { c->cg_r, c->cg_p };
            if ( c->overlap_halo && c->n_interior > 0 )
            {
                note_rc( c, halo_cells_begin( c, fl, 2, 1 ) );
                n += launch_cg_fused( c, 1 );
                note_rc( c, halo_cells_end( c ) );
                n += launch_cg_fused( c, 2 );
            }
            else
            {
                note_rc( c, halo_cells_begin( c, fl, 2, 1 ) );
                note_rc( c, halo_cells_end( c ) );
                n += launch_cg_fused( c, 0 );
            }
            if ( e )
                cudaEventRecord( e[5], c->stream );
            note_rc( c, cg_global_sum( c, 0 ) );
        }
        else
            n += launch_cg_fused( c, 0 );
        cg_select_p( c, c->pcur ^ 1 ); // phase B wrote the new p into the other buffer
        if ( e )
        {
            if ( !( peer || c->cfg.use_nccl ) )
                cudaEventRecord( e[5], c->stream );
            cudaEventRecord( e[3], c->stream );
        }
        return n;
    }
    if ( e )
        cudaEventRecord( e[0], c->stream );
    n += launch_cg_axpy( c );
    if ( e )
    {
        cudaEventRecord( e[4], c->stream );
        cudaEventRecord( e[1], c->stream );
    }
    n += launch_cg_pupdate( c );
    if ( e )
        cudaEventRecord( e[2], c->stream );
    if ( c->cfg.use_nccl )
        note_rc( c, halo_exchange_cells( c, c->cg_p, 1 ) );
    n += launch_stencil_dot( c );
    if ( e )
        cudaEventRecord( e[5], c->stream );
    if ( c->cfg.use_nccl )
        note_rc( c, cg_global_sum( c, 0 ) );
    if ( e )
        cudaEventRecord( e[3], c->stream );
    return n;
}

void collect_kernel_times( cfb_ctx* c )
{
    for ( int i = 0; i < c->ktimed; ++i )
    {
        float a = 0, b = 0, d = 0;
        cudaEventElapsedTime( &a, c->kev[i][0], c->kev[i][1] );
        cudaEventElapsedTime( &b, c->kev[i][1], c->kev[i][2] );
        cudaEventElapsedTime( &d, c->kev[i][2], c->kev[i][3] );
        c->stats.ms_k_axpy += a;
        c->stats.ms_k_pupdate += b;
        c->stats.ms_k_stencil += d;
        float xa = 0, xb = 0;
        if ( cudaEventElapsedTime( &xa, c->kev[i][4], c->kev[i][1] ) == cudaSuccess )
            c->stats.ms_k_exch_a += xa;
        if ( cudaEventElapsedTime( &xb, c->kev[i][5], c->kev[i][3] ) == cudaSuccess )
            c->stats.ms_k_exch_b += xb;
        cudaGetLastError();
        c->stats.k_timed_iters++;
    }
    c->ktimed = 0;
}

constexpr size_t STATE_HEAD = offsetof( CgState, hist );

// Jacobi-PCG from x0 = 0 on the current RHS.  `fixed_iters` > 0: exactly that many iterations.
int pcg_solve_form( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid );

// Jacobi-PCG from x0 = 0 on the current RHS in the CG form in force: the one chosen with "cg_variant", or the
// automatic choice for this block and these options (cg_variant_auto), fixed for the duration of the solve.
int pcg_solve( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid )
{
    if ( c->precond == CFB_PRECOND_MG )
        return mg_pcg_solve( c, fixed_iters, num_iter, resid );
    const int chosen = c->cg_variant;
    c->cg_variant = cg_variant_auto( c );
    const int rc = pcg_solve_form( c, fixed_iters, num_iter, resid );
    c->cg_variant = chosen;
    return rc;
}

int pcg_solve_form( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid )
{
    if ( c->cg_variant == 3 )
        return cg1_pcg_solve( c, fixed_iters, num_iter, resid ); // opt-in single-reduction form (kernels_cg1.cu)
    const int fixed = fixed_iters > 0;
    const int max_it = fixed ? fixed_iters : c->cfg.cg_max_iter;
    long long launches = 0;
    c->sticky_rc = 0;
    const int p_start = c->pcur;
    const bool peer = cg_peer_mode( c );
    launches += launch_cg_init( c, fixed ); // peer mode: includes the exchange of p0's faces
    if ( c->cfg.use_nccl && !peer )
        note_rc( c, halo_exchange_cells( c, c->cg_p, 1 ) );
    launches += launch_stencil_dot( c );
    if ( peer )
        note_rc( c, peer_exchange( c, 0, false, -1, false ) );
    else if ( c->cfg.use_nccl )
        note_rc( c, cg_global_sum( c, 0 ) );

    // a launcher refused its configuration or an exchange call failed: nothing (more) is enqueued, the
    // error goes back to the caller instead of a wrong x with CFB_OK
    auto bail = [&]() -> int {
        const int rc = take_sticky_rc( c );
        peer_faces_join( c );
        cudaStreamSynchronize( c->stream );
        c->ktimed = 0;
        return rc;
    };
    if ( c->sticky_rc )
        return bail();
    const bool persist = cg_persist_applies( c ); // small block: batches of iterations in one cooperative launch
    const int batch = persist ? std::max( poll_batch( c ), 32 ) : poll_batch( c );
    int enq = 0;
    bool done = false;
    // Pipelined polling: the state of batch i is inspected while batch i+1 is already queued, so
    // the GPU never waits for the host.  Kernels launched after convergence return immediately
    // (CgState::done), which keeps x and the iteration count exact.
    int pending = 0; // number of state snapshots in flight (0..1)
    while ( enq < max_it && !done )
    {
        int b = std::min( batch, max_it - enq );
        if ( persist )
        {
            launches += launch_cg_persistent( c, b );
            cg_select_p( c, c->pcur ^ ( b & 1 ) ); // every iteration writes the new direction into the other buffer
        }
        else
            for ( int i = 0; i < b && !c->sticky_rc; ++i )
                launches += enqueue_iteration( c );
        if ( c->sticky_rc )
            return bail();
        enq += b;
        if ( fixed )
            continue;
        if ( pending )
        {
            CFB_CUDA( c, cudaEventSynchronize( c->ev[EV_POLL] ) );
            const CgState* hs = c->h_state;
            // two-kernel form: `done` is recorded one phase after the tolerance was met
            done = hs->done != 0 || ( hs->iter > 0 && std::sqrt( hs->rr ) <= hs->thresh );
            pending = 0;
        }
        if ( !done )
        {
            CFB_CUDA( c, cudaMemcpyAsync( c->h_state, c->d_state, STATE_HEAD, cudaMemcpyDeviceToHost, c->stream ) );
            CFB_CUDA( c, cudaEventRecord( c->ev[EV_POLL], c->stream ) );
            pending = 1;
        }
    }
    note_rc( c, peer_faces_join( c ) ); // overlapped exchange: the side stream's last transfers
    if ( c->cg_variant >= 1 )
        launches += launch_cg_finish( c );
    CFB_CUDA( c, cudaMemcpyAsync( c->h_state, c->d_state, STATE_HEAD, cudaMemcpyDeviceToHost, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    int rc = check_async( c, "pcg_solve" );
    if ( rc )
        return rc;
    collect_kernel_times( c );
    c->stats.kernel_launches += launches;
    if ( c->h_state->xerror )
        return cfb_fail( c, CFB_ERR_NCCL, "peer-memory exchange timed out: a rank never published its CG sums" );
    c->last_iters = c->h_state->iter;
    c->last_resid = std::sqrt( c->h_state->rr );
    if ( c->cg_variant >= 1 )
    {
        // launches enqueued after convergence were no-ops on the device but flipped the host's idea
        // of the current p buffer: every executed phase B except a converging one wrote a new p
        const int swaps = c->last_iters - ( ( c->h_state->done && c->last_iters > 0 ) ? 1 : 0 );
        cg_select_p( c, p_start + swaps );
    }
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

} // namespace

extern "C" {

int cfb_abi_version( void ) { return CFB_ABI_VERSION; }

const char* cfb_last_error( const cfb_ctx* c ) { return c ? c->err.c_str() : g_cfb_error.c_str(); }

int cfb_default_config( cfb_config* cfg, int dim )
{
    if ( !cfg || ( dim != 2 && dim != 3 ) )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "dim must be 2 or 3" );
    std::memset( cfg, 0, sizeof( *cfg ) );
    cfg->struct_size = (int32_t)sizeof( cfb_config );
    cfg->dim = dim;
    // examples/advection.cpp:174-184
    const double loc[3] = { 0.2, 0.45, 0.45 }, size[3] = { 0.02, 0.1, 0.1 }, vel[3] = { 1.0, 0.0, 0.0 };
    for ( int d = 0; d < 3; ++d )
    {
        cfg->global_num_cell[d] = d < dim ? 128 : 1;
        cfg->global_bounding_box[d] = 0.0;
        cfg->global_bounding_box[3 + d] = d < dim ? 1.0 : 0.0;
        cfg->ranks_per_dim[d] = 1;
        cfg->inflow_location[d] = d < dim ? loc[d] : 0.0;
        cfg->inflow_size[d] = d < dim ? size[d] : 0.0;
        cfg->inflow_velocity[d] = d < dim ? vel[d] : 0.0;
    }
    cfg->halo_cell_width = 3; // src/Solver.hpp:78
    cfg->world_size = 1;
    cfg->density = 0.1;
    cfg->delta_t = 0.005;
    cfg->clamp_dt = 1;
    for ( int i = 0; i < 6; ++i )
        cfg->boundary_type[i] = CFB_SOLID; // examples/advection.cpp:446-448
    cfg->inflow_quantity = 3.0;
    cfg->cg_tolerance = 1.0e-6; // src/VelocityCorrector.hpp:103-105
    cfg->cg_max_iter = 2000;
    cfg->cg_print_level = 0;
    cfg->cg_stop_rule = CFB_STOP_ABS;
    cfg->field_interp_order = 3; // src/TimeIntegrator.hpp:113
    cfg->quirk_applypressure_bc = dim == 2 ? 1 : 0;
    cfg->quirk_rk3_stage3_v0 = 1;
    return CFB_OK;
}

int cfb_partition( int n, int nb, int block, int* owned, int* offset )
{
    if ( nb < 1 || block < 0 || block >= nb )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "bad block" );
    int o, f;
    partition( n, nb, block, o, f );
    if ( owned )
        *owned = o;
    if ( offset )
        *offset = f;
    return CFB_OK;
}

int cfb_create( const cfb_config* cfg, cfb_ctx** out )
{
    if ( out )
        *out = nullptr;
    if ( !cfg || !out || cfg->struct_size != (int32_t)sizeof( cfb_config ) )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "cfb_config size mismatch (ABI)" );
    if ( cfg->dim != 2 && cfg->dim != 3 )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "dim must be 2 or 3" );
    if ( cfg->halo_cell_width < 1 || cfg->halo_cell_width > 8 )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "halo_cell_width out of range" );
    if ( cfg->field_interp_order != 1 && cfg->field_interp_order != 3 )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "field_interp_order must be 1 or 3" );
    const int D = cfg->dim;
    for ( int d = 0; d < D; ++d )
        if ( cfg->global_num_cell[d] < 1 || cfg->ranks_per_dim[d] < 1 || cfg->block_id[d] < 0 ||
             cfg->block_id[d] >= cfg->ranks_per_dim[d] )
            return cfb_fail( nullptr, CFB_ERR_INVALID, "bad cell count / block grid" );

    // Mesh ctor: src/Mesh.hpp:50-64
    const double cell = ( cfg->global_bounding_box[3] - cfg->global_bounding_box[0] ) / cfg->global_num_cell[0];
    for ( int d = 0; d < D; ++d )
    {
        double extent = cfg->global_num_cell[d] * cell;
        if ( std::abs( extent - ( cfg->global_bounding_box[3 + d] - cfg->global_bounding_box[d] ) ) >
             10.0 * std::numeric_limits<double>::epsilon() )
            return cfb_fail( nullptr, CFB_ERR_MESH_EXTENT, "Extent not evenly divisible by uniform cell size" );
    }

    int ndev = 0;
    if ( cudaGetDeviceCount( &ndev ) != cudaSuccess || ndev < 1 )
    {
        cudaGetLastError();
        return cfb_fail( nullptr, CFB_ERR_NO_DEVICE, "no CUDA device: cajitafluids_b200 has no CPU fallback" );
    }
    if ( cfg->device_id < 0 || cfg->device_id >= ndev )
        return cfb_fail( nullptr, CFB_ERR_INVALID, "device_id out of range" );
    cudaDeviceProp prop{};
    cudaGetDeviceProperties( &prop, cfg->device_id );
    if ( prop.major != 10 )
        return cfb_fail( nullptr, CFB_ERR_NO_DEVICE,
                         std::string( "device is sm_" ) + std::to_string( prop.major * 10 + prop.minor ) +
                             ", this library is built for sm_100a only" );

    cfb_ctx* c = new cfb_ctx();
    *out = c;
    c->cfg = *cfg;
    c->device = cfg->device_id;
    c->sm_count = prop.multiProcessorCount;
    CFB_CUDA( c, cudaSetDevice( c->device ) );

    c->flat_2d = D == 2;
    Geo& g = c->g;
    g.D = D;
    g.h = cfg->halo_cell_width;
    g.cell = cell;
    for ( int d = 0; d < 3; ++d )
    {
        if ( d < D )
        {
            partition( cfg->global_num_cell[d], cfg->ranks_per_dim[d], cfg->block_id[d], g.n[d], g.off[d] );
            g.gn[d] = cfg->global_num_cell[d];
            g.lo_bd[d] = cfg->block_id[d] == 0;
            g.hi_bd[d] = cfg->block_id[d] == cfg->ranks_per_dim[d] - 1;
            g.bt[d] = cfg->boundary_type[d];
            g.bt[3 + d] = cfg->boundary_type[D + d];
            // UniformGlobalMesh: per-dimension cell size; LocalMesh: own low corner = global low +
            // cell_d * offset; ghosted low = own low - halo * cell_d
            g.celld[d] = ( cfg->global_bounding_box[3 + d] - cfg->global_bounding_box[d] ) / cfg->global_num_cell[d];
            g.rdxd[d] = 1.0 / g.celld[d];
            double own_low = cfg->global_bounding_box[d] + g.celld[d] * g.off[d];
            g.ghost_low[d] = own_low - g.h * g.celld[d];
        }
        else
        {
            // 2-D: a single plane between two SOLID z walls (see Geo)
            g.n[d] = 1;
            g.off[d] = 0;
            g.gn[d] = 1;
            g.lo_bd[d] = g.hi_bd[d] = 1;
            g.bt[d] = g.bt[3 + d] = CFB_SOLID;
            g.ghost_low[d] = 0.0;
            g.celld[d] = cell;
            g.rdxd[d] = 1.0 / cell;
        }
        g.nf[d] = g.n[d] + ( ( d < D && g.hi_bd[d] ) ? 1 : 0 );
        if ( g.n[d] < 1 )
            return cfb_fail( c, CFB_ERR_INVALID, "a block owns no cells" );
    }
    const int HX = 16;
    g.sy = ( ( HX + g.n[0] + 1 + g.h ) + 15 ) / 16 * 16;
    g.ay = g.n[1] + 1 + 2 * g.h;
    g.az = g.n[2] + 1 + 2 * g.h;
    g.sz = g.sy * g.ay;
    g.origin = (long long)g.h * g.sz + (long long)g.h * g.sy + HX;
    g.total = g.sz * g.az;

    // Solver ctor dt clamp: src/Solver.hpp:96-106
    double dt = cfg->delta_t;
    if ( cfg->clamp_dt )
    {
        double f2 = 0, vmax = 0;
        for ( int d = 0; d < D; ++d )
        {
            f2 += cfg->body_force[d] * cfg->body_force[d];
            vmax = std::fmax( vmax, std::fabs( cfg->inflow_velocity[d] ) );
        }
        double umax = vmax + std::sqrt( std::sqrt( f2 ) * cell );
        if ( umax > 0 && dt > cell / umax )
        {
            dt = cell / umax;
            if ( cfg->world_rank == 0 && cfg->cg_print_level > 0 )
                std::fprintf( stderr, "Reducting timestep to %g given mesh size and inflow velocity.\n", dt );
        }
    }
    g.dt = dt;
    g.time = 0.0;

    // VelocityCorrector::initializeMatrixValues scale (src/VelocityCorrector.hpp:128) and the
    // possible diagonals: 2*D*scale, minus scale per SOLID wall (BoundaryConditions.hpp:56-97).
    // In 2-D the kernels see two extra SOLID z walls, so index = 2 + (x,y walls): 6s - s - s == 4s.
    OpConst& op = c->op;
    op.scale = dt / ( cfg->density * cell * cell );
    op.neg_scale = -1.0 * op.scale;
    for ( int cnt = 0; cnt < 8; ++cnt )
    {
        double dgl;
        if ( D == 3 )
        {
            dgl = 6.0 * op.scale;
            for ( int i = 0; i < cnt; ++i )
                dgl -= op.scale;
        }
        else
        {
            dgl = 4.0 * op.scale; // src/VelocityCorrector.hpp:137
            for ( int i = 0; i < cnt - 2; ++i )
                dgl -= op.scale;
        }
        op.diag[cnt] = dgl;
        op.minv[cnt] = 1.0 / dgl; // src/VelocityCorrector.hpp:178
    }

    InflowConst& s = c->inflow;
    for ( int d = 0; d < 3; ++d )
    {
        s.lo[d] = cfg->inflow_location[d];
        s.hi[d] = cfg->inflow_location[d] + cfg->inflow_size[d]; // src/InflowSource.hpp:84-87
        s.vel[d] = cfg->inflow_velocity[d];
        s.force_dt[d] = cfg->body_force[d] * dt; // src/BodyForce.hpp:50
    }
    s.quantity = cfg->inflow_quantity;

    CFB_CUDA( c, cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) );
    {
        // the side stream carries the exchanges that run under compute kernels: its blocks go first
        int lo_prio = 0, hi_prio = 0;
        CFB_CUDA( c, cudaDeviceGetStreamPriorityRange( &lo_prio, &hi_prio ) );
        CFB_CUDA( c, cudaStreamCreateWithPriority( &c->comm_stream, cudaStreamNonBlocking, hi_prio ) );
    }
    for ( auto& e : c->ev )
        CFB_CUDA( c, cudaEventCreate( &e ) );

    const size_t bytes = (size_t)g.total * sizeof( double );
    auto alloc0 = [&]( double** p ) -> cudaError_t {
        cudaError_t e = cudaMalloc( p, bytes );
        if ( e != cudaSuccess )
            return e;
        return cudaMemsetAsync( *p, 0, bytes, c->stream ); // assign( 0.0, Ghost() )  ProblemManager.hpp:149-165
    };
    for ( int f = 0; f <= D; ++f )
        for ( int v = 0; v < 2; ++v )
            CFB_CUDA( c, alloc0( &c->fld[f][v] ) );
    CFB_CUDA( c, alloc0( &c->lhs ) );
    CFB_CUDA( c, alloc0( &c->rhs ) );
    CFB_CUDA( c, alloc0( &c->cg_r ) );
    CFB_CUDA( c, alloc0( &c->cg_pbuf[0] ) );
    CFB_CUDA( c, alloc0( &c->cg_pbuf[1] ) );
    c->cg_p = c->cg_pbuf[0];
    CFB_CUDA( c, alloc0( &c->cg_q ) );
    CFB_CUDA( c, cudaMalloc( &c->d_state, sizeof( CgState ) ) );
    CFB_CUDA( c, cudaMemsetAsync( c->d_state, 0, sizeof( CgState ), c->stream ) );
    {
        const int one = 1;
        CFB_CUDA( c, cudaMemcpyAsync( &c->d_state->world, &one, sizeof( int ), cudaMemcpyHostToDevice, c->stream ) );
    }
    CFB_CUDA( c, cudaMallocHost( &c->h_state, sizeof( CgState ) ) );
    std::memset( c->h_state, 0, sizeof( CgState ) );
    {
        const int rc = ensure_partials( c, CFB_MAX_PARTIALS );
        if ( rc )
            return rc;
    }
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );

    // ProblemManager::initialize with the constant MeshInitFunc (examples/advection.cpp:382-435)
    for ( int f = 0; f <= D; ++f )
    {
        double val = f == 0 ? cfg->init_quantity : cfg->init_velocity[f - 1];
        if ( val == 0.0 )
            continue;
        int ext[3];
        owned_extent( g, f, ext );
        std::vector<double> hostv( (size_t)ext[0] * ext[1] * ext[2], val );
        int rc = copy3d( c, f, CFB_CURRENT, CFB_OWNED, hostv.data(), true );
        if ( rc )
            return rc;
    }

    // tile shape of the stencil kernel: 64 x 16 x 4 stages by default
    int rc = stencil_setup( c );
    if ( rc )
        return rc;
    if ( cfg->use_nccl && cfg->world_size > 1 )
    {
        rc = halo_init( c );
        if ( rc )
            return rc;
    }
    else
        c->cfg.use_nccl = 0;
    // two-kernel CG iteration: tensor maps of cg_r / cg_p and the interior/boundary unit list
    // (needs the neighbour ranks halo_init just set)
    return fused_setup( c );
}

int cfb_destroy( cfb_ctx* c )
{
    if ( !c )
        return CFB_OK;
    cudaSetDevice( c->device );
    if ( c->stream )
        cudaStreamSynchronize( c->stream );
    output_destroy( c ); // writes a pending output first
    mg_destroy( c );
    halo_destroy( c );
    for ( int f = 0; f < 4; ++f )
        for ( int v = 0; v < 2; ++v )
            if ( c->fld[f][v] )
                cudaFree( c->fld[f][v] );
    for ( double* p : { c->lhs, c->rhs, c->cg_r, c->cg_pbuf[0], c->cg_pbuf[1], c->cg_q, c->d_partials } )
        if ( p )
            cudaFree( p );
    if ( c->d_units )
        cudaFree( c->d_units );
    if ( c->d_state )
        cudaFree( c->d_state );
    if ( c->h_state )
        cudaFreeHost( c->h_state );
    for ( auto& e : c->ev )
        if ( e )
            cudaEventDestroy( e );
    for ( cudaEvent_t e : { c->ev_phase[0], c->ev_phase[1], c->ev_ghost, c->ev_bnd } )
        if ( e )
            cudaEventDestroy( e );
    for ( auto& row : c->kev )
        for ( auto& e : row )
            if ( e )
                cudaEventDestroy( e );
    if ( c->stream )
        cudaStreamDestroy( c->stream );
    if ( c->comm_stream )
        cudaStreamDestroy( c->comm_stream );
    delete c;
    return CFB_OK;
}

int cfb_get_scalars( const cfb_ctx* c, double* cell, double* dt, double* time )
{
    if ( cell )
        *cell = c->g.cell;
    if ( dt )
        *dt = c->g.dt;
    if ( time )
        *time = c->g.time;
    return CFB_OK;
}

int cfb_owned_extent( const cfb_ctx* c, int field, int ext[3] )
{
    owned_extent( c->g, field, ext );
    return CFB_OK;
}

int cfb_global_offset( const cfb_ctx* c, int off[3] )
{
    for ( int d = 0; d < 3; ++d )
        off[d] = c->g.off[d];
    return CFB_OK;
}

int cfb_field_ptr( cfb_ctx* c, int field, int version, double** dev_ptr, int64_t* origin, int64_t* stride_y,
                   int64_t* stride_z )
{
    double* p = field_ptr( c, field, version );
    if ( !p )
        return cfb_fail( c, CFB_ERR_INVALID, "invalid field id" );
    *dev_ptr = p;
    if ( origin )
        *origin = c->g.origin;
    if ( stride_y )
        *stride_y = c->g.sy;
    if ( stride_z )
        *stride_z = c->g.sz;
    return CFB_OK;
}

int cfb_upload( cfb_ctx* c, int field, int version, int region, const double* host )
{
    return copy3d( c, field, version, region, const_cast<double*>( host ), true );
}
int cfb_download( cfb_ctx* c, int field, int version, int region, double* host )
{
    return copy3d( c, field, version, region, host, false );
}

int cfb_advance( cfb_ctx* c, int field )
{
    if ( field < 0 || field > c->g.D )
        return cfb_fail( c, CFB_ERR_INVALID, "advance: invalid field" );
    c->cur[field] = 1 - c->cur[field];
    return CFB_OK;
}

int cfb_gather( cfb_ctx* c, int version )
{
    if ( c->cfg.use_nccl )
        return halo_exchange_fields( c, version );
    return CFB_OK; // single block: physical-wall ghosts stay zero
}

int cfb_add_inputs( cfb_ctx* c )
{
    timer_start( c, PH_INPUTS );
    c->stats.kernel_launches += launch_add_inputs( c );
    timer_stop( c, PH_INPUTS );
    return check_async( c, "add_inputs" );
}

int cfb_time_integrator_step( cfb_ctx* c )
{
    timer_start( c, PH_ADVECT );
    int rc = cfb_gather( c, CFB_CURRENT ); // src/TimeIntegrator.hpp:131
    if ( rc )
        return rc;
    c->stats.kernel_launches += launch_advect( c );
    for ( int f = 0; f <= c->g.D; ++f ) // pm.advance  src/TimeIntegrator.hpp:167-172
        c->cur[f] = 1 - c->cur[f];
    timer_stop( c, PH_ADVECT );
    return check_async( c, "time_integrator_step" );
}

int cfb_build_rhs( cfb_ctx* c )
{
    timer_start( c, PH_RHS );
    // The divergence reads only the +1 face neighbour; the reference performs a full 3-deep
    // gather here (src/VelocityCorrector.hpp:190).
    int rc = cfb_gather( c, CFB_CURRENT );
    if ( rc )
        return rc;
    c->stats.kernel_launches += launch_divergence( c );
    timer_stop( c, PH_RHS );
    return check_async( c, "build_rhs" );
}

int cfb_pcg_solve( cfb_ctx* c, int* num_iter, double* residual_norm )
{
    timer_start( c, PH_PCG );
    int rc = pcg_solve( c, c->cfg.cg_fixed_iters, num_iter, residual_norm );
    timer_stop( c, PH_PCG );
    return rc;
}

int cfb_apply_pressure( cfb_ctx* c )
{
    timer_start( c, PH_APPLY );
    if ( c->cfg.use_nccl )
        halo_exchange_cells( c, c->lhs, 1 ); // _pressure_halo->gather  src/VelocityCorrector.hpp:236
    c->stats.kernel_launches += launch_apply_pressure( c );
    timer_stop( c, PH_APPLY );
    return check_async( c, "apply_pressure" );
}

int cfb_correct_velocity( cfb_ctx* c, int* num_iter, double* residual_norm )
{
    int rc = cfb_build_rhs( c );
    if ( rc )
        return rc;
    rc = cfb_pcg_solve( c, num_iter, residual_norm );
    if ( rc )
        return rc;
    return cfb_apply_pressure( c );
}

int cfb_setup( cfb_ctx* c )
{
    int rc = cfb_add_inputs( c );
    if ( rc )
        return rc;
    return cfb_correct_velocity( c, nullptr, nullptr );
}

int cfb_step( cfb_ctx* c )
{
    int rc = cfb_time_integrator_step( c );
    if ( rc )
        return rc;
    rc = cfb_add_inputs( c );
    if ( rc )
        return rc;
    rc = cfb_correct_velocity( c, nullptr, nullptr );
    c->g.time += c->g.dt; // src/Solver.hpp:146
    c->stats.steps++;
    return rc;
}

int cfb_solve( cfb_ctx* c, double t_final, int write_freq, int* steps_taken )
{
    int t = 0;
    const int num_step = (int)( t_final / c->g.dt ); // src/Solver.hpp:158 (print only)
    // _silo->siloWrite before setup and after every write_freq-th step (src/Solver.hpp:156,170-173),
    // when an output directory has been set (cfb_set_output_dir); both writes of t == 0 go to the same
    // files, the second replaces the first, as in the reference (DB_CLOBBER)
    const char* odir = output_solve_dir( c );
    const std::string out_dir = odir ? odir : "";
    int rc = CFB_OK;
    if ( odir && write_freq > 0 )
    {
        rc = output_write( c, out_dir.c_str(), t );
        if ( rc )
            return rc;
    }
    rc = cfb_setup( c );
    if ( rc )
        return rc;
    do
    {
        if ( c->cfg.world_rank == 0 && write_freq > 0 && 0 == t % write_freq )
            std::printf( "Step %d / %d at time = %f\n", t, num_step, c->g.time );
        rc = cfb_step( c );
        if ( rc )
            return rc;
        if ( odir && write_freq > 0 && 0 == t % write_freq )
        {
            rc = output_write( c, out_dir.c_str(), t );
            if ( rc )
                return rc;
        }
        t++;
    } while ( c->g.time < t_final );
    rc = output_flush( c );
    if ( rc )
        return rc;
    if ( steps_taken )
        *steps_taken = t;
    return CFB_OK;
}

int cfb_pcg_solve_host( cfb_ctx* c, const double* b_host, double* x_host, int* num_iter, double* residual_norm )
{
    int rc = cfb_upload( c, CFB_RHS, CFB_CURRENT, CFB_OWNED, b_host );
    if ( rc )
        return rc;
    rc = cfb_pcg_solve( c, num_iter, residual_norm );
    if ( rc )
        return rc;
    return cfb_download( c, CFB_PRESSURE, CFB_CURRENT, CFB_OWNED, x_host );
}

int cfb_stencil_dot( cfb_ctx* c, int reps, double* dot, double* ms_per_launch )
{
    if ( reps < 1 )
        reps = 1;
    c->sticky_rc = 0;
    // make sure a previous converged solve does not turn the launches into no-ops
    CFB_CUDA( c, cudaMemsetAsync( &c->d_state->done, 0, sizeof( int ), c->stream ) );
    if ( c->cfg.use_nccl )
        note_rc( c, halo_exchange_cells( c, c->cg_p, 1 ) );
    c->stats.kernel_launches += launch_stencil_dot( c ); // warm-up, untimed
    CFB_CUDA( c, cudaEventRecord( c->ev[EV_BENCH0], c->stream ) );
    for ( int i = 0; i < reps; ++i )
        c->stats.kernel_launches += launch_stencil_dot( c );
    CFB_CUDA( c, cudaEventRecord( c->ev[EV_BENCH1], c->stream ) );
    if ( c->cfg.use_nccl )
        note_rc( c, cg_global_sum( c, 0 ) );
    CFB_CUDA( c, cudaMemcpyAsync( c->h_state, c->d_state, STATE_HEAD, cudaMemcpyDeviceToHost, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    int rc = take_sticky_rc( c );
    if ( rc )
        return rc;
    rc = check_async( c, "stencil_dot" );
    if ( rc )
        return rc;
    float ms = 0;
    cudaEventElapsedTime( &ms, c->ev[EV_BENCH0], c->ev[EV_BENCH1] );
    if ( dot )
        *dot = c->h_state->pAp;
    if ( ms_per_launch )
        *ms_per_launch = ms / reps;
    return CFB_OK;
}

int cfb_pcg_fixed( cfb_ctx* c, int iters, double* ms_total, double* residual_norm )
{
    if ( iters < 1 )
        return cfb_fail( c, CFB_ERR_INVALID, "iters must be >= 1" );
    CFB_CUDA( c, cudaEventRecord( c->ev[EV_BENCH0], c->stream ) );
    int it = 0;
    int rc = pcg_solve( c, iters, &it, residual_norm );
    if ( rc )
        return rc;
    CFB_CUDA( c, cudaEventRecord( c->ev[EV_BENCH1], c->stream ) );
    CFB_CUDA( c, cudaEventSynchronize( c->ev[EV_BENCH1] ) );
    float ms = 0;
    cudaEventElapsedTime( &ms, c->ev[EV_BENCH0], c->ev[EV_BENCH1] );
    if ( ms_total )
        *ms_total = ms;
    return CFB_OK;
}

int cfb_fill_synthetic_velocity( cfb_ctx* c, int variant, uint64_t seed )
{
    c->stats.kernel_launches += launch_fill_synthetic( c, variant, seed );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    return check_async( c, "fill_synthetic_velocity" );
}

int cfb_get_stats( const cfb_ctx* c, cfb_stats* out )
{
    cfb_ctx* m = const_cast<cfb_ctx*>( c );
    for ( int s = 0; s < PH_COUNT; ++s )
        timer_collect( m, s );
    *out = c->stats;
    {
        // what a solve would run with the options as they are now
        const int chosen = m->cg_variant;
        m->cg_variant = cg_variant_auto( c );
        out->peer_mode = cg_peer_mode( c ) ? 1 : 0;
        out->peer_overlap = peer_overlapped( c ) ? 1 : 0;
        out->cg_variant = m->cg_variant;
        out->cg_persist = cg_persist_applies( c ) ? 1 : 0;
        m->cg_variant = chosen;
    }
    return CFB_OK;
}
int cfb_reset_stats( cfb_ctx* c )
{
    for ( int s = 0; s < PH_COUNT; ++s )
        timer_collect( c, s );
    c->stats = cfb_stats{};
    return CFB_OK;
}

int cfb_residual_history( const cfb_ctx* c, double* hist, int n, int* count )
{
    cfb_ctx* m = const_cast<cfb_ctx*>( c );
    int total = std::min( c->last_iters, CFB_HIST_MAX );
    int k = std::min( n, total );
    if ( k > 0 )
    {
        cudaError_t e = cudaMemcpy( hist, reinterpret_cast<const char*>( c->d_state ) + STATE_HEAD,
                                    (size_t)k * sizeof( double ), cudaMemcpyDeviceToHost );
        if ( e != cudaSuccess )
            return cfb_fail( m, CFB_ERR_CUDA, cudaGetErrorString( e ) );
    }
    if ( count )
        *count = total;
    return CFB_OK;
}

int cfb_set_cg_params( cfb_ctx* c, double tolerance, int max_iter, int print_level )
{
    if ( tolerance < 0 || max_iter < 0 )
        return cfb_fail( c, CFB_ERR_INVALID, "negative tolerance / max_iter" );
    c->cfg.cg_tolerance = tolerance;
    c->cfg.cg_max_iter = max_iter;
    c->cfg.cg_print_level = print_level;
    return CFB_OK;
}

int cfb_set_tuning( cfb_ctx* c, const char* key, int value )
{
    std::string k = key ? key : "";
    // range of every integer key (switches take any value: non-zero is on).  Tile shapes are set one key at a
    // time, so a combination nobody instantiated is refused where it is used: fused_setup for the extents, the
    // launch for the (tx, ty, stages) triple — pcg_solve / cfb_stencil_dot then return CFB_ERR_INVALID.
    struct Range
    {
        const char* key;
        int lo, hi;
    };
    static const Range ranges[] = { { "stencil_variant", 0, 1 }, { "stencil_tx", 64, 128 }, { "stencil_ty", 8, 32 },
                                    { "stencil_stages", 3, 6 },  { "stencil_zc", 0, 1 << 20 }, { "poll_every", 0, 1 << 20 },
                                    { "cg_variant", -1, 3 },      { "cg_persist", -1, 1 },      { "fused_tx", 64, 128 },   { "fused_ty", 8, 32 },
                                    { "fused_stages", 2, 4 },    { "fused_zc", 0, 1 << 20 }, { "rupdate_ctas", 1, 8 } };
    for ( const Range& r : ranges )
        if ( k == r.key && ( value < r.lo || value > r.hi ) )
            return cfb_fail( c, CFB_ERR_INVALID,
                             "tuning key " + k + ": value " + std::to_string( value ) + " outside [" +
                                 std::to_string( r.lo ) + ", " + std::to_string( r.hi ) + "]" );
    // a set-up that fails puts the previous value back
    auto set_checked = [&]( int& field, int ( *setup )( cfb_ctx* ) ) -> int {
        const int old = field;
        const bool old_auto = c->fu_auto;
        field = value;
        if ( setup == fused_setup )
            c->fu_auto = false;
        const int rc = setup( c );
        if ( rc )
        {
            const std::string msg = c->err;
            field = old;
            c->fu_auto = old_auto;
            setup( c );
            return cfb_fail( c, rc, msg );
        }
        return CFB_OK;
    };
    if ( k == "stencil_variant" )
        c->st_variant = value;
    else if ( k == "stencil_tx" )
        return set_checked( c->st_tx, stencil_setup );
    else if ( k == "stencil_ty" )
        return set_checked( c->st_ty, stencil_setup );
    else if ( k == "stencil_stages" )
        c->st_stages = value;
    else if ( k == "stencil_zc" )
    {
        c->st_zc = value;
        c->st_zc_auto = false;
    }
    else if ( k == "poll_every" )
        c->poll_every = value;
    else if ( k == "cg_variant" )
        c->cg_variant = value;
    else if ( k == "cg_persist" )
        c->cg_persist = value;
    else if ( k == "fused_tx" )
        return set_checked( c->fu_tx, fused_setup );
    else if ( k == "fused_ty" )
        return set_checked( c->fu_ty, fused_setup );
    else if ( k == "fused_stages" )
    {
        c->fu_stages = value;
        c->fu_auto = false;
    }
    else if ( k == "fused_zc" )
        return set_checked( c->fu_zc, fused_setup );
    else if ( k == "fused_yc" )
    {
        if ( value < 1 || value > 4096 )
            return cfb_fail( c, CFB_ERR_INVALID, "tuning key fused_yc: 1 ... 4096 tile rows per unit" );
        const bool old_auto = c->fu_yc_auto;
        const int old = c->fu_yc;
        c->fu_yc_auto = false;
        c->fu_yc = value;
        const int rc = fused_setup( c ); // the unit list follows
        if ( rc )
        {
            const std::string msg = c->err;
            c->fu_yc_auto = old_auto;
            c->fu_yc = old;
            fused_setup( c );
            return cfb_fail( c, rc, msg );
        }
        return CFB_OK;
    }
    else if ( k == "fused_nt" )
    {
        if ( value != 0 && value != 256 && value != 512 )
            return cfb_fail( c, CFB_ERR_INVALID, "tuning key fused_nt: 0 (automatic), 256 or 512" );
        c->fu_nt = value;
    }
    else if ( k == "fused_reverse" )
        c->fu_reverse = value != 0;
    else if ( k == "rupdate_ctas" )
        c->ru_ctas = value;
    else if ( k == "overlap_halo" )
        c->overlap_halo = value != 0;
    else if ( k == "peer_halo" )
    {
        c->use_peer = value != 0;
        // leaving peer mode after a timed-out exchange: clear the sticky error so the NCCL path can run
        if ( c->d_state )
            CFB_CUDA( c, cudaMemsetAsync( &c->d_state->xerror, 0, sizeof( int ), c->stream ) );
    }
    else if ( k == "flat_2d" )
    {
        // the unit list follows the kernels in use (runs of tile rows with the FLAT kernels)
        const bool old = c->flat_2d;
        c->flat_2d = value != 0;
        const int rc = fused_setup( c );
        if ( rc )
        {
            c->flat_2d = old;
            fused_setup( c );
        }
        return rc;
    }
    else if ( k == "advect_tile" )
        c->advect_tile = value != 0;
    else if ( k == "peer_overlap" )
        c->peer_overlap = value != 0;
    else if ( k == "mg_tma" )
    {
        if ( value < -1 || value > 1 )
            return cfb_fail( c, CFB_ERR_INVALID, "tuning key mg_tma: -1 (automatic), 0 or 1" );
        c->mg_tma = value;
        if ( c->mg_tma && c->mg ) // a hierarchy built while the key was off: size the march's scratch now
            return mg_tma_prepare( c );
    }
    else if ( k == "mg_tma_prolong" )
        c->mg_tma_prolong = value != 0;
    else if ( k == "mg_graph" )
        return mg_set_graph( c, value != 0 );
    else if ( k == "mg_coarse_kernel" )
        return mg_set_coarse_kernel( c, value != 0 );
    else if ( k == "peer_xstage" )
        c->peer_xstage_reads = value != 0;
    else if ( k == "time_kernels" )
    {
        c->time_kernels = value != 0;
        if ( c->time_kernels && !c->kev[0][0] )
            for ( auto& row : c->kev )
                for ( auto& e : row )
                    CFB_CUDA( c, cudaEventCreate( &e ) );
    }
    else if ( k == "fused_auto" )
    {
        const bool old = c->fu_auto;
        c->fu_auto = value != 0;
        const int rc = fused_setup( c );
        if ( rc )
            c->fu_auto = old;
        return rc;
    }
    else
        return cfb_fail( c, CFB_ERR_INVALID, "unknown tuning key: " + k );
    return CFB_OK;
}

} // extern "C"
