// emul_glue.cpp — tests/emul only: the emulated launcher, and host stand-ins for the parts of the
// library that cannot be emulated (TMA / mbarrier kernels, NCCL / peer-memory exchange).
//   * q = A p, sum p.q  (the TMA stencil kernel, verified on the GPU)  -> a plain loop with the same row
//     arithmetic (apply_row) and the same exactly-accumulated dot product
//   * the two-kernel fused CG form  -> plain loops with the kernels' semantics (phase A, phase B, finish)
//   * multi-GPU: halo.cu itself is compiled; NCCL is the in-process stand-in of nccl_emul.cpp (ranks are
//     threads), the NVLink peer-memory path reports "not available" (cudaIpc stand-ins fail)
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_cg1.cuh"
#include "device_reduce.cuh"

#include <chrono>
#include <sched.h>
#include <ucontext.h>
#include <vector>

namespace cfb_emul
{
thread_local uint3 g_threadIdx{ 0, 0, 0 }, g_blockIdx{ 0, 0, 0 };
thread_local dim3 g_blockDim, g_gridDim;

namespace
{
struct GraphNode
{
    dim3 grid, block;
    std::function<void()> body;
    bool coop;
};
thread_local bool t_capturing = false;
thread_local std::vector<GraphNode> t_nodes;
} // namespace

void launch( dim3 grid, dim3 block, const std::function<void()>& body )
{
    if ( t_capturing )
    {
        t_nodes.push_back( { grid, block, body, false } );
        return;
    }
    g_gridDim = grid;
    g_blockDim = block;
    for ( unsigned bz = 0; bz < grid.z; ++bz )
        for ( unsigned by = 0; by < grid.y; ++by )
            for ( unsigned bx = 0; bx < grid.x; ++bx )
            {
                g_blockIdx = uint3{ bx, by, bz };
                for ( unsigned tz = block.z; tz-- > 0; )
                    for ( unsigned ty = block.y; ty-- > 0; )
                        for ( unsigned tx = block.x; tx-- > 0; )
                        {
                            g_threadIdx = uint3{ tx, ty, tz };
                            body();
                        }
            }
}

long long clock_ns_real()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>( std::chrono::steady_clock::now().time_since_epoch() ).count();
}

// clock64() of the emulation: only the bounded spins of the peer-memory kernels call it, once per look at a
// flag another rank thread has to set — so it also gives the processor away (the ranks may outnumber the
// cores) and it runs 8x slow, which stretches the kernels' ~10 s timeouts to minutes on a loaded machine
long long clock_ns()
{
    sched_yield();
    return std::chrono::duration_cast<std::chrono::nanoseconds>( std::chrono::steady_clock::now().time_since_epoch() )
               .count() / 8;
}

// ---- cooperative launch: one fiber per CUDA thread -------------------------------------------------
namespace
{
struct Fiber
{
    ucontext_t ctx;
    std::vector<char> stack;
    bool finished = false;
};
struct Coop
{
    bool active = false;
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()>* body = nullptr;
};
thread_local Coop t_coop;

void fiber_entry()
{
    Coop& c = t_coop;
    ( *c.body )();
    c.fibers[c.current].finished = true;
    swapcontext( &c.fibers[c.current].ctx, &c.sched );
}
} // namespace

bool coop_active() { return t_coop.active; }

namespace
{
thread_local char t_smem_anchor;
thread_local __attribute__( ( aligned( 128 ) ) ) unsigned char t_dyn_smem[232 * 1024];
} // namespace
unsigned char* dyn_smem() { return t_dyn_smem; }
uint32_t smem_addr( const void* p )
{
    return (uint32_t)( reinterpret_cast<uintptr_t>( p ) - reinterpret_cast<uintptr_t>( &t_smem_anchor ) );
}
void* smem_ptr( uint32_t a ) { return &t_smem_anchor + (int32_t)a; }

void sync_threads()
{
    Coop& c = t_coop;
    if ( !c.active )
        return; // sequential launch: kernels that get here do not depend on the barrier
    swapcontext( &c.fibers[c.current].ctx, &c.sched ); // yield; resumed when every live fiber has arrived
}

void launch_coop( dim3 grid, dim3 block, const std::function<void()>& body )
{
    if ( t_capturing )
    {
        t_nodes.push_back( { grid, block, body, true } );
        return;
    }
    Coop& c = t_coop;
    const unsigned nt = block.x; // 1-D blocks only
    g_gridDim = grid;
    g_blockDim = block;
    c.body = &body;
    if ( c.fibers.size() < nt )
        c.fibers.resize( nt );
    for ( unsigned bx = 0; bx < grid.x; ++bx )
    {
        g_blockIdx = uint3{ bx, 0, 0 };
        for ( unsigned t = 0; t < nt; ++t )
        {
            Fiber& f = c.fibers[t];
            if ( f.stack.empty() )
                f.stack.resize( 256 * 1024 );
            f.finished = false;
            getcontext( &f.ctx );
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = nullptr;
            makecontext( &f.ctx, fiber_entry, 0 );
        }
        c.active = true;
        unsigned live = nt;
        while ( live > 0 )
        {
            // one pass = every live fiber runs up to its next barrier (or to its end), in thread order
            for ( unsigned t = 0; t < nt; ++t )
            {
                if ( c.fibers[t].finished )
                    continue;
                c.current = (int)t;
                g_threadIdx = uint3{ t, 0, 0 };
                swapcontext( &c.sched, &c.fibers[t].ctx );
                if ( c.fibers[t].finished )
                    --live;
            }
        }
        c.active = false;
    }
}
} // namespace cfb_emul

struct cfb_emul_graph
{
    std::vector<cfb_emul::GraphNode> nodes;
};
cudaError_t cudaStreamBeginCapture( cudaStream_t, cudaStreamCaptureMode )
{
    if ( cfb_emul::t_capturing )
        return cudaErrorEmul;
    cfb_emul::t_nodes.clear();
    cfb_emul::t_capturing = true;
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture( cudaStream_t, cudaGraph_t* g )
{
    if ( !cfb_emul::t_capturing )
        return cudaErrorEmul;
    cfb_emul::t_capturing = false;
    *g = new cfb_emul_graph{ cfb_emul::t_nodes };
    cfb_emul::t_nodes.clear();
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate( cudaGraphExec_t* e, cudaGraph_t g, unsigned long long )
{
    *e = new cfb_emul_graph{ g->nodes };
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch( cudaGraphExec_t e, cudaStream_t )
{
    for ( const auto& n : e->nodes )
    {
        if ( n.coop )
            cfb_emul::launch_coop( n.grid, n.block, n.body );
        else
            cfb_emul::launch( n.grid, n.block, n.body );
    }
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy( cudaGraph_t g )
{
    delete g;
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy( cudaGraphExec_t e )
{
    delete e;
    return cudaSuccess;
}

// cuTensorMapEncodeTiled as far as the product uses it: 3-D float64 tiles, no interleave / swizzle
static CUresult emul_encode_tiled( CUtensorMap* m, CUtensorMapDataType, cuuint32_t rank, void* base,
                                   const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill )
{
    if ( rank != 3 || ( reinterpret_cast<uintptr_t>( base ) & 15 ) || ( gstride[0] & 15 ) || ( gstride[1] & 15 ) ||
         box[0] > 256 || box[1] > 256 || box[2] > 256 || ( box[0] * 8 ) % 16 != 0 )
        return CUDA_ERROR_INVALID_VALUE; // the constraints the real encoder enforces
    m->base = base;
    for ( int d = 0; d < 3; ++d )
    {
        m->dim[d] = gdim[d];
        m->box[d] = box[d];
    }
    m->stride[0] = gstride[0];
    m->stride[1] = gstride[1];
    return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint( const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res )
{
    if ( std::string( name ) == "cuTensorMapEncodeTiled" )
    {
        *fn = reinterpret_cast<void*>( emul_encode_tiled );
        *res = cudaDriverEntryPointSuccess;
    }
    else
    {
        *fn = nullptr;
        *res = cudaDriverEntryPointSymbolNotFound;
    }
    return cudaSuccess;
}

#ifndef CFB_EMUL_REAL_TMA
// the multigrid's fine-level sweeps on the TMA march exist in the "tma" library only: here mg.cu keeps its own kernels
bool mg_tma_applies( const cfb_ctx* ) { return false; }
int mg_tma_prepare( cfb_ctx* ) { return CFB_OK; }
int launch_mg_smooth_tma( cfb_ctx*, const OpConst&, double, double*, double*, double*, int, int ) { return -1; }
int launch_mg_smooth02_tma( cfb_ctx*, const OpConst&, double, double, double*, double*, double* ) { return -1; }
int launch_mg_prolong_smooth_tma( cfb_ctx*, const OpConst&, double, double*, double*, double*, int, double*, long long, long long,
                                  const int*, int )
{
    return -1;
}

int stencil_setup( cfb_ctx* c )
{
    c->tmap_ok = true;
    return CFB_OK;
}

int launch_stencil_dot( cfb_ctx* c )
{
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( S->done )
        return 1;
    const double* p = c->cg_p;
    double* q = c->cg_q;
    dd_t acc{ 0.0, 0.0 };
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                const int w = wall_count( g, 0, i + g.off[0] ) + wall_count( g, 1, j + g.off[1] ) +
                              wall_count( g, 2, k + g.off[2] );
                const double a = apply_row( op.diag[w], op.neg_scale, p[o], p[o - 1], p[o + 1], p[o - g.sy],
                                            p[o + g.sy], p[o - g.sz], p[o + g.sz] );
                q[o] = a;
                dd_acc( acc, p[o] * a );
            }
    // publish_pAp of kernels_stencil.cu: several blocks keep the local double-double for the exact combine
    if ( S->world > 1 )
    {
        S->loc[0] = acc.hi;
        S->loc[1] = acc.lo;
    }
    else
        S->pAp = acc.hi + acc.lo;
    S->rz_old = S->rz_new;
    return 1;
}

// ---- the two-kernel CG form (kernels_fused.cu: TMA, not emulable) as plain loops -------------------------
// Same semantics, statement for statement, as cg_rupdate_kernel / cg_fused_kernel / cg_finish_kernel, so that
// the host orchestration around them (p double-buffering, done / finish logic, polling, and above all the
// NVLink peer-memory exchange, whose kernel IS the product's) runs on the CPU.
int fused_setup( cfb_ctx* c )
{
    c->fused_ok = true;
    return CFB_OK;
}

namespace
{
inline bool cg_converged( const CgState* S ) { return !S->fixed && std::sqrt( S->rr ) <= S->thresh; }
inline int walls_at( const Geo& g, int i, int j, int k )
{
    return wall_count( g, 0, i + g.off[0] ) + wall_count( g, 1, j + g.off[1] ) + wall_count( g, 2, k + g.off[2] );
}
} // namespace

// phase A: alpha = zr_old / pAp ; r -= alpha q ; sum r^2 ; sum r.M^-1 r
int launch_cg_rupdate( cfb_ctx* c )
{
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( S->done || ( S->iter > 0 && cg_converged( S ) ) )
    {
        S->done = 1;
        return 1;
    }
    const double alpha = S->rz_old / S->pAp, nalpha = -alpha;
    S->alpha = alpha;
    dd_t rr{ 0.0, 0.0 }, rz{ 0.0, 0.0 };
    double* r = c->cg_r;
    const double* q = c->cg_q;
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                const double v = fma( nalpha, q[o], r[o] );
                r[o] = v;
                dd_acc( rr, v * v );
                dd_acc( rz, ( op.minv[walls_at( g, i, j, k )] * v ) * v );
            }
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
    return 1;
}

// phase A' of the 64-byte iteration (kernels_stencil.cu MODE 1): q = A p recomputed, never stored
int launch_stencil_rupdate( cfb_ctx* c )
{
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( S->done || ( S->iter > 0 && cg_converged( S ) ) )
    {
        S->done = 1;
        return 1;
    }
    const double alpha = S->rz_old / S->pAp, nalpha = -alpha;
    S->alpha = alpha;
    dd_t rr{ 0.0, 0.0 }, rz{ 0.0, 0.0 };
    double* r = c->cg_r;
    const double* p = c->cg_p;
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                const int w = walls_at( g, i, j, k );
                const double a = apply_row( op.diag[w], op.neg_scale, p[o], p[o - 1], p[o + 1], p[o - g.sy], p[o + g.sy],
                                            p[o - g.sz], p[o + g.sz] );
                const double v = fma( nalpha, a, r[o] );
                r[o] = v;
                dd_acc( rr, v * v );
                dd_acc( rz, ( op.minv[w] * v ) * v );
            }
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
    return 1;
}

// the stencil kernel of the single-reduction CG (kernels_stencil.cu MODE 2): u = M^-1 r, w = A u, three sums
int launch_cg1_stencil( cfb_ctx* c, int init, bool mail )
{
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( mail )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "plain-loop stand-in: no mailbox form of the single-reduction stencil" ) );
        return 0;
    }
    if ( S->done )
        return 1;
    const double* r = c->cg_r;
    double* w = c->cg_q;
    auto u_at = [&]( int i, int j, int k ) { return op.minv[walls_at( g, i, j, k )] * r[geo_off( g, i, j, k )]; };
    dd_t rr{ 0.0, 0.0 }, gm{ 0.0, 0.0 }, dl{ 0.0, 0.0 };
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                const double uc = u_at( i, j, k );
                const double a = apply_row( op.diag[walls_at( g, i, j, k )], op.neg_scale, uc, u_at( i - 1, j, k ), u_at( i + 1, j, k ),
                                            u_at( i, j - 1, k ), u_at( i, j + 1, k ), u_at( i, j, k - 1 ), u_at( i, j, k + 1 ) );
                w[o] = a;
                dd_acc( rr, r[o] * r[o] );
                dd_acc( gm, uc * r[o] );
                dd_acc( dl, uc * a );
            }
    if ( S->world > 1 )
    {
        const dd_t v[3] = { rr, gm, dl };
        for ( int q = 0; q < 3; ++q )
        {
            S->loc[2 * q] = v[q].hi;
            S->loc[2 * q + 1] = v[q].lo;
        }
    }
    else
        cg1_finish( S, rr.hi + rr.lo, gm.hi + gm.lo, dl.hi + dl.lo, init );
    return 1;
}

// the persistent form exists in the TMA kernels only: never chosen with the plain-loop stand-ins
bool cg_persist_supported( const cfb_ctx* ) { return false; }
int launch_cg_persistent( cfb_ctx* c, int )
{
    note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "plain-loop stand-in: no persistent form" ) );
    return 0;
}

int launch_cg_finish( cfb_ctx* c )
{
    CgState* S = c->d_state;
    if ( !S->done && S->iter > 0 && cg_converged( S ) )
        S->done = 1;
    return 1;
}

// phase B: convergence test ; x += alpha p ; beta ; p_new = M^-1 r + beta p (also on the one-cell ghost ring,
// recomputed from the ghosts of r and of the old p, never stored there) ; q = A p_new ; sum p.q
// which: 0 = all units, 1 = "interior" (everything here), 2 = "boundary" (nothing left)
// (overlapped exchange: the ghosts arrive between the two launches, so there "interior" is nothing and the
// boundary launch, launch_cg_fused_mail below, does everything)
int launch_cg_fused( cfb_ctx* c, int which )
{
    if ( which == 2 || ( which == 1 && peer_overlapped( c ) ) )
        return 0;
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( S->done )
        return 1;
    double* x = c->lhs;
    const double* p_old = c->cg_pbuf[c->pcur];
    double* p_new = c->cg_pbuf[c->pcur ^ 1];
    const double* r = c->cg_r;
    double* q = c->cg_q;
    const double alpha = S->alpha;
    const double resid = std::sqrt( S->rr );
    const bool conv = !S->fixed && resid <= S->thresh;
    {
        const int it = S->iter;
        if ( it < CFB_HIST_MAX )
            S->hist[it] = resid;
        S->iter = it + 1;
    }
    if ( conv )
    {
        for ( int k = 0; k < g.n[2]; ++k )
            for ( int j = 0; j < g.n[1]; ++j )
                for ( int i = 0; i < g.n[0]; ++i )
                {
                    const long long o = geo_off( g, i, j, k );
                    x[o] = fma( alpha, p_old[o], x[o] );
                }
        return 1;
    }
    const double beta = S->rz_new / S->rz_old;
    std::vector<double> pn( (size_t)g.total, 0.0 );
    for ( int k = -1; k <= g.n[2]; ++k )
        for ( int j = -1; j <= g.n[1]; ++j )
            for ( int i = -1; i <= g.n[0]; ++i )
            {
                const int out = ( i < 0 || i >= g.n[0] ) + ( j < 0 || j >= g.n[1] ) + ( k < 0 || k >= g.n[2] );
                if ( out > 1 )
                    continue; // edges and corners are never read by the 7-point operator
                const long long o = geo_off( g, i, j, k );
                pn[o] = fma( beta, p_old[o], op.minv[walls_at( g, i, j, k )] * r[o] );
            }
    dd_t acc{ 0.0, 0.0 };
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                x[o] = fma( alpha, p_old[o], x[o] );
                p_new[o] = pn[o];
                const double a = apply_row( op.diag[walls_at( g, i, j, k )], op.neg_scale, pn[o], pn[o - 1], pn[o + 1],
                                            pn[o - g.sy], pn[o + g.sy], pn[o - g.sz], pn[o + g.sz] );
                if ( c->cg_variant != 2 )
                    q[o] = a;
                dd_acc( acc, pn[o] * a );
            }
    if ( S->world > 1 )
    {
        S->loc[0] = acc.hi;
        S->loc[1] = acc.lo;
    }
    else
        S->pAp = acc.hi + acc.lo;
    S->rz_old = S->rz_new;
    return 1;
}
// "peer_overlap" with the plain-loop stand-ins: the kernel followed by the reduction-only exchange its last block runs
int launch_cg_rupdate_mail( cfb_ctx* c )
{
    const int n = launch_cg_rupdate( c );
    peer_exchange( c, 1, false, -1, false );
    return n;
}
int launch_cg_fused_mail( cfb_ctx* c, int which, bool )
{
    if ( which != 2 )
        return 0;
    const int n = launch_cg_fused( c, 0 );
    peer_exchange( c, 0, false, -1, false );
    return n;
}
#endif // !CFB_EMUL_REAL_TMA
