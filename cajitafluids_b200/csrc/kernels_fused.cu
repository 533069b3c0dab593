// kernels_fused.cu — the two-kernel form of one Jacobi-PCG iteration (72 algorithmic bytes per cell
// instead of the 88 of the three-kernel form in kernels_cg.cu / kernels_stencil.cu, and instead of
// the 168 of the reference's stored-coefficient four-kernel form).
//
// Same arithmetic, statement by statement, as Cajita::ReferenceConjugateGradient::solve (SURVEY.md
// §3.3; driven from src/VelocityCorrector.hpp:276), only scheduled differently between the two
// reduction points an iteration of CG has:
//
//   phase A  cg_rupdate   alpha = zr_old / pAp ; r -= alpha q ; sum r^2 ; sum r.M^-1 r
//                         (24 B/cell: read r, q; write r)
//   phase B  cg_fused     convergence test ; x += alpha p (the update phase A deferred: p is read
//                         here anyway) ; beta = zr_new / zr_old ; p = M^-1 r + beta p ; q = A p ;
//                         sum p.q            (48 B/cell: read r, p, x; write p, x, q)
//
// Phase B is the 2.5-D z-marching stencil of kernels_stencil.cu with the p-update moved in front of
// it: the r and p planes of a (TX+4) x (TY+2) box are staged by TMA into a shared-memory ring, the
// new p is computed for the tile AND its one-cell x/y halo ring (recomputed, not exchanged between
// CTAs), kept in a 3-deep shared ring for the x/y neighbours and in registers for the z neighbours.
// x is streamed with 128-bit LDG/STG, prefetched one plane ahead.  The value of every element is
// produced by exactly the same expression as in the three-kernel form, so results are bit-identical
// to it and to the checker.
//
// After every phase B x is complete, so a solve may end after any whole iteration.
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"
#include "device_tma.cuh"
#include "device_peer.cuh"

#include <algorithm>

namespace
{

constexpr int NT = 256;

__device__ __forceinline__ bool cg_converged( const CgState* S )
{
    return !S->fixed && sqrt( S->rr ) <= S->thresh;
}

// ---------------------------------------------------------------------------------------------
// phase A.  Thread layout: 256 threads = TXP (power of two, 32..256) along x times 256/TXP rows; a
// block walks over batches of (256/TXP) * RU rows, every thread issuing the 128-bit loads of its RU
// rows before touching any of them (2 * RU * 16 bytes in flight per thread: this kernel has no
// reuse at all, it lives on memory-level parallelism), one integer division per RU column pairs.
constexpr int RU = 4;

// The rows of phase A for one block of a grid of `nblocks`: r -= alpha q on the block's batches of rows, the thread's
// shares of sum r^2 and sum r.M^-1 r added to rr / rz.  Shared by cg_rupdate_kernel and the persistent kernel.
__device__ __forceinline__ void rupdate_rows( const Geo& g, const OpConst& op, const double* __restrict__ q, double* __restrict__ r,
                                              const double nalpha, const int txp_log2, const int block, const int nblocks,
                                              dd_t& rr, dd_t& rz )
{
    const int txp = 1 << txp_log2, tyr = NT >> txp_log2;
    const int lx = threadIdx.x & ( txp - 1 ), ry = threadIdx.x >> txp_log2;
    const int npx = ( g.n[0] + 1 ) >> 1;
    const int rows = g.n[1] * g.n[2];
    const int batch = tyr * RU;
    const bool odd = g.n[0] & 1;
    for ( int row0 = block * batch; row0 < rows; row0 += nblocks * batch )
    {
        // rows of this thread: row0 + ry + u * tyr
        int cyz[RU];
        long long ro[RU];
        bool ok[RU];
#pragma unroll
        for ( int u = 0; u < RU; ++u )
        {
            const int row = row0 + ry + u * tyr;
            ok[u] = row < rows;
            const int k = row / g.n[1];
            const int j = row - k * g.n[1];
            cyz[u] = wall_count( g, 1, j + g.off[1] ) + wall_count( g, 2, k + g.off[2] );
            ro[u] = geo_off( g, 0, j, k );
        }
        for ( int ip = lx; ip < npx; ip += txp )
        {
            const int i = 2 * ip;
            const bool two = !odd || ip + 1 < npx;
            double2 qv[RU], rv[RU];
#pragma unroll
            for ( int u = 0; u < RU; ++u )
            {
                qv[u] = make_double2( 0.0, 0.0 );
                rv[u] = make_double2( 0.0, 0.0 );
                if ( ok[u] )
                {
                    if ( two )
                    {
                        qv[u] = *reinterpret_cast<const double2*>( q + ro[u] + i );
                        rv[u] = *reinterpret_cast<const double2*>( r + ro[u] + i );
                    }
                    else
                    {
                        qv[u].x = q[ro[u] + i];
                        rv[u].x = r[ro[u] + i];
                    }
                }
            }
            const int wx0 = wall_count( g, 0, i + g.off[0] ), wx1 = wall_count( g, 0, i + 1 + g.off[0] );
#pragma unroll
            for ( int u = 0; u < RU; ++u )
            {
                if ( !ok[u] )
                    continue;
                double2 v = rv[u];
                v.x = fma( nalpha, qv[u].x, v.x );
                dd_acc( rr, v.x * v.x );
                dd_acc( rz, ( op.minv[cyz[u] + wx0] * v.x ) * v.x );
                if ( two )
                {
                    v.y = fma( nalpha, qv[u].y, v.y );
                    *reinterpret_cast<double2*>( r + ro[u] + i ) = v;
                    dd_acc( rr, v.y * v.y );
                    dd_acc( rz, ( op.minv[cyz[u] + wx1] * v.y ) * v.y );
                }
                else
                    r[ro[u] + i] = v.x;
            }
        }
    }
}

template <bool PF>
__global__ void __launch_bounds__( NT, 3 )
    cg_rupdate_kernel( const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                       const double* __restrict__ q, double* __restrict__ r, CgState* S, double* partials,
                       int txp_log2, const __grid_constant__ typename PeerSel<PF>::type pf )
{
    // `done` is only ever written by this kernel, cg_check0 and cg_finish (never by phase B, whose
    // late CTAs would otherwise see it mid-launch); phase B of the iteration that met the tolerance
    // has already applied its x update, so there is nothing left to do.
    if ( S->done || ( S->iter > 0 && cg_converged( S ) ) )
    {
        if ( blockIdx.x == 0 && threadIdx.x == 0 )
            S->done = 1;
        return;
    }
    const double alpha = S->rz_old / S->pAp;
    const double nalpha = -alpha;
    if ( blockIdx.x == 0 && threadIdx.x == 0 )
        S->alpha = alpha;
    dd_t rr = { 0.0, 0.0 }, rz = { 0.0, 0.0 };
    rupdate_rows( g, op, q, r, nalpha, txp_log2, (int)blockIdx.x, (int)gridDim.x, rr, rz );
    dd_t vals[2] = { rr, rz };
    if ( block_reduce_finalize<NT, 2>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
        {
            if ( S->world > 1 )
            {
                S->loc[2] = vals[1].hi;
                S->loc[3] = vals[1].lo;
                S->loc[4] = vals[0].hi;
                S->loc[5] = vals[0].lo;
            }
            else
            {
                S->rr = vals[0].hi + vals[0].lo;
                S->rz_new = vals[1].hi + vals[1].lo;
            }
        }
        if constexpr ( PF )
        {
            dd_t sum[2]; // reduction point 1: (r.z, r.r) from S->loc[2..5]
            peer_mail_exchange<2>( S, pf, 1, 2, sum );
            if ( threadIdx.x == 0 )
            {
                S->rz_new = sum[0].hi + sum[0].lo;
                S->rr = sum[1].hi + sum[1].lo;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// end of a solve: after any phase B x is up to date; only the `done` flag may be missing when the
// tolerance was met by the very last iteration enqueued (no phase A followed to record it).
__global__ void cg_finish_kernel( CgState* S )
{
    if ( !S->done && S->iter > 0 && cg_converged( S ) )
        S->done = 1;
}

// ---------------------------------------------------------------------------------------------
// phase B
template <int TX_, int TY_, int NS_, int NT_ = 256>
struct FusedCfg
{
    static constexpr int TX = TX_, TY = TY_, NS = NS_;
    static constexpr int NT = NT_; // (the persistent kernel's phase A assumes 256)
    static constexpr int LX = TX / 2;  // threads along x (one column pair each)
    static constexpr int WY = NT / LX; // thread rows
    static constexpr int RY = TY / WY; // rows per thread
    static constexpr int PX = TX + 4;  // smem row pitch: 2-wide x halo keeps pairs 16-B aligned
    static constexpr int PY = TY + 2;
    static constexpr int BOX_BYTES = PX * PY * 8;
    static constexpr int BOX_PAD = ( BOX_BYTES + 127 ) / 128 * 128;
    static constexpr int STAGE_BYTES = 2 * BOX_PAD; // r box, p box
    static constexpr int NPN = 3;                   // new-p planes kept in shared memory
    static constexpr int SMEM_BYTES = NS * STAGE_BYTES + NPN * BOX_PAD + 128;
    static constexpr int CTAS = SMEM_BYTES <= 56 * 1024 ? 4 : ( SMEM_BYTES <= 75 * 1024 ? 3 : ( SMEM_BYTES <= 113 * 1024 ? 2 : 1 ) );
    static_assert( TX % 2 == 0 && NT % LX == 0 && TY % WY == 0 && RY >= 1 && LX % 32 == 0, "bad tile" );
};

struct FusedArgs
{
    double *x, *p, *q; // p: the buffer the NEW search direction is written to
    const double* p_old;
    CgState* S;
    double* partials;
    const int* units; // optional unit list (tile_x, tile_y, chunk) triples; nullptr = all units in order
    int tiles_x, tiles_y, zc, hx;
    int yc;          // two-dimensional runs (FLAT): tile rows a unit marches through along y (fused_unit_flat); else 1
    int reverse;     // walk the units from the top of the block down (see launch_cg_fused)
    // XS kernels (peer mode, x neighbours): the x ghosts of r and of the old p are read from the dense
    // staging areas [k * ny + j] the x neighbours fill over NVLink, [0] low side, [1] high side
    const double* gxr[2];
    const double* gxp[2];
    int store_q;     // 0: the 64-byte iteration (cg_variant 2) recomputes q = A p in its phase A' and never reads it
    int unit_base;   // index of this launch's first unit in the partial-sum scratch
    int units_total; // units of all launches that make up one phase B (last-block ticket target)
};

// One unit of phase B: the z-march over the planes [kbeg, kend) of tile (x0, y0) — new search direction on the tile
// and its halo ring, x += alpha p_old, q = A p_new — for the calling block; returns the thread's share of p.q.
// Shared by the one-unit-per-block kernel (lbase == 0) and by the persistent kernel of the small grids, whose blocks
// walk through several units with the same shared-memory ring and stage barriers (lbase: loads issued so far).
template <class C, bool XS>
__device__ __forceinline__ dd_t fused_unit( const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const Geo& g, const OpConst& op,
                                            const FusedArgs& a, const int x0, const int y0, const int kbeg, const int kend,
                                            const double alpha, const double beta, double* stage0, double* pn0,
                                            unsigned long long* full_bar, const uint32_t smem_base, const int lbase )
{
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS, PX = C::PX, RY = C::RY, WY = C::WY, LX = C::LX;
    constexpr int BOXD = C::BOX_PAD / 8, STAGED = C::STAGE_BYTES / 8;
    const int tid = threadIdx.x;
    const int lx = tid % LX, wy = tid / LX;
    const int nplanes = kend - kbeg;
    const int nloads = nplanes + 2; // planes kbeg-1 .. kend
    const int i0 = x0 + 2 * lx;
    const bool vx0 = i0 < g.n[0], vx1 = i0 + 1 < g.n[0];
    bool vy[RY];
#pragma unroll
    for ( int r = 0; r < RY; ++r )
        vy[r] = y0 + wy + r * WY < g.n[1];

    // TMA box origin (array coordinates): 2 columns left of the tile, 1 row below, plane kbeg-1
    const int cx = a.hx + x0 - 2;
    const int cy = g.h + y0 - 1;
    const int cz = g.h + kbeg - 1;

    // load index across the units a persistent block walks through: every stage barrier completes once per load, so
    // the parity a waiter needs is that of the running index (lbase == 0 in the one-unit-per-block kernels)
    auto issue = [&]( int l ) {
        const int s = ( lbase + l ) % NS;
        const uint32_t bar = smem_u32( &full_bar[s] );
        mbar_expect_tx( bar, 2 * C::BOX_BYTES );
        tma_load_3d( smem_base + s * C::STAGE_BYTES, &tmap_r, bar, cx, cy, cz + l );
        tma_load_3d( smem_base + s * C::STAGE_BYTES + C::BOX_PAD, &tmap_p, bar, cx, cy, cz + l );
    };

    if ( tid == 0 )
    {
        const int n0 = nloads < NS ? nloads : NS;
        for ( int l = 0; l < n0; ++l )
            issue( l );
    }

    // per-thread constants: SOLID-wall counts of my cells and of my share of the halo ring
    const int wx0 = wall_count( g, 0, i0 + g.off[0] );
    const int wx1 = wall_count( g, 0, i0 + 1 + g.off[0] );
    int wyc[RY];
#pragma unroll
    for ( int r = 0; r < RY; ++r )
        wyc[r] = wall_count( g, 1, y0 + wy + r * WY + g.off[1] );
    // y halo rows: thread row 0 takes the row below the tile, thread row WY-1 the row above
    const bool do_yh = ( wy == 0 ) || ( wy == WY - 1 );
    const int yh_row = ( wy == 0 ) ? 0 : TY + 1;
    const int yh_w = wall_count( g, 1, y0 + yh_row - 1 + g.off[1] );
    // x halo columns: 2 * TY single cells, taken by the lanes of one middle warp
    constexpr int XH_WARP = ( C::NT / 32 ) / 2;
    const bool do_xh = ( tid >> 5 ) == XH_WARP;
    const double ns = op.neg_scale;

    // x of my cells, prefetched one plane ahead
    double2 xn[RY];
    double* xrow = a.x + geo_off( g, i0, y0 + wy, kbeg );
    double* prow = a.p + geo_off( g, i0, y0 + wy, kbeg );
    double* qrow = a.q + geo_off( g, i0, y0 + wy, kbeg );
    auto load_x = [&]( const double* xr ) {
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            xn[r] = make_double2( 0.0, 0.0 );
            if ( vy[r] )
            {
                const double* xp = xr + (long long)( r * WY ) * g.sy;
                if ( vx1 )
                    xn[r] = *reinterpret_cast<const double2*>( xp );
                else if ( vx0 )
                    xn[r].x = *xp;
            }
        }
    };
    load_x( xrow );

    // new p of plane-load l for my cells (-> out[]); with `full` also for my share of the halo ring
    // into the shared new-p plane, and p, x of my cells are written back (plane owned by this chunk).
    double2 zm[RY], cc[RY], zp[RY];
    auto new_p = [&]( int l, double2 out[RY], bool full ) {
        const double* R = stage0 + ( ( lbase + l ) % NS ) * STAGED;
        const double* P = R + BOXD;
        double* PN = pn0 + ( l % C::NPN ) * BOXD;
        const int wz = wall_count( g, 2, kbeg + l - 1 + g.off[2] );
        double2 xc[RY];
        if ( full )
        {
#pragma unroll
            for ( int r = 0; r < RY; ++r )
                xc[r] = xn[r];
            if ( l < nplanes ) // next owned plane
                load_x( xrow + g.sz );
        }
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            const int o = ( wy + r * WY + 1 ) * PX + 2 * lx + 2;
            const double2 rv = *reinterpret_cast<const double2*>( R + o );
            const double2 pv = *reinterpret_cast<const double2*>( P + o );
            double2 v;
            v.x = fma( beta, pv.x, op.minv[wx0 + wyc[r] + wz] * rv.x );
            v.y = fma( beta, pv.y, op.minv[wx1 + wyc[r] + wz] * rv.y );
            out[r] = v;
            if ( full )
            {
                *reinterpret_cast<double2*>( PN + o ) = v;
                if ( vy[r] )
                {
                    const long long go = (long long)( r * WY ) * g.sy;
                    if ( vx1 )
                    {
                        *reinterpret_cast<double2*>( prow + go ) = v;
                        double2 xv = xc[r];
                        xv.x = fma( alpha, pv.x, xv.x );
                        xv.y = fma( alpha, pv.y, xv.y );
                        *reinterpret_cast<double2*>( xrow + go ) = xv;
                    }
                    else if ( vx0 )
                    {
                        prow[go] = v.x;
                        xrow[go] = fma( alpha, pv.x, xc[r].x );
                    }
                }
            }
        }
        if ( full )
        {
            if ( do_yh )
            {
                const int o = yh_row * PX + 2 * lx + 2;
                const double2 rv = *reinterpret_cast<const double2*>( R + o );
                const double2 pv = *reinterpret_cast<const double2*>( P + o );
                double2 v;
                v.x = fma( beta, pv.x, op.minv[wx0 + yh_w + wz] * rv.x );
                v.y = fma( beta, pv.y, op.minv[wx1 + yh_w + wz] * rv.y );
                *reinterpret_cast<double2*>( PN + o ) = v;
            }
            if ( do_xh )
            {
                for ( int it = tid & 31; it < 2 * TY; it += 32 )
                {
                    const int side = it / TY, row = it - side * TY; // side 0: column x0-1, 1: x0+TX
                    const int col = side ? TX + 2 : 1;
                    const int gi = x0 + ( side ? TX : -1 ) + g.off[0];
                    const int cw = wall_count( g, 0, gi ) + wall_count( g, 1, y0 + row + g.off[1] ) + wz;
                    const int o = ( row + 1 ) * PX + col;
                    double rv = R[o], pv = P[o];
                    if ( XS )
                    {
                        const bool ghost = side ? ( x0 + TX == g.n[0] && a.gxr[1] ) : ( x0 == 0 && a.gxr[0] );
                        if ( ghost && y0 + row < g.n[1] )
                        {
                            const size_t e = (size_t)( kbeg + l - 1 ) * g.n[1] + ( y0 + row );
                            rv = a.gxr[side][e];
                            pv = a.gxp[side][e];
                        }
                    }
                    PN[o] = fma( beta, pv, op.minv[cw] * rv );
                }
            }
            xrow += g.sz;
            prow += g.sz;
        }
    };

    // prologue: plane kbeg-1 (z neighbour only), plane kbeg (first owned plane)
    mbar_wait( smem_u32( &full_bar[lbase % NS] ), ( lbase / NS ) & 1 );
    new_p( 0, zm, false );
    mbar_wait( smem_u32( &full_bar[( lbase + 1 ) % NS] ), ( ( lbase + 1 ) / NS ) & 1 );
    new_p( 1, cc, true );
    __syncthreads();
    if ( tid == 0 )
    {
        if ( NS < nloads )
            issue( NS );
        if ( NS + 1 < nloads && NS > 1 )
            issue( NS + 1 );
    }

    dd_t acc = { 0.0, 0.0 };
    for ( int it = 0; it < nplanes; ++it )
    {
        const int l = it + 2; // plane k+1
        mbar_wait( smem_u32( &full_bar[( lbase + l ) % NS] ), ( ( lbase + l ) / NS ) & 1 );
        new_p( l, zp, l <= nplanes );
        // q = A p on plane k = kbeg + it: x/y neighbours from the shared new-p plane of load it+1
        const double* PN = pn0 + ( ( it + 1 ) % C::NPN ) * BOXD;
        const int wz = wall_count( g, 2, kbeg + it + g.off[2] );
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            const double* pc = PN + ( wy + r * WY + 1 ) * PX + 2 * lx + 2;
            const double xl = pc[-1];
            const double xr = pc[2];
            const double2 ym = *reinterpret_cast<const double2*>( pc - PX );
            const double2 yp = *reinterpret_cast<const double2*>( pc + PX );
            const double2 c = cc[r];
            const double d0 = op.diag[wx0 + wyc[r] + wz];
            const double d1 = op.diag[wx1 + wyc[r] + wz];
            const double a0 = apply_row( d0, ns, c.x, xl, c.y, ym.x, yp.x, zm[r].x, zp[r].x );
            const double a1 = apply_row( d1, ns, c.y, c.x, xr, ym.y, yp.y, zm[r].y, zp[r].y );
            if ( vy[r] )
            {
                double* qp = qrow + (long long)( r * WY ) * g.sy;
                if ( vx1 )
                {
                    if ( a.store_q )
                        *reinterpret_cast<double2*>( qp ) = make_double2( a0, a1 );
                    dd_acc( acc, c.x * a0 );
                    dd_acc( acc, c.y * a1 );
                }
                else if ( vx0 )
                {
                    if ( a.store_q )
                        *qp = a0;
                    dd_acc( acc, c.x * a0 );
                }
            }
            zm[r] = c;
            cc[r] = zp[r];
        }
        qrow += g.sz;
        __syncthreads(); // stage of load l is consumed, new-p plane of load l is complete
        if ( tid == 0 && l + NS < nloads )
            issue( l + NS );
    }

    return acc;
}

// One unit of phase B of a two-dimensional run (FLAT): there is ONE owned plane, so the march goes along y instead —
// a unit is a run of `ntile` tiles (x0, y0 + t TY), their boxes of r and p travel through the same shared-memory ring
// NS - 1 tiles ahead of the one being computed, and the prologue, the reduction and the launch of a block are paid
// once per run instead of once per tile.  (One tile per block, the round-2 form, sat at 2.7 TB/s whatever the tile
// shape — profiles/r2_sweep_2d_tilings.log: a block's life was one load latency after the other.)  Every tile
// recomputes its own halo ring of the new search direction like a 3-D plane does; z neighbours are the zero ghost
// planes.  Same statements on the same values as the one-tile form: the same bits.
template <class C, bool XS>
__device__ __forceinline__ dd_t fused_unit_flat( const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const Geo& g, const OpConst& op,
                                                 const FusedArgs& a, const int x0, const int y0, const int ntile,
                                                 const double alpha, const double beta, double* stage0, double* pn0,
                                                 unsigned long long* full_bar, const uint32_t smem_base, const int lbase )
{
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS, PX = C::PX, RY = C::RY, WY = C::WY, LX = C::LX;
    constexpr int BOXD = C::BOX_PAD / 8, STAGED = C::STAGE_BYTES / 8;
    const int tid = threadIdx.x;
    const int lx = tid % LX, wy = tid / LX;
    const int i0 = x0 + 2 * lx;
    const bool vx0 = i0 < g.n[0], vx1 = i0 + 1 < g.n[0];

    // TMA box origin of tile 0 (array coordinates): 2 columns left of the tile, 1 row below, the owned plane
    const int cx = a.hx + x0 - 2;
    const int cy = g.h + y0 - 1;
    const int cz = g.h;
    // load t = tile t of the run; running index across the units a persistent block walks through (see fused_unit)
    auto issue = [&]( int t ) {
        const int s = ( lbase + t ) % NS;
        const uint32_t bar = smem_u32( &full_bar[s] );
        mbar_expect_tx( bar, 2 * C::BOX_BYTES );
        tma_load_3d( smem_base + s * C::STAGE_BYTES, &tmap_r, bar, cx, cy + t * TY, cz );
        tma_load_3d( smem_base + s * C::STAGE_BYTES + C::BOX_PAD, &tmap_p, bar, cx, cy + t * TY, cz );
    };
    if ( tid == 0 )
    {
        const int n0 = ntile < NS ? ntile : NS;
        for ( int t = 0; t < n0; ++t )
            issue( t );
    }

    const int wx0 = wall_count( g, 0, i0 + g.off[0] );
    const int wx1 = wall_count( g, 0, i0 + 1 + g.off[0] );
    const int wz = wall_count( g, 2, g.off[2] );
    const bool do_yh = ( wy == 0 ) || ( wy == WY - 1 );
    const int yh_row = ( wy == 0 ) ? 0 : TY + 1;
    constexpr int XH_WARP = ( C::NT / 32 ) / 2;
    const bool do_xh = ( tid >> 5 ) == XH_WARP;
    const double ns = op.neg_scale;

    // x of my cells, prefetched one tile ahead
    double2 xn[RY];
    double* xrow = a.x + geo_off( g, i0, y0 + wy, 0 );
    double* prow = a.p + geo_off( g, i0, y0 + wy, 0 );
    double* qrow = a.q + geo_off( g, i0, y0 + wy, 0 );
    auto load_x = [&]( const double* xr, int yt ) {
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            xn[r] = make_double2( 0.0, 0.0 );
            if ( yt + wy + r * WY < g.n[1] )
            {
                const double* xp = xr + (long long)( r * WY ) * g.sy;
                if ( vx1 )
                    xn[r] = *reinterpret_cast<const double2*>( xp );
                else if ( vx0 )
                    xn[r].x = *xp;
            }
        }
    };
    load_x( xrow, y0 );

    dd_t acc = { 0.0, 0.0 };
    for ( int t = 0; t < ntile; ++t )
    {
        const int yt = y0 + t * TY;
        const long long tstep = (long long)TY * g.sy;
        mbar_wait( smem_u32( &full_bar[( lbase + t ) % NS] ), ( ( lbase + t ) / NS ) & 1 );
        const double* R = stage0 + ( ( lbase + t ) % NS ) * STAGED;
        const double* P = R + BOXD;
        double* PN = pn0 + ( t % C::NPN ) * BOXD;
        double2 xc[RY], cc[RY];
        bool vy[RY];
        int wyc[RY];
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            xc[r] = xn[r];
            vy[r] = yt + wy + r * WY < g.n[1];
            wyc[r] = wall_count( g, 1, yt + wy + r * WY + g.off[1] );
        }
        if ( t + 1 < ntile )
            load_x( xrow + tstep, yt + TY );
        // new search direction of the tile: my cells (-> registers, shared plane, global; x += alpha p_old) ...
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            const int o = ( wy + r * WY + 1 ) * PX + 2 * lx + 2;
            const double2 rv = *reinterpret_cast<const double2*>( R + o );
            const double2 pv = *reinterpret_cast<const double2*>( P + o );
            double2 v;
            v.x = fma( beta, pv.x, op.minv[wx0 + wyc[r] + wz] * rv.x );
            v.y = fma( beta, pv.y, op.minv[wx1 + wyc[r] + wz] * rv.y );
            cc[r] = v;
            *reinterpret_cast<double2*>( PN + o ) = v;
            if ( vy[r] )
            {
                const long long go = (long long)( r * WY ) * g.sy;
                if ( vx1 )
                {
                    *reinterpret_cast<double2*>( prow + go ) = v;
                    double2 xv = xc[r];
                    xv.x = fma( alpha, pv.x, xv.x );
                    xv.y = fma( alpha, pv.y, xv.y );
                    *reinterpret_cast<double2*>( xrow + go ) = xv;
                }
                else if ( vx0 )
                {
                    prow[go] = v.x;
                    xrow[go] = fma( alpha, pv.x, xc[r].x );
                }
            }
        }
        // ... and my share of its halo ring (shared plane only)
        if ( do_yh )
        {
            const int yh_w = wall_count( g, 1, yt + yh_row - 1 + g.off[1] );
            const int o = yh_row * PX + 2 * lx + 2;
            const double2 rv = *reinterpret_cast<const double2*>( R + o );
            const double2 pv = *reinterpret_cast<const double2*>( P + o );
            double2 v;
            v.x = fma( beta, pv.x, op.minv[wx0 + yh_w + wz] * rv.x );
            v.y = fma( beta, pv.y, op.minv[wx1 + yh_w + wz] * rv.y );
            *reinterpret_cast<double2*>( PN + o ) = v;
        }
        if ( do_xh )
        {
            for ( int it = tid & 31; it < 2 * TY; it += 32 )
            {
                const int side = it / TY, row = it - side * TY; // side 0: column x0-1, 1: x0+TX
                const int col = side ? TX + 2 : 1;
                const int gi = x0 + ( side ? TX : -1 ) + g.off[0];
                const int cw = wall_count( g, 0, gi ) + wall_count( g, 1, yt + row + g.off[1] ) + wz;
                const int o = ( row + 1 ) * PX + col;
                double rv = R[o], pv = P[o];
                if ( XS )
                {
                    const bool ghost = side ? ( x0 + TX == g.n[0] && a.gxr[1] ) : ( x0 == 0 && a.gxr[0] );
                    if ( ghost && yt + row < g.n[1] )
                    {
                        const size_t e = (size_t)( yt + row );
                        rv = a.gxr[side][e];
                        pv = a.gxp[side][e];
                    }
                }
                PN[o] = fma( beta, pv, op.minv[cw] * rv );
            }
        }
        __syncthreads(); // the stage of tile t is consumed, its new-p plane is complete
        if ( tid == 0 && t + NS < ntile )
            issue( t + NS );
        // q = A p on the tile: x/y neighbours from the shared new-p plane, z neighbours are the zero ghost planes
        // (plane t % NPN is written again by tile t + NPN, two block barriers after this one: NPN >= 2 suffices)
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            const double* pc = PN + ( wy + r * WY + 1 ) * PX + 2 * lx + 2;
            const double xl = pc[-1];
            const double xr = pc[2];
            const double2 ym = *reinterpret_cast<const double2*>( pc - PX );
            const double2 yp = *reinterpret_cast<const double2*>( pc + PX );
            const double2 c = cc[r];
            const double d0 = op.diag[wx0 + wyc[r] + wz];
            const double d1 = op.diag[wx1 + wyc[r] + wz];
            const double a0 = apply_row( d0, ns, c.x, xl, c.y, ym.x, yp.x, 0.0, 0.0 );
            const double a1 = apply_row( d1, ns, c.y, c.x, xr, ym.y, yp.y, 0.0, 0.0 );
            if ( vy[r] )
            {
                double* qp = qrow + (long long)( r * WY ) * g.sy;
                if ( vx1 )
                {
                    if ( a.store_q )
                        *reinterpret_cast<double2*>( qp ) = make_double2( a0, a1 );
                    dd_acc( acc, c.x * a0 );
                    dd_acc( acc, c.y * a1 );
                }
                else if ( vx0 )
                {
                    if ( a.store_q )
                        *qp = a0;
                    dd_acc( acc, c.x * a0 );
                }
            }
        }
        xrow += tstep;
        prow += tstep;
        qrow += tstep;
    }
    // a persistent block goes on to its next run: its first tile writes new-p plane 0, which the last tile of a short
    // run (ntile = 1, 4, ...) may still be reading
    __syncthreads();
    return acc;
}

// FLAT: two-dimensional runs (one owned plane between two zero ghost planes, see Geo): the z neighbours are
// zero by construction, so the two ghost planes are neither loaded nor recomputed — the 2-D traffic of r and p
// drops from three planes to the one that exists.  A template flag: the 3-D instantiations are untouched.
// PF: the block that draws the phase's last ticket runs the mailbox reduction of p.Ap (device_peer.cuh); a template
// flag, the other instantiations are untouched.
template <class C, bool XS, bool FLAT, bool PF>
__global__ void __launch_bounds__( C::NT, C::CTAS )
    cg_fused_kernel( const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_p,
                     const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                     const __grid_constant__ FusedArgs a, const __grid_constant__ typename PeerSel<PF>::type pf )
{
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS, PX = C::PX, RY = C::RY, WY = C::WY, LX = C::LX;
    constexpr int BOXD = C::BOX_PAD / 8, STAGED = C::STAGE_BYTES / 8;
    CgState* S = a.S;
    if ( S->done )
        return;

    const int tid = threadIdx.x;
    const int lx = tid % LX, wy = tid / LX;

    // unit -> (tile_x, tile_y, z chunk)
    int tx, ty, ch;
    if ( a.units )
    {
        tx = a.units[3 * blockIdx.x + 0];
        ty = a.units[3 * blockIdx.x + 1];
        ch = a.units[3 * blockIdx.x + 2];
    }
    else
    {
        const int u = a.reverse ? a.units_total - 1 - (int)blockIdx.x : (int)blockIdx.x;
        tx = u % a.tiles_x;
        ty = ( u / a.tiles_x ) % a.tiles_y;
        ch = u / ( a.tiles_x * a.tiles_y );
    }
    // FLAT: `ty` counts runs of a.yc tile rows (the unit marches along y: fused_unit_flat)
    const int x0 = tx * TX, y0 = ty * ( FLAT ? a.yc : 1 ) * TY;
    const int ntile = FLAT ? min( a.yc, ( g.n[1] + TY - 1 ) / TY - ty * a.yc ) : 1;
    const int kbeg = ch * a.zc;
    const int kend = min( kbeg + a.zc, g.n[2] );

    const int i0 = x0 + 2 * lx;
    const bool vx0 = i0 < g.n[0], vx1 = i0 + 1 < g.n[0];

    // ---- convergence bookkeeping of the iteration (the reference's test after kernel 1) ----------
    const double alpha = S->alpha;
    const double resid = sqrt( S->rr );
    const bool conv = !S->fixed && resid <= S->thresh;
    if ( a.unit_base + blockIdx.x == 0 && tid == 0 )
    {
        const int it = S->iter;
        if ( it < CFB_HIST_MAX )
            S->hist[it] = resid;
        S->iter = it + 1;
    }
    if ( conv )
    {
        // x += alpha p of the converged iteration; nothing else (the loop breaks here)
        for ( int k = kbeg; k < kend; ++k )
          for ( int t = 0; t < ntile; ++t )
#pragma unroll
            for ( int r = 0; r < RY; ++r )
            {
                const int j = y0 + t * TY + wy + r * WY;
                if ( j >= g.n[1] || !vx0 )
                    continue;
                const long long o = geo_off( g, i0, j, k );
                if ( vx1 )
                {
                    const double2 pv = *reinterpret_cast<const double2*>( a.p_old + o );
                    double2 xv = *reinterpret_cast<double2*>( a.x + o );
                    xv.x = fma( alpha, pv.x, xv.x );
                    xv.y = fma( alpha, pv.y, xv.y );
                    *reinterpret_cast<double2*>( a.x + o ) = xv;
                }
                else
                    a.x[o] = fma( alpha, a.p_old[o], a.x[o] );
            }
        return;
    }
    const double beta = S->rz_new / S->rz_old;

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__( 8 ) unsigned long long full_bar[NS];
    const uint32_t smem_base = ( smem_u32( smem_raw ) + 127u ) & ~127u;
    double* stage0 = reinterpret_cast<double*>( smem_raw + ( smem_base - smem_u32( smem_raw ) ) );
    double* pn0 = stage0 + NS * STAGED;

    if ( tid == 0 )
    {
        prefetch_tmap( &tmap_r );
        prefetch_tmap( &tmap_p );
#pragma unroll
        for ( int s = 0; s < NS; ++s )
            mbar_init( smem_u32( &full_bar[s] ), 1 );
        fence_barrier_init();
    }
    __syncthreads();
    dd_t acc;
    if constexpr ( FLAT )
        acc = fused_unit_flat<C, XS>( tmap_r, tmap_p, g, op, a, x0, y0, ntile, alpha, beta, stage0, pn0, full_bar, smem_base, 0 );
    else
        acc = fused_unit<C, XS>( tmap_r, tmap_p, g, op, a, x0, y0, kbeg, kend, alpha, beta, stage0, pn0, full_bar, smem_base, 0 );

    // p.Ap: block partials at [unit_base + blockIdx.x], the block drawing the last of
    // `units_total` tickets finalises (deterministic: double-double sums, order-independent)
    {
        __shared__ dd_t s_red[C::NT / 32];
        __shared__ bool s_last;
        dd_t s = dd_block_sum<C::NT>( acc, s_red );
        const unsigned bid = (unsigned)a.unit_base + blockIdx.x;
        if ( tid == 0 )
        {
            a.partials[(size_t)bid * 2 + 0] = s.hi;
            a.partials[(size_t)bid * 2 + 1] = s.lo;
            __threadfence();
            const unsigned t = atomicAdd( &S->ticket[1], 1u );
            s_last = ( t == (unsigned)a.units_total - 1 );
        }
        __syncthreads();
        if ( !s_last )
            return;
        __threadfence();
        dd_t tot = { 0.0, 0.0 };
        for ( unsigned b = tid; b < (unsigned)a.units_total; b += C::NT )
        {
            dd_t w;
            w.hi = __ldcg( a.partials + (size_t)b * 2 + 0 );
            w.lo = __ldcg( a.partials + (size_t)b * 2 + 1 );
            tot = dd_add( tot, w );
        }
        tot = dd_block_sum<C::NT>( tot, s_red );
        if ( tid == 0 )
        {
            S->ticket[1] = 0u;
            if ( S->world > 1 )
            {
                S->loc[0] = tot.hi;
                S->loc[1] = tot.lo;
            }
            else
                S->pAp = tot.hi + tot.lo;
            S->rz_old = S->rz_new; // "zTr_old = zTr_new"
        }
        if constexpr ( PF )
        {
            dd_t sum[1];
            peer_mail_exchange<1>( S, pf, 0, 0, sum );
            if ( tid == 0 )
                S->pAp = sum[0].hi + sum[0].lo;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent form of the two-kernel iteration for blocks whose CG vectors stay in the L2 (64^3 ... ~160^3): there
// an iteration is not bandwidth- but latency-bound — every kernel of the launch-per-phase form pays its launch, its
// prologue and the tail of its last-block reduction (~12.7 us per kernel whatever the size, measured at 64^3:
// profiles/r1_sweep_fused2.log), i.e. ~25 of the ~41 us of an iteration at 128^3.  Here ONE cooperative launch runs
// a batch of iterations: phase A on the block's rows, grid barrier (grid_barrier below), every block adds the block partials (exact
// double-double sums: the same value everywhere), phase B on the block's units (fused_unit above, the shared-memory
// ring and its stage barriers carried from unit to unit), grid barrier, sums again.  The CG scalars live in
// registers, identical in every block; block 0 writes them back at the end, in the layout the launch-per-phase
// kernels read, so the two forms can take turns inside one solve.  Same statements on the same values: the results
// are bit-identical to the other forms and to the checker.
struct PersistArgs
{
    double *x, *r, *q;
    double* pbuf[2];
    int pcur; // pbuf[pcur] holds the search direction when the launch starts
    CgState* S;
    double* partials;
    int pstride; // entries per value of the partial-sum scratch: values 0, 1 phase A, value 2 phase B, per BLOCK
    int tiles_x, tiles_y, zc, hx, units_total;
    int yc; // FLAT: tile rows per unit (see FusedArgs)
    int txp_log2;
    int niters; // iterations of this launch (fewer when the tolerance is met)
};

__device__ __forceinline__ void fence_proxy_async_global()
{
#ifdef __CUDA_ARCH__
    // my generic-proxy stores (r, p, x, q) before the TMA (async-proxy) loads other blocks issue behind the barrier
    asm volatile( "fence.proxy.async;" ::: "memory" );
#endif
}

// Barrier over the blocks of a cooperative launch (all resident): arrivals are counted, the last arrival resets the
// count and bumps the generation the others spin on.  One atomic and one spinning thread per block — cheaper than
// the general-purpose cooperative-groups barrier where an iteration has ~10 us to spend on two of them.  The fences
// make every thread's earlier global stores visible to every block behind the barrier (cumulativity through the
// block barriers on either side).  Bounded: a block that never arrives surfaces as an error, not as a hung GPU.
__device__ __forceinline__ void grid_barrier( CgState* S, const unsigned nblocks, unsigned& gen )
{
    __syncthreads();
    if ( threadIdx.x == 0 )
    {
        __threadfence();
        volatile unsigned* vgen = &S->gbar[1];
        if ( atomicAdd( &S->gbar[0], 1u ) == nblocks - 1 )
        {
            S->gbar[0] = 0u;
            __threadfence();
            atomicAdd( &S->gbar[1], 1u );
        }
        else
        {
            const long long t0 = clock64();
            while ( *vgen == gen )
                if ( clock64() - t0 > 8000000000ll )
                {
                    S->xerror = 1;
                    break;
                }
        }
        __threadfence();
    }
    ++gen;
    __syncthreads();
}

template <class C, bool FLAT>
__global__ void __launch_bounds__( C::NT, C::CTAS )
    cg_persistent_kernel( const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_p0,
                          const __grid_constant__ CUtensorMap tmap_p1, const __grid_constant__ Geo g,
                          const __grid_constant__ OpConst op, const __grid_constant__ PersistArgs a )
{
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS;
    constexpr int STAGED = C::STAGE_BYTES / 8;
    CgState* S = a.S;
    const int tid = threadIdx.x;
    const int bid = (int)blockIdx.x, nb = (int)gridDim.x;
    unsigned gen = *reinterpret_cast<volatile unsigned*>( &S->gbar[1] ); // read before this block's first arrival

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__( 8 ) unsigned long long full_bar[NS];
    __shared__ dd_t s_red[C::NT / 32];
    __shared__ double s_bc[2];
    const uint32_t smem_base = ( smem_u32( smem_raw ) + 127u ) & ~127u;
    double* stage0 = reinterpret_cast<double*>( smem_raw + ( smem_base - smem_u32( smem_raw ) ) );
    double* pn0 = stage0 + NS * STAGED;
    if ( tid == 0 )
    {
        prefetch_tmap( &tmap_r );
        prefetch_tmap( &tmap_p0 );
        prefetch_tmap( &tmap_p1 );
#pragma unroll
        for ( int s = 0; s < NS; ++s )
            mbar_init( smem_u32( &full_bar[s] ), 1 );
        fence_barrier_init();
    }
    __syncthreads();

    // the CG scalars: the same in every block at every moment
    double rz_old = S->rz_old, pAp = S->pAp, rr = S->rr, rz_new = S->rz_new, alpha = S->alpha;
    int iter = S->iter;
    const int fixed = S->fixed;
    const double thresh = S->thresh;
    bool done = S->done != 0;
    int pc = a.pcur, lbase = 0;

    // sums of the block partials of values v0 (and v1 >= 0), the same doubles in every thread of every block; the
    // loads of both values are in flight together
    auto totals = [&]( int v0, int v1, double& out0, double& out1 ) {
        dd_t t0 = { 0.0, 0.0 }, t1 = { 0.0, 0.0 };
        for ( int b = tid; b < nb; b += C::NT )
        {
            dd_t w0, w1 = { 0.0, 0.0 };
            w0.hi = __ldcg( a.partials + ( (size_t)v0 * a.pstride + b ) * 2 + 0 );
            w0.lo = __ldcg( a.partials + ( (size_t)v0 * a.pstride + b ) * 2 + 1 );
            if ( v1 >= 0 )
            {
                w1.hi = __ldcg( a.partials + ( (size_t)v1 * a.pstride + b ) * 2 + 0 );
                w1.lo = __ldcg( a.partials + ( (size_t)v1 * a.pstride + b ) * 2 + 1 );
            }
            t0 = dd_add( t0, w0 );
            t1 = dd_add( t1, w1 );
        }
        t0 = dd_block_sum<C::NT>( t0, s_red );
        if ( v1 >= 0 )
            t1 = dd_block_sum<C::NT>( t1, s_red );
        if ( tid == 0 )
        {
            s_bc[0] = t0.hi + t0.lo;
            s_bc[1] = t1.hi + t1.lo;
        }
        __syncthreads();
        out0 = s_bc[0];
        out1 = s_bc[1];
        __syncthreads(); // s_bc is free again
    };
    auto publish = [&]( int value, dd_t v ) {
        v = dd_block_sum<C::NT>( v, s_red );
        if ( tid == 0 )
        {
            a.partials[( (size_t)value * a.pstride + bid ) * 2 + 0] = v.hi;
            a.partials[( (size_t)value * a.pstride + bid ) * 2 + 1] = v.lo;
        }
    };

    for ( int n = 0; n < a.niters && !done; ++n )
    {
        // ---- phase A (cg_rupdate_kernel): the tolerance met by the previous iteration ends the solve
        if ( iter > 0 && !fixed && sqrt( rr ) <= thresh )
        {
            done = true;
            break;
        }
        alpha = rz_old / pAp;
        dd_t arr = { 0.0, 0.0 }, arz = { 0.0, 0.0 };
        rupdate_rows( g, op, a.q, a.r, -alpha, a.txp_log2, bid, nb, arr, arz );
        publish( 0, arr );
        publish( 1, arz );
        fence_proxy_async_global();
        grid_barrier( S, (unsigned)nb, gen );
        totals( 0, 1, rr, rz_new );

        // ---- phase B (cg_fused_kernel)
        const double resid = sqrt( rr );
        const bool conv = !fixed && resid <= thresh;
        if ( bid == 0 && tid == 0 && iter < CFB_HIST_MAX )
            S->hist[iter] = resid;
        ++iter;
        const double* p_old = a.pbuf[pc];
        if ( conv )
        {
            // x += alpha p of the converged iteration; nothing else (the loop breaks here)
            const unsigned npx = (unsigned)( ( g.n[0] + 1 ) >> 1 );
            const unsigned tot = npx * (unsigned)g.n[1] * (unsigned)g.n[2];
            for ( unsigned t = (unsigned)bid * C::NT + tid; t < tot; t += (unsigned)nb * C::NT )
            {
                const unsigned row = t / npx;
                const int i = 2 * (int)( t - row * npx );
                const int k = (int)( row / (unsigned)g.n[1] );
                const int j = (int)( row - (unsigned)k * (unsigned)g.n[1] );
                const long long o = geo_off( g, i, j, k );
                if ( i + 1 < g.n[0] )
                {
                    const double2 pv = *reinterpret_cast<const double2*>( p_old + o );
                    double2 xv = *reinterpret_cast<double2*>( a.x + o );
                    xv.x = fma( alpha, pv.x, xv.x );
                    xv.y = fma( alpha, pv.y, xv.y );
                    *reinterpret_cast<double2*>( a.x + o ) = xv;
                }
                else
                    a.x[o] = fma( alpha, p_old[o], a.x[o] );
            }
            done = true;
            break;
        }
        const double beta = rz_new / rz_old;
        FusedArgs fa{};
        fa.x = a.x;
        fa.p = a.pbuf[pc ^ 1];
        fa.q = a.q;
        fa.p_old = p_old;
        fa.hx = a.hx;
        fa.zc = a.zc;
        fa.store_q = 1;
        dd_t acc = { 0.0, 0.0 };
        for ( int u = bid; u < a.units_total; u += nb )
        {
            const int tx = u % a.tiles_x, ty = ( u / a.tiles_x ) % a.tiles_y, ch = u / ( a.tiles_x * a.tiles_y );
            const int kbeg = ch * a.zc, kend = min( kbeg + a.zc, g.n[2] );
            dd_t au;
            int nloads;
            if constexpr ( FLAT )
            {
                nloads = min( a.yc, ( g.n[1] + TY - 1 ) / TY - ty * a.yc ); // tiles of the run
                au = fused_unit_flat<C, false>( tmap_r, pc ? tmap_p1 : tmap_p0, g, op, fa, tx * TX, ty * a.yc * TY, nloads, alpha, beta,
                                                stage0, pn0, full_bar, smem_base, lbase );
            }
            else
            {
                nloads = kend - kbeg + 2;
                au = fused_unit<C, false>( tmap_r, pc ? tmap_p1 : tmap_p0, g, op, fa, tx * TX, ty * TY, kbeg, kend, alpha, beta,
                                           stage0, pn0, full_bar, smem_base, lbase );
            }
            acc = dd_add( acc, au );
            // every load index of the unit was issued and consumed
            lbase = ( lbase + nloads ) % ( 2 * NS );
        }
        publish( 2, acc );
        fence_proxy_async_global();
        grid_barrier( S, (unsigned)nb, gen );
        {
            double unused;
            totals( 2, -1, pAp, unused );
        }
        rz_old = rz_new;
        pc ^= 1;
    }
    if ( bid == 0 && tid == 0 )
    {
        S->rz_old = rz_old;
        S->pAp = pAp;
        S->rr = rr;
        S->rz_new = rz_new;
        S->alpha = alpha;
        S->iter = iter;
        if ( done )
            S->done = 1;
    }
}

typedef CUresult ( *PFN_encodeTiled )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill );

template <class C>
int launch_fused_cfg( cfb_ctx* c, const FusedArgs& a, int grid, const PeerFusedArgs* pf, cudaStream_t st )
{
    static bool attr_set = false;
    if ( !attr_set )
    {
        cudaFuncSetAttribute( cg_fused_kernel<C, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
        cudaFuncSetAttribute( cg_fused_kernel<C, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
        cudaFuncSetAttribute( cg_fused_kernel<C, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
        cudaFuncSetAttribute( cg_fused_kernel<C, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
        cudaFuncSetAttribute( cg_fused_kernel<C, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
        attr_set = true;
    }
    const bool xs = a.gxr[0] || a.gxr[1];
    const bool flat = c->g.D == 2 && c->flat_2d; // one owned plane between two zero ghost planes
    const NoPeerArgs none{};
    if ( pf ) // (peer_overlapped() has made sure that neither xs nor flat applies)
        cg_fused_kernel<C, false, false, true><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur],
                                                                                          c->g, c->op, a, *pf );
    else if ( xs && flat )
        cg_fused_kernel<C, true, true, false><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur],
                                                                                         c->g, c->op, a, none );
    else if ( xs )
        cg_fused_kernel<C, true, false, false><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur],
                                                                                          c->g, c->op, a, none );
    else if ( flat )
        cg_fused_kernel<C, false, true, false><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur],
                                                                                          c->g, c->op, a, none );
    else
        cg_fused_kernel<C, false, false, false><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur],
                                                                                           c->g, c->op, a, none );
    return 1;
}

int dispatch_fused( cfb_ctx* c, const FusedArgs& a, int grid, const PeerFusedArgs* pf = nullptr, cudaStream_t st = nullptr )
{
    if ( !st )
        st = c->stream;
    const int key = c->fu_tx * 10000 + c->fu_ty * 100 + c->fu_stages;
    // "fused_nt": sixteen warps instead of eight on the one CTA an SM holds of the 128 x 16 x 3 tiling (two rows per
    // thread instead of four).  Measured at 512^3 (profiles/r2_sweep_phase_b.log): without the q store (64-byte form)
    // 871 - 880 us against 904 - 917, with it (72-byte form) 1122 - 1125 against 1083 - 1084 — so the automatic
    // choice (0) takes it in the 64-byte form only.  Plain instantiation only: the 64-byte form has no x-staging
    // reads and no in-kernel mailbox reduction, and two-dimensional runs use the 72-byte form.
    const bool plain = !pf && !( a.gxr[0] || a.gxr[1] ) && !( c->g.D == 2 && c->flat_2d );
    if ( c->fu_nt == 512 && !( key == 1281603 && plain ) )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "fused_nt 512 exists for the plain 128 x 16 x 3 tiling only" ) );
        return 0;
    }
    if ( key == 1281603 && plain && ( c->fu_nt == 512 || ( c->fu_nt == 0 && !a.store_q ) ) )
    {
        using C = FusedCfg<128, 16, 3, 512>;
        static bool attr_set = false;
        if ( !attr_set )
        {
            cudaFuncSetAttribute( cg_fused_kernel<C, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES );
            attr_set = true;
        }
        const NoPeerArgs none{};
        cg_fused_kernel<C, false, false, false><<<grid, C::NT, C::SMEM_BYTES, st>>>( c->tmap_fr, c->tmap_fp[c->pcur], c->g, c->op, a, none );
        return 1;
    }
    switch ( key )
    {
    case 641603:
        return launch_fused_cfg<FusedCfg<64, 16, 3>>( c, a, grid, pf, st );
    case 641604:
        return launch_fused_cfg<FusedCfg<64, 16, 4>>( c, a, grid, pf, st );
    case 640803:
        return launch_fused_cfg<FusedCfg<64, 8, 3>>( c, a, grid, pf, st );
    case 640804:
        return launch_fused_cfg<FusedCfg<64, 8, 4>>( c, a, grid, pf, st );
    case 643202:
        return launch_fused_cfg<FusedCfg<64, 32, 2>>( c, a, grid, pf, st );
    case 643203:
        return launch_fused_cfg<FusedCfg<64, 32, 3>>( c, a, grid, pf, st );
    case 1280803:
        return launch_fused_cfg<FusedCfg<128, 8, 3>>( c, a, grid, pf, st );
    case 1280804:
        return launch_fused_cfg<FusedCfg<128, 8, 4>>( c, a, grid, pf, st );
    case 1281603:
        return launch_fused_cfg<FusedCfg<128, 16, 3>>( c, a, grid, pf, st );
    default:
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "unsupported fused tile configuration" ) );
        return 0;
    }
}

} // namespace

// Tiling of phase B for the current block: tiles in x/y, z chunk, number of units.
// Two-dimensional runs with the FLAT kernels: a unit is a run of `yc` tile rows (fused_unit_flat) and `tiles_y`
// counts runs; everywhere else yc = 1.
void fused_tiling( const cfb_ctx* c, int& tiles_x, int& tiles_y, int& zc, int& chunks, int& yc )
{
    const Geo& g = c->g;
    tiles_x = ( g.n[0] + c->fu_tx - 1 ) / c->fu_tx;
    tiles_y = ( g.n[1] + c->fu_ty - 1 ) / c->fu_ty;
    yc = ( g.D == 2 && c->flat_2d ) ? std::max( 1, std::min( c->fu_yc, tiles_y ) ) : 1;
    tiles_y = ( tiles_y + yc - 1 ) / yc;
    zc = c->fu_zc > 0 ? c->fu_zc : g.n[2];
    chunks = ( g.n[2] + zc - 1 ) / zc;
}

// (Re)build the tensor maps of cg_r and cg_p for the current fused tile shape.
int fused_setup( cfb_ctx* c )
{
    static PFN_encodeTiled encode = nullptr;
    if ( !encode )
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres );
        if ( e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn )
            return cfb_fail( c, CFB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available" );
        encode = (PFN_encodeTiled)fn;
    }
    const Geo& g = c->g;
    if ( !c->fu_auto && ( c->fu_tx < 2 || c->fu_ty < 1 || c->fu_zc < 0 ) )
        return cfb_fail( c, CFB_ERR_INVALID, "fused tile extents must be positive" );
    if ( c->fu_auto )
    {
        // Rules distilled from the sweeps in profiles/r1_sweep_fused*.log (64^3 ... 512^3):
        //  * 128x16 tiles (1 CTA/SM, 3 stages) win from 512 cells along x; below that 64x16 (2 CTAs/SM);
        //  * 64-plane chunks, halved until there is at least one unit per SM (every chunk re-reads
        //    two planes, so chunks stay as long as the SM count allows);
        //  * blocks too small even for that take 64x8 tiles (3 CTAs/SM, 4 stages).
        auto units_of = [&]( int tx, int ty, int zc ) {
            return (long long)( ( g.n[0] + tx - 1 ) / tx ) * ( ( g.n[1] + ty - 1 ) / ty ) * ( ( g.n[2] + zc - 1 ) / zc );
        };
        int tx = ( g.n[0] >= 512 && g.n[1] >= 64 ) ? 128 : 64, ty = 16, st = 3, zc = 64;
        while ( units_of( tx, ty, zc ) < c->sm_count && zc > 4 )
            zc /= 2;
        if ( units_of( tx, ty, zc ) < c->sm_count )
        {
            tx = 64;
            ty = 8;
            st = 4;
            zc = 64;
            while ( units_of( tx, ty, zc ) < c->sm_count && zc > 4 )
                zc /= 2;
        }
        // very many units (large blocks): longer chunks re-read fewer planes; bounded — a large cross-section
        // alone can exceed any unit budget (two-dimensional grids), the partial-sum scratch follows the unit count
        while ( units_of( tx, ty, zc ) > CFB_MAX_PARTIALS && zc < g.n[2] )
            zc *= 2;
        c->fu_tx = tx;
        c->fu_ty = ty;
        c->fu_stages = st;
        c->fu_zc = zc;
    }
    if ( c->fu_yc_auto )
    {
        // Two-dimensional runs: the run length that needs the fewest rounds of resident blocks, a run costing its tiles
        // plus about two for filling the ring.  Matches the measured best at every size (profiles/r2_sweep_2d_march.log,
        // 128 x 16 x 3 tiles, one block per SM): 8192^2 -> 32 (532 us; 16: 534, 8: 552, 64: 566), 4096^2 -> 64 (143;
        // 32: 147, 8: 153), 2048^2 -> 16 (47; 8: 49, 32: 65), 1024^2 -> 4 (20; 8: 26, 1: 28).
        const long long tiles_x = ( g.n[0] + c->fu_tx - 1 ) / c->fu_tx, tiles_y = ( g.n[1] + c->fu_ty - 1 ) / c->fu_ty;
        const long long box = ( (long long)( c->fu_tx + 4 ) * ( c->fu_ty + 2 ) * 8 + 127 ) / 128 * 128;
        const long long smem = ( 2LL * c->fu_stages + 3 ) * box + 128; // FusedCfg::SMEM_BYTES
        const long long slots = (long long)c->sm_count * ( smem <= 56 * 1024 ? 4 : ( smem <= 75 * 1024 ? 3 : ( smem <= 113 * 1024 ? 2 : 1 ) ) );
        long long best = -1;
        for ( int yc = 64; yc >= 1; yc /= 2 )
        {
            const long long units = tiles_x * ( ( tiles_y + yc - 1 ) / yc );
            const long long cost = ( ( units + slots - 1 ) / slots ) * ( std::min<long long>( yc, tiles_y ) + 2 );
            if ( best < 0 || cost < best )
            {
                best = cost;
                c->fu_yc = yc;
            }
        }
    }
    cuuint64_t gdim[3] = { (cuuint64_t)g.sy, (cuuint64_t)g.ay, (cuuint64_t)g.az };
    cuuint64_t gstride[2] = { (cuuint64_t)g.sy * 8, (cuuint64_t)g.sz * 8 };
    cuuint32_t box[3] = { (cuuint32_t)( c->fu_tx + 4 ), (cuuint32_t)( c->fu_ty + 2 ), 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUtensorMap* maps[3] = { &c->tmap_fr, &c->tmap_fp[0], &c->tmap_fp[1] };
    double* base[3] = { c->cg_r, c->cg_pbuf[0], c->cg_pbuf[1] };
    for ( int m = 0; m < 3; ++m )
    {
        CUresult r = encode( maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base[m], gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
        if ( r != CUDA_SUCCESS )
            return cfb_fail( c, CFB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string( (int)r ) );
    }
    // unit list: units whose tile touches a face with a neighbour rank (they read exchanged ghosts)
    // go last, so that everything before them can run while the ghosts are in flight
    int tiles_x, tiles_y, zc, chunks, yc;
    fused_tiling( c, tiles_x, tiles_y, zc, chunks, yc );
    {
        const int rc = ensure_partials( c, (long long)tiles_x * tiles_y * chunks );
        if ( rc )
            return rc;
    }
    std::vector<int> inner, outer;
    for ( int ch = 0; ch < chunks; ++ch )
        for ( int ty = 0; ty < tiles_y; ++ty )
            for ( int tx = 0; tx < tiles_x; ++tx )
            {
                const bool bd = ( tx == 0 && c->nbr[0] >= 0 ) || ( tx == tiles_x - 1 && c->nbr[1] >= 0 ) ||
                                ( ty == 0 && c->nbr[2] >= 0 ) || ( ty == tiles_y - 1 && c->nbr[3] >= 0 ) ||
                                ( ch == 0 && c->nbr[4] >= 0 ) || ( ch == chunks - 1 && c->nbr[5] >= 0 );
                std::vector<int>& v = bd ? outer : inner;
                v.push_back( tx );
                v.push_back( ty );
                v.push_back( ch );
            }
    c->n_interior = (int)inner.size() / 3;
    inner.insert( inner.end(), outer.begin(), outer.end() );
    c->n_units = (int)inner.size() / 3;
    if ( c->d_units )
        cudaFree( c->d_units );
    c->d_units = nullptr;
    if ( cudaMalloc( &c->d_units, inner.size() * sizeof( int ) ) != cudaSuccess ||
         cudaMemcpy( c->d_units, inner.data(), inner.size() * sizeof( int ), cudaMemcpyHostToDevice ) != cudaSuccess )
        return cfb_fail( c, CFB_ERR_CUDA, "fused unit list upload failed" );
    c->fused_ok = true;
    return CFB_OK;
}

static int launch_rupdate_impl( cfb_ctx* c, const PeerFusedArgs* pf )
{
    const Geo& g = c->g;
    const int npx = ( g.n[0] + 1 ) / 2;
    int txp_log2 = 5;
    while ( ( 1 << txp_log2 ) < npx && txp_log2 < 8 )
        ++txp_log2;
    const int batch = ( NT >> txp_log2 ) * RU;
    const long long rows = (long long)g.n[1] * g.n[2];
    long long grid = ( rows + batch - 1 ) / batch;
    const long long cap = std::min<long long>( (long long)c->sm_count * c->ru_ctas, CFB_MAX_PARTIALS );
    if ( grid > cap )
        grid = cap;
    if ( pf )
        cg_rupdate_kernel<true><<<(int)grid, NT, 0, c->stream>>>( g, c->op, c->cg_q, c->cg_r, c->d_state, c->d_partials,
                                                                 txp_log2, *pf );
    else
        cg_rupdate_kernel<false><<<(int)grid, NT, 0, c->stream>>>( g, c->op, c->cg_q, c->cg_r, c->d_state,
                                                                  c->d_partials, txp_log2, NoPeerArgs{} );
    return 1;
}

int launch_cg_rupdate( cfb_ctx* c ) { return launch_rupdate_impl( c, nullptr ); }

// Phase A whose last block runs the mailbox reduction of (r.z, r.r); no faces ("peer_overlap")
int launch_cg_rupdate_mail( cfb_ctx* c )
{
    PeerFusedArgs pf{};
    peer_mail_only( c, pf );
    return launch_rupdate_impl( c, &pf );
}

int launch_cg_finish( cfb_ctx* c )
{
    cg_finish_kernel<<<1, 1, 0, c->stream>>>( c->d_state );
    return 1;
}

// ---- persistent form (small blocks, one GPU) ---------------------------------------------------
namespace
{
template <class C, bool FLAT>
int launch_persist_cfg( cfb_ctx* c, PersistArgs& a )
{
    auto* fn = &cg_persistent_kernel<C, FLAT>;
    static int occ = -1;
    if ( occ < 0 )
    {
        if ( cudaFuncSetAttribute( fn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES ) != cudaSuccess ||
             cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, fn, C::NT, C::SMEM_BYTES ) != cudaSuccess || occ < 1 )
        {
            occ = -1;
            cudaGetLastError();
            note_rc( c, cfb_fail( c, CFB_ERR_CUDA, "persistent CG kernel: no resident block fits an SM" ) );
            return 0;
        }
    }
    // every block resident at once (the grid barrier needs it; the cooperative launch checks it)
    const int grid = std::min( occ * c->sm_count, CFB_MAX_PARTIALS );
    void* args[] = { (void*)&c->tmap_fr, (void*)&c->tmap_fp[0], (void*)&c->tmap_fp[1], (void*)&c->g, (void*)&c->op, (void*)&a };
    const cudaError_t e = cudaLaunchCooperativeKernel( fn, dim3( grid ), dim3( C::NT ), args, C::SMEM_BYTES, c->stream );
    if ( e != cudaSuccess )
    {
        cudaGetLastError();
        note_rc( c, cfb_fail( c, CFB_ERR_CUDA, std::string( "persistent CG kernel: " ) + cudaGetErrorString( e ) ) );
        return 0;
    }
    return 1;
}
} // namespace

// the tile shapes the persistent kernel is instantiated for (those fused_setup picks for small blocks)
bool cg_persist_supported( const cfb_ctx* c )
{
    const int key = c->fu_tx * 10000 + c->fu_ty * 100 + c->fu_stages;
    return key == 641603 || key == 640804 || key == 640803 || key == 1281603;
}

// `iters` iterations of the two-kernel form in one cooperative launch (the caller has checked cg_persist_applies)
int launch_cg_persistent( cfb_ctx* c, int iters )
{
    const Geo& g = c->g;
    PersistArgs a{};
    a.x = c->lhs;
    a.r = c->cg_r;
    a.q = c->cg_q;
    a.pbuf[0] = c->cg_pbuf[0];
    a.pbuf[1] = c->cg_pbuf[1];
    a.pcur = c->pcur;
    a.S = c->d_state;
    a.partials = c->d_partials;
    a.pstride = c->partials_cap;
    a.hx = 16;
    int chunks;
    fused_tiling( c, a.tiles_x, a.tiles_y, a.zc, chunks, a.yc );
    a.units_total = a.tiles_x * a.tiles_y * chunks;
    const int npx = ( g.n[0] + 1 ) / 2;
    a.txp_log2 = 5;
    while ( ( 1 << a.txp_log2 ) < npx && a.txp_log2 < 8 )
        ++a.txp_log2;
    a.niters = iters;
    const bool flat = g.D == 2 && c->flat_2d;
    const int key = c->fu_tx * 10000 + c->fu_ty * 100 + c->fu_stages;
    switch ( key )
    {
    case 641603:
        return flat ? launch_persist_cfg<FusedCfg<64, 16, 3>, true>( c, a ) : launch_persist_cfg<FusedCfg<64, 16, 3>, false>( c, a );
    case 640804:
        return flat ? launch_persist_cfg<FusedCfg<64, 8, 4>, true>( c, a ) : launch_persist_cfg<FusedCfg<64, 8, 4>, false>( c, a );
    case 640803:
        return flat ? launch_persist_cfg<FusedCfg<64, 8, 3>, true>( c, a ) : launch_persist_cfg<FusedCfg<64, 8, 3>, false>( c, a );
    case 1281603:
        return flat ? launch_persist_cfg<FusedCfg<128, 16, 3>, true>( c, a ) : launch_persist_cfg<FusedCfg<128, 16, 3>, false>( c, a );
    default:
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "persistent CG kernel: tile configuration not instantiated" ) );
        return 0;
    }
}

// phase B over all units (which = 0), or over the device unit list `c->d_units` split into
// interior units [0, n_interior) (which = 1) and boundary units [n_interior, n_units) (which = 2).
static int launch_cg_fused_impl( cfb_ctx* c, int which, const PeerFusedArgs* pf, cudaStream_t st = nullptr )
{
    FusedArgs a{};
    a.x = c->lhs;
    a.p_old = c->cg_pbuf[c->pcur];
    a.p = c->cg_pbuf[c->pcur ^ 1];
    a.q = c->cg_q;
    a.S = c->d_state;
    a.partials = c->d_partials;
    a.hx = 16;
    int chunks;
    fused_tiling( c, a.tiles_x, a.tiles_y, a.zc, chunks, a.yc );
    const int total = a.tiles_x * a.tiles_y * chunks;
    a.units_total = total;
    // optional top-down walk (phase A sweeps bottom-up and leaves the top of r in the L2); measured
    // neutral from 64^3 to 512^3 (profiles/r1_sweep_fused2.log), so it is off by default
    a.reverse = c->fu_reverse ? 1 : 0;
    a.store_q = c->cg_variant == 2 ? 0 : 1;
    if ( peer_xstaged( c ) )
        for ( int side = 0; side < 2; ++side )
            if ( c->nbr[side] >= 0 )
            {
                const size_t slot = (size_t)c->g.n[1] * c->g.n[2];
                a.gxr[side] = c->xstage_self + ( side * 3 + 0 ) * slot;
                a.gxp[side] = c->xstage_self + ( side * 3 + 1 + c->pcur ) * slot;
            }
    int grid = total;
    if ( which != 0 )
    {
        if ( !c->d_units || c->n_units != total )
        {
            note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "fused unit list not built" ) );
            return 0;
        }
        a.unit_base = which == 1 ? 0 : c->n_interior;
        grid = which == 1 ? c->n_interior : total - c->n_interior;
        a.units = c->d_units + 3 * a.unit_base;
        if ( grid == 0 )
            return 0;
    }
    return dispatch_fused( c, a, grid, pf, st );
}

int launch_cg_fused( cfb_ctx* c, int which ) { return launch_cg_fused_impl( c, which, nullptr ); }

// Phase B units whose last block (the one that draws the last of ALL the phase's tickets, whichever launch it
// belongs to) runs the mailbox reduction of p.Ap ("peer_overlap").  which = 1: interior units on the main stream,
// 2: boundary units — `side`: on the side stream, behind the face transfers that deliver their ghosts; the two
// launches run concurrently and share the ticket, so splitting the phase costs no extra wave of blocks.
int launch_cg_fused_mail( cfb_ctx* c, int which, bool side )
{
    PeerFusedArgs pf{};
    peer_mail_only( c, pf );
    return launch_cg_fused_impl( c, which, &pf, side ? c->comm_stream : c->stream );
}
