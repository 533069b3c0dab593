// kernels_stencil.cu — q = A p fused with sum(p.q): CG kernel 4, the dominant kernel of the path.
//
// A is the reference's pressure matrix (src/VelocityCorrector.hpp:116-144 with the boundary fix-up
// of src/BoundaryConditions.hpp:56-97) applied matrix-free: no coefficient array is stored; the
// diagonal follows from the number of SOLID walls a cell touches, off-domain neighbours are ghost
// zeros (physical walls) or the neighbour rank's cells (block boundaries).
//
// Design (sm_100a, HBM-bound: 16 algorithmic bytes per cell = read p, write q):
//   * one CTA owns a TX x TY tile of the x-y plane and marches ZC planes along z (2.5-D blocking):
//     k-1 / k / k+1 centre values live in registers, the x/y neighbours of plane k come from
//     shared memory;
//   * planes are staged into a NS-deep shared-memory ring by TMA (cp.async.bulk.tensor.3d, box
//     (TX+4) x (TY+2) x 1 with mbarrier complete_tx), issued NS-1 planes ahead by one thread, so
//     the load path costs no registers and no per-thread address arithmetic; out-of-range box
//     parts are zero-filled by the TMA unit;
//   * every thread owns column pairs: 128-bit LDS for the centre / y-neighbours, 128-bit
//     coalesced STG of q (a warp stores 512 contiguous bytes);
//   * p.q is accumulated per thread over its z-march, reduced with warp shuffles + one smem pass
//     per CTA, and finalised deterministically by the last CTA (device_reduce.cuh).
//   * a second variant (LDG, no shared memory) is kept for A/B measurements via cfb_set_tuning.
//   * phase A' of the 64-byte CG form is this kernel in MODE 1: every ring slot then also carries the tile of r.
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"
#include "device_tma.cuh"
#include "device_peer.cuh"
#include "device_cg1.cuh"
#include <algorithm>

namespace
{

// End of kernel 4: publish p.Ap (see publish_rr_rz in kernels_cg.cu for the multi-GPU path) and
// perform the CG loop's "zTr_old = zTr_new".
__device__ __forceinline__ void publish_pAp( CgState* S, dd_t v )
{
    if ( S->world > 1 )
    {
        S->loc[0] = v.hi;
        S->loc[1] = v.lo;
    }
    else
        S->pAp = v.hi + v.lo;
    S->rz_old = S->rz_new;
}

template <int TX_, int TY_, int NS_>
struct TileCfg
{
    static constexpr int TX = TX_, TY = TY_, NS = NS_;
    static constexpr int NT = 256;
    static constexpr int LX = TX / 2;  // threads along x (one column pair each)
    static constexpr int WY = NT / LX; // thread rows
    static constexpr int RY = TY / WY; // rows per thread
    static constexpr int PX = TX + 4;  // smem row pitch in doubles: 2-wide x halo keeps pairs 16-B aligned
    static constexpr int PY = TY + 2;
    static constexpr int BOX_BYTES = PX * PY * 8;
    static constexpr int STAGE_BYTES = ( BOX_BYTES + 127 ) / 128 * 128;
    static constexpr int SMEM_BYTES = NS * STAGE_BYTES + 128;
    // phase A' with r staged by TMA as well (RT): every stage carries the TX x TY tile of r behind the box of p
    static constexpr int R_BYTES = TX * TY * 8;
    static constexpr int STAGE_RT_BYTES = STAGE_BYTES + R_BYTES;
    static constexpr int SMEM_RT_BYTES = NS * STAGE_RT_BYTES + 128;
    static_assert( R_BYTES % 128 == 0, "r tile must keep the stages 128-byte aligned" );
    // MODE 5 (multigrid: prolongation + first post-sweep): behind them the box of the coarse correction whose cells are
    // the parents of the tile and of its halo ring, (TX / 2 + 2) x (TY / 2 + 2) entries of one coarse plane
    static constexpr int EX = TX / 2 + 2, EY = TY / 2 + 2;
    static constexpr int E_BYTES = EX * EY * 8;
    static constexpr int E_PAD = ( E_BYTES + 127 ) / 128 * 128;
    static constexpr int STAGE_E_BYTES = STAGE_RT_BYTES + E_PAD;
    static constexpr int SMEM_E_BYTES = NS * STAGE_E_BYTES + 128;
    static_assert( E_BYTES % 16 == 0, "coarse box rows must be multiples of 16 bytes" );
    static_assert( TX % 2 == 0 && NT % LX == 0 && TY % WY == 0 && RY >= 1, "bad tile" );
};

struct StencilArgs
{
    double* r; // MODE 1 only
    double* q;
    CgState* S;
    double* partials;
    int tiles_x, tiles_y, zc, hx;
    int pstride; // entries per value in the partial-sum scratch (cfb_ctx::partials_cap)
    int init;    // MODE 2: the launch that starts a solve (sets alpha, beta; counts no iteration)
    // MODE 3 / 4 (fine-level smoothing sweeps of the multigrid preconditioner, mg.cu): damping of the sweep(s); the
    // result goes to `q`; dot != 0: the sweep also leaves sum z.b in S->rz_new (mg_smooth_dot_kernel's job)
    double om1, om2;
    int dot;
    int ehz; // MODE 5: ghost width of the coarse array (array plane of coarse plane K = K + ehz; likewise rows, columns)
};

// MODE 0: q = A p, sum p.q (CG kernel 4).
// MODE 1: phase A' of the 64-byte iteration (cg_variant 2): the same march over p, but q = A p is consumed
//         on the spot instead of being stored: alpha = zr_old / pAp ; r -= alpha (A p) ; sum r^2 ; sum r.M^-1 r
//         (reads p and r through TMA, writes r: 24 B/cell, and phase B no longer writes q).
//         Statement for statement cg_rupdate_kernel with q recomputed by the same row expression that
//         produced it, hence bit-identical.
// MODE 2: the stencil kernel of the single-reduction CG (cg_variant 3, kernels_cg1.cu): the planes staged by TMA are
//         those of r; u = M^-1 r is formed on the fly for the centre and its six neighbours (the diagonal is a
//         function of the cell's wall count, so nothing is stored), w = A u is written, and the THREE sums of the
//         iteration's only reduction point are taken on the march: sum r^2, sum r.u, sum w.u (16 B/cell).
// MODE 3: one damped-Jacobi sweep of the multigrid preconditioner on the fine level (mg.cu: cell_smooth): the planes
//         staged by TMA are those of the iterate xi, every slot also carries the tile of the right-hand side b (as in
//         MODE 1); xo = xi + (omega D^-1)(b - A xi) is written to `q` (24 B/cell), optionally with sum xo.b.
// MODE 4: the first TWO sweeps from a zero initial guess in one pass (mg.cu: cell_smooth02): the planes are those of b;
//         x1 = (omega1 D^-1) b for the centre and its six neighbours (each with its own wall count, as MODE 2 forms
//         M^-1 r), xo = x1 + (omega2 D^-1)(b - A x1) (16 B/cell).  Statement for statement the one-thread-per-cell
//         kernels they replace on the fine level — which reached 42 - 50 % of the bandwidth this march reaches —
//         hence bit-identical.
// MODE 5: prolongation + correction + the first post-smoothing sweep in one pass (mg.cu: cell_prolong_smooth): MODE 3
//         on x' = xi + P e, the piecewise-constant prolongation of the coarse correction e added on the fly to the cell
//         and to its six neighbours; every ring slot also carries the box of e with the parents of the plane's tile and
//         halo ring (third tensor map, over the coarse level's own array).
// FLAT: two-dimensional runs (one owned plane between two zero ghost planes): the z neighbours are zero by
// construction and their planes are not loaded.  A template flag: the 3-D instantiations are untouched.
// PF (MODE 2): the block that draws the last ticket runs the mailbox reduction of the kernel's sums over NVLink
// peer memory (device_peer.cuh).
// MODE 1 stages the TX x TY tile of r of every plane through the TMA ring too (second tensor map, same mbarrier as
// the plane of p it is consumed with): the loads then run NS - 1 planes ahead whatever the number of resident warps.
// Round 2 streamed r with 128-bit loads one plane ahead: long_scoreboard 4.3 stalled warps per issue, 629 us = 77 %
// of the measured bandwidth at 512^3 against 528 - 540 us = 91 - 93 % now (profiles/r2_sweep_rtma.log, measured
// side by side through the "stencil_rtma" key, which went with the loser).
template <class C, int MODE>
constexpr int stencil_smem_bytes()
{
    return MODE == 5 ? C::SMEM_E_BYTES : ( ( MODE == 1 || MODE == 3 ) ? C::SMEM_RT_BYTES : C::SMEM_BYTES );
}
template <class C, int MODE>
constexpr int stencil_min_ctas()
{
    // 228 KB of shared memory per SM, 1 KB reserved per resident CTA
    constexpr int B = stencil_smem_bytes<C, MODE>();
    return ( MODE == 1 || MODE == 3 || MODE == 5 ) ? ( B <= 75 * 1024 ? 3 : ( B <= 113 * 1024 ? 2 : 1 ) )
                                                   : ( B <= 56 * 1024 ? 3 : ( B <= 110 * 1024 ? 2 : 1 ) );
}
template <class C, int MODE, bool FLAT, bool PF>
__global__ void __launch_bounds__( C::NT, stencil_min_ctas<C, MODE>() )
    stencil7_dot_tma( const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_r,
                      const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                      const __grid_constant__ StencilArgs a, const __grid_constant__ typename PeerSel<PF>::type pf )
{
    static_assert( !PF || MODE == 2, "only the single-reduction form reduces over the mailboxes from this kernel" );
    constexpr bool RT = MODE == 1 || MODE == 3 || MODE == 5; // the slots also carry a halo-free tile (of r / of b)
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS, PX = C::PX, RY = C::RY, WY = C::WY, LX = C::LX;
    constexpr int STAGE = MODE == 5 ? C::STAGE_E_BYTES : ( RT ? C::STAGE_RT_BYTES : C::STAGE_BYTES ); // bytes per ring slot
    double nalpha = 0.0;
    if ( MODE == 0 || MODE == 2 )
    {
        if ( a.S->done )
            return;
    }
    else if ( MODE == 1 )
    {
        // as in cg_rupdate_kernel: `done` is only ever written here, by cg_check0 and by cg_finish
        CgState* S = a.S;
        if ( S->done || ( S->iter > 0 && !S->fixed && sqrt( S->rr ) <= S->thresh ) )
        {
            if ( blockIdx.x == 0 && threadIdx.x == 0 )
                S->done = 1;
            return;
        }
        const double alpha = S->rz_old / S->pAp;
        nalpha = -alpha;
        if ( blockIdx.x == 0 && threadIdx.x == 0 )
            S->alpha = alpha;
    }

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__( 8 ) unsigned long long full_bar[NS];

    const uint32_t smem_base = ( smem_u32( smem_raw ) + 127u ) & ~127u;
    const double* stage0 =
        reinterpret_cast<const double*>( smem_raw + ( smem_base - smem_u32( smem_raw ) ) );

    const int tid = threadIdx.x;
    const int lx = tid % LX, wy = tid / LX;

    // unit -> (tile_x, tile_y, z chunk)
    const int u = blockIdx.x;
    const int tx = u % a.tiles_x;
    const int ty = ( u / a.tiles_x ) % a.tiles_y;
    const int ch = u / ( a.tiles_x * a.tiles_y );
    const int x0 = tx * TX, y0 = ty * TY;
    const int kbeg = ch * a.zc;
    const int kend = min( kbeg + a.zc, g.n[2] );
    const int nplanes = kend - kbeg; // planes to compute
    const int nloads = nplanes + 2;  // planes kbeg-1 .. kend

    // TMA box origin (array coordinates): 2 columns left of the tile, 1 row below, plane kbeg-1
    const int cx = a.hx + x0 - 2;
    const int cy = g.h + y0 - 1;
    const int cz = g.h + kbeg - 1;

    // load l = plane kbeg - 1 + l of p into slot l % NS; RT: with it the tile of r of the same plane, which is
    // consumed in the same iteration (planes kbeg .. kend - 1 only: l = 1 .. nplanes)
    auto issue = [&]( int l ) {
        const int s = l % NS;
        const uint32_t bar = smem_u32( &full_bar[s] );
        const bool with_r = RT && l >= 1 && l <= nplanes;
        mbar_expect_tx( bar, C::BOX_BYTES + ( with_r ? C::R_BYTES : 0 ) + ( MODE == 5 ? C::E_BYTES : 0 ) );
        tma_load_3d( smem_base + s * STAGE, &tmap, bar, cx, cy, cz + l );
        if ( with_r )
            tma_load_3d( smem_base + s * STAGE + C::STAGE_BYTES, &tmap_r, bar, a.hx + x0, g.h + y0, cz + l );
        if ( MODE == 5 ) // the coarse plane of the parents of plane kbeg - 1 + l (floor division: ghost -1 -> ghost -1)
            tma_load_3d( smem_base + s * STAGE + C::STAGE_RT_BYTES, &tmap_e, bar, x0 / 2 - 1 + a.ehz, y0 / 2 - 1 + a.ehz,
                         ( ( kbeg - 1 + l + 2 ) >> 1 ) - 1 + a.ehz );
    };
    if ( tid == 0 )
    {
        prefetch_tmap( &tmap );
        if ( RT )
            prefetch_tmap( &tmap_r );
        if ( MODE == 5 )
            prefetch_tmap( &tmap_e );
#pragma unroll
        for ( int s = 0; s < NS; ++s )
            mbar_init( smem_u32( &full_bar[s] ), 1 );
        fence_barrier_init();
    }
    __syncthreads();
    if ( tid == 0 )
    {
        const int n0 = nloads < NS ? nloads : NS;
        for ( int l = FLAT ? 1 : 0; l < ( FLAT ? 2 : n0 ); ++l )
            issue( l );
    }

    // per-thread constants: validity and SOLID-wall counts of my cells
    const int i0 = x0 + 2 * lx; // owned x index of the first cell of my pair
    const bool vx0 = i0 < g.n[0], vx1 = i0 + 1 < g.n[0];
    const int wx0 = wall_count( g, 0, i0 + g.off[0] );
    const int wx1 = wall_count( g, 0, i0 + 1 + g.off[0] );
    int wyc[RY];
    bool vy[RY];
#pragma unroll
    for ( int r = 0; r < RY; ++r )
    {
        const int j = y0 + wy + r * WY;
        vy[r] = j < g.n[1];
        wyc[r] = wall_count( g, 1, j + g.off[1] );
    }
    const double ns = op.neg_scale;
    // MODE 2: wall counts of the x and y neighbours (their M^-1 differs from the centre's next to a wall)
    const int wxm = wall_count( g, 0, i0 - 1 + g.off[0] ), wxp = wall_count( g, 0, i0 + 2 + g.off[0] );
    int wym[RY], wyp[RY];
#pragma unroll
    for ( int r = 0; r < RY; ++r )
    {
        const int j = y0 + wy + r * WY;
        wym[r] = wall_count( g, 1, j - 1 + g.off[1] );
        wyp[r] = wall_count( g, 1, j + 1 + g.off[1] );
    }

    // MODE 5: rows of the coarse box holding the parents of my rows and of their y neighbours (box row 0 = coarse row
    // y0 / 2 - 1); my pair's parent sits in box column lx + 1, its x neighbours' parents in columns lx and lx + 2
    int jc[RY], jm[RY], jp[RY];
#pragma unroll
    for ( int r = 0; r < RY; ++r )
    {
        const int t = wy + r * WY; // row of the tile
        jc[r] = ( t >> 1 ) + 1;
        jm[r] = ( t + 1 ) >> 1;
        jp[r] = ( ( t + 1 ) >> 1 ) + 1;
    }
    auto ebox = [&]( int slot ) { return stage0 + slot * ( STAGE / 8 ) + C::STAGE_RT_BYTES / 8; };
    // MODE 2: M^-1 of a cell with `idx` walls; MODE 4: the first sweep's factor omega1 D^-1 (the product first, as in
    // mg.cu's cell_smooth02)
    auto mfac = [&]( int idx ) { return MODE == 4 ? a.om1 * op.minv[idx] : op.minv[idx]; };
    double2 zm[RY], cc[RY];
    if ( !FLAT )
        mbar_wait( smem_u32( &full_bar[0] ), 0 );
    mbar_wait( smem_u32( &full_bar[1 % NS] ), ( 1 / NS ) & 1 );
#pragma unroll
    for ( int r = 0; r < RY; ++r )
    {
        const int row = wy + r * WY + 1;
        zm[r] = FLAT ? make_double2( 0.0, 0.0 ) : *reinterpret_cast<const double2*>( stage0 + row * PX + 2 * lx + 2 );
        cc[r] = *reinterpret_cast<const double2*>( stage0 + ( 1 % NS ) * ( STAGE / 8 ) + row * PX + 2 * lx + 2 );
        if ( MODE == 5 )
        {
            // xi -> x' = xi + P e of planes kbeg - 1 and kbeg (my pair shares its parent)
            const double e0 = ebox( 0 )[jc[r] * C::EX + lx + 1], e1 = ebox( 1 % NS )[jc[r] * C::EX + lx + 1];
            zm[r].x = zm[r].x + e0;
            zm[r].y = zm[r].y + e0;
            cc[r].x = cc[r].x + e1;
            cc[r].y = cc[r].y + e1;
        }
        if ( MODE == 2 || MODE == 4 )
        {
            // r -> u = M^-1 r (MODE 4: b -> x1 = (omega1 D^-1) b) of planes kbeg - 1 and kbeg
            const int wzl = wall_count( g, 2, kbeg - 1 + g.off[2] ), wz0 = wall_count( g, 2, kbeg + g.off[2] );
            zm[r].x = mfac( wx0 + wyc[r] + wzl ) * zm[r].x;
            zm[r].y = mfac( wx1 + wyc[r] + wzl ) * zm[r].y;
            cc[r].x = mfac( wx0 + wyc[r] + wz0 ) * cc[r].x;
            cc[r].y = mfac( wx1 + wyc[r] + wz0 ) * cc[r].y;
        }
    }

    // slot 0 (plane kbeg-1) only feeds the zm registers: once every thread has read it, it takes
    // load NS.  From here on slot (it+1) % NS is released at the end of iteration `it`.
    __syncthreads();
    if ( tid == 0 && !FLAT && NS < nloads )
        issue( NS );

    dd_t acc = { 0.0, 0.0 }, acc2 = { 0.0, 0.0 }, acc3 = { 0.0, 0.0 }; // MODE 0: p.q | 1: r.r, r.M^-1 r | 2: r.r, r.u, w.u
    double* qrow = ( MODE != 1 ? a.q : a.r ) + geo_off( g, i0, y0 + wy, kbeg ); // MODE 1 writes r back in place
    for ( int it = 0; it < nplanes; ++it )
    {
        const int lc = it + 1, ln = it + 2; // load indices of plane k and plane k+1
        const int sc = lc % NS, sn = ln % NS;
        if ( !FLAT )
            mbar_wait( smem_u32( &full_bar[sn] ), ( ln / NS ) & 1 );
        const double* P = stage0 + sc * ( STAGE / 8 );
        const double* N = stage0 + sn * ( STAGE / 8 );
        const int wz = wall_count( g, 2, kbeg + it + g.off[2] );
#pragma unroll
        for ( int r = 0; r < RY; ++r )
        {
            const int row = wy + r * WY + 1;
            const double* pc = P + row * PX + 2 * lx + 2;
            double2 zp = FLAT ? make_double2( 0.0, 0.0 ) : *reinterpret_cast<const double2*>( N + row * PX + 2 * lx + 2 );
            double xl = pc[-1];
            double xr = pc[2];
            double2 ym = *reinterpret_cast<const double2*>( pc - PX );
            double2 yp = *reinterpret_cast<const double2*>( pc + PX );
            if ( MODE == 5 )
            {
                // the staged values are xi: x' = xi + P e of every neighbour, each with its own parent
                const double* En = ebox( sn );
                const double* Ec = ebox( sc );
                const double en = En[jc[r] * C::EX + lx + 1];
                zp.x = zp.x + en;
                zp.y = zp.y + en;
                xl = xl + Ec[jc[r] * C::EX + lx];
                xr = xr + Ec[jc[r] * C::EX + lx + 2];
                const double em = Ec[jm[r] * C::EX + lx + 1], ep = Ec[jp[r] * C::EX + lx + 1];
                ym.x = ym.x + em;
                ym.y = ym.y + em;
                yp.x = yp.x + ep;
                yp.y = yp.y + ep;
            }
            if ( MODE == 2 || MODE == 4 )
            {
                // the staged values are r (b): u = M^-1 r (x1 = omega1 D^-1 b) of every neighbour, each with its own
                // wall count
                const int wzn = wall_count( g, 2, kbeg + it + 1 + g.off[2] );
                zp.x = mfac( wx0 + wyc[r] + wzn ) * zp.x;
                zp.y = mfac( wx1 + wyc[r] + wzn ) * zp.y;
                xl = mfac( wxm + wyc[r] + wz ) * xl;
                xr = mfac( wxp + wyc[r] + wz ) * xr;
                ym.x = mfac( wx0 + wym[r] + wz ) * ym.x;
                ym.y = mfac( wx1 + wym[r] + wz ) * ym.y;
                yp.x = mfac( wx0 + wyp[r] + wz ) * yp.x;
                yp.y = mfac( wx1 + wyp[r] + wz ) * yp.y;
            }
            const double2 c = cc[r];
            const double d0 = op.diag[wx0 + wyc[r] + wz];
            const double d1 = op.diag[wx1 + wyc[r] + wz];
            const double a0 = apply_row( d0, ns, c.x, xl, c.y, ym.x, yp.x, zm[r].x, zp.x );
            const double a1 = apply_row( d1, ns, c.y, c.x, xr, ym.y, yp.y, zm[r].y, zp.y );
            if ( vy[r] )
            {
                double* qp = qrow + (long long)( r * WY ) * g.sy;
                if ( MODE == 0 )
                {
                    if ( vx1 )
                    {
                        *reinterpret_cast<double2*>( qp ) = make_double2( a0, a1 );
                        dd_acc( acc, c.x * a0 );
                        dd_acc( acc, c.y * a1 );
                    }
                    else if ( vx0 )
                    {
                        *qp = a0;
                        dd_acc( acc, c.x * a0 );
                    }
                }
                else if ( MODE == 2 )
                {
                    const double2 rc = *reinterpret_cast<const double2*>( pc ); // r of my pair
                    if ( vx1 )
                    {
                        *reinterpret_cast<double2*>( qp ) = make_double2( a0, a1 );
                        dd_acc( acc, rc.x * rc.x );
                        dd_acc( acc2, c.x * rc.x );
                        dd_acc( acc3, c.x * a0 );
                        dd_acc( acc, rc.y * rc.y );
                        dd_acc( acc2, c.y * rc.y );
                        dd_acc( acc3, c.y * a1 );
                    }
                    else if ( vx0 )
                    {
                        *qp = a0;
                        dd_acc( acc, rc.x * rc.x );
                        dd_acc( acc2, c.x * rc.x );
                        dd_acc( acc3, c.x * a0 );
                    }
                }
                else if ( MODE == 3 || MODE == 5 )
                {
                    // cell_smooth / cell_prolong_smooth (mg.cu): xo = x + (omega D^-1)(b - A x) with x = xi (MODE 5: xi + P e),
                    // b from the tile behind the box of xi
                    const int w0 = wx0 + wyc[r] + wz, w1 = wx1 + wyc[r] + wz;
                    const double2 bv = *reinterpret_cast<const double2*>( P + C::STAGE_BYTES / 8 + ( row - 1 ) * TX + 2 * lx );
                    const double z0 = fma( a.om1 * op.minv[w0], bv.x - a0, c.x );
                    if ( vx1 )
                    {
                        const double z1 = fma( a.om1 * op.minv[w1], bv.y - a1, c.y );
                        *reinterpret_cast<double2*>( qp ) = make_double2( z0, z1 );
                        if ( a.dot )
                        {
                            dd_acc( acc, z0 * bv.x );
                            dd_acc( acc, z1 * bv.y );
                        }
                    }
                    else if ( vx0 )
                    {
                        *qp = z0;
                        if ( a.dot )
                            dd_acc( acc, z0 * bv.x );
                    }
                }
                else if ( MODE == 4 )
                {
                    // cell_smooth02 (mg.cu): c = x1 of my pair; xo = x1 + (omega2 D^-1)(b - A x1), b from the plane itself
                    const int w0 = wx0 + wyc[r] + wz, w1 = wx1 + wyc[r] + wz;
                    const double2 bv = *reinterpret_cast<const double2*>( pc );
                    const double z0 = fma( a.om2 * op.minv[w0], bv.x - a0, c.x );
                    if ( vx1 )
                    {
                        const double z1 = fma( a.om2 * op.minv[w1], bv.y - a1, c.y );
                        *reinterpret_cast<double2*>( qp ) = make_double2( z0, z1 );
                    }
                    else if ( vx0 )
                        *qp = z0;
                }
                else
                {
                    // kernel 1's residual update + kernel 2's reduction (z = M^-1 r is never stored)
                    const int w0 = wx0 + wyc[r] + wz, w1 = wx1 + wyc[r] + wz;
                    // my pair of r from the tile behind the box of p (slot sc: its barrier has been waited on)
                    const double2 rcur = *reinterpret_cast<const double2*>( P + C::STAGE_BYTES / 8 + ( row - 1 ) * TX + 2 * lx );
                    if ( vx1 )
                    {
                        double2 rv = rcur;
                        rv.x = fma( nalpha, a0, rv.x );
                        rv.y = fma( nalpha, a1, rv.y );
                        *reinterpret_cast<double2*>( qp ) = rv;
                        dd_acc( acc, rv.x * rv.x );
                        dd_acc( acc2, ( op.minv[w0] * rv.x ) * rv.x );
                        dd_acc( acc, rv.y * rv.y );
                        dd_acc( acc2, ( op.minv[w1] * rv.y ) * rv.y );
                    }
                    else if ( vx0 )
                    {
                        const double rv = fma( nalpha, a0, rcur.x );
                        *qp = rv;
                        dd_acc( acc, rv * rv );
                        dd_acc( acc2, ( op.minv[w0] * rv ) * rv );
                    }
                }
            }
            zm[r] = c;
            cc[r] = zp;
        }
        qrow += g.sz;
        __syncthreads(); // every thread is done with slot sc -> it can be refilled
        if ( tid == 0 && !FLAT && lc + NS < nloads )
            issue( lc + NS );
    }

    if ( MODE == 0 )
    {
        dd_t vals[1] = { acc };
        if ( block_reduce_finalize<C::NT, 1>( vals, a.partials, a.pstride, &a.S->ticket[1] ) )
        {
            if ( tid == 0 )
                publish_pAp( a.S, vals[0] );
        }
    }
    else if ( MODE == 2 )
    {
        dd_t vals[3] = { acc, acc2, acc3 }; // r.r, gamma = r.u, delta = w.u
        if ( block_reduce_finalize<C::NT, 3>( vals, a.partials, a.pstride, &a.S->ticket[1] ) )
        {
            CgState* S = a.S;
            if ( tid == 0 && S->world > 1 )
                for ( int v = 0; v < 3; ++v )
                {
                    S->loc[2 * v] = vals[v].hi;
                    S->loc[2 * v + 1] = vals[v].lo;
                }
            if constexpr ( PF )
            {
                dd_t sum[3];
                peer_mail_exchange<3>( S, pf, 1, 0, sum );
                if ( tid == 0 )
                    cg1_finish( S, sum[0].hi + sum[0].lo, sum[1].hi + sum[1].lo, sum[2].hi + sum[2].lo, a.init );
            }
            else if ( tid == 0 && S->world == 1 )
                cg1_finish( S, vals[0].hi + vals[0].lo, vals[1].hi + vals[1].lo, vals[2].hi + vals[2].lo, a.init );
            // (several ranks over NCCL: cg_global_sum( c, 2 + init ) combines the local sums and finishes)
        }
    }
    else if ( MODE == 3 || MODE == 5 )
    {
        if ( a.dot ) // (uniform over the grid) sum z.b -> rz_new, as mg_smooth_dot_kernel / mg_publish leave it
        {
            dd_t vals[1] = { acc };
            if ( block_reduce_finalize<C::NT, 1>( vals, a.partials, a.pstride, &a.S->ticket[1] ) )
            {
                if ( tid == 0 )
                {
                    CgState* S = a.S;
                    if ( S->world > 1 )
                    {
                        S->loc[0] = vals[0].hi;
                        S->loc[1] = vals[0].lo;
                    }
                    else
                        S->rz_new = vals[0].hi + vals[0].lo;
                }
            }
        }
    }
    else if ( MODE == 4 )
    {
    }
    else
    {
        dd_t vals[2] = { acc, acc2 };
        if ( block_reduce_finalize<C::NT, 2>( vals, a.partials, a.pstride, &a.S->ticket[1] ) )
        {
            if ( tid == 0 )
            {
                // publish_rr_rz of kernels_cg.cu
                CgState* S = a.S;
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
        }
    }
}

// ---- variant 1: no shared memory, register z-blocking, neighbours through L1/L2 -------------
// block = 32 x 8 threads, each thread owns one column pair and marches ZC planes.
__global__ void __launch_bounds__( 256, 4 )
    stencil7_dot_ldg( const __grid_constant__ Geo g, const __grid_constant__ OpConst op,
                      const double* __restrict__ p, const __grid_constant__ StencilArgs a )
{
    if ( a.S->done )
        return;
    const int lx = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int u = blockIdx.x;
    const int tx = u % a.tiles_x;
    const int ty = ( u / a.tiles_x ) % a.tiles_y;
    const int ch = u / ( a.tiles_x * a.tiles_y );
    const int i0 = tx * 64 + 2 * lx, j = ty * 8 + wy;
    const int kbeg = ch * a.zc, kend = min( kbeg + a.zc, g.n[2] );
    const bool vy = j < g.n[1], vx0 = i0 < g.n[0], vx1 = i0 + 1 < g.n[0];
    const int wxy0 = wall_count( g, 0, i0 + g.off[0] ) + wall_count( g, 1, j + g.off[1] );
    const int wxy1 = wall_count( g, 0, i0 + 1 + g.off[0] ) + wall_count( g, 1, j + g.off[1] );
    const double ns = op.neg_scale;
    dd_t acc = { 0.0, 0.0 };
    if ( vy && vx0 )
    {
        const double* pp = p + geo_off( g, i0, j, kbeg );
        double* qp = a.q + geo_off( g, i0, j, kbeg );
        double2 zm = *reinterpret_cast<const double2*>( pp - g.sz );
        double2 c = *reinterpret_cast<const double2*>( pp );
        for ( int k = kbeg; k < kend; ++k )
        {
            const double2 zp = __ldg( reinterpret_cast<const double2*>( pp + g.sz ) );
            const double2 ym = __ldg( reinterpret_cast<const double2*>( pp - g.sy ) );
            const double2 yp = __ldg( reinterpret_cast<const double2*>( pp + g.sy ) );
            const double xl = __ldg( pp - 1 );
            const double xr = __ldg( pp + 2 );
            const int wz = wall_count( g, 2, k + g.off[2] );
            const double a0 = apply_row( op.diag[wxy0 + wz], ns, c.x, xl, c.y, ym.x, yp.x, zm.x, zp.x );
            const double a1 = apply_row( op.diag[wxy1 + wz], ns, c.y, c.x, xr, ym.y, yp.y, zm.y, zp.y );
            if ( vx1 )
            {
                *reinterpret_cast<double2*>( qp ) = make_double2( a0, a1 );
                dd_acc( acc, c.x * a0 );
                dd_acc( acc, c.y * a1 );
            }
            else
            {
                *qp = a0;
                dd_acc( acc, c.x * a0 );
            }
            zm = c;
            c = zp;
            pp += g.sz;
            qp += g.sz;
        }
    }
    dd_t vals[1] = { acc };
    if ( block_reduce_finalize<256, 1>( vals, a.partials, a.pstride, &a.S->ticket[1] ) )
    {
        if ( threadIdx.x == 0 )
            publish_pAp( a.S, vals[0] );
    }
}

typedef CUresult ( *PFN_encodeTiled )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill );

template <class C, int MODE>
int launch_tma_mode( cfb_ctx* c, const StencilArgs& a, int grid, const PeerFusedArgs* pf )
{
    constexpr int SMEM = stencil_smem_bytes<C, MODE>();
    static bool attr_set = false;
    if ( !attr_set )
    {
        cudaFuncSetAttribute( stencil7_dot_tma<C, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM );
        cudaFuncSetAttribute( stencil7_dot_tma<C, MODE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM );
        if constexpr ( MODE == 2 )
            cudaFuncSetAttribute( stencil7_dot_tma<C, MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM );
        if constexpr ( MODE == 1 )
        {
            // three (two) CTAs of the small (medium) tilings fill the SM's shared memory: ask for all of it
            cudaFuncSetAttribute( stencil7_dot_tma<C, MODE, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared );
            cudaFuncSetAttribute( stencil7_dot_tma<C, MODE, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared );
        }
        attr_set = true;
    }
    const NoPeerArgs none{};
    const CUtensorMap& tm = MODE == 2 ? c->tmap_sr : c->tmap_p; // MODE 2 marches over r, the others over p
    if constexpr ( MODE == 2 )
    {
        if ( pf ) // (the callers have made sure that flat does not apply)
        {
            stencil7_dot_tma<C, MODE, false, true><<<grid, C::NT, SMEM, c->stream>>>( tm, c->tmap_r1, c->tmap_r1, c->g, c->op, a, *pf );
            return 1;
        }
    }
    if ( c->g.D == 2 && c->flat_2d )
        stencil7_dot_tma<C, MODE, true, false><<<grid, C::NT, SMEM, c->stream>>>( tm, c->tmap_r1, c->tmap_r1, c->g, c->op, a, none );
    else
        stencil7_dot_tma<C, MODE, false, false><<<grid, C::NT, SMEM, c->stream>>>( tm, c->tmap_r1, c->tmap_r1, c->g, c->op, a, none );
    return 1;
}
template <class C>
int launch_tma( cfb_ctx* c, const StencilArgs& a, int grid, int mode, const PeerFusedArgs* pf )
{
    if ( mode == 0 )
        return launch_tma_mode<C, 0>( c, a, grid, nullptr );
    if ( mode == 1 )
        return launch_tma_mode<C, 1>( c, a, grid, nullptr );
    return launch_tma_mode<C, 2>( c, a, grid, pf );
}

} // namespace

// One 3-D float64 tensor map over an array in the layout of the CG vectors (Geo): boxes of bx x by x 1 entries.
static int encode_map_dims( cfb_ctx* c, CUtensorMap* map, double* base, int bx, int by, long long sy, long long sz, int ay, int az );
static int encode_map( cfb_ctx* c, CUtensorMap* map, double* base, int bx, int by )
{
    return encode_map_dims( c, map, base, bx, by, c->g.sy, c->g.sz, c->g.ay, c->g.az );
}
// ... over any x-fastest array with row stride sy, plane stride sz (doubles), ay rows, az planes
static int encode_map_dims( cfb_ctx* c, CUtensorMap* map, double* base, int bx, int by, long long sy, long long sz, int ay, int az )
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
    cuuint64_t gdim[3] = { (cuuint64_t)sy, (cuuint64_t)ay, (cuuint64_t)az };
    cuuint64_t gstride[2] = { (cuuint64_t)sy * 8, (cuuint64_t)sz * 8 };
    cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = encode( map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
    if ( r != CUDA_SUCCESS )
        return cfb_fail( c, CFB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string( (int)r ) );
    return CFB_OK;
}

// (Re)build the tensor maps of the CG vectors for the current tile shape: the stencil box (tile + halo) of both
// buffers of the search direction and of r (single-reduction form), the halo-free tile of r (phase A').
int stencil_setup( cfb_ctx* c )
{
    for ( int b = 0; b < 2; ++b )
        if ( int rc = encode_map( c, &c->tmap_pbuf[b], c->cg_pbuf[b], c->st_tx + 4, c->st_ty + 2 ) )
            return rc;
    if ( int rc = encode_map( c, &c->tmap_sr, c->cg_r, c->st_tx + 4, c->st_ty + 2 ) )
        return rc;
    if ( int rc = encode_map( c, &c->tmap_r1, c->cg_r, c->st_tx, c->st_ty ) )
        return rc;
    c->tmap_p = c->tmap_pbuf[c->pcur];
    c->tmap_ok = true;
    return CFB_OK;
}

// z chunk of the marches that run three CTAs per SM on 64 x 16 tiles (phase A', the multigrid sweeps) — measured with
// phase A' (profiles/r2_sweep_forms2.log): 16-plane chunks once they make three waves or more — many short units keep
// the tail of the last wave short (512^3: 529 us against 540 / 570 with 32 / 64 planes; 320^3: 149 against 157 / 175)
// —, otherwise ONE wave of at most two units per SM, chunks of equal length (256^3: 4 chunks of 64 planes 81 us, 1.15
// waves of 32-plane chunks 103 us; 192^3: 8 chunks of 24 planes 42.5 us against 56 with 64-plane chunks)
static int march_chunk( const cfb_ctx* c, long long tiles )
{
    const Geo& g = c->g;
    if ( tiles * ( ( g.n[2] + 15 ) / 16 ) >= 9LL * c->sm_count )
        return 16;
    long long nch = 2LL * c->sm_count / tiles;
    nch = std::max( 1LL, std::min( nch, (long long)( g.n[2] + 7 ) / 8 ) );
    return (int)( ( g.n[2] + nch - 1 ) / nch );
}

static int launch_stencil( cfb_ctx* c, int mode, const PeerFusedArgs* pf = nullptr, int init = 0 )
{
    const Geo& g = c->g;
    StencilArgs a{};
    a.r = c->cg_r;
    a.q = c->cg_q;
    a.S = c->d_state;
    a.init = init;
    a.hx = 16;
    const int tx = c->st_variant == 1 ? 64 : c->st_tx;
    const int ty = c->st_variant == 1 ? 8 : c->st_ty;
    if ( tx < 2 || ty < 1 )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "unsupported stencil tile configuration" ) );
        return 0;
    }
    a.tiles_x = ( g.n[0] + tx - 1 ) / tx;
    a.tiles_y = ( g.n[1] + ty - 1 ) / ty;
    // (64-plane chunks whatever the block: shortening them until every CTA slot has a unit was measured and loses at
    // 256^3 — 4262 vs 4418 iterations/s in the 64-byte form, profiles/r2_cg_forms_by_size.json — every chunk re-reads
    // two planes)
    int zc = c->st_zc > 0 ? c->st_zc : g.n[2];
    // phase A' picks its own chunks unless "stencil_zc" was set
    if ( mode == 1 && c->st_zc_auto && g.D == 3 )
        zc = march_chunk( c, (long long)a.tiles_x * a.tiles_y );
    a.zc = zc;
    // one block per unit, one partial sum per block: the scratch follows the unit count (large cross-sections,
    // e.g. two-dimensional grids beyond 2048^2, have more than CFB_MAX_PARTIALS tiles in a single plane)
    const long long units = (long long)a.tiles_x * a.tiles_y * ( ( g.n[2] + zc - 1 ) / zc );
    if ( note_rc( c, ensure_partials( c, units ) ) )
        return 0;
    a.partials = c->d_partials;
    a.pstride = c->partials_cap;
    const int grid = (int)units;
    if ( c->st_variant == 1 )
    {
        if ( mode != 0 )
        {
            note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "the LDG stencil variant has no phase A' / single-reduction form" ) );
            return 0;
        }
        stencil7_dot_ldg<<<grid, 256, 0, c->stream>>>( g, c->op, c->cg_p, a );
        return 1;
    }
    const int key = c->st_tx * 10000 + c->st_ty * 100 + c->st_stages;
    switch ( key )
    {
    case 641604:
        return launch_tma<TileCfg<64, 16, 4>>( c, a, grid, mode, pf );
    case 641606:
        return launch_tma<TileCfg<64, 16, 6>>( c, a, grid, mode, pf );
    case 640804:
        return launch_tma<TileCfg<64, 8, 4>>( c, a, grid, mode, pf );
    case 643204:
        return launch_tma<TileCfg<64, 32, 4>>( c, a, grid, mode, pf );
    case 643203:
        return launch_tma<TileCfg<64, 32, 3>>( c, a, grid, mode, pf );
    case 1281604:
        return launch_tma<TileCfg<128, 16, 4>>( c, a, grid, mode, pf );
    case 1281603:
        return launch_tma<TileCfg<128, 16, 3>>( c, a, grid, mode, pf );
    case 1283203:
        return launch_tma<TileCfg<128, 32, 3>>( c, a, grid, mode, pf );
    case 1280804:
        return launch_tma<TileCfg<128, 8, 4>>( c, a, grid, mode, pf );
    default:
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "unsupported stencil tile configuration" ) );
        return 0;
    }
}

int launch_stencil_dot( cfb_ctx* c ) { return launch_stencil( c, 0 ); }

// cg_variant 3 (single-reduction CG): w = A M^-1 r, sum r^2, sum r.u, sum w.u
int launch_cg1_stencil( cfb_ctx* c, int init, bool mail )
{
    if ( c->st_variant == 1 )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "the LDG stencil variant has no single-reduction form" ) );
        return 0;
    }
    if ( mail )
    {
        if ( c->g.D == 2 && c->flat_2d )
        {
            note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "flat_2d has no mailbox instantiation" ) );
            return 0;
        }
        PeerFusedArgs pf{};
        peer_mail_only( c, pf );
        return launch_stencil( c, 2, &pf, init );
    }
    return launch_stencil( c, 2, nullptr, init );
}

// phase A' of the 64-byte iteration (cg_variant 2): r -= alpha (A p) with q recomputed, sum r^2, sum r.M^-1 r
int launch_stencil_rupdate( cfb_ctx* c ) { return launch_stencil( c, 1 ); }

// ---------------------------------------------------------------------------------------------
// Fine-level smoothing sweeps of the multigrid preconditioner on the TMA z-march (MODE 3 / 4 above; mg.cu calls these
// for level 0 of three-dimensional runs, whose arrays live in the layout of the CG vectors).  64 x 16 tiles, four
// stages, three CTAs per SM, the chunk rule of phase A'.  Returns the number of launches, or -1 when the march does
// not apply (mg.cu then runs its one-thread-per-cell kernels).
namespace
{
using MgTile = TileCfg<64, 16, 4>;

// shared-memory attributes of a mode's kernel, once per process — from mg_tma_prepare (when the hierarchy is built), not
// from the first launch, which may sit inside the stream capture of "mg_graph"
template <int MODE>
void mg_mode_attrs()
{
    static bool attr_set = false;
    if ( attr_set )
        return;
    cudaFuncSetAttribute( stencil7_dot_tma<MgTile, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          stencil_smem_bytes<MgTile, MODE>() );
    cudaFuncSetAttribute( stencil7_dot_tma<MgTile, MODE, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                          cudaSharedmemCarveoutMaxShared );
    attr_set = true;
}

template <int MODE>
int launch_mg_mode( cfb_ctx* c, const CUtensorMap& box, const CUtensorMap& tile, const CUtensorMap& ebox, const OpConst& op,
                    StencilArgs& a )
{
    constexpr int SMEM = stencil_smem_bytes<MgTile, MODE>();
    mg_mode_attrs<MODE>();
    const Geo& g = c->g;
    a.S = c->d_state;
    a.hx = 16;
    a.tiles_x = ( g.n[0] + MgTile::TX - 1 ) / MgTile::TX;
    a.tiles_y = ( g.n[1] + MgTile::TY - 1 ) / MgTile::TY;
    a.zc = march_chunk( c, (long long)a.tiles_x * a.tiles_y );
    const long long units = (long long)a.tiles_x * a.tiles_y * ( ( g.n[2] + a.zc - 1 ) / a.zc );
    // (the scratch was sized by mg_tma_prepare when the hierarchy was built: nothing is allocated here, where a
    // graph capture may be under way)
    if ( units > c->partials_cap )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "multigrid march: partial-sum scratch not prepared" ) );
        return 0;
    }
    a.partials = c->d_partials;
    a.pstride = c->partials_cap;
    stencil7_dot_tma<MgTile, MODE, false, false><<<(int)units, MgTile::NT, SMEM, c->stream>>>( box, tile, ebox, g, op, a, NoPeerArgs{} );
    return 1;
}

// the tensor maps of the three fine-level arrays a cycle touches, rebuilt when the arrays change (a new hierarchy)
int mg_maps( cfb_ctx* c, double* b, double* x0, double* x1 )
{
    if ( c->mg_map_ptr[0] == b && c->mg_map_ptr[1] == x0 && c->mg_map_ptr[2] == x1 )
        return CFB_OK;
    double* base[3] = { b, x0, x1 };
    for ( int m = 0; m < 3; ++m )
        if ( int rc = encode_map( c, &c->mg_map_box[m], base[m], MgTile::TX + 4, MgTile::TY + 2 ) )
            return rc;
    if ( int rc = encode_map( c, &c->mg_map_tile_b, b, MgTile::TX, MgTile::TY ) )
        return rc;
    for ( int m = 0; m < 3; ++m )
        c->mg_map_ptr[m] = base[m];
    c->mg_map_e_ptr[0] = c->mg_map_e_ptr[1] = nullptr; // (the coarse level's buffers belong to the same hierarchy)
    return CFB_OK;
}
} // namespace

// the partial-sum scratch for the march's unit count, before any launch (mg_build)
int mg_tma_prepare( cfb_ctx* c )
{
    if ( !mg_tma_applies( c ) )
        return CFB_OK;
    mg_mode_attrs<3>();
    mg_mode_attrs<4>();
    mg_mode_attrs<5>();
    const Geo& g = c->g;
    const long long tiles = (long long)( ( g.n[0] + MgTile::TX - 1 ) / MgTile::TX ) * ( ( g.n[1] + MgTile::TY - 1 ) / MgTile::TY );
    const int zc = march_chunk( c, tiles );
    return ensure_partials( c, tiles * ( ( g.n[2] + zc - 1 ) / zc ) );
}

bool mg_tma_applies( const cfb_ctx* c )
{
    const bool on = c->mg_tma > 0 || ( c->mg_tma < 0 && !c->cfg.use_nccl );
    return on && c->g.D == 3 && c->st_variant == 0 && c->g.n[0] >= 16 && c->g.n[1] >= 4;
}

// xo = xi + (omega D^-1)(b - A xi) on the fine level; dot: also sum xo.b -> rz_new (mg_smooth_kernel / mg_smooth_dot_kernel).
// b, x0, x1: the level's right-hand side and its two iterate buffers; xi / xo are x0 / x1 in one order or the other.
int launch_mg_smooth_tma( cfb_ctx* c, const OpConst& op, double omega, double* b, double* x0, double* x1, int xi_is, int dot )
{
    if ( !mg_tma_applies( c ) )
        return -1;
    if ( note_rc( c, mg_maps( c, b, x0, x1 ) ) )
        return 0;
    StencilArgs a{};
    a.q = xi_is == 0 ? x1 : x0;
    a.om1 = omega;
    a.dot = dot;
    return launch_mg_mode<3>( c, c->mg_map_box[1 + xi_is], c->mg_map_tile_b, c->mg_map_tile_b, op, a );
}

// the first two sweeps from a zero initial guess in one pass: xo = x1 + (omega2 D^-1)(b - A x1), x1 = (omega1 D^-1) b
// (mg_smooth02_kernel); the result goes to the level's second buffer
int launch_mg_smooth02_tma( cfb_ctx* c, const OpConst& op, double omega1, double omega2, double* b, double* x0, double* x1 )
{
    if ( !mg_tma_applies( c ) )
        return -1;
    if ( note_rc( c, mg_maps( c, b, x0, x1 ) ) )
        return 0;
    StencilArgs a{};
    a.q = x1;
    a.om1 = omega1;
    a.om2 = omega2;
    return launch_mg_mode<4>( c, c->mg_map_box[0], c->mg_map_tile_b, c->mg_map_tile_b, op, a );
}

// prolongation + correction + first post-smoothing sweep: xo = x' + (omega D^-1)(b - A x'), x' = xi + P e
// (mg_prolong_smooth_kernel).  e: the coarse level's current iterate, an array with one ghost layer, row stride csy,
// plane stride csz, cn[] owned cells.  The coarse rows must be multiples of 16 bytes (TMA): fine extents divisible by 4.
int launch_mg_prolong_smooth_tma( cfb_ctx* c, const OpConst& op, double omega, double* b, double* x0, double* x1, int xi_is,
                                  double* e, long long csy, long long csz, const int cn[3], int dot )
{
    if ( !mg_tma_applies( c ) || !c->mg_tma_prolong || ( csy & 1 ) || ( csz & 1 ) || ( c->g.n[2] & 1 ) )
        return -1;
    if ( note_rc( c, mg_maps( c, b, x0, x1 ) ) )
        return 0;
    int slot = -1;
    for ( int q = 0; q < 2; ++q )
        if ( c->mg_map_e_ptr[q] == e )
            slot = q;
    if ( slot < 0 )
    {
        slot = c->mg_map_e_ptr[0] ? 1 : 0; // the coarse level has two iterate buffers
        if ( note_rc( c, encode_map_dims( c, &c->mg_map_e[slot], e, MgTile::EX, MgTile::EY, csy, csz, cn[1] + 2, cn[2] + 2 ) ) )
            return 0;
        c->mg_map_e_ptr[slot] = e;
    }
    StencilArgs a{};
    a.q = xi_is == 0 ? x1 : x0;
    a.om1 = omega;
    a.dot = dot;
    a.ehz = 1;
    return launch_mg_mode<5>( c, c->mg_map_box[1 + xi_is], c->mg_map_tile_b, c->mg_map_e[slot], op, a );
}

