// mg.cu — opt-in geometric multigrid preconditioner for the pressure CG.
//
// The reference's default path hands the pressure system to HYPRE (PCG + PFMG,
// examples/advection.cpp:186-189, src/VelocityCorrector.hpp:324-338); its "Reference" path — the one
// this library reproduces bit for bit — is Jacobi-PCG, whose iteration count grows like n (292 at
// 64^3, 582 at 128^3, > 2000 at 512^3).  This file adds what PFMG is there for, behind
// cfb_set_preconditioner( CFB_PRECOND_MG ): the same CG (same matrix, same absolute stopping test,
// same exactly-accumulated dot products) with   z = M^-1 r := one V(nu1, nu2) cycle   instead of
// z = D^-1 r.  It is NOT a restatement of HYPRE; the CPU checker (the mg_* functions under oracle/) states
// the same algorithm operation for operation, so the two agree bit for bit, and the solution agrees
// with the Jacobi path to the solver tolerance (tests/test_zz_multigrid.py).  Never the default:
// north_star pins "same preconditioner as the reference" for the parity runs.
//
//   levels      : cell-centred 2:1 coarsening while every extent stays even and >= 2 after halving;
//                 level l re-discretises the same 2*D+1-point operator (same wall logic) with
//                 scale_l = scale_0 / 4^l
//   smoother    : damped Jacobi  x += omega D^-1 (b - A x)  (first sweep from x = 0: x = omega D^-1 b), one
//                 omega per sweep: by default the reciprocals of the Chebyshev nodes of [0.4, 2] in 3-D with
//                 2..4 sweeps (0.566, 1.577 for V(2,2): 8 / 8 / 9 instead of 11 / 12 / 14 iterations at
//                 32^3 / 64^3 / 128^3 for nothing), reversed after the coarse correction; 6/7 (3-D, one sweep),
//                 0.8 (2-D) otherwise,
//                 ping-pong between two arrays per level
//   restriction : mean of the 2^D children of the residual (residual fused into the kernel);
//                 prolongation: piecewise constant, fused with the correction
//   coarsest    : nuc Jacobi sweeps
//   several GPUs: the SAME global cycle, block-decomposed like the CG (not a block-local preconditioner: that
//                 loses the coarse space across blocks, 49 instead of 12 iterations at 64^3 on 2x2x2 blocks);
//                 every level is coarsened block by block (identical block extents required), ghost layers
//                 are refreshed before each operator application by a one-layer face exchange (pack kernel,
//                 one grouped NCCL send/recv, unpack kernel), sums are all-gathered and combined exactly, so
//                 iteration counts and results do not depend on the decomposition
//   null space  : with all walls SOLID the operator is singular; x is shifted at the end so that
//                 sum_i d_i x_i = 0, the component Jacobi-PCG produces (quirk Q1 of the reference leaks
//                 that constant into v, src/VelocityCorrector.hpp:260)
//
// All kernels here are plain one-thread-per-cell streaming kernels (grid-stride walks with MgCursor; HBM-bound on the fine level,
// launch-bound on the coarse ones); q = A p stays the TMA stencil kernel.  Bytes per fine cell and CG
// iteration with V(2,2): both pre-smoothing sweeps in one pass 16, residual + restriction 17, prolongation
// fused with the first post-smoothing sweep 24, second sweep fused with z.r 24, p-update 24, stencil 16,
// axpy 48 = 169 (+ 1/7 of the 81 of the cycle for the coarse levels) against 72 for a Jacobi iteration,
// for ~10 iterations instead of ~2500 at 512^3.
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"

#include <algorithm>
#include <cmath>

namespace
{

constexpr int NT = 256;

struct MgLevelDev
{
    int n[3];
    int cz;             // coarsening factor to the next level along z (1 in 2-D)
    int slo[3], shi[3]; // the low / high end of dim d is a SOLID physical wall
    int nlo[3], nhi[3]; // a neighbouring block continues the grid there: the ghost layer holds its values
                        // after an exchange (otherwise ghosts are the zeros the operator reads off the domain)
    long long sy, sz, origin;
    double ns;          // off-diagonal coefficient, -scale_l
    double diag[8], minv[8]; // by number of SOLID walls touched: diagonal, 1 / diagonal
};

struct MgLevelHost
{
    MgLevelDev d;
    double* b = nullptr;
    double* x[2] = { nullptr, nullptr };
    int cur = 0;
    bool owns_b = false;
    long long cells = 0;
};

} // namespace

struct MgStage
{
    std::vector<MgLevelHost> lv;
    int nu1 = 2, nu2 = 2, nuc = 8;
    int max_levels = 0; // 0 = as many as the block allows
    double omega = 0.0;               // the caller's argument (<= 0: default schedule)
    std::vector<double> wpre, wpost;  // damping of every pre- / post-smoothing sweep
    double wc = 0.0;                  // damping on the coarsest level
    bool singular = false;
    cudaEvent_t ev_poll = nullptr;
    // "mg_graph" tuning key (one block only): the ~60 launches of a V-cycle replayed as one CUDA graph.  All
    // pointers, the ping-pong sequence and the launch shapes of a cycle are the same every time, so the graph
    // is captured once per form ([0] plain cycle, [1] with the fused z.r of the CG) and reused.
    // several blocks with NVLink peer memory: the neighbours' level arrays mapped into this process, so that a
    // ghost exchange is ONE kernel that stores my boundary layers into their ghost layers and meets them at a
    // flag barrier (mg_xchg_kernel) instead of pack kernel + NCCL send/recv + unpack kernel
    bool peer = false;
    std::vector<double*> own_arrays;  // what was exported, in the order all ranks use
    std::vector<double*> peer_arrays; // [face s][array a] -> peer_arrays[s * own_arrays.size() + a]
    unsigned long long xseq = 0;      // exchanges so far (the same on every rank)
    const double* last_xchg = nullptr; // the array of the previous exchange
    // "mg_coarse_kernel" tuning key (one block only): levels [coarse_start, last] run in one single-CTA kernel
    bool use_coarse = false;
    int coarse_start = -1;
    bool use_graph = false;
    cudaGraphExec_t graph[2] = { nullptr, nullptr };
    int graph_launches[2] = { 0, 0 };
    bool graph_dotted[2] = { false, false };
    int graph_cur0[2] = { 0, 0 }; // where the fine level's result ends up
};

namespace
{

__device__ __forceinline__ int mg_walls( const MgLevelDev& L, int i, int j, int k )
{
    return ( i == 0 && L.slo[0] ) + ( i == L.n[0] - 1 && L.shi[0] ) + ( j == 0 && L.slo[1] ) +
           ( j == L.n[1] - 1 && L.shi[1] ) + ( k == 0 && L.slo[2] ) + ( k == L.n[2] - 1 && L.shi[2] );
}

__device__ __forceinline__ long long mg_off( const MgLevelDev& L, int i, int j, int k )
{
    return L.origin + (long long)k * L.sz + (long long)j * L.sy + i;
}

// v -> (i, j, k) of a level with x extent n0 and y extent n1: two divisions, 32-bit whenever v fits
__device__ __forceinline__ void mg_split( long long v, int n0, int n1, int& i, int& j, int& k )
{
    if ( v < 0x7fffffffll )
    {
        const unsigned u = (unsigned)v, r = u / (unsigned)n0;
        i = (int)( u - r * (unsigned)n0 );
        k = (int)( r / (unsigned)n1 );
        j = (int)( r - (unsigned)k * (unsigned)n1 );
    }
    else
    {
        const long long r = v / n0;
        i = (int)( v - r * n0 );
        k = (int)( r / n1 );
        j = (int)( r - (long long)k * n1 );
    }
}

// A thread's walk over the cells t = first, first + stride, ... < total of a level (x fastest) with (i, j, k)
// carried along: the divisions are done once per thread — for the start and, if the thread has a second cell at
// all, for the stride — and every further cell costs adds and compares.  (A 64-bit division per cell and
// coordinate is ~100 instructions: more issue slots than the 16-24 bytes of a fine-level cell take from HBM.)
struct MgCursor
{
    long long t, stride;
    int i, j, k, di, dj, dk, n0, n1;
    __device__ __forceinline__ MgCursor( const MgLevelDev& L, long long first, long long stride_, long long total )
        : t( first ), stride( stride_ ), di( 0 ), dj( 0 ), dk( 0 ), n0( L.n[0] ), n1( L.n[1] )
    {
        mg_split( first, n0, n1, i, j, k );
        if ( first + stride_ < total )
            mg_split( stride_, n0, n1, di, dj, dk );
    }
    __device__ __forceinline__ void next()
    {
        t += stride;
        i += di;
        int c = i >= n0 ? 1 : 0;
        i -= c ? n0 : 0;
        j += dj + c;
        c = j >= n1 ? 1 : 0;
        j -= c ? n1 : 0;
        k += dk + c;
    }
};
// all cells of a level, by the threads of the grid / of one CTA
#define MG_GRID_CELLS( cu, L, total )                                                                              \
    for ( MgCursor cu( L, blockIdx.x * (long long)NT + threadIdx.x, (long long)gridDim.x * NT, total ); cu.t < total; cu.next() )
#define MG_CTA_CELLS( cu, L, total ) for ( MgCursor cu( L, tid, NT, total ); cu.t < total; cu.next() )

// coarse index of fine index i (floor division, so that the ghost index -1 maps to the coarse ghost -1)
__device__ __forceinline__ int mg_parent( int i, int f ) { return f == 1 ? i : ( ( i + 2 ) >> 1 ) - 1; }

// (A x)(i,j,k): diag * x, then one fused multiply-add per neighbour in stencil order
__device__ __forceinline__ double mg_Ax( const MgLevelDev& L, const double* __restrict__ x, long long o, int w )
{
    return apply_row( L.diag[w], L.ns, x[o], x[o - 1], x[o + 1], x[o - L.sy], x[o + L.sy], x[o - L.sz],
                      x[o + L.sz] );
}

// ---- the element-wise operations of the cycle, one cell each (shared by the grid-wide kernels below and by
// the single-CTA kernel that runs the coarse end of the cycle) ------------------------------------------
__device__ __forceinline__ void cell_smooth0( const MgLevelDev& L, double omega, const double* __restrict__ b,
                                              double* __restrict__ x, const MgCursor& cu )
{
    const int i = cu.i, j = cu.j, k = cu.k;
    const long long o = mg_off( L, i, j, k );
    x[o] = ( omega * L.minv[mg_walls( L, i, j, k )] ) * b[o];
}

__device__ __forceinline__ double cell_smooth( const MgLevelDev& L, double omega, const double* __restrict__ b,
                                               const double* __restrict__ xi, double* __restrict__ xo, const MgCursor& cu,
                                               double* bv_out = nullptr )
{
    const int i = cu.i, j = cu.j, k = cu.k;
    const long long o = mg_off( L, i, j, k );
    const int w = mg_walls( L, i, j, k );
    const double bv = b[o];
    const double res = bv - mg_Ax( L, xi, o, w );
    const double z = fma( omega * L.minv[w], res, xi[o] );
    xo[o] = z;
    if ( bv_out )
        *bv_out = bv;
    return z;
}

// The first TWO sweeps from a zero initial guess in one pass (16 instead of 16 + 24 bytes per cell):
// x1 = (omega D^-1) b is recomputed for the six neighbours from b itself (neighbours off the block are
// the ghost zeros the unfused sweep would read), then x2 = x1 + omega D^-1 (b - A x1).
__device__ __forceinline__ void cell_smooth02( const MgLevelDev& L, double omega1, double omega2,
                                               const double* __restrict__ b, double* __restrict__ xo, const MgCursor& cu )
{
    const int i = cu.i, j = cu.j, k = cu.k;
    const long long o = mg_off( L, i, j, k );
    // Fast path (all but the two outermost layers of the block, i.e. ~98 % of a 512^3 level): neither the cell nor
    // any of its six neighbours touches a wall, so every D^-1 is minv[0] and every neighbour exists — the same
    // products and the same row as below, without seven wall counts (this kernel is issue-bound, not HBM-bound:
    // profiles/r2_launches_mg512.csv)
    if ( i >= 2 && i < L.n[0] - 2 && j >= 2 && j < L.n[1] - 2 && k >= 2 && k < L.n[2] - 2 )
    {
        const double m1 = omega1 * L.minv[0];
        const double bc = b[o];
        const double xc = m1 * bc;
        const double res = bc - apply_row( L.diag[0], L.ns, xc, m1 * b[o - 1], m1 * b[o + 1], m1 * b[o - L.sy], m1 * b[o + L.sy],
                                           m1 * b[o - L.sz], m1 * b[o + L.sz] );
        xo[o] = fma( omega2 * L.minv[0], res, xc );
        return;
    }
    const int w = mg_walls( L, i, j, k );
    const double bc = b[o];
    const double xc = ( omega1 * L.minv[w] ) * bc;
    // a neighbour across a block interface is a ghost entry of b (exchanged by the caller); its wall
    // count is that of an interior cell along the interface normal, which mg_walls returns for -1 / n
    const double xm = ( i > 0 || L.nlo[0] ) ? ( omega1 * L.minv[mg_walls( L, i - 1, j, k )] ) * b[o - 1] : 0.0;
    const double xp = ( i < L.n[0] - 1 || L.nhi[0] ) ? ( omega1 * L.minv[mg_walls( L, i + 1, j, k )] ) * b[o + 1] : 0.0;
    const double ym = ( j > 0 || L.nlo[1] ) ? ( omega1 * L.minv[mg_walls( L, i, j - 1, k )] ) * b[o - L.sy] : 0.0;
    const double yp = ( j < L.n[1] - 1 || L.nhi[1] ) ? ( omega1 * L.minv[mg_walls( L, i, j + 1, k )] ) * b[o + L.sy] : 0.0;
    const double zm = ( k > 0 || L.nlo[2] ) ? ( omega1 * L.minv[mg_walls( L, i, j, k - 1 )] ) * b[o - L.sz] : 0.0;
    const double zp = ( k < L.n[2] - 1 || L.nhi[2] ) ? ( omega1 * L.minv[mg_walls( L, i, j, k + 1 )] ) * b[o + L.sz] : 0.0;
    const double res = bc - apply_row( L.diag[w], L.ns, xc, xm, xp, ym, yp, zm, zp );
    xo[o] = fma( omega2 * L.minv[w], res, xc );
}

__global__ void __launch_bounds__( NT )
    mg_smooth0_kernel( const __grid_constant__ MgLevelDev L, double omega, const double* __restrict__ b,
                       double* __restrict__ x )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    MG_GRID_CELLS( cu, L, total )
        cell_smooth0( L, omega, b, x, cu );
}

__global__ void __launch_bounds__( NT )
    mg_smooth_kernel( const __grid_constant__ MgLevelDev L, double omega, const double* __restrict__ b,
                      const double* __restrict__ xi, double* __restrict__ xo )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    MG_GRID_CELLS( cu, L, total )
        cell_smooth( L, omega, b, xi, xo, cu );
}

__global__ void __launch_bounds__( NT )
    mg_smooth02_kernel( const __grid_constant__ MgLevelDev L, double omega1, double omega2,
                        const double* __restrict__ b, double* __restrict__ xo )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    MG_GRID_CELLS( cu, L, total )
        cell_smooth02( L, omega1, omega2, b, xo, cu );
}

// End of a reduction.  One block: the rounded sum is final.  Several blocks (ranks): keep the local
// double-double in S->loc[0..1]; the host all-gathers them and mgcg_combine_kernel adds them exactly in
// rank order, so the global value is again the correctly rounded exact sum, whatever the decomposition.
__device__ __forceinline__ bool mg_publish( CgState* S, dd_t v, int slot, double* direct )
{
    if ( S->world > 1 )
    {
        S->loc[2 * slot] = v.hi;
        S->loc[2 * slot + 1] = v.lo;
        return false;
    }
    *direct = v.hi + v.lo;
    return true;
}

// One smoothing sweep that also accumulates sum xo . b — on the fine level b is the CG residual r and the
// last sweep's xo is z, so this is the z.r of CG kernel 2 without another pass over z and r.
__global__ void __launch_bounds__( NT )
    mg_smooth_dot_kernel( const __grid_constant__ MgLevelDev L, double omega, const double* __restrict__ b,
                          const double* __restrict__ xi, double* __restrict__ xo, CgState* S, double* partials )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    dd_t rz = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, L, total )
    {
        double bv;
        const double z = cell_smooth( L, omega, b, xi, xo, cu, &bv );
        dd_acc( rz, z * bv );
    }
    dd_t vals[1] = { rz };
    if ( block_reduce_finalize<NT, 1>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
            mg_publish( S, vals[0], 0, &S->rz_new );
    }
}

__device__ __forceinline__ double mg_res( const MgLevelDev& F, const double* __restrict__ b,
                                          const double* __restrict__ x, int i, int j, int k )
{
    const long long o = mg_off( F, i, j, k );
    return b[o] - mg_Ax( F, x, o, mg_walls( F, i, j, k ) );
}

// coarse b = mean of the children's residuals, summed pairwise: x pairs, then y, then z
__device__ __forceinline__ void cell_restrict( const MgLevelDev& F, const MgLevelDev& C, const double* __restrict__ bf,
                                               const double* __restrict__ xf, double* __restrict__ bc, const MgCursor& cu )
{
    const int I = cu.i, J = cu.j, K = cu.k;
    const int i = 2 * I, j = 2 * J, k = F.cz * K;
    double s = ( mg_res( F, bf, xf, i, j, k ) + mg_res( F, bf, xf, i + 1, j, k ) ) +
               ( mg_res( F, bf, xf, i, j + 1, k ) + mg_res( F, bf, xf, i + 1, j + 1, k ) );
    if ( F.cz == 2 )
    {
        const double u = ( mg_res( F, bf, xf, i, j, k + 1 ) + mg_res( F, bf, xf, i + 1, j, k + 1 ) ) +
                         ( mg_res( F, bf, xf, i, j + 1, k + 1 ) + mg_res( F, bf, xf, i + 1, j + 1, k + 1 ) );
        s = ( s + u ) * 0.125;
    }
    else
        s = s * 0.25;
    bc[mg_off( C, I, J, K )] = s;
}

__device__ __forceinline__ void cell_prolong( const MgLevelDev& F, const MgLevelDev& C, double* __restrict__ xf,
                                              const double* __restrict__ ec, const MgCursor& cu )
{
    const int i = cu.i, j = cu.j, k = cu.k;
    const long long o = mg_off( F, i, j, k );
    xf[o] = xf[o] + ec[mg_off( C, i / 2, j / 2, k / F.cz )];
}

// Prolongation + correction + the first post-smoothing sweep in one pass (24 instead of 17 + 24 bytes per
// cell): x' = x + P e is formed on the fly for the cell and its six neighbours (neighbours off the block:
// the ghost zeros of x', which the separate prolongation never writes), then xo = x' + omega D^-1 (b - A x').
__device__ __forceinline__ double cell_prolong_smooth( const MgLevelDev& F, const MgLevelDev& C, double omega,
                                                       const double* __restrict__ b, const double* __restrict__ xi,
                                                       const double* __restrict__ ec, double* __restrict__ xo,
                                                       const MgCursor& cu, double* bv_out = nullptr )
{
    const int i = cu.i, j = cu.j, k = cu.k;
    const long long o = mg_off( F, i, j, k );
    const int I = i / 2, J = j / 2, K = k / F.cz;
    // Fast path (cells with all six neighbours inside the block and no wall: every layer but the outermost): the
    // parents of the neighbours sit at +-1 / +-sy / +-sz from the cell's own parent, chosen by the parity of the
    // index — one coarse offset instead of seven; the same sums and the same row as below.
    if ( i >= 1 && i < F.n[0] - 1 && j >= 1 && j < F.n[1] - 1 && k >= 1 && k < F.n[2] - 1 )
    {
        const long long oc = mg_off( C, I, J, K );
        const long long cxm = ( i & 1 ) ? 0 : -1, cxp = ( i & 1 ) ? 1 : 0;
        const long long cym = ( j & 1 ) ? 0 : -C.sy, cyp = ( j & 1 ) ? C.sy : 0;
        const long long czm = F.cz == 1 ? -C.sz : ( ( k & 1 ) ? 0 : -C.sz ), czp = F.cz == 1 ? C.sz : ( ( k & 1 ) ? C.sz : 0 );
        const double xc = xi[o] + ec[oc];
        const double xm = xi[o - 1] + ec[oc + cxm], xp = xi[o + 1] + ec[oc + cxp];
        const double ym = xi[o - F.sy] + ec[oc + cym], yp = xi[o + F.sy] + ec[oc + cyp];
        const double zm = xi[o - F.sz] + ec[oc + czm], zp = xi[o + F.sz] + ec[oc + czp];
        const double bv = b[o];
        const double res = bv - apply_row( F.diag[0], F.ns, xc, xm, xp, ym, yp, zm, zp );
        const double z = fma( omega * F.minv[0], res, xc );
        xo[o] = z;
        if ( bv_out )
            *bv_out = bv;
        return z;
    }
    const int w = mg_walls( F, i, j, k );
    const double xc = xi[o] + ec[mg_off( C, I, J, K )];
    // across a block interface: the ghost entries of x and of the coarse correction (both exchanged)
    const double xm = ( i > 0 || F.nlo[0] ) ? xi[o - 1] + ec[mg_off( C, mg_parent( i - 1, 2 ), J, K )] : 0.0;
    const double xp = ( i < F.n[0] - 1 || F.nhi[0] ) ? xi[o + 1] + ec[mg_off( C, mg_parent( i + 1, 2 ), J, K )] : 0.0;
    const double ym = ( j > 0 || F.nlo[1] ) ? xi[o - F.sy] + ec[mg_off( C, I, mg_parent( j - 1, 2 ), K )] : 0.0;
    const double yp = ( j < F.n[1] - 1 || F.nhi[1] ) ? xi[o + F.sy] + ec[mg_off( C, I, mg_parent( j + 1, 2 ), K )] : 0.0;
    const double zm = ( k > 0 || F.nlo[2] ) ? xi[o - F.sz] + ec[mg_off( C, I, J, mg_parent( k - 1, F.cz ) )] : 0.0;
    const double zp = ( k < F.n[2] - 1 || F.nhi[2] ) ? xi[o + F.sz] + ec[mg_off( C, I, J, mg_parent( k + 1, F.cz ) )] : 0.0;
    const double bv = b[o];
    const double res = bv - apply_row( F.diag[w], F.ns, xc, xm, xp, ym, yp, zm, zp );
    const double z = fma( omega * F.minv[w], res, xc );
    xo[o] = z;
    if ( bv_out )
        *bv_out = bv;
    return z;
}

__global__ void __launch_bounds__( NT )
    mg_restrict_kernel( const __grid_constant__ MgLevelDev F, const __grid_constant__ MgLevelDev C,
                        const double* __restrict__ bf, const double* __restrict__ xf, double* __restrict__ bc )
{
    const long long total = (long long)C.n[0] * C.n[1] * C.n[2];
    MG_GRID_CELLS( cu, C, total )
        cell_restrict( F, C, bf, xf, bc, cu );
}

__global__ void __launch_bounds__( NT )
    mg_prolong_kernel( const __grid_constant__ MgLevelDev F, const __grid_constant__ MgLevelDev C,
                       double* __restrict__ xf, const double* __restrict__ ec )
{
    const long long total = (long long)F.n[0] * F.n[1] * F.n[2];
    MG_GRID_CELLS( cu, F, total )
        cell_prolong( F, C, xf, ec, cu );
}

// DOT: also sum xo . b (see mg_smooth_dot_kernel), for cycles whose only post-smoothing sweep this is.
template <bool DOT>
__global__ void __launch_bounds__( NT )
    mg_prolong_smooth_kernel( const __grid_constant__ MgLevelDev F, const __grid_constant__ MgLevelDev C, double omega,
                              const double* __restrict__ b, const double* __restrict__ xi,
                              const double* __restrict__ ec, double* __restrict__ xo, CgState* S, double* partials )
{
    const long long total = (long long)F.n[0] * F.n[1] * F.n[2];
    dd_t rz = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, F, total )
    {
        double bv;
        const double z = cell_prolong_smooth( F, C, omega, b, xi, ec, xo, cu, &bv );
        if ( DOT )
            dd_acc( rz, z * bv );
    }
    if ( DOT )
    {
        dd_t vals[1] = { rz };
        if ( block_reduce_finalize<NT, 1>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
        {
            if ( threadIdx.x == 0 )
                mg_publish( S, vals[0], 0, &S->rz_new );
        }
    }
}

// ---- the coarse end of the cycle in ONE kernel ------------------------------------------------------------
// From the first level with at most MG_COARSE_CELLS cells down to the coarsest and back up, a single CTA runs
// every sweep / transfer with a block barrier in between ("mg_coarse_kernel" tuning key; one block only).  The
// arrays of these levels (<= 3 x 46 KB each, smaller below) live in the L2; what is saved is ~30 of the ~60
// launches of a cycle, which is what bounds the cycle on grids up to ~128^3.  Same element-wise operations in
// the same order as the launch-per-operation form: bit-identical.
#define MG_COARSE_CELLS 4096
#define MG_COARSE_LEVELS 8
struct MgCoarseArgs
{
    int nlev, nu1, nu2, nuc;
    double wpre[8], wpost[8], wc; // damping per sweep (nu1, nu2 <= 8 here)
    MgLevelDev lv[MG_COARSE_LEVELS];
    double* b[MG_COARSE_LEVELS];
    double* x[MG_COARSE_LEVELS][2];
};

__global__ void __launch_bounds__( NT )
    mg_coarse_cycle_kernel( const __grid_constant__ MgCoarseArgs a )
{
    __shared__ int cur[MG_COARSE_LEVELS];
    const int tid = threadIdx.x;
    // down: pre-smoothing (the coarsest level: its nuc sweeps), residual + restriction
    for ( int l = 0; l < a.nlev; ++l )
    {
        const MgLevelDev& L = a.lv[l];
        const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
        const bool coarsest = l == a.nlev - 1;
        const int sweeps = coarsest ? a.nuc : a.nu1;
        int c, done;
        if ( sweeps >= 2 )
        {
            MG_CTA_CELLS( cu, L, total )
                cell_smooth02( L, coarsest ? a.wc : a.wpre[0], coarsest ? a.wc : a.wpre[1], a.b[l], a.x[l][1], cu );
            c = 1;
            done = 2;
        }
        else
        {
            MG_CTA_CELLS( cu, L, total )
                cell_smooth0( L, coarsest ? a.wc : a.wpre[0], a.b[l], a.x[l][0], cu );
            c = 0;
            done = 1;
        }
        __syncthreads();
        for ( ; done < sweeps; ++done )
        {
            MG_CTA_CELLS( cu, L, total )
                cell_smooth( L, coarsest ? a.wc : a.wpre[done], a.b[l], a.x[l][c], a.x[l][1 - c], cu );
            c = 1 - c;
            __syncthreads();
        }
        if ( tid == 0 )
            cur[l] = c;
        if ( l + 1 < a.nlev )
        {
            const MgLevelDev& C = a.lv[l + 1];
            const long long ctotal = (long long)C.n[0] * C.n[1] * C.n[2];
            MG_CTA_CELLS( cu, C, ctotal )
                cell_restrict( L, C, a.b[l], a.x[l][c], a.b[l + 1], cu );
        }
        __syncthreads();
    }
    // up: prolongation + correction (+ first post-smoothing sweep), remaining post-smoothing sweeps
    for ( int l = a.nlev - 2; l >= 0; --l )
    {
        const MgLevelDev& L = a.lv[l];
        const MgLevelDev& C = a.lv[l + 1];
        const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
        int c = cur[l];
        const double* ec = a.x[l + 1][cur[l + 1]];
        if ( a.nu2 == 0 )
        {
            MG_CTA_CELLS( cu, L, total )
                cell_prolong( L, C, a.x[l][c], ec, cu );
            __syncthreads();
            continue;
        }
        MG_CTA_CELLS( cu, L, total )
            cell_prolong_smooth( L, C, a.wpost[0], a.b[l], a.x[l][c], ec, a.x[l][1 - c], cu );
        c = 1 - c;
        __syncthreads();
        for ( int s2 = 1; s2 < a.nu2; ++s2 )
        {
            MG_CTA_CELLS( cu, L, total )
                cell_smooth( L, a.wpost[s2], a.b[l], a.x[l][c], a.x[l][1 - c], cu );
            c = 1 - c;
            __syncthreads();
        }
        if ( tid == 0 )
            cur[l] = c;
        __syncthreads();
    }
}

// ---- the CG around it (Cajita::ReferenceConjugateGradient::solve with a general M^-1) ---------------
// start of solve: x0 = 0, r0 = b, sum r0^2
__global__ void __launch_bounds__( NT )
    mgcg_init_kernel( const __grid_constant__ MgLevelDev L, const double* __restrict__ b, double* __restrict__ x,
                      double* __restrict__ r, CgState* S, double* partials, int fixed )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    dd_t rr = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const long long o = mg_off( L, i, j, k );
        const double bv = b[o];
        x[o] = 0.0;
        r[o] = bv;
        dd_acc( rr, bv * bv );
    }
    dd_t vals[1] = { rr };
    if ( block_reduce_finalize<NT, 1>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
        {
            mg_publish( S, vals[0], 0, &S->rr );
            S->iter = 0;
            S->done = 0;
            S->fixed = fixed;
        }
    }
}

__global__ void mgcg_check0_kernel( CgState* S, double tol, int stop_rel )
{
    const double bnorm = sqrt( S->rr );
    S->bnorm = bnorm;
    S->thresh = stop_rel ? tol * bnorm : tol;
    if ( !S->fixed && bnorm <= S->thresh )
        S->done = 1;
}

// kernel 2's reduction: sum z.r
__global__ void __launch_bounds__( NT )
    mgcg_dot_kernel( const __grid_constant__ MgLevelDev L, const double* __restrict__ z,
                     const double* __restrict__ r, CgState* S, double* partials )
{
    if ( S->done )
        return;
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    dd_t rz = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const long long o = mg_off( L, i, j, k );
        dd_acc( rz, z[o] * r[o] );
    }
    dd_t vals[1] = { rz };
    if ( block_reduce_finalize<NT, 1>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
            mg_publish( S, vals[0], 0, &S->rz_new );
    }
}

// kernel 3: p = z + beta p   (first: p0 = z0); the stencil kernel that follows sets rz_old = rz_new
__global__ void __launch_bounds__( NT )
    mgcg_pupdate_kernel( const __grid_constant__ MgLevelDev L, const double* __restrict__ z, double* __restrict__ p,
                         const CgState* S, int first )
{
    if ( S->done )
        return;
    const double beta = first ? 0.0 : S->rz_new / S->rz_old;
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const long long o = mg_off( L, i, j, k );
        p[o] = first ? z[o] : fma( beta, p[o], z[o] );
    }
}

// after kernel 1's sum r^2 is known globally: history, iteration count, stopping test
__device__ __forceinline__ void mgcg_bookkeeping( CgState* S )
{
    const double resid = sqrt( S->rr );
    const int it = S->iter;
    if ( it < CFB_HIST_MAX )
        S->hist[it] = resid;
    S->iter = it + 1;
    if ( !S->fixed && resid <= S->thresh )
        S->done = 1;
}

// kernel 1: x += alpha p, r -= alpha q, sum r^2, then the iteration's bookkeeping and stopping test
__global__ void __launch_bounds__( NT )
    mgcg_axpy_kernel( const __grid_constant__ MgLevelDev L, const double* __restrict__ p,
                      const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, CgState* S,
                      double* partials )
{
    if ( S->done )
        return;
    const double alpha = S->rz_old / S->pAp;
    const double nalpha = -alpha;
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    dd_t rr = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const long long o = mg_off( L, i, j, k );
        x[o] = fma( alpha, p[o], x[o] );
        const double rv = fma( nalpha, q[o], r[o] );
        r[o] = rv;
        dd_acc( rr, rv * rv );
    }
    dd_t vals[1] = { rr };
    if ( block_reduce_finalize<NT, 1>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 && mg_publish( S, vals[0], 0, &S->rr ) )
            mgcg_bookkeeping( S );
    }
}

// null-space pinning: sum d_i x_i and sum d_i -> S->loc[4..5]
__global__ void __launch_bounds__( NT )
    mgcg_nullsum_kernel( const __grid_constant__ MgLevelDev L, const double* __restrict__ x, CgState* S,
                         double* partials )
{
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    dd_t dx = { 0.0, 0.0 }, ds = { 0.0, 0.0 };
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const double dg = L.diag[mg_walls( L, i, j, k )];
        dd_acc( dx, dg * x[mg_off( L, i, j, k )] );
        dd_acc( ds, dg );
    }
    dd_t vals[2] = { dx, ds };
    if ( block_reduce_finalize<NT, 2>( vals, partials, CFB_MAX_PARTIALS, &S->ticket[0] ) )
    {
        if ( threadIdx.x == 0 )
        {
            mg_publish( S, vals[0], 0, &S->loc[4] );
            mg_publish( S, vals[1], 1, &S->loc[5] );
        }
    }
}

__global__ void __launch_bounds__( NT )
    mgcg_shift_kernel( const __grid_constant__ MgLevelDev L, double* __restrict__ x, const CgState* S )
{
    const double shift = S->loc[4] / S->loc[5];
    const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
    MG_GRID_CELLS( cu, L, total )
    {
        const int i = cu.i, j = cu.j, k = cu.k;
        const long long o = mg_off( L, i, j, k );
        x[o] = x[o] - shift;
    }
}

// Several blocks: exact combination, in rank order, of the all-gathered local double-doubles (S->gath,
// nv values per rank).  what: 0 = r0.r0 of the start, 1 = z.r, 2 = r.r of an iteration (+ bookkeeping),
// 3 = the two null-space sums.
__global__ void mgcg_combine_kernel( CgState* S, int what )
{
    if ( what != 0 && what != 3 && S->done )
        return;
    const int nv = what == 3 ? 2 : 1;
    dd_t acc[2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
    for ( int r = 0; r < S->world; ++r )
        for ( int v = 0; v < nv; ++v )
        {
            dd_t w = { S->gath[( r * nv + v ) * 2], S->gath[( r * nv + v ) * 2 + 1] };
            acc[v] = dd_add( acc[v], w );
        }
    const double v0 = acc[0].hi + acc[0].lo;
    if ( what == 0 )
        S->rr = v0;
    else if ( what == 1 )
        S->rz_new = v0;
    else if ( what == 2 )
    {
        S->rr = v0;
        mgcg_bookkeeping( S );
    }
    else
    {
        S->loc[4] = v0;
        S->loc[5] = acc[1].hi + acc[1].lo;
    }
}

// One-layer face exchange of a level array: PACK copies my boundary layers into the send buffers, !PACK
// scatters the received layers into my ghost layers.  blockIdx.y = face.
struct MgFaces
{
    int nface;
    int dim[6], src[6], dst[6]; // normal direction; owned index of the layer sent; ghost index filled
    double* buf[6];
};

template <bool PACK>
__global__ void __launch_bounds__( NT )
    mg_face_kernel( const __grid_constant__ MgLevelDev L, const __grid_constant__ MgFaces a, double* __restrict__ arr )
{
    const int f = blockIdx.y;
    if ( f >= a.nface )
        return;
    const int d = a.dim[f];
    const int e0 = d == 0 ? L.n[1] : L.n[0];               // fastest tangential extent
    const int e1 = d == 2 ? L.n[1] : L.n[2];               // slowest tangential extent
    const long long total = (long long)e0 * e1;
    const int fix = PACK ? a.src[f] : a.dst[f];
    for ( long long t = blockIdx.x * (long long)NT + threadIdx.x; t < total; t += (long long)gridDim.x * NT )
    {
        const int u = (int)( t % e0 ), v = (int)( t / e0 );
        const int i = d == 0 ? fix : u;
        const int j = d == 0 ? u : ( d == 1 ? fix : v );
        const int k = d == 2 ? fix : v;
        const long long o = mg_off( L, i, j, k );
        if ( PACK )
            a.buf[f][t] = arr[o];
        else
            arr[o] = a.buf[f][t];
    }
}

// The same exchange over NVLink peer memory, one kernel: (1) my boundary layers go straight into the
// neighbours' ghost layers (same level geometry on every rank); (2) system-scope fence, ticket; the last block
// (3) tells every neighbour "my stores of exchange #seq are done" in its mailbox and (4) waits until every
// neighbour has said the same in mine (bounded spin).  The protocol of cg_xchg_kernel (halo.cu), between face
// neighbours only.  Consecutive exchanges never target the same array, so a neighbour that is still reading the
// ghosts of the previous exchange is never overwritten: it cannot be more than one exchange behind.
struct MgXchg
{
    int nface;
    int dim[6], src[6], dst[6];
    double* peer[6];                    // the neighbour's copy of the array being exchanged
    unsigned long long* tell[6];        // &mail[neighbour]->mseq[me]
    const unsigned long long* hear[6];  // &mail_self->mseq[neighbour]
    unsigned long long seq;
    unsigned int* ticket;
    CgState* S;
    long long timeout_cycles;
};

__device__ __forceinline__ unsigned long long mg_ld_volatile_u64( const unsigned long long* p )
{
    return *reinterpret_cast<const volatile unsigned long long*>( p );
}

__global__ void __launch_bounds__( NT )
    mg_xchg_kernel( const __grid_constant__ MgLevelDev L, const __grid_constant__ MgXchg a, const double* arr )
{
    for ( int f = 0; f < a.nface; ++f )
    {
        const int d = a.dim[f];
        const int e0 = d == 0 ? L.n[1] : L.n[0];
        const int e1 = d == 2 ? L.n[1] : L.n[2];
        const long long total = (long long)e0 * e1;
        double* dst = a.peer[f];
        for ( long long t = blockIdx.x * (long long)NT + threadIdx.x; t < total; t += (long long)gridDim.x * NT )
        {
            const int u = (int)( t % e0 ), v = (int)( t / e0 );
            const int i = d == 0 ? a.src[f] : u, j = d == 0 ? u : ( d == 1 ? a.src[f] : v ), k = d == 2 ? a.src[f] : v;
            const int I = d == 0 ? a.dst[f] : i, J = d == 1 ? a.dst[f] : j, K = d == 2 ? a.dst[f] : k;
            dst[mg_off( L, I, J, K )] = arr[mg_off( L, i, j, k )];
        }
    }
    __threadfence_system();
    __shared__ bool s_last;
    __syncthreads();
    if ( threadIdx.x == 0 )
        s_last = atomicAdd( a.ticket, 1u ) == gridDim.x - 1;
    __syncthreads();
    if ( !s_last )
        return;
    __threadfence_system();
    if ( threadIdx.x == 0 )
        *a.ticket = 0u;
    if ( threadIdx.x < a.nface )
    {
        *reinterpret_cast<volatile unsigned long long*>( a.tell[threadIdx.x] ) = a.seq;
        if ( !a.S->xerror )
        {
            const long long t0 = clock64();
            while ( mg_ld_volatile_u64( a.hear[threadIdx.x] ) < a.seq )
                if ( clock64() - t0 > a.timeout_cycles )
                {
                    a.S->xerror = 1;
                    break;
                }
        }
    }
    __threadfence_system();
}

// ---- host side -------------------------------------------------------------------------------------
inline int grid_for( const cfb_ctx* c, long long cells )
{
    long long b = ( cells + NT - 1 ) / NT;
    long long cap = (long long)c->sm_count * 8;
    if ( cap > CFB_MAX_PARTIALS )
        cap = CFB_MAX_PARTIALS;
    return (int)( b < 1 ? 1 : ( b > cap ? cap : b ) );
}

// (Launch order: one block per 256 cells launched in order, instead of this capped grid whose threads stride through the
// level, was measured and is slower — 9.88 vs 9.48 ms per MG-PCG iteration at 512^3, profiles/r2_mg_launch_order.log.)
void mg_free( cfb_ctx* c )
{
    MgStage* m = c->mg;
    if ( !m )
        return;
    if ( m->peer )
    {
        // close my mappings of the neighbours' arrays; nobody frees before everybody has closed (a collective:
        // every rank rebuilds / destroys its preconditioner at the same point)
        cudaStreamSynchronize( c->stream );
        for ( double* p : m->peer_arrays )
            if ( p )
            {
                cudaIpcCloseMemHandle( p );
                for ( void*& q : c->ipc_opened )
                    if ( q == p )
                        q = nullptr;
            }
        c->ipc_opened.erase( std::remove( c->ipc_opened.begin(), c->ipc_opened.end(), nullptr ), c->ipc_opened.end() );
        halo_allreduce( c, &c->d_state->gath[0], 1 );
        cudaStreamSynchronize( c->stream );
    }
    for ( auto& L : m->lv )
    {
        if ( L.owns_b && L.b )
            cudaFree( L.b );
        for ( double* p : L.x )
            if ( p )
                cudaFree( p );
    }
    if ( m->ev_poll )
        cudaEventDestroy( m->ev_poll );
    for ( cudaGraphExec_t g : m->graph )
        if ( g )
            cudaGraphExecDestroy( g );
    delete m;
    c->mg = nullptr;
}

int mg_build( cfb_ctx* c, int nu1, int nu2, int nuc, double omega, int max_levels )
{
    mg_free( c );
    MgStage* m = new MgStage();
    c->mg = m;
    m->max_levels = max_levels;
    m->use_graph = c->mg_graph;
    m->use_coarse = c->mg_coarse;
    m->nu1 = nu1;
    m->nu2 = nu2;
    m->nuc = nuc;
    const Geo& g = c->g;
    const int D = g.D;
    m->omega = omega;
    {
        // damping per sweep; literals so that every implementation uses the same bits:
        // 1 / ( 1.2 + 0.8 cos( (2k - 1) pi / (2 nu) ) ), k = 1..nu
        static const double cheb[5][4] = { { 0, 0, 0, 0 },
                                           { 0, 0, 0, 0 },
                                           { 0.5663522991524661, 1.576504843704677, 0, 0 },
                                           { 0.5283121635129678, 0.8333333333333334, 1.9716878364870327, 0 },
                                           { 0.515702196925985, 0.6639459287266942, 1.118751870515935, 2.169685110214365 } };
        const double fixed = omega > 0.0 ? omega : ( D == 3 ? 6.0 / 7.0 : 0.8 );
        m->wc = fixed;
        auto fill = [&]( std::vector<double>& w, int nu, bool reverse ) {
            w.assign( nu > 0 ? nu : 0, fixed );
            // only for the symmetric cycle (nu1 == nu2): the big per-sweep factors are harmless as a product,
            // not one by one, and CG needs M symmetric positive definite
            if ( omega <= 0.0 && D == 3 && nu >= 2 && nu <= 4 && nu1 == nu2 )
                for ( int q = 0; q < nu; ++q )
                    w[q] = cheb[nu][reverse ? nu - 1 - q : q];
        };
        fill( m->wpre, nu1, false );
        fill( m->wpost, nu2, true );
    }
    m->singular = true;
    for ( int d = 0; d < D; ++d )
        m->singular = m->singular && g.bt[d] == CFB_SOLID && g.bt[3 + d] == CFB_SOLID;
    CFB_CUDA( c, cudaEventCreateWithFlags( &m->ev_poll, cudaEventDisableTiming ) );
    int n[3] = { g.n[0], g.n[1], D == 3 ? g.n[2] : 1 };
    double scale = c->op.scale;
    for ( int l = 0; l < 16; ++l )
    {
        m->lv.emplace_back();
        MgLevelHost& H = m->lv.back();
        MgLevelDev& L = H.d;
        for ( int d = 0; d < 3; ++d )
        {
            L.n[d] = n[d];
            L.slo[d] = d < D ? ( g.lo_bd[d] && g.bt[d] == CFB_SOLID ) : 1;
            L.shi[d] = d < D ? ( g.hi_bd[d] && g.bt[3 + d] == CFB_SOLID ) : 1;
            L.nlo[d] = ( d < D && c->cfg.use_nccl && c->nbr[2 * d] >= 0 ) ? 1 : 0;
            L.nhi[d] = ( d < D && c->cfg.use_nccl && c->nbr[2 * d + 1] >= 0 ) ? 1 : 0;
        }
        L.cz = D == 3 ? 2 : 1;
        L.ns = -1.0 * scale;
        for ( int cnt = 0; cnt < 8; ++cnt )
        {
            // the reference's own sequence (src/VelocityCorrector.hpp:137, BoundaryConditions.hpp:56-97); in
            // 2-D the two z walls of the single plane are part of the count (see Geo)
            double dgl;
            if ( D == 3 )
            {
                dgl = 6.0 * scale;
                for ( int i = 0; i < cnt; ++i )
                    dgl -= scale;
            }
            else
            {
                dgl = 4.0 * scale;
                for ( int i = 0; i < cnt - 2; ++i )
                    dgl -= scale;
            }
            L.diag[cnt] = dgl;
            L.minv[cnt] = 1.0 / dgl;
        }
        H.cells = (long long)n[0] * n[1] * n[2];
        size_t elems;
        if ( l == 0 )
        {
            // the fine level lives in the layout of the CG vectors: b is the residual itself
            L.sy = g.sy;
            L.sz = g.sz;
            L.origin = g.origin;
            elems = (size_t)g.total;
            H.b = nullptr; // bound per solve to cg_r
        }
        else
        {
            L.sy = n[0] + 2;
            L.sz = L.sy * ( n[1] + 2 );
            L.origin = L.sz + L.sy + 1;
            elems = (size_t)L.sz * ( n[2] + 2 );
            CFB_CUDA( c, cudaMalloc( &H.b, elems * sizeof( double ) ) );
            CFB_CUDA( c, cudaMemsetAsync( H.b, 0, elems * sizeof( double ), c->stream ) );
            H.owns_b = true;
        }
        for ( int q = 0; q < 2; ++q )
        {
            // ghosts are never written: they stay zero, which is what the operator reads off the domain
            CFB_CUDA( c, cudaMalloc( &H.x[q], elems * sizeof( double ) ) );
            CFB_CUDA( c, cudaMemsetAsync( H.x[q], 0, elems * sizeof( double ), c->stream ) );
        }
        // every block has the same extents here (cfb_set_preconditioner checks), so all ranks stop together
        bool can = m->max_levels <= 0 || l + 1 < m->max_levels;
        for ( int d = 0; d < D; ++d )
            can = can && n[d] % 2 == 0 && n[d] / 2 >= 2;
        if ( !can )
            break;
        for ( int d = 0; d < D; ++d )
            n[d] /= 2;
        scale = scale * 0.25;
    }
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    if ( c->cfg.use_nccl && c->peer_ok && c->use_peer )
    {
        // level 0: the two x arrays (its right-hand side is cg_r, mapped since start-up); below: b and both x
        for ( size_t l = 0; l < m->lv.size(); ++l )
        {
            if ( l > 0 )
                m->own_arrays.push_back( m->lv[l].b );
            m->own_arrays.push_back( m->lv[l].x[0] );
            m->own_arrays.push_back( m->lv[l].x[1] );
        }
        int rc = peer_map_arrays( c, (int)m->own_arrays.size(), m->own_arrays.data(), m->peer_arrays, &m->peer );
        if ( rc )
            return rc;
    }
    if ( int rc = mg_tma_prepare( c ) ) // the partial-sum scratch of the fine level's TMA sweeps
        return rc;
    m->coarse_start = -1;
    for ( int l = 0; l < (int)m->lv.size(); ++l )
        if ( m->lv[l].cells <= MG_COARSE_CELLS && (int)m->lv.size() - l <= MG_COARSE_LEVELS )
        {
            m->coarse_start = l;
            break;
        }
    return CFB_OK;
}

// Several blocks: refresh the one-cell ghost layers of a level array from the face neighbours (pack kernel,
// one grouped NCCL send/recv, unpack kernel; 7-point operator and cell-local transfers: faces only).
int mg_exchange( cfb_ctx* c, MgLevelHost& H, double* arr, int* launches )
{
    if ( !c->cfg.use_nccl )
        return CFB_OK;
    const MgLevelDev& L = H.d;
    MgFaces pk{}, up{};
    size_t counts[6] = { 0, 0, 0, 0, 0, 0 };
    long long mx = 1;
    for ( int s = 0; s < 6; ++s )
    {
        if ( c->nbr[s] < 0 )
            continue;
        const int d = s / 2, side = s % 2;
        const long long cnt = (long long)L.n[0] * L.n[1] * L.n[2] / L.n[d];
        counts[s] = (size_t)cnt;
        mx = cnt > mx ? cnt : mx;
        pk.dim[pk.nface] = up.dim[up.nface] = d;
        pk.src[pk.nface] = side == 0 ? 0 : L.n[d] - 1; // my first layer -> low neighbour, my last -> high
        up.dst[up.nface] = side == 0 ? -1 : L.n[d];    // its last layer -> my low ghost, its first -> my high
        pk.buf[pk.nface++] = c->d_halo_send[s];
        up.buf[up.nface++] = c->d_halo_recv[s];
    }
    if ( pk.nface == 0 )
        return CFB_OK;
    // Two exchanges in a row on the same array (only V(nu1, 0) cycles with a one-sweep coarsest level do that)
    // would let a fast rank overwrite ghosts a neighbour is still reading: those go through NCCL, whose
    // receives are ordered behind the neighbour's own stream.
    const bool repeat = c->mg->last_xchg == arr;
    c->mg->last_xchg = arr;
    if ( c->mg->peer && c->use_peer && !repeat )
    {
        MgStage* m = c->mg;
        // which exported array is this?  (level 0's right-hand side is cg_r: the start-up mapping)
        int idx = -1;
        for ( size_t a = 0; a < m->own_arrays.size(); ++a )
            if ( m->own_arrays[a] == arr )
                idx = (int)a;
        if ( idx < 0 && arr != c->cg_r )
            return cfb_fail( c, CFB_ERR_INVALID, "multigrid exchange of an array that was not exported" );
        MgXchg x{};
        int f = 0;
        for ( int s = 0; s < 6; ++s )
        {
            if ( c->nbr[s] < 0 )
                continue;
            const int d = s / 2, side = s % 2;
            x.dim[f] = d;
            x.src[f] = side == 0 ? 0 : L.n[d] - 1; // my first layer -> the low neighbour's high ghost
            x.dst[f] = side == 0 ? L.n[d] : -1;    // (index n there), my last -> the high neighbour's low ghost
            x.peer[f] = idx >= 0 ? m->peer_arrays[(size_t)s * m->own_arrays.size() + idx] : c->peer_r[s];
            x.tell[f] = &c->mail[c->nbr[s]]->mseq[c->cfg.world_rank];
            x.hear[f] = &c->mail_self->mseq[c->nbr[s]];
            ++f;
        }
        x.nface = f;
        x.seq = ++m->xseq;
        x.ticket = c->d_xticket;
        x.S = c->d_state;
        x.timeout_cycles = 20000000000ll;
        long long gx = ( mx + 4 * NT - 1 ) / ( 4 * NT );
        const long long gcap = 2ll * c->sm_count;
        mg_xchg_kernel<<<(int)( gx < 1 ? 1 : ( gx > gcap ? gcap : gx ) ), NT, 0, c->stream>>>( L, x, arr );
        *launches += 1;
        return CFB_OK;
    }
    long long bx = ( mx + NT - 1 ) / NT;
    const long long cap = (long long)c->sm_count * 4;
    dim3 grid( (unsigned)( bx > cap ? cap : bx ), (unsigned)pk.nface );
    mg_face_kernel<true><<<grid, NT, 0, c->stream>>>( L, pk, arr );
    int rc = halo_sendrecv_slots( c, counts, c->stream );
    if ( rc )
        return rc;
    mg_face_kernel<false><<<grid, NT, 0, c->stream>>>( L, up, arr );
    *launches += 2;
    return CFB_OK;
}

// Several blocks: all-gather the local double-doubles left in S->loc by the last reduction and combine
// them exactly (see mgcg_combine_kernel for `what`).
int mg_global_sum( cfb_ctx* c, int what, int* launches )
{
    if ( !c->cfg.use_nccl )
        return CFB_OK;
    int rc = halo_allgather( c, &c->d_state->loc[0], c->d_state->gath, what == 3 ? 4 : 2 );
    if ( rc )
        return rc;
    mgcg_combine_kernel<<<1, 1, 0, c->stream>>>( c->d_state, what );
    *launches += 1;
    return CFB_OK;
}

#define MG_TRY( expr )                                                                             \
    do                                                                                             \
    {                                                                                              \
        int _rc = ( expr );                                                                        \
        if ( _rc )                                                                                 \
            return _rc;                                                                            \
    } while ( 0 )

// the first `sweeps` (>= 1) smoothing sweeps of a level from a zero initial guess; w[s] = damping of sweep s
// (nullptr: the coarsest level's fixed damping)
// the level's operator constants in the form the TMA stencil kernels take
static OpConst level_op( const MgLevelDev& L )
{
    OpConst op{};
    op.neg_scale = L.ns;
    op.scale = -L.ns;
    for ( int i = 0; i < 8; ++i )
    {
        op.diag[i] = L.diag[i];
        op.minv[i] = L.minv[i];
    }
    return op;
}

int launch_presmooth( cfb_ctx* c, MgLevelHost& H, int sweeps, const double* w, int* n )
{
    const int grid = grid_for( c, H.cells );
    const double wc = c->mg->wc;
    int done;
    const bool fine = &H == &c->mg->lv[0];
    if ( sweeps >= 2 )
    {
        MG_TRY( mg_exchange( c, H, H.b, n ) ); // the fused pair reads b of the six neighbours
        // fine level of a 3-D run: the TMA z-march (kernels_stencil.cu MODE 4), same statements
        if ( !fine || launch_mg_smooth02_tma( c, level_op( H.d ), w ? w[0] : wc, w ? w[1] : wc, H.b, H.x[0], H.x[1] ) < 0 )
            mg_smooth02_kernel<<<grid, NT, 0, c->stream>>>( H.d, w ? w[0] : wc, w ? w[1] : wc, H.b, H.x[1] );
        H.cur = 1; // where the unfused pair of sweeps leaves its result
        done = 2;
    }
    else
    {
        mg_smooth0_kernel<<<grid, NT, 0, c->stream>>>( H.d, w ? w[0] : wc, H.b, H.x[0] );
        H.cur = 0;
        done = 1;
    }
    *n += 1;
    for ( ; done < sweeps; ++done )
    {
        MG_TRY( mg_exchange( c, H, H.x[H.cur], n ) );
        if ( !fine || launch_mg_smooth_tma( c, level_op( H.d ), w ? w[done] : wc, H.b, H.x[0], H.x[1], H.cur, 0 ) < 0 )
            mg_smooth_kernel<<<grid, NT, 0, c->stream>>>( H.d, w ? w[done] : wc, H.b, H.x[H.cur], H.x[1 - H.cur] );
        H.cur = 1 - H.cur;
        *n += 1;
    }
    return CFB_OK;
}

// One V-cycle on level l; *n counts the launches.  `dot` (fine level only, from the CG): the sweep that
// produces the level's result also accumulates sum z.r; *dotted tells whether it did (it cannot when there
// is no post-smoothing sweep).
int vcycle( cfb_ctx* c, int l, bool dot, bool* dotted, int* n )
{
    MgStage* m = c->mg;
    MgLevelHost& H = m->lv[l];
    const bool last = l + 1 == (int)m->lv.size();
    const int grid = grid_for( c, H.cells );
    if ( dotted )
        *dotted = false;
    if ( m->use_coarse && !c->cfg.use_nccl && l == m->coarse_start && m->nu1 <= 8 && m->nu2 <= 8 )
    {
        // the rest of the cycle, down to the coarsest level and back up to this one, in one single-CTA kernel
        MgCoarseArgs a{};
        a.nlev = (int)m->lv.size() - l;
        a.nu1 = m->nu1;
        a.nu2 = m->nu2;
        a.nuc = m->nuc;
        a.wc = m->wc;
        for ( int q = 0; q < m->nu1; ++q )
            a.wpre[q] = m->wpre[q];
        for ( int q = 0; q < m->nu2; ++q )
            a.wpost[q] = m->wpost[q];
        for ( int q = 0; q < a.nlev; ++q )
        {
            MgLevelHost& Q = m->lv[l + q];
            a.lv[q] = Q.d;
            a.b[q] = Q.b;
            a.x[q][0] = Q.x[0];
            a.x[q][1] = Q.x[1];
        }
        mg_coarse_cycle_kernel<<<1, NT, 0, c->stream>>>( a );
        *n += 1;
        // where the kernel leaves this level's result: the same ping-pong sequence as the launches below
        const int sweeps = last ? m->nuc : m->nu1;
        int cur = sweeps >= 2 ? 1 : 0;
        cur ^= ( sweeps - ( sweeps >= 2 ? 2 : 1 ) ) & 1;
        if ( !last )
            cur ^= m->nu2 & 1;
        H.cur = cur;
        return CFB_OK;
    }
    MG_TRY( launch_presmooth( c, H, last ? m->nuc : m->nu1, last ? nullptr : m->wpre.data(), n ) );
    if ( last )
        return CFB_OK;
    MgLevelHost& C = m->lv[l + 1];
    MG_TRY( mg_exchange( c, H, H.x[H.cur], n ) ); // the residual reads x of the six neighbours
    mg_restrict_kernel<<<grid_for( c, C.cells ), NT, 0, c->stream>>>( H.d, C.d, H.b, H.x[H.cur], C.b );
    *n += 1;
    MG_TRY( vcycle( c, l + 1, false, nullptr, n ) );
    if ( m->nu2 == 0 )
    {
        mg_prolong_kernel<<<grid, NT, 0, c->stream>>>( H.d, C.d, H.x[H.cur], C.x[C.cur] );
        *n += 1;
        return CFB_OK;
    }
    // x[cur] still has the ghosts exchanged for the restriction; the coarse correction needs its own
    MG_TRY( mg_exchange( c, C, C.x[C.cur], n ) );
    const bool dot_here = dot && m->nu2 == 1;
    // fine level of a 3-D run: the TMA z-march (kernels_stencil.cu MODE 5), same statements
    if ( l == 0 && launch_mg_prolong_smooth_tma( c, level_op( H.d ), m->wpost[0], H.b, H.x[0], H.x[1], H.cur, C.x[C.cur], C.d.sy, C.d.sz,
                                                 C.d.n, dot_here ? 1 : 0 ) >= 0 )
        ;
    else if ( dot_here )
        mg_prolong_smooth_kernel<true><<<grid, NT, 0, c->stream>>>( H.d, C.d, m->wpost[0], H.b, H.x[H.cur], C.x[C.cur],
                                                                   H.x[1 - H.cur], c->d_state, c->d_partials );
    else
        mg_prolong_smooth_kernel<false><<<grid_for( c, H.cells ), NT, 0, c->stream>>>( H.d, C.d, m->wpost[0], H.b, H.x[H.cur],
                                                                                         C.x[C.cur], H.x[1 - H.cur], c->d_state,
                                                                                         c->d_partials );
    H.cur = 1 - H.cur;
    *n += 1;
    for ( int s = 1; s < m->nu2; ++s )
    {
        MG_TRY( mg_exchange( c, H, H.x[H.cur], n ) );
        const bool dot_now = dot && s == m->nu2 - 1;
        if ( l == 0 && launch_mg_smooth_tma( c, level_op( H.d ), m->wpost[s], H.b, H.x[0], H.x[1], H.cur, dot_now ? 1 : 0 ) >= 0 )
            ; // fine level of a 3-D run: the TMA z-march (kernels_stencil.cu MODE 3), same statements
        else if ( dot_now )
            mg_smooth_dot_kernel<<<grid, NT, 0, c->stream>>>( H.d, m->wpost[s], H.b, H.x[H.cur], H.x[1 - H.cur], c->d_state,
                                                             c->d_partials );
        else
            mg_smooth_kernel<<<grid_for( c, H.cells ), NT, 0, c->stream>>>( H.d, m->wpost[s], H.b, H.x[H.cur], H.x[1 - H.cur] );
        H.cur = 1 - H.cur;
        *n += 1;
    }
    if ( dotted )
        *dotted = dot;
    return CFB_OK;
}

// The fine-level cycle as the callers use it: plain launches, or (mg_graph, one block) the captured graph.
int run_cycle( cfb_ctx* c, bool dot, bool* dotted, int* n )
{
    MgStage* m = c->mg;
    if ( !m->use_graph || c->cfg.use_nccl )
        return vcycle( c, 0, dot, dotted, n );
    const int f = dot ? 1 : 0;
    if ( !m->graph[f] )
    {
        cudaGraph_t g = nullptr;
        int launches = 0;
        bool d = false;
        CFB_CUDA( c, cudaStreamBeginCapture( c->stream, cudaStreamCaptureModeThreadLocal ) );
        int rc = vcycle( c, 0, dot, &d, &launches );
        cudaError_t e = cudaStreamEndCapture( c->stream, &g );
        if ( rc )
            return rc;
        if ( e != cudaSuccess || !g )
            return cfb_fail( c, CFB_ERR_CUDA, std::string( "V-cycle graph capture failed: " ) + cudaGetErrorString( e ) );
        e = cudaGraphInstantiate( &m->graph[f], g, 0 );
        cudaGraphDestroy( g );
        if ( e != cudaSuccess )
            return cfb_fail( c, CFB_ERR_CUDA, std::string( "cudaGraphInstantiate: " ) + cudaGetErrorString( e ) );
        m->graph_launches[f] = launches;
        m->graph_dotted[f] = d;
        m->graph_cur0[f] = m->lv[0].cur;
    }
    CFB_CUDA( c, cudaGraphLaunch( m->graph[f], c->stream ) );
    m->lv[0].cur = m->graph_cur0[f];
    if ( dotted )
        *dotted = m->graph_dotted[f];
    *n += m->graph_launches[f];
    return CFB_OK;
}

} // namespace

void mg_destroy( cfb_ctx* c ) { mg_free( c ); }

// "mg_graph" / "mg_coarse_kernel" tuning keys
int mg_set_graph( cfb_ctx* c, bool on )
{
    if ( c->mg )
        c->mg->use_graph = on;
    c->mg_graph = on;
    return CFB_OK;
}
int mg_set_coarse_kernel( cfb_ctx* c, bool on )
{
    if ( c->mg )
    {
        c->mg->use_coarse = on;
        for ( cudaGraphExec_t& g : c->mg->graph ) // a captured cycle has the other form baked in
            if ( g )
            {
                cudaGraphExecDestroy( g );
                g = nullptr;
            }
    }
    c->mg_coarse = on;
    return CFB_OK;
}

// Cajita::ReferenceConjugateGradient::solve( b, x ) from x0 = 0 with z = V-cycle( r ).
int mg_pcg_solve( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid )
{
    MgStage* m = c->mg;
    if ( !m )
        return cfb_fail( c, CFB_ERR_INVALID, "multigrid preconditioner not set up" );
    const int fixed = fixed_iters > 0;
    const int max_it = fixed ? fixed_iters : c->cfg.cg_max_iter;
    CgState* S = c->d_state;
    MgLevelHost& F = m->lv[0];
    F.b = c->cg_r;
    const MgLevelDev& L = F.d;
    const int grid = grid_for( c, F.cells );
    const size_t head = offsetof( CgState, hist );
    long long launches = 0;

    int nl = 0; // launches counted by the helpers
    c->sticky_rc = 0;
    mgcg_init_kernel<<<grid, NT, 0, c->stream>>>( L, c->rhs, c->lhs, c->cg_r, S, c->d_partials, fixed );
    MG_TRY( mg_global_sum( c, 0, &nl ) );
    mgcg_check0_kernel<<<1, 1, 0, c->stream>>>( S, c->cfg.cg_tolerance, c->cfg.cg_stop_rule == CFB_STOP_REL );
    launches += 2;

    int enq = 0;
    bool done = false, pending = false;
    while ( enq < max_it && !done )
    {
        // one iteration: z = M^-1 r ; z.r ; p = z + beta p ; q = A p, p.q ; x, r, r.r, stopping test
        bool dotted = false;
        MG_TRY( run_cycle( c, true, &dotted, &nl ) );
        const double* z = F.x[F.cur];
        if ( !dotted ) // no post-smoothing sweep to fuse z.r into
        {
            mgcg_dot_kernel<<<grid, NT, 0, c->stream>>>( L, z, c->cg_r, S, c->d_partials );
            launches += 1;
        }
        MG_TRY( mg_global_sum( c, 1, &nl ) );
        mgcg_pupdate_kernel<<<grid, NT, 0, c->stream>>>( L, z, c->cg_p, S, enq == 0 ? 1 : 0 );
        launches += 1;
        if ( c->cfg.use_nccl )
            MG_TRY( halo_exchange_cells( c, c->cg_p, 1 ) );
        launches += launch_stencil_dot( c );
        if ( c->sticky_rc ) // the stencil launcher refused its tile configuration
        {
            cudaStreamSynchronize( c->stream );
            return take_sticky_rc( c );
        }
        if ( c->cfg.use_nccl )
            MG_TRY( cg_global_sum( c, 0 ) );
        mgcg_axpy_kernel<<<grid, NT, 0, c->stream>>>( L, c->cg_p, c->cg_q, c->lhs, c->cg_r, S, c->d_partials );
        launches += 1;
        MG_TRY( mg_global_sum( c, 2, &nl ) );
        ++enq;
        if ( fixed )
            continue;
        // the state after iteration enq - 1 is looked at while iteration enq is already queued
        if ( pending )
        {
            CFB_CUDA( c, cudaEventSynchronize( m->ev_poll ) );
            done = c->h_state->done != 0;
            pending = false;
        }
        if ( !done )
        {
            CFB_CUDA( c, cudaMemcpyAsync( c->h_state, S, head, cudaMemcpyDeviceToHost, c->stream ) );
            CFB_CUDA( c, cudaEventRecord( m->ev_poll, c->stream ) );
            pending = true;
        }
    }
    if ( m->singular )
    {
        mgcg_nullsum_kernel<<<grid, NT, 0, c->stream>>>( L, c->lhs, S, c->d_partials );
        MG_TRY( mg_global_sum( c, 3, &nl ) );
        mgcg_shift_kernel<<<grid, NT, 0, c->stream>>>( L, c->lhs, S );
        launches += 2;
    }
    CFB_CUDA( c, cudaMemcpyAsync( c->h_state, S, head, cudaMemcpyDeviceToHost, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    cudaError_t e = cudaGetLastError();
    if ( e != cudaSuccess )
        return cfb_fail( c, CFB_ERR_CUDA, std::string( "mg_pcg_solve: " ) + cudaGetErrorString( e ) );
    c->stats.kernel_launches += launches + nl;
    if ( c->h_state->xerror )
        return cfb_fail( c, CFB_ERR_NCCL, "peer-memory exchange timed out: a neighbour never finished its ghost stores" );
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

extern "C" int cfb_set_preconditioner( cfb_ctx* c, int kind, int nu_pre, int nu_post, int nu_coarse, double omega )
{
    if ( kind != CFB_PRECOND_JACOBI && kind != CFB_PRECOND_MG )
        return cfb_fail( c, CFB_ERR_INVALID, "unknown preconditioner" );
    if ( kind == CFB_PRECOND_JACOBI )
    {
        c->precond = CFB_PRECOND_JACOBI;
        return CFB_OK;
    }
    if ( nu_pre < 1 || nu_post < 0 || nu_coarse < 1 || omega >= 2.0 )
        return cfb_fail( c, CFB_ERR_INVALID,
                         "multigrid preconditioner: nu_pre >= 1, nu_post >= 0, nu_coarse >= 1, omega < 2" );
    // several blocks: the levels are coarsened block by block, which needs identical block extents
    for ( int d = 0; d < c->g.D; ++d )
        if ( c->cfg.global_num_cell[d] % c->cfg.ranks_per_dim[d] != 0 )
            return cfb_fail( c, CFB_ERR_INVALID,
                             "multigrid preconditioner: the cells of every dimension must divide evenly among the blocks" );
    int rc = mg_build( c, nu_pre, nu_post, nu_coarse, omega, c->mg_max_levels );
    if ( rc )
        return rc;
    c->precond = CFB_PRECOND_MG;
    return CFB_OK;
}

extern "C" int cfb_set_mg_max_levels( cfb_ctx* c, int max_levels )
{
    c->mg_max_levels = max_levels;
    if ( c->precond == CFB_PRECOND_MG && c->mg )
        return mg_build( c, c->mg->nu1, c->mg->nu2, c->mg->nuc, c->mg->omega, max_levels );
    return CFB_OK;
}

extern "C" int cfb_mg_num_levels( cfb_ctx* c, int* levels )
{
    *levels = c->mg ? (int)c->mg->lv.size() : 0;
    return CFB_OK;
}

// z = M^-1 r for dense owned-cell host arrays: one V-cycle on its own (introspection / tests).
// Uses the CG work vector r as the fine-level right-hand side.
extern "C" int cfb_mg_apply( cfb_ctx* c, const double* r_host, double* z_host )
{
    MgStage* m = c->mg;
    if ( c->precond != CFB_PRECOND_MG || !m )
        return cfb_fail( c, CFB_ERR_INVALID, "multigrid preconditioner not set" );
    int rc = cfb_upload( c, CFB_CG_R, CFB_CURRENT, CFB_OWNED, r_host );
    if ( rc )
        return rc;
    MgLevelHost& F = m->lv[0];
    F.b = c->cg_r;
    int nl = 0;
    rc = run_cycle( c, false, nullptr, &nl );
    if ( rc )
        return rc;
    c->stats.kernel_launches += nl;
    const Geo& g = c->g;
    const double* z = F.x[F.cur] + g.origin;
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr( const_cast<double*>( z ), (size_t)g.sy * 8, (size_t)g.sy, (size_t)g.ay );
    p.dstPtr = make_cudaPitchedPtr( z_host, (size_t)g.n[0] * 8, (size_t)g.n[0], (size_t)g.n[1] );
    p.extent = make_cudaExtent( (size_t)g.n[0] * 8, (size_t)g.n[1], (size_t)g.n[2] );
    p.kind = cudaMemcpyDeviceToHost;
    CFB_CUDA( c, cudaMemcpy3DAsync( &p, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    return CFB_OK;
}
