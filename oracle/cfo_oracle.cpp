// cfo_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A dependency-free C++17 (+OpenMP) restatement of the cajitafluids hot path
// (divergence -> Jacobi-PCG -> gradient subtraction, RK3 semi-Lagrangian advection,
// inflow/body-force inputs) used ONLY as the checker in tests/, in
// __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference leg.
// Nothing in cajitafluids_b200/ (the product) links, loads or calls this file.
//
// PARITY STATUS.  The reference cannot be built as shipped (Kokkos, Cabana/Cajita, MPI, Silo absent;
// SURVEY.md F3) and none of its tests touches this path.  It IS run here, though:
// `make -C oracle ref` compiles the unmodified reference sources from /root/reference against
// single-rank stand-ins for those dependencies (oracle/refshim/, see its README) into
// oracle/_ref/libcfref.so, and tests/test_reference_shim.py holds this file against it BIT FOR BIT
// (whole 2-D runs, every stage alone on seeded fields, the stored matrix, ghosts, the CG residual
// history, the solve loop, the error paths); tests/golden/refrun_*.npz are that library's outputs.
// So everything that lives in the reference tree is PINNED against the reference's own statements.
// PARITY UNPINNED for the rest: the third-party arithmetic that is NOT in the reference tree —
// Cabana (ECP-copa/Cabana, Cajita sub-library, unpinned: `cabana@master` in
// configs/llnl-lassen/spack.yaml:12, API level ~0.5): CG loop, B-splines, G2P, LocalMesh
// coordinates, index spaces — is restated from its published algorithm, marked [Cajita-mem] here
// and in the stand-in, and anchored on (a) the reference's geometry assertions (tests/tstMesh.cpp,
// tests/tstProblemManager.cpp: run unmodified on the stand-in, and re-expressed in
// tests/test_oracle_geometry.py), (b) the analytic known-answer tests of SURVEY.md §8c and (c) an
// independent numpy/scipy model (tests/ref2d_numpy.py).  The 3-D extension has no reference at all.
//
// The reference is 2-D only.  dim == 2 follows it statement by statement;
// dim == 3 is the obvious extension, every 3-D choice is marked [3D-ext].
//
// Data structures deliberately mirror the reference so this doubles as the CPU baseline:
// ghosted arrays (halo 3 on every side, also on physical walls), a stored 2*D+1 coefficient
// matrix + stored inverse diagonal, the 4-kernel / 3-reduction Cajita CG, one loop nest per
// advected field, a full field gather before the divergence.
//
// All file:line citations are relative to the cajitafluids tree.

#include "../include/cfb.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{

using clk = std::chrono::steady_clock;

// Accumulator of the global sums (r.r, z.r, p.Ap).  The reference leaves the summation order to
// Kokkos::parallel_reduce (it changes with backend and thread count, and with it the last bits of
// alpha/beta, which CG then amplifies by ~1e4-1e5 over a few hundred iterations: measured 1e-11
// relative in the pressure after 20 steps at 64^3 between 1 and 8 OpenMP threads).  In its default
// mode the oracle therefore accumulates the (double-rounded) products in double-double (TwoSum),
// i.e. it returns the correctly rounded EXACT sum, the centre of all admissible orders; the CUDA
// path does the same, which makes whole runs bit-comparable.  Mode 0 (plain double, the reference's
// arithmetic) is kept for the CPU-baseline timing.
struct acc_t
{
    double hi = 0.0, lo = 0.0;
    bool exact = true;
    acc_t() {}
    explicit acc_t( bool e ) : exact( e ) {}
    inline void add( double x )
    {
        if ( exact )
        {
            const double s = hi + x;
            const double bb = s - hi;
            const double e = ( hi - ( s - bb ) ) + ( x - bb );
            hi = s;
            lo += e;
        }
        else
            hi += x;
    }
    inline void merge( const acc_t& b )
    {
        if ( exact )
        {
            const double s = hi + b.hi;
            const double bb = s - hi;
            double e = ( hi - ( s - bb ) ) + ( b.hi - bb );
            e += lo + b.lo;
            hi = s + e;
            lo = e - ( hi - s );
        }
        else
            hi += b.hi;
    }
    inline double value() const { return hi + lo; }
};
#pragma omp declare reduction( accsum : acc_t : omp_out.merge( omp_in ) ) initializer( omp_priv = acc_t( omp_orig.exact ) )

// ---------------------------------------------------------------------------------------
// A ghosted array of one entity type (Cajita::Array with one dof), x fastest.
// Entity: 0 = Cell, 1 = Face<I>, 2 = Face<J>, 3 = Face<K>.
struct Arr
{
    int e[3] = { 1, 1, 1 }; // ghosted extents
    std::vector<double> a;
    void alloc( const int ext[3] )
    {
        for ( int d = 0; d < 3; ++d )
            e[d] = ext[d];
        // Cajita::ArrayOp::assign( ..., 0.0, Ghost() )  src/ProblemManager.hpp:149-165
        a.assign( (size_t)e[0] * e[1] * e[2], 0.0 );
    }
    inline size_t idx( int i, int j, int k ) const
    {
        return ( (size_t)k * e[1] + j ) * e[0] + i;
    }
    inline double& operator()( int i, int j, int k ) { return a[idx( i, j, k )]; }
    inline double operator()( int i, int j, int k ) const { return a[idx( i, j, k )]; }
};

struct Space
{
    int lo[3], hi[3];
};

typedef void ( *gather_cb )( void* user, int what );
typedef void ( *allreduce_cb )( void* user, double* vals, int n );

} // namespace

struct cfo_ctx
{
    cfb_config cfg;
    int D;
    int h;         // halo cell width
    int n[3];      // owned cells of this block
    int off[3];    // global cell offset of this block
    bool hi_bd[3]; // block touches the high physical boundary in dim d
    bool lo_bd[3];
    double cell;   // Mesh::cellSize == globalMesh().cellSize( 0 )  (src/Mesh.hpp:119-122)
    // Cajita's UniformGlobalMesh keeps one cell size PER DIMENSION, (hi_d - lo_d) / n_d [Cajita-mem]
    // (createUniformGlobalMesh( low, high, num_cell ), src/Mesh.hpp:79-80); the Mesh ctor only checks
    // that they agree with cellSize( 0 ) to 10 eps (:56-64).  LocalMesh::coordinates and the spline
    // logical coordinates use these; the operator scales, the dt clamp and the divergence use `cell`.
    double celld[3];
    double dt;     // clamped
    double time;
    double ghost_low[3]; // LocalMesh ghosted low corner [Cajita-mem]
    int bc_min[3], bc_max[3];

    // ProblemManager state: [entity][version]
    Arr fld[4][2];
    int cur[4];

    // VelocityCorrector
    Arr lhs, rhs;
    int nst;                // 2*D+1
    std::vector<double> A;  // (cell, c) c fastest  -- getMatrixValues()
    std::vector<double> Mi; // getPreconditionerValues()
    Arr cg_p, cg_z, cg_r, cg_q;
    Arr cg_s;          // single-reduction CG only (cg_single_reduction): s = A p by recurrence; allocated on first use
    int cg_algorithm = 0; // 0 = Cajita's ReferenceConjugateGradient loop, 1 = single-reduction (Chronopoulos-Gear) form
    int num_iter;
    double resid; // sqrt(sum r^2)
    std::vector<double> hist;

    // opt-in multigrid preconditioner (cfo_set_preconditioner; NOT the reference's: see mg_* below)
    int precond = 0; // 0 = the reference's diagonal (Jacobi) preconditioner, 1 = geometric multigrid V-cycle
    int mg_nu1 = 2, mg_nu2 = 2, mg_nuc = 8;
    int mg_max_levels = 0; // 0 = as many as the grid allows
    double mg_omega = 0.0;
    std::vector<double> mg_wpre, mg_wpost; // damping of every pre- / post-smoothing sweep
    double mg_wc = 0.0;                    // damping on the coarsest level
    struct Mg* mg = nullptr;

    // distributed hooks (tests drive halo exchange / allreduce over gloo)
    gather_cb gcb = nullptr;
    allreduce_cb acb = nullptr;
    void* cb_user = nullptr;

    double t_phase[8] = { 0 };
    long long cg_total = 0;
    long long steps = 0;
    bool accum_exact = true; // see acc_t
    std::string err;
};

namespace
{

std::string g_err;

// ---------------------------------------------------------------------------------------
// Cajita GlobalGrid block partition [Cajita-mem]: n/nb cells per block, the first n%nb
// blocks get one more.
void partition( int n, int nb, int b, int& owned, int& offset )
{
    int base = n / nb, rem = n % nb;
    owned = base + ( b < rem ? 1 : 0 );
    offset = b * base + std::min( b, rem );
}

// Ghosted extents of an entity: owned cells + 2*halo (+1 along the face normal).
// tests/tstMesh.cpp:61-68 pins n + 2*halo + 1 by n + 2*halo for Face<I> on one rank.
void ghost_ext( const cfo_ctx& c, int ent, int ext[3] )
{
    for ( int d = 0; d < 3; ++d )
    {
        if ( d < c.D )
            ext[d] = c.n[d] + 2 * c.h + ( ent - 1 == d ? 1 : 0 );
        else
            ext[d] = 1;
    }
}

// LocalGrid::indexSpace( Own(), entity, Local() ) [Cajita-mem]: cells [h, h+n); faces get one
// more along their normal only on the block touching the high (non-periodic) wall.
Space own_space( const cfo_ctx& c, int ent )
{
    Space s;
    for ( int d = 0; d < 3; ++d )
    {
        if ( d < c.D )
        {
            s.lo[d] = c.h;
            s.hi[d] = c.h + c.n[d] + ( ( ent - 1 == d && c.hi_bd[d] ) ? 1 : 0 );
        }
        else
        {
            s.lo[d] = 0;
            s.hi[d] = 1;
        }
    }
    return s;
}

// LocalMesh::coordinates( entity, idx, x ) [Cajita-mem]:
//   x[d] = ghost_low[d] + (idx[d] + 0.5) * cell   for cells and tangential face dirs
//   x[d] = ghost_low[d] +  idx[d]        * cell   along the face normal
inline void coordinates( const cfo_ctx& c, int ent, const int idx[3], double x[3] )
{
    for ( int d = 0; d < c.D; ++d )
    {
        if ( ent - 1 == d )
            x[d] = c.ghost_low[d] + double( idx[d] ) * c.celld[d];
        else
            x[d] = c.ghost_low[d] + ( double( idx[d] ) + 0.5 ) * c.celld[d];
    }
}

// IndexConversion::createL2G [Cajita-mem]
inline int l2g( const cfo_ctx& c, int d, int i ) { return i - c.h + c.off[d]; }

// ---------------------------------------------------------------------------------------
// Cajita::Spline<1> / Spline<3> + evaluateSpline + G2P::value [Cajita-mem]
//   logical coordinate  xl = (x - x_of_entity_0) * (1/cell)
//   order 1: s = int(xl), int(xl)+1 ; w = 1-f, f            with f = xl - int(xl)
//   order 3: s = int(xl)-1 .. int(xl)+2 ; cubic B-spline weights evaluated from the distance
//            to the first knot, xn = f + 1, stepping xn -= 1 per knot.
inline void spline1( double xl, int s[2], double w[2] )
{
    int i0 = static_cast<int>( xl );
    s[0] = i0;
    s[1] = i0 + 1;
    double xn = xl - double( i0 );
    w[0] = 1.0 - xn;
    w[1] = xn;
}
inline void spline3( double xl, int s[4], double w[4] )
{
    int i0 = static_cast<int>( xl );
    s[0] = i0 - 1;
    s[1] = i0;
    s[2] = i0 + 1;
    s[3] = i0 + 2;
    const double one_sixth = 1.0 / 6.0;
    const double two_thirds = one_sixth * 4.0;
    const double four_thirds = 2.0 * two_thirds;
    double xn = xl - double( i0 ) + 1.0;
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

// Interpolation::interpolateField<D, order, Entity>  src/Interpolation.hpp:30-41
// The reference reads whatever index the spline produces (no clamping; leaving the halo is
// UB there).  We clamp the index into the allocation so a bad CFL cannot segfault the
// checker; parity is only defined for foot points inside the halo.
template <int ORDER>
inline double interpolate_field( const cfo_ctx& c, int ent, const double loc[3],
                                 const Arr& f )
{
    constexpr int NK = ORDER + 1;
    int s[3][NK];
    double w[3][NK];
    const int zero[3] = { 0, 0, 0 };
    double low[3];
    coordinates( c, ent, zero, low );
    for ( int d = 0; d < c.D; ++d )
    {
        const double rdx = 1.0 / c.celld[d]; // evaluateSpline: 1 / local_mesh.cellSize( d ) [Cajita-mem]
        double xl = ( loc[d] - low[d] ) * rdx;
        if ( ORDER == 1 )
            spline1( xl, s[d], w[d] );
        else
            spline3( xl, s[d], w[d] );
        for ( int a = 0; a < NK; ++a )
            s[d][a] = std::min( std::max( s[d][a], 0 ), f.e[d] - 1 );
    }
    double value = 0.0;
    if ( c.D == 2 )
    {
        for ( int a = 0; a < NK; ++a )
            for ( int b = 0; b < NK; ++b )
                value += f( s[0][a], s[1][b], 0 ) * w[0][a] * w[1][b];
    }
    else
    {
        for ( int a = 0; a < NK; ++a )
            for ( int b = 0; b < NK; ++b )
                for ( int g = 0; g < NK; ++g )
                    value += f( s[0][a], s[1][b], s[2][g] ) * w[0][a] * w[1][b] * w[2][g];
    }
    return value;
}

// Interpolation::interpolateVelocity<D,1>  src/Interpolation.hpp:43-54  (+ w: [3D-ext])
inline void interpolate_velocity( const cfo_ctx& c, const double loc[3], double vel[3] )
{
    for ( int d = 0; d < c.D; ++d )
        vel[d] = interpolate_field<1>( c, 1 + d, loc, c.fld[1 + d][c.cur[1 + d]] );
}

// TimeIntegrator::rk3  src/TimeIntegrator.hpp:36-77
// Q2: the third stage uses v0 (":57-58"); the 3-D stub at :59-60 writes x1[2] instead of
// x2[2] — [3D-ext] fills x2[2] properly but keeps the Q2 choice of v0.
inline void rk3( const cfo_ctx& c, const double x0[3], double dt, double trace[3] )
{
    double v0[3], x1[3], v1[3], x2[3], v2[3];
    interpolate_velocity( c, x0, v0 );
    for ( int d = 0; d < c.D; ++d )
        x1[d] = x0[d] - 0.5 * dt * v0[d];
    interpolate_velocity( c, x1, v1 );
    const double* vs = c.cfg.quirk_rk3_stage3_v0 ? v0 : v1;
    for ( int d = 0; d < c.D; ++d )
        x2[d] = x0[d] - 0.75 * dt * vs[d];
    interpolate_velocity( c, x2, v2 );
    for ( int d = 0; d < c.D; ++d )
        trace[d] = x0[d] - dt * ( ( 2.0 / 9.0 ) * v0[d] + ( 3.0 / 9.0 ) * v1[d] +
                                  ( 4.0 / 9.0 ) * v2[d] );
}

// TimeIntegrator::advect  src/TimeIntegrator.hpp:81-116
void advect( cfo_ctx& c, int ent )
{
    const Arr& fc = c.fld[ent][c.cur[ent]];
    Arr& fn = c.fld[ent][1 - c.cur[ent]];
    Space s = own_space( c, ent );
    const int order = c.cfg.field_interp_order;
    const double dt = c.dt;
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                int idx[3] = { i, j, k };
                double start[3], trace[3];
                coordinates( c, ent, idx, start );
                rk3( c, start, dt, trace );
                fn( i, j, k ) = ( order == 1 ) ? interpolate_field<1>( c, ent, trace, fc )
                                               : interpolate_field<3>( c, ent, trace, fc );
            }
}

inline void do_gather( cfo_ctx& c, int what )
{
    if ( c.gcb )
        c.gcb( c.cb_user, what );
}
inline void do_allreduce( cfo_ctx& c, double* v, int n )
{
    if ( c.acb )
        c.acb( c.cb_user, v, n );
}

// TimeIntegrator::step  src/TimeIntegrator.hpp:120-177
void time_integrator_step( cfo_ctx& c )
{
    auto t0 = clk::now();
    do_gather( c, 0 ); // pm.gather( Version::Current() )   :131
    for ( int ent = 0; ent <= c.D; ++ent )
        advect( c, ent ); // :137-160
    for ( int ent = 0; ent <= c.D; ++ent )
        c.cur[ent] = 1 - c.cur[ent]; // pm.advance   :167-172
    c.t_phase[0] += std::chrono::duration<double>( clk::now() - t0 ).count();
}

// BoundaryCondition::operator()( Face<d>, ... )  src/BoundaryConditions.hpp:102-129
inline void bc_face( const cfo_ctx& c, int d, Arr& f, const int g[3], int i, int j, int k )
{
    if ( g[d] <= c.bc_min[d] && c.cfg.boundary_type[d] == CFB_SOLID )
        f( i, j, k ) = 0;
    if ( g[d] > c.bc_max[d] && c.cfg.boundary_type[c.D + d] == CFB_SOLID )
        f( i, j, k ) = 0;
}

// InflowSource box test  src/InflowSource.hpp:40-41 (+ z: [3D-ext])
inline bool in_inflow( const cfo_ctx& c, const double x[3] )
{
    for ( int d = 0; d < c.D; ++d )
    {
        double lo = c.cfg.inflow_location[d];
        double hi = c.cfg.inflow_location[d] + c.cfg.inflow_size[d];
        if ( !( x[d] >= lo && x[d] < hi ) )
            return false;
    }
    return true;
}

// Solver::_addInputs  src/Solver.hpp:181-263
void add_inputs( cfo_ctx& c )
{
    auto t0 = clk::now();
    {
        Arr& q = c.fld[0][c.cur[0]];
        Space s = own_space( c, 0 );
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    int idx[3] = { i, j, k };
                    double x[3];
                    coordinates( c, 0, idx, x );
                    // InflowSource( Cell )  src/InflowSource.hpp:33-50
                    if ( in_inflow( c, x ) && q( i, j, k ) < c.cfg.inflow_quantity )
                        q( i, j, k ) = c.cfg.inflow_quantity;
                    // BodyForce( Cell ) is empty  src/BodyForce.hpp:33-41
                }
    }
    for ( int d = 0; d < c.D; ++d )
    {
        Arr& u = c.fld[1 + d][c.cur[1 + d]];
        Space s = own_space( c, 1 + d );
        const double V = c.cfg.inflow_velocity[d];
        const double fdt = c.cfg.body_force[d] * c.dt;
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    int idx[3] = { i, j, k };
                    double x[3];
                    coordinates( c, 1 + d, idx, x );
                    int g[3] = { l2g( c, 0, i ), l2g( c, 1, j ), c.D == 3 ? l2g( c, 2, k ) : 0 };
                    // InflowSource( Face )  src/InflowSource.hpp:52-78
                    if ( in_inflow( c, x ) && std::fabs( u( i, j, k ) ) < std::fabs( V ) )
                        u( i, j, k ) = V;
                    // BodyForce( Face )  src/BodyForce.hpp:43-60
                    u( i, j, k ) += fdt;
                    bc_face( c, d, u, g, i, j, k );
                }
    }
    c.t_phase[1] += std::chrono::duration<double>( clk::now() - t0 ).count();
}

// VelocityCorrector::initializeMatrixValues + BoundaryCondition::build_matrix +
// fillMatrixValues( reference )   src/VelocityCorrector.hpp:116-144,158-180
// src/BoundaryConditions.hpp:56-97.   Stencil order {0},{-x},{+x},{-y},{+y}(,{-z},{+z}).
void fill_matrix( cfo_ctx& c )
{
    Space s = own_space( c, 0 );
    const double scale = c.dt / ( c.cfg.density * c.cell * c.cell ); // :128
    const int nst = c.nst;
    const Arr& L = c.lhs;
    c.A.assign( L.a.size() * nst, 0.0 );
    c.Mi.assign( L.a.size(), 0.0 );
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                double* m = &c.A[L.idx( i, j, k ) * nst];
                int g[3] = { l2g( c, 0, i ), l2g( c, 1, j ), c.D == 3 ? l2g( c, 2, k ) : 0 };
                m[0] = ( 2.0 * c.D ) * scale; // 4.0 * scale  :137  (6.0 in 3-D [3D-ext])
                for ( int e = 1; e < nst; ++e )
                    m[e] = -1.0 * scale;
                for ( int d = 0; d < c.D; ++d )
                {
                    if ( g[d] <= c.bc_min[d] )
                    { // low wall of dim d
                        m[1 + 2 * d] = 0;
                        if ( c.cfg.boundary_type[d] == CFB_SOLID )
                            m[0] -= scale;
                    }
                    if ( g[d] > c.bc_max[d] - 1 )
                    { // high wall of dim d
                        m[2 + 2 * d] = 0;
                        if ( c.cfg.boundary_type[c.D + d] == CFB_SOLID )
                            m[0] -= scale;
                    }
                }
                c.Mi[L.idx( i, j, k )] = 1.0 / m[0]; // :178
            }
}

// VelocityCorrector::_buildRHS  src/VelocityCorrector.hpp:182-212
void build_rhs( cfo_ctx& c )
{
    auto t0 = clk::now();
    do_gather( c, 0 ); // _pm->gather( Version::Current() )  :190  (Q7)
    Space s = own_space( c, 0 );
    const double scale = 1.0 / c.cell;
    const Arr& u = c.fld[1][c.cur[1]];
    const Arr& v = c.fld[2][c.cur[2]];
    const Arr* w = c.D == 3 ? &c.fld[3][c.cur[3]] : nullptr;
    Arr& rhs = c.rhs;
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                double div = u( i + 1, j, k ) - u( i, j, k ) + v( i, j + 1, k ) - v( i, j, k );
                if ( w )
                    div = div + ( *w )( i, j, k + 1 ) - ( *w )( i, j, k ); // [3D-ext]
                rhs( i, j, k ) = -scale * div;
            }
    // Cajita::ArrayOp::assign( *_lhs, 0.0, Own() )   :272
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                c.lhs( i, j, k ) = 0.0;
    c.t_phase[2] += std::chrono::duration<double>( clk::now() - t0 ).count();
}

// sum_c A(cell,c) * x(cell + off_c), in stencil order, each term accumulated with one fused
// multiply-add (what `Ax += A * x` compiles to on an FMA machine).
inline double apply_A( const cfo_ctx& c, const Arr& x, int i, int j, int k )
{
    const double* m = &c.A[x.idx( i, j, k ) * c.nst];
    double Ax = 0.0;
    Ax = std::fma( m[0], x( i, j, k ), Ax );
    Ax = std::fma( m[1], x( i - 1, j, k ), Ax );
    Ax = std::fma( m[2], x( i + 1, j, k ), Ax );
    Ax = std::fma( m[3], x( i, j - 1, k ), Ax );
    Ax = std::fma( m[4], x( i, j + 1, k ), Ax );
    if ( c.D == 3 )
    {
        Ax = std::fma( m[5], x( i, j, k - 1 ), Ax );
        Ax = std::fma( m[6], x( i, j, k + 1 ), Ax );
    }
    return Ax;
}


// ---------------------------------------------------------------------------------------
// Opt-in geometric multigrid preconditioner for the CG (the role HYPRE's PFMG
// plays in the reference's default path, examples/advection.cpp:186-189, src/VelocityCorrector.hpp:
// 324-338).  This is NOT a restatement of anything in the reference tree or of HYPRE: it is the
// CPU statement of the product's own algorithm (cajitafluids_b200/csrc/mg.cu), operation for
// operation, so that the two can be compared bit for bit.  Parity with the reference is defined
// only through the solution of the same linear system (same matrix, same stopping test).
//
//   z = M^-1 r  :=  one V(nu1, nu2) cycle on A z = r from a zero initial guess
//   levels      : cell-centred 2:1 coarsening in every dimension while all extents stay even and
//                 >= 2 after halving; level l uses the SAME 2*D+1-point operator (same boundary
//                 logic) with scale_l = scale_0 / 4^l (re-discretisation, not Galerkin)
//   smoother    : damped Jacobi, x += omega D^-1 (b - A x); the first pre-smoothing sweep starts
//                 from x = 0 and is x = (omega D^-1) b.  Default damping: one omega per sweep, the reciprocals
//                 of the Chebyshev nodes of [0.4, 2] (the upper part of the spectrum of D^-1 A) in 3-D with
//                 2..4 sweeps — 0.566, 1.577 for V(2,2): 8 instead of 12 iterations at 64^3, 9 instead of 14
//                 at 128^3, for nothing — post-smoothing in reverse order; 6/7 (3-D, one sweep) and 0.8 (2-D,
//                 where the schedule does not help) otherwise; a positive omega argument fixes it for all sweeps
//   restriction : mean of the 2^D children of the residual; prolongation: piecewise constant
//   coarsest    : nuc Jacobi sweeps
// With nu1 == nu2 the cycle is a symmetric operator (R is a multiple of P^T, the smoother is
// symmetric), which CG needs.
struct MgLevel
{
    int n[3];
    int cz;                  // coarsening factor to the next level along z (1 in 2-D)
    bool slo[3], shi[3];     // the low / high end of dim d is a SOLID physical wall
    double scale, ns;
    double diag[8], minv[8]; // by number of SOLID walls touched: diagonal, 1 / diagonal
    Arr b, x[2];             // ghosted by one layer (ghosts stay zero)
    int cur = 0;             // x[cur] holds the level's result
    inline int walls( int i, int j, int k ) const
    {
        return ( i == 0 && slo[0] ) + ( i == n[0] - 1 && shi[0] ) + ( j == 0 && slo[1] ) +
               ( j == n[1] - 1 && shi[1] ) + ( k == 0 && slo[2] ) + ( k == n[2] - 1 && shi[2] );
    }
    // the row of apply_A, matrix-free: diag * x, then one fused multiply-add per neighbour in stencil
    // order (off-domain neighbours are ghost zeros); (i, j, k) are owned indices
    inline double Ax( const Arr& v, int i, int j, int k ) const
    {
        const int I = i + 1, J = j + 1, K = k + 1;
        double a = diag[walls( i, j, k )] * v( I, J, K );
        a = std::fma( ns, v( I - 1, J, K ), a );
        a = std::fma( ns, v( I + 1, J, K ), a );
        a = std::fma( ns, v( I, J - 1, K ), a );
        a = std::fma( ns, v( I, J + 1, K ), a );
        a = std::fma( ns, v( I, J, K - 1 ), a );
        a = std::fma( ns, v( I, J, K + 1 ), a );
        return a;
    }
};

} // namespace

struct Mg
{
    std::vector<MgLevel> lv;
};

namespace
{

// damping per sweep (see the header comment of this section); literals, so that every implementation uses the
// same bits: 1 / ( 1.2 + 0.8 cos( (2k - 1) pi / (2 nu) ) ), k = 1..nu
void mg_schedule( cfo_ctx& c )
{
    static const double cheb[5][4] = { { 0, 0, 0, 0 },
                                       { 0, 0, 0, 0 },
                                       { 0.5663522991524661, 1.576504843704677, 0, 0 },
                                       { 0.5283121635129678, 0.8333333333333334, 1.9716878364870327, 0 },
                                       { 0.515702196925985, 0.6639459287266942, 1.118751870515935, 2.169685110214365 } };
    const double fixed = c.mg_omega > 0.0 ? c.mg_omega : ( c.D == 3 ? 6.0 / 7.0 : 0.8 );
    c.mg_wc = fixed;
    auto fill = [&]( std::vector<double>& w, int nu, bool reverse ) {
        w.assign( nu > 0 ? nu : 0, fixed );
        // only for the symmetric cycle (nu1 == nu2): the big per-sweep factors are harmless as a product,
        // not one by one, and CG needs M symmetric positive definite
        if ( c.mg_omega <= 0.0 && c.D == 3 && nu >= 2 && nu <= 4 && c.mg_nu1 == c.mg_nu2 )
            for ( int s = 0; s < nu; ++s )
                w[s] = cheb[nu][reverse ? nu - 1 - s : s];
    };
    fill( c.mg_wpre, c.mg_nu1, false );
    fill( c.mg_wpost, c.mg_nu2, true );
}

void mg_build( cfo_ctx& c )
{
    delete c.mg;
    c.mg = new Mg();
    const int D = c.D;
    mg_schedule( c );
    int n[3] = { c.n[0], c.n[1], D == 3 ? c.n[2] : 1 };
    double scale = c.dt / ( c.cfg.density * c.cell * c.cell ); // src/VelocityCorrector.hpp:128
    for ( int l = 0; l < 16; ++l )
    {
        c.mg->lv.emplace_back();
        MgLevel& L = c.mg->lv.back();
        for ( int d = 0; d < 3; ++d )
        {
            L.n[d] = n[d];
            // 2-D: one plane between two SOLID z walls, the same trick as the CUDA kernels (6 - 2 = 4)
            L.slo[d] = d < D ? ( c.lo_bd[d] && c.cfg.boundary_type[d] == CFB_SOLID ) : true;
            L.shi[d] = d < D ? ( c.hi_bd[d] && c.cfg.boundary_type[D + d] == CFB_SOLID ) : true;
        }
        L.cz = D == 3 ? 2 : 1;
        L.scale = scale;
        L.ns = -1.0 * scale;
        for ( int cnt = 0; cnt < 8; ++cnt )
        {
            double dgl = 6.0 * scale; // 2-D: 4 * scale == 6 * scale - scale - scale only up to rounding,
            if ( D == 2 )             // so follow the reference's own sequence (:137, BoundaryConditions.hpp:56-97)
            {
                dgl = 4.0 * scale;
                for ( int i = 0; i < cnt - 2; ++i )
                    dgl -= scale;
            }
            else
                for ( int i = 0; i < cnt; ++i )
                    dgl -= scale;
            L.diag[cnt] = dgl;
            L.minv[cnt] = 1.0 / dgl;
        }
        const int ext[3] = { n[0] + 2, n[1] + 2, n[2] + 2 };
        L.b.alloc( ext );
        L.x[0].alloc( ext );
        L.x[1].alloc( ext );
        bool can = c.mg_max_levels <= 0 || l + 1 < c.mg_max_levels;
        for ( int d = 0; d < D; ++d )
            can = can && n[d] % 2 == 0 && n[d] / 2 >= 2;
        if ( !can )
            break;
        for ( int d = 0; d < D; ++d )
            n[d] /= 2;
        scale = scale * 0.25;
    }
}

inline void mg_smooth0( MgLevel& L, double omega )
{
    Arr& x = L.x[0];
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = 0; k < L.n[2]; ++k )
        for ( int j = 0; j < L.n[1]; ++j )
            for ( int i = 0; i < L.n[0]; ++i )
                x( i + 1, j + 1, k + 1 ) = ( omega * L.minv[L.walls( i, j, k )] ) * L.b( i + 1, j + 1, k + 1 );
    L.cur = 0;
}

inline void mg_smooth( MgLevel& L, double omega )
{
    const Arr& xi = L.x[L.cur];
    Arr& xo = L.x[1 - L.cur];
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = 0; k < L.n[2]; ++k )
        for ( int j = 0; j < L.n[1]; ++j )
            for ( int i = 0; i < L.n[0]; ++i )
            {
                const double res = L.b( i + 1, j + 1, k + 1 ) - L.Ax( xi, i, j, k );
                xo( i + 1, j + 1, k + 1 ) = std::fma( omega * L.minv[L.walls( i, j, k )], res, xi( i + 1, j + 1, k + 1 ) );
            }
    L.cur = 1 - L.cur;
}

// coarse b = mean of the children's residuals, summed pairwise: x pairs, then y, then z
inline void mg_restrict( const MgLevel& F, MgLevel& C )
{
    const Arr& x = F.x[F.cur];
    auto res = [&]( int i, int j, int k ) { return F.b( i + 1, j + 1, k + 1 ) - F.Ax( x, i, j, k ); };
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int K = 0; K < C.n[2]; ++K )
        for ( int J = 0; J < C.n[1]; ++J )
            for ( int I = 0; I < C.n[0]; ++I )
            {
                const int i = 2 * I, j = 2 * J, k = F.cz * K;
                double s = ( res( i, j, k ) + res( i + 1, j, k ) ) + ( res( i, j + 1, k ) + res( i + 1, j + 1, k ) );
                if ( F.cz == 2 )
                {
                    const double t = ( res( i, j, k + 1 ) + res( i + 1, j, k + 1 ) ) +
                                     ( res( i, j + 1, k + 1 ) + res( i + 1, j + 1, k + 1 ) );
                    s = ( s + t ) * 0.125;
                }
                else
                    s = s * 0.25;
                C.b( I + 1, J + 1, K + 1 ) = s;
            }
}

inline void mg_prolong( MgLevel& F, const MgLevel& C )
{
    Arr& x = F.x[F.cur];
    const Arr& e = C.x[C.cur];
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = 0; k < F.n[2]; ++k )
        for ( int j = 0; j < F.n[1]; ++j )
            for ( int i = 0; i < F.n[0]; ++i )
                x( i + 1, j + 1, k + 1 ) = x( i + 1, j + 1, k + 1 ) + e( i / 2 + 1, j / 2 + 1, k / F.cz + 1 );
}

void mg_vcycle( cfo_ctx& c, int l )
{
    Mg& m = *c.mg;
    MgLevel& L = m.lv[l];
    const bool last = l + 1 == (int)m.lv.size();
    // the coarsest level: nuc sweeps with the fixed damping; elsewhere the pre-smoothing schedule
    mg_smooth0( L, last ? c.mg_wc : c.mg_wpre[0] );
    for ( int s = 1; s < ( last ? c.mg_nuc : c.mg_nu1 ); ++s )
        mg_smooth( L, last ? c.mg_wc : c.mg_wpre[s] );
    if ( last )
        return;
    mg_restrict( L, m.lv[l + 1] );
    mg_vcycle( c, l + 1 );
    mg_prolong( L, m.lv[l + 1] );
    for ( int s = 0; s < c.mg_nu2; ++s )
        mg_smooth( L, c.mg_wpost[s] );
}

// z = M^-1 r on the owned cells of the ghosted CG arrays
void mg_apply( cfo_ctx& c, const Arr& r, Arr& z )
{
    if ( !c.mg )
        mg_build( c );
    MgLevel& L = c.mg->lv[0];
    const Space s = own_space( c, 0 );
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = 0; k < L.n[2]; ++k )
        for ( int j = 0; j < L.n[1]; ++j )
            for ( int i = 0; i < L.n[0]; ++i )
                L.b( i + 1, j + 1, k + 1 ) = r( s.lo[0] + i, s.lo[1] + j, s.lo[2] + k );
    mg_vcycle( c, 0 );
    const Arr& x = L.x[L.cur];
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = 0; k < L.n[2]; ++k )
        for ( int j = 0; j < L.n[1]; ++j )
            for ( int i = 0; i < L.n[0]; ++i )
                z( s.lo[0] + i, s.lo[1] + j, s.lo[2] + k ) = x( i + 1, j + 1, k + 1 );
}

// Cajita::ReferenceConjugateGradient::solve( b, x ) [Cajita-mem]  
// driven from src/VelocityCorrector.hpp:276 with tol 1e-6 / max_iter 2000 (:103-104),
// diagonal preconditioner (:166-179).  Absolute 2-norm stopping test.
// gather ids: 1 = x halo, 2 = r halo (width 0 for a diagonal M: no-op), 3 = p halo.
int cg_solve_single_reduction( cfo_ctx& c );
int cg_solve( cfo_ctx& c )
{
    if ( c.cg_algorithm == 1 && c.precond == 0 )
        return cg_solve_single_reduction( c );
    auto t0 = clk::now();
    Space s = own_space( c, 0 );
    const Arr& b = c.rhs;
    Arr& x = c.lhs;
    Arr &p = c.cg_p, &z = c.cg_z, &r = c.cg_r, &q = c.cg_q;
    const double tol = c.cfg.cg_tolerance;
    const int max_iter = c.cfg.cg_fixed_iters > 0 ? c.cfg.cg_fixed_iters : c.cfg.cg_max_iter;
    const bool fixed = c.cfg.cg_fixed_iters > 0;
    c.num_iter = 0;
    c.hist.clear();
    double thresh = tol;

    // r0 = b - A x0 ; rr
    do_gather( c, 1 );
    acc_t rr_acc( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : rr_acc )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                double r_new = b( i, j, k ) - apply_A( c, x, i, j, k );
                r( i, j, k ) = r_new;
                rr_acc.add( r_new * r_new );
            }
    double rr = rr_acc.value();
    do_allreduce( c, &rr, 1 );
    c.resid = std::sqrt( rr );
    if ( c.cfg.cg_stop_rule == CFB_STOP_REL )
        thresh = tol * c.resid; // x0 = 0 => r0 = b
    if ( !fixed && c.resid <= thresh )
    {
        c.t_phase[3] += std::chrono::duration<double>( clk::now() - t0 ).count();
        return CFB_OK;
    }

    // z0 = M r0 ; p0 = z0 ; zTr
    do_gather( c, 2 );
    const bool mgp = c.precond == 1; // opt-in multigrid V-cycle instead of the reference's diagonal M
    if ( mgp )
        mg_apply( c, r, z );
    acc_t zTr_acc( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : zTr_acc )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                double Mr = mgp ? z( i, j, k ) : c.Mi[r.idx( i, j, k )] * r( i, j, k );
                z( i, j, k ) = Mr;
                p( i, j, k ) = Mr;
                zTr_acc.add( Mr * r( i, j, k ) );
            }
    double zTr_old = zTr_acc.value();
    do_allreduce( c, &zTr_old, 1 );

    // q0 = A p0 ; pTAp
    do_gather( c, 3 );
    acc_t pTAp_acc( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : pTAp_acc )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                double Ap = apply_A( c, p, i, j, k );
                q( i, j, k ) = Ap;
                pTAp_acc.add( p( i, j, k ) * Ap );
            }
    double pTAp = pTAp_acc.value();
    do_allreduce( c, &pTAp, 1 );

    bool converged = false;
    while ( c.num_iter < max_iter )
    {
        // kernel 1: x += alpha p ; r -= alpha q ; rr
        const double alpha = zTr_old / pTAp;
        rr_acc = acc_t( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : rr_acc )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    x( i, j, k ) = std::fma( alpha, p( i, j, k ), x( i, j, k ) );
                    double r_new = std::fma( -alpha, q( i, j, k ), r( i, j, k ) );
                    r( i, j, k ) = r_new;
                    rr_acc.add( r_new * r_new );
                }
        rr = rr_acc.value();
        do_allreduce( c, &rr, 1 );
        c.resid = std::sqrt( rr );
        ++c.num_iter;
        c.hist.push_back( c.resid );
        if ( c.cfg.cg_print_level == 2 && c.cfg.world_rank == 0 )
            std::printf( "Cajita CG Iteration %d: |r|_2 = %g\n", c.num_iter, c.resid );
        if ( !fixed && c.resid <= thresh )
        {
            converged = true;
            break;
        }

        // kernel 2: z = M r ; zTr
        do_gather( c, 2 );
        if ( mgp )
            mg_apply( c, r, z );
        zTr_acc = acc_t( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : zTr_acc )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    double Mr = mgp ? z( i, j, k ) : c.Mi[r.idx( i, j, k )] * r( i, j, k );
                    z( i, j, k ) = Mr;
                    zTr_acc.add( Mr * r( i, j, k ) );
                }
        double zTr_new = zTr_acc.value();
        do_allreduce( c, &zTr_new, 1 );

        // kernel 3: p = z + beta p
        const double beta = zTr_new / zTr_old;
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                    p( i, j, k ) = std::fma( beta, p( i, j, k ), z( i, j, k ) );

        // kernel 4: q = A p ; pTAp
        do_gather( c, 3 );
        pTAp_acc = acc_t( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : pTAp_acc )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    double Ap = apply_A( c, p, i, j, k );
                    q( i, j, k ) = Ap;
                    pTAp_acc.add( p( i, j, k ) * Ap );
                }
        pTAp = pTAp_acc.value();
        do_allreduce( c, &pTAp, 1 );
        zTr_old = zTr_new;
    }
    // Multigrid preconditioner on the singular (all-SOLID) operator: pin the null-space component of x
    // to the one the diagonal preconditioner produces.  Jacobi-PCG from x0 = 0 keeps x in
    // D^-1 range(A), i.e. sum_i d_i x_i = 0 (d = diagonal of A), whereas a V-cycle does not.  The
    // constant is not arbitrary for the reference's results: quirk Q1 leaks it into v on the y walls
    // (src/VelocityCorrector.hpp:260).  x -= (sum d_i x_i / sum d_i) restores it.
    if ( mgp )
    {
        bool singular = true;
        for ( int d = 0; d < c.D; ++d )
            singular = singular && c.cfg.boundary_type[d] == CFB_SOLID && c.cfg.boundary_type[c.D + d] == CFB_SOLID;
        if ( singular )
        {
            const MgLevel& L = c.mg->lv[0];
            acc_t dx_acc( c.accum_exact ), d_acc( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : dx_acc, d_acc )
            for ( int k = s.lo[2]; k < s.hi[2]; ++k )
                for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                    for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                    {
                        const double dg = L.diag[L.walls( i - s.lo[0], j - s.lo[1], k - s.lo[2] )];
                        dx_acc.add( dg * x( i, j, k ) );
                        d_acc.add( dg );
                    }
            const double shift = dx_acc.value() / d_acc.value();
#pragma omp parallel for collapse( 2 ) schedule( static )
            for ( int k = s.lo[2]; k < s.hi[2]; ++k )
                for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                    for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                        x( i, j, k ) = x( i, j, k ) - shift;
        }
    }
    c.cg_total += c.num_iter;
    if ( c.cfg.cg_print_level > 0 && c.cfg.world_rank == 0 )
        std::printf( "Cajita CG Finished in %d iterations, |r|_2 = %g\n", c.num_iter, c.resid );
    c.t_phase[3] += std::chrono::duration<double>( clk::now() - t0 ).count();
    if ( !converged && !fixed )
    {
        c.err = "Cajita CG solver did not converge";
        return CFB_ERR_NOT_CONVERGED;
    }
    return CFB_OK;
}

// Opt-in single-reduction form of the same Jacobi-PCG (NOT the reference's loop: Chronopoulos and
// Gear's rearrangement, "s-step iterative methods for symmetric linear systems", J. Comput. Appl. Math. 25, 1989).
// In exact arithmetic the iterates equal those of cg_solve; per iteration there is ONE point where global sums are
// needed (r.r, r.u, w.u together) instead of three (two in the product's default form):
//     u = M^-1 r ; w = A u ; gamma = r.u ; delta = w.u ; [r.r]                    <- the one reduction
//     beta = gamma / gamma_old ; alpha = gamma / ( delta - beta gamma / alpha_old )   (first: beta = 0, alpha = gamma / delta)
//     p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s
// Statement order and fma placement are the product's (kernels_cg1.cu); single block (the r halo a decomposed run
// needs before w = A u is gather id 2).
int cg_solve_single_reduction( cfo_ctx& c )
{
    auto t0 = clk::now();
    Space sp = own_space( c, 0 );
    const Arr& b = c.rhs;
    Arr& x = c.lhs;
    Arr &p = c.cg_p, &u = c.cg_z, &r = c.cg_r, &w = c.cg_q, &s = c.cg_s;
    if ( s.a.size() != r.a.size() )
        s.alloc( r.e ); // same ghosted extents as the other CG vectors
    const double tol = c.cfg.cg_tolerance;
    const int max_iter = c.cfg.cg_fixed_iters > 0 ? c.cfg.cg_fixed_iters : c.cfg.cg_max_iter;
    const bool fixed = c.cfg.cg_fixed_iters > 0;
    c.num_iter = 0;
    c.hist.clear();
    double thresh = tol;
    // x0 = 0 is the product's contract for this form: r0 = b
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = sp.lo[2]; k < sp.hi[2]; ++k )
        for ( int j = sp.lo[1]; j < sp.hi[1]; ++j )
            for ( int i = sp.lo[0]; i < sp.hi[0]; ++i )
            {
                x( i, j, k ) = 0.0;
                r( i, j, k ) = b( i, j, k );
            }
    double gamma_old = 0.0, alpha = 0.0, beta = 0.0;
    bool converged = false;
    for ( ;; )
    {
        // u = M^-1 r (ghost cells of a single block: zero) ; w = A u ; the three sums
        do_gather( c, 2 );
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = sp.lo[2]; k < sp.hi[2]; ++k )
            for ( int j = sp.lo[1]; j < sp.hi[1]; ++j )
                for ( int i = sp.lo[0]; i < sp.hi[0]; ++i )
                    u( i, j, k ) = c.Mi[r.idx( i, j, k )] * r( i, j, k );
        acc_t rr_acc( c.accum_exact ), g_acc( c.accum_exact ), d_acc( c.accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : rr_acc, g_acc, d_acc )
        for ( int k = sp.lo[2]; k < sp.hi[2]; ++k )
            for ( int j = sp.lo[1]; j < sp.hi[1]; ++j )
                for ( int i = sp.lo[0]; i < sp.hi[0]; ++i )
                {
                    const double Au = apply_A( c, u, i, j, k );
                    w( i, j, k ) = Au;
                    const double rv = r( i, j, k ), uv = u( i, j, k );
                    rr_acc.add( rv * rv );
                    g_acc.add( uv * rv );
                    d_acc.add( uv * Au );
                }
        double sums[3] = { rr_acc.value(), g_acc.value(), d_acc.value() };
        do_allreduce( c, sums, 3 );
        const double rr = sums[0], gamma = sums[1], delta = sums[2];
        c.resid = std::sqrt( rr );
        if ( c.num_iter == 0 )
        {
            if ( c.cfg.cg_stop_rule == CFB_STOP_REL )
                thresh = tol * c.resid;
            if ( !fixed && c.resid <= thresh )
            {
                converged = true;
                break;
            }
            beta = 0.0;
            alpha = gamma / delta;
        }
        else
        {
            c.hist.push_back( c.resid );
            if ( c.cfg.cg_print_level == 2 && c.cfg.world_rank == 0 )
                std::printf( "Cajita CG Iteration %d: |r|_2 = %g\n", c.num_iter, c.resid );
            if ( !fixed && c.resid <= thresh )
            {
                converged = true;
                break;
            }
            if ( c.num_iter >= max_iter )
                break;
            beta = gamma / gamma_old;
            alpha = gamma / ( delta - ( beta * gamma ) / alpha );
        }
        gamma_old = gamma;
        const bool first = c.num_iter == 0;
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = sp.lo[2]; k < sp.hi[2]; ++k )
            for ( int j = sp.lo[1]; j < sp.hi[1]; ++j )
                for ( int i = sp.lo[0]; i < sp.hi[0]; ++i )
                {
                    const double pv = first ? u( i, j, k ) : std::fma( beta, p( i, j, k ), u( i, j, k ) );
                    const double sv = first ? w( i, j, k ) : std::fma( beta, s( i, j, k ), w( i, j, k ) );
                    p( i, j, k ) = pv;
                    s( i, j, k ) = sv;
                    x( i, j, k ) = std::fma( alpha, pv, x( i, j, k ) );
                    r( i, j, k ) = std::fma( -alpha, sv, r( i, j, k ) );
                }
        ++c.num_iter;
    }
    c.cg_total += c.num_iter;
    if ( c.cfg.cg_print_level > 0 && c.cfg.world_rank == 0 )
        std::printf( "Cajita CG Finished in %d iterations, |r|_2 = %g\n", c.num_iter, c.resid );
    c.t_phase[3] += std::chrono::duration<double>( clk::now() - t0 ).count();
    if ( !converged && !fixed )
    {
        c.err = "Cajita CG solver did not converge";
        return CFB_ERR_NOT_CONVERGED;
    }
    return CFB_OK;
}

// VelocityCorrector::_applyPressure  src/VelocityCorrector.hpp:214-264
void apply_pressure( cfo_ctx& c )
{
    auto t0 = clk::now();
    const double scale = c.dt / ( c.cfg.density * c.cell ); // :217
    do_gather( c, 4 ); // _pressure_halo->gather( lhs )  :236
    const Arr& p = c.lhs;
    Arr& u = c.fld[1][c.cur[1]];
    for ( int d = 0; d < c.D; ++d )
    {
        Arr& f = c.fld[1 + d][c.cur[1 + d]];
        Space s = own_space( c, 1 + d );
        const int di = d == 0, dj = d == 1, dk = d == 2;
        // Q1 (src/VelocityCorrector.hpp:260): the FaceJ kernel hands `u` to the boundary
        // functor, so the FaceJ wall test zeroes u(i,j) instead of v(i,j).
        const bool q1 = ( d == 1 ) && c.cfg.quirk_applypressure_bc;
        // The quirk writes u(i, j) for J-face indices; rows of different j never alias, and
        // within this loop nest only u (not f == v) is written by the functor: race-free.
#pragma omp parallel for collapse( 2 ) schedule( static )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    f( i, j, k ) -= scale * ( p( i, j, k ) - p( i - di, j - dj, k - dk ) );
                    int g[3] = { l2g( c, 0, i ), l2g( c, 1, j ), c.D == 3 ? l2g( c, 2, k ) : 0 };
                    if ( q1 )
                        bc_face( c, 1, u, g, i, j, k );
                    else
                        bc_face( c, d, f, g, i, j, k );
                }
    }
    c.t_phase[4] += std::chrono::duration<double>( clk::now() - t0 ).count();
}

// VelocityCorrector::correctVelocity  src/VelocityCorrector.hpp:266-282
int correct_velocity( cfo_ctx& c )
{
    build_rhs( c );
    int rc = cg_solve( c );
    if ( rc != CFB_OK )
        return rc;
    apply_pressure( c );
    return CFB_OK;
}

int field_arr( cfo_ctx& c, int field, int version, Arr** out )
{
    if ( field >= CFB_QUANTITY && field <= CFB_W )
    {
        if ( field > c.D )
            return CFB_ERR_INVALID;
        int v = ( version == CFB_CURRENT ) ? c.cur[field] : 1 - c.cur[field];
        *out = &c.fld[field][v];
        return CFB_OK;
    }
    switch ( field )
    {
    case CFB_PRESSURE:
        *out = &c.lhs;
        return CFB_OK;
    case CFB_RHS:
        *out = &c.rhs;
        return CFB_OK;
    case CFB_CG_R:
        *out = &c.cg_r;
        return CFB_OK;
    case CFB_CG_P:
        *out = &c.cg_p;
        return CFB_OK;
    case CFB_CG_Q:
        *out = &c.cg_q;
        return CFB_OK;
    }
    return CFB_ERR_INVALID;
}
inline int field_entity( int field ) { return ( field >= CFB_U && field <= CFB_W ) ? field : 0; }

} // namespace

extern "C" {

const char* cfo_last_error( const cfo_ctx* c ) { return c ? c->err.c_str() : g_err.c_str(); }

int cfo_create( const cfb_config* cfg, cfo_ctx** out )
{
    *out = nullptr;
    if ( !cfg || cfg->struct_size != (int)sizeof( cfb_config ) || ( cfg->dim != 2 && cfg->dim != 3 ) )
    {
        g_err = "invalid config";
        return CFB_ERR_INVALID;
    }
    cfo_ctx* c = new cfo_ctx();
    c->cfg = *cfg;
    c->D = cfg->dim;
    c->h = cfg->halo_cell_width;
    c->time = 0.0;
    const int D = c->D;

    // Mesh ctor  src/Mesh.hpp:41-103
    c->cell = ( cfg->global_bounding_box[3] - cfg->global_bounding_box[0] ) /
              cfg->global_num_cell[0]; // :50-51
    for ( int d = 0; d < D; ++d )
    {
        double extent = cfg->global_num_cell[d] * c->cell;
        if ( std::abs( extent - ( cfg->global_bounding_box[3 + d] - cfg->global_bounding_box[d] ) ) >
             10.0 * std::numeric_limits<double>::epsilon() ) // :56-64
        {
            g_err = "Extent not evenly divisible by uniform cell size";
            delete c;
            return CFB_ERR_MESH_EXTENT;
        }
    }
    for ( int d = 0; d < 3; ++d )
    {
        if ( d < D )
        {
            partition( cfg->global_num_cell[d], cfg->ranks_per_dim[d], cfg->block_id[d], c->n[d],
                       c->off[d] );
            c->lo_bd[d] = cfg->block_id[d] == 0;
            c->hi_bd[d] = cfg->block_id[d] == cfg->ranks_per_dim[d] - 1;
            c->bc_min[d] = 0;                           // :74-78
            c->bc_max[d] = cfg->global_num_cell[d] - 1; // src/Solver.hpp:109-110
            // LocalMesh [Cajita-mem]: own low corner = global low + cell * global offset;
            // ghosted low corner = own low corner - halo * cell.
            c->celld[d] = ( cfg->global_bounding_box[3 + d] - cfg->global_bounding_box[d] ) / cfg->global_num_cell[d];
            double own_low = cfg->global_bounding_box[d] + c->celld[d] * c->off[d];
            c->ghost_low[d] = own_low - c->h * c->celld[d];
        }
        else
        {
            c->n[d] = 1;
            c->off[d] = 0;
            c->lo_bd[d] = c->hi_bd[d] = true;
            c->bc_min[d] = c->bc_max[d] = 0;
            c->ghost_low[d] = 0;
            c->celld[d] = c->cell;
        }
    }

    // Solver ctor dt clamp  src/Solver.hpp:96-106   (+ z components: [3D-ext])
    c->dt = cfg->delta_t;
    if ( cfg->clamp_dt )
    {
        double f2 = 0, vmax = 0;
        for ( int d = 0; d < D; ++d )
        {
            f2 += cfg->body_force[d] * cfg->body_force[d];
            vmax = std::fmax( vmax, std::fabs( cfg->inflow_velocity[d] ) );
        }
        double forcemax = std::sqrt( f2 );
        double umax = vmax + std::sqrt( forcemax * c->cell );
        if ( umax > 0 && c->dt > c->cell / umax )
            c->dt = c->cell / umax;
    }

    // ProblemManager ctor  src/ProblemManager.hpp:127-180
    for ( int ent = 0; ent <= D; ++ent )
    {
        int ext[3];
        ghost_ext( *c, ent, ext );
        c->fld[ent][0].alloc( ext );
        c->fld[ent][1].alloc( ext );
        c->cur[ent] = 0;
        // initialize( create_functor )  :186-263 with MeshInitFunc  examples/advection.cpp:382-435
        Space s = own_space( *c, ent );
        double val = ent == 0 ? cfg->init_quantity : cfg->init_velocity[ent - 1];
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                    c->fld[ent][0]( i, j, k ) = val;
    }

    // VelocityCorrector ctor  src/VelocityCorrector.hpp:70-114
    {
        int ext[3];
        ghost_ext( *c, 0, ext );
        c->lhs.alloc( ext );
        c->rhs.alloc( ext );
        c->cg_p.alloc( ext );
        c->cg_z.alloc( ext );
        c->cg_r.alloc( ext );
        c->cg_q.alloc( ext );
        c->nst = 2 * D + 1;
        fill_matrix( *c );
    }
    *out = c;
    return CFB_OK;
}

int cfo_destroy( cfo_ctx* c )
{
    if ( c )
        delete c->mg;
    delete c;
    return CFB_OK;
}

int cfo_set_callbacks( cfo_ctx* c, gather_cb g, allreduce_cb a, void* user )
{
    c->gcb = g;
    c->acb = a;
    c->cb_user = user;
    return CFB_OK;
}

int cfo_get_scalars( const cfo_ctx* c, double* cell, double* dt, double* time )
{
    if ( cell )
        *cell = c->cell;
    if ( dt )
        *dt = c->dt;
    if ( time )
        *time = c->time;
    return CFB_OK;
}

int cfo_owned_extent( const cfo_ctx* c, int field, int ext[3] )
{
    Space s = own_space( *c, field_entity( field ) );
    for ( int d = 0; d < 3; ++d )
        ext[d] = s.hi[d] - s.lo[d];
    return CFB_OK;
}
int cfo_global_offset( const cfo_ctx* c, int off[3] )
{
    for ( int d = 0; d < 3; ++d )
        off[d] = c->off[d];
    return CFB_OK;
}

// Borrowed pointer to the ghosted array (reference layout: local ghosted indices).
int cfo_field_ptr( cfo_ctx* c, int field, int version, double** ptr, int ext[3] )
{
    Arr* a;
    int rc = field_arr( *c, field, version, &a );
    if ( rc )
        return rc;
    *ptr = a->a.data();
    for ( int d = 0; d < 3; ++d )
        ext[d] = a->e[d];
    return CFB_OK;
}

int cfo_upload( cfo_ctx* c, int field, int version, int region, const double* host )
{
    Arr* a;
    int rc = field_arr( *c, field, version, &a );
    if ( rc )
        return rc;
    if ( region == CFB_GHOSTED )
    {
        std::memcpy( a->a.data(), host, a->a.size() * sizeof( double ) );
        return CFB_OK;
    }
    Space s = own_space( *c, field_entity( field ) );
    size_t n = 0;
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                ( *a )( i, j, k ) = host[n++];
    return CFB_OK;
}
int cfo_download( cfo_ctx* c, int field, int version, int region, double* host )
{
    Arr* a;
    int rc = field_arr( *c, field, version, &a );
    if ( rc )
        return rc;
    if ( region == CFB_GHOSTED )
    {
        std::memcpy( host, a->a.data(), a->a.size() * sizeof( double ) );
        return CFB_OK;
    }
    Space s = own_space( *c, field_entity( field ) );
    size_t n = 0;
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                host[n++] = ( *a )( i, j, k );
    return CFB_OK;
}

int cfo_advance( cfo_ctx* c, int field )
{
    if ( field < 0 || field > c->D )
        return CFB_ERR_INVALID;
    c->cur[field] = 1 - c->cur[field];
    return CFB_OK;
}

int cfo_add_inputs( cfo_ctx* c )
{
    add_inputs( *c );
    return CFB_OK;
}
int cfo_time_integrator_step( cfo_ctx* c )
{
    time_integrator_step( *c );
    return CFB_OK;
}
int cfo_build_rhs( cfo_ctx* c )
{
    build_rhs( *c );
    return CFB_OK;
}
int cfo_pcg_solve( cfo_ctx* c, int* num_iter, double* resid )
{
    int rc = cg_solve( *c );
    if ( num_iter )
        *num_iter = c->num_iter;
    if ( resid )
        *resid = c->resid;
    return rc;
}
int cfo_apply_pressure( cfo_ctx* c )
{
    apply_pressure( *c );
    return CFB_OK;
}
int cfo_correct_velocity( cfo_ctx* c, int* num_iter, double* resid )
{
    int rc = correct_velocity( *c );
    if ( num_iter )
        *num_iter = c->num_iter;
    if ( resid )
        *resid = c->resid;
    return rc;
}
// Solver::setup  src/Solver.hpp:125-133
int cfo_setup( cfo_ctx* c )
{
    add_inputs( *c );
    return correct_velocity( *c );
}
// Solver::step  src/Solver.hpp:135-147
int cfo_step( cfo_ctx* c )
{
    time_integrator_step( *c );
    add_inputs( *c );
    int rc = correct_velocity( *c );
    c->time += c->dt;
    c->steps++;
    return rc;
}
// Solver::solve  src/Solver.hpp:149-177 (Silo output omitted)
int cfo_solve( cfo_ctx* c, double t_final, int write_freq, int* steps_taken )
{
    int t = 0;
    int num_step = (int)( t_final / c->dt );
    int rc = cfo_setup( c );
    if ( rc )
        return rc;
    do
    {
        if ( c->cfg.world_rank == 0 && write_freq > 0 && 0 == t % write_freq )
            std::printf( "Step %d / %d at time = %f\n", t, num_step, c->time );
        rc = cfo_step( c );
        if ( rc )
            return rc;
        t++;
    } while ( c->time < t_final );
    if ( steps_taken )
        *steps_taken = t;
    return CFB_OK;
}

// q = A p on the CG work vectors + sum(p*q): CG kernel 4 in isolation.
int cfo_stencil_dot( cfo_ctx* c, int reps, double* dot, double* ms_per_launch )
{
    Space s = own_space( *c, 0 );
    const Arr& p = c->cg_p;
    Arr& q = c->cg_q;
    acc_t pTAp( c->accum_exact );
    auto t0 = clk::now();
    for ( int it = 0; it < reps; ++it )
    {
        pTAp = acc_t( c->accum_exact );
#pragma omp parallel for collapse( 2 ) schedule( static ) reduction( accsum : pTAp )
        for ( int k = s.lo[2]; k < s.hi[2]; ++k )
            for ( int j = s.lo[1]; j < s.hi[1]; ++j )
                for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                {
                    double Ap = apply_A( *c, p, i, j, k );
                    q( i, j, k ) = Ap;
                    pTAp.add( p( i, j, k ) * Ap );
                }
    }
    double sec = std::chrono::duration<double>( clk::now() - t0 ).count();
    if ( dot )
        *dot = pTAp.value();
    if ( ms_per_launch )
        *ms_per_launch = 1e3 * sec / std::max( reps, 1 );
    return CFB_OK;
}

int cfo_residual_history( const cfo_ctx* c, double* hist, int n, int* count )
{
    int m = std::min<int>( n, (int)c->hist.size() );
    for ( int i = 0; i < m; ++i )
        hist[i] = c->hist[i];
    if ( count )
        *count = (int)c->hist.size();
    return CFB_OK;
}

int cfo_get_stats( const cfo_ctx* c, cfb_stats* out )
{
    std::memset( out, 0, sizeof( *out ) );
    out->ms_advect = 1e3 * c->t_phase[0];
    out->ms_add_inputs = 1e3 * c->t_phase[1];
    out->ms_build_rhs = 1e3 * c->t_phase[2];
    out->ms_pcg = 1e3 * c->t_phase[3];
    out->ms_apply_pressure = 1e3 * c->t_phase[4];
    out->cg_iterations = c->cg_total;
    out->steps = c->steps;
    return CFB_OK;
}
int cfo_reset_stats( cfo_ctx* c )
{
    for ( double& t : c->t_phase )
        t = 0;
    c->cg_total = 0;
    c->steps = 0;
    return CFB_OK;
}

// 1 (default): exact double-double sums; 0: plain double sums like the reference (CPU baseline timing)
int cfo_set_accumulation( cfo_ctx* c, int exact )
{
    c->accum_exact = exact != 0;
    return CFB_OK;
}

// Opt-in preconditioner of the product (cfb_set_preconditioner): kind 0 = the reference's diagonal,
// 1 = multigrid V(nu_pre, nu_post) cycle with nu_coarse sweeps on the coarsest level; omega <= 0 picks
// the default damping (6/7 in 3-D, 0.8 in 2-D).  Single block only.
int cfo_set_preconditioner( cfo_ctx* c, int kind, int nu_pre, int nu_post, int nu_coarse, double omega )
{
    if ( kind != 0 && kind != 1 )
    {
        c->err = "unknown preconditioner";
        return CFB_ERR_INVALID;
    }
    if ( kind == 1 && ( c->cfg.world_size > 1 || nu_pre < 1 || nu_post < 0 || nu_coarse < 1 || omega >= 2.0 ) )
    {
        c->err = "multigrid preconditioner: single block, nu_pre >= 1, nu_post >= 0, nu_coarse >= 1, omega < 2";
        return CFB_ERR_INVALID;
    }
    c->precond = kind;
    if ( kind == 1 )
    {
        c->mg_nu1 = nu_pre;
        c->mg_nu2 = nu_post;
        c->mg_nuc = nu_coarse;
        c->mg_omega = omega;
        mg_build( *c );
    }
    return CFB_OK;
}

// Cap on the number of multigrid levels (0 = as many as the grid allows).  A block-decomposed run of the
// product coarsens while every BLOCK stays even; the single-block checker is given the same depth.
int cfo_set_mg_max_levels( cfo_ctx* c, int max_levels )
{
    c->mg_max_levels = max_levels;
    if ( c->precond == 1 )
        mg_build( *c );
    return CFB_OK;
}
int cfo_mg_num_levels( cfo_ctx* c, int* levels )
{
    *levels = c->mg ? (int)c->mg->lv.size() : 0;
    return CFB_OK;
}

// z = M^-1 r for dense owned-cell host arrays: one V-cycle on its own (introspection / tests)
int cfo_mg_apply( cfo_ctx* c, const double* r_host, double* z_host )
{
    if ( c->precond != 1 || !c->mg )
    {
        c->err = "multigrid preconditioner not set";
        return CFB_ERR_INVALID;
    }
    const Space s = own_space( *c, 0 );
    const int ex = s.hi[0] - s.lo[0], ey = s.hi[1] - s.lo[1];
    Arr &r = c->cg_r, &z = c->cg_z;
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                r( i, j, k ) = r_host[( (size_t)( k - s.lo[2] ) * ey + ( j - s.lo[1] ) ) * ex + ( i - s.lo[0] )];
    mg_apply( *c, r, z );
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
                z_host[( (size_t)( k - s.lo[2] ) * ey + ( j - s.lo[1] ) ) * ex + ( i - s.lo[0] )] = z( i, j, k );
    return CFB_OK;
}

int cfo_num_threads( void )
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// 0: Cajita's ReferenceConjugateGradient loop (default) ; 1: the single-reduction form (the product's cg_variant 3)
int cfo_set_cg_algorithm( cfo_ctx* c, int algorithm )
{
    if ( algorithm != 0 && algorithm != 1 )
        return CFB_ERR_INVALID;
    c->cg_algorithm = algorithm;
    return CFB_OK;
}

// Fixed-iteration mode of the following solves (0: back to the stopping test); the benchmark's CPU arm times solves
// of two lengths on one context to separate the per-solve set-up from the per-iteration cost.
int cfo_set_fixed_iters( cfo_ctx* c, int iters )
{
    c->cfg.cg_fixed_iters = iters > 0 ? iters : 0;
    return CFB_OK;
}

// Number of OpenMP threads for everything that follows (n < 1: all the cores the process may run on).  The
// benchmark's CPU arm calls it: a launcher such as torchrun exports OMP_NUM_THREADS=1, which would time this
// restatement of the reference on a single core.
int cfo_set_num_threads( int n )
{
#ifdef _OPENMP
    if ( n < 1 )
        n = omp_get_num_procs();
    omp_set_num_threads( n );
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// Matrix / preconditioner read-back for the known-answer tests (ghosted layout).
int cfo_matrix_ptr( cfo_ctx* c, double** A, double** Minv, int* nst )
{
    *A = c->A.data();
    *Minv = c->Mi.data();
    *nst = c->nst;
    return CFB_OK;
}

// Expose the spline weights and a point interpolation for known-answer tests.
int cfo_spline_weights( int order, double xl, int* s, double* w )
{
    if ( order == 1 )
        spline1( xl, s, w );
    else if ( order == 3 )
        spline3( xl, s, w );
    else
        return CFB_ERR_INVALID;
    return CFB_OK;
}
int cfo_interpolate( cfo_ctx* c, int field, int order, const double* loc, double* value )
{
    Arr* a;
    int rc = field_arr( *c, field, CFB_CURRENT, &a );
    if ( rc )
        return rc;
    double l[3] = { loc[0], loc[1], c->D == 3 ? loc[2] : 0.0 };
    *value = order == 1 ? interpolate_field<1>( *c, field_entity( field ), l, *a )
                        : interpolate_field<3>( *c, field_entity( field ), l, *a );
    return CFB_OK;
}
int cfo_rk3( cfo_ctx* c, const double* x0, double* trace )
{
    double a[3] = { x0[0], x0[1], c->D == 3 ? x0[2] : 0.0 }, t[3] = { 0, 0, 0 };
    rk3( *c, a, c->dt, t );
    for ( int d = 0; d < c->D; ++d )
        trace[d] = t[d];
    return CFB_OK;
}
// SiloWriter::writeFile  src/SiloWriter.hpp:56-197: what the reference hands to Silo for this block.
//   :109-123  node coordinates of the owned cells, per dim: coordinates( Node(), {0,..,i,..,0} )[d]
//             for i = cell_domain.min(d) .. cell_domain.max(d) inclusive (extent + 1 nodes)
//   :136-156  owned quantity (ghosts dropped)
//   :172-186  cell-centred velocity: coordinates( Cell(), idx ) -> interpolateVelocity<D,1>
// No gather precedes it in the reference; the samples that fall on ghost entities carry weight
// (1 - f) or f with f = 0 up to rounding.  Dense x-fastest outputs: quantity[nz][ny][nx],
// velocity[D][nz][ny][nx], nodes_d[n_d + 1]; NULL skips an output.
int cfo_output_extract( cfo_ctx* c, double* quantity, double* velocity, double* nodes_x, double* nodes_y,
                        double* nodes_z )
{
    const Space s = own_space( *c, 0 );
    const int ex = s.hi[0] - s.lo[0], ey = s.hi[1] - s.lo[1], ez = s.hi[2] - s.lo[2];
    const size_t ncell = (size_t)ex * ey * ez;
    double* nodes[3] = { nodes_x, nodes_y, nodes_z };
    for ( int d = 0; d < c->D; ++d )
    {
        if ( !nodes[d] )
            continue;
        for ( int i = s.lo[d]; i < s.hi[d] + 1; ++i )
            nodes[d][i - s.lo[d]] = c->ghost_low[d] + double( i ) * c->celld[d]; // Node: every dim is "normal"
    }
    const Arr& q = c->fld[0][c->cur[0]];
#pragma omp parallel for collapse( 2 ) schedule( static )
    for ( int k = s.lo[2]; k < s.hi[2]; ++k )
        for ( int j = s.lo[1]; j < s.hi[1]; ++j )
            for ( int i = s.lo[0]; i < s.hi[0]; ++i )
            {
                const size_t o = ( (size_t)( k - s.lo[2] ) * ey + ( j - s.lo[1] ) ) * ex + ( i - s.lo[0] );
                if ( quantity )
                    quantity[o] = q( i, j, k );
                if ( velocity )
                {
                    const int idx[3] = { i, j, k };
                    double loc[3] = { 0, 0, 0 }, vel[3] = { 0, 0, 0 };
                    coordinates( *c, 0, idx, loc );
                    interpolate_velocity( *c, loc, vel );
                    for ( int d = 0; d < c->D; ++d )
                        velocity[(size_t)d * ncell + o] = vel[d];
                }
            }
    return CFB_OK;
}

int cfo_coordinates( cfo_ctx* c, int field, const int* idx, double* x )
{
    int id[3] = { idx[0], idx[1], c->D == 3 ? idx[2] : 0 };
    double xx[3] = { 0, 0, 0 };
    coordinates( *c, field_entity( field ), id, xx );
    for ( int d = 0; d < c->D; ++d )
        x[d] = xx[d];
    return CFB_OK;
}

} // extern "C"
