// kernels_fields.cu — one-pass-per-step kernels: inputs, semi-Lagrangian advection, gradient
// subtraction, synthetic fields.  FP64 throughout; built with -fmad=false so that every
// expression rounds exactly like the statement it restates (explicit fma() only where noted).
//
// Reference statements restated here (cajitafluids tree):
//   Solver::_addInputs            src/Solver.hpp:181-263  (+ InflowSource.hpp:33-78, BodyForce.hpp:33-60)
//   TimeIntegrator::rk3 / advect  src/TimeIntegrator.hpp:36-116
//   Interpolation::*              src/Interpolation.hpp:30-54  (Cajita splines, see device_geo.cuh)
//   VelocityCorrector::_applyPressure  src/VelocityCorrector.hpp:214-264
//   BoundaryCondition::operator() src/BoundaryConditions.hpp:102-129
#include "cfb_internal.h"
#include "device_geo.cuh"

namespace
{

// ---------------------------------------------------------------------------------------------
// Solver::_addInputs: one thread per (i,j,k) of the (n+1)^D box; it handles the cell and the D
// faces that carry that owned index.  Statement order per entity: source -> body -> bc.
template <int D>
__global__ void __launch_bounds__( 256 )
    add_inputs_kernel( const __grid_constant__ Geo g, const __grid_constant__ InflowConst s,
                       double* __restrict__ q, double* __restrict__ u, double* __restrict__ v,
                       double* __restrict__ w )
{
    const int bx = g.n[0] + 1, by = g.n[1] + 1;
    const int bz = D == 3 ? g.n[2] + 1 : 1;
    const long long total = (long long)bx * by * bz;
    for ( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
          t += (long long)gridDim.x * blockDim.x )
    {
        const int i = (int)( t % bx );
        const int j = (int)( ( t / bx ) % by );
        const int k = (int)( t / ( (long long)bx * by ) );
        const int idx[3] = { i, j, k };
        const long long o = geo_off( g, i, j, k );
        const int gidx[3] = { i + g.off[0], j + g.off[1], k + g.off[2] };

        // cells: InflowSource( Cell )  src/InflowSource.hpp:33-50 ; BodyForce( Cell ) is empty
        if ( i < g.n[0] && j < g.n[1] && ( D == 2 || k < g.n[2] ) )
        {
            double x[3];
            geo_coordinates<D>( g, 0, idx, x );
            if ( in_box<D>( s, x ) )
            {
                double qv = q[o];
                if ( qv < s.quantity )
                    q[o] = s.quantity;
            }
        }
#pragma unroll
        for ( int d = 0; d < D; ++d )
        {
            const bool own = idx[0] < ( d == 0 ? g.nf[0] : g.n[0] ) &&
                             idx[1] < ( d == 1 ? g.nf[1] : g.n[1] ) &&
                             ( D == 2 || idx[2] < ( d == 2 ? g.nf[2] : g.n[2] ) );
            if ( !own )
                continue;
            double* f = d == 0 ? u : ( d == 1 ? v : w );
            double x[3];
            geo_coordinates<D>( g, 1 + d, idx, x );
            double val = f[o];
            // InflowSource( Face )  src/InflowSource.hpp:52-78
            if ( in_box<D>( s, x ) && fabs( val ) < fabs( s.vel[d] ) )
                val = s.vel[d];
            // BodyForce( Face )  src/BodyForce.hpp:43-60
            val += s.force_dt[d];
            val = bc_face( g, d, gidx[d], val );
            f[o] = val;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TimeIntegrator::advect for every entity kind (0 = Cell/q, 1.. = faces) in one launch,
// i.e. the whole advection of TimeIntegrator::step.
struct AdvectArgs
{
    const double* cur[4];
    double* next[4];
};

// Work decomposition.  One block per work item, items in launch order = (chunk of the volume, entity) with the
// ENTITY FASTEST: the blocks resident at any moment work on the same few planes of the volume for all D + 1 entity
// kinds, so the velocity planes they gather from are pulled from HBM once and shared through the L2 (the round-1
// form — a capped grid-stride launch with the entity in blockIdx.y — spread its resident blocks over the whole
// volume and swept it once per entity kind: 40 GB of DRAM reads for 8.6 GB compulsory, profiles/r2_advect_*).
// An item is 128 consecutive entities of the flattened (i, j, k) index of its kind, or — TILED ("advect_tile"
// tuning key) — a 32 x 2 x 2 (3-D) / 32 x 4 (2-D) tile, whose 4^D-point neighbourhoods overlap in y and z too.
// Only the thread -> entity map differs: the same values.
// 5 resident blocks per SM (96 registers, nothing spilled).  Compiled for 6 (80 registers) and 8 (64, a few values
// spilled) it measured slower at every size: 38.2 / 40.9 / 46.7 ms at 512^3 (profiles/r2_advect_occupancy_variants.json).
template <int D, int ORDER, bool TILED>
__global__ void __launch_bounds__( 128, 5 )
    advect_kernel( const __grid_constant__ Geo g, const __grid_constant__ AdvectArgs a, int quirk_v0 )
{
    const int ent = (int)( blockIdx.x % (unsigned)( D + 1 ) );
    const long long item = (long long)( blockIdx.x / (unsigned)( D + 1 ) );
    const int ex = ent == 1 ? g.nf[0] : g.n[0];
    const int ey = ent == 2 ? g.nf[1] : g.n[1];
    const int ez = D == 3 ? ( ent == 3 ? g.nf[2] : g.n[2] ) : 1;
    const double* fc = a.cur[ent];
    double* fn = a.next[ent];
    const double dt = g.dt;
    constexpr int TY = D == 3 ? 2 : 4, TZ = D == 3 ? 2 : 1;
    {
        int i, j, k;
        if ( TILED )
        {
            const int nbx = ( ex + 31 ) / 32, nby = ( ey + TY - 1 ) / TY, nbz = ( ez + TZ - 1 ) / TZ;
            if ( item >= (long long)nbx * nby * nbz )
                return;
            const int bx = (int)( item % nbx ), by = (int)( ( item / nbx ) % nby ), bz = (int)( item / ( (long long)nbx * nby ) );
            const int tid = threadIdx.x;
            i = bx * 32 + ( tid & 31 );
            j = by * TY + ( ( tid >> 5 ) % TY );
            k = bz * TZ + ( tid >> 5 ) / TY;
            if ( i >= ex || j >= ey || k >= ez )
                return;
        }
        else
        {
            const long long t = item * 128 + threadIdx.x;
            if ( t >= (long long)ex * ey * ez )
                return;
            i = (int)( t % ex );
            j = (int)( ( t / ex ) % ey );
            k = (int)( t / ( (long long)ex * ey ) );
        }
        const int idx[3] = { i, j, k };
        double x0[3], v0[3], x1[3], v1[3], x2[3], v2[3], trace[3];
        // 1. location of the entity            src/TimeIntegrator.hpp:105
        geo_coordinates<D>( g, ent, idx, x0 );
        // 2. rk3 back-trace                    src/TimeIntegrator.hpp:36-77
        interp_velocity<D>( g, a.cur, x0, v0 );
#pragma unroll
        for ( int d = 0; d < D; ++d )
            x1[d] = x0[d] - 0.5 * dt * v0[d];
        interp_velocity<D>( g, a.cur, x1, v1 );
#pragma unroll
        for ( int d = 0; d < D; ++d )
            x2[d] = x0[d] - 0.75 * dt * ( quirk_v0 ? v0[d] : v1[d] ); // Q2
        interp_velocity<D>( g, a.cur, x2, v2 );
#pragma unroll
        for ( int d = 0; d < D; ++d )
            trace[d] = x0[d] - dt * ( ( 2.0 / 9.0 ) * v0[d] + ( 3.0 / 9.0 ) * v1[d] +
                                      ( 4.0 / 9.0 ) * v2[d] );
        // 3. sample the advected field         src/TimeIntegrator.hpp:112-114
        fn[geo_off( g, i, j, k )] = interp_field<D, ORDER>( g, ent, fc, trace );
    }
}

// ---------------------------------------------------------------------------------------------
// VelocityCorrector::_applyPressure: all velocity components in one pass.
template <int D>
__global__ void __launch_bounds__( 256 )
    apply_pressure_kernel( const __grid_constant__ Geo g, double scale, int quirk_q1,
                           const double* __restrict__ p, double* u, double* __restrict__ v,
                           double* __restrict__ w )
{
    const int bx = g.n[0] + 1, by = g.n[1] + 1;
    const int bz = D == 3 ? g.n[2] + 1 : 1;
    const long long total = (long long)bx * by * bz;
    for ( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
          t += (long long)gridDim.x * blockDim.x )
    {
        const int i = (int)( t % bx );
        const int j = (int)( ( t / bx ) % by );
        const int k = (int)( t / ( (long long)bx * by ) );
        const int idx[3] = { i, j, k };
        const long long o = geo_off( g, i, j, k );
        const int gidx[3] = { i + g.off[0], j + g.off[1], k + g.off[2] };
        const double pc = p[o];
#pragma unroll
        for ( int d = 0; d < D; ++d )
        {
            const bool own = idx[0] < ( d == 0 ? g.nf[0] : g.n[0] ) &&
                             idx[1] < ( d == 1 ? g.nf[1] : g.n[1] ) &&
                             ( D == 2 || idx[2] < ( d == 2 ? g.nf[2] : g.n[2] ) );
            if ( !own )
                continue;
            double* f = d == 0 ? u : ( d == 1 ? v : w );
            const long long st = d == 0 ? 1 : ( d == 1 ? g.sy : g.sz );
            double val = f[o];
            val -= scale * ( pc - p[o - st] ); // :245 / :256
            if ( d == 1 && quirk_q1 )
            {
                // Q1 (src/VelocityCorrector.hpp:260): bc( FaceJ(), u, ... ) — the J-face wall
                // test zeroes u(i,j), v keeps its pressure correction.  This thread already
                // finished u(i,j,k) in the d == 0 pass, which reproduces the reference's
                // "u kernel, then v kernel" order.
                f[o] = val;
                if ( bc_face( g, 1, gidx[1], 1.0 ) == 0.0 )
                    u[o] = 0.0;
            }
            else
            {
                f[o] = bc_face( g, d, gidx[d], val );
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic MAC velocity: variant 0 = sin/cos product, variant 1 = hashed U(-1,1).
__device__ __forceinline__ unsigned long long splitmix64( unsigned long long x )
{
    x += 0x9E3779B97F4A7C15ull;
    x = ( x ^ ( x >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    x = ( x ^ ( x >> 27 ) ) * 0x94D049BB133111EBull;
    return x ^ ( x >> 31 );
}

template <int D>
__global__ void __launch_bounds__( 256 )
    synthetic_velocity_kernel( const __grid_constant__ Geo g, int variant, unsigned long long seed,
                               double lo0, double lo1, double lo2, double* __restrict__ u,
                               double* __restrict__ v, double* __restrict__ w )
{
    const int bx = g.n[0] + 1, by = g.n[1] + 1;
    const int bz = D == 3 ? g.n[2] + 1 : 1;
    const long long total = (long long)bx * by * bz;
    const double lo[3] = { lo0, lo1, lo2 };
    const double PI = 3.14159265358979323846;
    for ( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
          t += (long long)gridDim.x * blockDim.x )
    {
        const int i = (int)( t % bx );
        const int j = (int)( ( t / bx ) % by );
        const int k = (int)( t / ( (long long)bx * by ) );
        const int idx[3] = { i, j, k };
        const long long o = geo_off( g, i, j, k );
        const int gidx[3] = { i + g.off[0], j + g.off[1], k + g.off[2] };
#pragma unroll
        for ( int d = 0; d < D; ++d )
        {
            const bool own = idx[0] < ( d == 0 ? g.nf[0] : g.n[0] ) &&
                             idx[1] < ( d == 1 ? g.nf[1] : g.n[1] ) &&
                             ( D == 2 || idx[2] < ( d == 2 ? g.nf[2] : g.n[2] ) );
            if ( !own )
                continue;
            double* f = d == 0 ? u : ( d == 1 ? v : w );
            double val;
            if ( variant == 0 )
            {
                val = 1.0;
                for ( int e = 0; e < D; ++e )
                {
                    // global position of the entity (decomposition independent)
                    double xe = lo[e] + ( gidx[e] + ( e == d ? 0.0 : 0.5 ) ) * g.cell;
                    double L = g.gn[e] * g.cell;
                    val *= ( e == d ) ? sin( PI * ( xe - lo[e] ) / L ) : cos( 2.0 * PI * ( xe - lo[e] ) / L );
                }
            }
            else
            {
                unsigned long long key =
                    ( ( (unsigned long long)gidx[2] * 4099ull + gidx[1] ) * 4099ull + gidx[0] ) * 4ull + d;
                unsigned long long r = splitmix64( key ^ splitmix64( seed ) );
                val = ( (double)( r >> 11 ) ) * ( 2.0 / 9007199254740992.0 ) - 1.0;
            }
            // wall-normal component exactly zero on both physical walls
            if ( gidx[d] <= 0 || gidx[d] > g.gn[d] - 1 )
                val = 0.0;
            f[o] = val;
        }
    }
}

inline int grid_for( long long total, int block, int sm_count )
{
    long long b = ( total + block - 1 ) / block;
    long long cap = (long long)sm_count * 32;
    return (int)( b < 1 ? 1 : ( b > cap ? cap : b ) );
}

} // namespace

int launch_add_inputs( cfb_ctx* c )
{
    const Geo& g = c->g;
    long long total = (long long)( g.n[0] + 1 ) * ( g.n[1] + 1 ) * ( g.D == 3 ? g.n[2] + 1 : 1 );
    int grid = grid_for( total, 256, c->sm_count );
    double* q = field_ptr( c, CFB_QUANTITY, CFB_CURRENT );
    double* u = field_ptr( c, CFB_U, CFB_CURRENT );
    double* v = field_ptr( c, CFB_V, CFB_CURRENT );
    double* w = field_ptr( c, CFB_W, CFB_CURRENT );
    if ( g.D == 2 )
        add_inputs_kernel<2><<<grid, 256, 0, c->stream>>>( g, c->inflow, q, u, v, w );
    else
        add_inputs_kernel<3><<<grid, 256, 0, c->stream>>>( g, c->inflow, q, u, v, w );
    return 1;
}

int launch_advect( cfb_ctx* c )
{
    const Geo& g = c->g;
    AdvectArgs a{};
    for ( int e = 0; e <= g.D && e < 4; ++e ) // (D <= 3; the second bound is for the compiler's range analysis)
    {
        a.cur[e] = field_ptr( c, e, CFB_CURRENT );
        a.next[e] = field_ptr( c, e, CFB_NEXT );
    }
    const int qv0 = c->cfg.quirk_rk3_stage3_v0;
    const int order = c->cfg.field_interp_order;
    // items per entity kind (the largest kind decides; the others leave their surplus blocks at once)
    long long items;
    if ( c->advect_tile )
        items = (long long)( ( g.n[0] + 1 + 31 ) / 32 ) * ( ( g.n[1] + 1 + ( g.D == 3 ? 1 : 3 ) ) / ( g.D == 3 ? 2 : 4 ) ) *
                ( g.D == 3 ? ( g.n[2] + 1 + 1 ) / 2 : 1 );
    else
    {
        // the largest flattened entity count among cells and faces
        long long mx = 0;
        for ( int e = 0; e <= g.D; ++e )
        {
            long long t = 1;
            for ( int d = 0; d < g.D; ++d )
                t *= ( e - 1 == d ) ? g.nf[d] : g.n[d];
            mx = t > mx ? t : mx;
        }
        items = ( mx + 127 ) / 128;
    }
    const long long blocks = items * ( g.D + 1 );
    if ( blocks > 2147483647ll )
    {
        note_rc( c, cfb_fail( c, CFB_ERR_INVALID, "advection: block too large for one launch" ) );
        return 0;
    }
    const dim3 grid( (unsigned)blocks );
    const int key = ( g.D == 3 ? 100 : 0 ) + ( order == 3 ? 10 : 0 ) + ( c->advect_tile ? 1 : 0 );
#define CFB_ADVECT( D_, O_, T_ )                                                                                          \
    do                                                                                                                    \
    {                                                                                                                     \
        advect_kernel<D_, O_, T_><<<grid, 128, 0, c->stream>>>( g, a, qv0 );                                              \
    } while ( 0 )
    switch ( key )
    {
    case 0:
        CFB_ADVECT( 2, 1, false );
        break;
    case 1:
        CFB_ADVECT( 2, 1, true );
        break;
    case 10:
        CFB_ADVECT( 2, 3, false );
        break;
    case 11:
        CFB_ADVECT( 2, 3, true );
        break;
    case 100:
        CFB_ADVECT( 3, 1, false );
        break;
    case 101:
        CFB_ADVECT( 3, 1, true );
        break;
    case 110:
        CFB_ADVECT( 3, 3, false );
        break;
    default:
        CFB_ADVECT( 3, 3, true );
        break;
    }
#undef CFB_ADVECT
    return 1;
}

int launch_apply_pressure( cfb_ctx* c )
{
    const Geo& g = c->g;
    long long total = (long long)( g.n[0] + 1 ) * ( g.n[1] + 1 ) * ( g.D == 3 ? g.n[2] + 1 : 1 );
    int grid = grid_for( total, 256, c->sm_count );
    const double scale = g.dt / ( c->cfg.density * g.cell ); // src/VelocityCorrector.hpp:217
    double* u = field_ptr( c, CFB_U, CFB_CURRENT );
    double* v = field_ptr( c, CFB_V, CFB_CURRENT );
    double* w = field_ptr( c, CFB_W, CFB_CURRENT );
    if ( g.D == 2 )
        apply_pressure_kernel<2><<<grid, 256, 0, c->stream>>>( g, scale, c->cfg.quirk_applypressure_bc,
                                                               c->lhs, u, v, w );
    else
        apply_pressure_kernel<3><<<grid, 256, 0, c->stream>>>( g, scale, c->cfg.quirk_applypressure_bc,
                                                               c->lhs, u, v, w );
    return 1;
}

int launch_fill_synthetic( cfb_ctx* c, int variant, uint64_t seed )
{
    const Geo& g = c->g;
    long long total = (long long)( g.n[0] + 1 ) * ( g.n[1] + 1 ) * ( g.D == 3 ? g.n[2] + 1 : 1 );
    int grid = grid_for( total, 256, c->sm_count );
    double* u = field_ptr( c, CFB_U, CFB_CURRENT );
    double* v = field_ptr( c, CFB_V, CFB_CURRENT );
    double* w = field_ptr( c, CFB_W, CFB_CURRENT );
    const double* lo = c->cfg.global_bounding_box;
    if ( g.D == 2 )
        synthetic_velocity_kernel<2><<<grid, 256, 0, c->stream>>>( g, variant, seed, lo[0], lo[1], lo[2], u, v, w );
    else
        synthetic_velocity_kernel<3><<<grid, 256, 0, c->stream>>>( g, variant, seed, lo[0], lo[1], lo[2], u, v, w );
    return 1;
}
