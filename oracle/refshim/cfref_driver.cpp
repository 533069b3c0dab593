// cfref_driver.cpp — C entry points around the UNMODIFIED cajitafluids reference (its src/*.hpp are
// compiled from where they lie under /root/reference; see Makefile.ref) built against the
// single-rank stand-ins for Kokkos / Cajita / MPI / Silo in this directory.
//
// TEST INFRASTRUCTURE ONLY.  The output, oracle/_ref/libcfref.so, is the "reference run here" that
// pins oracle/cfo_oracle.cpp (tests/test_reference_shim.py) and that produced the golden fixtures
// under tests/golden/ref_*.npz (tests/golden/make_golden_ref.py).  It exists only in a container
// that has /root/reference; nothing in the product, in `-m gpu` tests, in smoke() or in bench.py
// needs it at run time.
//
// The entry points carry the names and signatures of include/cfb.h with the prefix cfref_, so the
// same Python `Context` class drives the product (cfb_), the oracle (cfo_) and the reference
// (cfref_).  What runs behind them is the reference's own code:
//   cfref_create            -> CajitaFluids::createSolver( "serial" | "openmp", ... )  src/Solver.hpp:283-350
//   cfref_setup/step/solve  -> SolverBase::setup / step / solve                        src/Solver.hpp:125-177
//   cfref_add_inputs        -> Solver::_addInputs                                      src/Solver.hpp:181-263
//   cfref_time_integrator_step -> TimeIntegrator::step<2>                              src/TimeIntegrator.hpp:120-177
//   cfref_build_rhs / apply_pressure / correct_velocity -> VelocityCorrector::_buildRHS (+ lhs = 0),
//                              _applyPressure, correctVelocity                         src/VelocityCorrector.hpp:182-282
//   cfref_pcg_solve         -> _pressure_solver->solve( *_rhs, *_lhs )                 src/VelocityCorrector.hpp:276
// The reference keeps its state private; this file is compiled with -fno-access-control instead of
// editing or macro-patching the reference.
#include "../../include/cfb.h"

#include <Solver.hpp>

#include <cstring>
#include <memory>
#include <string>

#include <silo.h>

namespace
{

using namespace CajitaFluids;
using MemorySpace = Kokkos::HostSpace;
using mesh2 = Cajita::UniformMesh<double, 2>;
using cg_type = Cajita::ReferenceConjugateGradient<double, Cajita::Cell, mesh2, MemorySpace>;
using ref_solver_type = Cajita::ReferenceStructuredSolver<double, Cajita::Cell, mesh2, MemorySpace>;
using Cell = Cajita::Cell;
using FaceI = Cajita::Face<Cajita::Dim::I>;
using FaceJ = Cajita::Face<Cajita::Dim::J>;

std::string g_err;

// constant initial state, the role of MeshInitFunc in examples/advection.cpp:382-435
struct ConstantInit
{
    double q, u[2];
    bool operator()( Cell, Field::Quantity, const int*, const double*, double& v ) const
    {
        v = q;
        return true;
    }
    bool operator()( FaceI, Field::Velocity, const int*, const double*, double& v ) const
    {
        v = u[0];
        return true;
    }
    bool operator()( FaceJ, Field::Velocity, const int*, const double*, double& v ) const
    {
        v = u[1];
        return true;
    }
};

// type-erased access to the pieces of Solver<2, Exec, HostSpace>
struct Handle
{
    virtual ~Handle() = default;
    virtual void addInputs() = 0;
    virtual void timeIntegratorStep() = 0;
    virtual void buildRHS() = 0;
    virtual void pcg() = 0;
    virtual void applyPressure() = 0;
    virtual void correctVelocity() = 0;
    virtual void advance( int field ) = 0;
    virtual double& time() = 0;
    virtual double dt() = 0;
    virtual double cellSize() = 0;
    virtual int halo() = 0;
    // view of a field: data(i, j, 0) through a callback-free accessor
    virtual double* at( int field, int version, int i, int j ) = 0;
    virtual void ownedSpace( int field, int lo[2], int hi[2] ) = 0;
    virtual void ghostExtent( int field, int ext[2] ) = 0;
    virtual cg_type* cg() = 0;
    virtual void siloWrite( int time_step ) = 0;
    SolverBase* base = nullptr;
};

template <class Exec>
struct HandleT : Handle
{
    using solver_t = Solver<2, Exec, MemorySpace>;
    using vc_t = VelocityCorrector<2, Exec, MemorySpace, ref_solver_type>;
    std::shared_ptr<solver_t> s;
    vc_t* vc = nullptr;

    explicit HandleT( const std::shared_ptr<solver_t>& sp )
        : s( sp )
    {
        base = s.get();
        vc = dynamic_cast<vc_t*>( s->_vc.get() );
        if ( !vc )
            throw std::runtime_error( "cfref: the velocity corrector is not the Reference-solver one" );
    }
    void addInputs() override { s->_addInputs(); }
    void timeIntegratorStep() override { TimeIntegrator::step<2>( Exec(), *s->_pm, s->_dt, s->_bc ); }
    void buildRHS() override
    {
        vc->_buildRHS();
        Cajita::ArrayOp::assign( *vc->_lhs, 0.0, Cajita::Own() ); // src/VelocityCorrector.hpp:272
    }
    void pcg() override { vc->_pressure_solver->solve( *vc->_rhs, *vc->_lhs ); }
    void applyPressure() override { vc->_applyPressure(); }
    void correctVelocity() override { vc->correctVelocity(); }
    void advance( int field ) override
    {
        if ( field == CFB_QUANTITY )
            s->_pm->advance( Cell(), Field::Quantity() );
        else if ( field == CFB_U )
            s->_pm->advance( FaceI(), Field::Velocity() );
        else
            s->_pm->advance( FaceJ(), Field::Velocity() );
    }
    double& time() override { return s->_time; }
    double dt() override { return s->_dt; }
    double cellSize() override { return s->_mesh->cellSize(); }
    int halo() override { return s->_mesh->localGrid()->haloCellWidth(); }
    double* at( int field, int version, int i, int j ) override
    {
        auto& pm = *s->_pm;
        switch ( field )
        {
        case CFB_QUANTITY:
            return version == CFB_CURRENT ? &pm.get( Cell(), Field::Quantity(), Version::Current() )( i, j, 0 )
                                          : &pm.get( Cell(), Field::Quantity(), Version::Next() )( i, j, 0 );
        case CFB_U:
            return version == CFB_CURRENT ? &pm.get( FaceI(), Field::Velocity(), Version::Current() )( i, j, 0 )
                                          : &pm.get( FaceI(), Field::Velocity(), Version::Next() )( i, j, 0 );
        case CFB_V:
            return version == CFB_CURRENT ? &pm.get( FaceJ(), Field::Velocity(), Version::Current() )( i, j, 0 )
                                          : &pm.get( FaceJ(), Field::Velocity(), Version::Next() )( i, j, 0 );
        case CFB_PRESSURE:
            return &vc->_lhs->view()( i, j, 0 );
        case CFB_RHS:
            return &vc->_rhs->view()( i, j, 0 );
        }
        return nullptr;
    }
    template <class Entity>
    void spaceOf( Entity, int lo[2], int hi[2] )
    {
        auto sp = s->_mesh->localGrid()->indexSpace( Cajita::Own(), Entity(), Cajita::Local() );
        for ( int d = 0; d < 2; ++d )
        {
            lo[d] = (int)sp.min( d );
            hi[d] = (int)sp.max( d );
        }
    }
    void ownedSpace( int field, int lo[2], int hi[2] ) override
    {
        if ( field == CFB_U )
            spaceOf( FaceI(), lo, hi );
        else if ( field == CFB_V )
            spaceOf( FaceJ(), lo, hi );
        else
            spaceOf( Cell(), lo, hi );
    }
    template <class Entity>
    void ghostOf( Entity, int ext[2] )
    {
        auto sp = s->_mesh->localGrid()->indexSpace( Cajita::Ghost(), Entity(), Cajita::Local() );
        for ( int d = 0; d < 2; ++d )
            ext[d] = (int)sp.extent( d );
    }
    void ghostExtent( int field, int ext[2] ) override
    {
        if ( field == CFB_U )
            ghostOf( FaceI(), ext );
        else if ( field == CFB_V )
            ghostOf( FaceJ(), ext );
        else
            ghostOf( Cell(), ext );
    }
    cg_type* cg() override { return dynamic_cast<cg_type*>( vc->_pressure_solver.get() ); }
    // the call Solver::solve makes (src/Solver.hpp:156,170-173)
    void siloWrite( int time_step ) override { s->_silo->siloWrite( strdup( "Mesh" ), time_step, s->_time, s->_dt ); }
};

} // namespace

struct cfref_ctx
{
    cfb_config cfg;
    std::unique_ptr<Handle> h;
    std::string err;
    long long cg_total = 0;
    long long steps = 0;
    int last_iters = 0;
};

namespace
{
int fail( cfref_ctx* c, int code, const std::string& msg )
{
    if ( c )
        c->err = msg;
    g_err = msg;
    return code;
}

bool valid_field( int f ) { return f == CFB_QUANTITY || f == CFB_U || f == CFB_V || f == CFB_PRESSURE || f == CFB_RHS; }

// run a piece of the reference, translating its exceptions into the cfb status codes
template <class F>
int guarded( cfref_ctx* c, F&& f )
{
    try
    {
        f();
    }
    catch ( const std::logic_error& e )
    {
        return fail( c, CFB_ERR_MESH_EXTENT, e.what() );
    }
    catch ( const std::runtime_error& e )
    {
        const std::string w = e.what();
        return fail( c, w.find( "did not converge" ) != std::string::npos ? CFB_ERR_NOT_CONVERGED : CFB_ERR_INVALID, w );
    }
    catch ( const std::exception& e )
    {
        return fail( c, CFB_ERR_INVALID, e.what() );
    }
    return CFB_OK;
}

void account_cg( cfref_ctx* c )
{
    c->last_iters = c->h->cg()->getNumIter();
    c->cg_total += c->last_iters;
}

int copy_field( cfref_ctx* c, int field, int version, int region, double* host, bool to_ref )
{
    if ( !valid_field( field ) )
        return fail( c, CFB_ERR_INVALID, "invalid field id" );
    int lo[2], hi[2];
    if ( region == CFB_GHOSTED )
    {
        int ext[2];
        c->h->ghostExtent( field, ext );
        lo[0] = lo[1] = 0;
        hi[0] = ext[0];
        hi[1] = ext[1];
    }
    else
        c->h->ownedSpace( field, lo, hi );
    const int ex = hi[0] - lo[0];
    for ( int j = lo[1]; j < hi[1]; ++j )
        for ( int i = lo[0]; i < hi[0]; ++i )
        {
            double* p = c->h->at( field, version, i, j );
            double& hv = host[(size_t)( j - lo[1] ) * ex + ( i - lo[0] )]; // dense, x fastest (cfb.h)
            if ( to_ref )
                *p = hv;
            else
                hv = *p;
        }
    return CFB_OK;
}
} // namespace

extern "C" {

const char* cfref_last_error( const cfref_ctx* c ) { return c ? c->err.c_str() : g_err.c_str(); }

// CG arithmetic of the Cajita stand-in: 0 = plain double (default), 1 = the oracle's (bit-comparable)
int cfref_set_cg_arithmetic( int exact )
{
    cfref::knobs().cg_exact = exact ? 1 : 0;
    return CFB_OK;
}

int cfref_create( const cfb_config* cfg, cfref_ctx** out )
{
    if ( out )
        *out = nullptr;
    if ( !cfg || !out || cfg->struct_size != (int32_t)sizeof( cfb_config ) )
        return fail( nullptr, CFB_ERR_INVALID, "cfb_config size mismatch (ABI)" );
    if ( cfg->dim != 2 )
        return fail( nullptr, CFB_ERR_INVALID, "the reference is 2-D only (SURVEY.md F1)" );
    if ( cfg->world_size != 1 )
        return fail( nullptr, CFB_ERR_INVALID, "refshim: one rank only" );
    if ( cfg->field_interp_order != 3 )
        return fail( nullptr, CFB_ERR_INVALID, "the reference hard-codes order 3 (src/TimeIntegrator.hpp:113)" );
    if ( !cfg->quirk_applypressure_bc || !cfg->quirk_rk3_stage3_v0 || !cfg->clamp_dt )
        return fail( nullptr, CFB_ERR_INVALID, "the reference always has its quirks Q1, Q2 and the dt clamp" );
    if ( cfg->halo_cell_width != 3 )
        return fail( nullptr, CFB_ERR_INVALID, "the reference hard-codes halo 3 (src/Solver.hpp:78)" );
    if ( cfg->cg_stop_rule != CFB_STOP_ABS )
        return fail( nullptr, CFB_ERR_INVALID, "the reference stops on the absolute residual norm" );

    auto* c = new cfref_ctx();
    *out = c;
    c->cfg = *cfg;
    cfref::knobs().cg_print = cfg->cg_print_level;
    int rc = guarded( c, [&]() {
        Kokkos::Array<double, 4> box = { cfg->global_bounding_box[0], cfg->global_bounding_box[1],
                                         cfg->global_bounding_box[3], cfg->global_bounding_box[4] };
        std::array<int, 2> ncell = { cfg->global_num_cell[0], cfg->global_num_cell[1] };
        Cajita::DimBlockPartitioner<2> partitioner;
        BoundaryCondition<2> bc;
        bc.boundary_type = { cfg->boundary_type[0], cfg->boundary_type[1], cfg->boundary_type[2],
                             cfg->boundary_type[3] };
        InflowSource<2> source( { cfg->inflow_location[0], cfg->inflow_location[1] },
                                { cfg->inflow_size[0], cfg->inflow_size[1] },
                                { cfg->inflow_velocity[0], cfg->inflow_velocity[1] }, cfg->inflow_quantity );
        BodyForce<2> body( cfg->body_force[0], cfg->body_force[1] );
        ConstantInit init{ cfg->init_quantity, { cfg->init_velocity[0], cfg->init_velocity[1] } };
#ifdef _OPENMP
        const std::string device = "openmp";
        using Exec = Kokkos::OpenMP;
#else
        const std::string device = "serial";
        using Exec = Kokkos::Serial;
#endif
        auto sb = createSolver( device, MPI_COMM_WORLD, box, ncell, partitioner, cfg->density, init, bc, source,
                                body, cfg->delta_t, "Reference", "none" );
        auto sp = std::dynamic_pointer_cast<Solver<2, Exec, MemorySpace>>( sb );
        if ( !sp )
            throw std::runtime_error( "cfref: unexpected solver type" );
        c->h.reset( new HandleT<Exec>( sp ) );
        // the reference fixes tol 1e-6 / max_iter 2000 / print 1 in the VelocityCorrector ctor
        // (src/VelocityCorrector.hpp:103-105); other values are applied through the solver's own setters
        cg_type* cg = c->h->cg();
        if ( !cg )
            throw std::runtime_error( "cfref: unexpected pressure solver type" );
        cg->setTolerance( cfg->cg_tolerance );
        cg->setMaxIter( cfg->cg_max_iter );
        cg->setFixedIterations( cfg->cg_fixed_iters );
    } );
    return rc;
}

int cfref_destroy( cfref_ctx* c )
{
    delete c;
    return CFB_OK;
}

int cfref_get_scalars( const cfref_ctx* c, double* cell, double* dt, double* time )
{
    if ( cell )
        *cell = c->h->cellSize();
    if ( dt )
        *dt = c->h->dt();
    if ( time )
        *time = c->h->time();
    return CFB_OK;
}

int cfref_owned_extent( const cfref_ctx* c, int field, int ext[3] )
{
    int lo[2], hi[2];
    c->h->ownedSpace( field, lo, hi );
    ext[0] = hi[0] - lo[0];
    ext[1] = hi[1] - lo[1];
    ext[2] = 1;
    return CFB_OK;
}

int cfref_global_offset( const cfref_ctx*, int off[3] )
{
    off[0] = off[1] = off[2] = 0;
    return CFB_OK;
}

int cfref_upload( cfref_ctx* c, int field, int version, int region, const double* host )
{
    return copy_field( c, field, version, region, const_cast<double*>( host ), true );
}
int cfref_download( cfref_ctx* c, int field, int version, int region, double* host )
{
    return copy_field( c, field, version, region, host, false );
}

int cfref_advance( cfref_ctx* c, int field )
{
    if ( field < CFB_QUANTITY || field > CFB_V )
        return fail( c, CFB_ERR_INVALID, "advance: invalid field" );
    c->h->advance( field );
    return CFB_OK;
}

int cfref_add_inputs( cfref_ctx* c )
{
    return guarded( c, [&]() { c->h->addInputs(); } );
}
int cfref_time_integrator_step( cfref_ctx* c )
{
    return guarded( c, [&]() { c->h->timeIntegratorStep(); } );
}
int cfref_build_rhs( cfref_ctx* c )
{
    return guarded( c, [&]() { c->h->buildRHS(); } );
}
int cfref_pcg_solve( cfref_ctx* c, int* num_iter, double* resid )
{
    int rc = guarded( c, [&]() { c->h->pcg(); } );
    account_cg( c );
    if ( num_iter )
        *num_iter = c->last_iters;
    if ( resid )
        *resid = c->h->cg()->getFinalRelativeResidualNorm();
    return rc;
}
int cfref_apply_pressure( cfref_ctx* c )
{
    return guarded( c, [&]() { c->h->applyPressure(); } );
}
int cfref_correct_velocity( cfref_ctx* c, int* num_iter, double* resid )
{
    int rc = guarded( c, [&]() { c->h->correctVelocity(); } );
    account_cg( c );
    if ( num_iter )
        *num_iter = c->last_iters;
    if ( resid )
        *resid = c->h->cg()->getFinalRelativeResidualNorm();
    return rc;
}
int cfref_setup( cfref_ctx* c )
{
    int rc = guarded( c, [&]() { c->h->base->setup(); } );
    account_cg( c );
    return rc;
}
int cfref_step( cfref_ctx* c )
{
    int rc = guarded( c, [&]() { c->h->base->step(); } );
    account_cg( c );
    c->steps++;
    return rc;
}
// SolverBase::solve, Silo writes included (captured in memory by the silo.h stand-in).  The step
// count is recovered from the clock: solve() advances _time by _dt per step (src/Solver.hpp:146).
int cfref_solve( cfref_ctx* c, double t_final, int write_freq, int* steps_taken )
{
    const double t0 = c->h->time();
    int rc = guarded( c, [&]() { c->h->base->solve( t_final, write_freq > 0 ? write_freq : 1 << 30 ); } );
    int n = 0;
    double t = t0;
    while ( t < c->h->time() && n < ( 1 << 30 ) )
    {
        t += c->h->dt();
        ++n;
    }
    c->steps += n;
    if ( steps_taken )
        *steps_taken = n;
    return rc;
}

int cfref_stencil_dot( cfref_ctx* c, int, double*, double* )
{
    return fail( c, CFB_ERR_INVALID, "stencil_dot is a micro-benchmark entry of the product, not of the reference" );
}

int cfref_get_stats( const cfref_ctx* c, cfb_stats* out )
{
    std::memset( out, 0, sizeof( *out ) );
    out->cg_iterations = c->cg_total;
    out->steps = c->steps;
    return CFB_OK;
}
int cfref_reset_stats( cfref_ctx* c )
{
    c->cg_total = 0;
    c->steps = 0;
    return CFB_OK;
}

int cfref_residual_history( const cfref_ctx* c, double* hist, int n, int* count )
{
    const auto& h = c->h->cg()->history();
    const int k = std::min<int>( n, (int)h.size() );
    for ( int i = 0; i < k; ++i )
        hist[i] = h[i];
    if ( count )
        *count = (int)h.size();
    return CFB_OK;
}

// The reference's stored matrix: (ghosted i, ghosted j, 5 coefficients) and inverse diagonal, copied
// for the owned cells into dense x-fastest arrays A[j][i][5], Minv[j][i].
int cfref_matrix( cfref_ctx* c, double* A, double* Minv )
{
    int lo[2], hi[2];
    c->h->ownedSpace( CFB_QUANTITY, lo, hi );
    auto a = c->h->cg()->getMatrixValues().view();
    auto m = c->h->cg()->getPreconditionerValues().view();
    const int ex = hi[0] - lo[0];
    for ( int j = lo[1]; j < hi[1]; ++j )
        for ( int i = lo[0]; i < hi[0]; ++i )
        {
            const size_t o = (size_t)( j - lo[1] ) * ex + ( i - lo[0] );
            for ( int s = 0; s < 5; ++s )
                A[o * 5 + s] = a( i, j, s );
            Minv[o] = m( i, j, 0 );
        }
    return CFB_OK;
}

// What the reference's SiloWriter handed to Silo on its most recent write (src/SiloWriter.hpp:56-197):
// owned quantity and cell-centred velocity (order-1 interpolation of the face velocities), x fastest.
int cfref_silo_last( int* writes, int* cycle, double* time, int dims[2], double* quantity, double* ucc, double* vcc,
                     double* xnodes, double* ynodes )
{
    const auto& s = cfref::silo_capture();
    if ( writes )
        *writes = s.writes;
    if ( cycle )
        *cycle = s.cycle;
    if ( time )
        *time = s.time;
    if ( dims )
    {
        dims[0] = s.zone_dims[0];
        dims[1] = s.zone_dims[1];
    }
    auto put = []( double* dst, const std::vector<double>& src ) {
        if ( dst && !src.empty() )
            std::memcpy( dst, src.data(), src.size() * sizeof( double ) );
    };
    put( quantity, s.quantity );
    put( ucc, s.velocity[0] );
    put( vcc, s.velocity[1] );
    put( xnodes, s.coords[0] );
    put( ynodes, s.coords[1] );
    return CFB_OK;
}

// One SiloWriter::siloWrite of the current state (the call of src/Solver.hpp:170-173), handed back in the
// shape of cfb_output_extract: quantity[ny][nx], velocity[2][ny][nx], node coordinates per dim.
int cfref_output_extract( cfref_ctx* c, double* quantity, double* velocity, double* nodes_x, double* nodes_y,
                          double* /*nodes_z*/ )
{
    int rc = guarded( c, [&]() { c->h->siloWrite( (int)c->steps ); } );
    if ( rc )
        return rc;
    const auto& s = cfref::silo_capture();
    const size_t n = (size_t)s.zone_dims[0] * s.zone_dims[1];
    if ( s.quantity.size() != n || s.velocity[0].size() != n || s.velocity[1].size() != n )
        return fail( c, CFB_ERR_INVALID, "silo capture is incomplete" );
    if ( quantity )
        std::memcpy( quantity, s.quantity.data(), n * sizeof( double ) );
    if ( velocity )
    {
        std::memcpy( velocity, s.velocity[0].data(), n * sizeof( double ) );
        std::memcpy( velocity + n, s.velocity[1].data(), n * sizeof( double ) );
    }
    if ( nodes_x )
        std::memcpy( nodes_x, s.coords[0].data(), s.coords[0].size() * sizeof( double ) );
    if ( nodes_y )
        std::memcpy( nodes_y, s.coords[1].data(), s.coords[1].size() * sizeof( double ) );
    return CFB_OK;
}

int cfref_num_threads( void )
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
