// CajitaFluids.hpp — header-only C++ mirror of the reference's host interfaces on top of the C ABI
// (include/cfb.h).  Same namespace, class and method names as cajitafluids, minus the Kokkos /
// Cajita / MPI template parameters, so a driver written against the reference (e.g.
// examples/advection.cpp) switches to the B200 path by changing its includes and passing the
// device string "b200".
//
//   reference                                              here
//   ------------------------------------------------------ ------------------------------------------
//   Cajita::Cell, Cajita::Face<Cajita::Dim::I|J|K>          CajitaFluids::Cell, FaceI, FaceJ, FaceK
//   Field::Quantity/Velocity, Version::Current/Next          same tags   (src/ProblemManager.hpp:35-85)
//   Cajita::DimBlockPartitioner<D>                           CajitaFluids::DimBlockPartitioner<D>
//   MPI_Comm                                                 CajitaFluids::Comm {rank, size, nccl ids}
//   BoundaryCondition<D>, InflowSource<D>, BodyForce<D>      same PODs   (src/BoundaryConditions.hpp,
//                                                            InflowSource.hpp, BodyForce.hpp)
//   Mesh<D,Exec,Mem>                                         Mesh<D>     (src/Mesh.hpp:41-136)
//   ProblemManager<D,Exec,Mem>                               ProblemManager<D>  get/advance/gather
//   Cajita::ReferenceConjugateGradient                       B200ConjugateGradient (setTolerance,
//                                                            setMaxIter, setPrintLevel, solve, ...)
//   VelocityCorrectorBase / createVelocityCorrector          same        (src/VelocityCorrector.hpp)
//   TimeIntegrator::step                                     same        (src/TimeIntegrator.hpp:120)
//   SiloWriter<D,Exec,Mem>::siloWrite                        SiloWriter<D>::siloWrite (src/SiloWriter.hpp:355;
//                                                            .npy + .json instead of Silo/PMPIO, asynchronous)
//   SolverBase / Solver<D,...> / createSolver                same        (src/Solver.hpp:41-48,283-350)
//
// Errors: the reference throws std::runtime_error / std::logic_error; the shims translate the C
// status codes back into the same exception types.
#ifndef CAJITAFLUIDS_B200_HPP
#define CAJITAFLUIDS_B200_HPP

#include "../cfb.h"

#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace CajitaFluids
{

// ---- entity / field / version tags ---------------------------------------------------------------
struct Cell
{
    static constexpr int id = CFB_QUANTITY;
};
struct FaceI
{
    static constexpr int id = CFB_U;
};
struct FaceJ
{
    static constexpr int id = CFB_V;
};
struct FaceK
{
    static constexpr int id = CFB_W;
};
namespace Field
{
struct Quantity
{
};
struct Velocity
{
};
} // namespace Field
namespace Version
{
struct Current
{
    static constexpr int id = CFB_CURRENT;
};
struct Next
{
    static constexpr int id = CFB_NEXT;
};
} // namespace Version

// src/BoundaryConditions.hpp:30-37
struct BoundaryType
{
    enum Values
    {
        SOLID = CFB_SOLID,
        FREE = CFB_FREE,
    };
};

// src/BoundaryConditions.hpp:131-134: plain data; min/max are filled by the Solver.
template <std::size_t NumSpaceDim>
struct BoundaryCondition
{
    std::array<int, 2 * NumSpaceDim> boundary_type{}; // [d] low wall, [NumSpaceDim + d] high wall
    std::array<int, NumSpaceDim> min{};
    std::array<int, NumSpaceDim> max{};
};

// src/InflowSource.hpp:80-101
template <std::size_t NumSpaceDim>
struct InflowSource
{
    InflowSource( std::array<double, NumSpaceDim> location, std::array<double, NumSpaceDim> size,
                  std::array<double, NumSpaceDim> velocity, double quantity )
        : _quantity( quantity )
    {
        for ( std::size_t d = 0; d < NumSpaceDim; ++d )
        {
            _bounding_box[d] = location[d];
            _bounding_box[NumSpaceDim + d] = location[d] + size[d];
            _velocity[d] = velocity[d];
        }
    }
    std::array<double, 2 * NumSpaceDim> _bounding_box;
    double _quantity;
    std::array<double, NumSpaceDim> _velocity;
};

// src/BodyForce.hpp:62-69
template <std::size_t NumSpaceDim>
struct BodyForce
{
    BodyForce( double fx, double fy, double fz = 0.0 )
    {
        const double f[3] = { fx, fy, fz };
        for ( std::size_t d = 0; d < NumSpaceDim; ++d )
            _force[d] = f[d];
    }
    std::array<double, NumSpaceDim> _force;
};

// Stand-in for MPI_Comm: which rank of how many this process is, plus the two NCCL ids that rank 0
// created with cfb_nccl_unique_id and broadcast out of band.
struct Comm
{
    int rank = 0;
    int size = 1;
    std::array<unsigned char, 2 * CFB_NCCL_ID_BYTES> nccl_id{};
};

// Cajita::DimBlockPartitioner: near-cubic block grid; the split order is z, y, x so that the largest
// faces are the contiguous ones.
template <std::size_t NumSpaceDim>
struct DimBlockPartitioner
{
    std::array<int, NumSpaceDim> ranksPerDimension( int comm_size ) const
    {
        std::array<int, NumSpaceDim> r;
        r.fill( 1 );
        int d = NumSpaceDim - 1;
        while ( comm_size > 1 )
        {
            if ( comm_size % 2 )
                throw std::runtime_error( "DimBlockPartitioner: comm size must be a power of two" );
            r[d] *= 2;
            comm_size /= 2;
            d = ( d + NumSpaceDim - 1 ) % NumSpaceDim;
        }
        return r;
    }
};

namespace detail
{
inline void check( int rc, const cfb_ctx* ctx )
{
    if ( rc == CFB_OK )
        return;
    std::string msg = cfb_last_error( ctx );
    if ( rc == CFB_ERR_MESH_EXTENT )
        throw std::logic_error( msg ); // src/Mesh.hpp:62-63
    throw std::runtime_error( msg );
}

struct CtxHolder
{
    cfb_ctx* ctx = nullptr;
    cfb_config cfg{};
    ~CtxHolder()
    {
        if ( ctx )
            cfb_destroy( ctx );
    }
};
} // namespace detail

// A non-owning view of one field: device pointer into the padded array + indexing data.
// view( i, j, k ) indexing uses the reference's LOCAL (ghosted) indices: owned entities start at
// halo width, exactly like the Kokkos views returned by ProblemManager::get.
struct FieldView
{
    double* dev_ptr = nullptr;
    int64_t origin = 0, stride_y = 0, stride_z = 0;
    int halo = 0;
    int extent[3] = { 0, 0, 0 }; // owned extent
    int64_t offset( int i, int j, int k = 0 ) const
    {
        return origin + ( k - ( stride_z ? halo : 0 ) ) * stride_z + ( j - halo ) * stride_y + ( i - halo );
    }
};

// ---- Mesh (src/Mesh.hpp) ---------------------------------------------------------------------------
template <std::size_t NumSpaceDim>
class Mesh
{
  public:
    explicit Mesh( std::shared_ptr<detail::CtxHolder> h )
        : _h( std::move( h ) )
    {
    }
    double cellSize() const
    {
        double c;
        cfb_get_scalars( _h->ctx, &c, nullptr, nullptr );
        return c;
    }
    std::array<int, NumSpaceDim> minDomainGlobalCellIndex() const
    {
        std::array<int, NumSpaceDim> m;
        m.fill( 0 );
        return m;
    }
    std::array<int, NumSpaceDim> maxDomainGlobalCellIndex() const
    {
        std::array<int, NumSpaceDim> m;
        for ( std::size_t d = 0; d < NumSpaceDim; ++d )
            m[d] = _h->cfg.global_num_cell[d] - 1;
        return m;
    }
    int rank() const { return _h->cfg.world_rank; }
    int haloCellWidth() const { return _h->cfg.halo_cell_width; }
    // owned extent of an entity (Cajita::Own index space) and the block's global cell offset (L2G)
    template <class Entity>
    std::array<int, 3> ownedExtent( Entity ) const
    {
        std::array<int, 3> e;
        cfb_owned_extent( _h->ctx, Entity::id, e.data() );
        return e;
    }
    std::array<int, 3> globalOffset() const
    {
        std::array<int, 3> o;
        cfb_global_offset( _h->ctx, o.data() );
        return o;
    }

  private:
    std::shared_ptr<detail::CtxHolder> _h;
};

// ---- ProblemManager (src/ProblemManager.hpp) -------------------------------------------------------
template <std::size_t NumSpaceDim>
class ProblemManager
{
  public:
    using mesh_type = Mesh<NumSpaceDim>;
    explicit ProblemManager( std::shared_ptr<detail::CtxHolder> h )
        : _h( std::move( h ) )
        , _mesh( std::make_shared<mesh_type>( _h ) )
    {
    }
    const std::shared_ptr<mesh_type>& mesh() const { return _mesh; }

    template <class Entity, class FieldTag, class VersionTag>
    FieldView get( Entity, FieldTag, VersionTag ) const
    {
        FieldView v;
        detail::check( cfb_field_ptr( _h->ctx, Entity::id, VersionTag::id, &v.dev_ptr, &v.origin, &v.stride_y,
                                      &v.stride_z ),
                       _h->ctx );
        v.halo = _h->cfg.halo_cell_width;
        cfb_owned_extent( _h->ctx, Entity::id, v.extent );
        return v;
    }
    template <class Entity, class FieldTag>
    void advance( Entity, FieldTag )
    {
        detail::check( cfb_advance( _h->ctx, Entity::id ), _h->ctx );
    }
    void gather( Version::Current ) const { detail::check( cfb_gather( _h->ctx, CFB_CURRENT ), _h->ctx ); }
    void gather( Version::Next ) const { detail::check( cfb_gather( _h->ctx, CFB_NEXT ), _h->ctx ); }

    // host mirrors for tests / output (Kokkos::create_mirror_view_and_copy in the reference's tests)
    template <class Entity, class VersionTag>
    std::vector<double> copyToHost( Entity, VersionTag, bool ghosted = false ) const
    {
        int e[3];
        cfb_owned_extent( _h->ctx, ghosted ? CFB_QUANTITY : Entity::id, e );
        size_t n = 1;
        for ( std::size_t d = 0; d < 3; ++d )
        {
            int ext = e[d];
            if ( ghosted && d < NumSpaceDim )
                ext += 2 * _h->cfg.halo_cell_width + ( Entity::id - 1 == (int)d ? 1 : 0 );
            n *= (size_t)ext;
        }
        std::vector<double> out( n );
        detail::check( cfb_download( _h->ctx, Entity::id, VersionTag::id, ghosted ? CFB_GHOSTED : CFB_OWNED,
                                     out.data() ),
                       _h->ctx );
        return out;
    }
    template <class Entity, class VersionTag>
    void copyFromHost( Entity, VersionTag, const std::vector<double>& in, bool ghosted = false )
    {
        detail::check( cfb_upload( _h->ctx, Entity::id, VersionTag::id, ghosted ? CFB_GHOSTED : CFB_OWNED,
                                   in.data() ),
                       _h->ctx );
    }

    // ProblemManager::initialize (src/ProblemManager.hpp:186-263): the functor is evaluated on the
    // host over the owned entities (a device lambda cannot cross the C ABI) and uploaded.
    template <class InitFunctor>
    void initialize( const InitFunctor& f )
    {
        init_entity( Cell(), Field::Quantity(), f );
        init_entity( FaceI(), Field::Velocity(), f );
        init_entity( FaceJ(), Field::Velocity(), f );
        if constexpr ( NumSpaceDim == 3 )
            init_entity( FaceK(), Field::Velocity(), f );
    }

    const std::shared_ptr<detail::CtxHolder>& holder() const { return _h; }

  private:
    template <class Entity, class FieldTag, class InitFunctor>
    void init_entity( Entity ent, FieldTag tag, const InitFunctor& f )
    {
        int e[3], off[3];
        cfb_owned_extent( _h->ctx, Entity::id, e );
        cfb_global_offset( _h->ctx, off );
        const int halo = _h->cfg.halo_cell_width;
        double cell;
        cfb_get_scalars( _h->ctx, &cell, nullptr, nullptr );
        std::vector<double> host( (size_t)e[0] * e[1] * e[2] );
        size_t n = 0;
        for ( int k = 0; k < e[2]; ++k )
            for ( int j = 0; j < e[1]; ++j )
                for ( int i = 0; i < e[0]; ++i, ++n )
                {
                    const int own[3] = { i, j, k };
                    int coords[NumSpaceDim];
                    double x[NumSpaceDim];
                    for ( std::size_t d = 0; d < NumSpaceDim; ++d )
                    {
                        coords[d] = own[d] + halo; // local index, as the reference passes it
                        const double g = own[d] + off[d] + ( Entity::id - 1 == (int)d ? 0.0 : 0.5 );
                        x[d] = _h->cfg.global_bounding_box[d] + g * cell;
                    }
                    double v = 0.0;
                    f( ent, tag, coords, x, v );
                    host[n] = v;
                }
        copyFromHost( ent, Version::Current(), host );
    }
    std::shared_ptr<detail::CtxHolder> _h;
    std::shared_ptr<mesh_type> _mesh;
};

// ---- the solver plug-in surface (Cajita::ReferenceConjugateGradient) -------------------------------
class B200ConjugateGradient
{
  public:
    explicit B200ConjugateGradient( std::shared_ptr<detail::CtxHolder> h )
        : _h( std::move( h ) )
    {
    }
    // The matrix / preconditioner are implied by (dt, density, cell size, boundary types): the
    // stencil setters only validate that the caller asks for the operator the kernels implement.
    template <class Stencil>
    void setMatrixStencil( const Stencil& s, bool is_symmetric = false )
    {
        (void)is_symmetric;
        if ( s.size() != 2 * (size_t)_h->cfg.dim + 1 )
            throw std::runtime_error( "B200ConjugateGradient: only the 2*D+1 point Laplacian stencil is supported" );
    }
    template <class Stencil>
    void setPreconditionerStencil( const Stencil& s, bool is_symmetric = false )
    {
        (void)is_symmetric;
        if ( s.size() != 1 )
            throw std::runtime_error( "B200ConjugateGradient: only the diagonal (Jacobi) preconditioner is supported" );
    }
    void setTolerance( double tol )
    {
        _tol = tol;
        push();
    }
    void setMaxIter( int n )
    {
        _max_iter = n;
        push();
    }
    void setPrintLevel( int l )
    {
        _print = l;
        push();
    }
    void setup() { push(); }
    // solve( b, x ) with both vectors resident on the device as the ctx's RHS / PRESSURE fields
    void solve()
    {
        detail::check( cfb_pcg_solve( _h->ctx, &_num_iter, &_resid ), _h->ctx );
    }
    // solve( b, x ) with host vectors over the owned cells
    void solve( const std::vector<double>& b, std::vector<double>& x )
    {
        x.resize( b.size() );
        detail::check( cfb_pcg_solve_host( _h->ctx, b.data(), x.data(), &_num_iter, &_resid ), _h->ctx );
    }
    int getNumIter() const { return _num_iter; }
    double getFinalResidualNorm() const { return _resid; }
    double tolerance() const { return _tol; }
    int maxIter() const { return _max_iter; }
    int printLevel() const { return _print; }

  private:
    void push() { detail::check( cfb_set_cg_params( _h->ctx, _tol, _max_iter, _print ), _h->ctx ); }
    std::shared_ptr<detail::CtxHolder> _h;
    double _tol = 1.0e-6, _resid = 0.0;
    int _max_iter = 2000, _print = 1, _num_iter = 0;
};

// ---- VelocityCorrector (src/VelocityCorrector.hpp) -------------------------------------------------
class VelocityCorrectorBase
{
  public:
    virtual ~VelocityCorrectorBase() = default;
    virtual void correctVelocity() = 0;
};

template <std::size_t NumSpaceDim>
class VelocityCorrector : public VelocityCorrectorBase
{
  public:
    VelocityCorrector( std::shared_ptr<detail::CtxHolder> h, std::shared_ptr<B200ConjugateGradient> solver )
        : _h( std::move( h ) )
        , _pressure_solver( std::move( solver ) )
    {
        // src/VelocityCorrector.hpp:96-106: stencil, Jacobi preconditioner, tol / max_iter / print
        const std::size_t D = NumSpaceDim;
        std::vector<std::array<int, NumSpaceDim>> stencil( 2 * D + 1 );
        for ( auto& s : stencil )
            s.fill( 0 );
        for ( std::size_t d = 0; d < D; ++d )
        {
            stencil[1 + 2 * d][d] = -1;
            stencil[2 + 2 * d][d] = 1;
        }
        _pressure_solver->setMatrixStencil( stencil, false );
        std::vector<std::array<int, NumSpaceDim>> diag_stencil( 1 );
        diag_stencil[0].fill( 0 );
        _pressure_solver->setPreconditionerStencil( diag_stencil, false );
        _pressure_solver->setTolerance( 1.0e-6 );
        _pressure_solver->setMaxIter( 2000 );
        _pressure_solver->setPrintLevel( 1 );
        _pressure_solver->setup();
    }
    void _buildRHS() { detail::check( cfb_build_rhs( _h->ctx ), _h->ctx ); }
    void _applyPressure() { detail::check( cfb_apply_pressure( _h->ctx ), _h->ctx ); }
    void correctVelocity() override
    {
        _buildRHS();
        _pressure_solver->solve();
        _applyPressure();
    }
    const std::shared_ptr<B200ConjugateGradient>& pressureSolver() const { return _pressure_solver; }

  private:
    std::shared_ptr<detail::CtxHolder> _h;
    std::shared_ptr<B200ConjugateGradient> _pressure_solver;
};

// ---- TimeIntegrator (src/TimeIntegrator.hpp:120-177) ------------------------------------------------
namespace TimeIntegrator
{
template <std::size_t NumSpaceDim>
void step( std::shared_ptr<detail::CtxHolder>& h )
{
    detail::check( cfb_time_integrator_step( h->ctx ), h->ctx );
}
} // namespace TimeIntegrator

// ---- SiloWriter (src/SiloWriter.hpp) -------------------------------------------------------------
// Same constructor argument and siloWrite signature as the reference's writer.  The extraction (owned q,
// cell-centred velocity) is one kernel and the copy to the host is asynchronous: siloWrite returns at
// once, the files of a write appear at the next siloWrite / flush / destruction of the solver.  Silo and
// PMPIO are not available, so the container is .npy per variable and block + a .json master per step
// under `directory()` ("data" like src/SiloWriter.hpp:379-384).
template <std::size_t NumSpaceDim>
class SiloWriter
{
  public:
    using pm_type = ProblemManager<NumSpaceDim>;
    explicit SiloWriter( const std::shared_ptr<pm_type>& pm )
        : _h( pm->holder() )
    {
    }
    explicit SiloWriter( const std::shared_ptr<detail::CtxHolder>& h )
        : _h( h )
    {
    }
    void setDirectory( const std::string& dir ) { _dir = dir; }
    const std::string& directory() const { return _dir; }
    // name ("Mesh"), time and dt are what the reference passes; time and dt are read from the solver state
    void siloWrite( const char* /*name*/, int time_step, double /*time*/, double /*dt*/ )
    {
        detail::check( cfb_write_output( _h->ctx, _dir.c_str(), time_step ), _h->ctx );
    }
    void flush() { detail::check( cfb_output_flush( _h->ctx ), _h->ctx ); }
    // what writeFile hands to Silo, as host arrays (x fastest): quantity[ncell], velocity[D * ncell]
    void extract( std::vector<double>& quantity, std::vector<double>& velocity )
    {
        int ext[3];
        cfb_owned_extent( _h->ctx, CFB_QUANTITY, ext );
        const size_t n = (size_t)ext[0] * ext[1] * ext[2];
        quantity.resize( n );
        velocity.resize( n * NumSpaceDim );
        detail::check( cfb_output_extract( _h->ctx, quantity.data(), velocity.data(), nullptr, nullptr, nullptr ),
                       _h->ctx );
    }

  private:
    std::shared_ptr<detail::CtxHolder> _h;
    std::string _dir = "data";
};

// ---- Solver (src/Solver.hpp) -----------------------------------------------------------------------
class SolverBase
{
  public:
    virtual ~SolverBase() = default;
    virtual void setup( void ) = 0;
    virtual void step( void ) = 0;
    virtual void solve( const double t_final, const int write_freq ) = 0;
};

template <std::size_t NumSpaceDim>
class Solver : public SolverBase
{
  public:
    using pm_type = ProblemManager<NumSpaceDim>;
    using bc_type = BoundaryCondition<NumSpaceDim>;

    template <class InitFunc>
    Solver( const Comm& comm, const std::array<double, 2 * NumSpaceDim>& global_bounding_box,
            const std::array<int, NumSpaceDim>& global_num_cell,
            const DimBlockPartitioner<NumSpaceDim>& partitioner, const double density,
            const InitFunc& create_functor, const BoundaryCondition<NumSpaceDim>& bc,
            const InflowSource<NumSpaceDim>& source, const BodyForce<NumSpaceDim>& body, const double delta_t,
            const std::string& matrix_solver, const std::string& preconditioner, int device_id = -1 )
        : _bc( bc )
    {
        // the Reference path always builds its Jacobi preconditioner itself (the reference ignores the
        // preconditioner string there); "MG" is this backend's opt-in multigrid V-cycle (cfb_set_preconditioner)
        if ( matrix_solver != "Reference" )
            throw std::runtime_error( "cajitafluids_b200 implements only the 'Reference' matrix solver "
                                      "(HYPRE is out of scope)" );
        _h = std::make_shared<detail::CtxHolder>();
        cfb_config& c = _h->cfg;
        cfb_default_config( &c, (int)NumSpaceDim );
        auto ranks = partitioner.ranksPerDimension( comm.size );
        int r = comm.rank;
        for ( std::size_t d = 0; d < NumSpaceDim; ++d )
        {
            c.global_num_cell[d] = global_num_cell[d];
            c.global_bounding_box[d] = global_bounding_box[d];
            c.global_bounding_box[3 + d] = global_bounding_box[NumSpaceDim + d];
            c.ranks_per_dim[d] = ranks[d];
            c.block_id[d] = r % ranks[d]; // rank = (bz*py + by)*px + bx
            r /= ranks[d];
            c.boundary_type[d] = bc.boundary_type[d];
            c.boundary_type[NumSpaceDim + d] = bc.boundary_type[NumSpaceDim + d];
            c.inflow_location[d] = source._bounding_box[d];
            c.inflow_size[d] = source._bounding_box[NumSpaceDim + d] - source._bounding_box[d];
            c.inflow_velocity[d] = source._velocity[d];
            c.body_force[d] = body._force[d];
        }
        for ( std::size_t d = NumSpaceDim; d < 3; ++d )
        {
            c.inflow_location[d] = c.inflow_size[d] = c.inflow_velocity[d] = c.body_force[d] = 0.0;
        }
        c.inflow_quantity = source._quantity;
        c.world_rank = comm.rank;
        c.world_size = comm.size;
        c.density = density;
        c.delta_t = delta_t; // clamped inside cfb_create like src/Solver.hpp:96-106
        c.cg_print_level = 1; // src/VelocityCorrector.hpp:105
        c.device_id = device_id >= 0 ? device_id : comm.rank;
        c.use_nccl = comm.size > 1;
        std::memcpy( c.nccl_id, comm.nccl_id.data(), sizeof( c.nccl_id ) );
        detail::check( cfb_create( &c, &_h->ctx ), _h->ctx );

        _bc.min = Mesh<NumSpaceDim>( _h ).minDomainGlobalCellIndex(); // src/Solver.hpp:109-110
        _bc.max = Mesh<NumSpaceDim>( _h ).maxDomainGlobalCellIndex();
        _pm = std::make_shared<pm_type>( _h );
        _pm->initialize( create_functor );
        if ( preconditioner == "MG" )
            detail::check( cfb_set_preconditioner( _h->ctx, CFB_PRECOND_MG, 2, 2, 8, 0.0 ), _h->ctx );
        auto cg = std::make_shared<B200ConjugateGradient>( _h );
        _vc = std::make_shared<VelocityCorrector<NumSpaceDim>>( _h, cg );
        _silo = std::make_shared<SiloWriter<NumSpaceDim>>( _h ); // src/Solver.hpp:121-122
    }

    void setup() override { detail::check( cfb_setup( _h->ctx ), _h->ctx ); }
    void step() override { detail::check( cfb_step( _h->ctx ), _h->ctx ); }
    // src/Solver.hpp:149-177, _silo->siloWrite before setup and after every write_freq-th step included
    // (into siloWriter()->directory(), "data" by default; an empty directory name turns the writes off)
    void solve( const double t_final, const int write_freq ) override
    {
        int steps = 0;
        detail::check( cfb_set_output_dir( _h->ctx, _silo->directory().c_str() ), _h->ctx );
        detail::check( cfb_solve( _h->ctx, t_final, write_freq, &steps ), _h->ctx );
        _steps = steps;
    }
    const std::shared_ptr<SiloWriter<NumSpaceDim>>& siloWriter() const { return _silo; }
    void _addInputs() { detail::check( cfb_add_inputs( _h->ctx ), _h->ctx ); }

    const std::shared_ptr<pm_type>& problemManager() const { return _pm; }
    const std::shared_ptr<VelocityCorrectorBase>& velocityCorrector() const { return _vc; }
    double time() const
    {
        double t;
        cfb_get_scalars( _h->ctx, nullptr, nullptr, &t );
        return t;
    }
    double deltaT() const
    {
        double t;
        cfb_get_scalars( _h->ctx, nullptr, &t, nullptr );
        return t;
    }
    int stepsTaken() const { return _steps; }
    cfb_stats stats() const
    {
        cfb_stats s;
        cfb_get_stats( _h->ctx, &s );
        return s;
    }

  private:
    std::shared_ptr<detail::CtxHolder> _h;
    bc_type _bc;
    std::shared_ptr<pm_type> _pm;
    std::shared_ptr<VelocityCorrectorBase> _vc;
    std::shared_ptr<SiloWriter<NumSpaceDim>> _silo;
    int _steps = 0;
};

// src/VelocityCorrector.hpp:297-339 — string dispatch; HYPRE names are rejected.
template <std::size_t NumSpaceDims>
std::shared_ptr<VelocityCorrectorBase>
createVelocityCorrector( const std::shared_ptr<Solver<NumSpaceDims>>& solver, std::string matrix_solver,
                         std::string precon )
{
    (void)precon;
    if ( matrix_solver.compare( "Reference" ) != 0 )
        throw std::runtime_error( "only the 'Reference' solver is available on the b200 backend" );
    return solver->velocityCorrector();
}

// src/Solver.hpp:283-350 — device-string dispatch.  One backend: "b200" ("cuda" accepted as alias).
template <std::size_t NumSpaceDim, class InitFunc>
std::shared_ptr<SolverBase>
createSolver( const std::string& device, const Comm& comm,
              const std::array<double, 2 * NumSpaceDim>& global_bounding_box,
              const std::array<int, NumSpaceDim>& global_num_cell,
              const DimBlockPartitioner<NumSpaceDim>& partitioner, const double density,
              const InitFunc& create_functor, const BoundaryCondition<NumSpaceDim>& bc,
              const InflowSource<NumSpaceDim>& source, const BodyForce<NumSpaceDim>& body, const double delta_t,
              const std::string& matrix_solver, const std::string& preconditioner )
{
    if ( 0 == device.compare( "b200" ) || 0 == device.compare( "cuda" ) )
    {
        return std::make_shared<Solver<NumSpaceDim>>( comm, global_bounding_box, global_num_cell, partitioner,
                                                      density, create_functor, bc, source, body, delta_t,
                                                      matrix_solver, preconditioner );
    }
    else if ( 0 == device.compare( "serial" ) )
        throw std::runtime_error( "Serial Backend Not Enabled" );
    else if ( 0 == device.compare( "openmp" ) )
        throw std::runtime_error( "OpenMP Backend Not Enabled" );
    else if ( 0 == device.compare( "hip" ) )
        throw std::runtime_error( "HIP Backend Not Enabled" );
    throw std::runtime_error( "invalid backend" );
}

} // namespace CajitaFluids

#endif // CAJITAFLUIDS_B200_HPP
