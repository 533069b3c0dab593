// Cajita.hpp — STAND-IN for the absent Cabana/Cajita dependency.  TEST INFRASTRUCTURE ONLY.
//
// Single-rank, host-only, non-periodic subset of the Cajita (Cabana ~0.5) API that the cajitafluids
// sources use, written from scratch so that the UNMODIFIED reference (src/*.hpp,
// examples/advection.cpp, tests/tst*.cpp — compiled from /root/reference, never copied) runs here.
// The reference's own statements (divergence, matrix fill, boundary conditions, gradient
// subtraction, RK3, interpolation call sites, inflow, body force, dt clamp, step orchestration,
// Silo extraction) therefore execute as written; what this file supplies is the third-party part,
// restated from Cajita's published behaviour and marked [Cajita-mem] exactly like in
// oracle/cfo_oracle.cpp:
//   * grid bookkeeping: owned / ghosted index spaces with halo cells allocated also on physical
//     walls (pinned by the reference's tests/tstMesh.cpp:47-68, which this shim passes),
//     LocalMesh::coordinates, IndexConversion::createL2G;
//   * Array / ArrayLayout / ArrayOp::assign, Halo (one rank, non-periodic: nothing to exchange);
//   * B-spline data (orders 1 and 3), evaluateSpline, G2P::value;
//   * ReferenceStructuredSolver / ReferenceConjugateGradient (preconditioned CG, absolute 2-norm
//     stopping test), HypreStructuredSolver as a type that throws when used.
#ifndef CFREF_SHIM_CAJITA_HPP
#define CFREF_SHIM_CAJITA_HPP

#include <Kokkos_Core.hpp>
#include <mpi.h>

#include <array>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
// Knobs for the test driver (oracle/refshim/cfref_driver.cpp).
namespace cfref
{
struct Knobs
{
    // CG arithmetic: 0 = plain double sums, products and updates as the published loop writes them;
    // 1 = the oracle's arithmetic (fused multiply-adds where `a += b * c` appears, double-double
    //     accumulation of the dot products), which makes whole runs bit-comparable with it.
    int cg_exact = 0;
    int cg_print = -1; // >= 0 overrides the level set through setPrintLevel
};
inline Knobs& knobs()
{
    static Knobs k;
    return k;
}
} // namespace cfref

namespace Cajita
{

// ---- tags -------------------------------------------------------------------------------------
struct Dim
{
    enum Values
    {
        I = 0,
        J = 1,
        K = 2
    };
};
struct Cell
{
};
struct Node
{
};
template <int D>
struct Face;
template <>
struct Face<Dim::I>
{
    static constexpr int dim = Dim::I;
};
template <>
struct Face<Dim::J>
{
    static constexpr int dim = Dim::J;
};
template <>
struct Face<Dim::K>
{
    static constexpr int dim = Dim::K;
};
struct Own
{
};
struct Ghost
{
};
struct Local
{
};
struct Global
{
};

namespace Impl
{
// number of entities in dimension d = number of cells + extra(d)
template <class Entity>
struct EntityExtra;
template <>
struct EntityExtra<Cell>
{
    static constexpr int extra( int ) { return 0; }
};
template <>
struct EntityExtra<Node>
{
    static constexpr int extra( int ) { return 1; }
};
template <int D>
struct EntityExtra<Face<D>>
{
    static constexpr int extra( int d ) { return d == D ? 1 : 0; }
};
} // namespace Impl

// ---- meshes -----------------------------------------------------------------------------------
template <class Scalar, std::size_t NumSpaceDim = 3>
struct UniformMesh
{
    using scalar_type = Scalar;
    static constexpr std::size_t num_space_dim = NumSpaceDim;
};

template <std::size_t N>
class BlockPartitioner
{
  public:
    virtual ~BlockPartitioner() = default;
    virtual std::array<int, N> ranksPerDimension( MPI_Comm comm,
                                                  const std::array<int, N>& global_cells_per_dim ) const = 0;
};
template <std::size_t N>
class DimBlockPartitioner : public BlockPartitioner<N>
{
  public:
    DimBlockPartitioner() = default;
    std::array<int, N> ranksPerDimension( MPI_Comm, const std::array<int, N>& ) const override
    {
        std::array<int, N> r;
        r.fill( 1 ); // MPI_Dims_create on one rank
        return r;
    }
};
template <std::size_t N>
class ManualBlockPartitioner : public BlockPartitioner<N>
{
  public:
    explicit ManualBlockPartitioner( const std::array<int, N>& r )
        : _r( r )
    {
    }
    std::array<int, N> ranksPerDimension( MPI_Comm, const std::array<int, N>& ) const override { return _r; }

  private:
    std::array<int, N> _r;
};

template <class MeshType>
class GlobalMesh
{
  public:
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    GlobalMesh( const std::array<double, num_space_dim>& lo, const std::array<double, num_space_dim>& hi,
                const std::array<int, num_space_dim>& n )
        : _lo( lo )
        , _hi( hi )
        , _n( n )
    {
        for ( std::size_t d = 0; d < num_space_dim; ++d )
            _cell[d] = ( _hi[d] - _lo[d] ) / _n[d];
    }
    double lowCorner( int d ) const { return _lo[d]; }
    double highCorner( int d ) const { return _hi[d]; }
    double extent( int d ) const { return _hi[d] - _lo[d]; }
    int globalNumCell( int d ) const { return _n[d]; }
    double cellSize( int d ) const { return _cell[d]; }

  private:
    std::array<double, num_space_dim> _lo, _hi;
    std::array<int, num_space_dim> _n;
    std::array<double, num_space_dim> _cell;
};

template <class Scalar, std::size_t N>
std::shared_ptr<GlobalMesh<UniformMesh<Scalar, N>>>
createUniformGlobalMesh( const std::array<Scalar, N>& lo, const std::array<Scalar, N>& hi,
                         const std::array<int, N>& num_cell )
{
    return std::make_shared<GlobalMesh<UniformMesh<Scalar, N>>>( lo, hi, num_cell );
}

template <class MeshType>
class GlobalGrid
{
  public:
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    GlobalGrid( MPI_Comm comm, const std::shared_ptr<GlobalMesh<MeshType>>& mesh,
                const std::array<bool, num_space_dim>& periodic, const BlockPartitioner<num_space_dim>& part )
        : _comm( comm )
        , _mesh( mesh )
        , _periodic( periodic )
    {
        std::array<int, num_space_dim> n;
        for ( std::size_t d = 0; d < num_space_dim; ++d )
            n[d] = mesh->globalNumCell( d );
        _ranks = part.ranksPerDimension( comm, n );
        int total = 1;
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            total *= _ranks[d];
            if ( _periodic[d] )
                throw std::logic_error( "refshim: periodic grids are not supported" );
        }
        if ( total != 1 )
            throw std::logic_error( "refshim: one rank only" );
    }
    MPI_Comm comm() const { return _comm; }
    const GlobalMesh<MeshType>& globalMesh() const { return *_mesh; }
    bool isPeriodic( int d ) const { return _periodic[d]; }
    bool onLowBoundary( int ) const { return true; }
    bool onHighBoundary( int ) const { return true; }
    int totalNumBlock() const { return 1; }
    int blockId() const { return 0; }
    int dimNumBlock( int d ) const { return _ranks[d]; }
    int dimBlockId( int ) const { return 0; }
    int globalNumEntity( Cell, int d ) const { return _mesh->globalNumCell( d ); }
    int globalNumEntity( Node, int d ) const { return _mesh->globalNumCell( d ) + 1; }
    template <int D>
    int globalNumEntity( Face<D>, int d ) const
    {
        return _mesh->globalNumCell( d ) + ( d == D ? 1 : 0 );
    }
    int ownedNumCell( int d ) const { return _mesh->globalNumCell( d ); }
    int globalOffset( int ) const { return 0; }

  private:
    MPI_Comm _comm;
    std::shared_ptr<GlobalMesh<MeshType>> _mesh;
    std::array<bool, num_space_dim> _periodic;
    std::array<int, num_space_dim> _ranks;
};

template <class MeshType>
std::shared_ptr<GlobalGrid<MeshType>>
createGlobalGrid( MPI_Comm comm, const std::shared_ptr<GlobalMesh<MeshType>>& mesh,
                  const std::array<bool, MeshType::num_space_dim>& periodic,
                  const BlockPartitioner<MeshType::num_space_dim>& partitioner )
{
    return std::make_shared<GlobalGrid<MeshType>>( comm, mesh, periodic, partitioner );
}

template <long N>
class IndexSpace
{
  public:
    IndexSpace()
    {
        for ( long d = 0; d < N; ++d )
            _min[d] = _max[d] = 0;
    }
    IndexSpace( const std::array<long, N>& lo, const std::array<long, N>& hi )
        : _min( lo )
        , _max( hi )
    {
    }
    long min( long d ) const { return _min[d]; }
    long max( long d ) const { return _max[d]; }
    long extent( long d ) const { return _max[d] - _min[d]; }
    long size() const
    {
        long s = 1;
        for ( long d = 0; d < N; ++d )
            s *= extent( d );
        return s;
    }

  private:
    std::array<long, N> _min, _max;
};

template <long N, class ExecutionSpace>
Kokkos::MDRangePolicy<ExecutionSpace, Kokkos::Rank<N>> createExecutionPolicy( const IndexSpace<N>& s,
                                                                             const ExecutionSpace& )
{
    static_assert( N == 2, "refshim: 2-D index spaces only" );
    Kokkos::MDRangePolicy<ExecutionSpace, Kokkos::Rank<N>> p;
    for ( int d = 0; d < 2; ++d )
    {
        p.lo[d] = s.min( d );
        p.hi[d] = s.max( d );
    }
    return p;
}

// [Cajita-mem] LocalGrid: halo cells are allocated on every side of the block, physical walls
// included (tests/tstMesh.cpp:61-68 asserts n + 2*halo (+1 along the face normal) for the ghosted
// Face<I> space of a single-rank non-periodic grid).  Local indices: owned cells [halo, halo + n);
// owned nodes / faces get one more along their normal on the block that touches the high wall.
template <class MeshType>
class LocalGrid
{
  public:
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    LocalGrid( const std::shared_ptr<GlobalGrid<MeshType>>& gg, int halo )
        : _gg( gg )
        , _halo( halo )
    {
    }
    const GlobalGrid<MeshType>& globalGrid() const { return *_gg; }
    int haloCellWidth() const { return _halo; }

    template <class Entity>
    IndexSpace<num_space_dim> indexSpace( Own, Entity, Local ) const
    {
        std::array<long, num_space_dim> lo, hi;
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            lo[d] = _halo;
            hi[d] = _halo + _gg->ownedNumCell( d ) +
                    ( _gg->onHighBoundary( d ) ? Impl::EntityExtra<Entity>::extra( d ) : 0 );
        }
        return IndexSpace<num_space_dim>( lo, hi );
    }
    template <class Entity>
    IndexSpace<num_space_dim> indexSpace( Ghost, Entity, Local ) const
    {
        std::array<long, num_space_dim> lo, hi;
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            lo[d] = 0;
            hi[d] = _gg->ownedNumCell( d ) + 2 * _halo + Impl::EntityExtra<Entity>::extra( d );
        }
        return IndexSpace<num_space_dim>( lo, hi );
    }
    template <class Entity>
    IndexSpace<num_space_dim> indexSpace( Own, Entity, Global ) const
    {
        std::array<long, num_space_dim> lo, hi;
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            lo[d] = _gg->globalOffset( d );
            hi[d] = lo[d] + _gg->ownedNumCell( d ) +
                    ( _gg->onHighBoundary( d ) ? Impl::EntityExtra<Entity>::extra( d ) : 0 );
        }
        return IndexSpace<num_space_dim>( lo, hi );
    }
    // one rank, non-periodic: the only neighbour is myself at offset 0
    int neighborRank( const std::array<int, num_space_dim>& off ) const
    {
        for ( std::size_t d = 0; d < num_space_dim; ++d )
            if ( off[d] != 0 )
                return -1;
        return 0;
    }
    template <class Decomposition, class Entity>
    IndexSpace<num_space_dim> sharedIndexSpace( Decomposition, Entity, const std::array<int, num_space_dim>& off,
                                                int = -1 ) const
    {
        if ( neighborRank( off ) < 0 )
            return IndexSpace<num_space_dim>(); // no neighbour: empty
        return indexSpace( Own(), Entity(), Local() );
    }

  private:
    std::shared_ptr<GlobalGrid<MeshType>> _gg;
    int _halo;
};

template <class MeshType>
std::shared_ptr<LocalGrid<MeshType>> createLocalGrid( const std::shared_ptr<GlobalGrid<MeshType>>& gg, int halo )
{
    return std::make_shared<LocalGrid<MeshType>>( gg, halo );
}

// [Cajita-mem] LocalMesh: geometry of the ghosted block.  Ghosted low corner = global low corner +
// cell * (global offset - halo); an entity's coordinate is low + (index + 1/2) * cell in the
// directions where it is cell-centred and low + index * cell where it is node-centred.
template <class Device, class MeshType>
class LocalMesh
{
  public:
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    LocalMesh() = default;
    explicit LocalMesh( const LocalGrid<MeshType>& lg )
    {
        const auto& gg = lg.globalGrid();
        const auto& gm = gg.globalMesh();
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            _cell[d] = gm.cellSize( d );
            _own_low[d] = gm.lowCorner( d ) + _cell[d] * gg.globalOffset( d );
            _own_high[d] = gm.lowCorner( d ) + _cell[d] * ( gg.globalOffset( d ) + gg.ownedNumCell( d ) );
            _ghost_low[d] = _own_low[d] - lg.haloCellWidth() * _cell[d];
            _ghost_high[d] = _own_high[d] + lg.haloCellWidth() * _cell[d];
        }
    }
    double lowCorner( Own, int d ) const { return _own_low[d]; }
    double highCorner( Own, int d ) const { return _own_high[d]; }
    double lowCorner( Ghost, int d ) const { return _ghost_low[d]; }
    double highCorner( Ghost, int d ) const { return _ghost_high[d]; }
    double cellSize( int d ) const { return _cell[d]; }

    template <class Entity, class Int, class Scalar>
    inline void coordinates( Entity, const Int index[], Scalar x[] ) const
    {
        for ( std::size_t d = 0; d < num_space_dim; ++d )
        {
            if ( Impl::EntityExtra<Entity>::extra( d ) )
                x[d] = _ghost_low[d] + static_cast<Scalar>( index[d] ) * _cell[d];
            else
                x[d] = _ghost_low[d] + ( static_cast<Scalar>( index[d] ) + Scalar( 0.5 ) ) * _cell[d];
        }
    }

  private:
    double _cell[num_space_dim], _own_low[num_space_dim], _own_high[num_space_dim], _ghost_low[num_space_dim],
        _ghost_high[num_space_dim];
};

template <class Device, class MeshType>
LocalMesh<Device, MeshType> createLocalMesh( const LocalGrid<MeshType>& lg )
{
    return LocalMesh<Device, MeshType>( lg );
}

// [Cajita-mem] IndexConversion::createL2G: global = local - (first owned local index) + global offset
namespace IndexConversion
{
template <class MeshType, class Entity>
struct L2G
{
    int own_min[MeshType::num_space_dim];
    int global_off[MeshType::num_space_dim];
    inline void operator()( const int li, const int lj, int& gi, int& gj ) const
    {
        gi = li - own_min[0] + global_off[0];
        gj = lj - own_min[1] + global_off[1];
    }
};
template <class MeshType, class Entity>
L2G<MeshType, Entity> createL2G( const LocalGrid<MeshType>& lg, Entity )
{
    L2G<MeshType, Entity> f;
    auto own = lg.indexSpace( Own(), Entity(), Local() );
    for ( std::size_t d = 0; d < MeshType::num_space_dim; ++d )
    {
        f.own_min[d] = static_cast<int>( own.min( d ) );
        f.global_off[d] = lg.globalGrid().globalOffset( d );
    }
    return f;
}
} // namespace IndexConversion

// ---- arrays -----------------------------------------------------------------------------------
template <class Entity, class MeshType>
class ArrayLayout
{
  public:
    using entity_type = Entity;
    using mesh_type = MeshType;
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    ArrayLayout( const std::shared_ptr<LocalGrid<MeshType>>& lg, int dofs )
        : _lg( lg )
        , _dofs( dofs )
    {
    }
    const std::shared_ptr<LocalGrid<MeshType>> localGrid() const { return _lg; }
    int dofsPerEntity() const { return _dofs; }
    template <class Decomposition, class IndexType>
    IndexSpace<num_space_dim> indexSpace( Decomposition dec, IndexType idx ) const
    {
        return _lg->indexSpace( dec, Entity(), idx );
    }

  private:
    std::shared_ptr<LocalGrid<MeshType>> _lg;
    int _dofs;
};

template <class Entity, class MeshType>
std::shared_ptr<ArrayLayout<Entity, MeshType>> createArrayLayout( const std::shared_ptr<LocalGrid<MeshType>>& lg,
                                                                  int dofs, Entity )
{
    return std::make_shared<ArrayLayout<Entity, MeshType>>( lg, dofs );
}

template <class Scalar, class Entity, class MeshType, class MemorySpace>
class Array
{
  public:
    static_assert( MeshType::num_space_dim == 2, "refshim: the reference is 2-D" );
    using value_type = Scalar;
    using entity_type = Entity;
    using mesh_type = MeshType;
    using memory_space = MemorySpace;
    using execution_space = Kokkos::DefaultHostExecutionSpace;
    using device_type = Kokkos::Device<execution_space, MemorySpace>;
    using array_layout = ArrayLayout<Entity, MeshType>;
    using view_type = Kokkos::View<Scalar***, Kokkos::LayoutRight, device_type>;

    Array( const std::string& label, const std::shared_ptr<array_layout>& layout )
        : _label( label )
        , _layout( layout )
    {
        auto g = layout->indexSpace( Ghost(), Local() );
        _view = view_type( label, g.extent( 0 ), g.extent( 1 ), layout->dofsPerEntity() );
    }
    const std::string& label() const { return _label; }
    std::shared_ptr<array_layout> layout() const { return _layout; }
    view_type view() const { return _view; }

  private:
    std::string _label;
    std::shared_ptr<array_layout> _layout;
    view_type _view;
};

template <class Scalar, class MemorySpace, class Entity, class MeshType>
std::shared_ptr<Array<Scalar, Entity, MeshType, MemorySpace>>
createArray( const std::string& label, const std::shared_ptr<ArrayLayout<Entity, MeshType>>& layout )
{
    return std::make_shared<Array<Scalar, Entity, MeshType, MemorySpace>>( label, layout );
}

namespace ArrayOp
{
template <class Array_t, class Decomposition>
void assign( Array_t& a, const typename Array_t::value_type value, Decomposition dec )
{
    auto v = a.view();
    auto s = a.layout()->indexSpace( dec, Local() );
    for ( long i = s.min( 0 ); i < s.max( 0 ); ++i )
        for ( long j = s.min( 1 ); j < s.max( 1 ); ++j )
            for ( int k = 0; k < v.extent_int( 2 ); ++k )
                v( (int)i, (int)j, k ) = value;
}
} // namespace ArrayOp

// ---- halos ------------------------------------------------------------------------------------
template <std::size_t N>
struct NodeHaloPattern
{
};
template <std::size_t N>
struct FaceHaloPattern
{
};

// One rank, non-periodic: every neighbour rank is -1, so a gather moves nothing and the ghost
// entities on the physical walls keep the zeros they were assigned at allocation.
template <class MemorySpace>
class Halo
{
  public:
    template <class ExecutionSpace, class... Arrays>
    void gather( const ExecutionSpace&, const Arrays&... ) const
    {
    }
    template <class ExecutionSpace, class... Arrays>
    void scatter( const ExecutionSpace&, const Arrays&... ) const
    {
    }
};

namespace Impl
{
template <class A0, class...>
struct FirstArray
{
    using type = A0;
};
} // namespace Impl

// createHalo( pattern, width, arrays... )   (src/ProblemManager.hpp:174-176)
template <std::size_t N, class... Arrays>
std::shared_ptr<Halo<typename Impl::FirstArray<Arrays...>::type::memory_space>>
createHalo( const NodeHaloPattern<N>&, const int, const Arrays&... )
{
    return std::make_shared<Halo<typename Impl::FirstArray<Arrays...>::type::memory_space>>();
}
// createHalo<Scalar, MemorySpace>( layout, pattern, width )   (src/VelocityCorrector.hpp:112-113)
template <class Scalar, class MemorySpace, class Entity, class MeshType, std::size_t N>
std::shared_ptr<Halo<MemorySpace>> createHalo( const ArrayLayout<Entity, MeshType>&, const FaceHaloPattern<N>&,
                                               const int )
{
    return std::make_shared<Halo<MemorySpace>>();
}

// ---- splines ----------------------------------------------------------------------------------
// [Cajita-mem] Spline<Order>: logical coordinate x = (p - position of entity 0) / cell;
//   order 1: knots int(x), int(x)+1, weights (1 - f, f), f = x - int(x);
//   order 3: knots int(x)-1 .. int(x)+2, cubic B-spline weights evaluated from the distance to the
//            first knot xn = f + 1, then xn -= 1 per knot.
template <int Order>
struct Spline;
template <>
struct Spline<1>
{
    static constexpr int num_knot = 2;
    template <class Scalar>
    static inline void stencil( const Scalar x0, int indices[2] )
    {
        indices[0] = static_cast<int>( x0 );
        indices[1] = indices[0] + 1;
    }
    template <class Scalar>
    static inline void value( const Scalar x0, Scalar values[2] )
    {
        const Scalar xn = x0 - static_cast<int>( x0 );
        values[0] = Scalar( 1.0 ) - xn;
        values[1] = xn;
    }
};
template <>
struct Spline<3>
{
    static constexpr int num_knot = 4;
    template <class Scalar>
    static inline void stencil( const Scalar x0, int indices[4] )
    {
        indices[0] = static_cast<int>( x0 ) - 1;
        indices[1] = indices[0] + 1;
        indices[2] = indices[1] + 1;
        indices[3] = indices[2] + 1;
    }
    template <class Scalar>
    static inline void value( const Scalar x0, Scalar values[4] )
    {
        const Scalar one_sixth = 1.0 / 6.0;
        const Scalar two_thirds = one_sixth * 4.0;
        const Scalar four_thirds = 2.0 * two_thirds;
        // knot at i - 1
        Scalar xn = x0 - static_cast<int>( x0 ) + 1.0;
        Scalar xn2 = xn * xn;
        values[0] = -xn * xn2 * one_sixth + xn2 - 2.0 * xn + four_thirds;
        // knot at i
        xn -= 1.0;
        xn2 = xn * xn;
        values[1] = 0.5 * xn * xn2 - xn2 + two_thirds;
        // knot at i + 1
        xn -= 1.0;
        xn2 = xn * xn;
        values[2] = -0.5 * xn * xn2 - xn2 + two_thirds;
        // knot at i + 2
        xn -= 1.0;
        xn2 = xn * xn;
        values[3] = xn * xn2 * one_sixth + xn2 + 2.0 * xn + four_thirds;
    }
};

template <class Scalar, int Order, std::size_t NumSpaceDim, class Entity>
struct SplineData
{
    static constexpr int order = Order;
    static constexpr int num_knot = Spline<Order>::num_knot;
    static constexpr std::size_t num_space_dim = NumSpaceDim;
    using entity_type = Entity;
    Scalar w[NumSpaceDim][num_knot];
    int s[NumSpaceDim][num_knot];
};

template <class LocalMeshType, class Scalar, int Order, std::size_t N, class Entity>
inline void evaluateSpline( const LocalMeshType& local_mesh, const Scalar p[N],
                            SplineData<Scalar, Order, N, Entity>& data )
{
    const int zero[N] = {};
    Scalar low[N];
    local_mesh.coordinates( Entity(), zero, low );
    for ( std::size_t d = 0; d < N; ++d )
    {
        const Scalar rdx = Scalar( 1.0 ) / local_mesh.cellSize( d );
        const Scalar x = ( p[d] - low[d] ) * rdx;
        Spline<Order>::stencil( x, data.s[d] );
        Spline<Order>::value( x, data.w[d] );
    }
}

namespace G2P
{
// scalar grid value -> point, 2-D
template <class View_t, class Scalar, int Order, class Entity>
inline void value( const View_t& view, const SplineData<Scalar, Order, 2, Entity>& sd, Scalar& result )
{
    result = 0.0;
    for ( int i = 0; i < SplineData<Scalar, Order, 2, Entity>::num_knot; ++i )
        for ( int j = 0; j < SplineData<Scalar, Order, 2, Entity>::num_knot; ++j )
            result += view( sd.s[0][i], sd.s[1][j], 0 ) * sd.w[0][i] * sd.w[1][j];
}
} // namespace G2P

// ---- structured solvers -----------------------------------------------------------------------
template <class Scalar, class Entity, class MeshType, class MemorySpace>
class ReferenceStructuredSolver
{
  public:
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;
    using Array_t = Array<Scalar, Entity, MeshType, MemorySpace>;
    virtual ~ReferenceStructuredSolver() = default;
    virtual void setMatrixStencil( const std::vector<std::array<int, num_space_dim>>& stencil,
                                   const bool is_symmetric = false ) = 0;
    virtual const Array_t& getMatrixValues() = 0;
    virtual void setPreconditionerStencil( const std::vector<std::array<int, num_space_dim>>& stencil,
                                           const bool is_symmetric = false ) = 0;
    virtual const Array_t& getPreconditionerValues() = 0;
    virtual void setTolerance( const double tol ) = 0;
    virtual void setMaxIter( const int max_iter ) = 0;
    virtual void setPrintLevel( const int print_level ) = 0;
    virtual void setup() = 0;
    virtual void solve( const Array_t& b, Array_t& x ) = 0;
    virtual int getNumIter() = 0;
    virtual double getFinalRelativeResidualNorm() = 0;
};

namespace Impl
{
// double-double accumulator (TwoSum), the arithmetic of oracle/cfo_oracle.cpp's acc_t
struct DDAcc
{
    double hi = 0.0, lo = 0.0;
    bool exact;
    explicit DDAcc( bool e )
        : exact( e )
    {
    }
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
    inline double value() const { return hi + lo; }
};
} // namespace Impl

// [Cajita-mem] ReferenceConjugateGradient: preconditioned CG on a stencil matrix stored per cell,
// x0 = whatever x holds on entry, absolute stopping test sqrt(sum r^2) <= tolerance, at most
// max_iter iterations, std::runtime_error when it does not converge.  Loop structure per
// residual kernel, preconditioner kernel, direction kernel, operator kernel, one
// global sum after each of the three reductions.  (One rank: the gathers and all-reduces vanish.)
template <class Scalar, class Entity, class MeshType, class MemorySpace>
class ReferenceConjugateGradient : public ReferenceStructuredSolver<Scalar, Entity, MeshType, MemorySpace>
{
  public:
    using base = ReferenceStructuredSolver<Scalar, Entity, MeshType, MemorySpace>;
    using Array_t = typename base::Array_t;
    using layout_t = ArrayLayout<Entity, MeshType>;
    static constexpr std::size_t num_space_dim = MeshType::num_space_dim;

    explicit ReferenceConjugateGradient( const layout_t& layout )
        : _lg( layout.localGrid() )
    {
    }
    void setMatrixStencil( const std::vector<std::array<int, num_space_dim>>& stencil, const bool = false ) override
    {
        _A_stencil = stencil;
        auto l = createArrayLayout( _lg, (int)stencil.size(), Entity() );
        _A = createArray<Scalar, MemorySpace>( "matrix", l );
        ArrayOp::assign( *_A, 0.0, Ghost() );
    }
    const Array_t& getMatrixValues() override { return *_A; }
    void setPreconditionerStencil( const std::vector<std::array<int, num_space_dim>>& stencil,
                                   const bool = false ) override
    {
        _M_stencil = stencil;
        auto l = createArrayLayout( _lg, (int)stencil.size(), Entity() );
        _M = createArray<Scalar, MemorySpace>( "preconditioner", l );
        ArrayOp::assign( *_M, 0.0, Ghost() );
    }
    const Array_t& getPreconditionerValues() override { return *_M; }
    void setTolerance( const double tol ) override { _tol = tol; }
    void setMaxIter( const int n ) override { _max_iter = n; }
    void setPrintLevel( const int p ) override { _print = p; }
    void setup() override
    {
        auto l = createArrayLayout( _lg, 1, Entity() );
        for ( auto* v : { &_r, &_z, &_p, &_q } )
        {
            *v = createArray<Scalar, MemorySpace>( "cg_work", l );
            ArrayOp::assign( **v, 0.0, Ghost() );
        }
    }
    int getNumIter() override { return _num_iter; }
    double getFinalRelativeResidualNorm() override { return _resid; }
    const std::vector<double>& history() const { return _hist; }
    void setFixedIterations( int n ) { _fixed = n; }

    void solve( const Array_t& b_arr, Array_t& x_arr ) override
    {
        if ( !_M )
            throw std::logic_error( "refshim CG: no preconditioner set" );
        const bool exact = cfref::knobs().cg_exact != 0;
        const int print = cfref::knobs().cg_print >= 0 ? cfref::knobs().cg_print : _print;
        auto own = _lg->indexSpace( Own(), Entity(), Local() );
        const int i0 = (int)own.min( 0 ), i1 = (int)own.max( 0 ), j0 = (int)own.min( 1 ), j1 = (int)own.max( 1 );
        auto A = _A->view();
        auto M = _M->view();
        auto b = b_arr.view();
        auto x = x_arr.view();
        auto r = _r->view(), z = _z->view(), p = _p->view(), q = _q->view();
        const int na = (int)_A_stencil.size(), nm = (int)_M_stencil.size();
        // `acc += a * b`: one fused multiply-add in the oracle's arithmetic, two roundings otherwise
        auto mac = [exact]( double a, double bb, double acc ) { return exact ? std::fma( a, bb, acc ) : acc + a * bb; };
        auto apply = [&]( const std::vector<std::array<int, num_space_dim>>& st, int n, const auto& C,
                          const auto& v, int i, int j ) {
            double s = 0.0;
            for ( int c = 0; c < n; ++c )
                s = mac( C( i, j, c ), v( i + st[c][0], j + st[c][1], 0 ), s );
            return s;
        };
        const int max_iter = _fixed > 0 ? _fixed : _max_iter;
        _num_iter = 0;
        _hist.clear();

        // r = b - A x ; sum r^2
        Impl::DDAcc rr( exact );
        for ( int i = i0; i < i1; ++i )
            for ( int j = j0; j < j1; ++j )
            {
                const double rn = b( i, j, 0 ) - apply( _A_stencil, na, A, x, i, j );
                r( i, j, 0 ) = rn;
                rr.add( rn * rn );
            }
        _resid = std::sqrt( rr.value() );
        bool converged = false;
        if ( _fixed <= 0 && _resid <= _tol )
            converged = true;
        else
        {
            // z = M r ; p = z ; sum z.r
            Impl::DDAcc zr( exact );
            for ( int i = i0; i < i1; ++i )
                for ( int j = j0; j < j1; ++j )
                {
                    const double zn = precondition( exact, nm, M, r, i, j );
                    z( i, j, 0 ) = zn;
                    p( i, j, 0 ) = zn;
                    zr.add( zn * r( i, j, 0 ) );
                }
            double zr_old = zr.value();
            // q = A p ; sum p.q
            Impl::DDAcc pq( exact );
            for ( int i = i0; i < i1; ++i )
                for ( int j = j0; j < j1; ++j )
                {
                    const double Ap = apply( _A_stencil, na, A, p, i, j );
                    q( i, j, 0 ) = Ap;
                    pq.add( p( i, j, 0 ) * Ap );
                }
            double pAp = pq.value();
            while ( _num_iter < max_iter )
            {
                const double alpha = zr_old / pAp;
                Impl::DDAcc rr2( exact );
                for ( int i = i0; i < i1; ++i )
                    for ( int j = j0; j < j1; ++j )
                    {
                        x( i, j, 0 ) = mac( alpha, p( i, j, 0 ), x( i, j, 0 ) );
                        const double rn = mac( -alpha, q( i, j, 0 ), r( i, j, 0 ) );
                        r( i, j, 0 ) = rn;
                        rr2.add( rn * rn );
                    }
                _resid = std::sqrt( rr2.value() );
                ++_num_iter;
                _hist.push_back( _resid );
                if ( print == 2 )
                    std::printf( "Cajita CG Iteration %d: |r|_2 = %g\n", _num_iter, _resid );
                if ( _fixed <= 0 && _resid <= _tol )
                {
                    converged = true;
                    break;
                }
                Impl::DDAcc zr2( exact );
                for ( int i = i0; i < i1; ++i )
                    for ( int j = j0; j < j1; ++j )
                    {
                        const double zn = precondition( exact, nm, M, r, i, j );
                        z( i, j, 0 ) = zn;
                        zr2.add( zn * r( i, j, 0 ) );
                    }
                const double zr_new = zr2.value();
                const double beta = zr_new / zr_old;
                for ( int i = i0; i < i1; ++i )
                    for ( int j = j0; j < j1; ++j )
                        p( i, j, 0 ) = mac( beta, p( i, j, 0 ), z( i, j, 0 ) );
                Impl::DDAcc pq2( exact );
                for ( int i = i0; i < i1; ++i )
                    for ( int j = j0; j < j1; ++j )
                    {
                        const double Ap = apply( _A_stencil, na, A, p, i, j );
                        q( i, j, 0 ) = Ap;
                        pq2.add( p( i, j, 0 ) * Ap );
                    }
                pAp = pq2.value();
                zr_old = zr_new;
            }
        }
        if ( print >= 1 )
            std::printf( "Cajita CG Finished in %d iterations, |r|_2 = %g\n", _num_iter, _resid );
        if ( !converged && _fixed <= 0 )
            throw std::runtime_error( "Cajita CG solver did not converge" );
    }

  private:
    // z = sum_c M(c) * r(+off_c).  A one-entry (diagonal) stencil is a plain product in both
    // arithmetics: 0 + M*r needs no accumulation.
    template <class MView, class RView>
    inline double precondition( bool exact, int nm, const MView& M, const RView& r, int i, int j ) const
    {
        if ( nm == 1 && _M_stencil[0][0] == 0 && _M_stencil[0][1] == 0 )
            return M( i, j, 0 ) * r( i, j, 0 );
        double s = 0.0;
        for ( int c = 0; c < nm; ++c )
        {
            const double t = r( i + _M_stencil[c][0], j + _M_stencil[c][1], 0 );
            s = exact ? std::fma( M( i, j, c ), t, s ) : s + M( i, j, c ) * t;
        }
        return s;
    }

    std::shared_ptr<LocalGrid<MeshType>> _lg;
    std::vector<std::array<int, num_space_dim>> _A_stencil, _M_stencil;
    std::shared_ptr<Array_t> _A, _M, _r, _z, _p, _q;
    double _tol = 1.0e-6;
    int _max_iter = 1000;
    int _print = 0;
    int _fixed = 0;
    int _num_iter = 0;
    double _resid = 0.0;
    std::vector<double> _hist;
};

template <class Scalar, class MemorySpace, class Entity, class MeshType>
std::shared_ptr<ReferenceConjugateGradient<Scalar, Entity, MeshType, MemorySpace>>
createReferenceConjugateGradient( const ArrayLayout<Entity, MeshType>& layout, const bool = false )
{
    return std::make_shared<ReferenceConjugateGradient<Scalar, Entity, MeshType, MemorySpace>>( layout );
}

// HYPRE is out of scope (north_star: "no HYPRE"): the type exists so that the reference's
// createVelocityCorrector compiles; using it throws.
template <class Scalar, class Entity, class MemorySpace>
class HypreStructuredSolver
{
  public:
    template <class Stencil>
    void setMatrixStencil( const Stencil&, const bool = false )
    {
        fail();
    }
    template <class Array_t>
    void setMatrixValues( const Array_t& )
    {
        fail();
    }
    void setTolerance( const double ) { fail(); }
    void setMaxIter( const int ) { fail(); }
    void setPrintLevel( const int ) { fail(); }
    void setPreconditioner( const std::shared_ptr<HypreStructuredSolver>& ) { fail(); }
    void setup() { fail(); }
    template <class Array_t>
    void solve( const Array_t&, Array_t& )
    {
        fail();
    }
    int getNumIter() { return 0; }
    double getFinalRelativeResidualNorm() { return 0.0; }

  private:
    [[noreturn]] static void fail()
    {
        throw std::runtime_error( "refshim: HYPRE solvers are not available (use -m Reference)" );
    }
};

template <class Scalar, class MemorySpace, class Entity, class MeshType>
std::shared_ptr<HypreStructuredSolver<Scalar, Entity, MemorySpace>>
createHypreStructuredSolver( const std::string&, const ArrayLayout<Entity, MeshType>&, const bool = false )
{
    throw std::runtime_error( "refshim: HYPRE solvers are not available (use -m Reference)" );
}

} // namespace Cajita

#endif
