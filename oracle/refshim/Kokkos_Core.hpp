// Kokkos_Core.hpp — STAND-IN for the absent Kokkos dependency.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: let the UNMODIFIED cajitafluids sources (src/*.hpp, examples/advection.cpp, compiled
// from where they lie under /root/reference — never copied) build and run on one host rank, so
// that the CPU oracle (oracle/cfo_oracle.cpp) can be pinned against the reference's own statements
// instead of against a hand restatement only.  See oracle/refshim/README.md.
//
// This is NOT Kokkos: it is the minimal host-only subset of the Kokkos 3.x API surface that the
// reference touches, written from scratch:
//   Kokkos::Array, Kokkos::View<T***> (LayoutRight / LayoutLeft, reference-counted),
//   Kokkos::Device / HostSpace / Serial / OpenMP / DefaultHostExecutionSpace,
//   Kokkos::MDRangePolicy<Exec, Rank<2>> + parallel_for (serial, or `omp parallel for` over the
//   slow index for Kokkos::OpenMP), create_mirror_view_and_copy, Profiling::push/popRegion,
//   initialize / finalize / ScopeGuard, KOKKOS_LAMBDA / KOKKOS_(INLINE_)FUNCTION.
#ifndef CFREF_SHIM_KOKKOS_CORE_HPP
#define CFREF_SHIM_KOKKOS_CORE_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_LAMBDA [=]
#define KOKKOS_ENABLE_SERIAL 1
#ifdef _OPENMP
#define KOKKOS_ENABLE_OPENMP 1
#endif

namespace Kokkos
{

struct HostSpace
{
    using memory_space = HostSpace;
};
struct Serial
{
    using execution_space = Serial;
    using memory_space = HostSpace;
};
struct OpenMP
{
    using execution_space = OpenMP;
    using memory_space = HostSpace;
};
#ifdef _OPENMP
using DefaultHostExecutionSpace = OpenMP;
#else
using DefaultHostExecutionSpace = Serial;
#endif
using DefaultExecutionSpace = DefaultHostExecutionSpace;

template <class ExecutionSpace, class MemorySpace>
struct Device
{
    using execution_space = ExecutionSpace;
    using memory_space = MemorySpace;
    using device_type = Device;
};

struct LayoutLeft
{
};
struct LayoutRight
{
};

// Aggregate, like Kokkos::Array: `Array<int,4> a = { 0, 1, 2, 3 };` and `a = { ... };` work.
template <class T, std::size_t N>
struct Array
{
    T m_internal_implementation_private_member_data[N];
    using value_type = T;
    static constexpr std::size_t size() { return N; }
    T& operator[]( std::size_t i ) { return m_internal_implementation_private_member_data[i]; }
    const T& operator[]( std::size_t i ) const
    {
        return m_internal_implementation_private_member_data[i];
    }
    T* data() { return m_internal_implementation_private_member_data; }
    const T* data() const { return m_internal_implementation_private_member_data; }
};

namespace Impl
{
template <class... P>
struct has_layout_left : std::false_type
{
};
template <class P0, class... P>
struct has_layout_left<P0, P...>
    : std::conditional_t<std::is_same<P0, LayoutLeft>::value, std::true_type, has_layout_left<P...>>
{
};
template <class... P>
struct first_device
{
    using type = Device<DefaultHostExecutionSpace, HostSpace>;
};
template <class E, class M, class... P>
struct first_device<Device<E, M>, P...>
{
    using type = Device<E, M>;
};
template <class P0, class... P>
struct first_device<P0, P...>
{
    using type = typename first_device<P...>::type;
};
} // namespace Impl

template <class DataType, class... Properties>
class View;

// Rank-3 view of a reference-counted allocation.  operator() is const and returns a mutable
// reference (Kokkos view semantics: a const View is a const HANDLE), which is what lets the
// reference's by-value lambda captures write through their copies.
template <class T, class... Properties>
class View<T***, Properties...>
{
  public:
    using value_type = T;
    using device_type = typename Impl::first_device<Properties...>::type;
    using memory_space = typename device_type::memory_space;
    using execution_space = typename device_type::execution_space;
    static constexpr bool is_layout_left = Impl::has_layout_left<Properties...>::value;

    View() = default;
    View( const std::string& label, std::size_t n0, std::size_t n1, std::size_t n2 )
        : _label( label )
        , _n{ n0, n1, n2 }
        , _store( std::make_shared<std::vector<T>>( n0 * n1 * n2, T() ) )
        , _data( _store->data() )
    {
    }

    inline T& operator()( const int i, const int j, const int k ) const
    {
        if ( is_layout_left )
            return _data[( static_cast<std::size_t>( k ) * _n[1] + j ) * _n[0] + i];
        return _data[( static_cast<std::size_t>( i ) * _n[1] + j ) * _n[2] + k];
    }
    std::size_t extent( const int d ) const { return _n[d]; }
    int extent_int( const int d ) const { return static_cast<int>( _n[d] ); }
    std::size_t size() const { return _n[0] * _n[1] * _n[2]; }
    std::size_t span() const { return size(); }
    T* data() const { return _data; }
    const std::string& label() const { return _label; }

  private:
    std::string _label;
    std::size_t _n[3] = { 0, 0, 0 };
    std::shared_ptr<std::vector<T>> _store;
    T* _data = nullptr;
};

template <class T, class... P, class... Q>
void deep_copy( const View<T***, P...>& dst, const View<T***, Q...>& src )
{
    for ( int i = 0; i < dst.extent_int( 0 ); ++i )
        for ( int j = 0; j < dst.extent_int( 1 ); ++j )
            for ( int k = 0; k < dst.extent_int( 2 ); ++k )
                dst( i, j, k ) = src( i, j, k );
}

// Host-only: the mirror of a host view is a fresh allocation with the same layout and contents.
template <class Space, class T, class... P>
View<T***, P...> create_mirror_view_and_copy( const Space&, const View<T***, P...>& src )
{
    View<T***, P...> dst( src.label() + "_mirror", src.extent( 0 ), src.extent( 1 ), src.extent( 2 ) );
    deep_copy( dst, src );
    return dst;
}

template <unsigned N>
struct Rank
{
    static constexpr unsigned rank = N;
};

// 2-D iteration range [lo, hi) per dimension on an execution space.
template <class ExecutionSpace, class RankType = Rank<2>>
struct MDRangePolicy
{
    using execution_space = ExecutionSpace;
    long lo[2];
    long hi[2];
};

namespace Impl
{
template <class Functor>
inline void run_2d( Serial, const long lo[2], const long hi[2], const Functor& f )
{
    for ( long i = lo[0]; i < hi[0]; ++i )
        for ( long j = lo[1]; j < hi[1]; ++j )
            f( static_cast<int>( i ), static_cast<int>( j ) );
}
template <class Functor>
inline void run_2d( OpenMP, const long lo[2], const long hi[2], const Functor& f )
{
#pragma omp parallel for schedule( static )
    for ( long i = lo[0]; i < hi[0]; ++i )
        for ( long j = lo[1]; j < hi[1]; ++j )
            f( static_cast<int>( i ), static_cast<int>( j ) );
}
} // namespace Impl

template <class ExecutionSpace, class RankType, class Functor>
inline void parallel_for( const std::string&, const MDRangePolicy<ExecutionSpace, RankType>& policy,
                          const Functor& functor )
{
    Impl::run_2d( ExecutionSpace(), policy.lo, policy.hi, functor );
}
template <class ExecutionSpace, class RankType, class Functor>
inline void parallel_for( const MDRangePolicy<ExecutionSpace, RankType>& policy, const Functor& functor )
{
    Impl::run_2d( ExecutionSpace(), policy.lo, policy.hi, functor );
}

inline void fence() {}

namespace Profiling
{
inline void pushRegion( const std::string& ) {}
inline void popRegion() {}
} // namespace Profiling

inline void initialize( int&, char** ) {}
inline void initialize() {}
inline void finalize() {}
struct ScopeGuard
{
    ScopeGuard( int&, char** ) {}
    ScopeGuard() {}
};

} // namespace Kokkos

#endif
