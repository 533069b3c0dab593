// cuda_runtime.h — HOST STAND-IN for the CUDA runtime, used ONLY by tests/emul (see README there).
//
// The product's plain streaming kernels (no TMA, no shared memory beyond the reduction helper) are
// compiled as ordinary C++ and executed one CUDA thread after the other, so that their LOGIC — index
// arithmetic, statement order, host-side orchestration — can be checked bit for bit against the CPU
// oracle in a container without a GPU.  This is test infrastructure: nothing under cajitafluids_b200/
// builds, loads or ships it, and nothing measured ever runs through it.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <utility>
#include <type_traits>

#define CFB_HOST_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__( ... )
#define __grid_constant__
#define __shared__ static thread_local /* one block at a time per rank thread; fibers share it */
#define __align__( n ) __attribute__( ( aligned( n ) ) )

struct uint3
{
    unsigned x, y, z;
};
struct dim3
{
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3( unsigned a, unsigned b = 1, unsigned c = 1 ) : x( a ), y( b ), z( c ) {}
    dim3( int a ) : x( (unsigned)a ) {}
    dim3( long long a ) : x( (unsigned)a ) {}
};
struct double2
{
    double x, y;
};
inline double2 make_double2( double a, double b ) { return double2{ a, b }; }

namespace cfb_emul
{
extern thread_local uint3 g_threadIdx, g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;
// blocks in order; inside a block the threads run one after the other from the LAST to the first, so
// that thread 0 — the one that finalises block reductions — sees every other thread's contribution
void launch( dim3 grid, dim3 block, const std::function<void()>& body );
// cooperative form for kernels whose threads really meet at __syncthreads() (the peer-memory exchange
// kernel): every CUDA thread of a block is a fiber (ucontext); a fiber that reaches __syncthreads() yields,
// and all live fibers of the block pass the barrier together.  Blocks run in order.
void launch_coop( dim3 grid, dim3 block, const std::function<void()>& body );
void sync_threads(); // no-op outside launch_coop
bool coop_active();
long long clock_ns();
long long clock_ns_real();
// "shared memory" of the emulated block: thread-local storage of the rank thread.  Addresses in the
// shared window are 32-bit offsets from a thread-local anchor (all shared objects of a kernel — the static
// thread_local arrays that __shared__ turns into and the dynamic buffer — live in one TLS block).
unsigned char* dyn_smem();
uint32_t smem_addr( const void* p );
void* smem_ptr( uint32_t a );
} // namespace cfb_emul
#define threadIdx cfb_emul::g_threadIdx
#define blockIdx cfb_emul::g_blockIdx
#define blockDim cfb_emul::g_blockDim
#define gridDim cfb_emul::g_gridDim

template <class T>
inline T __ldg( const T* p ) { return *p; }
template <class T>
inline T __ldcg( const T* p ) { return *p; }
inline int min( int a, int b ) { return a < b ? a : b; }
inline int max( int a, int b ) { return a > b ? a : b; }
inline void __syncthreads() { cfb_emul::sync_threads(); }
inline void __threadfence() { __atomic_thread_fence( __ATOMIC_SEQ_CST ); }
inline void __threadfence_system() { __atomic_thread_fence( __ATOMIC_SEQ_CST ); }
inline unsigned atomicAdd( unsigned* p, unsigned v )
{
    unsigned o = *p;
    *p = o + v;
    return o;
}

// ---- runtime API (everything is synchronous host memory) ---------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmul = 1 };
typedef struct cfb_emul_stream* cudaStream_t;
typedef struct cfb_emul_event* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct cudaDeviceProp
{
    int major = 10, minor = 0, multiProcessorCount = 2;
};
struct cudaPitchedPtr
{
    void* ptr;
    size_t pitch, xsize, ysize;
};
struct cudaExtent
{
    size_t width, height, depth;
};
struct cudaMemcpy3DParms
{
    cudaPitchedPtr srcPtr{}, dstPtr{};
    cudaExtent extent{};
    cudaMemcpyKind kind = cudaMemcpyHostToHost;
};
inline cudaPitchedPtr make_cudaPitchedPtr( void* p, size_t pitch, size_t xs, size_t ys ) { return { p, pitch, xs, ys }; }
inline cudaExtent make_cudaExtent( size_t w, size_t h, size_t d ) { return { w, h, d }; }

enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
cudaError_t cudaGetDriverEntryPoint( const char* name, void** fn, unsigned long long flags,
                                     cudaDriverEntryPointQueryResult* res );
template <class F>
inline cudaError_t cudaFuncSetAttribute( F, cudaFuncAttribute, int ) { return cudaSuccess; }
inline const char* cudaGetErrorString( cudaError_t ) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount( int* n )
{
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties( cudaDeviceProp* p, int )
{
    *p = cudaDeviceProp();
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice( int ) { return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc( T** p, size_t bytes )
{
    *p = static_cast<T*>( std::malloc( bytes ? bytes : 1 ) );
    return *p ? cudaSuccess : cudaErrorEmul;
}
template <class T>
inline cudaError_t cudaMallocHost( T** p, size_t bytes ) { return cudaMalloc( p, bytes ); }
inline cudaError_t cudaFree( void* p )
{
    std::free( p );
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost( void* p ) { return cudaFree( p ); }
inline cudaError_t cudaMemsetAsync( void* p, int v, size_t n, cudaStream_t )
{
    std::memset( p, v, n );
    return cudaSuccess;
}
inline cudaError_t cudaMemset( void* p, int v, size_t n )
{
    std::memset( p, v, n );
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy( void* d, const void* s, size_t n, cudaMemcpyKind )
{
    std::memcpy( d, s, n );
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync( void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t )
{
    std::memcpy( d, s, n );
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy3DAsync( const cudaMemcpy3DParms* p, cudaStream_t )
{
    const char* s = static_cast<const char*>( p->srcPtr.ptr );
    char* d = static_cast<char*>( p->dstPtr.ptr );
    for ( size_t z = 0; z < p->extent.depth; ++z )
        for ( size_t y = 0; y < p->extent.height; ++y )
            std::memcpy( d + ( z * p->dstPtr.ysize + y ) * p->dstPtr.pitch,
                         s + ( z * p->srcPtr.ysize + y ) * p->srcPtr.pitch, p->extent.width );
    return cudaSuccess;
}
// "peer memory": the ranks of an emulated run are threads of one process, so a handle is the pointer itself
struct cudaIpcMemHandle_t
{
    char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle( cudaIpcMemHandle_t* h, void* p )
{
    std::memset( h, 0, sizeof( *h ) );
    std::memcpy( h->reserved, &p, sizeof( p ) );
    return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle( void** out, cudaIpcMemHandle_t h, unsigned )
{
    std::memcpy( out, h.reserved, sizeof( *out ) );
    return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle( void* ) { return cudaSuccess; }
inline long long clock64() { return cfb_emul::clock_ns(); }

inline cudaError_t cudaStreamCreateWithFlags( cudaStream_t* s, unsigned )
{
    *s = reinterpret_cast<cudaStream_t>( std::malloc( 1 ) );
    return cudaSuccess;
}
// cooperative launch: ONE block of fibers (a kernel with a grid barrier must work for any grid size, which is what lets
// it be checked here: its barrier degenerates to the block barrier); the argument pointers are
// dereferenced by the kernel's own parameter types
template <class... A, size_t... I>
inline void cfb_emul_call( void ( *f )( A... ), void** args, std::index_sequence<I...> )
{
    f( *static_cast<typename std::remove_cv<typename std::remove_reference<A>::type>::type*>( args[I] )... );
}
template <class... A>
inline cudaError_t cudaLaunchCooperativeKernel( void ( *f )( A... ), dim3, dim3 block, void** args, size_t, cudaStream_t )
{
    cfb_emul::launch_coop( dim3( 1 ), block, [=]() { cfb_emul_call( f, args, std::index_sequence_for<A...>() ); } );
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor( int* n, F, int, size_t )
{
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange( int* lo, int* hi )
{
    *lo = 0;
    *hi = -1;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithPriority( cudaStream_t* s, unsigned flags, int )
{
    return cudaStreamCreateWithFlags( s, flags );
}
inline cudaError_t cudaStreamDestroy( cudaStream_t s )
{
    std::free( s );
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize( cudaStream_t ) { return cudaSuccess; }
// CUDA graphs, kernel nodes only: while a capture is open the emulated launches are recorded (arguments by
// value, as a real launch copies them) instead of executed; cudaGraphLaunch replays them.
typedef struct cfb_emul_graph* cudaGraph_t;
typedef struct cfb_emul_graph* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1 };
cudaError_t cudaStreamBeginCapture( cudaStream_t, cudaStreamCaptureMode );
cudaError_t cudaStreamEndCapture( cudaStream_t, cudaGraph_t* );
cudaError_t cudaGraphInstantiate( cudaGraphExec_t*, cudaGraph_t, unsigned long long );
cudaError_t cudaGraphLaunch( cudaGraphExec_t, cudaStream_t );
cudaError_t cudaGraphDestroy( cudaGraph_t );
cudaError_t cudaGraphExecDestroy( cudaGraphExec_t );
inline cudaError_t cudaStreamWaitEvent( cudaStream_t, cudaEvent_t, unsigned ) { return cudaSuccess; }
// events carry the host time of their record, so that the library's CUDA-event timers return something
struct cfb_emul_event
{
    long long ns;
};
inline cudaError_t cudaEventCreate( cudaEvent_t* e )
{
    *e = new cfb_emul_event{ 0 };
    return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags( cudaEvent_t* e, unsigned ) { return cudaEventCreate( e ); }
inline cudaError_t cudaEventDestroy( cudaEvent_t e )
{
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord( cudaEvent_t e, cudaStream_t )
{
    e->ns = cfb_emul::clock_ns_real();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize( cudaEvent_t ) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime( float* ms, cudaEvent_t a, cudaEvent_t b )
{
    *ms = (float)( ( b->ns - a->ns ) * 1.0e-6 );
    return cudaSuccess;
}
