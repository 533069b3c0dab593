// halo.cu — multi-GPU plumbing: one rank == one process == one GPU (like one MPI rank == one Kokkos
// device in the reference), ghost exchange and CG scalar reductions over NCCL (NVLink 5 / NVSwitch).
//
// Replaces:
//   Cajita::Halo::gather of the advection halo   src/ProblemManager.hpp:173-176,394-403
//       NodeHaloPattern (26 neighbours in 3-D), width = halo cell width, q,u,v(,w) jointly
//       -> three axis sweeps (x, then y incl. x ghosts, then z incl. x,y ghosts): 6 messages,
//          corners and edges arrive transitively; all fields travel in one message per neighbour.
//   Cajita::Halo::gather of the pressure halo    src/VelocityCorrector.hpp:112-113,236
//   the CG's per-iteration gather of p (width 1, face neighbours only)
//       -> one pack kernel for all faces, one ncclGroup of send/recv, one unpack kernel.
//   MPI_Allreduce of the CG scalars               (Cajita ReferenceConjugateGradient)
//       -> ncclAllReduce on the device-resident CgState; nothing comes back to the host.
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): a single-GPU run has no NCCL dependency and
// a torch process shares the NCCL it already loaded.  Two communicators are created so that halo
// traffic (side stream) and scalar reductions (main stream) never share a communicator.
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

namespace
{

struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t ( *GetUniqueId )( ncclUniqueId* ) = nullptr;
    ncclResult_t ( *CommInitRank )( ncclComm_t*, int, ncclUniqueId, int ) = nullptr;
    ncclResult_t ( *CommDestroy )( ncclComm_t ) = nullptr;
    ncclResult_t ( *Send )( const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t ) = nullptr;
    ncclResult_t ( *Recv )( void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t ) = nullptr;
    ncclResult_t ( *AllReduce )( const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                                 cudaStream_t ) = nullptr;
    ncclResult_t ( *AllGather )( const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t ) = nullptr;
    ncclResult_t ( *GroupStart )() = nullptr;
    ncclResult_t ( *GroupEnd )() = nullptr;
    const char* ( *GetErrorString )( ncclResult_t ) = nullptr;
};

NcclApi g_nccl;

bool nccl_load( std::string& err )
{
    if ( g_nccl.handle )
        return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for ( const char* n : names )
    {
        h = dlopen( n, RTLD_NOW | RTLD_GLOBAL );
        if ( h )
            break;
    }
    if ( !h )
    {
        err = std::string( "dlopen(libnccl.so.2) failed: " ) + dlerror();
        return false;
    }
#define BIND( field, sym )                                                                         \
    g_nccl.field = reinterpret_cast<decltype( g_nccl.field )>( dlsym( h, sym ) );                  \
    if ( !g_nccl.field )                                                                           \
    {                                                                                              \
        err = std::string( "NCCL symbol missing: " ) + sym;                                        \
        return false;                                                                              \
    }
    BIND( GetUniqueId, "ncclGetUniqueId" )
    BIND( CommInitRank, "ncclCommInitRank" )
    BIND( CommDestroy, "ncclCommDestroy" )
    BIND( Send, "ncclSend" )
    BIND( Recv, "ncclRecv" )
    BIND( AllReduce, "ncclAllReduce" )
    BIND( AllGather, "ncclAllGather" )
    BIND( GroupStart, "ncclGroupStart" )
    BIND( GroupEnd, "ncclGroupEnd" )
    BIND( GetErrorString, "ncclGetErrorString" )
#undef BIND
    g_nccl.handle = h;
    return true;
}

struct Comm
{
    ncclComm_t halo = nullptr; // send/recv on the side stream
    ncclComm_t red = nullptr;  // allreduce on the main stream
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};

#define CFB_NCCL( c, expr )                                                                        \
    do                                                                                             \
    {                                                                                              \
        ncclResult_t _r = ( expr );                                                                \
        if ( _r != ncclSuccess )                                                                   \
            return cfb_fail( ( c ), CFB_ERR_NCCL,                                                  \
                             std::string( #expr ) + ": " + g_nccl.GetErrorString( _r ) );          \
    } while ( 0 )

// A box of owned-index space [lo, hi) per dim, copied for `nf` fields into / out of a dense buffer.
struct Box
{
    int lo[3], hi[3];
    __host__ __device__ long long count() const
    {
        return (long long)( hi[0] - lo[0] ) * ( hi[1] - lo[1] ) * ( hi[2] - lo[2] );
    }
};

struct PackArgs
{
    int nbox;        // number of boxes handled by this launch (<= 6)
    int nf;          // fields per box
    Box box[6];
    double* buf[6];  // dense buffer of box b: [field][k][j][i]
    double* fld[4];  // field base pointers
};

template <bool PACK>
__global__ void __launch_bounds__( 256 )
    pack_kernel( const __grid_constant__ Geo g, const __grid_constant__ PackArgs a )
{
    const int b = blockIdx.y;
    if ( b >= a.nbox )
        return;
    const Box& bx = a.box[b];
    const int ex = bx.hi[0] - bx.lo[0], ey = bx.hi[1] - bx.lo[1];
    const long long per = bx.count();
    const long long total = per * a.nf;
    double* buf = a.buf[b];
    for ( long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (long long)gridDim.x * 256 )
    {
        const int f = (int)( t / per );
        const long long r = t - f * per;
        const int i = (int)( r % ex ) + bx.lo[0];
        const int j = (int)( ( r / ex ) % ey ) + bx.lo[1];
        const int k = (int)( r / ( (long long)ex * ey ) ) + bx.lo[2];
        double* p = a.fld[f] + geo_off( g, i, j, k );
        if ( PACK )
            buf[t] = *p;
        else
            *p = buf[t];
    }
}

int launch_pack( cfb_ctx* c, const PackArgs& a, bool pack, cudaStream_t st )
{
    long long mx = 1;
    for ( int b = 0; b < a.nbox; ++b )
        mx = std::max( mx, a.box[b].count() * a.nf );
    int gx = (int)std::min<long long>( ( mx + 255 ) / 256, (long long)c->sm_count * 8 );
    dim3 grid( gx, a.nbox );
    if ( pack )
        pack_kernel<true><<<grid, 256, 0, st>>>( c->g, a );
    else
        pack_kernel<false><<<grid, 256, 0, st>>>( c->g, a );
    c->stats.kernel_launches += 1;
    return 1;
}

inline int rank_of( const cfb_config& cfg, int bx, int by, int bz )
{
    return ( bz * cfg.ranks_per_dim[1] + by ) * cfg.ranks_per_dim[0] + bx;
}

// Exchange along the listed dims: boxes `send_lo/hi` go to the low/high neighbour, `recv_lo/hi` are
// filled from them.  All transfers of one call form a single NCCL group on stream `st`.
int exchange( cfb_ctx* c, int nf, double* const fld[4], int ndims, const int dims[3], const Box send_lo[3],
              const Box send_hi[3], const Box recv_lo[3], const Box recv_hi[3], cudaStream_t st )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    PackArgs pk{}, up{};
    pk.nf = up.nf = nf;
    for ( int f = 0; f < nf; ++f )
        pk.fld[f] = up.fld[f] = fld[f];
    struct Xfer
    {
        int peer;
        double *sbuf, *rbuf;
        size_t scount, rcount;
    } x[6];
    int nx = 0;
    for ( int q = 0; q < ndims; ++q )
    {
        const int d = dims[q];
        for ( int side = 0; side < 2; ++side )
        {
            const int peer = c->nbr[2 * d + side];
            if ( peer < 0 )
                continue;
            const Box& sb = side == 0 ? send_lo[q] : send_hi[q];
            const Box& rb = side == 0 ? recv_lo[q] : recv_hi[q];
            const int slot = 2 * d + side;
            if ( (size_t)( sb.count() * nf ) > c->halo_buf_elems || (size_t)( rb.count() * nf ) > c->halo_buf_elems )
                return cfb_fail( c, CFB_ERR_INVALID, "halo buffer too small" );
            pk.box[pk.nbox] = sb;
            pk.buf[pk.nbox++] = c->d_halo_send[slot];
            up.box[up.nbox] = rb;
            up.buf[up.nbox++] = c->d_halo_recv[slot];
            x[nx++] = { peer, c->d_halo_send[slot], c->d_halo_recv[slot], (size_t)( sb.count() * nf ),
                        (size_t)( rb.count() * nf ) };
        }
    }
    if ( nx == 0 )
        return CFB_OK;
    launch_pack( c, pk, true, st );
    CFB_NCCL( c, g_nccl.GroupStart() );
    for ( int i = 0; i < nx; ++i )
    {
        CFB_NCCL( c, g_nccl.Send( x[i].sbuf, x[i].scount, ncclDouble, x[i].peer, cm->halo, st ) );
        CFB_NCCL( c, g_nccl.Recv( x[i].rbuf, x[i].rcount, ncclDouble, x[i].peer, cm->halo, st ) );
    }
    CFB_NCCL( c, g_nccl.GroupEnd() );
    launch_pack( c, up, false, st );
    return CFB_OK;
}


// ---------------------------------------------------------------------------------------------
// NVLink peer-memory path (one kernel per reduction point of the CG iteration).
//
//   1. every block copies a share of my boundary layers straight into the neighbours' ghost layers
//      (their arrays are mapped here through cudaIpc; the stores travel over NVLink / NVSwitch);
//   2. all threads fence at system scope, the block draws a ticket; the block that draws the last one
//   3. publishes my local double-double sums and then the sequence number into every rank's mailbox,
//   4. waits until every rank's sequence number has arrived in MY mailbox (bounded spin), and
//   5. combines the W double-doubles in rank order: the global value is the correctly rounded exact
//      sum on every rank, bit for bit the same (see device_reduce.cuh).
//
// The two publications per iteration double as the barriers that make the ghost stores safe: a rank
// enters phase B only after every rank has finished phase A and its ghost stores (and vice versa), and
// p is double-buffered, so no ghost layer is overwritten while a neighbour may still read it.
struct XFace
{
    const double* src;
    double* dst;
    long long dorigin, dsy, dsz;
    int lo[3], ext[3], shift[3]; // my box (owned index space), peer index = my index - shift
    int packed;                  // 1: dst is a dense staging area, element t of the box goes to dst[t]
};

struct XchgArgs
{
    int nface;
    XFace f[12];
    CgState* S;
    PeerMail* mail[CFB_MAX_PEERS];
    unsigned int* ticket;
    int which, rank, world;
    long long timeout_cycles;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64( const unsigned long long* p )
{
    unsigned long long v;
    asm volatile( "ld.volatile.global.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );
    return v;
}

__global__ void __launch_bounds__( 256 )
    cg_xchg_kernel( const __grid_constant__ Geo g, const __grid_constant__ XchgArgs a )
{
    // 1. ghost stores into the neighbours
    for ( int fi = 0; fi < a.nface; ++fi )
    {
        const XFace& f = a.f[fi];
        const long long total = (long long)f.ext[0] * f.ext[1] * f.ext[2];
        for ( long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (long long)gridDim.x * 256 )
        {
            const int i = (int)( t % f.ext[0] ) + f.lo[0];
            const int j = (int)( ( t / f.ext[0] ) % f.ext[1] ) + f.lo[1];
            const int k = (int)( t / ( (long long)f.ext[0] * f.ext[1] ) ) + f.lo[2];
            const double v = f.src[geo_off( g, i, j, k )];
            if ( f.packed )
                f.dst[t] = v;
            else
                f.dst[f.dorigin + (long long)( k - f.shift[2] ) * f.dsz + (long long)( j - f.shift[1] ) * f.dsy +
                      ( i - f.shift[0] )] = v;
        }
    }
    // 2. my stores are performed system-wide before the ticket is drawn
    __threadfence_system();
    __shared__ bool s_last;
    __syncthreads();
    if ( threadIdx.x == 0 )
    {
        const unsigned t = atomicAdd( a.ticket, 1u );
        s_last = ( t == gridDim.x - 1 );
    }
    __syncthreads();
    if ( !s_last )
        return;
    __threadfence_system();

    CgState* S = a.S;
    __shared__ unsigned long long s_seq;
    if ( threadIdx.x == 0 )
    {
        *a.ticket = 0u;
        s_seq = ++S->seq[a.which];
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int nd = a.which == 0 ? 2 : 4; // doubles published
    const double* loc = &S->loc[a.which == 0 ? 0 : 2];
    // 3. publish: data, fence, sequence number (one thread per destination rank)
    if ( threadIdx.x < a.world )
    {
        PeerMail* m = a.mail[threadIdx.x];
        for ( int q = 0; q < nd; ++q )
            m->v[a.which][a.rank][q] = loc[q];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>( &m->seq[a.which][a.rank] ) = seq;
    }
    // 4. wait for every rank's publication in my mailbox
    PeerMail* me = a.mail[a.rank];
    if ( threadIdx.x < a.world && !S->xerror )
    {
        const long long t0 = clock64();
        while ( ld_volatile_u64( &me->seq[a.which][threadIdx.x] ) < seq )
        {
            if ( clock64() - t0 > a.timeout_cycles )
            {
                S->xerror = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    // 5. exact combination in rank order
    if ( threadIdx.x == 0 )
    {
        dd_t acc[2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
        const int nv = nd / 2;
        for ( int r = 0; r < a.world; ++r )
            for ( int v = 0; v < nv; ++v )
            {
                const volatile double* src = &me->v[a.which][r][2 * v];
                dd_t w = { src[0], src[1] };
                acc[v] = dd_add( acc[v], w );
            }
        if ( a.which == 0 )
            S->pAp = acc[0].hi + acc[0].lo;
        else
        {
            S->rz_new = acc[0].hi + acc[0].lo;
            S->rr = acc[1].hi + acc[1].lo;
        }
    }
}

// staging area -> ghost columns i = -1 / i = n[0] of up to 4 (side, array) pairs
struct XUnpackArgs
{
    int n;
    const double* src[4];
    double* dst[4];
    int col[4];
};

__global__ void __launch_bounds__( 256 )
    cg_xunpack_kernel( const __grid_constant__ Geo g, const __grid_constant__ XUnpackArgs a )
{
    const long long per = (long long)g.n[1] * g.n[2];
    for ( long long t = blockIdx.x * 256ll + threadIdx.x; t < per * a.n; t += (long long)gridDim.x * 256 )
    {
        const int q = (int)( t / per );
        const long long e = t - q * per;
        const int j = (int)( e % g.n[1] ), k = (int)( e / g.n[1] );
        a.dst[q][geo_off( g, a.col[q], j, k )] = a.src[q][e];
    }
}

// ---------------------------------------------------------------------------------------------
// Overlapped exchange ("peer_overlap"): the ghost stores of cg_xchg_kernel without its reduction, run on the side
// stream while the main stream computes cells that do not read ghosts.
//   1. every block stores a share of my boundary layers into the neighbours' ghost layers / x staging areas;
//   2. system fence, ticket; the block that draws the last one
//   3. tells every face neighbour "transfer number n of this kind has landed" (its mailbox, slot of the face I sit on),
//   4. waits until every face neighbour has told me the same (bounded spin).
// What follows on the side stream (the x scatter, the event the boundary units wait for) therefore sees complete
// ghost layers.  Why nobody overwrites a layer that is still being read: cfb_api.cu, enqueue_iteration.
struct FaceArgs
{
    int nface;
    XFace f[6];
    PeerMail* nbr_mail[6]; // mailbox of the neighbour behind face i of this launch
    int nbr_slot[6];       // ... and the slot I write there: the neighbour's face that looks at me
    int my_slot[6];        // the slot of my own mailbox that neighbour writes
    PeerMail* mail_self;
    CgState* S;
    unsigned int* ticket;
    int kind;
    long long timeout_cycles;
};

__global__ void __launch_bounds__( 256 )
    cg_face_kernel( const __grid_constant__ Geo g, const __grid_constant__ FaceArgs a )
{
    for ( int fi = 0; fi < a.nface; ++fi )
    {
        const XFace& f = a.f[fi];
        const long long total = (long long)f.ext[0] * f.ext[1] * f.ext[2];
        for ( long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (long long)gridDim.x * 256 )
        {
            const int i = (int)( t % f.ext[0] ) + f.lo[0];
            const int j = (int)( ( t / f.ext[0] ) % f.ext[1] ) + f.lo[1];
            const int k = (int)( t / ( (long long)f.ext[0] * f.ext[1] ) ) + f.lo[2];
            const double v = f.src[geo_off( g, i, j, k )];
            if ( f.packed )
                f.dst[t] = v;
            else
                f.dst[f.dorigin + (long long)( k - f.shift[2] ) * f.dsz + (long long)( j - f.shift[1] ) * f.dsy +
                      ( i - f.shift[0] )] = v;
        }
    }
    __threadfence_system();
    __shared__ bool s_last;
    __syncthreads();
    if ( threadIdx.x == 0 )
    {
        const unsigned t = atomicAdd( a.ticket, 1u );
        s_last = ( t == gridDim.x - 1 );
    }
    __syncthreads();
    if ( !s_last )
        return;
    __threadfence_system();
    CgState* S = a.S;
    __shared__ unsigned long long s_seq;
    if ( threadIdx.x == 0 )
    {
        *a.ticket = 0u;
        s_seq = ++S->fseq[a.kind];
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    if ( threadIdx.x < a.nface )
    {
        *reinterpret_cast<volatile unsigned long long*>( &a.nbr_mail[threadIdx.x]->fseq[a.kind][a.nbr_slot[threadIdx.x]] ) = seq;
        if ( !S->xerror )
        {
            const long long t0 = clock64();
            while ( ld_volatile_u64( &a.mail_self->fseq[a.kind][a.my_slot[threadIdx.x]] ) < seq )
            {
                if ( clock64() - t0 > a.timeout_cycles )
                {
                    S->xerror = 1;
                    break;
                }
            }
        }
    }
    __threadfence_system();
}

// what every rank tells the others at start-up
struct PeerInfo
{
    cudaIpcMemHandle_t h_r, h_p0, h_p1, h_mail, h_xstage;
    long long origin, sy, sz;
    int n[3];
    int device;
    int ok;
};

int peer_setup( cfb_ctx* c )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    const int W = c->cfg.world_size, me = c->cfg.world_rank;
    c->peer_ok = false;
    if ( W > CFB_MAX_PEERS )
        return CFB_OK;
    const char* env = std::getenv( "CFB_PEER" );
    if ( env && env[0] == '0' )
        return CFB_OK;
    CFB_CUDA( c, cudaMalloc( &c->mail_self, sizeof( PeerMail ) ) );
    CFB_CUDA( c, cudaMemset( c->mail_self, 0, sizeof( PeerMail ) ) );
    // [0]: cg_xchg_kernel (main stream), [1]: cg_face_kernel (side stream)
    CFB_CUDA( c, cudaMalloc( &c->d_xticket, 2 * sizeof( unsigned int ) ) );
    CFB_CUDA( c, cudaMemset( c->d_xticket, 0, 2 * sizeof( unsigned int ) ) );
    // [2 sides][r, pbuf 0, pbuf 1] slots of ny * nz doubles each (see xslot)
    const size_t xstage_elems = (size_t)6 * c->g.n[1] * c->g.n[2];
    CFB_CUDA( c, cudaMalloc( &c->xstage_self, xstage_elems * sizeof( double ) ) );
    CFB_CUDA( c, cudaMemset( c->xstage_self, 0, xstage_elems * sizeof( double ) ) );
    PeerInfo mine{};
    mine.ok = 1;
    if ( cudaIpcGetMemHandle( &mine.h_xstage, c->xstage_self ) != cudaSuccess ||
         cudaIpcGetMemHandle( &mine.h_r, c->cg_r ) != cudaSuccess ||
         cudaIpcGetMemHandle( &mine.h_p0, c->cg_pbuf[0] ) != cudaSuccess ||
         cudaIpcGetMemHandle( &mine.h_p1, c->cg_pbuf[1] ) != cudaSuccess ||
         cudaIpcGetMemHandle( &mine.h_mail, c->mail_self ) != cudaSuccess )
    {
        cudaGetLastError();
        mine.ok = 0;
    }
    mine.origin = c->g.origin;
    mine.sy = c->g.sy;
    mine.sz = c->g.sz;
    for ( int d = 0; d < 3; ++d )
        mine.n[d] = c->g.n[d];
    mine.device = c->device;
    // all-gather the PeerInfo records through NCCL (bytes)
    PeerInfo* d_all = nullptr;
    CFB_CUDA( c, cudaMalloc( &d_all, sizeof( PeerInfo ) * ( W + 1 ) ) );
    CFB_CUDA( c, cudaMemcpy( d_all + W, &mine, sizeof( PeerInfo ), cudaMemcpyHostToDevice ) );
    CFB_NCCL( c, g_nccl.AllGather( d_all + W, d_all, sizeof( PeerInfo ), ncclChar, cm->red, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    std::vector<PeerInfo> all( W );
    CFB_CUDA( c, cudaMemcpy( all.data(), d_all, sizeof( PeerInfo ) * W, cudaMemcpyDeviceToHost ) );
    cudaFree( d_all );
    bool ok = true;
    for ( int r = 0; r < W; ++r )
        ok = ok && all[r].ok;
    auto open = [&]( const cudaIpcMemHandle_t& h, void** out ) -> bool {
        if ( cudaIpcOpenMemHandle( out, h, cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess )
        {
            cudaGetLastError();
            return false;
        }
        c->ipc_opened.push_back( *out );
        return true;
    };
    if ( ok )
    {
        for ( int r = 0; r < W && ok; ++r )
        {
            if ( r == me )
                c->mail[r] = c->mail_self;
            else
                ok = open( all[r].h_mail, reinterpret_cast<void**>( &c->mail[r] ) );
        }
        // a neighbour may appear on several faces (periodic-free block grids: it does not); map once
        for ( int s = 0; s < 6 && ok; ++s )
        {
            const int r = c->nbr[s];
            if ( r < 0 )
                continue;
            if ( s < 2 )
                ok = open( all[r].h_xstage, reinterpret_cast<void**>( &c->peer_xstage[s] ) );
            ok = ok && open( all[r].h_r, reinterpret_cast<void**>( &c->peer_r[s] ) ) &&
                 open( all[r].h_p0, reinterpret_cast<void**>( &c->peer_p[0][s] ) ) &&
                 open( all[r].h_p1, reinterpret_cast<void**>( &c->peer_p[1][s] ) );
            c->peer_origin[s] = all[r].origin;
            c->peer_sy[s] = all[r].sy;
            c->peer_sz[s] = all[r].sz;
            for ( int d = 0; d < 3; ++d )
                c->peer_n[s][d] = all[r].n[d];
        }
    }
    // everybody or nobody: agree on the outcome (sum of failures over ranks)
    double flag = ok ? 0.0 : 1.0, *d_flag = nullptr;
    CFB_CUDA( c, cudaMalloc( &d_flag, sizeof( double ) ) );
    CFB_CUDA( c, cudaMemcpy( d_flag, &flag, sizeof( double ), cudaMemcpyHostToDevice ) );
    CFB_NCCL( c, g_nccl.AllReduce( d_flag, d_flag, 1, ncclDouble, ncclSum, cm->red, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    CFB_CUDA( c, cudaMemcpy( &flag, d_flag, sizeof( double ), cudaMemcpyDeviceToHost ) );
    cudaFree( d_flag );
    c->peer_ok = flag == 0.0;
    if ( c->peer_ok && !c->ev_ghost ) // events of the overlapped exchange (cfb_api.cu: enqueue_iteration)
    {
        CFB_CUDA( c, cudaEventCreateWithFlags( &c->ev_ghost, cudaEventDisableTiming ) );
        CFB_CUDA( c, cudaEventCreateWithFlags( &c->ev_phase[0], cudaEventDisableTiming ) );
        CFB_CUDA( c, cudaEventCreateWithFlags( &c->ev_phase[1], cudaEventDisableTiming ) );
        CFB_CUDA( c, cudaEventCreateWithFlags( &c->ev_bnd, cudaEventDisableTiming ) );
    }
    return CFB_OK;
}

void peer_destroy( cfb_ctx* c )
{
    for ( void* p : c->ipc_opened )
        cudaIpcCloseMemHandle( p );
    c->ipc_opened.clear();
    if ( c->mail_self )
        cudaFree( c->mail_self );
    if ( c->xstage_self )
        cudaFree( c->xstage_self );
    c->xstage_self = nullptr;
    if ( c->d_xticket )
        cudaFree( c->d_xticket );
    c->mail_self = nullptr;
    c->d_xticket = nullptr;
    c->peer_ok = false;
}

} // namespace

// staging slot of (side, kind) inside an xstage allocation; kind 0: r, 1: pbuf 0, 2: pbuf 1
static inline size_t xslot( const Geo& g, int side, int kind )
{
    return (size_t)( side * 3 + kind ) * g.n[1] * g.n[2];
}

int peer_exchange( cfb_ctx* c, int which, bool copy_r, int copy_pbuf, bool unpack )
{
    const Geo& g = c->g;
    XchgArgs a{};
    a.S = c->d_state;
    a.ticket = c->d_xticket;
    a.which = which;
    a.rank = c->cfg.world_rank;
    a.world = c->cfg.world_size;
    a.timeout_cycles = 20000000000ll; // ~10 s at 1.9 GHz: a dead peer must not hang the GPU
    for ( int r = 0; r < a.world; ++r )
        a.mail[r] = c->mail[r];
    long long cells = 0;
    XUnpackArgs u{};
    for ( int s = 0; s < 2 * g.D; ++s )
    {
        if ( c->nbr[s] < 0 )
            continue;
        const int d = s / 2, side = s % 2;
        for ( int fld = 0; fld < 2; ++fld )
        {
            if ( ( fld == 0 && !copy_r ) || ( fld == 1 && copy_pbuf < 0 ) )
                continue;
            const int kind = fld == 0 ? 0 : 1 + copy_pbuf;
            double* mine = fld == 0 ? c->cg_r : c->cg_pbuf[copy_pbuf];
            XFace& f = a.f[a.nface++];
            f.src = mine;
            f.dst = fld == 0 ? c->peer_r[s] : c->peer_p[copy_pbuf][s];
            f.dorigin = c->peer_origin[s];
            f.dsy = c->peer_sy[s];
            f.dsz = c->peer_sz[s];
            for ( int e = 0; e < 3; ++e )
            {
                f.lo[e] = 0;
                f.ext[e] = g.n[e];
                f.shift[e] = 0;
            }
            f.ext[d] = 1;
            if ( side == 0 )
            {
                f.lo[d] = 0; // my first layer -> the low neighbour's high ghost (index n_peer)
                f.shift[d] = -c->peer_n[s][d];
            }
            else
            {
                f.lo[d] = g.n[d] - 1; // my last layer -> the high neighbour's low ghost (index -1)
                f.shift[d] = g.n[d];
            }
            if ( d == 0 )
            {
                // x faces travel packed into the neighbour's staging slot [its side facing me]; what the
                // neighbour left in mine is scattered into my ghost column after the barrier.  (Storing them
                // straight into the neighbour's ghost column — 8-byte stores at the row stride over NVLink, no
                // staging, no scatter launch — was measured and is slower: 1071 vs 1151 iterations/s on an x split of
                // 2 x 512^3, 170 vs 44 us exposed after phase B; profiles/r2_bench_n2_blocks211_xdirect{0,1}.json.)
                f.packed = 1;
                f.dst = c->peer_xstage[s] + xslot( g, 1 - side, kind );
                if ( unpack )
                {
                    u.src[u.n] = c->xstage_self + xslot( g, side, kind );
                    u.dst[u.n] = mine;
                    u.col[u.n] = side == 0 ? -1 : g.n[0];
                    ++u.n;
                }
            }
            cells += (long long)f.ext[0] * f.ext[1] * f.ext[2];
        }
    }
    int grid = (int)std::min<long long>( std::max<long long>( ( cells + 1023 ) / 1024, 1 ), 2ll * c->sm_count );
    cg_xchg_kernel<<<grid, 256, 0, c->stream>>>( g, a );
    c->stats.kernel_launches += 1;
    if ( u.n > 0 )
    {
        const long long tot = (long long)g.n[1] * g.n[2] * u.n;
        const int ug = (int)std::min<long long>( ( tot + 255 ) / 256, 4ll * c->sm_count );
        cg_xunpack_kernel<<<ug, 256, 0, c->stream>>>( g, u );
        c->stats.kernel_launches += 1;
    }
    return CFB_OK;
}

int peer_faces_async( cfb_ctx* c, int kind, int pbuf, cudaEvent_t after )
{
    const Geo& g = c->g;
    FaceArgs a{};
    a.S = c->d_state;
    a.ticket = c->d_xticket + 1;
    a.kind = kind;
    a.mail_self = c->mail_self;
    a.timeout_cycles = 20000000000ll; // ~10 s at 1.9 GHz: a dead peer must not hang the GPU
    double* mine = kind == 0 ? c->cg_r : c->cg_pbuf[pbuf];
    const int stage_kind = kind == 0 ? 0 : 1 + pbuf;
    long long cells = 0;
    XUnpackArgs u{};
    for ( int s = 0; s < 2 * g.D; ++s )
    {
        if ( c->nbr[s] < 0 )
            continue;
        const int d = s / 2, side = s % 2;
        const int i = a.nface++;
        a.nbr_mail[i] = c->mail[c->nbr[s]];
        a.nbr_slot[i] = s ^ 1;
        a.my_slot[i] = s;
        XFace& f = a.f[i];
        f.src = mine;
        f.dst = kind == 0 ? c->peer_r[s] : c->peer_p[pbuf][s];
        f.dorigin = c->peer_origin[s];
        f.dsy = c->peer_sy[s];
        f.dsz = c->peer_sz[s];
        for ( int e = 0; e < 3; ++e )
        {
            f.lo[e] = 0;
            f.ext[e] = g.n[e];
            f.shift[e] = 0;
        }
        f.ext[d] = 1;
        if ( side == 0 )
        {
            f.lo[d] = 0; // my first layer -> the low neighbour's high ghost (index n_peer)
            f.shift[d] = -c->peer_n[s][d];
        }
        else
        {
            f.lo[d] = g.n[d] - 1; // my last layer -> the high neighbour's low ghost (index -1)
            f.shift[d] = g.n[d];
        }
        if ( d == 0 )
        {
            // x faces travel packed (see peer_exchange) and are scattered by the receiver behind the flag wait
            f.packed = 1;
            f.dst = c->peer_xstage[s] + xslot( g, 1 - side, stage_kind );
            u.src[u.n] = c->xstage_self + xslot( g, side, stage_kind );
            u.dst[u.n] = mine;
            u.col[u.n] = side == 0 ? -1 : g.n[0];
            ++u.n;
        }
        cells += (long long)f.ext[0] * f.ext[1] * f.ext[2];
    }
    if ( after )
        CFB_CUDA( c, cudaStreamWaitEvent( c->comm_stream, after, 0 ) );
    const int grid = (int)std::min<long long>( std::max<long long>( ( cells + 1023 ) / 1024, 1 ), 2ll * c->sm_count );
    cg_face_kernel<<<grid, 256, 0, c->comm_stream>>>( g, a );
    c->stats.kernel_launches += 1;
    if ( u.n > 0 )
    {
        const long long tot = (long long)g.n[1] * g.n[2] * u.n;
        const int ug = (int)std::min<long long>( ( tot + 255 ) / 256, 4ll * c->sm_count );
        cg_xunpack_kernel<<<ug, 256, 0, c->comm_stream>>>( g, u );
        c->stats.kernel_launches += 1;
    }
    CFB_CUDA( c, cudaEventRecord( c->ev_ghost, c->comm_stream ) );
    c->side_busy = true;
    return CFB_OK;
}

int peer_faces_join( cfb_ctx* c )
{
    if ( !c->side_busy )
        return CFB_OK;
    CFB_CUDA( c, cudaStreamWaitEvent( c->stream, c->ev_ghost, 0 ) );
    c->side_busy = false;
    return CFB_OK;
}

extern "C" int cfb_nccl_unique_id( unsigned char* id )
{
    std::string err;
    if ( !nccl_load( err ) )
        return cfb_fail( nullptr, CFB_ERR_NCCL, err );
    static_assert( sizeof( ncclUniqueId ) == CFB_NCCL_ID_BYTES, "ncclUniqueId size" );
    for ( int i = 0; i < 2; ++i )
    {
        ncclUniqueId u;
        ncclResult_t r = g_nccl.GetUniqueId( &u );
        if ( r != ncclSuccess )
            return cfb_fail( nullptr, CFB_ERR_NCCL, g_nccl.GetErrorString( r ) );
        std::memcpy( id + i * CFB_NCCL_ID_BYTES, &u, CFB_NCCL_ID_BYTES );
    }
    return CFB_OK;
}

int halo_init( cfb_ctx* c )
{
    std::string err;
    if ( !nccl_load( err ) )
        return cfb_fail( c, CFB_ERR_NCCL, err );
    const cfb_config& cfg = c->cfg;
    const Geo& g = c->g;
    int nranks = 1;
    for ( int d = 0; d < g.D; ++d )
        nranks *= cfg.ranks_per_dim[d];
    if ( nranks != cfg.world_size )
        return cfb_fail( c, CFB_ERR_INVALID, "ranks_per_dim does not multiply to world_size" );
    const int b[3] = { cfg.block_id[0], cfg.block_id[1], g.D == 3 ? cfg.block_id[2] : 0 };
    if ( rank_of( cfg, b[0], b[1], b[2] ) != cfg.world_rank )
        return cfb_fail( c, CFB_ERR_INVALID, "world_rank must equal (bz*py + by)*px + bx" );
    for ( int d = 0; d < g.D; ++d )
    {
        int lo[3] = { b[0], b[1], b[2] }, hi[3] = { b[0], b[1], b[2] };
        lo[d] -= 1;
        hi[d] += 1;
        c->nbr[2 * d] = b[d] > 0 ? rank_of( cfg, lo[0], lo[1], lo[2] ) : -1;
        c->nbr[2 * d + 1] = b[d] < cfg.ranks_per_dim[d] - 1 ? rank_of( cfg, hi[0], hi[1], hi[2] ) : -1;
    }
    Comm* cm = new Comm();
    c->nccl = cm;
    ncclUniqueId id0, id1;
    std::memcpy( &id0, cfg.nccl_id, CFB_NCCL_ID_BYTES );
    std::memcpy( &id1, cfg.nccl_id + CFB_NCCL_ID_BYTES, CFB_NCCL_ID_BYTES );
    CFB_NCCL( c, g_nccl.CommInitRank( &cm->halo, cfg.world_size, id0, cfg.world_rank ) );
    CFB_NCCL( c, g_nccl.CommInitRank( &cm->red, cfg.world_size, id1, cfg.world_rank ) );
    CFB_CUDA( c, cudaEventCreateWithFlags( &cm->ev_ready, cudaEventDisableTiming ) );
    CFB_CUDA( c, cudaEventCreateWithFlags( &cm->ev_done, cudaEventDisableTiming ) );
    // buffers: (D+1) fields x (halo+1) layers x the largest ghosted face
    long long e[3] = { g.n[0] + 1 + 2 * g.h, g.n[1] + 1 + 2 * g.h, g.n[2] + 1 + 2 * g.h };
    long long face = std::max( e[0] * e[1], std::max( e[0] * e[2], e[1] * e[2] ) );
    c->halo_buf_elems = (size_t)( face * ( g.h + 1 ) * ( g.D + 1 ) );
    for ( int s = 0; s < 2 * g.D; ++s )
    {
        if ( c->nbr[s] < 0 )
            continue;
        CFB_CUDA( c, cudaMalloc( &c->d_halo_send[s], c->halo_buf_elems * sizeof( double ) ) );
        CFB_CUDA( c, cudaMalloc( &c->d_halo_recv[s], c->halo_buf_elems * sizeof( double ) ) );
    }
    // warm both communicators up (connection setup happens on first use)
    CFB_NCCL( c, g_nccl.AllReduce( &c->d_state->pAp, &c->d_state->pAp, 1, ncclDouble, ncclSum, cm->red, c->stream ) );
    CFB_NCCL( c, g_nccl.AllReduce( &c->d_state->rr, &c->d_state->rr, 1, ncclDouble, ncclSum, cm->halo, c->comm_stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->comm_stream ) );
    CFB_CUDA( c, cudaMemsetAsync( c->d_state, 0, sizeof( CgState ), c->stream ) );
    if ( cfg.world_size > 64 )
        return cfb_fail( c, CFB_ERR_INVALID, "at most 64 ranks" );
    CFB_CUDA( c, cudaMemcpyAsync( &c->d_state->world, &cfg.world_size, sizeof( int ), cudaMemcpyHostToDevice, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    return peer_setup( c );
}

void halo_destroy( cfb_ctx* c )
{
    peer_destroy( c );
    Comm* cm = static_cast<Comm*>( c->nccl );
    if ( cm )
    {
        if ( cm->halo )
            g_nccl.CommDestroy( cm->halo );
        if ( cm->red )
            g_nccl.CommDestroy( cm->red );
        if ( cm->ev_ready )
            cudaEventDestroy( cm->ev_ready );
        if ( cm->ev_done )
            cudaEventDestroy( cm->ev_done );
        delete cm;
        c->nccl = nullptr;
    }
    for ( int s = 0; s < 6; ++s )
    {
        if ( c->d_halo_send[s] )
            cudaFree( c->d_halo_send[s] );
        if ( c->d_halo_recv[s] )
            cudaFree( c->d_halo_recv[s] );
        c->d_halo_send[s] = c->d_halo_recv[s] = nullptr;
    }
}

// Width-`w` face-neighbour exchange of one cell array (p of the CG, the pressure before
// _applyPressure), enqueued on the SIDE stream: halo_cells_begin() forks from the main stream,
// halo_cells_end() joins.  Between the two the caller may launch work that does not read ghosts.
int halo_cells_begin( cfb_ctx* c, double* const* fields, int nf, int w )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    const Geo& g = c->g;
    CFB_CUDA( c, cudaEventRecord( cm->ev_ready, c->stream ) );
    CFB_CUDA( c, cudaStreamWaitEvent( c->comm_stream, cm->ev_ready, 0 ) );
    int dims[3] = { 0, 1, 2 };
    Box slo[3], shi[3], rlo[3], rhi[3];
    for ( int d = 0; d < g.D; ++d )
    {
        Box full;
        for ( int e = 0; e < 3; ++e )
        {
            full.lo[e] = 0;
            full.hi[e] = g.n[e];
        }
        slo[d] = shi[d] = rlo[d] = rhi[d] = full;
        slo[d].lo[d] = 0;
        slo[d].hi[d] = w; // my first w layers -> low neighbour's high ghosts
        shi[d].lo[d] = g.n[d] - w;
        shi[d].hi[d] = g.n[d];
        rlo[d].lo[d] = -w;
        rlo[d].hi[d] = 0;
        rhi[d].lo[d] = g.n[d];
        rhi[d].hi[d] = g.n[d] + w;
    }
    double* fl[4] = { nullptr, nullptr, nullptr, nullptr };
    for ( int f = 0; f < nf && f < 4; ++f )
        fl[f] = fields[f];
    int rc = exchange( c, nf, fl, g.D, dims, slo, shi, rlo, rhi, c->comm_stream );
    if ( rc )
        return rc;
    CFB_CUDA( c, cudaEventRecord( cm->ev_done, c->comm_stream ) );
    return CFB_OK;
}

int halo_cells_end( cfb_ctx* c )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    CFB_CUDA( c, cudaStreamWaitEvent( c->stream, cm->ev_done, 0 ) );
    return CFB_OK;
}

int halo_exchange_cells( cfb_ctx* c, double* field, int width )
{
    int rc = halo_cells_begin( c, &field, 1, width );
    if ( rc )
        return rc;
    return halo_cells_end( c );
}

// ProblemManager::gather: width-h exchange of q,u,v(,w) including edges and corners, as three axis
// sweeps on the main stream.  Along the sweep axis d the low ghosts [-h, 0) come from the low
// neighbour's last h layers and the high ghosts [n, n+h] (h+1 layers: the face on the block
// boundary is owned by the upper block) from the high neighbour's first h+1 layers; tangentially the
// sweeps cover owned+1, then the x ghosts, then x and y ghosts.
int halo_exchange_fields( cfb_ctx* c, int version )
{
    const Geo& g = c->g;
    double* fl[4] = { nullptr, nullptr, nullptr, nullptr };
    for ( int f = 0; f <= g.D; ++f )
        fl[f] = field_ptr( c, f, version );
    for ( int d = 0; d < g.D; ++d )
    {
        Box t; // tangential extent of this sweep
        for ( int e = 0; e < 3; ++e )
        {
            if ( e >= g.D )
            {
                t.lo[e] = 0;
                t.hi[e] = 1;
            }
            else if ( e < d )
            {
                t.lo[e] = -g.h;
                t.hi[e] = g.n[e] + g.h + 1;
            }
            else
            {
                t.lo[e] = 0;
                t.hi[e] = g.n[e] + 1;
            }
        }
        Box slo = t, shi = t, rlo = t, rhi = t;
        slo.lo[d] = 0;
        slo.hi[d] = g.h + 1; // -> low neighbour's high ghosts [n, n+h]
        shi.lo[d] = g.n[d] - g.h;
        shi.hi[d] = g.n[d]; // -> high neighbour's low ghosts [-h, 0)
        rlo.lo[d] = -g.h;
        rlo.hi[d] = 0;
        rhi.lo[d] = g.n[d];
        rhi.hi[d] = g.n[d] + g.h + 1;
        int dims[1] = { d };
        int rc = exchange( c, g.D + 1, fl, 1, dims, &slo, &shi, &rlo, &rhi, c->stream );
        if ( rc )
            return rc;
    }
    return CFB_OK;
}

int peer_map_arrays( cfb_ctx* c, int count, double* const* mine, std::vector<double*>& mapped, bool* ok_out )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    const int W = c->cfg.world_size, me = c->cfg.world_rank;
    *ok_out = false;
    mapped.assign( (size_t)6 * count, nullptr );
    if ( !cm || count < 1 )
        return CFB_OK;
    // [flag][count handles] per rank, all-gathered as bytes
    const size_t rec = sizeof( cudaIpcMemHandle_t ) * count + 64;
    std::vector<char> mine_rec( rec, 0 );
    int good = c->peer_ok ? 1 : 0; // no peer mailboxes, no peer exchanges
    for ( int a = 0; a < count && good; ++a )
        if ( cudaIpcGetMemHandle( reinterpret_cast<cudaIpcMemHandle_t*>( mine_rec.data() + 64 ) + a, mine[a] ) != cudaSuccess )
        {
            cudaGetLastError();
            good = 0;
        }
    std::memcpy( mine_rec.data(), &good, sizeof( int ) );
    char* d_all = nullptr;
    CFB_CUDA( c, cudaMalloc( &d_all, rec * ( W + 1 ) ) );
    CFB_CUDA( c, cudaMemcpy( d_all + rec * W, mine_rec.data(), rec, cudaMemcpyHostToDevice ) );
    CFB_NCCL( c, g_nccl.AllGather( d_all + rec * W, d_all, rec, ncclChar, cm->red, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    std::vector<char> all( rec * W );
    CFB_CUDA( c, cudaMemcpy( all.data(), d_all, rec * W, cudaMemcpyDeviceToHost ) );
    cudaFree( d_all );
    bool ok = true;
    for ( int r = 0; r < W; ++r )
    {
        int g = 0;
        std::memcpy( &g, all.data() + rec * r, sizeof( int ) );
        ok = ok && g != 0;
    }
    if ( ok )
        for ( int s = 0; s < 6 && ok; ++s )
        {
            const int r = c->nbr[s];
            if ( r < 0 || r == me )
                continue;
            const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>( all.data() + rec * r + 64 );
            for ( int a = 0; a < count && ok; ++a )
            {
                void* p = nullptr;
                if ( cudaIpcOpenMemHandle( &p, h[a], cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess )
                {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                c->ipc_opened.push_back( p );
                mapped[(size_t)s * count + a] = static_cast<double*>( p );
            }
        }
    // everybody or nobody
    double flag = ok ? 0.0 : 1.0, *d_flag = nullptr;
    CFB_CUDA( c, cudaMalloc( &d_flag, sizeof( double ) ) );
    CFB_CUDA( c, cudaMemcpy( d_flag, &flag, sizeof( double ), cudaMemcpyHostToDevice ) );
    CFB_NCCL( c, g_nccl.AllReduce( d_flag, d_flag, 1, ncclDouble, ncclSum, cm->red, c->stream ) );
    CFB_CUDA( c, cudaStreamSynchronize( c->stream ) );
    CFB_CUDA( c, cudaMemcpy( &flag, d_flag, sizeof( double ), cudaMemcpyDeviceToHost ) );
    cudaFree( d_flag );
    *ok_out = flag == 0.0;
    if ( !*ok_out )
        mapped.assign( (size_t)6 * count, nullptr );
    return CFB_OK;
}

// One grouped exchange of the per-neighbour buffers: d_halo_send[s] (counts[s] doubles) goes to the neighbour on
// face s, d_halo_recv[s] is filled by it; used by callers that pack / unpack with their own kernels (the
// multigrid levels of mg.cu, whose arrays do not have the layout of the CG vectors).
int halo_sendrecv_slots( cfb_ctx* c, const size_t counts[6], cudaStream_t st )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    CFB_NCCL( c, g_nccl.GroupStart() );
    for ( int s = 0; s < 6; ++s )
    {
        if ( c->nbr[s] < 0 || counts[s] == 0 )
            continue;
        if ( counts[s] > c->halo_buf_elems )
            return cfb_fail( c, CFB_ERR_INVALID, "halo buffer too small" );
        CFB_NCCL( c, g_nccl.Send( c->d_halo_send[s], counts[s], ncclDouble, c->nbr[s], cm->halo, st ) );
        CFB_NCCL( c, g_nccl.Recv( c->d_halo_recv[s], counts[s], ncclDouble, c->nbr[s], cm->halo, st ) );
    }
    CFB_NCCL( c, g_nccl.GroupEnd() );
    return CFB_OK;
}

int halo_allreduce( cfb_ctx* c, double* dev_vals, int n )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    CFB_NCCL( c, g_nccl.AllReduce( dev_vals, dev_vals, (size_t)n, ncclDouble, ncclSum, cm->red, c->stream ) );
    return CFB_OK;
}

int halo_allgather( cfb_ctx* c, const double* dev_send, double* dev_recv, int n_per_rank )
{
    Comm* cm = static_cast<Comm*>( c->nccl );
    CFB_NCCL( c, g_nccl.AllGather( dev_send, dev_recv, (size_t)n_per_rank, ncclDouble, cm->red, c->stream ) );
    return CFB_OK;
}
