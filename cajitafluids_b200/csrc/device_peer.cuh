// device_peer.cuh — the ghost / reduction exchange of a CG reduction point done by the compute kernel itself
// ("peer_fused" tuning key; several blocks over NVLink peer memory).  Shared by phase A (cg_rupdate_kernel<true>,
// stencil7_dot_tma<C, 1, false, true>) and phase B (cg_fused_kernel<C, false, false, true>).
#pragma once
#include "cfb_internal.h"
#include "device_reduce.cuh"

namespace
{

// Several blocks over NVLink peer memory.  Instead of phase B followed by the exchange kernel (halo.cu:
// cg_xchg_kernel), the boundary tiles store the cells of the new search direction that lie on a block face
// straight into the neighbour's ghost layer as they are computed — the transfer overlaps the z-march tile by
// tile — and the block that draws the last ticket (the one that already closes the p.Ap reduction) publishes
// the local double-double into every rank's mailbox, waits for everybody's and combines them: steps 3-5 of the
// exchange kernel, same mailboxes, same sequence numbers, hence the same barrier semantics (a rank leaves phase
// B only when every rank's ghost stores of this phase are done; p is double-buffered, so nobody still reads the
// ghost layers written here).  Phase A (cg_rupdate_kernel<true>) does the same with the faces of r and the
// (r.z, r.r) pair; r is single-buffered, and safe for the same reason: a rank enters phase A only when every
// rank has left phase B, the last reader of the old ghost layers of r.  No exchange launch in the iteration.
struct PeerFace
{
    double* dst;                 // the neighbour's copy of the array the new p is written to
    long long dorigin, dsy, dsz; // its layout
    int lo[3], ext[3], shift[3]; // my box (owned index space); peer index = my index - shift
};

struct PeerFusedArgs
{
    int nface;
    PeerFace f[6];
    PeerMail* mail[CFB_MAX_PEERS];
    int rank, world;
    long long timeout_cycles;
};

struct NoPeerArgs
{
};

template <bool PF>
struct PeerSel
{
    typedef NoPeerArgs type;
};
template <>
struct PeerSel<true>
{
    typedef PeerFusedArgs type;
};

__device__ __forceinline__ void peer_store_cell( const PeerFusedArgs& pf, unsigned fmask, int i, int j, int k, double v )
{
#pragma unroll
    for ( int f = 0; f < 6; ++f )
    {
        if ( !( ( fmask >> f ) & 1u ) )
            continue;
        const PeerFace& F = pf.f[f];
        if ( i >= F.lo[0] && i < F.lo[0] + F.ext[0] && j >= F.lo[1] && j < F.lo[1] + F.ext[1] && k >= F.lo[2] &&
             k < F.lo[2] + F.ext[2] )
            F.dst[F.dorigin + (long long)( k - F.shift[2] ) * F.dsz + (long long)( j - F.shift[1] ) * F.dsy +
                  ( i - F.shift[0] )] = v;
    }
}

// Steps 3-5 of cg_xchg_kernel (halo.cu), run by all threads of the block that drew the last ticket of a kernel:
// publish S->loc[first .. first + nd) into slot [which][my rank] of every rank's mailbox (data, system fence,
// sequence number), wait for every rank's publication in mine (bounded), and hand back the exact sums in rank
// order: out[v] = sum over ranks of the v-th double-double (thread 0 only).
template <int NV>
__device__ __forceinline__ void peer_mail_exchange( CgState* S, const PeerFusedArgs& pf, int which, int first, dd_t out[NV] )
{
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    __threadfence_system();
    if ( tid == 0 )
        s_seq = ++S->seq[which];
    __syncthreads(); // also: what thread 0 left in S->loc is visible to the publishing threads
    const unsigned long long seq = s_seq;
    if ( tid < pf.world )
    {
        PeerMail* m = pf.mail[tid];
        for ( int q = 0; q < 2 * NV; ++q )
            m->v[which][pf.rank][q] = S->loc[first + q];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>( &m->seq[which][pf.rank] ) = seq;
    }
    PeerMail* me = pf.mail[pf.rank];
    if ( tid < pf.world && !S->xerror )
    {
        const long long t0 = clock64();
        while ( *reinterpret_cast<const volatile unsigned long long*>( &me->seq[which][tid] ) < seq )
            if ( clock64() - t0 > pf.timeout_cycles )
            {
                S->xerror = 1;
                break;
            }
    }
    __threadfence_system();
    __syncthreads();
    if ( tid == 0 )
    {
        for ( int v = 0; v < NV; ++v )
        {
            dd_t sum = { 0.0, 0.0 };
            for ( int r = 0; r < pf.world; ++r )
            {
                const volatile double* src = &me->v[which][r][2 * v];
                dd_t w = { src[0], src[1] };
                sum = dd_add( sum, w );
            }
            out[v] = sum;
        }
    }
}

// no faces: the kernel only runs the mailbox reduction in its last block ("peer_overlap": the faces travel on the
// side stream, halo.cu: cg_face_kernel)
inline void peer_mail_only( cfb_ctx* c, PeerFusedArgs& pf )
{
    pf.nface = 0;
    pf.rank = c->cfg.world_rank;
    pf.world = c->cfg.world_size;
    pf.timeout_cycles = 20000000000ll; // ~10 s at 1.9 GHz: a dead peer must not hang the GPU
    for ( int r = 0; r < pf.world; ++r )
        pf.mail[r] = c->mail[r];
}

// the faces of block `c` towards its neighbours, as destinations of `array` (one of the neighbours' mapped copies)
inline void peer_faces( cfb_ctx* c, PeerFusedArgs& pf, double* const dst_of_side[6] )
{
    const Geo& g = c->g;
    pf.rank = c->cfg.world_rank;
    pf.world = c->cfg.world_size;
    pf.timeout_cycles = 20000000000ll; // ~10 s at 1.9 GHz: a dead peer must not hang the GPU
    for ( int r = 0; r < pf.world; ++r )
        pf.mail[r] = c->mail[r];
    for ( int s = 0; s < 2 * g.D; ++s )
    {
        if ( c->nbr[s] < 0 )
            continue;
        const int d = s / 2, side = s % 2;
        PeerFace& f = pf.f[pf.nface++];
        f.dst = dst_of_side[s];
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
    }
}


} // namespace
