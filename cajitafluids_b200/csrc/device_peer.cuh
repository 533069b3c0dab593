// device_peer.cuh — the global sums of a CG reduction point taken by the compute kernel itself: the block that draws
// the kernel's last ticket (the one that already closes the local reduction) publishes the local double-doubles into
// every rank's mailbox over NVLink peer memory, waits for everybody's and combines them in rank order — steps 3-5 of
// the exchange kernel (halo.cu: cg_xchg_kernel), same mailboxes, same sequence numbers.  Used by the overlapped
// exchange schedule ("peer_overlap": the faces travel on the side stream, halo.cu: cg_face_kernel) and by the
// single-reduction CG (kernels_stencil.cu MODE 2).  Template flag PF of the kernels; the other instantiations are
// untouched.
//
// Round 1 also had the boundary tiles store their face cells into the neighbours' ghost layers from inside the
// compute kernels ("peer_fused").  It measured 0.57 - 0.75 x the exchange-kernel schedule on 2 - 8 GPUs (per-cell
// 8-byte system-scope stores, a system fence in every boundary block: SCALE_r01.json) and was removed in round 2.
#pragma once
#include "cfb_internal.h"
#include "device_reduce.cuh"

namespace
{

struct PeerFusedArgs
{
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

// the mailboxes of every rank, for a kernel whose last block runs the reduction
inline void peer_mail_only( cfb_ctx* c, PeerFusedArgs& pf )
{
    pf.rank = c->cfg.world_rank;
    pf.world = c->cfg.world_size;
    pf.timeout_cycles = 20000000000ll; // ~10 s at 1.9 GHz: a dead peer must not hang the GPU
    for ( int r = 0; r < pf.world; ++r )
        pf.mail[r] = c->mail[r];
}


} // namespace
