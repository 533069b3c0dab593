// device_cg1.cuh — the scalar step of the single-reduction CG (cg_variant 3; kernels_cg1.cu, kernels_stencil.cu
// MODE 2), run by ONE thread once the three global sums of an iteration are known: on one GPU and on the NVLink
// peer path the last block of the stencil kernel, over NCCL the combine kernel behind the all-gather.
#pragma once
#include "cfb_internal.h"

// rr = r.r, gamma = r.M^-1 r, delta = (A M^-1 r).(M^-1 r) of the residual the update kernel just produced
// (init: of r0 = b).  Same statements, same order as the checker's cg_solve_single_reduction.
__device__ __forceinline__ void cg1_finish( CgState* S, double rr, double gamma, double delta, int init )
{
    S->rr = rr;
    if ( init )
    {
        // (the threshold and an r0 that already meets it: cg_check0_kernel, before this launch)
        S->beta = 0.0;
        S->alpha = gamma / delta;
        S->rz_old = gamma;
        return;
    }
    const double resid = sqrt( rr );
    const int it = S->iter;
    if ( it < CFB_HIST_MAX )
        S->hist[it] = resid;
    S->iter = it + 1;
    if ( !S->fixed && resid <= S->thresh )
    {
        S->done = 1;
        return;
    }
    const double beta = gamma / S->rz_old;
    const double alpha = gamma / ( delta - ( beta * gamma ) / S->alpha );
    S->beta = beta;
    S->alpha = alpha;
    S->rz_old = gamma;
}
