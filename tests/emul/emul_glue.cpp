// emul_glue.cpp — tests/emul only: the emulated launcher, and host stand-ins for the parts of the
// library that cannot be emulated (TMA / mbarrier kernels, NCCL / peer-memory exchange).
//   * q = A p, sum p.q  (the TMA stencil kernel, verified on the GPU)  -> a plain loop with the same row
//     arithmetic (apply_row) and the same exactly-accumulated dot product
//   * the two-kernel fused CG form  -> not available: the emulated context runs the three-kernel form
//   * multi-GPU: halo.cu itself is compiled; NCCL is the in-process stand-in of nccl_emul.cpp (ranks are
//     threads), the NVLink peer-memory path reports "not available" (cudaIpc stand-ins fail)
#include "cfb_internal.h"
#include "device_geo.cuh"
#include "device_reduce.cuh"

namespace cfb_emul
{
thread_local uint3 g_threadIdx{ 0, 0, 0 }, g_blockIdx{ 0, 0, 0 };
thread_local dim3 g_blockDim, g_gridDim;

void launch( dim3 grid, dim3 block, const std::function<void()>& body )
{
    g_gridDim = grid;
    g_blockDim = block;
    for ( unsigned bz = 0; bz < grid.z; ++bz )
        for ( unsigned by = 0; by < grid.y; ++by )
            for ( unsigned bx = 0; bx < grid.x; ++bx )
            {
                g_blockIdx = uint3{ bx, by, bz };
                for ( unsigned tz = block.z; tz-- > 0; )
                    for ( unsigned ty = block.y; ty-- > 0; )
                        for ( unsigned tx = block.x; tx-- > 0; )
                        {
                            g_threadIdx = uint3{ tx, ty, tz };
                            body();
                        }
            }
}
} // namespace cfb_emul

int stencil_setup( cfb_ctx* c )
{
    c->tmap_ok = true;
    return CFB_OK;
}

int launch_stencil_dot( cfb_ctx* c )
{
    const Geo& g = c->g;
    const OpConst& op = c->op;
    CgState* S = c->d_state;
    if ( S->done )
        return 1;
    const double* p = c->cg_p;
    double* q = c->cg_q;
    dd_t acc{ 0.0, 0.0 };
    for ( int k = 0; k < g.n[2]; ++k )
        for ( int j = 0; j < g.n[1]; ++j )
            for ( int i = 0; i < g.n[0]; ++i )
            {
                const long long o = geo_off( g, i, j, k );
                const int w = wall_count( g, 0, i + g.off[0] ) + wall_count( g, 1, j + g.off[1] ) +
                              wall_count( g, 2, k + g.off[2] );
                const double a = apply_row( op.diag[w], op.neg_scale, p[o], p[o - 1], p[o + 1], p[o - g.sy],
                                            p[o + g.sy], p[o - g.sz], p[o + g.sz] );
                q[o] = a;
                dd_acc( acc, p[o] * a );
            }
    // publish_pAp of kernels_stencil.cu: several blocks keep the local double-double for the exact combine
    if ( S->world > 1 )
    {
        S->loc[0] = acc.hi;
        S->loc[1] = acc.lo;
    }
    else
        S->pAp = acc.hi + acc.lo;
    S->rz_old = S->rz_new;
    return 1;
}

int fused_setup( cfb_ctx* c )
{
    c->cg_variant = 0; // three-kernel CG form: the fused TMA kernels cannot be emulated
    c->fused_ok = false;
    return CFB_OK;
}
int launch_cg_rupdate( cfb_ctx* c ) { return cfb_fail( c, CFB_ERR_INVALID, "emul: fused CG form unavailable" ), 0; }
int launch_cg_fused( cfb_ctx* c, int ) { return cfb_fail( c, CFB_ERR_INVALID, "emul: fused CG form unavailable" ), 0; }
int launch_cg_finish( cfb_ctx* ) { return 0; }
