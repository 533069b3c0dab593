// cfb_internal.h — shared declarations of the sm_100a implementation behind include/cfb.h.
#pragma once

#include "../../include/cfb.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
// Geometry of one rank's block, passed by value to every kernel.
//
// Device layout (all fields of a ctx share it):  offset(i,j,k) = origin + k*sz + j*sy + i
// with (i,j,k) the 0-based OWNED index.  hx = 16 doubles (first owned entity of every row is
// 128-byte aligned), hy = hz = halo width; rows/planes are allocated for owned + 1 (faces) +
// 2*halo, so every reference ghost index (local index l <-> owned index l - halo) exists.
// In 2-D nz = 1 and the z ghost planes stay zero, which makes the 7-point operator with SOLID
// z walls identical to the reference's 5-point operator (diag 6-2 = 4).
struct Geo
{
    int D;
    int n[3];     // owned cells
    int nf[3];    // owned faces along their normal: n[d] + (block touches the high wall)
    int off[3];   // global offset of owned cell 0
    int gn[3];    // global cells
    int h;        // halo cell width
    int lo_bd[3]; // block touches the low / high physical wall
    int hi_bd[3];
    int bt[6];    // boundary types, [d] low, [3 + d] high  (NB: always stride 3 here)
    long long sy, sz, origin, total; // strides / allocation size in doubles
    int ay, az;                      // allocated rows / planes
    double cell, dt; // Mesh::cellSize() == cell size of dim 0 (src/Mesh.hpp:119-122): operator scales, dt clamp
    // Cajita's UniformGlobalMesh keeps one cell size per dimension, (hi_d - lo_d) / n_d; LocalMesh::coordinates
    // and the spline logical coordinates use it (the Mesh ctor only checks agreement to 10 eps, :56-64)
    double celld[3], rdxd[3];
    double ghost_low[3]; // Cajita LocalMesh ghosted low corner of this block
    double time;
};

// Pressure operator constants (src/VelocityCorrector.hpp:128,137 + BoundaryConditions.hpp:56-97).
// diag[c] = 2*D*scale with `scale` subtracted c times (c = number of SOLID walls the cell touches),
// minv[c] = 1.0 / diag[c]  (src/VelocityCorrector.hpp:178).
struct OpConst
{
    double scale;     // dt / (rho h^2)
    double neg_scale; // off-diagonal coefficient
    double diag[8];
    double minv[8];
};

// Device-resident CG scalars (one per ctx).  All kernels read/write it; the host polls it.
#define CFB_HIST_MAX 8192
struct CgState
{
    // global values (after the allreduce when multi-GPU); (rz_new, rr) are adjacent on purpose:
    // they travel in one 2-element allreduce, pAp in a 1-element one.
    double rz_old, pAp, rz_new, rr;
    double loc[6];                  // multi-GPU: local double-double sums  pAp | rz_new | rr
    double gath[64 * 6];            // multi-GPU: all-gathered local sums (<= 64 ranks, <= 3 double-doubles each)
    int world;                      // number of ranks
    int pad0;
    double bnorm;                   // sqrt(rr) of r0
    double thresh;                  // absolute threshold in use
    double alpha;                   // two-kernel form: alpha of the running iteration (phase A -> B)
    double beta;                    // single-reduction form (cg_variant 3): beta of the running iteration
    unsigned long long seq[2];      // peer-memory exchange: publications so far of (pAp | rz_new, rr)
    unsigned long long fseq[2];     // overlapped exchange: face transfers completed so far of (r | the search direction)
    int xerror;                     // sticky: a peer never published (timeout)
    int pad1;
    int iter;                       // completed iterations (kernel-1 executions)
    int done;                       // 1 once sqrt(rr) <= thresh
    int fixed;                      // fixed-iteration mode: never set done
    int pad;
    unsigned int ticket[4];         // last-block tickets: [0] init/axpy, [1] stencil
    unsigned int gbar[2];           // persistent kernel: grid barrier (arrivals of the running barrier | generation)
    double hist[CFB_HIST_MAX];
};

struct InflowConst
{
    double lo[3], hi[3], vel[3], quantity, force_dt[3];
};

// Mailbox of one rank for the peer-memory reductions: slot [which][writer rank], written by the
// writer over NVLink (data, system fence, then the sequence number), read locally.
#define CFB_MAX_PEERS 16
struct PeerMail
{
    double v[2][CFB_MAX_PEERS][6]; // <= 3 double-doubles per publication
    unsigned long long seq[2][CFB_MAX_PEERS];
    // overlapped exchange (halo.cu: cg_face_kernel): slot [0: r, 1: search direction][my face the writer sits on] =
    // number of face transfers of that kind the neighbour has completed into my ghost layers / staging areas
    unsigned long long fseq[2][6];
    // multigrid ghost exchanges (mg.cu): slot [writer rank] = number of exchanges whose stores that rank has
    // completed into my arrays
    unsigned long long mseq[CFB_MAX_PEERS];
};

// Minimum capacity (entries per value) of the block-partial scratch; the grid-stride kernels never launch more
// blocks than this.  The tiled kernels (one block per unit) grow the scratch to their unit count: ensure_partials.
#define CFB_MAX_PARTIALS 4096
#define CFB_MAX_UNITS ( 1 << 22 ) // more tiles than this in one launch is refused (CFB_ERR_INVALID)
#define CFB_KTIMED 64

struct OutputStage; // output.cu
struct MgStage;     // mg.cu

struct cfb_ctx
{
    cfb_config cfg{};
    Geo g{};
    OpConst op{};
    InflowConst inflow{};
    std::string err;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, comm_stream = nullptr;

    // fields: [id][version]; only q,u,v,w have two versions
    double* fld[4][2] = { { nullptr } };
    int cur[4] = { 0, 0, 0, 0 };
    double *lhs = nullptr, *rhs = nullptr, *cg_r = nullptr, *cg_p = nullptr, *cg_q = nullptr;
    // the search direction is double-buffered: the two-kernel iteration recomputes the new p on tile
    // halos from the OLD p of neighbouring tiles, so it cannot be updated in place.  cg_p (and
    // tmap_p) always name the current buffer cg_pbuf[pcur].
    double* cg_pbuf[2] = { nullptr, nullptr };
    int pcur = 0;

    CgState* d_state = nullptr;
    CgState* h_state = nullptr; // pinned mirror (first bytes only are copied)
    double* d_partials = nullptr; // [3 values][partials_cap][hi,lo] scratch for block partial sums
    int partials_cap = 0;         // >= CFB_MAX_PARTIALS; value n of block b at [( n * stride + b ) * 2], stride <= cap
    // first error raised while enqueuing work (a launcher refusing a configuration, a failed exchange call): the
    // launchers return launch counts, so the code travels here and pcg_solve / the C entry points hand it out
    int sticky_rc = 0;

    // stencil TMA descriptor + tiling
    CUtensorMap tmap_p{};
    CUtensorMap tmap_pbuf[2] = {}; // stencil-box maps of cg_pbuf[0 / 1]
    CUtensorMap tmap_sr{};         // stencil-box map of cg_r (single-reduction form: the stencil runs on M^-1 r)
    CUtensorMap tmap_r1{};         // halo-free tile map of cg_r (phase A' of the 64-byte form with r staged by TMA)
    bool tmap_ok = false;
    int st_variant = 0; // 0 = TMA z-march (default)
    int st_tx = 64, st_ty = 16, st_stages = 4, st_zc = 64;
    bool st_zc_auto = true; // phase A' picks its own z chunk (launch_stencil) until "stencil_zc" is set
    int poll_every = 0; // 0 = auto
    // "flat_2d" tuning key: two-dimensional runs skip the loads of the two zero ghost planes in the TMA kernels
    // (FLAT instantiations).  On for every 2-D context since it was measured (8192^2, 50 fixed iterations:
    // 658 vs 581 iterations/s, profiles/r2_bench_n1.json extra.flat_2d_probe); same values bit for bit.
    bool flat_2d = false; // set in cfb_create
    // "advect_tile" tuning key: 32 x 2 x 2 entity tiles per block in the advection kernel instead of rows
    bool advect_tile = false;
    // "peer_overlap" tuning key (several blocks, NVLink peer memory, two-kernel form): the reduction of each phase
    // runs in the last block of the compute kernel (mailboxes), the faces travel on the side stream under the
    // interior units of phase B (r) and under the next phase A (search direction); boundary units run last
    bool peer_overlap = false;
    cudaEvent_t ev_phase[2] = { nullptr, nullptr }; // main stream: phase A / phase B of the running iteration done
    cudaEvent_t ev_ghost = nullptr;                 // side stream: everything enqueued there so far is done
    cudaEvent_t ev_bnd = nullptr;                   // side stream: the boundary units of the running phase B are done
    bool side_busy = false;                         // face transfers enqueued since the last join

    // two-kernel CG iteration (kernels_fused.cu): tensor maps of cg_r / cg_p, tiling, unit list
    // 1 = two kernels / 72 B per cell, 0 = three kernels / 88 B, 2 = two kernels / 64 B: q is
    // never stored, phase A' recomputes A p (kernels_stencil.cu MODE 1); 3 = single-reduction (Chronopoulos-Gear)
    // form, opt-in: two kernels / 88 B per cell, ONE reduction point and one ghost exchange per iteration
    // (kernels_cg1.cu; iteration counts within +-1 of the other forms, which are bit-identical to one another)
    int cg_variant = -1; // -1: chosen per solve by cg_variant_auto() below
    CUtensorMap tmap_fr{}, tmap_fp[2] = {}; // fused-box maps of cg_r and of cg_pbuf[0 / 1]
    bool fused_ok = false;
    bool fu_auto = true; // pick the tiling from the block size; any "fused_*" tuning key turns it off
    int fu_tx = 64, fu_ty = 16, fu_stages = 3, fu_zc = 64;
    // two-dimensional runs (FLAT kernels): tile rows a unit of phase B marches through along y ("fused_yc"; picked by
    // fused_setup until the key is set)
    int fu_yc = 1;
    bool fu_yc_auto = true;
    int fu_nt = 0; // "fused_nt": threads per CTA of phase B: 256, 512 (128 x 16 x 3 tiling only), 0 = dispatch_fused picks
    int* d_units = nullptr; // (tile_x, tile_y, chunk) triples: interior units first, then boundary
    int n_units = 0, n_interior = 0;
    // "cg_persist" tuning key: 1 = run batches of iterations of the two-kernel form in ONE cooperative launch
    // (kernels_fused.cu: cg_persistent_kernel), 0 = never, -1 = where it pays: one block of ~80^3 ... ~170^3 cells
    // (the launch-per-phase kernels are latency-bound there; cg_persist_applies below)
    int cg_persist = -1;
    bool fu_reverse = false;  // phase B walks the units top-down (L2 reuse between the phases)
    int ru_ctas = 3;          // phase A: CTAs per SM of the grid-stride launch
    // NCCL path: run interior units while the r/p ghosts are in flight (measured slower than
    // halo-first at 2 x 512^3: profiles/r1_bench_n2_nccl_overlap.json vs r1_bench_n2_nccl.json; off)
    bool overlap_halo = false;

    // stats
    cudaEvent_t ev[16] = { nullptr };
    bool ev_pending[8] = { false };
    bool time_kernels = false;
    int ktimed = 0;
    cudaEvent_t kev[CFB_KTIMED][6] = { { nullptr } }; // A start, A end, B start, B end | A kernel end, B kernel end
    cfb_stats stats{};
    int last_iters = 0;
    double last_resid = 0;
    std::vector<double> hist;

    // multi-GPU
    void* nccl = nullptr; // NcclComm*, see halo.cu
    double* d_halo_send[6] = { nullptr };
    double* d_halo_recv[6] = { nullptr };
    size_t halo_buf_elems = 0;
    int nbr[6] = { -1, -1, -1, -1, -1, -1 }; // neighbour ranks: [2*d] low, [2*d+1] high

    // NVLink peer memory (cudaIpc): the neighbours' cg_r / cg_pbuf arrays and every rank's mailbox are
    // mapped into this process, so the per-iteration ghost exchange and the CG reductions are plain
    // stores into peer HBM from one kernel (halo.cu: cg_xchg_kernel) — no pack/unpack, no NCCL call.
    bool peer_ok = false;  // mappings established
    bool use_peer = true;  // use them for the CG iterations ("peer_halo" tuning key)
    double* peer_r[6] = { nullptr };
    double* peer_p[2][6] = { { nullptr } };
    long long peer_origin[6] = { 0 }, peer_sy[6] = { 0 }, peer_sz[6] = { 0 };
    int peer_n[6][3] = { { 0 } };
    // x faces are strided in memory: they travel packed (contiguous NVLink stores) into the neighbour's
    // staging area [side][0: r, 1: pbuf 0, 2: pbuf 1][nz * ny] and are scattered into the ghost column
    // by the receiver (an experiment that let phases A / B store their faces themselves and read the x
    // ghosts from staging was bit-exact but ~7 % slower: branch exp/peer-inkernel)
    double* xstage_self = nullptr;
    // "peer_xstage" tuning key; measured slower than scattering into the ghost columns (x split of 2 x
    // 512^3: 539 vs 581 it/s): the halo warp waits for the global loads plane by plane
    bool peer_xstage_reads = false;
    double* peer_xstage[2] = { nullptr, nullptr };
    PeerMail* mail_self = nullptr;
    PeerMail* mail[CFB_MAX_PEERS] = { nullptr };
    unsigned int* d_xticket = nullptr;
    std::vector<void*> ipc_opened;

    // output stage (output.cu), created on first use
    OutputStage* out = nullptr;

    // opt-in multigrid preconditioner (mg.cu; cfb_set_preconditioner)
    int precond = CFB_PRECOND_JACOBI;
    int mg_max_levels = 0; // 0 = as many as the block allows
    // "mg_tma" tuning key: the fine-level smoothing sweeps of three-dimensional runs on the TMA z-march
    // (kernels_stencil.cu MODE 3 / 4) instead of the one-thread-per-cell kernels
    // -1 (default): on for one block — measured and validated on a B200 — and off for several blocks, where the march
    // has run in the host emulation only (every block grid, bit-identical) but not yet on several GPUs; 1 forces it on
    int mg_tma = -1;
    CUtensorMap mg_map_box[3] = {}, mg_map_tile_b{}; // stencil boxes of the fine level's b, x[0], x[1]; tile of b
    double* mg_map_ptr[3] = { nullptr, nullptr, nullptr };
    // "mg_tma_prolong" tuning key: prolongation + first post-sweep on the march too (MODE 5); maps of the coarse
    // level's two iterate buffers
    bool mg_tma_prolong = true;
    CUtensorMap mg_map_e[2] = {};
    double* mg_map_e_ptr[2] = { nullptr, nullptr };
    bool mg_graph = false; // "mg_graph" tuning key: replay the V-cycle as a CUDA graph (one block)
    bool mg_coarse = false; // "mg_coarse_kernel" tuning key: the coarse end of the cycle in one single-CTA kernel
    MgStage* mg = nullptr;
};

extern std::string g_cfb_error;

int cfb_fail( cfb_ctx* c, int code, const std::string& msg );
// remember the first non-zero status of a call whose return value cannot travel (see cfb_ctx::sticky_rc)
inline int note_rc( cfb_ctx* c, int rc )
{
    if ( rc && !c->sticky_rc )
        c->sticky_rc = rc;
    return rc;
}
inline int take_sticky_rc( cfb_ctx* c )
{
    const int rc = c->sticky_rc;
    c->sticky_rc = 0;
    return rc;
}
// make room for `units` block partials per value (host-synchronising when it has to grow; set-up time only)
int ensure_partials( cfb_ctx* c, long long units );

#define CFB_CUDA( c, expr )                                                                        \
    do                                                                                             \
    {                                                                                              \
        cudaError_t _e = ( expr );                                                                 \
        if ( _e != cudaSuccess )                                                                   \
            return cfb_fail( ( c ), CFB_ERR_CUDA,                                                  \
                             std::string( #expr ) + ": " + cudaGetErrorString( _e ) + " at " +     \
                                 __FILE__ + ":" + std::to_string( __LINE__ ) );                    \
    } while ( 0 )

// make cg_pbuf[which] the current search-direction buffer
inline void cg_select_p( cfb_ctx* c, int which )
{
    c->pcur = which & 1;
    c->cg_p = c->cg_pbuf[c->pcur];
    c->tmap_p = c->tmap_pbuf[c->pcur];
}

inline double* field_ptr( cfb_ctx* c, int field, int version )
{
    if ( field >= CFB_QUANTITY && field <= CFB_W )
    {
        if ( field > c->g.D )
            return nullptr;
        int v = version == CFB_CURRENT ? c->cur[field] : 1 - c->cur[field];
        return c->fld[field][v];
    }
    switch ( field )
    {
    case CFB_PRESSURE:
        return c->lhs;
    case CFB_RHS:
        return c->rhs;
    case CFB_CG_R:
        return c->cg_r;
    case CFB_CG_P:
        return c->cg_p;
    case CFB_CG_Q:
        return c->cg_q;
    }
    return nullptr;
}

// ---- kernel launchers (each returns the number of kernels it enqueued) -----------------------
// kernels_fields.cu
int launch_add_inputs( cfb_ctx* c );
int launch_advect( cfb_ctx* c );
int launch_apply_pressure( cfb_ctx* c );
int launch_fill_synthetic( cfb_ctx* c, int variant, uint64_t seed );
// kernels_cg.cu
int launch_divergence( cfb_ctx* c );          // rhs = -div/h ; lhs = 0
int launch_cg_init( cfb_ctx* c, int fixed );  // x=0, r=b, p=Minv r, rr, rz ; + check kernel
inline bool cg_peer_mode( const cfb_ctx* c )
{
    return c->cfg.use_nccl && c->peer_ok && c->use_peer && c->cg_variant >= 1;
}
// peer mode with the exchange taken off the critical path (cfb_api.cu: enqueue_iteration); the 72-byte form only
// (phase A' of the 64-byte form reads the ghosts of the search direction), not with the staging-area reads or the
// two-dimensional FLAT kernels (no mailbox instantiation of those)
inline bool peer_xstaged( const cfb_ctx* c );
inline bool peer_overlapped( const cfb_ctx* c )
{
    return cg_peer_mode( c ) && c->peer_overlap && c->cg_variant == 1 && !c->peer_xstage_reads &&
           !( c->g.D == 2 && c->flat_2d );
}
// peer mode with x neighbours and tiles that end exactly on the block: phase B reads its x ghosts
// straight from the staging areas, so the iterations never scatter them into the ghost columns
inline bool peer_xstaged( const cfb_ctx* c )
{
    // (not with cg_variant 2: its phase A' reads the x ghosts of p from the ghost columns through TMA)
    return cg_peer_mode( c ) && c->peer_xstage_reads && c->cg_variant == 1 && ( c->nbr[0] >= 0 || c->nbr[1] >= 0 ) &&
           c->g.n[0] % c->fu_tx == 0;
}
int launch_cg_axpy( cfb_ctx* c );             // kernel 1 (+ fused kernel-2 reduction)
int launch_cg_pupdate( cfb_ctx* c );          // convergence bookkeeping + kernel 3
// kernels_stencil.cu
int stencil_setup( cfb_ctx* c );              // builds the tensor map for cg_p
int launch_stencil_dot( cfb_ctx* c );         // kernel 4
int launch_stencil_rupdate( cfb_ctx* c );     // cg_variant 2, phase A': r -= alpha (A p) without a stored q
// cg_variant 3: w = A M^-1 r with the three sums of the iteration (init: the launch that starts a solve); mail: the
// last block runs the mailbox reduction (NVLink peer path)
int launch_cg1_stencil( cfb_ctx* c, int init, bool mail );
// multigrid: fine-level sweeps on the TMA march (-1: does not apply, the caller runs its own kernels)
bool mg_tma_applies( const cfb_ctx* c );
int mg_tma_prepare( cfb_ctx* c );
int launch_mg_smooth_tma( cfb_ctx* c, const OpConst& op, double omega, double* b, double* x0, double* x1, int xi_is, int dot );
int launch_mg_smooth02_tma( cfb_ctx* c, const OpConst& op, double omega1, double omega2, double* b, double* x0, double* x1 );
int launch_mg_prolong_smooth_tma( cfb_ctx* c, const OpConst& op, double omega, double* b, double* x0, double* x1, int xi_is,
                                  double* e, long long csy, long long csz, const int cn[3], int dot );
// kernels_cg1.cu: p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s
int launch_cg1_update( cfb_ctx* c );
int cg1_pcg_solve( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid );
// kernels_fused.cu
int fused_setup( cfb_ctx* c );                // tensor maps of cg_r, cg_p + the unit list
int launch_cg_rupdate( cfb_ctx* c );          // phase A: r -= alpha q, sum r^2, sum r.Minv r
int launch_cg_fused( cfb_ctx* c, int which ); // phase B: 0 = all units, 1 = interior, 2 = boundary
int launch_cg_rupdate_mail( cfb_ctx* c );      // phase A, its last block runs the mailbox reduction (peer_overlap)
int launch_cg_fused_mail( cfb_ctx* c, int which, bool side ); // phase B units (1 interior / 2 boundary), the same
int launch_cg_finish( cfb_ctx* c );
bool cg_persist_supported( const cfb_ctx* c ); // tile shape instantiated for the persistent kernel
int launch_cg_persistent( cfb_ctx* c, int iters ); // `iters` iterations in one cooperative launch
// the persistent form applies: two-kernel form, Jacobi, one block, no per-kernel timing; automatic choice by size
inline bool cg_persist_eligible( const cfb_ctx* c ) // ... apart from the CG form
{
    if ( c->cfg.use_nccl || c->time_kernels || c->cg_persist == 0 || !cg_persist_supported( c ) )
        return false;
    if ( c->cg_persist > 0 )
        return true;
    // automatic: where it measured faster than the launch-per-phase kernels (profiles/r2_small_grids.json: 96^3
    // 22.8 vs 28.9 us per iteration, 128^3 35.6 vs 41.0, 160^3 85.0 vs 87.9; 64^3 21.9 vs 20.7 and 192^3 and above
    // no gain: there the data movement, not the launch / reduction latency, is what an iteration costs)
    const double cells = (double)c->g.n[0] * c->g.n[1] * c->g.n[2];
    return cells >= 4.0e5 && cells <= 5.0e6;
}
inline bool cg_persist_applies( const cfb_ctx* c ) { return c->cg_variant == 1 && cg_persist_eligible( c ); }
// The CG form a solve runs when none was chosen ("cg_variant" -1, the default).  Forms 0, 1 and 2 produce identical bits,
// so this is a pure performance choice, made from measurements (profiles/r2_sweep_forms2.log, r2_sweep_rtma.log,
// r2_small_grids.json): the 64-byte form (2) for three-dimensional blocks of 7e6 cells (192^3) and more — since its
// phase A' stages r by TMA it is ahead at every size measured there: 131 vs 143 us per iteration at 192^3, 191 vs 202 at
// 224^3, 229 vs 257 at 256^3, 415 vs 471 at 320^3, 652 vs 755 at 384^3, 1014 vs 1136 at 448^3, 1526 vs 1658 at 512^3; the
// 72-byte form (1) everywhere else: in two dimensions far ahead (8192^2: 1526 vs 1870 us: the one-plane z-march of phase
// A' has nothing to pipeline), small blocks run its persistent single-launch form, and the overlapped exchange and the
// staging-area reads are schedules of this form.
inline int cg_variant_auto( const cfb_ctx* c )
{
    if ( c->cg_variant >= 0 )
        return c->cg_variant;
    const bool peer = c->cfg.use_nccl && c->peer_ok && c->use_peer;
    if ( cg_persist_eligible( c ) || ( peer && ( c->peer_overlap || c->peer_xstage_reads ) ) )
        return 1;
    const double cells = (double)c->g.n[0] * c->g.n[1] * c->g.n[2];
    return ( c->g.D == 3 && cells >= 7.0e6 ) ? 2 : 1;
}
// output.cu: SiloWriter::siloWrite re-designed (extraction kernel + asynchronous copy now, files later)
int output_write( cfb_ctx* c, const char* dir, int time_step );
int output_flush( cfb_ctx* c );
const char* output_solve_dir( const cfb_ctx* c ); // directory set with cfb_set_output_dir, or nullptr
void output_destroy( cfb_ctx* c );
// mg.cu: the CG with z = V-cycle( r ) instead of z = D^-1 r
int mg_pcg_solve( cfb_ctx* c, int fixed_iters, int* num_iter, double* resid );
void mg_destroy( cfb_ctx* c );
int mg_set_graph( cfb_ctx* c, bool on );
int mg_set_coarse_kernel( cfb_ctx* c, bool on );
// halo.cu
int halo_init( cfb_ctx* c );
void halo_destroy( cfb_ctx* c );
int halo_exchange_cells( cfb_ctx* c, double* field, int width ); // face-neighbour exchange of a cell array
// the same for up to two cell arrays in one message per neighbour, split into begin (fork to the side
// stream) / end (join), so that work not reading ghosts can be launched in between
int halo_cells_begin( cfb_ctx* c, double* const* fields, int nf, int width );
int halo_cells_end( cfb_ctx* c );
// peer-memory exchange of the two-kernel iteration: store my boundary layers of the given cell arrays
// (0: cg_r, 1: the p buffer `pbuf`) into the neighbours' ghost layers, publish my local double-double
// sums (which = 0: pAp, 1: rz_new and rr) to every rank, wait for theirs, combine exactly
// (copy_r / copy_pbuf >= 0: which arrays' faces travel: r, and / or that p buffer; unpack: scatter the
// received x faces from the staging area into the ghost columns — not needed when phase B reads them
// from staging, always needed for the plain stencil at the start of a solve)
int peer_exchange( cfb_ctx* c, int which, bool copy_r, int copy_pbuf, bool unpack );
// overlapped exchange: on the side stream, once `after` (an event of the main stream) has happened, store my
// boundary layers of cg_r (kind 0) or of search-direction buffer `pbuf` (kind 1) into the neighbours' ghost
// layers, tell each neighbour, wait until each neighbour has told me, scatter the x faces I received; then record
// c->ev_ghost.  No reduction here: the compute kernels' last blocks run the mailboxes (device_peer.cuh).
int peer_faces_async( cfb_ctx* c, int kind, int pbuf, cudaEvent_t after ); // after == nullptr: no wait
int peer_faces_join( cfb_ctx* c ); // main stream waits for everything peer_faces_async has enqueued
int halo_exchange_fields( cfb_ctx* c, int version );             // width-h exchange of q,u,v,w
int halo_sendrecv_slots( cfb_ctx* c, const size_t counts[6], cudaStream_t st ); // d_halo_send/recv[s] <-> nbr[s]
// NVLink peer memory for further arrays: every rank passes its `count` device allocations (same count, same
// order on all ranks); on success mapped[s * count + a] is the face-s neighbour's array a in this process
// (nullptr where there is no neighbour) and *ok is true on every rank, otherwise *ok is false on every rank.
int peer_map_arrays( cfb_ctx* c, int count, double* const* mine, std::vector<double*>& mapped, bool* ok );
int halo_allreduce( cfb_ctx* c, double* dev_vals, int n );
int halo_allgather( cfb_ctx* c, const double* dev_send, double* dev_recv, int n_per_rank );
// kernels_cg.cu: all-gather + exact combine of the local CG sums (which = 0: pAp, 1: rz_new and rr)
int cg_global_sum( cfb_ctx* c, int which );
