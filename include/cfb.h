/* cfb.h — C ABI of the B200-native pressure-projection / advection path.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`,
 * takes plain pointers / sizes / PODs and returns an int status (0 == CFB_OK).
 * Each function cites the reference interface it replaces (paths relative to the
 * cajitafluids source tree).  The host-side C++ shims in include/cajitafluids_b200/
 * rebuild the reference's `Mesh / ProblemManager / BoundaryCondition / InflowSource /
 * BodyForce / VelocityCorrector / Solver` classes on top of these calls, and
 * cajitafluids_b200/_capi.py binds the same symbols with ctypes for the tests.
 *
 * Memory model: a cfb_ctx owns all device memory, streams, graphs and the NCCL
 * communicator of ONE rank == ONE GPU (the reference: one MPI rank == one Kokkos
 * device).  Pointers handed out by cfb_field_ptr are borrowed device pointers,
 * invalidated by cfb_destroy; Current/Next pointers swap on cfb_advance /
 * cfb_time_integrator_step exactly like ProblemManager::advance.
 * A ctx is not thread-safe (one caller thread, like one MPI rank).
 *
 * Array layout (device): every field lives in a padded 3-D box, x (i) fastest:
 *     offset(i,j,k) = (k + hz) * stride_z + (j + hy) * stride_y + (i + hx)
 * with (i,j,k) the 0-based OWNED index of the block, hx = 16 (128-byte aligned
 * first owned entity), hy = hz = halo width (dim == 2: one owned plane between zero ghost planes).
 * All fields of one ctx share the same strides.  Ghost entities on physical
 * walls are allocated, zero-filled once and never written (reference behaviour:
 * src/ProblemManager.hpp:149-165, tests/tstMesh.cpp:61-68).
 */
#ifndef CFB_H
#define CFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFB_ABI_VERSION 1
#define CFB_NCCL_ID_BYTES 128

/* status codes */
enum
{
    CFB_OK = 0,
    CFB_ERR_INVALID = 1,       /* bad argument / config   (std::runtime_error "invalid backend" etc.) */
    CFB_ERR_MESH_EXTENT = 2,   /* src/Mesh.hpp:56-64  std::logic_error "Extent not evenly divisible..." */
    CFB_ERR_CUDA = 3,          /* CUDA runtime error, text in cfb_last_error */
    CFB_ERR_NCCL = 4,          /* NCCL error */
    CFB_ERR_NOT_CONVERGED = 5, /* Cajita CG: std::runtime_error "CG solver did not converge" */
    CFB_ERR_NO_DEVICE = 6      /* no sm_100 GPU: the product has no CPU fallback */
};

/* src/BoundaryConditions.hpp:30-37 */
enum
{
    CFB_SOLID = 0,
    CFB_FREE = 1
};

/* entity/field ids: (Cell,Quantity), (FaceI,Velocity), (FaceJ,Velocity), (FaceK,Velocity)
 * src/ProblemManager.hpp:281-350; PRESSURE/RHS are VelocityCorrector::_lhs/_rhs
 * (src/VelocityCorrector.hpp:83-90); CG_* are the solver work vectors. */
enum
{
    CFB_QUANTITY = 0,
    CFB_U = 1,
    CFB_V = 2,
    CFB_W = 3,
    CFB_PRESSURE = 4,
    CFB_RHS = 5,
    CFB_CG_R = 6,
    CFB_CG_P = 7,
    CFB_CG_Q = 8,
    CFB_NUM_FIELDS = 9
};

/* src/ProblemManager.hpp:35-60 Version::Current / Version::Next */
enum
{
    CFB_CURRENT = 0,
    CFB_NEXT = 1
};

/* region selector for upload/download */
enum
{
    CFB_OWNED = 0,  /* Cajita::Own index space of the entity */
    CFB_GHOSTED = 1 /* Cajita::Ghost index space: owned + halo_cell_width each side */
};

/* CG stopping rule.  Reference: absolute 2-norm, sqrt(sum r^2) <= tol */
enum
{
    CFB_STOP_ABS = 0,
    CFB_STOP_REL = 1 /* sqrt(sum r^2) <= tol * sqrt(sum b^2); extension, never default */
};

/* Preconditioner of the CG.  The reference's "Reference" solver always uses the diagonal one
 * (src/VelocityCorrector.hpp:166-179); MG is this library's opt-in extension in the role HYPRE PFMG
 * plays in the reference's default path (examples/advection.cpp:186-189). */
enum
{
    CFB_PRECOND_JACOBI = 0,
    CFB_PRECOND_MG = 1
};

/* Plain-data mirror of everything createSolver(...) receives
 * (src/Solver.hpp:283-293, examples/advection.cpp:438-459). */
typedef struct cfb_config
{
    int32_t struct_size; /* = sizeof(cfb_config); ABI guard */
    int32_t dim;         /* 2 (the reference) or 3 (our extension) */

    /* Mesh (src/Mesh.hpp:41-103) */
    int32_t global_num_cell[3];
    double global_bounding_box[6]; /* lo[0..2], hi[0..2] */
    int32_t halo_cell_width;       /* 3: src/Solver.hpp:78 */

    /* Cajita::DimBlockPartitioner result: block grid and this rank's block */
    int32_t ranks_per_dim[3];
    int32_t block_id[3];
    int32_t world_rank;
    int32_t world_size;

    /* Solver scalars */
    double density;
    double delta_t;   /* requested; clamped like src/Solver.hpp:96-106 when clamp_dt */
    int32_t clamp_dt; /* 1 = reference behaviour */

    /* BoundaryCondition<D>::boundary_type  (src/BoundaryConditions.hpp:131-134)
     * index d = low wall of dim d, index dim + d = high wall of dim d
     * (2D: [-x,-y,+x,+y] == the reference's order). */
    int32_t boundary_type[6];

    /* InflowSource (src/InflowSource.hpp:80-90) */
    double inflow_location[3];
    double inflow_size[3];
    double inflow_velocity[3];
    double inflow_quantity;

    /* BodyForce (src/BodyForce.hpp:62-66) */
    double body_force[3];

    /* MeshInitFunc constant initial state (examples/advection.cpp:382-435);
     * arbitrary initial fields go through cfb_upload. */
    double init_quantity;
    double init_velocity[3];

    /* Cajita ReferenceConjugateGradient knobs set at src/VelocityCorrector.hpp:103-105 */
    double cg_tolerance;    /* 1e-6 */
    int32_t cg_max_iter;    /* 2000 */
    int32_t cg_print_level; /* 1 */
    int32_t cg_stop_rule;   /* CFB_STOP_ABS */
    int32_t cg_fixed_iters; /* >0: run exactly this many iterations, ignore tol (benchmarks) */

    /* src/TimeIntegrator.hpp:113 hard-codes 3; README advertises 1 or 3 */
    int32_t field_interp_order;

    /* Reference quirks, replicated behind switches */
    int32_t quirk_applypressure_bc; /* Q1: src/VelocityCorrector.hpp:260 */
    int32_t quirk_rk3_stage3_v0;    /* Q2: src/TimeIntegrator.hpp:57-58  */

    /* device / communicator */
    int32_t device_id;                         /* CUDA ordinal */
    int32_t use_nccl;                          /* 0 when world_size == 1 */
    /* two ncclUniqueIds from rank 0 (cfb_nccl_unique_id): halo communicator, reduction communicator */
    unsigned char nccl_id[2 * CFB_NCCL_ID_BYTES];
} cfb_config;

typedef struct cfb_ctx cfb_ctx;

/* Per-phase device timers (CUDA events) and launch accounting; the NVTX-range
 * equivalent of the reference's Kokkos::Profiling regions. */
typedef struct cfb_stats
{
    double ms_advect;         /* "TimeIntegrator::Step" */
    double ms_add_inputs;     /* "Solve::AddInputs::*" */
    double ms_build_rhs;      /* "VelocityCorrector::BuildProblem" */
    double ms_pcg;            /* "VelocityCorrector::PresureSolve" */
    double ms_apply_pressure; /* "VelocityCorrector::ApplyPressure" */
    double ms_halo;           /* exposed halo + allreduce time on the main stream */
    int64_t kernel_launches;  /* our own kernels launched since create / reset */
    int64_t cg_iterations;    /* total CG iterations since create / reset */
    int64_t steps;            /* Solver::step calls */
    /* per-kernel device time of the CG iterations bracketed with events ("time_kernels" tuning) */
    double ms_k_axpy;
    double ms_k_pupdate;
    double ms_k_stencil;
    int64_t k_timed_iters;
    /* multi-GPU: 1 = the CG iterations exchange ghosts and sums by direct stores into peer memory over
     * NVLink (cudaIpc mappings), 0 = NCCL send/recv + all-gather (or single GPU) */
    int64_t peer_mode;
    /* several GPUs, same bracketing: time on the main stream between the end of the phase-A / phase-B compute
     * kernel and the end of the exchange that follows it (ghost stores + global sums); 0 when the exchange runs
     * inside / beside the compute kernels ("peer_overlap") — ms_k_axpy and ms_k_stencil include these */
    double ms_k_exch_a;
    double ms_k_exch_b;
    /* 1 = overlapped exchange schedule in use for the CG iterations (see the "peer_overlap" tuning key) */
    int64_t peer_overlap;
    /* the CG form a solve runs with the options as they are ("cg_variant": the chosen one, or the automatic choice),
     * and whether batches of its iterations run in the persistent single-launch kernel ("cg_persist") */
    int64_t cg_variant;
    int64_t cg_persist;
} cfb_stats;

/* ---- lifecycle ------------------------------------------------------------------ */

/* Fill cfg with examples/advection.cpp:168-189,446-454 defaults generalised to `dim`. */
int cfb_default_config( cfb_config* cfg, int dim );

/* Split n cells over nb blocks the way Cajita's GlobalGrid does; returns owned count and offset. */
int cfb_partition( int n, int nb, int block, int* owned, int* offset );

/* rank-0 helper: two ncclGetUniqueId results into a 2 * CFB_NCCL_ID_BYTES buffer; the caller
 * broadcasts them to all ranks (torch.distributed / MPI_Bcast) and stores them in cfb_config. */
int cfb_nccl_unique_id( unsigned char* id );

/* Replaces createSolver(...) + Solver ctor (src/Solver.hpp:69-123,283-350):
 * builds mesh, clamps dt, allocates the ProblemManager arrays, the VelocityCorrector
 * vectors, streams, graphs, NCCL comm.  *ctx is set even on failure when possible so
 * cfb_last_error(ctx) can be read; cfb_last_error(NULL) returns the global error. */
int cfb_create( const cfb_config* cfg, cfb_ctx** ctx );
int cfb_destroy( cfb_ctx* ctx );
const char* cfb_last_error( const cfb_ctx* ctx );

/* ---- Mesh / ProblemManager accessors ------------------------------------------- */

/* Mesh::cellSize (src/Mesh.hpp:119-122), the clamped dt (src/Solver.hpp:96-106), _time. */
int cfb_get_scalars( const cfb_ctx* ctx, double* cell_size, double* delta_t, double* time );

/* Cajita index spaces of this block: owned extent of an entity (Own, Local) and the
 * block's global cell offset (IndexConversion::createL2G). */
int cfb_owned_extent( const cfb_ctx* ctx, int field, int ext[3] );
int cfb_global_offset( const cfb_ctx* ctx, int off[3] );

/* ProblemManager::get (src/ProblemManager.hpp:281-350): borrowed device pointer to the
 * padded array; `origin` = element offset of owned entity (0,0,0); strides in elements. */
int cfb_field_ptr( cfb_ctx* ctx, int field, int version, double** dev_ptr,
                   int64_t* origin, int64_t* stride_y, int64_t* stride_z );

/* Host <-> device copy of a dense x-fastest host array covering `region` of the entity.
 * OWNED: ext from cfb_owned_extent; GHOSTED: Cajita's Ghost index space = owned CELLS + 2*halo in
 * each spatial dim, +1 along a face normal on every block (tests/tstMesh.cpp:61-68). */
int cfb_upload( cfb_ctx* ctx, int field, int version, int region, const double* host );
int cfb_download( cfb_ctx* ctx, int field, int version, int region, double* host );

/* ProblemManager::advance (src/ProblemManager.hpp:357-380): swap Current/Next. */
int cfb_advance( cfb_ctx* ctx, int field );
/* ProblemManager::gather (src/ProblemManager.hpp:394-403): width-halo exchange of q,u,v(,w). */
int cfb_gather( cfb_ctx* ctx, int version );

/* ---- the hot path ----------------------------------------------------------------- */

/* Solver::_addInputs (src/Solver.hpp:181-263). */
int cfb_add_inputs( cfb_ctx* ctx );
/* TimeIntegrator::step (src/TimeIntegrator.hpp:120-177): gather, advect all, advance all. */
int cfb_time_integrator_step( cfb_ctx* ctx );
/* VelocityCorrector::_buildRHS + lhs = 0 (src/VelocityCorrector.hpp:182-212,272). */
int cfb_build_rhs( cfb_ctx* ctx );
/* _pressure_solver->solve(rhs, lhs) (src/VelocityCorrector.hpp:276): Jacobi-PCG from
 * x0 = 0 on the current RHS.  Returns iteration count and final sqrt(sum r^2). */
int cfb_pcg_solve( cfb_ctx* ctx, int* num_iter, double* residual_norm );
/* VelocityCorrector::_applyPressure (src/VelocityCorrector.hpp:214-264). */
int cfb_apply_pressure( cfb_ctx* ctx );
/* VelocityCorrectorBase::correctVelocity (src/VelocityCorrector.hpp:266-282). */
int cfb_correct_velocity( cfb_ctx* ctx, int* num_iter, double* residual_norm );
/* SolverBase::setup / step / solve (src/Solver.hpp:125-177). solve returns steps taken. */
int cfb_setup( cfb_ctx* ctx );
int cfb_step( cfb_ctx* ctx );
int cfb_solve( cfb_ctx* ctx, double t_final, int write_freq, int* steps_taken );

/* The solver plug-in surface with HOST vectors: owned-cell dense arrays b (in) and x (out);
 * copies are part of the call (bench.py's e2e leg).  Mirrors
 * ReferenceConjugateGradient::solve(b, x) for callers that keep state on the host. */
int cfb_pcg_solve_host( cfb_ctx* ctx, const double* b_host, double* x_host, int* num_iter,
                        double* residual_norm );

/* ---- output stage (the step after the hot path in Solver::solve) ---- */

/* SiloWriter::writeFile (src/SiloWriter.hpp:56-197): what the reference hands to Silo for this block.
 * One extraction kernel drops the ghosts of q (:136-156) and interpolates the MAC velocity to the cell
 * centres (Interpolation::interpolateVelocity<D,1> at LocalMesh::coordinates( Cell ), :172-186); the
 * node coordinates of the owned cells (:109-123) are computed on the host.  Dense x-fastest HOST
 * arrays: quantity[nz][ny][nx], velocity[D][nz][ny][nx], nodes_d[n_d + 1]; NULL skips an output.
 * Blocking form (tests, callers that want the arrays). */
int cfb_output_extract( cfb_ctx* ctx, double* quantity, double* velocity, double* nodes_x, double* nodes_y,
                        double* nodes_z );
/* SiloWriter::siloWrite( "Mesh", time_step, time, dt ) (src/SiloWriter.hpp:355-417), asynchronous:
 * enqueues the extraction kernel behind the work already queued and the device -> pinned-host copy on
 * an I/O stream, and returns; the files are written at the next cfb_write_output / cfb_output_flush /
 * cfb_destroy.  Silo and PMPIO are absent, so the container is
 *     <dir>/raw/CajitaFluidsOutput<rank:05d><step:05d>.{quantity,velocity,nodes_x,nodes_y[,nodes_z]}.npy
 *     <dir>/CajitaFluids<step:05d>.json      (rank 0: cycle, time, dtime, every block's offset/extent/files;
 *                                             the role of writeMultiObjects, :292-346)
 * following the reference's name pattern (:379-384).  dir == NULL or "" means "data" like the reference. */
int cfb_write_output( cfb_ctx* ctx, const char* dir, int time_step );
int cfb_output_flush( cfb_ctx* ctx );
/* Makes cfb_solve write like Solver::solve does (src/Solver.hpp:156,170-173): once before setup and
 * after every step t with t % write_freq == 0.  NULL / "" turns the writes off (the default). */
int cfb_set_output_dir( cfb_ctx* ctx, const char* dir );
/* Host-only helper: a dense little-endian float64 array as a numpy .npy (version 1.0) file, C order. */
int cfb_write_npy( const char* path, const double* data, int ndim, const int64_t* shape );

/* ---- micro-benchmark / introspection entry points -------------------------------- */

/* q = A p and sum(p*q) on the CG work vectors (fields CFB_CG_P -> CFB_CG_Q), `reps` launches;
 * returns the dot product of the last launch and the mean device time per launch. */
int cfb_stencil_dot( cfb_ctx* ctx, int reps, double* dot, double* ms_per_launch );
/* One fixed-count CG run timed on the device: kernels only, no convergence polling. */
int cfb_pcg_fixed( cfb_ctx* ctx, int iters, double* ms_total, double* residual_norm );
/* Synthetic MAC velocity field of SURVEY §8d (sin/cos product, wall-normal zero) into Current u,v,w. */
int cfb_fill_synthetic_velocity( cfb_ctx* ctx, int variant, uint64_t seed );

int cfb_get_stats( const cfb_ctx* ctx, cfb_stats* out );
int cfb_reset_stats( cfb_ctx* ctx );
/* CG residual history of the last solve (sqrt(sum r^2) after each iteration), up to n entries. */
int cfb_residual_history( const cfb_ctx* ctx, double* hist, int n, int* count );
/* Tuning hook (unknown keys and out-of-range values return CFB_ERR_INVALID; a tile combination nobody instantiated is
 * refused by the solve that would use it).  No key changes results, except that cg_variant 3 is a different —
 * mathematically equivalent — recurrence (iteration counts within +-1 of the others).  Keys:
 *   cg_variant -1|0|1|2|3 CG iteration form: 1 = two kernels 72 B/cell, 0 = three kernels 88 B, 2 = two kernels 64 B (q never
 *                         stored); -1 (default) = 2 for 3-D blocks of >= 7e6 cells (192^3), 1 elsewhere (the three produce
 *                         identical bits: a choice by measurement); 3 = opt-in single-reduction (Chronopoulos-Gear) form: two kernels
 *                         88 B, ONE reduction point and one ghost exchange per iteration
 *   stencil_variant, stencil_tx, stencil_ty, stencil_stages, stencil_zc      tiling of the stencil7 + dot kernel (and of
 *                         phase A' of the 64-byte form, which picks its own z chunk until stencil_zc is set)
 *   fused_auto, fused_tx, fused_ty, fused_stages, fused_zc, fused_reverse, rupdate_ctas   tiling of the two-kernel form
 *   fused_nt 0|256|512    threads per block of phase B (512: the 128 x 16 x 3 tiling only; 0 = 512 in the 64-byte form there)
 *   fused_yc n            2-D runs: tile rows a unit of phase B marches through along y (default: picked from the grid)
 *   flat_2d 0|1           2-D runs: do not load the two zero ghost planes in the TMA kernels (default 1 in 2-D)
 *   advect_tile 0|1       advection kernel: 32 x 2 x 2 entity tiles per block instead of rows (default 0)
 *   poll_every n          convergence polling interval in iterations (0 = auto)
 *   peer_halo 0|1         ghost exchange over NVLink peer memory (default when available) or NCCL send/recv
 *   peer_overlap 0|1      NVLink path: faces on a side stream under interior work, reductions through mailboxes in the
 *                         compute kernels' last blocks, instead of one exchange kernel after each phase
 *   overlap_halo, peer_xstage      further exchange schedules (see csrc/halo.cu)
 *   mg_graph, mg_coarse_kernel     multigrid V-cycle as a CUDA graph / coarse levels in one kernel (one block)
 *   mg_tma -1|0|1, mg_tma_prolong 0|1   multigrid, 3-D: the fine level's smoothing sweeps / its prolongation + first
 *                         post-sweep on the TMA z-march (mg_tma -1, the default: on for one block, off for several
 *                         blocks, where it has not run on hardware yet; mg_tma_prolong default 1)
 *   time_kernels 0|1      record CUDA events around each CG kernel (cfb_stats.ms_k_*) */
int cfb_set_tuning( cfb_ctx* ctx, const char* key, int value );
/* ReferenceConjugateGradient::setTolerance / setMaxIter / setPrintLevel
 * (src/VelocityCorrector.hpp:103-105) after construction. */
int cfb_set_cg_params( cfb_ctx* ctx, double tolerance, int max_iter, int print_level );

/* Choose the CG preconditioner (default CFB_PRECOND_JACOBI == the reference).  CFB_PRECOND_MG: one
 * geometric multigrid V(nu_pre, nu_post) cycle per CG iteration (damped-Jacobi smoother with damping
 * `omega`; <= 0 picks the default: in 3-D with 2..4 sweeps per side one damping per sweep, the reciprocals
 * of the Chebyshev nodes of [0.4, 2], else 6/7 in 3-D and 0.8 in 2-D; nu_coarse sweeps on the coarsest
 * level; 2:1 cell-centred coarsening while all extents stay even).  Same matrix, same stopping test, same solution to the
 * solver tolerance, O(10) iterations instead of O(n).  With several blocks the same global cycle runs
 * block-decomposed (one-layer face exchange per operator application; the cells of every dimension must
 * divide evenly among the blocks), so iteration counts and results do not depend on the decomposition. */
int cfb_set_preconditioner( cfb_ctx* ctx, int kind, int nu_pre, int nu_post, int nu_coarse, double omega );

/* Cap on the number of multigrid levels (0 = as many as the grid allows: coarsening goes on while every
 * block's extents stay even and >= 2 after halving) and the depth in use. */
int cfb_set_mg_max_levels( cfb_ctx* ctx, int max_levels );
int cfb_mg_num_levels( cfb_ctx* ctx, int* levels );
/* One application of the multigrid preconditioner on its own, z = M^-1 r, for dense owned-cell HOST
 * arrays (introspection / tests; overwrites the CG work vector r). */
int cfb_mg_apply( cfb_ctx* ctx, const double* r_host, double* z_host );

int cfb_abi_version( void );

#ifdef __cplusplus
}
#endif
#endif /* CFB_H */
