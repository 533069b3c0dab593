"""ctypes mirror of `cfb_config` / `cfb_stats` (include/cfb.h) and the reference defaults.

The defaults are those of examples/advection.cpp:168-189 (cells 128, box [0,1]^D, inflow at
(0.2, 0.45) size (0.02, 0.1) velocity (1, 0) quantity 3.0, density 0.1, gravity 0, dt 0.005) and
advection.cpp:446-454 (all walls SOLID, zero initial state), generalised to `dim` as in
SURVEY.md §8d: the inflow box is extended by (0.45, 0.1) in z.
"""
import ctypes as C

NCCL_ID_BYTES = 128

# status codes (include/cfb.h)
OK, ERR_INVALID, ERR_MESH_EXTENT, ERR_CUDA, ERR_NCCL, ERR_NOT_CONVERGED, ERR_NO_DEVICE = range(7)
SOLID, FREE = 0, 1
QUANTITY, U, V, W, PRESSURE, RHS, CG_R, CG_P, CG_Q = range(9)
CURRENT, NEXT = 0, 1
OWNED, GHOSTED = 0, 1
STOP_ABS, STOP_REL = 0, 1
PRECOND_JACOBI, PRECOND_MG = 0, 1


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("dim", C.c_int32),
        ("global_num_cell", C.c_int32 * 3),
        ("global_bounding_box", C.c_double * 6),
        ("halo_cell_width", C.c_int32),
        ("ranks_per_dim", C.c_int32 * 3),
        ("block_id", C.c_int32 * 3),
        ("world_rank", C.c_int32),
        ("world_size", C.c_int32),
        ("density", C.c_double),
        ("delta_t", C.c_double),
        ("clamp_dt", C.c_int32),
        ("boundary_type", C.c_int32 * 6),
        ("inflow_location", C.c_double * 3),
        ("inflow_size", C.c_double * 3),
        ("inflow_velocity", C.c_double * 3),
        ("inflow_quantity", C.c_double),
        ("body_force", C.c_double * 3),
        ("init_quantity", C.c_double),
        ("init_velocity", C.c_double * 3),
        ("cg_tolerance", C.c_double),
        ("cg_max_iter", C.c_int32),
        ("cg_print_level", C.c_int32),
        ("cg_stop_rule", C.c_int32),
        ("cg_fixed_iters", C.c_int32),
        ("field_interp_order", C.c_int32),
        ("quirk_applypressure_bc", C.c_int32),
        ("quirk_rk3_stage3_v0", C.c_int32),
        ("device_id", C.c_int32),
        ("use_nccl", C.c_int32),
        ("nccl_id", C.c_ubyte * (2 * NCCL_ID_BYTES)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_advect", C.c_double),
        ("ms_add_inputs", C.c_double),
        ("ms_build_rhs", C.c_double),
        ("ms_pcg", C.c_double),
        ("ms_apply_pressure", C.c_double),
        ("ms_halo", C.c_double),
        ("kernel_launches", C.c_int64),
        ("cg_iterations", C.c_int64),
        ("steps", C.c_int64),
        ("ms_k_axpy", C.c_double),
        ("ms_k_pupdate", C.c_double),
        ("ms_k_stencil", C.c_double),
        ("k_timed_iters", C.c_int64),
        ("peer_mode", C.c_int64),
        ("ms_k_exch_a", C.c_double),
        ("ms_k_exch_b", C.c_double),
        ("peer_overlap", C.c_int64),
        ("cg_variant", C.c_int64),
        ("cg_persist", C.c_int64),
    ]


def default_config(dim=3, cells=128, *, box=1.0, dt=0.005, density=0.1, gravity=0.0,
                   interp_order=3, quirks=None, print_level=0):
    """Python twin of cfb_default_config (the C function is the authority; a test compares them).

    quirks: None -> reference behaviour where the reference defines it (2-D: Q1 and Q2 on;
    3-D: Q1 off, Q2 on).
    """
    if dim not in (2, 3):
        raise ValueError("dim must be 2 or 3")
    n = (cells,) * 3 if isinstance(cells, int) else tuple(cells) + (1,) * (3 - len(cells))
    cfg = Config()
    cfg.struct_size = C.sizeof(Config)
    cfg.dim = dim
    h = box / n[0] if not isinstance(box, (tuple, list)) else None
    for d in range(3):
        cfg.global_num_cell[d] = n[d] if d < dim else 1
        cfg.global_bounding_box[d] = 0.0
        if isinstance(box, (tuple, list)):
            cfg.global_bounding_box[3 + d] = box[d] if d < dim else 0.0
        else:
            # cubic cells: extent_d = n_d * h  (src/Mesh.hpp:56-64 requires it)
            cfg.global_bounding_box[3 + d] = (n[d] * h if n[d] != n[0] else box) if d < dim else 0.0
        cfg.ranks_per_dim[d] = 1
        cfg.block_id[d] = 0
    cfg.halo_cell_width = 3
    cfg.world_rank, cfg.world_size = 0, 1
    cfg.density, cfg.delta_t, cfg.clamp_dt = density, dt, 1
    for i in range(6):
        cfg.boundary_type[i] = SOLID
    loc, size, vel = (0.2, 0.45, 0.45), (0.02, 0.1, 0.1), (1.0, 0.0, 0.0)
    for d in range(3):
        cfg.inflow_location[d] = loc[d] if d < dim else 0.0
        cfg.inflow_size[d] = size[d] if d < dim else 0.0
        cfg.inflow_velocity[d] = vel[d] if d < dim else 0.0
        cfg.body_force[d] = 0.0
        cfg.init_velocity[d] = 0.0
    cfg.body_force[1] = -gravity  # BodyForce( 0.0, -cl.gravity )  advection.cpp:452
    cfg.inflow_quantity = 3.0
    cfg.init_quantity = 0.0
    cfg.cg_tolerance, cfg.cg_max_iter, cfg.cg_print_level = 1.0e-6, 2000, print_level
    cfg.cg_stop_rule, cfg.cg_fixed_iters = STOP_ABS, 0
    cfg.field_interp_order = interp_order
    if quirks is None:
        cfg.quirk_applypressure_bc = 1 if dim == 2 else 0
        cfg.quirk_rk3_stage3_v0 = 1
    else:
        cfg.quirk_applypressure_bc, cfg.quirk_rk3_stage3_v0 = (1 if quirks[0] else 0), (1 if quirks[1] else 0)
    cfg.device_id, cfg.use_nccl = 0, 0
    return cfg


def copy_config(cfg):
    out = Config()
    C.memmove(C.byref(out), C.byref(cfg), C.sizeof(Config))
    return out
