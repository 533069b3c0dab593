"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): element-wise kernels are compared BIT-EXACTLY (the library is
built with -fmad=false and the oracle with -ffp-contract=off, both use explicit fma() in the same
places); anything that goes through a reduction (CG) is compared at 1e-10 relative L2 with the
iteration count within +-1.
"""
import numpy as np
import pytest

from cajitafluids_b200 import Solver, config as K
from helpers import assert_same, fields_of, make_cfg, random_cells, rel_l2, set_both, smooth_velocity, eigen_tol
from oracle_api import Oracle

pytestmark = pytest.mark.gpu

TOL_FIELD = 1e-10  # north_star: fields within 1e-10 relative L2 in FP64

SHAPES = [
    (3, (32, 32, 32)),
    (3, (40, 24, 20)),   # ragged: not a multiple of any tile
    (3, (66, 17, 9)),    # odd sizes: last column pair is half-valid
    (3, (130, 36, 5)),
    (2, (64, 64)),
    (2, (50, 30)),
    (2, (33, 47)),
]


def pair(cfg):
    return Solver(cfg), Oracle(cfg)


def box_for(cells):
    # cubic cells of size 1/n0: extent_d = n_d / n0   (src/Mesh.hpp:56-64)
    return tuple(c / cells[0] for c in cells)


@pytest.mark.parametrize("dim,cells", SHAPES)
@pytest.mark.parametrize("walls", ["solid", "mixed"])
def test_stencil_dot_bit_exact(dim, cells, walls):
    bt = [K.SOLID] * 6 if walls == "solid" else [K.FREE, K.SOLID, K.SOLID, K.FREE, K.SOLID, K.FREE]
    cfg = make_cfg(dim, cells, box=box_for(cells), boundary_type=bt)
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(1234)
    p = random_cells(gpu, rng, K.CG_P)
    set_both(gpu, ora, K.CG_P, p)
    dg, _ = gpu.stencil_dot(1)
    do, _ = ora.stencil_dot(1)
    qg, qo = gpu.get(K.CG_Q), ora.get(K.CG_Q)
    assert np.array_equal(qg, qo), f"max abs diff {np.abs(qg - qo).max()}"
    assert abs(dg - do) <= 1e-13 * max(abs(do), np.abs(p * qo).sum() * 1e-3 + 1e-300)


@pytest.mark.parametrize("variant,tx,ty,st", [(0, 64, 16, 4), (0, 64, 8, 4), (0, 64, 32, 3), (0, 128, 16, 3),
                                              (0, 128, 8, 4), (0, 64, 16, 6), (1, 64, 8, 0)])
def test_stencil_variants_agree(variant, tx, ty, st):
    cells = (70, 50, 21)
    cfg = make_cfg(3, cells, box=box_for(cells))
    gpu, ora = pair(cfg)
    gpu.set_tuning("stencil_variant", variant)
    if variant == 0:
        gpu.set_tuning("stencil_stages", st)
        gpu.set_tuning("stencil_tx", tx)
        gpu.set_tuning("stencil_ty", ty)
    gpu.set_tuning("stencil_zc", 8)  # several z chunks
    rng = np.random.default_rng(7)
    p = random_cells(gpu, rng, K.CG_P)
    set_both(gpu, ora, K.CG_P, p)
    dg, _ = gpu.stencil_dot(2)
    do, _ = ora.stencil_dot(1)
    assert np.array_equal(gpu.get(K.CG_Q), ora.get(K.CG_Q))
    assert abs(dg - do) <= 1e-12 * abs(do) + 1e-300


def test_stencil_eigenvector_known_answer():
    """p = prod cos(pi m (g+1/2)/n) is an eigenvector of the all-SOLID operator."""
    n = (48, 32, 40)
    cfg = make_cfg(3, n, box=box_for(n))
    gpu = Solver(cfg)
    m = (3, 2, 5)
    ax = [np.cos(np.pi * m[d] * (np.arange(n[d]) + 0.5) / n[d]) for d in range(3)]
    p = ax[2][:, None, None] * ax[1][None, :, None] * ax[0][None, None, :]
    gpu.set(K.CG_P, p)
    gpu.stencil_dot(1)
    q = gpu.get(K.CG_Q)
    h, dt, _ = gpu.scalars()
    scale = dt / (cfg.density * h * h)
    lam = scale * sum(2 - 2 * np.cos(np.pi * m[d] / n[d]) for d in range(3))
    assert rel_l2(q, lam * p) < 1e-13


@pytest.mark.parametrize("dim,cells", SHAPES[:5])
def test_build_rhs_and_apply_pressure_bit_exact(dim, cells):
    for q1 in (0, 1):
        cfg = make_cfg(dim, cells, box=box_for(cells), quirks=(q1, 1))
        gpu, ora = pair(cfg)
        rng = np.random.default_rng(99)
        for f in fields_of(dim)[1:]:
            set_both(gpu, ora, f, rng.uniform(-1, 1, size=gpu.shape(f)))
        gpu.build_rhs()
        ora.build_rhs()
        assert np.array_equal(gpu.get(K.RHS), ora.get(K.RHS))
        p = random_cells(gpu, rng, K.PRESSURE)
        set_both(gpu, ora, K.PRESSURE, p)
        gpu.apply_pressure()
        ora.apply_pressure()
        for f in fields_of(dim)[1:]:
            assert np.array_equal(gpu.get(f, region=K.GHOSTED), ora.get(f, region=K.GHOSTED)), (f, q1)


@pytest.mark.parametrize("dim,cells", [(3, (32, 32, 32)), (3, (40, 24, 20)), (2, (64, 64)), (2, (50, 30))])
def test_add_inputs_bit_exact(dim, cells):
    cfg = make_cfg(dim, cells, box=box_for(cells), body_force=(0.3, -9.8, 0.1)[:dim] + (0.0,) * (3 - dim))
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(5)
    for f in fields_of(dim):
        set_both(gpu, ora, f, rng.uniform(-0.5, 0.5, size=gpu.shape(f)))
    gpu.add_inputs()
    ora.add_inputs()
    for f in fields_of(dim):
        assert np.array_equal(gpu.get(f), ora.get(f)), f
    assert gpu.dt == ora.dt  # same clamp (src/Solver.hpp:96-106)


@pytest.mark.parametrize("dim,cells", [(3, (32, 32, 32)), (3, (40, 24, 20)), (2, (64, 64)), (2, (50, 30))])
@pytest.mark.parametrize("order", [1, 3])
@pytest.mark.parametrize("q2", [0, 1])
def test_advection_bit_exact(dim, cells, order, q2):
    cfg = make_cfg(dim, cells, box=box_for(cells), interp_order=order, quirks=(0, q2))
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(2024)
    h, dt, _ = gpu.scalars()
    vel = smooth_velocity(gpu, rng, amp=0.9 * h / dt)  # CFL 0.9
    for f, a in vel.items():
        set_both(gpu, ora, f, a)
    set_both(gpu, ora, K.QUANTITY, random_cells(gpu, rng))
    gpu.time_integrator_step()
    ora.time_integrator_step()
    for f in fields_of(dim):
        a, b = gpu.get(f), ora.get(f)
        assert np.array_equal(a, b), (f, np.abs(a - b).max())


@pytest.mark.parametrize("dim,cells", [(3, (32, 32, 32)), (3, (40, 24, 20)), (2, (64, 64)), (2, (96, 40))])
def test_pcg_solve_parity(dim, cells):
    cfg = make_cfg(dim, cells, box=box_for(cells))
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(11)
    h, dt, _ = gpu.scalars()
    vel = smooth_velocity(gpu, rng, amp=1.0)
    for f, a in vel.items():
        set_both(gpu, ora, f, a)
    # wall-normal velocities must vanish on the SOLID walls or the all-Neumann system is inconsistent
    gpu.add_inputs()
    ora.add_inputs()
    gpu.build_rhs()
    ora.build_rhs()
    ig, rg = gpu.pcg_solve()
    io, ro = ora.pcg_solve()
    assert abs(ig - io) <= 1, (ig, io)  # the bar; exact sums on both sides give equality:
    assert ig == io
    assert rg <= cfg.cg_tolerance and ro <= cfg.cg_tolerance
    assert_same(gpu.get(K.PRESSURE), ora.get(K.PRESSURE), "pressure")
    hg, ho = gpu.residual_history(), ora.residual_history()
    assert len(hg) == len(ho) == ig
    assert np.array_equal(hg, ho)


FUSED_TILINGS = [(64, 16, 3), (64, 16, 4), (64, 8, 3), (64, 8, 4), (64, 32, 2), (64, 32, 3), (128, 8, 3),
                 (128, 8, 4), (128, 16, 3)]


@pytest.mark.parametrize("cells,walls", [((70, 50, 21), "solid"), ((130, 36, 5), "mixed"), ((33, 47), "solid")])
def test_cg_two_kernel_form_matches_three_kernel_form_and_oracle(cells, walls):
    """The 72 B/cell two-kernel iteration (kernels_fused.cu) against the 88 B/cell three-kernel one
    and the oracle: same iteration count, residual history and pressure, bit for bit, for every
    tiling (several z chunks, ragged tiles)."""
    dim = len(cells)
    bt = [K.SOLID] * 6 if walls == "solid" else [K.FREE, K.SOLID, K.SOLID, K.FREE, K.SOLID, K.FREE]
    cfg = make_cfg(dim, cells, box=box_for(cells), boundary_type=bt)
    ora = Oracle(cfg)
    rng = np.random.default_rng(21)
    vel = None
    runs = [(0, None)] + [(1, t) for t in FUSED_TILINGS]
    ref = None
    for variant, tiling in runs:
        gpu = Solver(cfg)
        gpu.set_tuning("cg_variant", variant)
        if tiling:
            gpu.set_tuning("fused_stages", tiling[2])
            gpu.set_tuning("fused_zc", 8)
            gpu.set_tuning("fused_tx", tiling[0])
            gpu.set_tuning("fused_ty", tiling[1])
        if vel is None:
            vel = smooth_velocity(gpu, rng, amp=1.0)
            for f, a in vel.items():
                ora.set(f, a)
            ora.add_inputs()
            ora.build_rhs()
            io, ro = ora.pcg_solve()
            ref = (io, ro, ora.get(K.PRESSURE), ora.residual_history())
        for f, a in vel.items():
            gpu.set(f, a)
        gpu.add_inputs()
        gpu.build_rhs()
        ig, rg = gpu.pcg_solve()
        assert ig == ref[0], (variant, tiling, ig, ref[0])
        assert rg == ref[1], (variant, tiling)
        assert np.array_equal(gpu.get(K.PRESSURE), ref[2]), (variant, tiling)
        assert np.array_equal(gpu.residual_history(), ref[3]), (variant, tiling)
        # the vectors the next iteration would use are the same too
        if variant == 0:
            keep = {f: gpu.get(f) for f in (K.CG_R,)}
        else:
            assert np.array_equal(gpu.get(K.CG_R), keep[K.CG_R]), (variant, tiling)


@pytest.mark.parametrize("variant", [0, 1])
def test_cg_fixed_iterations_both_forms(variant):
    cells = (48, 40, 36)
    cfg = make_cfg(3, cells, box=box_for(cells), fixed_iters=17)
    gpu, ora = pair(cfg)
    gpu.set_tuning("cg_variant", variant)
    rng = np.random.default_rng(8)
    b = random_cells(gpu, rng, K.RHS)
    b -= b.mean()
    set_both(gpu, ora, K.RHS, b)
    ig, rg = gpu.pcg_solve()
    io, ro = ora.pcg_solve()
    assert ig == io == 17 and rg == ro
    for f in (K.PRESSURE, K.CG_R, K.CG_P, K.CG_Q):
        assert np.array_equal(gpu.get(f), ora.get(f)), f


def test_pcg_zero_rhs_returns_immediately():
    cfg = make_cfg(3, 16)
    gpu = Solver(cfg)
    gpu.build_rhs()  # zero velocity -> zero rhs
    it, res = gpu.pcg_solve()
    assert it == 0 and res == 0.0
    assert not gpu.get(K.PRESSURE).any()


def test_pcg_not_converged_is_an_error():
    from cajitafluids_b200 import CfbError
    cfg = make_cfg(3, 32, max_iter=5)
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(3)
    for f, a in smooth_velocity(gpu, rng).items():
        set_both(gpu, ora, f, a)
    gpu.build_rhs()
    ora.build_rhs()
    with pytest.raises(CfbError) as e:
        gpu.pcg_solve()
    assert e.value.code == K.ERR_NOT_CONVERGED
    with pytest.raises(CfbError):
        ora.pcg_solve()
    # both stopped after exactly max_iter iterations with the same iterate
    assert rel_l2(gpu.get(K.PRESSURE), ora.get(K.PRESSURE)) < 1e-12


def test_pcg_fixed_iterations_and_host_entry():
    cfg = make_cfg(3, (48, 40, 36), box=box_for((48, 40, 36)), fixed_iters=25)
    gpu, ora = pair(cfg)
    rng = np.random.default_rng(8)
    b = random_cells(gpu, rng, K.RHS)
    b -= b.mean()
    ora.set(K.RHS, b)
    x, it, res = gpu.pcg_solve_host(b)
    io, ro = ora.pcg_solve()
    assert it == io == 25
    assert_same(x, ora.get(K.PRESSURE), "x")
    assert res == ro


@pytest.mark.parametrize("dim,cells,steps", [(2, 64, 25), (3, 32, 12)])
def test_full_steps_reference_defaults(dim, cells, steps):
    """setup() + steps of the default advection problem (examples/advection.cpp defaults; 2-D runs
    with the reference quirks Q1/Q2 on): fields 1e-10, CG iterations +-1 every step."""
    cfg = make_cfg(dim, cells)
    gpu, ora = pair(cfg)
    gpu.setup()
    ora.setup()
    ig, io = gpu.stats()["cg_iterations"], ora.stats()["cg_iterations"]
    assert abs(ig - io) <= 1
    for s in range(steps):
        gpu.step()
        ora.step()
        ng, no = gpu.stats()["cg_iterations"], ora.stats()["cg_iterations"]
        assert abs((ng - ig) - (no - io)) <= 1, f"step {s}: {ng - ig} vs {no - io}"
        ig, io = ng, no
    for f in fields_of(dim) + [K.PRESSURE]:
        assert_same(gpu.get(f), ora.get(f), f"field {f}")
    assert gpu.time == ora.time


def test_config0_64cubed_linear_interp():
    """BASELINE config 0 (shortened to 20 of its 100 steps to bound the oracle's CPU time):
    64^3, linear field interpolation, local PCG."""
    cfg = make_cfg(3, 64, interp_order=1)
    gpu, ora = pair(cfg)
    gpu.setup()
    ora.setup()
    for _ in range(20):
        gpu.step()
        ora.step()
    assert abs(gpu.stats()["cg_iterations"] - ora.stats()["cg_iterations"]) <= 21
    assert gpu.stats()["cg_iterations"] == ora.stats()["cg_iterations"]
    for f in fields_of(3) + [K.PRESSURE]:
        assert_same(gpu.get(f), ora.get(f), f"field {f}")


def test_solve_loop_matches_reference_step_count():
    cfg = make_cfg(2, 32)
    gpu, ora = pair(cfg)
    ng = gpu.solve(0.05, 0)
    no = ora.solve(0.05, 0)
    assert ng == no
    assert gpu.time == ora.time
    for f in fields_of(2):
        assert_same(gpu.get(f), ora.get(f), f"field {f}")


def test_projection_makes_velocity_divergence_free():
    """after correctVelocity (Q1 off) the discrete divergence is at CG-tolerance level
    and wall-normal velocities are exactly zero."""
    cfg = make_cfg(3, 40, quirks=(0, 1), tol=1e-10)
    gpu = Solver(cfg)
    rng = np.random.default_rng(21)
    for f, a in smooth_velocity(gpu, rng).items():
        gpu.set(f, a)
    gpu.add_inputs()
    gpu.correct_velocity()
    gpu.build_rhs()
    div = gpu.get(K.RHS)
    assert np.linalg.norm(div.ravel()) < 1e-8
    u, v, w = gpu.get(K.U), gpu.get(K.V), gpu.get(K.W)
    assert not u[:, :, 0].any() and not u[:, :, -1].any()
    assert not v[:, 0, :].any() and not v[:, -1, :].any()
    assert not w[0].any() and not w[-1].any()


def test_large_grid_properties_256():
    """Size-independent properties at a BASELINE size (256^3): eigenvector identity, constant
    null-space, and CG's recurrence residual equals the true residual b - A x."""
    n = 256
    cfg = make_cfg(3, n, fixed_iters=40)
    gpu = Solver(cfg)
    ax = np.cos(np.pi * 2 * (np.arange(n) + 0.5) / n)
    p = ax[:, None, None] * ax[None, :, None] * ax[None, None, :]
    gpu.set(K.CG_P, p)
    gpu.stencil_dot(1)
    h, dt, _ = gpu.scalars()
    lam = dt / (cfg.density * h * h) * 3 * (2 - 2 * np.cos(np.pi * 2 / n))
    assert rel_l2(gpu.get(K.CG_Q), lam * p) < eigen_tol(n)
    gpu.set(K.CG_P, np.ones_like(p))
    dot, _ = gpu.stencil_dot(1)
    assert np.abs(gpu.get(K.CG_Q)).max() < 1e-9 * dt / (cfg.density * h * h)
    # CG: true residual == recurrence residual after 40 iterations
    gpu.fill_synthetic_velocity(0)
    gpu.build_rhs()
    b = gpu.get(K.RHS)
    it, res = gpu.pcg_solve()
    assert it == 40
    x = gpu.get(K.PRESSURE)
    gpu.set(K.CG_P, x)
    gpu.stencil_dot(1)
    true_r = b - gpu.get(K.CG_Q)
    assert abs(np.linalg.norm(true_r.ravel()) - res) <= 1e-9 * np.linalg.norm(b.ravel())


def test_config1_128cubed_cubic_interp():
    """BASELINE config 1: 128^3 on one B200, cubic-spline advection + PCG with the reference's tolerance,
    field by field against the CPU run (shortened to setup + 2 steps, ~1750 CG iterations, to bound the
    oracle's CPU time)."""
    cfg = make_cfg(3, 128, interp_order=3)
    gpu, ora = pair(cfg)
    gpu.setup()
    ora.setup()
    for _ in range(2):
        gpu.step()
        ora.step()
    ig, io = gpu.stats()["cg_iterations"], ora.stats()["cg_iterations"]
    assert abs(ig - io) <= 3 and ig == io, (ig, io)
    for f in fields_of(3) + [K.PRESSURE]:
        assert_same(gpu.get(f), ora.get(f), f"field {f}")
    assert gpu.time == ora.time
