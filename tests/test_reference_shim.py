"""The oracle against THE REFERENCE ITSELF, run here.

oracle/_ref/libcfref.so is the unmodified cajitafluids source (src/*.hpp compiled from
/root/reference) on single-rank stand-ins for its absent dependencies (oracle/refshim/: Kokkos,
Cabana/Cajita, MPI, Silo, googletest).  Every statement of the hot path that lives in the reference
tree — Solver ctor + dt clamp, _addInputs, InflowSource, BodyForce, BoundaryCondition (matrix and
velocity forms, quirk Q1 included), VelocityCorrector matrix fill / _buildRHS / _applyPressure /
correctVelocity, TimeIntegrator rk3 (quirk Q2 included) / advect / step, Interpolation call sites,
ProblemManager state + advance, SiloWriter's cell-centred velocity — runs as written; the third-party
arithmetic (CG loop, B-splines, index spaces, coordinates) is the stand-in's restatement.

These tests need /root/reference (this container); on the GPU box they skip and the fixtures
generated from the same library (tests/golden/refrun_*.npz, test_golden_refrun.py) stand in.
"""
import os
import subprocess

import numpy as np
import pytest

import reference_api
from cajitafluids_b200 import config as K
from helpers import make_cfg, rel_l2
from oracle_api import Oracle

pytestmark = pytest.mark.skipif(not reference_api.available(),
                                reason="oracle/_ref/libcfref.so needs the reference tree (/root/reference)")

FIELDS = [K.QUANTITY, K.U, K.V, K.PRESSURE, K.RHS]


def cfg_default(n=32):
    return make_cfg(2, n)


def cfg_gravity_free_walls():
    return make_cfg(2, 40, boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE], body_force=(0.0, -9.8, 0.0))


def cfg_rectangular():
    # 48 x 24 cells on [0, 1] x [0, 0.5]: Mesh.hpp:50-64 wants extent_d == n_d * h
    return make_cfg(2, (48, 24), box=(1.0, 0.5))


def cfg_moving_start():
    c = make_cfg(2, 36, body_force=(0.5, -2.0, 0.0))
    c.init_quantity = 0.25
    c.init_velocity[0], c.init_velocity[1] = 0.3, -0.2
    c.inflow_velocity[1] = 0.4
    return c


def cfg_clamped_dt():
    # h = 1/256 < dt = 0.005: the Solver ctor clamps dt to h / umax (src/Solver.hpp:96-106)
    return make_cfg(2, 256, max_iter=3000)


CASES = {
    "default_n32": (cfg_default, 5),
    "gravity_free_walls_n40": (cfg_gravity_free_walls, 4),
    "rectangular_48x24": (cfg_rectangular, 4),
    "moving_start_n36": (cfg_moving_start, 4),
    "clamped_dt_n256": (cfg_clamped_dt, 1),
}


def run_both(cfg, steps, exact=True):
    r, o = reference_api.Reference(cfg, exact=exact), Oracle(cfg)
    its = []
    for c in (r, o):
        c.setup()
        it = [c.stats()["cg_iterations"]]
        for _ in range(steps):
            c.step()
            it.append(c.stats()["cg_iterations"])
        its.append(np.diff([0] + it))
    return r, o, its[0], its[1]


def test_reference_own_unit_tests_pass_on_the_stand_ins():
    """tests/tstMesh.cpp, tstProblemManager.cpp, tstBoundaryConditions.cpp of the reference, compiled
    unmodified: the stand-in's index spaces satisfy the reference's own assertions."""
    for t in ("tstMesh", "tstProblemManager", "tstBoundaryConditions"):
        exe = os.path.join(reference_api.REF_DIR, t)
        p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "FAILED" not in p.stdout and "PASSED" in p.stdout


@pytest.mark.parametrize("name", sorted(CASES))
def test_whole_runs_bit_identical_to_the_reference(name):
    """setup + steps: scalars, per-solve CG iteration counts, every field incl. ghosts, bit for bit."""
    mk, steps = CASES[name]
    cfg = mk()
    r, o, it_r, it_o = run_both(cfg, steps)
    assert r.scalars() == o.scalars()
    assert list(it_r) == list(it_o)
    for f in FIELDS:
        assert np.array_equal(r.get(f), o.get(f)), (name, f)
    for f in (K.QUANTITY, K.U, K.V):
        for ver in (K.CURRENT, K.NEXT):
            assert np.array_equal(r.get(f, ver, K.GHOSTED), o.get(f, ver, K.GHOSTED)), (name, f, ver)
    assert np.array_equal(r.residual_history(), o.residual_history())


@pytest.mark.parametrize("name", ["default_n32", "gravity_free_walls_n40", "moving_start_n36"])
def test_plain_double_reference_within_the_stated_bar(name):
    """The reference's own arithmetic (plain double sums in serial order, no fused multiply-adds in
    the CG updates): north_star's bar — CG iterations +-1, fields <= 1e-10 relative L2."""
    mk, steps = CASES[name]
    r, o, it_r, it_o = run_both(mk(), steps, exact=False)
    assert np.all(np.abs(it_r - it_o) <= 1), (it_r, it_o)
    for f in (K.QUANTITY, K.U, K.V, K.PRESSURE):
        e = rel_l2(o.get(f), r.get(f))
        assert e < 1e-10, (name, f, e)


def test_stored_matrix_and_jacobi_values_match():
    """initializeMatrixValues + BoundaryCondition::build_matrix + the 1/diag fill, SOLID and FREE walls."""
    for cfg in (cfg_default(20), cfg_gravity_free_walls()):
        r, o = reference_api.Reference(cfg), Oracle(cfg)
        A_r, Mi_r = r.matrix()
        A_o, Mi_o = o.matrix()  # ghosted layout [z, y, x, c]
        h = cfg.halo_cell_width
        assert np.array_equal(A_r, A_o[0, h:-h, h:-h, :])
        assert np.array_equal(Mi_r, Mi_o[0, h:-h, h:-h])


def seeded_state(ctx, rng, amp=0.8):
    """Random smooth-ish fields in the GHOSTED index space of every array (the advection reads ghosts)."""
    out = {}
    for f in (K.QUANTITY, K.U, K.V):
        shp = ctx.shape(f, K.GHOSTED)
        a = rng.uniform(-1.0, 1.0, size=shp)
        # a little smoothing keeps the CFL number of the random velocity below 1
        a = 0.25 * (a + np.roll(a, 1, 1) + np.roll(a, 1, 2) + np.roll(a, (1, 1), (1, 2)))
        out[f] = amp * a
    return out


def test_pieces_one_by_one_on_seeded_fields():
    """Each stage alone on the same seeded state: TimeIntegrator::step, _addInputs, _buildRHS, the CG
    solve, _applyPressure (Q1 included)."""
    cfg = make_cfg(2, 28, body_force=(0.3, -1.0, 0.0))
    rng = np.random.default_rng(20251017)
    r, o = reference_api.Reference(cfg), Oracle(cfg)
    st = seeded_state(o, rng)
    for c in (r, o):
        for f, a in st.items():
            c.set(f, a, K.CURRENT, K.GHOSTED)
    for stage in ("time_integrator_step", "add_inputs", "build_rhs", "pcg_solve", "apply_pressure"):
        res = [getattr(c, stage)() for c in (r, o)]
        if stage == "pcg_solve":
            assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
        for f in FIELDS:
            assert np.array_equal(r.get(f), o.get(f)), (stage, f)
        for f in (K.QUANTITY, K.U, K.V):
            assert np.array_equal(r.get(f, K.CURRENT, K.GHOSTED), o.get(f, K.CURRENT, K.GHOSTED)), (stage, f)


def test_not_converged_is_the_same_error():
    cfg = make_cfg(2, 32, max_iter=5)
    r, o = reference_api.Reference(cfg), Oracle(cfg)
    for c in (r, o):
        with pytest.raises(Exception) as e:
            c.setup()
        assert e.value.code == K.ERR_NOT_CONVERGED and "did not converge" in str(e.value)


def test_solve_loop_step_count_and_clock():
    """SolverBase::solve: `do { step; t++ } while ( _time < t_final )` with _time accumulated in FP."""
    cfg = cfg_default(24)
    r, o = reference_api.Reference(cfg), Oracle(cfg)
    n_r, n_o = r.solve(0.0349, 1000), o.solve(0.0349, 1000)
    assert n_r == n_o == 7
    assert r.time == o.time
    for f in FIELDS:
        assert np.array_equal(r.get(f), o.get(f)), f


def test_reference_cli_prints_the_oracles_iteration_counts():
    """The reference's own examples/advection.cpp (-m Reference), compiled unmodified."""
    exe = os.path.join(reference_api.REF_DIR, "advection_ref")
    p = subprocess.run([exe, "-n", "32", "-t", "0.02", "-m", "Reference", "-p", "serial"], capture_output=True,
                       text=True, timeout=300, cwd=reference_api.REF_DIR)
    assert p.returncode == 0, p.stderr
    its = [int(line.split()[4]) for line in p.stdout.splitlines() if line.startswith("Cajita CG Finished")]
    o = Oracle(cfg_default(32))
    o.setup()
    want = [o.stats()["cg_iterations"]]
    for _ in range(4):
        o.step()
        want.append(o.stats()["cg_iterations"])
    assert its == list(np.diff([0] + want))
