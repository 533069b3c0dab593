""""flat_2d" tuning key: two-dimensional runs (the reference's own dimensionality) go through the 3-D TMA kernels
with one owned plane between two zero ghost planes; the FLAT instantiations of `stencil7_dot_tma` and
`cg_fused_kernel` take the z neighbours as the zeros they are instead of loading the two ghost planes.  Same values,
every CG form (reference: Cajita::ReferenceConjugateGradient::solve, src/VelocityCorrector.hpp:124-143 in 2-D).

CPU: tests/test_emulated_kernels.py::test_emulated_two_dimensional_runs_without_the_ghost_plane_loads (the kernels
themselves on the host).  GPU: here.  (Sorts last on purpose: written after this round's GPU budget was spent.)
"""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import fields_of, make_cfg
from oracle_api import Oracle


def box_of(cells):
    return tuple(c / cells[0] for c in cells)


@pytest.mark.gpu
@pytest.mark.parametrize("cells,kw", [((64, 64), {}), ((150, 90), dict(boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE])),
                                      ((512, 320), {}), ((1024, 1024), {})])
def test_cuda_two_dimensional_solves_without_the_ghost_plane_loads(cells, kw):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(2, cells, box=box_of(cells), fixed_iters=25, **kw)
    g, o = Solver(cfg), Oracle(cfg)
    g.set_tuning("flat_2d", 1)
    for variant in (1, 2, 0):
        g.set_tuning("cg_variant", variant)
        for s in (g, o):
            s.add_inputs()
            s.build_rhs()
        assert g.pcg_solve() == o.pcg_solve(), variant
        assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE)), variant
        assert np.array_equal(g.get(K.CG_R), o.get(K.CG_R)), variant
    g.close()


@pytest.mark.gpu
def test_cuda_two_dimensional_steps_without_the_ghost_plane_loads():
    from cajitafluids_b200 import Solver
    cfg = make_cfg(2, 64)
    g, o = Solver(cfg), Oracle(cfg)
    g.set_tuning("flat_2d", 1)
    for s in (g, o):
        s.setup()
        for _ in range(3):
            s.step()
    assert g.stats()["cg_iterations"] == o.stats()["cg_iterations"]
    for f in fields_of(2) + [K.PRESSURE]:
        assert np.array_equal(g.get(f), o.get(f)), f
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,kw", [(2, (150, 90), {}), (3, (70, 33, 21), {}), (3, 48, dict(interp_order=1))])
def test_cuda_advection_in_entity_tiles(dim, cells, kw):
    """"advect_tile" tuning key (32 x 2 x 2 entity tiles per block in the advection kernel): the same values."""
    from cajitafluids_b200 import Solver
    box = 1.0 if isinstance(cells, int) else box_of(cells)
    cfg = make_cfg(dim, cells, box=box, **kw)
    g, o = Solver(cfg), Oracle(cfg)
    g.set_tuning("advect_tile", 1)
    for s in (g, o):
        s.setup()
        for _ in range(3):
            s.step()
    assert g.stats()["cg_iterations"] == o.stats()["cg_iterations"]
    for f in fields_of(dim) + [K.PRESSURE]:
        assert np.array_equal(g.get(f), o.get(f)), f
    g.close()


@pytest.mark.gpu
def test_two_dimensional_grids_with_more_tiles_than_the_minimum_scratch():
    """Round-1 host crash (SIGFPE in the tiling loops once one x-y plane had more than 4096 tiles: any 2-D grid
    beyond ~2048^2, the reference's own dimensionality — src/Solver.hpp:50-55).  4096 x 2304: 32 x 144 tiles of
    128 x 16 (phase B), 64 x 144 of 64 x 16 (plain stencil, phase A'), every CG form, with and without the
    ghost-plane loads, against the oracle bit for bit."""
    from cajitafluids_b200 import Solver
    cells = (4096, 2304)
    cfg = make_cfg(2, cells, box=box_of(cells), fixed_iters=6)
    g, o = Solver(cfg), Oracle(cfg)
    for s in (g, o):
        s.add_inputs()
    ref = None
    for flat in (0, 1):
        g.set_tuning("flat_2d", flat)
        for variant in (1, 2, 0):
            g.set_tuning("cg_variant", variant)
            g.build_rhs()
            res = g.pcg_solve()
            if ref is None:
                o.build_rhs()
                ref = (o.pcg_solve(), o.get(K.PRESSURE))
            assert res == ref[0], (flat, variant)
            assert np.array_equal(g.get(K.PRESSURE), ref[1]), (flat, variant)
    g.close()
    o.close()


@pytest.mark.gpu
def test_two_dimensional_8192_squared():
    """The probe size of bench.py (8192^2 = the cell count of 406^3): 64 x 512 = 32768 units in one plane.  The
    three CG forms and the FLAT instantiations agree with one another bit for bit, the operator has its
    eigenvector (size-independent checks: the oracle's 23 ghosted arrays would take 12 GB of host memory)."""
    from cajitafluids_b200 import Solver
    from helpers import eigen_tol, rel_l2
    n = 8192
    cfg = make_cfg(2, n, fixed_iters=20)
    g = Solver(cfg)
    ax = np.cos(np.pi * 2 * (np.arange(n) + 0.5) / n)
    p = ax[:, None] * ax[None, :]
    g.set(K.CG_P, p.reshape(g.shape(K.CG_P)))
    g.stencil_dot(1)
    h, dt, _ = g.scalars()
    lam = dt / (cfg.density * h * h) * 2 * (2 - 2 * np.cos(np.pi * 2 / n))
    assert rel_l2(g.get(K.CG_Q).reshape(n, n), lam * p) < eigen_tol(n, 2)
    del p
    g.fill_synthetic_velocity(0)
    ref = None
    for flat in (0, 1):
        g.set_tuning("flat_2d", flat)
        for variant in (1, 2, 0):
            g.set_tuning("cg_variant", variant)
            g.build_rhs()
            it, res = g.pcg_solve()
            x = g.get(K.PRESSURE)
            assert it == 20 and np.isfinite(res) and res > 0
            if ref is None:
                ref = (res, x)
            else:
                assert res == ref[0] and np.array_equal(x, ref[1]), (flat, variant)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cells,walls", [((150, 47), None), ((700, 520), [K.FREE, K.SOLID, K.SOLID, K.FREE]), ((1500, 900), None)])
def test_cuda_two_dimensional_units_march_along_y(cells, walls):
    """Two-dimensional runs: a unit of phase B is a run of "fused_yc" tile rows travelling through the TMA ring
    (fused_unit_flat): stage reuse and barrier parities over runs shorter and longer than the ring, in the
    one-unit-per-block kernel and in the persistent kernel (ring carried from run to run), against the oracle."""
    from cajitafluids_b200 import Solver
    cfg = make_cfg(2, cells, box=box_of(cells), fixed_iters=8, **(dict(boundary_type=walls) if walls else {}))
    g, o = Solver(cfg), Oracle(cfg)
    g.fill_synthetic_velocity(1, 3)
    for f in fields_of(2)[1:]:
        o.set(f, g.get(f))
    o.build_rhs()
    ro = o.pcg_solve()
    po = o.get(K.PRESSURE)

    def check(tag):
        g.build_rhs()
        assert g.pcg_solve() == ro, tag
        assert np.array_equal(g.get(K.PRESSURE), po), tag

    check("library's choices")
    g.set_tuning("cg_persist", 0)
    for tx, ty, st in [(64, 16, 3), (64, 8, 4), (128, 16, 3), (64, 32, 2), (128, 8, 3)]:
        for k, v in (("fused_stages", st), ("fused_ty", ty), ("fused_tx", tx)):
            g.set_tuning(k, v)
        for yc in (1, 2, 3, 7, 64):
            g.set_tuning("fused_yc", yc)
            for variant in (1, 2):
                g.set_tuning("cg_variant", variant)
                check((tx, ty, st, yc, variant))
    g.set_tuning("cg_variant", 1)
    for tx, ty, st in [(64, 16, 3), (64, 8, 4)]:
        for k, v in (("fused_stages", st), ("fused_ty", ty), ("fused_tx", tx), ("cg_persist", 1)):
            g.set_tuning(k, v)
        for yc in (1, 2, 5):
            g.set_tuning("fused_yc", yc)
            check(("persistent", tx, ty, st, yc))
    g.close()
