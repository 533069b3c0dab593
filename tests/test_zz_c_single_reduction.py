"""Opt-in single-reduction form of the Jacobi-PCG (`cg_variant` 3): Chronopoulos-Gear's
rearrangement of the recurrences of Cajita::ReferenceConjugateGradient::solve (driven from
src/VelocityCorrector.hpp:276) — the same iterates in exact arithmetic, ONE reduction point and one ghost exchange per
iteration (csrc/kernels_cg1.cu, kernels_stencil.cu MODE 2).  On the GPU against the checker's statement of the same
algorithm (bit for bit) and against the reference's loop (the stated bar: iterations +-1, fields <= 1e-10).
CPU: tests/test_emulated_kernels.py::test_emulated_single_reduction_cg (the kernels themselves on the host)."""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import fields_of, make_cfg, rel_l2
from oracle_api import Oracle


def box_of(cells):
    return 1.0 if isinstance(cells, int) else tuple(c / cells[0] for c in cells)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,kw", [(3, 64, {}), (3, (70, 50, 21), dict(boundary_type=[K.SOLID, K.FREE, K.SOLID, K.SOLID, K.SOLID, K.FREE])),
                                          (3, (130, 36, 5), {}), (2, (150, 90), {}), (2, 128, dict(boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE]))])
def test_cuda_single_reduction_cg_against_both_statements(dim, cells, kw):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(dim, cells, box=box_of(cells), **kw)
    g, o1, o0 = Solver(cfg), Oracle(cfg), Oracle(cfg)
    g.set_tuning("cg_variant", 3)
    o1.set_cg_algorithm(1)
    rng = np.random.default_rng(41)
    vel = {f: rng.uniform(-1, 1, size=g.shape(f)) for f in fields_of(dim)[1:]}
    for s in (g, o1, o0):
        for f, a in vel.items():
            s.set(f, a)
        s.add_inputs()
        s.build_rhs()
    rg, r1, r0 = g.pcg_solve(), o1.pcg_solve(), o0.pcg_solve()
    assert abs(rg[0] - r1[0]) <= 1 and rel_l2(g.get(K.PRESSURE), o1.get(K.PRESSURE)) < 1e-10
    assert rg == r1 and np.array_equal(g.get(K.PRESSURE), o1.get(K.PRESSURE))
    assert np.array_equal(g.residual_history(), o1.residual_history())
    assert abs(rg[0] - r0[0]) <= 1 and rel_l2(g.get(K.PRESSURE), o0.get(K.PRESSURE)) < 1e-10
    # whole steps of the default problem
    g2, p1, p0 = Solver(cfg), Oracle(cfg), Oracle(cfg)
    g2.set_tuning("cg_variant", 3)
    p1.set_cg_algorithm(1)
    for s in (g2, p1, p0):
        s.setup()
        for _ in range(2):
            s.step()
    assert g2.stats()["cg_iterations"] == p1.stats()["cg_iterations"]
    assert abs(g2.stats()["cg_iterations"] - p0.stats()["cg_iterations"]) <= 3
    for f in fields_of(dim) + [K.PRESSURE]:
        assert np.array_equal(g2.get(f), p1.get(f)), f
        assert rel_l2(g2.get(f), p0.get(f)) < 1e-10, f
    for s in (g, g2):
        s.close()


@pytest.mark.gpu
def test_cuda_single_reduction_cg_fixed_iterations_and_form_switching():
    from cajitafluids_b200 import Solver
    cfg = make_cfg(3, (96, 64, 40), box=box_of((96, 64, 40)), fixed_iters=25)
    g, o = Solver(cfg), Oracle(cfg)
    rng = np.random.default_rng(42)
    vel = {f: rng.uniform(-1, 1, size=g.shape(f)) for f in fields_of(3)[1:]}
    for variant, alg in ((3, 1), (1, 0), (2, 0), (3, 1), (0, 0)):
        g.set_tuning("cg_variant", variant)
        o.set_cg_algorithm(alg)
        for s in (g, o):
            for f, a in vel.items():
                s.set(f, a)
            s.add_inputs()
            s.build_rhs()
        assert g.pcg_solve() == o.pcg_solve(), variant
        assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE)), variant
    g.close()
