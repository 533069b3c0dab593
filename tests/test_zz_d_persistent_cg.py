"""The persistent form of the two-kernel CG iteration (`cg_persist`; csrc/kernels_fused.cu: cg_persistent_kernel): batches of
iterations of Cajita::ReferenceConjugateGradient::solve (driven from src/VelocityCorrector.hpp:276) in ONE cooperative
launch — phase A, grid barrier, phase B over the block's units, grid barrier — which is what blocks whose CG vectors
stay in the L2 run by default.  On the GPU this is the first place where several blocks meet at the grid barrier and
where the TMA stage barriers are carried from unit to unit; bit for bit against the oracle and the launch-per-phase form.
CPU: tests/test_emulated_kernels.py::test_emulated_persistent_cg_iterations (one block of fibers)."""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import fields_of, make_cfg
from oracle_api import Oracle


def box_of(cells):
    return 1.0 if isinstance(cells, int) else tuple(c / cells[0] for c in cells)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,kw", [(3, 64, {}), (3, (70, 50, 21), dict(boundary_type=[K.SOLID, K.FREE, K.SOLID, K.SOLID, K.SOLID, K.FREE])),
                                          (3, (130, 36, 5), {}), (3, (128, 96, 80), {}), (2, (150, 90), {}),
                                          (2, 256, dict(boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE]))])
def test_cuda_persistent_iterations_against_the_oracle(dim, cells, kw):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(dim, cells, box=box_of(cells), **kw)
    g, o = Solver(cfg), Oracle(cfg)
    g.set_tuning("cg_persist", 1)
    rng = np.random.default_rng(61)
    vel = {f: rng.uniform(-1, 1, size=g.shape(f)) for f in fields_of(dim)[1:]}
    for s in (g, o):
        for f, a in vel.items():
            s.set(f, a)
        s.add_inputs()
        s.build_rhs()
    l0 = g.stats()["kernel_launches"]
    res = g.pcg_solve()
    launches = g.stats()["kernel_launches"] - l0
    assert res == o.pcg_solve()
    assert launches < 12 + res[0] // 16, (launches, res)
    assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE)) and np.array_equal(g.residual_history(), o.residual_history())
    g.close()
    cfg2 = make_cfg(dim, cells, box=box_of(cells), fixed_iters=75, **kw)
    g2, o2 = Solver(cfg2), Oracle(cfg2)
    for persist in (1, 0, 1, 1):
        g2.set_tuning("cg_persist", persist)
        for s in (g2, o2):
            for f, a in vel.items():
                s.set(f, a)
            s.add_inputs()
            s.build_rhs()
        assert g2.pcg_solve() == o2.pcg_solve(), persist
        assert np.array_equal(g2.get(K.PRESSURE), o2.get(K.PRESSURE)), persist
        assert np.array_equal(g2.get(K.CG_R), o2.get(K.CG_R)), persist
    g2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells", [(3, 64), (2, 128)])
def test_cuda_persistent_iterations_whole_steps(dim, cells):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(dim, cells)
    g, o = Solver(cfg), Oracle(cfg)
    g.set_tuning("cg_persist", 1)
    for s in (g, o):
        s.setup()
        for _ in range(3):
            s.step()
    assert g.stats()["cg_iterations"] == o.stats()["cg_iterations"]
    for f in fields_of(dim) + [K.PRESSURE]:
        assert np.array_equal(g.get(f), o.get(f)), f
    g.close()
