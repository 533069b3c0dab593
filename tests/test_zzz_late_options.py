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
