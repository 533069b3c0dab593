"""The 64-byte CG iteration (`cg_variant` 2: q = A p is never stored; phase A' — the TMA stencil kernel in a second
mode — recomputes it while it updates r, src/VelocityCorrector.hpp:124-143 + Cajita's ReferenceConjugateGradient) on
the GPU against the other forms and the oracle, bit for bit.  CPU: tests/test_emulated_kernels.py (the kernels
themselves on the host).  (Sorts late: written after this round's GPU budget was spent.)"""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import fields_of, make_cfg, rel_l2, smooth_velocity
from oracle_api import Oracle

ALL = lambda dim: fields_of(dim) + [K.PRESSURE]  # noqa: E731


def run(ctx, steps):
    ctx.setup()
    its = [ctx.stats()["cg_iterations"]]
    for _ in range(steps):
        ctx.step()
        its.append(ctx.stats()["cg_iterations"])
    return np.diff([0] + its)


# ---------------------------------------------------------------------------------------------
# cg_variant 2 (64 B/cell: q never stored, phase A' recomputes A p) — written after the round's GPU budget was
# spent, bit-exact against the oracle in the host emulation (TMA kernels themselves); first GPU run here
@pytest.mark.gpu
@pytest.mark.parametrize("cells,walls", [((70, 50, 21), "solid"), ((130, 36, 5), "mixed"), ((33, 47), "solid"),
                                         ((64, 64, 64), "solid")])
def test_cuda_64_byte_iteration_matches_the_other_forms_and_the_oracle(cells, walls):
    from cajitafluids_b200 import Solver
    dim = len(cells)
    bt = None
    if walls == "mixed":
        bt = [K.SOLID, K.FREE, K.SOLID, K.SOLID, K.SOLID, K.FREE]
    box = tuple(c / cells[0] for c in cells)
    for fixed in (0, 9):
        cfg = make_cfg(dim, cells, box=box, fixed_iters=fixed, **(dict(boundary_type=bt) if bt else {}))
        o = Oracle(cfg)
        rng = np.random.default_rng(6)
        vel = smooth_velocity(o, rng)
        for f, a in vel.items():
            o.set(f, a)
        o.add_inputs()
        o.build_rhs()
        ro = o.pcg_solve()
        for variant in (2, 1):
            g = Solver(cfg)
            g.set_tuning("cg_variant", variant)
            for f, a in vel.items():
                g.set(f, a)
            g.add_inputs()
            g.build_rhs()
            rg = g.pcg_solve()
            assert abs(rg[0] - ro[0]) <= 1, (variant, rg, ro)
            assert rel_l2(g.get(K.PRESSURE), o.get(K.PRESSURE)) < 1e-10, variant
            assert rg == ro and np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE)), (variant, rg, ro)
            assert np.array_equal(g.residual_history(), o.residual_history())
            g.close()


@pytest.mark.gpu
def test_cuda_64_byte_iteration_whole_steps():
    from cajitafluids_b200 import Solver
    for dim, n in ((2, 64), (3, 48)):
        cfg = make_cfg(dim, n)
        g, o = Solver(cfg), Oracle(cfg)
        g.set_tuning("cg_variant", 2)
        assert list(run(g, 3)) == list(run(o, 3))
        for f in ALL(dim):
            assert np.array_equal(g.get(f), o.get(f)), f


@pytest.mark.gpu
@pytest.mark.parametrize("cells", [(70, 50, 21), (130, 36, 70), (150, 47)])
def test_cuda_phase_a_prime_tilings_with_staged_r(cells):
    """Phase A' on every stencil tiling, the tile of r travelling through the TMA ring behind the plane of p: the
    asynchronous side of the ring — stage reuse, one mbarrier for the two boxes of a slot, chunks shorter and longer
    than the ring — which the host emulation cannot see."""
    from cajitafluids_b200 import Solver
    dim = len(cells)
    cfg = make_cfg(dim, cells, box=tuple(c / cells[0] for c in cells), fixed_iters=6)
    g, o = Solver(cfg), Oracle(cfg)
    g.fill_synthetic_velocity(1, 7)
    for f in fields_of(dim)[1:]:
        o.set(f, g.get(f))
    o.build_rhs()
    ro = o.pcg_solve()
    po, rr = o.get(K.PRESSURE), o.get(K.CG_R)
    g.set_tuning("cg_variant", 2)
    tilings = [(64, 16, 4), (64, 16, 6), (64, 8, 4), (64, 32, 4), (64, 32, 3), (128, 16, 4), (128, 16, 3), (128, 32, 3),
               (128, 8, 4)]
    for rtma in (1,):  # (the form that streamed r with 128-bit loads was measured against this one and removed)
        for tx, ty, st in tilings:
            for zc in ((2, 5, 64) if dim == 3 else (64,)):
                for k, v in (("stencil_stages", st), ("stencil_ty", ty), ("stencil_tx", tx), ("stencil_zc", zc)):
                    g.set_tuning(k, v)
                g.build_rhs()
                assert g.pcg_solve() == ro, (rtma, tx, ty, st, zc)
                assert np.array_equal(g.get(K.PRESSURE), po) and np.array_equal(g.get(K.CG_R), rr), (rtma, tx, ty, st, zc)
    g.close()


@pytest.mark.gpu
def test_cuda_phase_b_with_sixteen_warps():
    """"fused_nt" 512: the 128 x 16 x 3 tiling of phase B with two rows per thread instead of four — same values."""
    from cajitafluids_b200 import Solver
    cells = (300, 50, 21)
    cfg = make_cfg(3, cells, box=tuple(c / cells[0] for c in cells), fixed_iters=6)
    g, o = Solver(cfg), Oracle(cfg)
    g.fill_synthetic_velocity(1, 7)
    for f in fields_of(3)[1:]:
        o.set(f, g.get(f))
    o.build_rhs()
    ro = o.pcg_solve()
    po = o.get(K.PRESSURE)
    for k, v in (("fused_stages", 3), ("fused_ty", 16), ("fused_tx", 128), ("fused_zc", 8), ("fused_nt", 512)):
        g.set_tuning(k, v)
    for variant in (1, 2):
        g.set_tuning("cg_variant", variant)
        g.build_rhs()
        assert g.pcg_solve() == ro, variant
        assert np.array_equal(g.get(K.PRESSURE), po), variant
    g.close()
