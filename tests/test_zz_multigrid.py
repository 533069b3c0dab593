"""Opt-in geometric multigrid preconditioner of the pressure CG (the role HYPRE
PFMG plays in the reference's default path).  It is this library's own algorithm, never the default:
north_star pins the reference's Jacobi preconditioner for the parity runs.

CPU: the checker's statement of the algorithm (oracle/cfo_oracle.cpp: mg_*) — same solution as the
Jacobi path to the solver tolerance (north_star's bar on the fields), iteration counts that do not grow
with n, symmetry / positivity of the V-cycle as an operator (what CG needs), error paths.
GPU: the CUDA statement (cajitafluids_b200/csrc/mg.cu) against the checker, bit for bit.

(Sorts last on purpose: the CUDA side was written after this round's GPU budget was spent.)
"""
import numpy as np
import pytest

from cajitafluids_b200 import CfbError, config as K
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


CASES = [
    (2, 64, {}),
    (2, (48, 40), {}),                                                    # 3 levels, coarsest 12 x 10
    (2, 32, dict(boundary_type=[K.FREE, K.SOLID, K.SOLID, K.SOLID])),     # non-singular operator: no pinning
    (2, 40, dict(boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE], body_force=(0.0, -9.8, 0.0))),
    (3, 32, {}),
    (3, (24, 20, 16), {}),
    (3, 16, dict(boundary_type=[K.SOLID, K.SOLID, K.FREE, K.SOLID, K.SOLID, K.SOLID])),
    (3, 18, {}),                                                          # 2 levels only (9 is odd)
]


def cfg_of(dim, cells, kw):
    box = 1.0 if isinstance(cells, int) else tuple(c / cells[0] for c in cells)
    return make_cfg(dim, cells, box=box, **kw)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,cells,kw", CASES)
def test_mg_pcg_gives_the_jacobi_pcg_solution_in_far_fewer_iterations(dim, cells, kw):
    jac, mg = Oracle(cfg_of(dim, cells, kw)), Oracle(cfg_of(dim, cells, kw))
    mg.set_preconditioner("mg")
    ij, im = run(jac, 3), run(mg, 3)
    assert np.all(im * 4 <= ij), (ij, im)
    # both stop at sqrt(sum r^2) <= 1e-6: the fields agree to the solver tolerance, including the
    # null-space component of the pressure (quirk Q1 leaks it into v in 2-D)
    for f in ALL(dim):
        a, b = mg.get(f), jac.get(f)
        assert rel_l2(a, b) < (1e-5 if f == K.PRESSURE else 1e-6), (f, rel_l2(a, b))
    assert mg.time == jac.time


def test_mg_pcg_tight_tolerance_meets_the_field_bar():
    """With the solves converged well below the bar, the two preconditioners give the same step to
    north_star's 1e-10."""
    for dim, n in ((2, 48), (3, 24)):
        jac, mg = (Oracle(make_cfg(dim, n, tol=1e-11, max_iter=20000)) for _ in range(2))
        mg.set_preconditioner("mg")
        run(jac, 2), run(mg, 2)
        for f in ALL(dim):
            assert rel_l2(mg.get(f), jac.get(f)) < 1e-10, f


def test_mg_iteration_count_does_not_grow_like_n():
    its = {}
    for n in (32, 64):
        o = Oracle(make_cfg(3, n))
        o.set_preconditioner("mg")
        its[n] = int(run(o, 0)[0])
    assert 0 < its[32] and its[64] <= its[32] + 4 and its[64] <= 16, its
    o = Oracle(make_cfg(2, 256))
    o.set_preconditioner("mg")
    assert int(run(o, 0)[0]) <= 12


def vcycle_as_matrix(o, n_cells):
    """Columns of M^-1: the V-cycle applied to the unit vectors."""
    shp = o.shape(K.RHS)
    cols = []
    for e in range(n_cells):
        b = np.zeros(n_cells)
        b[e] = 1.0
        cols.append(o.mg_apply(b.reshape(shp)).ravel().copy())
    return np.array(cols).T


@pytest.mark.parametrize("dim,n,bt", [(2, 8, None), (3, 4, None), (2, 8, [K.FREE, K.SOLID, K.SOLID, K.SOLID])])
def test_vcycle_is_a_symmetric_positive_operator(dim, n, bt):
    kw = dict(boundary_type=bt) if bt else {}
    o = Oracle(make_cfg(dim, n, **kw))
    o.set_preconditioner("mg", 2, 2, 4)
    M = vcycle_as_matrix(o, n ** dim)
    assert np.abs(M - M.T).max() <= 1e-13 * np.abs(M).max()
    w = np.linalg.eigvalsh(0.5 * (M + M.T))
    assert w.min() > 0.0, w.min()


def test_mg_parameters_and_errors():
    o = Oracle(make_cfg(2, 32))
    for bad in ((0, 2, 8, 0.0), (2, -1, 8, 0.0), (2, 2, 0, 0.0), (2, 2, 8, 2.5)):
        with pytest.raises(CfbError):
            o.set_preconditioner("mg", *bad)
    with pytest.raises(CfbError):
        o.set_preconditioner(7)
    o.set_preconditioner("mg", 1, 1, 4, 0.7)
    i11 = int(run(o, 0)[0])
    o2 = Oracle(make_cfg(2, 32))
    o2.set_preconditioner("mg", 3, 3, 4, 0.7)
    assert int(run(o2, 0)[0]) < i11
    # back to the reference's preconditioner: the reference's iteration count
    o3, o4 = Oracle(make_cfg(2, 32)), Oracle(make_cfg(2, 32))
    o3.set_preconditioner("mg")
    o3.set_preconditioner("jacobi")
    assert list(run(o3, 1)) == list(run(o4, 1))
    assert np.array_equal(o3.get(K.PRESSURE), o4.get(K.PRESSURE))


def test_mg_refuses_several_blocks():
    cfg = make_cfg(3, 16)
    cfg.ranks_per_dim[2], cfg.world_size = 2, 2
    o = Oracle(cfg)
    with pytest.raises(CfbError):
        o.set_preconditioner("mg")


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,kw", CASES)
def test_cuda_mg_pcg_matches_the_checker_bit_for_bit(dim, cells, kw):
    from cajitafluids_b200 import Solver
    g, o = Solver(cfg_of(dim, cells, kw)), Oracle(cfg_of(dim, cells, kw))
    g.set_preconditioner("mg")
    o.set_preconditioner("mg")
    ig, io = run(g, 3), run(o, 3)
    assert np.all(np.abs(ig - io) <= 1), (ig, io)
    for f in ALL(dim):
        assert rel_l2(g.get(f), o.get(f)) < 1e-10, f
    assert list(ig) == list(io)
    for f in ALL(dim):
        assert np.array_equal(g.get(f), o.get(f)), f
    hg, ho = g.residual_history(), o.residual_history()
    assert np.array_equal(hg, ho)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells", [(2, (32, 16)), (2, (12, 20)), (3, (16, 8, 8)), (3, (12, 20, 8)), (3, (6, 6, 6))])
def test_cuda_vcycle_alone_matches_the_checker_bit_for_bit(dim, cells):
    from cajitafluids_b200 import Solver
    cfg = cfg_of(dim, cells, {})
    g, o = Solver(cfg), Oracle(cfg)
    rng = np.random.default_rng(9)
    for nu in ((1, 0, 1), (1, 1, 2), (2, 2, 8), (3, 2, 4)):
        for s in (g, o):
            s.set_preconditioner("mg", *nu)
        r = rng.standard_normal(o.shape(K.PRESSURE))
        zg, zo = g.mg_apply(r), o.mg_apply(r)
        assert rel_l2(zg, zo) < 1e-13, (nu, rel_l2(zg, zo))
        assert np.array_equal(zg, zo), nu


@pytest.mark.gpu
@pytest.mark.parametrize("nu,fixed", [((1, 1, 3), 0), ((2, 2, 8), 0), ((3, 3, 5), 0), ((3, 1, 5), 6)])
def test_cuda_mg_single_solve_on_seeded_rhs(nu, fixed):
    """(V(3,1) is not a symmetric operator, hence no preconditioner for CG to converge with: a fixed number of
    iterations there, for the code path.)"""
    from cajitafluids_b200 import Solver
    for dim, cells in ((2, (40, 24)), (3, (20, 12, 16))):
        cfg = cfg_of(dim, cells, dict(fixed_iters=fixed))
        g, o = Solver(cfg), Oracle(cfg)
        rng = np.random.default_rng(5)
        # all walls SOLID: the system is singular, so the velocity must vanish on the walls of THIS box
        # (edge lengths cells / cells[0]) for the right-hand side to be compatible
        for f, a in smooth_velocity(o, rng, extent=[c / cells[0] for c in cells]).items():
            g.set(f, a)
            o.set(f, a)
        for s in (g, o):
            s.set_preconditioner("mg", *nu, 0.75)
            s.build_rhs()
        (ig, rg), (io, ro) = g.pcg_solve(), o.pcg_solve()
        assert ig == io and rg == ro, (ig, io, rg, ro)
        assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE))
        assert np.array_equal(g.get(K.CG_R), o.get(K.CG_R))


@pytest.mark.gpu
def test_cuda_mg_fixed_iterations_and_back_to_jacobi():
    from cajitafluids_b200 import Solver
    cfg = make_cfg(3, 32, fixed_iters=5)
    g, o = Solver(cfg), Oracle(cfg)
    for s in (g, o):
        s.set_preconditioner("mg")
        s.add_inputs()
        s.build_rhs()
    assert g.pcg_solve() == o.pcg_solve()
    assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE))
    for s in (g, o):
        s.set_preconditioner("jacobi")
        s.build_rhs()  # lhs = 0 (src/VelocityCorrector.hpp:272): the checker's solve starts from the x it is given
    assert g.pcg_solve() == o.pcg_solve()
    assert np.array_equal(g.get(K.PRESSURE), o.get(K.PRESSURE))


@pytest.mark.gpu
def test_cuda_mg_cycle_replayed_as_a_cuda_graph():
    """"mg_graph" tuning key: same results, one graph launch per V-cycle instead of ~60 kernel launches."""
    import time
    from cajitafluids_b200 import Solver
    cfg = make_cfg(3, 64)
    o = Oracle(cfg)
    o.set_preconditioner("mg")
    io = run(o, 2)
    t = {}
    for graph, coarse in ((0, 0), (1, 0), (0, 1), (1, 1)):
        g = Solver(cfg)
        g.set_tuning("mg_graph", graph)
        g.set_tuning("mg_coarse_kernel", coarse)  # levels <= 16^3 down and back up in one single-CTA kernel
        g.set_preconditioner("mg")
        t0 = time.perf_counter()
        assert list(run(g, 2)) == list(io), (graph, coarse)
        t[(graph, coarse)] = time.perf_counter() - t0
        for f in ALL(3):
            assert np.array_equal(g.get(f), o.get(f)), (graph, coarse, f)
        r = np.random.default_rng(1).standard_normal(o.shape(K.PRESSURE))
        assert np.array_equal(g.mg_apply(r), o.mg_apply(r)), (graph, coarse)
        g.close()
    print("64^3, setup + 2 steps with MG, seconds by (graph, coarse kernel):", t)


@pytest.mark.gpu
def test_cuda_mg_beats_jacobi_at_128_cubed():
    """Time to solution of one projection at 128^3 with the reference's tolerance: the number that the
    preconditioner is there for (reported; required to be at least 1.5x better — at this size the V-cycle's ~60
    launches per iteration are launch-bound until it is captured in a CUDA graph)."""
    import time
    from cajitafluids_b200 import Solver
    t = {}
    for kind in ("jacobi", "mg"):
        g = Solver(make_cfg(3, 128))
        g.set_preconditioner(kind)
        g.add_inputs()
        g.build_rhs()
        g.pcg_solve()  # warm
        t0 = time.perf_counter()
        it, res = g.pcg_solve()
        t[kind] = (time.perf_counter() - t0, it, res)
        g.close()
    print("128^3 projection:", t)
    assert t["mg"][1] * 10 < t["jacobi"][1] and t["mg"][0] * 1.5 < t["jacobi"][0], t


@pytest.mark.gpu
@pytest.mark.parametrize("cells,walls", [((72, 36, 24), None), ((130, 20, 12), [K.SOLID, K.FREE, K.SOLID, K.SOLID, K.FREE, K.SOLID]),
                                         ((192, 160, 96), None)])
def test_cuda_mg_fine_level_sweeps_on_the_tma_march(cells, walls):
    """"mg_tma" (default on, 3-D): the fine level's smoothing sweeps on the TMA z-march (kernels_stencil.cu MODE 4 / 3)
    against the one-thread-per-cell kernels and the checker, every sweep-count combination; then a whole step."""
    from cajitafluids_b200 import Solver
    cfg = make_cfg(3, cells, box=tuple(c / cells[0] for c in cells), **(dict(boundary_type=walls) if walls else {}))
    g, o = Solver(cfg), Oracle(cfg)
    rng = np.random.default_rng(3)
    big = cells[0] * cells[1] * cells[2] > 1_000_000
    for nu in ((2, 2, 8, 0.0),) if big else ((2, 2, 8, 0.0), (1, 1, 2, 0.0), (3, 2, 4, 0.7), (2, 0, 1, 0.0), (4, 4, 3, 0.0), (1, 3, 2, 0.6)):
        o.set_preconditioner("mg", *nu)
        r = rng.standard_normal(o.shape(K.PRESSURE))
        z = o.mg_apply(r)
        for tma, prolong in ((1, 1), (1, 0), (0, 0)):
            g.set_tuning("mg_tma", tma)
            g.set_tuning("mg_tma_prolong", prolong)  # prolongation + first post-sweep on the march as well (MODE 5)
            g.set_preconditioner("mg", *nu)
            assert np.array_equal(g.mg_apply(r), z), (nu, tma, prolong)
    g.set_tuning("mg_tma", 1)
    g.set_tuning("mg_tma_prolong", 1)
    if not big:
        for s in (g, o):
            s.set_preconditioner("mg")
        assert list(run(g, 1)) == list(run(o, 1))
        for f in ALL(3):
            assert np.array_equal(g.get(f), o.get(f)), f
        assert np.array_equal(g.residual_history(), o.residual_history())
    g.close()
