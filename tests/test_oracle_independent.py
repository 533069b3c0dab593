"""Cross-check of the C++ oracle against an independent numpy/scipy restatement of the reference's
2-D time step (tests/ref2d_numpy.py: written from the cajitafluids sources in a different style —
whole-array numpy, explicitly assembled sparse matrix, numpy dot products).

The two evaluate every sum in a different order, so the bar is the north_star tolerance, not bit
identity: CG iteration counts within +-1 per solve, fields within 1e-9 relative L2 after several
steps (the CG stops at |r| <= 1e-6, SURVEY H1: round-off level differences in the dot products
move the iterates by ~1e-11 relative).
"""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import make_cfg, rel_l2
from oracle_api import Oracle
from ref2d_numpy import FREE, SOLID, Ref2D

NAMES = {K.QUANTITY: "q", K.U: "u", K.V: "v", K.PRESSURE: "p"}


def compare(o, m, tol, what):
    for f, name in NAMES.items():
        a, b = o.get(f)[0], m.owned(name)
        assert a.shape == b.shape
        e = rel_l2(a, b)
        assert e < tol, f"{what}: field {name} rel l2 {e}"


@pytest.mark.parametrize("n", [32, 48])
@pytest.mark.parametrize("quirks", [(1, 1), (0, 0), (0, 1)])
def test_full_steps_2d_reference_defaults(n, quirks):
    steps = 6
    o = Oracle(make_cfg(2, n, quirks=quirks))
    m = Ref2D(n, quirk_q1=bool(quirks[0]), quirk_q2=bool(quirks[1]))
    assert o.dt == m.dt and o.cell_size == m.h
    o.setup()
    m.setup()
    assert abs(o.stats()["cg_iterations"] - m.cg_iterations) <= 1
    compare(o, m, 1e-9, "setup")
    for s in range(steps):
        io = o.stats()["cg_iterations"]
        o.step()
        m.step()
        assert abs((o.stats()["cg_iterations"] - io) - m.last_iters) <= 1, f"step {s}"
    compare(o, m, 1e-9, f"{steps} steps")
    assert abs(o.time - m.time) < 1e-15


def test_body_force_dt_clamp_and_free_walls():
    """gravity (dt clamp src/Solver.hpp:96-106, BodyForce) and FREE walls (Dirichlet rows of the matrix)"""
    n = 40
    bt = [FREE, SOLID, SOLID, FREE]  # [-x, -y, +x, +y]
    cfg = make_cfg(2, n, boundary_type=bt, body_force=(0.0, -9.8, 0.0), dt=0.05)
    o = Oracle(cfg)
    m = Ref2D(n, dt=0.05, force=(0.0, -9.8), btype=tuple(bt))
    assert o.dt == m.dt and m.dt < 0.05
    o.setup()
    m.setup()
    for _ in range(3):
        o.step()
        m.step()
    assert abs(o.stats()["cg_iterations"] - m.cg_iterations) <= 4
    compare(o, m, 1e-9, "gravity + free walls")


def test_linear_field_interpolation_order_1():
    n = 32
    o = Oracle(make_cfg(2, n, interp_order=1))
    m = Ref2D(n)
    o.setup()
    m.setup()
    for _ in range(4):
        o.step()
        m.step(order=1)
    compare(o, m, 1e-9, "order-1 advection")


def test_matrix_apply_against_assembled_sparse_matrix():
    n = 24
    o = Oracle(make_cfg(2, n, boundary_type=[SOLID, FREE, FREE, SOLID]))
    m = Ref2D(n, btype=(SOLID, FREE, FREE, SOLID))
    rng = np.random.default_rng(5)
    p = rng.uniform(-1, 1, size=(n, n))           # [j, i]
    o.set(K.CG_P, p[None])
    o.stencil_dot(1)
    q = o.get(K.CG_Q)[0]
    want = (m.A @ p.T.ravel()).reshape(n, n).T    # the model is [i, j]
    assert rel_l2(q, want) < 1e-14
