"""The 3-D oracle against an independent dimension-generic numpy/scipy restatement of the time step
(tests/refnd_numpy.py).  cajitafluids is 2-D only; the benchmark sizes are 3-D, and nothing else holds the oracle's
3-D extension (FaceK / w, 7-point operator, 64-point splines).

The generic model is pinned first: in 2-D it must agree with the 2-D model written from the reference
(tests/ref2d_numpy.py) to round-off.  Then the oracle is held against it in 3-D at the north_star tolerance (the two
evaluate every sum in a different order; CG stops at |r| <= 1e-6): iteration counts within +-1 per solve, fields
within 1e-9 relative L2 after several steps."""
import numpy as np
import pytest

from cajitafluids_b200 import config as K
from helpers import make_cfg, rel_l2
from oracle_api import Oracle
from ref2d_numpy import Ref2D
from refnd_numpy import FREE, SOLID, RefND

NAMES = {K.QUANTITY: "q", K.U: "u", K.V: "v", K.W: "w", K.PRESSURE: "p"}


@pytest.mark.parametrize("kw", [{}, dict(quirk_q1=False, quirk_q2=False),
                                dict(force=(0.0, -9.8), btype=(FREE, SOLID, SOLID, FREE), dt=0.05)])
def test_generic_model_agrees_with_the_2d_model(kw):
    a, b = RefND((32, 32), **kw), Ref2D(32, **kw)
    assert a.dt == b.dt and a.h == b.h
    assert abs(a.A - b.A).max() == 0.0
    a.setup()
    b.setup()
    for _ in range(4):
        a.step()
        b.step()
    assert abs(a.cg_iterations - b.cg_iterations) <= 2
    for name in "quvp":
        assert rel_l2(a.owned(name), b.owned(name)) < 1e-10, name


def compare(o, m, tol, what):
    for f, name in NAMES.items():
        a, b = o.get(f), m.owned(name)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        e = rel_l2(a, b)
        assert e < tol, f"{what}: field {name} rel l2 {e}"


@pytest.mark.parametrize("cells,quirks,order", [((32, 32, 32), (1, 1), 3), ((32, 32, 32), (0, 0), 3), ((32, 24, 16), (1, 1), 3),
                                                ((32, 32, 32), (1, 1), 1)])
def test_full_steps_3d(cells, quirks, order):
    steps = 4
    box = tuple(c / cells[0] for c in cells)
    o = Oracle(make_cfg(3, cells, box=box, quirks=quirks, interp_order=order))
    m = RefND(cells, quirk_q1=bool(quirks[0]), quirk_q2=bool(quirks[1]))
    assert o.dt == m.dt and o.cell_size == m.h
    o.setup()
    m.setup()
    assert abs(o.stats()["cg_iterations"] - m.cg_iterations) <= 1
    compare(o, m, 1e-9, "setup")
    for s in range(steps):
        io = o.stats()["cg_iterations"]
        o.step()
        m.step(order=order)
        assert abs((o.stats()["cg_iterations"] - io) - m.last_iters) <= 1, f"step {s}"
    compare(o, m, 1e-9, f"{steps} steps")
    assert m.owned("q").max() > 1.0 and abs(m.owned("w")).max() > 0.0  # something happened, in z too
    assert abs(o.time - m.time) < 1e-15


@pytest.mark.parametrize("quirks", [None, (1, 1)])
def test_3d_body_force_dt_clamp_and_free_walls(quirks):
    """gravity (dt clamp, BodyForce), FREE walls (Dirichlet rows), and the library's 3-D defaults (quirks None: Q1 —
    the FaceJ kernel's boundary call on u, a 2-D slip of the reference — off, Q2 on; cajitafluids_b200/config.py)."""
    cells = (32, 32, 32)
    bt = [FREE, SOLID, SOLID, SOLID, FREE, SOLID]  # [-x, -y, -z, +x, +y, +z]
    cfg = make_cfg(3, cells, boundary_type=bt, body_force=(0.0, -9.8, 2.0), dt=0.05, quirks=quirks)
    o = Oracle(cfg)
    m = RefND(cells, dt=0.05, force=(0.0, -9.8, 2.0), btype=bt, quirk_q1=quirks is not None, quirk_q2=True)
    assert o.dt == m.dt and m.dt < 0.05
    o.setup()
    m.setup()
    for _ in range(3):
        o.step()
        m.step()
    assert abs(o.stats()["cg_iterations"] - m.cg_iterations) <= 4
    compare(o, m, 1e-9, "gravity + free walls")


def test_3d_matrix_apply_against_assembled_sparse_matrix():
    cells = (12, 10, 8)
    bt = [SOLID, FREE, SOLID, FREE, SOLID, SOLID]
    o = Oracle(make_cfg(3, cells, box=tuple(c / cells[0] for c in cells), boundary_type=bt))
    m = RefND(cells, btype=bt)
    rng = np.random.default_rng(5)
    p = rng.uniform(-1, 1, size=cells[::-1])        # [k, j, i]
    o.set(K.CG_P, p)
    o.stencil_dot(1)
    want = (m.A @ p.T.ravel()).reshape(cells).T     # the model is [i, j, k]
    assert rel_l2(o.get(K.CG_Q), want) < 1e-14


def test_3d_output_stage_against_the_model():
    """src/SiloWriter.hpp:136-197 extended to 3-D: owned quantity + velocity interpolated to the cell centres."""
    cells = (32, 24, 16)
    o = Oracle(make_cfg(3, cells, box=tuple(c / cells[0] for c in cells)))
    m = RefND(cells, quirk_q1=False)
    o.setup()
    m.setup()
    for _ in range(2):
        o.step()
        m.step()
    q, vel, nodes = o.output()
    mq, mvel = m.output()
    assert rel_l2(q, mq) < 1e-9
    vel = np.asarray(vel).reshape((3,) + cells[::-1])
    for d in range(3):
        assert rel_l2(vel[d], mvel[d]) < 1e-9, d
    for d in range(3):
        assert np.allclose(nodes[d], np.arange(cells[d] + 1) / cells[0], rtol=0, atol=1e-15)
