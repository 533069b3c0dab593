"""The output stage that follows the hot path in Solver::solve (src/Solver.hpp:156,170-173; SURVEY.md
§8f rank 2): what SiloWriter::writeFile hands to Silo for a block (src/SiloWriter.hpp:109-186) — node
coordinates, the owned quantity, the MAC velocity interpolated to the cell centres.

CPU: the oracle's restatement against the reference run here (bit for bit; needs /root/reference),
against the reference-run fixtures (travel) and against known answers; the .npy writer of the library.
GPU: the extraction kernel through the C ABI against the oracle and the fixtures, bit for bit; the
asynchronous file writes of cfb_write_output / cfb_solve.

(The file sorts last on purpose: the CUDA side of this stage was written after the round's GPU budget
was spent, so a failure here must not hide the results of the verified tests under `pytest -x`.)
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)

import reference_api  # noqa: E402
from cajitafluids_b200 import config as K  # noqa: E402
from helpers import make_cfg, rel_l2, smooth_velocity  # noqa: E402
from oracle_api import Oracle  # noqa: E402
from refrun_cases import CASES  # noqa: E402
from test_golden_refrun import check as check_fixture  # noqa: E402

needs_ref = pytest.mark.skipif(not reference_api.available(), reason="needs the reference tree (/root/reference)")

# non-dyadic boxes: Cajita's per-dimension cell size differs from Mesh::cellSize() in the last bit
BOXES = [((32, 32), (1.0, 1.0)), ((40, 24), (1.0, 0.6)), ((30, 70), (0.3, 0.7)), ((24, 56), (0.3, 0.7))]


def run(ctx, steps):
    ctx.setup()
    for _ in range(steps):
        ctx.step()
    return ctx


def same_output(a, b):
    (qa, va, na), (qb, vb, nb) = a.output(), b.output()
    assert np.array_equal(qa, qb)
    assert np.array_equal(va, vb)
    assert len(na) == len(nb) and all(np.array_equal(x, y) for x, y in zip(na, nb))


# ---------------------------------------------------------------------------------------------
# CPU: oracle against the reference, fixtures, known answers
@needs_ref
@pytest.mark.parametrize("cells,box", BOXES)
def test_oracle_output_is_what_the_reference_hands_to_silo(cells, box):
    cfg = make_cfg(2, cells, box=box)
    o, r = run(Oracle(cfg), 3), run(reference_api.Reference(cfg), 3)
    same_output(o, r)
    # and the state behind it, per-dimension cell sizes included
    for f in (K.QUANTITY, K.U, K.V, K.PRESSURE):
        assert np.array_equal(o.get(f), r.get(f)), f


@needs_ref
def test_reference_solve_writes_through_silo_at_the_reference_cadence():
    """Solver::solve writes once before setup and after every step t with t % write_freq == 0."""
    cfg = make_cfg(2, 16)
    r = reference_api.Reference(cfg)
    w0 = r.silo_last()[0]
    steps = r.solve(5 * r.dt * 0.999, 2)  # 5 steps: t = 0..4 -> writes at t = 0, 2, 4 (+ the one before setup)
    assert steps == 5
    writes, cycle, time = r.silo_last()[:3]
    assert writes - w0 == 4 and cycle == 4


def test_cell_centred_velocity_known_answers():
    """The y (x) sample of u (v) falls on an entity centre: weight 1 there up to rounding, so the
    cell-centred velocity is the mean of the cell's two faces; a field linear in x is reproduced."""
    for dim, n in ((2, 24), (3, 12)):
        cfg = make_cfg(dim, n)
        o = Oracle(cfg)
        rng = np.random.default_rng(7)
        vel = smooth_velocity(o, rng)
        for f, a in vel.items():
            o.set(f, a)
        q = rng.uniform(-1, 1, size=o.shape(K.QUANTITY))
        o.set(K.QUANTITY, q)
        oq, ov, nodes = o.output()
        assert np.array_equal(oq, q)
        u, v = vel[K.U], vel[K.V]
        assert np.allclose(ov[0], 0.5 * (u[..., :-1] + u[..., 1:]), rtol=0, atol=1e-14)
        assert np.allclose(ov[1], 0.5 * (v[..., :-1, :] + v[..., 1:, :]), rtol=0, atol=1e-14)
        if dim == 3:
            w = vel[K.W]
            assert np.allclose(ov[2], 0.5 * (w[:-1] + w[1:]), rtol=0, atol=1e-14)
        h = o.cell_size
        for d in range(dim):
            assert nodes[d].shape == (n + 1,) and np.allclose(nodes[d], np.arange(n + 1) * h, rtol=0, atol=1e-15)
        # u = 2 + 3 x sampled at the cell centres
        x_face = np.arange(n + 1) * h
        o.set(K.U, np.broadcast_to(2.0 + 3.0 * x_face, o.shape(K.U)).copy())
        ov = o.output()[1]
        xc = (np.arange(n) + 0.5) * h
        assert np.allclose(ov[0], np.broadcast_to(2.0 + 3.0 * xc, ov[0].shape), rtol=0, atol=1e-14)


def test_write_npy_round_trip(tmp_path):
    """Host-only helper of the product library (no GPU needed): numpy reads what it writes."""
    from cajitafluids_b200 import load
    lib = load()
    rng = np.random.default_rng(3)
    for shape in ((7,), (3, 5), (2, 3, 4), (3, 2, 5, 4), (1, 1), (123,)):
        a = rng.standard_normal(shape)
        path = str(tmp_path / ("a%d.npy" % len(shape)))
        shp = (C.c_int64 * len(shape))(*shape)
        lib.check(lib.fn["write_npy"](os.fsencode(path), a.ctypes.data_as(C.POINTER(C.c_double)), len(shape), shp))
        b = np.load(path)
        assert b.dtype == np.float64 and b.shape == shape and np.array_equal(a, b)
        assert os.path.getsize(path) % 8 == 0 and (os.path.getsize(path) - a.nbytes) % 64 == 0
    assert lib.fn["write_npy"](None, None, 0, None) == K.ERR_INVALID


# ---------------------------------------------------------------------------------------------
# GPU: the extraction kernel and the asynchronous writer, through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_output_matches_reference_run_fixture(name):
    from cajitafluids_b200 import Solver
    check_fixture(Solver(CASES[name][0]()), name, with_output=True)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,box", [(2, c, b) for c, b in BOXES] +
                         [(3, (20, 16, 12), (1.0, 0.8, 0.6)), (3, (24, 24, 24), (1.0, 1.0, 1.0)),
                          (3, (16, 10, 18), (0.8, 0.5, 0.9))])
def test_cuda_output_matches_oracle_bit_for_bit(dim, cells, box):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(dim, cells, box=box)
    g, o = run(Solver(cfg), 2), run(Oracle(cfg), 2)
    for f in [K.QUANTITY, K.U, K.V, K.PRESSURE] + ([K.W] if dim == 3 else []):
        assert np.array_equal(g.get(f), o.get(f)), f
    same_output(g, o)


@pytest.mark.gpu
def test_cuda_output_on_seeded_fields():
    from cajitafluids_b200 import Solver
    for dim, n in ((2, 37), (3, 19)):
        cfg = make_cfg(dim, n)
        g, o = Solver(cfg), Oracle(cfg)
        rng = np.random.default_rng(11)
        for f, a in smooth_velocity(o, rng).items():
            g.set(f, a)
            o.set(f, a)
        q = rng.uniform(-1, 1, size=o.shape(K.QUANTITY))
        g.set(K.QUANTITY, q)
        o.set(K.QUANTITY, q)
        same_output(g, o)


@pytest.mark.gpu
def test_cuda_solve_writes_npy_files_at_the_reference_cadence(tmp_path):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(2, 32)
    g = Solver(cfg)
    out = str(tmp_path / "data")
    g.set_output_dir(out)
    steps = g.solve(5 * g.dt * 0.999, 2)
    assert steps == 5
    masters = sorted(f for f in os.listdir(out) if f.endswith(".json"))
    assert masters == ["CajitaFluids%05d.json" % t for t in (0, 2, 4)]
    m = json.load(open(os.path.join(out, masters[-1])))
    assert m["cycle"] == 4 and m["dim"] == 2 and m["global_num_cell"] == [32, 32] and len(m["blocks"]) == 1
    # the last write happened after step index 4, i.e. on the final state
    assert m["time"] == g.time and m["dtime"] == g.dt
    q, vel, nodes = g.output()
    blk = m["blocks"][0]
    assert blk["offset"] == [0, 0] and blk["extent"] == [32, 32]
    fq = np.load(os.path.join(out, blk["quantity"]))
    fv = np.load(os.path.join(out, blk["velocity"]))
    assert fq.shape == (32, 32) and fv.shape == (2, 32, 32)
    assert np.array_equal(fq, q[0]) and np.array_equal(fv, vel[:, 0])
    base = blk["quantity"].replace("quantity.npy", "")
    assert np.array_equal(np.load(os.path.join(out, base + "nodes_x.npy")), nodes[0])
    assert np.array_equal(np.load(os.path.join(out, base + "nodes_y.npy")), nodes[1])
    # an earlier write holds an earlier state: an oracle run stopped after step index 2
    o = run(Oracle(cfg), 3)
    oq, ov, _ = o.output()
    name2 = "raw/CajitaFluidsOutput%05d%05d." % (0, 2)
    assert np.array_equal(np.load(os.path.join(out, name2 + "quantity.npy")), oq[0])
    assert np.array_equal(np.load(os.path.join(out, name2 + "velocity.npy")), ov[:, 0])


@pytest.mark.gpu
def test_cuda_write_output_is_deferred_until_flush(tmp_path):
    from cajitafluids_b200 import Solver
    cfg = make_cfg(3, 16)
    g = run(Solver(cfg), 1)
    out = str(tmp_path / "o")
    q0, v0, _ = g.output()
    g.write_output(out, 7)
    g.step()  # the time loop goes on while the copy is in flight; the files hold the state of the write
    g.output_flush()
    fq = np.load(os.path.join(out, "raw", "CajitaFluidsOutput%05d%05d.quantity.npy" % (0, 7)))
    fv = np.load(os.path.join(out, "raw", "CajitaFluidsOutput%05d%05d.velocity.npy" % (0, 7)))
    assert fq.shape == (16, 16, 16) and fv.shape == (3, 16, 16, 16)
    assert np.array_equal(fq, q0) and np.array_equal(fv, v0)
    assert not np.array_equal(g.output()[0], q0)
