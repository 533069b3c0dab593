"""Several ranks of the host-emulated library (tests/emul) in one process: the block decomposition, the
three-sweep field halo, the NCCL-path CG (three-kernel form, all-gathered exact sums), the pressure halo
and the output stage's gather, held against the SINGLE-BLOCK oracle bit for bit — on the block grids of
1, 2, 4 and 8 GPUs (1x1x2, 1x2x2, 2x2x2) and on x / y splits.

Covers the host logic of cajitafluids_b200/csrc/halo.cu (boxes, neighbours, message sizes, sweep order)
that the GPU tests could only exercise on two GPUs so far.  Not covered here: the NVLink peer-memory
exchange and the fused two-kernel CG form (GPU tests only).  See tests/emul/README.md.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))

import build_emul  # noqa: E402
from cajitafluids_b200 import config as K  # noqa: E402
from cajitafluids_b200._capi import Library  # noqa: E402
from helpers import fields_of, make_cfg  # noqa: E402
from multirank import block_slices, run_ranks  # noqa: E402
from oracle_api import Oracle  # noqa: E402


# "loops": the TMA kernels (stencil, phase B of the two-kernel CG form) are plain-loop stand-ins;
# "tma": they are the product's kernels themselves (TMA loads as synchronous copies, one fiber per CUDA thread)
@pytest.fixture(scope="module", params=["loops", "tma"])
def emul(request):
    import ctypes
    lib = Library(build_emul.build(tma=request.param == "tma"), "cfb_", mode=ctypes.RTLD_LOCAL)
    lib.tma = request.param == "tma"  # fibers make these runs ~20x slower: the tests shorten themselves
    return lib


GRIDS = [(2, None), (4, None), (8, None), (2, (2, 1, 1)), (2, (1, 2, 1)), (4, (2, 2, 1)), (3, (3, 1, 1)), (6, (1, 3, 2))]


# under "tma" (fibers: ~20x slower) the block grids that add something over "loops": all three axes split, an x
# split with a y split, uneven blocks
TMA_GRIDS = {(8, None), (4, (2, 2, 1)), (3, (3, 1, 1)), (2, (2, 1, 1))}


def tma_subset(emul, world, blocks):
    if emul.tma and (world, blocks) not in TMA_GRIDS:
        pytest.skip("block grid covered by the loops library; tma runs a subset")


def cfg3(cells=(24, 20, 18), **kw):
    return make_cfg(3, cells, box=tuple(c / cells[0] for c in cells), **kw)


@pytest.mark.parametrize("world,blocks", GRIDS)
def test_field_gather_fills_every_ghost_including_edges_and_corners(emul, world, blocks):
    tma_subset(emul, world, blocks)
    cfg = cfg3()
    ora = Oracle(cfg)
    rng = np.random.default_rng(3)
    glob = {}
    for f in fields_of(3):
        a = rng.uniform(-1, 1, size=ora.shape(f))
        ora.set(f, a)
        glob[f] = (a, ora.get(f, region=K.GHOSTED))  # physical ghosts are zero

    def body(ctx, rank):
        for f in fields_of(3):
            ctx.set(f, glob[f][0][block_slices(ctx, f)])
        ctx.gather(K.CURRENT)
        off = ctx.global_offset()
        bad = []
        for f in fields_of(3):
            mine = ctx.get(f, region=K.GHOSTED)
            ez, ey, ex = mine.shape
            want = glob[f][1][off[2]:off[2] + ez, off[1]:off[1] + ey, off[0]:off[0] + ex]
            if not np.array_equal(mine, want):
                bad.append((f, int((mine != want).sum())))
        return bad

    assert run_ranks(emul, cfg, world, body, blocks) == [[]] * world


@pytest.mark.parametrize("world,blocks", GRIDS)
def test_decomposed_pcg_is_bit_identical_to_the_single_block_oracle(emul, world, blocks):
    tma_subset(emul, world, blocks)
    cfg = cfg3()
    ora = Oracle(cfg)
    rng = np.random.default_rng(77)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, bo, ho = ora.get(K.PRESSURE), ora.get(K.RHS), ora.residual_history()

    def body(ctx, rank):
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        ctx.add_inputs()
        ctx.build_rhs()
        ok_rhs = np.array_equal(ctx.get(K.RHS), bo[block_slices(ctx, K.RHS)])
        ig, rg = ctx.pcg_solve()
        ok_p = np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)])
        return ok_rhs, ig, rg, ok_p, np.array_equal(ctx.residual_history(), ho)

    for res in run_ranks(emul, cfg, world, body, blocks):
        assert res == (True, io, ro, True, True), res


@pytest.mark.parametrize("world,blocks", [(2, None), (4, None), (8, None), (4, (2, 2, 1))])
def test_decomposed_steps_and_output_match_the_single_block_oracle(emul, world, blocks):
    tma_subset(emul, world, blocks)
    cfg = cfg3(cells=(32, 24, 16), body_force=(0.0, -3.0, 0.5))
    ora = Oracle(cfg)
    ora.setup()
    want0 = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    nsteps = 1 if emul.tma else 2
    for _ in range(nsteps):
        ora.step()
    want = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    oq, ov, _ = ora.output()
    its = ora.stats()["cg_iterations"]

    gl = {f: max(np.linalg.norm(want[f].ravel()), 1e-300) for f in want}

    def body(ctx, rank):
        # Block-local LocalMesh coordinates (own low corner + local index * cell, like every rank of the
        # reference computes them) differ from the single block's in the last bit, so advected fields agree
        # to rounding, not bit for bit; everything before the first advection does (asserted below).
        ctx.setup()
        exact = [f for f in want0 if not np.array_equal(ctx.get(f), want0[f][block_slices(ctx, f)])]
        for _ in range(nsteps):
            ctx.step()
        err = {f: float(np.linalg.norm((ctx.get(f) - want[f][block_slices(ctx, f)]).ravel()) / gl[f]) for f in want}
        q, v, _ = ctx.output()
        sl = block_slices(ctx, K.QUANTITY)
        err["out_q"] = float(np.abs(q - oq[sl]).max())
        err["out_vel"] = float(np.abs(v - ov[(slice(None),) + sl]).max())
        return exact, err, ctx.stats()["cg_iterations"], ctx.time

    for exact, err, it, t in run_ranks(emul, cfg, world, body, blocks):
        assert exact == []
        assert max(err.values()) < 1e-12, err
        assert abs(it - its) <= 3 and t == ora.time


# ---- the opt-in multigrid preconditioner, block-decomposed (the same global cycle on every decomposition) ----
MG_GRIDS = [(2, None, (32, 32, 32)), (4, None, (32, 32, 32)), (8, None, (32, 32, 32)), (2, (2, 1, 1), (32, 24, 16)),
            (4, (2, 2, 1), (32, 24, 16)), (3, (1, 3, 1), (16, 24, 8)), (8, None, (16, 16, 16))]


def mg_levels_of(cells, blocks):
    """depth of the block-wise coarsening: while every block extent is even and >= 2 after halving"""
    n = [c // b for c, b in zip(cells, blocks)]
    lv = 1
    while all(e % 2 == 0 and e // 2 >= 2 for e in n):
        n = [e // 2 for e in n]
        lv += 1
    return lv


@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer"])
@pytest.mark.parametrize("world,blocks,cells", MG_GRIDS)
@pytest.mark.parametrize("nu", [(2, 2, 8), (1, 1, 3), (3, 0, 2), (2, 0, 1)])
def test_decomposed_vcycle_is_bit_identical_to_the_single_block_oracle(emul, world, blocks, cells, nu, peer):
    if emul.tma and (world, blocks) not in ((8, None), (2, (2, 1, 1))):
        pytest.skip("the fine-level sweeps on the TMA march: all axes split, and an x split")
    from cajitafluids_b200.distributed import block_grid
    cfg = cfg3(cells=cells)
    bl = blocks or block_grid(world, 3)
    ora = Oracle(cfg)
    ora.set_mg_max_levels(mg_levels_of(cells, bl))
    ora.set_preconditioner("mg", *nu)
    r = np.random.default_rng(4).standard_normal(ora.shape(K.PRESSURE))
    z = ora.mg_apply(r)

    def body(ctx, rank):
        ctx.set_tuning("mg_tma", 1)  # the fine level on the TMA march: off by default for several blocks (not yet on several GPUs)
        ctx.set_preconditioner("mg", *nu)
        sl = block_slices(ctx, K.PRESSURE)
        ok = [np.array_equal(ctx.mg_apply(r[sl]), z[sl]) for _ in range(2)]  # twice: sequence numbers carry over
        return ctx.mg_num_levels(), all(ok)

    assert run_ranks(emul, cfg, world, body, blocks, peer=peer) == [(ora.mg_num_levels(), True)] * world


@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer"])
@pytest.mark.parametrize("world,blocks,cells", MG_GRIDS)
def test_decomposed_mg_pcg_is_bit_identical_to_the_single_block_oracle(emul, world, blocks, cells, peer):
    from cajitafluids_b200.distributed import block_grid
    cfg = cfg3(cells=cells)
    bl = blocks or block_grid(world, 3)
    ora = Oracle(cfg)
    ora.set_mg_max_levels(mg_levels_of(cells, bl))
    ora.set_preconditioner("mg")
    rng = np.random.default_rng(77)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()
    assert io < 40

    def body(ctx, rank):
        ctx.set_tuning("mg_tma", 1)
        ctx.set_preconditioner("mg")
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        ctx.add_inputs()
        ctx.build_rhs()
        ig, rg = ctx.pcg_solve()
        return ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)]), \
            np.array_equal(ctx.residual_history(), ho)

    for res in run_ranks(emul, cfg, world, body, blocks, peer=peer):
        assert res == (io, ro, True, True), res


def test_decomposed_mg_steps_match_the_single_block_oracle(emul):
    cells = (32, 32, 32)
    cfg = cfg3(cells=cells)
    ora = Oracle(cfg)
    ora.set_mg_max_levels(mg_levels_of(cells, (2, 2, 2)))
    ora.set_preconditioner("mg")
    ora.setup()
    want0 = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    ora.step()
    want = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    gl = {f: max(np.linalg.norm(want[f].ravel()), 1e-300) for f in want}
    its = ora.stats()["cg_iterations"]

    def body(ctx, rank):
        ctx.set_tuning("mg_tma", 1)
        ctx.set_preconditioner("mg")
        ctx.setup()
        exact = [f for f in want0 if not np.array_equal(ctx.get(f), want0[f][block_slices(ctx, f)])]
        ctx.step()
        err = {f: float(np.linalg.norm((ctx.get(f) - want[f][block_slices(ctx, f)]).ravel()) / gl[f]) for f in want}
        return exact, err, ctx.stats()["cg_iterations"]

    for exact, err, it in run_ranks(emul, cfg, 8, body):
        assert exact == [] and max(err.values()) < 1e-12 and abs(it - its) <= 1, (exact, err, it, its)


def test_mg_needs_evenly_divided_blocks(emul):
    from cajitafluids_b200 import CfbError
    cfg = cfg3(cells=(30, 20, 18))  # 30 cells over 4 blocks

    def body(ctx, rank):
        try:
            ctx.set_preconditioner("mg")
        except CfbError as e:
            return "divide evenly" in str(e)
        return False

    assert run_ranks(emul, cfg, 4, body, (4, 1, 1)) == [True] * 4


# ---- the default multi-GPU path of the CG iterations: NVLink peer-memory exchange (halo.cu: peer_setup,
# peer_exchange, cg_xchg_kernel, cg_xunpack_kernel are the product's; phases A / B are the plain-loop stand-ins)
@pytest.mark.parametrize("world,blocks", GRIDS)
def test_peer_memory_exchange_pcg_is_bit_identical_to_the_single_block_oracle(emul, world, blocks):
    tma_subset(emul, world, blocks)
    cfg = cfg3()
    ora = Oracle(cfg)
    rng = np.random.default_rng(78)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()
    reps = 1 if emul.tma and world != 8 else 2

    def body(ctx, rank):
        peer = ctx.stats()["peer_mode"]
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        out = []
        for _ in range(reps):  # twice: sequence numbers and p buffers carry over from solve to solve
            ctx.add_inputs()
            ctx.build_rhs()
            ig, rg = ctx.pcg_solve()
            out.append((ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)]),
                        np.array_equal(ctx.residual_history(), ho)))
        return peer, out

    for peer, out in run_ranks(emul, cfg, world, body, blocks, peer=True):
        assert peer == 1
        assert out == [(io, ro, True, True)] * reps, out


@pytest.mark.parametrize("world,blocks", [(2, None), (8, None), (4, (2, 2, 1))])
def test_peer_memory_exchange_steps_and_fixed_iterations(emul, world, blocks):
    tma_subset(emul, world, blocks)
    cfg = cfg3(cells=(32, 24, 16), fixed_iters=25)
    ora = Oracle(cfg)
    ora.setup()
    want0 = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    ora.step()
    want = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    gl = {f: max(np.linalg.norm(want[f].ravel()), 1e-300) for f in want}

    def body(ctx, rank):
        ctx.setup()
        exact = [f for f in want0 if not np.array_equal(ctx.get(f), want0[f][block_slices(ctx, f)])]
        ctx.step()
        err = {f: float(np.linalg.norm((ctx.get(f) - want[f][block_slices(ctx, f)]).ravel()) / gl[f]) for f in want}
        return ctx.stats()["peer_mode"], exact, err

    for peer, exact, err in run_ranks(emul, cfg, world, body, blocks, peer=True):
        assert peer == 1 and exact == [] and max(err.values()) < 1e-12, (peer, exact, err)


@pytest.mark.parametrize("world,blocks,peer", [(2, (1, 2, 1), False), (4, (2, 2, 1), False), (2, (2, 1, 1), True),
                                               (4, (2, 2, 1), True), (3, (3, 1, 1), True)])
def test_two_dimensional_block_decomposition(emul, world, blocks, peer):
    """The reference's own dimensionality, several blocks (quirks Q1 / Q2 on): projection bit for bit, steps to
    rounding (block-local coordinates)."""
    if emul.tma and not (world == 4 and peer):
        pytest.skip("tma runs one 2-D case")
    cfg = make_cfg(2, (48, 36), box=(1.0, 0.75))
    ora = Oracle(cfg)
    ora.setup()
    want0 = {f: ora.get(f) for f in fields_of(2) + [K.PRESSURE]}
    nsteps = 1 if emul.tma else 2
    for _ in range(nsteps):
        ora.step()
    want = {f: ora.get(f) for f in fields_of(2) + [K.PRESSURE]}
    gl = {f: max(np.linalg.norm(want[f].ravel()), 1e-300) for f in want}

    def body(ctx, rank):
        ctx.setup()
        exact = [f for f in want0 if not np.array_equal(ctx.get(f), want0[f][block_slices(ctx, f)])]
        for _ in range(nsteps):
            ctx.step()
        err = {f: float(np.linalg.norm((ctx.get(f) - want[f][block_slices(ctx, f)]).ravel()) / gl[f]) for f in want}
        return ctx.stats()["peer_mode"], exact, err

    for pm, exact, err in run_ranks(emul, cfg, world, body, blocks, peer=peer):
        assert pm == (1 if peer else 0) and exact == [] and max(err.values()) < 1e-12, (pm, exact, err)


@pytest.mark.parametrize("world,blocks,peer", [(8, None, True), (8, None, False), (4, (2, 2, 1), True), (2, (2, 1, 1), True),
                                               (2, (1, 2, 1), False), (6, (1, 3, 2), True)])
def test_64_byte_iteration_block_decomposed(emul, world, blocks, peer):
    """cg_variant 2 on several blocks: phase A' reads the ghosts of the search direction that the exchange after
    phase B delivered (peer path) / that one more exchange delivers (NCCL path)."""
    if emul.tma and world not in (8, 4):
        pytest.skip("tma runs a subset")
    cfg = cfg3(fixed_iters=15) if emul.tma else cfg3()
    ora = Oracle(cfg)
    rng = np.random.default_rng(79)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()

    def body(ctx, rank):
        ctx.set_tuning("cg_variant", 2)
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        ctx.add_inputs()
        ctx.build_rhs()
        ig, rg = ctx.pcg_solve()
        return ctx.stats()["peer_mode"], ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)]), \
            np.array_equal(ctx.residual_history(), ho)

    for res in run_ranks(emul, cfg, world, body, blocks, peer=peer):
        assert res == (1 if peer else 0, io, ro, True, True), res


MODES = [("peer, small tiles", True, {"fused_stages": 4, "fused_zc": 4, "fused_tx": 64, "fused_ty": 8}),
         ("peer, phase B reads the x ghosts from the staging areas", True, {"peer_xstage": 1}),
         ("peer, faces on the side stream under interior work, reductions in the kernels' last blocks", True,
          {"peer_overlap": 1}),
         ("peer, overlapped exchange, small tiles", True,
          {"peer_overlap": 1, "fused_stages": 4, "fused_zc": 4, "fused_tx": 64, "fused_ty": 8}),
         ("NCCL, interior overlapped with the r/p halo", False, {"overlap_halo": 1}),
         ("NCCL, overlap, small tiles, several boundary units", False,
          {"overlap_halo": 1, "fused_stages": 4, "fused_zc": 4, "fused_tx": 64, "fused_ty": 8}),
         ("three-kernel, NCCL", False, {"cg_variant": 0})]


@pytest.mark.parametrize("name,peer,tune", MODES, ids=[m[0] for m in MODES])
@pytest.mark.parametrize("world,blocks,cells", [(8, None, (24, 20, 18)), (2, (2, 1, 1), (128, 24, 20))])
def test_every_exchange_schedule_of_the_cg_iterations(emul, world, blocks, cells, name, peer, tune):
    """The tuning modes the multi-GPU worker runs on real GPUs (tests/mgpu_worker.py section 3), on 2x2x2 blocks
    and on an x split whose blocks end exactly on a 64-wide tile (the staging-read variant's case)."""
    if not emul.tma:
        pytest.skip("unit lists and staging reads exist in the TMA kernels only")
    cfg = cfg3(cells=cells, fixed_iters=12)
    ora = Oracle(cfg)
    rng = np.random.default_rng(80)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po = ora.get(K.PRESSURE)

    def body(ctx, rank):
        for k, v in tune.items():
            ctx.set_tuning(k, v)
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        ctx.add_inputs()
        ctx.build_rhs()
        ig, rg = ctx.pcg_solve()
        return ctx.stats()["peer_mode"], ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)])

    want_peer = 1 if peer and tune.get("cg_variant", 1) >= 1 else 0
    for res in run_ranks(emul, cfg, world, body, blocks, peer=peer):
        assert res == (want_peer, io, ro, True), (name, res)


def test_decomposed_solve_writes_one_file_set_per_block_and_one_master(emul, tmp_path):
    """siloWrite on several ranks: every block writes its own arrays, rank 0 the master naming all of them
    (writeMultiObjects, src/SiloWriter.hpp:292-346); stitched together they are the single-block output."""
    import json
    if emul.tma:
        pytest.skip("output path: nothing TMA-specific")
    cells = (32, 24, 16)
    cfg = cfg3(cells=cells)
    out = str(tmp_path / "data")
    ora = Oracle(cfg)
    steps = ora.solve(3 * ora.dt * 0.999, 0)
    oq, ov, _ = ora.output()

    def body(ctx, rank):
        ctx.set_output_dir(out)
        n = ctx.solve(3 * ctx.dt * 0.999, 2)
        return n, ctx.global_offset(), ctx.owned_extent(K.QUANTITY)

    res = run_ranks(emul, cfg, 4, body, (2, 2, 1))
    assert [r[0] for r in res] == [steps] * 4
    m = json.load(open(os.path.join(out, "CajitaFluids%05d.json" % 2)))
    assert m["cycle"] == 2 and m["global_num_cell"] == list(cells) and len(m["blocks"]) == 4
    q = np.full(oq.shape, np.nan)
    v = np.full(ov.shape, np.nan)
    for blk, (_, off, ext) in zip(m["blocks"], res):
        assert blk["offset"] == list(off) and blk["extent"] == list(ext)
        sl = tuple(slice(off[d], off[d] + ext[d]) for d in (2, 1, 0))
        q[sl] = np.load(os.path.join(out, blk["quantity"]))
        v[(slice(None),) + sl] = np.load(os.path.join(out, blk["velocity"]))
    # written after step index 2 = the final state; advected fields agree to rounding across decompositions
    assert np.abs(q - oq).max() < 1e-12 and np.abs(v - ov).max() < 1e-12


# (world, blocks, fixed iterations: 0 = to convergence, once — fibers are slow)
GRIDS_FUSED = [(2, None, 0), (2, (2, 1, 1), 14), (2, (1, 2, 1), 14), (4, (2, 2, 1), 14), (6, (1, 3, 2), 14), (8, None, 14)]


@pytest.mark.parametrize("world,blocks,fixed", GRIDS_FUSED)
def test_overlapped_exchange_on_every_block_grid(emul, world, blocks, fixed):
    """"peer_overlap": the reduction of each phase in the last block of its compute kernel (mailboxes), the faces on
    the side stream (cg_face_kernel: neighbour-to-neighbour arrival flags), boundary units of phase B last.  Converged
    and repeated solves, switching between the schedules from solve to solve (mailbox and face sequence numbers,
    p buffers carry over), x / y / z / uneven splits."""
    if not emul.tma:
        pytest.skip("unit lists exist in the TMA kernels only")
    cfg = cfg3(cells=(23, 20, 21), fixed_iters=fixed)
    ora = Oracle(cfg)
    rng = np.random.default_rng(83)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()

    def body(ctx, rank):
        for f, a in vel.items():
            ctx.set(f, a[block_slices(ctx, f)])
        out = []
        for overlap in (1, 0, 1, 1):
            ctx.set_tuning("peer_overlap", overlap)
            ctx.add_inputs()
            ctx.build_rhs()
            ig, rg = ctx.pcg_solve()
            out.append((ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)]),
                        np.array_equal(ctx.residual_history(), ho)))
        return ctx.stats()["peer_mode"], out

    for peer, out in run_ranks(emul, cfg, world, body, blocks, peer=True):
        assert peer == 1
        assert out == [(io, ro, True, True)] * 4, out


def test_overlapped_exchange_whole_steps(emul):
    """(same bar as test_peer_memory_exchange_steps_and_fixed_iterations)"""
    if not emul.tma:
        pytest.skip("unit lists exist in the TMA kernels only")
    cfg = cfg3(cells=(32, 24, 16), fixed_iters=20)
    ora = Oracle(cfg)
    ora.setup()
    want0 = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    ora.step()
    want = {f: ora.get(f) for f in fields_of(3) + [K.PRESSURE]}
    gl = {f: max(np.linalg.norm(want[f].ravel()), 1e-300) for f in want}

    def body(ctx, rank):
        ctx.set_tuning("peer_overlap", 1)
        ctx.setup()
        exact = [f for f in want0 if not np.array_equal(ctx.get(f), want0[f][block_slices(ctx, f)])]
        ctx.step()
        err = {f: float(np.linalg.norm((ctx.get(f) - want[f][block_slices(ctx, f)]).ravel()) / gl[f]) for f in want}
        return exact, err

    for exact, err in run_ranks(emul, cfg, 8, body, None, peer=True):
        assert exact == [] and max(err.values()) < 1e-12, (exact, err)


@pytest.mark.parametrize("peer", [True, False], ids=["peer mailboxes", "NCCL all-gather"])
@pytest.mark.parametrize("world,blocks,cells", [(8, None, (24, 20, 18)), (2, (2, 1, 1), (128, 24, 20)), (4, (1, 2, 2), (23, 20, 21))])
def test_single_reduction_cg_block_decomposed(emul, world, blocks, cells, peer):
    """cg_variant 3 on several blocks: ONE reduction point (three sums through the mailboxes in the stencil kernel's
    last block, or one all-gather) and one ghost exchange (the faces of r) per iteration; bit for bit the
    single-block checker's statement of the same algorithm, to convergence and with fixed iterations."""
    if not emul.tma:
        pytest.skip("the plain-loop stand-in has no mailbox form")
    for fixed in (0, 9):
        cfg = cfg3(cells=cells, fixed_iters=fixed)
        ora = Oracle(cfg)
        ora.set_cg_algorithm(1)
        rng = np.random.default_rng(85)
        vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
        for f, a in vel.items():
            ora.set(f, a)
        ora.add_inputs()
        ora.build_rhs()
        io, ro = ora.pcg_solve()
        po, ho = ora.get(K.PRESSURE), ora.residual_history()

        def body(ctx, rank):
            ctx.set_tuning("cg_variant", 3)
            out = []
            for _ in range(2):  # twice: sequence numbers and vectors carry over
                for f, a in vel.items():
                    ctx.set(f, a[block_slices(ctx, f)])
                ctx.add_inputs()
                ctx.build_rhs()
                ig, rg = ctx.pcg_solve()
                out.append((ig, rg, np.array_equal(ctx.get(K.PRESSURE), po[block_slices(ctx, K.PRESSURE)]),
                            np.array_equal(ctx.residual_history(), ho)))
            return out

        for out in run_ranks(emul, cfg, world, body, blocks, peer=peer):
            assert out == [(io, ro, True, True)] * 2, (fixed, out)
