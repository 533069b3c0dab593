"""Multi-GPU parity worker, launched by tests/test_multigpu.py (or by hand) under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_worker.py [--cells 48 40 36] [--steps 4]

Every rank owns one block of the 3-D block decomposition on its own GPU (NCCL halo exchange and
allreduce inside the C library) and checks its block against the SINGLE-BLOCK CPU oracle run on the
same global problem: decomposition invariance + parity in one go.
Exit code 0 on success on every rank.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from cajitafluids_b200 import Solver, config as K
from cajitafluids_b200.distributed import attach_nccl, block_grid, decompose
from helpers import fields_of, make_cfg, rel_l2
from oracle_api import Oracle


def block_slices(gpu, field):
    """numpy slices (z, y, x) of this rank's owned entities inside the global owned array."""
    off = gpu.global_offset()
    ext = gpu.owned_extent(field)
    return tuple(slice(off[d], off[d] + ext[d]) for d in (2, 1, 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs=3, default=[48, 40, 36])
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--blocks", type=int, nargs=3, default=None, help="block grid (default: split z, y, x)")
    ap.add_argument("--skip-steps", action="store_true", help="only the stencil / gather / PCG sections")
    ap.add_argument("--quick", action="store_true", help="only the PCG section, three exchange modes")
    ap.add_argument("--mg", action="store_true", help="only the multigrid-preconditioner section")
    ap.add_argument("--cg1", action="store_true", help="only the single-reduction CG section (cg_variant 3)")
    ap.add_argument("--everything", action="store_true",
                    help="every section in one launch (default sections, multigrid, overlapped exchange, single-reduction CG): "
                         "one rendezvous instead of four when GPU time is charged per GPU")
    ap.add_argument("--overlap", action="store_true",
                    help="the PCG section and whole steps with the overlapped exchange (peer_overlap) on and off")
    args = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cells = tuple(args.cells)
    blocks = tuple(args.blocks) if args.blocks else block_grid(world)
    assert blocks[0] * blocks[1] * blocks[2] == world
    box = tuple(c / cells[0] for c in cells)

    def gcfg(**kw):
        return make_cfg(3, cells, box=box, **kw)

    def rank_cfg(**kw):
        c = decompose(gcfg(**kw), rank, world, blocks)
        c.device_id = local
        return attach_nccl(c, dist)

    failures = []

    def check(name, ok, detail=""):
        if not ok:
            failures.append(f"rank {rank}: {name} {detail}")

    rng = np.random.default_rng(77)
    if args.everything:
        sections_1_2(args, rank, gcfg, rank_cfg, rng, check)
        section_3(args, rank, gcfg, rank_cfg, rng, check)
        section_4(args, rank, gcfg, rank_cfg, check)
        args.overlap = True
        section_3(args, rank, gcfg, rank_cfg, rng, check)
        section_4(args, rank, gcfg, rank_cfg, check, tune={"peer_overlap": 1})
        section_6(args, rank, gcfg, rank_cfg, rng, check)
        if all(c % (2 * b) == 0 for c, b in zip(cells, blocks)):
            section_5(args, rank, blocks, gcfg, rank_cfg, rng, check)
    elif args.mg:
        section_5(args, rank, blocks, gcfg, rank_cfg, rng, check)
    elif args.cg1:
        section_6(args, rank, gcfg, rank_cfg, rng, check)
    elif args.overlap:
        section_3(args, rank, gcfg, rank_cfg, rng, check)
        for on in (1, 0):
            section_4(args, rank, gcfg, rank_cfg, check, tune={"peer_overlap": on})
    else:
        if not args.quick:
            sections_1_2(args, rank, gcfg, rank_cfg, rng, check)
        section_3(args, rank, gcfg, rank_cfg, rng, check)
        if not (args.quick or args.skip_steps):
            section_4(args, rank, gcfg, rank_cfg, check)

    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for msg in failures:
        print("FAIL", msg, flush=True)
    if rank == 0:
        print(f"mgpu_worker world={world} blocks={blocks} cells={cells}: "
              f"{'OK' if int(flag) == 0 else 'FAILED (%d)' % int(flag)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


def sections_1_2(args, rank, gcfg, rank_cfg, rng, check):
    # 1. stencil + dot on a random global p --------------------------------------------------------
    ora = Oracle(gcfg())
    gpu = Solver(rank_cfg())
    p = rng.uniform(-1, 1, size=ora.shape(K.CG_P))
    ora.set(K.CG_P, p)
    gpu.set(K.CG_P, p[block_slices(gpu, K.CG_P)])
    dg, _ = gpu.stencil_dot(1)
    do, _ = ora.stencil_dot(1)
    qo = ora.get(K.CG_Q)[block_slices(gpu, K.CG_Q)]
    qg = gpu.get(K.CG_Q)
    check("stencil q bit-exact", np.array_equal(qg, qo), f"max diff {np.abs(qg - qo).max()}")
    # each rank's pAp is its local sum until the allreduce; stencil_dot returns the allreduced value
    check("stencil dot", abs(dg - do) <= 1e-12 * abs(do), f"{dg} vs {do}")

    # 2. width-3 field gather incl. edges/corners -----------------------------------------------------
    for f in fields_of(3):
        a = rng.uniform(-1, 1, size=ora.shape(f))
        ora.set(f, a)
        gpu.set(f, a[block_slices(gpu, f)])
    gpu.gather(K.CURRENT)
    h = 3
    off = gpu.global_offset()
    for f in fields_of(3):
        glob = ora.get(f, region=K.GHOSTED)  # global ghosted array, physical ghosts are zero
        mine = gpu.get(f, region=K.GHOSTED)
        ez, ey, ex = mine.shape
        want = glob[off[2]:off[2] + ez, off[1]:off[1] + ey, off[0]:off[0] + ex]
        check(f"gather field {f}", np.array_equal(mine, want), f"mismatches {(mine != want).sum()}")
    gpu.close()
    ora.close()



def section_3(args, rank, gcfg, rank_cfg, rng, check):
    # 3. PCG on a synthetic divergence: every CG form / halo schedule, bit for bit ----------------------
    ora = Oracle(gcfg())
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()
    modes = [("two-kernel, NVLink peer stores (default)", {"cg_variant": 1, "peer_halo": 1}),
             ("two-kernel, NVLink peer stores, small tiles",
              {"cg_variant": 1, "peer_halo": 1, "fused_stages": 4, "fused_zc": 4, "fused_tx": 64, "fused_ty": 8}),
             ("two-kernel, NVLink peer stores, phase B reads the x ghosts from the staging areas",
              {"cg_variant": 1, "peer_halo": 1, "peer_xstage": 1}),
             ("two-kernel, NCCL, interior overlapped with the r/p halo",
              {"cg_variant": 1, "peer_halo": 0, "overlap_halo": 1}),
             ("two-kernel, NCCL, halo first", {"cg_variant": 1, "peer_halo": 0, "overlap_halo": 0}),
             ("two-kernel, NCCL, small tiles, several boundary units",
              {"cg_variant": 1, "peer_halo": 0, "overlap_halo": 1, "fused_stages": 4, "fused_zc": 4, "fused_tx": 64,
               "fused_ty": 8}),
             ("three-kernel, NCCL", {"cg_variant": 0}),
             ("two-kernel without a stored q (64 B/cell), NVLink peer stores", {"cg_variant": 2, "peer_halo": 1}),
             ("two-kernel without a stored q (64 B/cell), NCCL", {"cg_variant": 2, "peer_halo": 0})]
    if args.quick:
        modes = [modes[0], modes[1], modes[2], modes[4]]
    if args.overlap:
        modes = [("overlapped exchange (faces on the side stream, reductions in the kernels' last blocks)",
                  {"cg_variant": 1, "peer_halo": 1, "peer_overlap": 1}),
                 ("overlapped exchange, small tiles",
                  {"cg_variant": 1, "peer_halo": 1, "peer_overlap": 1, "fused_stages": 4, "fused_zc": 4, "fused_tx": 64,
                   "fused_ty": 8}),
                 ("exchange kernel after each phase", {"cg_variant": 1, "peer_halo": 1, "peer_overlap": 0}),
                 ("overlapped exchange requested, 64-byte form (falls back to the exchange kernel)",
                  {"cg_variant": 2, "peer_halo": 1, "peer_overlap": 1})]
    for name, tune in modes:
        gpu = Solver(rank_cfg())
        for k, v in tune.items():
            gpu.set_tuning(k, v)
        if tune.get("peer_halo") == 1:
            check(f"peer-memory mode active [{name}]", gpu.stats()["peer_mode"] == 1,
                  "cudaIpc mapping of the neighbours' arrays failed: NCCL fallback in use")
        for f, a in vel.items():
            gpu.set(f, a[block_slices(gpu, f)])
        gpu.add_inputs()
        gpu.build_rhs()
        check(f"rhs bit-exact [{name}]", np.array_equal(gpu.get(K.RHS), ora.get(K.RHS)[block_slices(gpu, K.RHS)]))
        ig, rg = gpu.pcg_solve()
        check(f"pcg iterations [{name}]", abs(ig - io) <= 1, f"{ig} vs {io}")
        if ig == io:
            e = rel_l2(gpu.get(K.PRESSURE), po[block_slices(gpu, K.PRESSURE)])
            check(f"pcg pressure 1e-10 [{name}]", e < 1e-10, f"rel l2 {e}")
            # exact (double-double) global sums make the decomposition invisible:
            check(f"pcg pressure bit-exact [{name}]",
                  np.array_equal(gpu.get(K.PRESSURE), po[block_slices(gpu, K.PRESSURE)]))
            check(f"pcg residual history bit-exact [{name}]", np.array_equal(gpu.residual_history(), ho))
        gpu.close()
    ora.close()



def section_4(args, rank, gcfg, rank_cfg, check, tune=None):
    # 4. whole timesteps of the default inflow problem -----------------------------------------------------
    ora = Oracle(gcfg())
    gpu = Solver(rank_cfg())
    for k, v in (tune or {}).items():
        gpu.set_tuning(k, v)
    ora.setup()
    gpu.setup()
    for _ in range(args.steps):
        ora.step()
        gpu.step()
    check("cg iteration total", abs(gpu.stats()["cg_iterations"] - ora.stats()["cg_iterations"]) <= args.steps + 1,
          f"{gpu.stats()['cg_iterations']} vs {ora.stats()['cg_iterations']}")
    for f in fields_of(3) + [K.PRESSURE]:
        e = rel_l2(gpu.get(f), ora.get(f)[block_slices(gpu, f)])
        # blocks far from the inflow may hold (near-)zero fields: compare against the global norm too
        gl = np.linalg.norm(ora.get(f).ravel())
        ea = np.linalg.norm((gpu.get(f) - ora.get(f)[block_slices(gpu, f)]).ravel()) / max(gl, 1e-300)
        check(f"step field {f} {tune or ''}", min(e, ea) < 1e-10, f"rel l2 {e} (vs global norm {ea})")
    gpu.close()
    ora.close()


def section_5(args, rank, blocks, gcfg, rank_cfg, rng, check):
    # 5. opt-in multigrid preconditioner: the same global V-cycle, block-decomposed --------------------------
    n = [c // b for c, b in zip(args.cells, blocks)]
    levels = 1
    while all(e % 2 == 0 and e // 2 >= 2 for e in n):
        n = [e // 2 for e in n]
        levels += 1
    ora = Oracle(gcfg())
    ora.set_mg_max_levels(levels)  # the blocks stop coarsening when a BLOCK extent becomes odd
    ora.set_preconditioner("mg")
    r = rng.standard_normal(ora.shape(K.PRESSURE))
    z = ora.mg_apply(r)
    vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
    for f, a in vel.items():
        ora.set(f, a)
    ora.add_inputs()
    ora.build_rhs()
    io, ro = ora.pcg_solve()
    po, ho = ora.get(K.PRESSURE), ora.residual_history()
    gpu = Solver(rank_cfg())
    gpu.set_preconditioner("mg")
    check("mg levels", gpu.mg_num_levels() == levels, f"{gpu.mg_num_levels()} vs {levels}")
    sl = block_slices(gpu, K.PRESSURE)
    zg = gpu.mg_apply(r[sl])
    check("V-cycle 1e-13", rel_l2(zg, z[sl]) < 1e-13, f"rel l2 {rel_l2(zg, z[sl])}")
    check("V-cycle bit-exact", np.array_equal(zg, z[sl]))
    # the fine level on the TMA march ("mg_tma" 1: off by default for several blocks until this section has run on GPUs)
    gpu.set_tuning("mg_tma", 1)
    gpu.set_preconditioner("mg")
    zt = gpu.mg_apply(r[sl])
    check("V-cycle on the TMA march bit-exact", np.array_equal(zt, z[sl]))
    gpu.set_tuning("mg_tma", -1)
    gpu.set_preconditioner("mg")
    for f, a in vel.items():
        gpu.set(f, a[block_slices(gpu, f)])
    gpu.add_inputs()
    gpu.build_rhs()
    ig, rg = gpu.pcg_solve()
    check("mg-pcg iterations", abs(ig - io) <= 1, f"{ig} vs {io}")
    e = rel_l2(gpu.get(K.PRESSURE), po[sl])
    check("mg-pcg pressure 1e-10", e < 1e-10, f"rel l2 {e}")
    check("mg-pcg pressure bit-exact", np.array_equal(gpu.get(K.PRESSURE), po[sl]))
    check("mg-pcg residual history bit-exact", np.array_equal(gpu.residual_history(), ho))
    gpu.close()
    ora.close()


def section_6(args, rank, gcfg, rank_cfg, rng, check):
    # 6. opt-in single-reduction CG (cg_variant 3): one reduction point and one ghost exchange per iteration ------
    for fixed in (0, 11):
        ora = Oracle(gcfg(fixed_iters=fixed))
        ora.set_cg_algorithm(1)
        vel = {f: rng.uniform(-1, 1, size=ora.shape(f)) for f in fields_of(3)[1:]}
        for f, a in vel.items():
            ora.set(f, a)
        ora.add_inputs()
        ora.build_rhs()
        io, ro = ora.pcg_solve()
        po, ho = ora.get(K.PRESSURE), ora.residual_history()
        for name, tune in (("mailboxes in the stencil kernel's last block", {"peer_halo": 1}), ("NCCL all-gather", {"peer_halo": 0})):
            gpu = Solver(rank_cfg(fixed_iters=fixed))
            gpu.set_tuning("cg_variant", 3)
            for k, v in tune.items():
                gpu.set_tuning(k, v)
            for rep in range(2):
                for f, a in vel.items():
                    gpu.set(f, a[block_slices(gpu, f)])
                gpu.add_inputs()
                gpu.build_rhs()
                ig, rg = gpu.pcg_solve()
                tag = f"[{name}, fixed={fixed}, solve {rep}]"
                check(f"cg1 iterations {tag}", abs(ig - io) <= 1, f"{ig} vs {io}")
                e = rel_l2(gpu.get(K.PRESSURE), po[block_slices(gpu, K.PRESSURE)])
                check(f"cg1 pressure 1e-10 {tag}", e < 1e-10, f"rel l2 {e}")
                check(f"cg1 pressure bit-exact {tag}", np.array_equal(gpu.get(K.PRESSURE), po[block_slices(gpu, K.PRESSURE)]))
                check(f"cg1 residual history bit-exact {tag}", ig == io and np.array_equal(gpu.residual_history(), ho))
            gpu.close()
        ora.close()


if __name__ == "__main__":
    main()
