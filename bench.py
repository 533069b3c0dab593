#!/usr/bin/env python
"""bench.py — headline benchmark of the pressure-projection hot path (BASELINE.json).

Metric: Jacobi-PCG iterations per second at 512^3 FP64 per GPU (weak scaling: every rank owns a
512^3 block of a 3-D block-decomposed box), with the achieved HBM GB/s of the dominant kernel
(stencil7 + dot) against the measured B200 roofline.  A "step" is one fixed-iteration PCG solve
(`--iters`, default 100; config 2/3 of BASELINE.json use fixed iteration counts because the
reference's own tol/max_iter would stop at 2000 < ~2500 needed) on a synthetic,
decomposition-independent divergence right-hand side that is already resident in HBM.

  python bench.py --gpus N --steps K --warmup W          our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   the CPU restatement of the reference
                                                         (oracle/, OpenMP) on the box's host cores

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pcg_iterations_per_second"
UNIT = "iterations/s"
BYTES_STENCIL = 16  # read p, write q                       
# algorithmic bytes per owned cell of one CG iteration / of its dominant kernel, per CG form
#   0: three kernels  axpy+norms 48 | p-update 24 | stencil7+dot 16            (kernels_cg/stencil.cu)
#   1: two kernels    r-update+norms 24 | x-update + p-update + stencil7 + dot 48   (kernels_fused.cu)
#   2: two kernels    stencil7 recomputed + r-update+norms 24 | x-update + p-update + stencil7 + dot, q not stored 40
# the exchange schedule of the CG iterations the library picks on several GPUs (cfb_internal.h: peer_overlap)
DEFAULT_PEER_OVERLAP = False
BYTES_ITER = {0: 88, 1: 72, 2: 64, 3: 88}
BYTES_DOMINANT = {0: 16, 1: 48, 2: 40, 3: 16}
KERNEL_DOMINANT = {0: "stencil7_dot_tma", 1: "cg_fused_kernel (x-update + p-update + stencil7 + p.Ap dot)",
                   2: "cg_fused_kernel (x-update + p-update + stencil7 + p.Ap dot, q not stored)",
                   3: "stencil7_dot_tma MODE 2 (u = M^-1 r, w = A u, three sums) of the single-reduction form"}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s from B200_PROFILING.md (of fallback)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def block_grid(n):
    """1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2, 8 -> 2x2x2 (split z, then y, then x)."""
    px = py = pz = 1
    dims = [1, 1, 1]
    k, d = n, 2
    while k > 1:
        dims[d] *= 2
        k //= 2
        d = (d - 1) % 3
    px, py, pz = dims
    assert px * py * pz == n, "gpus must be a power of two"
    return px, py, pz


def make_config(args, rank, world, blocks, scaling=None, cells=None, gravity=0.0, iters=None):
    from cajitafluids_b200 import default_config
    n = cells or args.cells
    if (scaling or args.scaling) == "weak":
        gcells = tuple(n * b for b in blocks)
    else:
        gcells = (n, n, n)
    # h = 1/512 fixed; gravity g: body force (0, -g, 0) as examples/advection.cpp:452
    cfg = default_config(3, gcells, box=tuple(c / 512.0 for c in gcells), gravity=gravity)
    bz, r = divmod(rank, blocks[0] * blocks[1])
    by, bx = divmod(r, blocks[0])
    for d, (b, bid) in enumerate(zip(blocks, (bx, by, bz))):
        cfg.ranks_per_dim[d] = b
        cfg.block_id[d] = bid
    cfg.world_rank, cfg.world_size = rank, world
    cfg.cg_fixed_iters = args.iters if iters is None else iters
    cfg.cg_print_level = 0
    return cfg, gcells


def workload_name(args):
    return f"pcg_{args.cells}cubed_fixed{args.iters}_jacobi_synthetic_divergence"


class CpuArm:
    """The CPU restatement of the reference (oracle/: stored-coefficient matrix, 4-kernel / 3-reduction CG, OpenMP)
    on the same 1-GPU workload, with every host core whatever OMP_NUM_THREADS the launcher exported (torchrun sets
    it to 1).  One context for the whole arm; a timed step is a solve of `iters_sample` fixed iterations and is
    normalised per iteration: the per-solve set-up (cg_init and the first A p, which the GPU arm amortises over its
    100 iterations per step) is measured with one-iteration solves and taken out."""

    def __init__(self, cells):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        import psutil
        import oracle_api
        from cajitafluids_b200 import config as K, default_config

        self.cores = oracle_api.set_num_threads(0)
        n = cells
        need = 23 * (n + 6) ** 3 * 8
        avail = psutil.virtual_memory().available
        self.note = ""
        while need > 0.6 * avail and n > 64:
            n //= 2
            need = 23 * (n + 6) ** 3 * 8
            self.note = f" (host RAM too small for the full grid, sampled at {n}^3)"
        self.n = n
        cfg = default_config(3, n, box=n / 512.0)
        cfg.cg_fixed_iters = 1
        o = oracle_api.Oracle(cfg)
        o.set_accumulation(False)  # plain double sums: the reference's arithmetic and cost
        # same synthetic MAC velocity family as the GPU arm (sin/cos product)
        h = o.cell_size
        ax = [np.arange(n + 1) * h, (np.arange(n) + 0.5) * h]
        L = n * h
        for d, f in enumerate((K.U, K.V, K.W)):
            g = [np.sin(np.pi * ax[0] / L) if e == d else np.cos(2 * np.pi * ax[1] / L) for e in range(3)]
            val = g[2][:, None, None] * g[1][None, :, None] * g[0][None, None, :]
            sl = [slice(None)] * 3
            sl[2 - d] = 0
            val[tuple(sl)] = 0.0
            sl[2 - d] = -1
            val[tuple(sl)] = 0.0
            o.set(f, val)
            del val
        o.build_rhs()
        self.o = o

    def solve(self, iters):
        self.o.set_fixed_iters(iters)
        self.o.build_rhs()  # (also resets the pressure: the checker's solve starts from the lhs it finds)
        t0 = time.perf_counter()
        self.o.pcg_solve()
        return time.perf_counter() - t0

    def measure(self, iters_sample, steps, warmup, budget_s):
        """-> dict(t_iter, t_setup, step_seconds[], iters_sample, steps).  Stops adding steps when `budget_s` is
        used up (never below one timed step)."""
        t_all = time.perf_counter()
        t1 = min(self.solve(1) for _ in range(2))
        probe = self.solve(2)
        t_it0 = max(probe - t1, 1e-6)
        # a step must leave room for all K + W of them inside the budget
        left = budget_s - (time.perf_counter() - t_all)
        fit = int((left / max(1, steps + warmup) - t1) / t_it0)
        iters_sample = max(2, min(iters_sample, fit))
        for _ in range(warmup):
            self.solve(iters_sample)
        secs = []
        for _ in range(steps):
            secs.append(self.solve(iters_sample))
            if time.perf_counter() - t_all > budget_s and secs:
                break
        # per-iteration time with the set-up taken out; timing noise on tiny grids must not push it outside what the
        # set-up can possibly be (between nothing and three iterations' worth: cg_init + the first A p)
        t_iter = [min(max((t - t1) / (iters_sample - 1), t / (iters_sample + 3)), t / iters_sample) for t in secs]
        mean_it = sum(t_iter) / len(t_iter)
        return {"t_iter": mean_it, "t_setup": max(t1 - mean_it, 0.0), "step_seconds": secs,
                "iters_sample": iters_sample, "steps": len(secs)}

    def describe(self, m):
        return (f"{m['steps']} solves of {m['iters_sample']} fixed Jacobi-PCG iterations at {self.n}^3 FP64 on "
                f"{self.cores} host cores{self.note}, per-iteration rate with the per-solve set-up "
                f"({m['t_setup']:.2f} s: cg_init + first A p, from one-iteration solves) taken out; CPU restatement of "
                "the reference (Kokkos/Cajita unavailable): stored 7-coefficient matrix + M^-1 array, 4 kernels / "
                "3 reductions per iteration, OpenMP")

    def close(self):
        self.o.close()


def cpu_baseline(args):
    """bounded sample for the main arm's line (N = 1 only)."""
    arm = CpuArm(args.cells)
    m = arm.measure(10, 2, 0, 40.0)
    out = {"value": 1.0 / m["t_iter"], "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.describe(m),
           "value_with_setup": m["iters_sample"] * m["steps"] / sum(m["step_seconds"])}
    arm.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cells
    arm = CpuArm(n)
    # the whole arm (K + W steps) within a few minutes: ~150 s of timed solves plus the set-up of the context
    m = arm.measure(min(args.iters, 20), args.steps, args.warmup, 150.0)
    value = 1.0 / m["t_iter"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.describe(m)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": m["steps"], "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(m["step_seconds"]) / m["steps"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "cells_per_gpu": [n, n, n] if args.scaling == "weak" else None,
                       "cg_iters_per_step": args.iters,
                       "cg_iters_timed_per_step": m["iters_sample"],
                       "normalisation": "iterations/s = 1 / per-iteration time; a step of this arm is a bounded sample "
                                        "of the workload's step (fewer iterations of the same solve), its per-solve "
                                        "set-up taken out",
                       "unit_of_work": "one CG iteration over one %d^3 block; the host cores work on one block at a "
                                       "time, so this figure does not depend on the number of GPUs of the other arm "
                                       "(whose value sums block-iterations over ranks)" % n,
                       "launched_world_size": world},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    arm.close()
    print(json.dumps(line), flush=True)


def mg_side_measurements(args):
    """`--side mg` (run by the main arm in a child process): time to solution of ONE projection of the default
    inflow problem on the bench grid with the reference's stopping test (sqrt(sum r^2) <= 1e-6), Jacobi (the
    reference's preconditioner; max_iter raised, the reference's 2000 do not converge at 512^3)
    against the opt-in multigrid V(2,2) preconditioner, then whole timesteps with multigrid.
    Prints one JSON object with the keys to merge into `extra`."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from cajitafluids_b200 import Solver, default_config
    torch.cuda.set_device(0)
    out = {}
    tts = {}
    for kind in ("jacobi", "mg", "mg_graph", "mg_graph_coarse"):
        try:
            c3 = default_config(3, args.cells, box=args.cells / 512.0)
            c3.cg_max_iter = 20000
            c3.cg_print_level = 0
            s3 = Solver(c3)
            if "graph" in kind:  # the V-cycle's ~60 launches replayed as one CUDA graph
                s3.set_tuning("mg_graph", 1)
            if "coarse" in kind:  # levels <= 16^3 down and back up in one single-CTA kernel
                s3.set_tuning("mg_coarse_kernel", 1)
            s3.set_preconditioner("jacobi" if kind == "jacobi" else "mg")
            s3.add_inputs()
            s3.build_rhs()
            s3.pcg_solve()  # warm-up (first-launch costs, graph capture)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it3, res3 = s3.pcg_solve()
            torch.cuda.synchronize()
            tts[kind] = {"seconds": time.perf_counter() - t0, "cg_iterations": it3, "final_residual": res3}
            s3.close()
        except Exception as e:  # noqa: BLE001
            tts[kind] = {"error": repr(e)[:300]}
    if "seconds" in tts["jacobi"] and "seconds" in tts["mg"]:
        tts["speedup"] = tts["jacobi"]["seconds"] / tts["mg"]["seconds"]
    tts["note"] = ("one pressure solve of the default inflow problem at %d^3 to sqrt(sum r^2) <= 1e-6, "
                   "wall clock incl. convergence polling; mg = opt-in V(2,2) cycle, never the default" % args.cells)
    out["projection_time_to_solution"] = tts
    # the same whole timesteps as extra.timesteps_per_s (reference tol / max_iter), with the opt-in
    # multigrid preconditioner
    if args.timestep_cells > 0:
        try:
            s4 = Solver(default_config(3, args.timestep_cells))
            s4.set_preconditioner("mg")
            s4.setup()
            for _ in range(2):
                s4.step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it0 = s4.stats()["cg_iterations"]
            for _ in range(5):
                s4.step()
            torch.cuda.synchronize()
            dt4 = time.perf_counter() - t0
            out["timesteps_per_s_mg"] = {"cells": [args.timestep_cells] * 3, "value": 5 / dt4,
                                         "cg_iters_per_step": (s4.stats()["cg_iterations"] - it0) / 5,
                                         "interp_order": 3, "preconditioner": "opt-in multigrid V(2,2)"}
            s4.close()
        except Exception as e:  # noqa: BLE001
            out["timesteps_per_s_mg"] = {"error": repr(e)[:300]}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())


def probe_advect(args):
    """`--side advect` (child process of the main arm): the advection kernel with rows vs entity tiles
    ("advect_tile") on the bench grid.  Prints one JSON object with the keys to merge into `extra`."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from cajitafluids_b200 import Solver, default_config
    torch.cuda.set_device(0)
    adv = {}
    for tile in (0, 1):
        try:
            cfg = default_config(3, args.cells, box=args.cells / 512.0)
            cfg.cg_fixed_iters = 5
            cfg.cg_print_level = 0
            s = Solver(cfg)
            s.set_tuning("advect_tile", tile)
            s.setup()
            s.step()
            s.reset_stats()
            for _ in range(3):
                s.step()
            adv["tile%d_ms" % tile] = s.stats()["ms_advect"] / 3
            s.close()
        except Exception as e:  # noqa: BLE001
            adv["tile%d_error" % tile] = repr(e)[:200]
    adv["note"] = "advection kernel (q, u, v, w in one launch) at %d^3, rows (tile0) vs 32x2x2 entity tiles (tile1)" % args.cells
    sys.stdout.flush()
    os.write(json_fd, (json.dumps({"advect_tile_probe": adv}) + "\n").encode())


def probe_flat2d(args):
    """`--side flat2d` (child process of the main arm): two-dimensional fixed-iteration solves (the reference's own
    dimensionality) with and without the ghost-plane loads ("flat_2d"), both two-kernel CG forms."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from cajitafluids_b200 import Solver, default_config
    torch.cuda.set_device(0)
    n2 = args.cells * 16 if args.cells >= 64 else args.cells  # 8192^2 for the 512^3 bench
    flat = {"cells": [n2, n2], "iters": 50}
    for fl in (0, 1):
        for variant in (1, 2):
            key = "flat%d_variant%d" % (fl, variant)
            try:
                cfg = default_config(2, n2, box=n2 / 512.0)
                cfg.cg_print_level = 0
                s = Solver(cfg)
                s.set_tuning("flat_2d", fl)
                s.set_tuning("cg_variant", variant)
                s.fill_synthetic_velocity(0)
                s.build_rhs()
                for _ in range(3):
                    s.pcg_fixed(50)
                ms, res = s.pcg_fixed(50)
                flat[key] = {"iterations_per_s": 50 / (ms * 1e-3), "residual": res}
                s.close()
            except Exception as e:  # noqa: BLE001
                flat[key] = {"error": repr(e)[:200]}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps({"flat_2d_probe": flat}) + "\n").encode())


def run_child(extra, key, argv, timeout):
    """a side measurement in a child process with a time limit; a failure is reported under ITS OWN key."""
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0 or not p.stdout.strip():
            raise RuntimeError("child exit %d: %s" % (p.returncode, p.stderr.strip()[-300:]))
        extra.update(json.loads(p.stdout.strip().splitlines()[-1]))
    except Exception as e:  # noqa: BLE001
        extra[key] = {"error": repr(e)[:400]}


def multi_gpu_extras(args, s, extra, dist, torch, np, rank, world, local, blocks, units, resid, fallback, barrier,
                     all_ranks_ok):
    """Several GPUs, measured in the same run so that the driver's scaling runs record them at every N:
    extra.exchange_schedules  the headline workload with each exchange schedule of the CG iterations;
    extra.strong_scaling      BASELINE configs[3]: the 1-GPU grid (cells^3 GLOBAL) block-decomposed over the N GPUs;
    extra.decomposition_parity  a small block-decomposed run (solve + 3 whole steps, 96^3 global) against the
                              single-block CPU oracle on rank 0: the analogue of the reference's only distributed
                              test (tests/tstProblemManager.cpp:61-98), for the peer path and the NCCL path."""
    from cajitafluids_b200 import Solver, config as K
    from cajitafluids_b200.distributed import attach_nccl

    def timed(solver, steps=3, warm=2):
        err, ms_sum, res = None, 0.0, None
        try:
            for _ in range(warm):
                solver.pcg_fixed(args.iters)
            for _ in range(steps):
                m, res = solver.pcg_fixed(args.iters)
                ms_sum += m
        except Exception as e:  # noqa: BLE001
            err = e
        if not all_ranks_ok(err is None):
            return {"error": repr(err)[:300] if err is not None else "another rank failed"}
        t = torch.tensor([ms_sum], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"iterations_per_s": steps * args.iters / (float(t[0]) * 1e-3), "final_residual": res, "steps": steps}

    # 1. exchange schedules on the headline workload (same right-hand side: equal residuals bit for bit)
    sched = {}
    if s.stats()["peer_mode"] and not fallback:
        for name, tune in (("exchange_kernel_64_byte_form", {"cg_variant": 2, "peer_overlap": 0}),
                           ("exchange_kernel_72_byte_form", {"cg_variant": 1, "peer_overlap": 0}),
                           ("overlapped_72_byte_form", {"cg_variant": 1, "peer_overlap": 1}),
                           ("nccl_sendrecv_allgather", {"cg_variant": args.cg_variant, "peer_overlap": 0, "peer_halo": 0})):
            try:
                for k, v in tune.items():
                    s.set_tuning(k, v)
                r = timed(s)
                if "iterations_per_s" in r:
                    r = {"value": r["iterations_per_s"] * units, "unit": UNIT,
                         "same_residual_as_headline": r["final_residual"] == resid, "steps": r["steps"]}
                sched[name] = r
            except Exception as e:  # noqa: BLE001
                sched[name] = {"error": repr(e)[:300]}
        s.set_tuning("peer_halo", 1)
        s.set_tuning("cg_variant", args.cg_variant_requested)
        s.set_tuning("peer_overlap", int(DEFAULT_PEER_OVERLAP))
        sched["headline_schedule"] = "exchange_kernel_%d_byte_form" % BYTES_ITER[args.cg_variant]
        extra["exchange_schedules"] = sched

    # 2. strong scaling (BASELINE configs[3]): cells^3 global over all ranks, each exchange schedule and the opt-in
    #    single-reduction form
    if args.scaling == "weak":
        try:
            cfg_s, g_s = make_config(args, rank, world, blocks, scaling="strong")
            cfg_s.device_id = local
            attach_nccl(cfg_s, dist)
            ss = Solver(cfg_s)
            ss.fill_synthetic_velocity(0)
            ss.build_rhs()
            out = {"global_cells": list(g_s), "blocks": list(blocks), "cg_iters_per_step": args.iters,
                   "cells_local": int(np.prod(ss.owned_extent(K.QUANTITY))), "peer_mode": ss.stats()["peer_mode"],
                   "unit": "global CG iterations/s of the fixed %d^3 problem on %d GPUs (divide by the N = 1 headline "
                           "value for the speed-up)" % (args.cells, world)}
            for name, tune in (("library_default", {"cg_variant": -1, "peer_overlap": 0}),
                               ("exchange_kernel_64_byte_form", {"cg_variant": 2, "peer_overlap": 0}),
                               ("exchange_kernel_72_byte_form", {"cg_variant": 1, "peer_overlap": 0}),
                               ("overlapped_72_byte_form", {"cg_variant": 1, "peer_overlap": 1}),
                               ("single_reduction_cg_variant3", {"cg_variant": 3, "peer_overlap": 0}),
                               ("nccl_sendrecv_allgather", {"cg_variant": -1, "peer_overlap": 0, "peer_halo": 0})):
                try:
                    for k, v in tune.items():
                        ss.set_tuning(k, v)
                    out[name] = timed(ss, steps=5, warm=3)
                except Exception as e:  # noqa: BLE001
                    out[name] = {"error": repr(e)[:300]}
            default_key = "library_default"
            if "iterations_per_s" in out.get(default_key, {}):
                out["iterations_per_s"] = out[default_key]["iterations_per_s"]
                out["schedule"] = default_key
            extra["strong_scaling"] = out
            ss.close()
        except Exception as e:  # noqa: BLE001
            extra["strong_scaling"] = {"error": repr(e)[:300]}

    # 2b. the same strong-scaled problem on a block grid that does not split x (its faces are strided in memory and
    #     travel through a staging area): 8 GPUs as 1 x 2 x 4 instead of 2 x 2 x 2
    if args.scaling == "weak" and world == 8 and not args.blocks:
        try:
            alt = (1, 2, 4)
            cfg_a, g_a = make_config(args, rank, world, alt, scaling="strong")
            cfg_a.device_id = local
            attach_nccl(cfg_a, dist)
            sa = Solver(cfg_a)
            sa.fill_synthetic_velocity(0)
            sa.build_rhs()
            r = timed(sa, steps=5, warm=3)
            r.update({"global_cells": list(g_a), "blocks": list(alt), "cells_local": int(np.prod(sa.owned_extent(K.QUANTITY)))})
            extra["strong_scaling_blocks_1x2x4"] = r
            sa.close()
        except Exception as e:  # noqa: BLE001
            extra["strong_scaling_blocks_1x2x4"] = {"error": repr(e)[:300]}

    # 3. decomposition parity against the single-block oracle
    try:
        extra["decomposition_parity"] = decomposition_parity(dist, torch, np, rank, world, local, blocks)
    except Exception as e:  # noqa: BLE001
        extra["decomposition_parity"] = {"error": repr(e)[:300]}

    # 4. whole timesteps with the opt-in multigrid preconditioner, weak-scaled: timestep_cells^3 per GPU of the default
    #    inflow problem (box edge = blocks, h = 1 / timestep_cells), the same global V-cycle block-decomposed
    if args.timestep_cells > 0:
        try:
            from cajitafluids_b200 import default_config
            from cajitafluids_b200.distributed import decompose
            n = args.timestep_cells
            gcfg = default_config(3, tuple(n * b for b in blocks), box=tuple(float(b) for b in blocks))
            cfg_m = decompose(gcfg, rank, world, blocks)
            cfg_m.device_id = local
            attach_nccl(cfg_m, dist)
            sm = Solver(cfg_m)
            sm.set_preconditioner("mg")
            err, dt_m, its = None, 0.0, 0
            try:
                sm.setup()
                for _ in range(2):
                    sm.step()
                barrier()
                it0 = sm.stats()["cg_iterations"]
                t0 = time.perf_counter()
                for _ in range(5):
                    sm.step()
                barrier()
                dt_m = time.perf_counter() - t0
                its = sm.stats()["cg_iterations"] - it0
            except Exception as e:  # noqa: BLE001
                err = e
            if all_ranks_ok(err is None):
                t = torch.tensor([dt_m], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                extra["timesteps_per_s_mg"] = {"global_cells": [n * b for b in blocks], "cells_per_gpu": [n] * 3,
                                               "value": 5 / float(t[0]), "cg_iters_per_step": its / 5, "interp_order": 3,
                                               "mg_levels": sm.mg_num_levels(), "peer_mode": sm.stats()["peer_mode"],
                                               "preconditioner": "opt-in multigrid V(2,2), the global cycle block-decomposed"}
            else:
                extra["timesteps_per_s_mg"] = {"error": repr(err)[:300] if err is not None else "another rank failed"}
            sm.close()
        except Exception as e:  # noqa: BLE001
            extra["timesteps_per_s_mg"] = {"error": repr(e)[:300]}


def decomposition_parity(dist, torch, np, rank, world, local, blocks, n=96, steps=3):
    from cajitafluids_b200 import Solver, config as K, default_config
    from cajitafluids_b200.distributed import attach_nccl, decompose
    gcfg = default_config(3, n)  # the default inflow problem, reference tol / max_iter
    fields = [K.QUANTITY, K.U, K.V, K.W, K.PRESSURE]
    # rank 0: the single-block oracle, every host core; its fields and iteration counts go to every rank
    shapes = {f: ((n + 1 if f == K.W else n), (n + 1 if f == K.V else n), (n + 1 if f == K.U else n)) for f in fields}
    ref = {f: torch.empty(shapes[f], dtype=torch.float64, device="cuda") for f in fields}
    its = torch.zeros(steps + 1, dtype=torch.float64, device="cuda")
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api
        oracle_api.set_num_threads(0)
        o = oracle_api.Oracle(gcfg)
        o.setup()
        cnt = [o.stats()["cg_iterations"]]
        for _ in range(steps):
            o.step()
            cnt.append(o.stats()["cg_iterations"])
        for f in fields:
            ref[f].copy_(torch.from_numpy(np.ascontiguousarray(o.get(f))))
        its.copy_(torch.tensor(np.diff([0] + cnt), dtype=torch.float64))
        o.close()
    for f in fields:
        dist.broadcast(ref[f], src=0)
    dist.broadcast(its, src=0)
    want_its = [int(v) for v in its.tolist()]
    out = {"global_cells": [n] * 3, "blocks": list(blocks), "steps": steps, "oracle_cg_iterations": want_its,
           "checker": "single-block CPU oracle on rank 0 (oracle/, exact sums)"}
    for name, tune in (("peer", {}), ("nccl", {"peer_halo": 0})):
        cfg = decompose(gcfg, rank, world, blocks)
        cfg.device_id = local
        attach_nccl(cfg, dist)
        g = Solver(cfg)
        for k, v in tune.items():
            g.set_tuning(k, v)
        g.setup()
        cnt = [g.stats()["cg_iterations"]]
        for _ in range(steps):
            g.step()
            cnt.append(g.stats()["cg_iterations"])
        got_its = [int(v) for v in np.diff([0] + cnt)]
        off = g.global_offset()
        num = torch.zeros(len(fields), dtype=torch.float64, device="cuda")
        den = torch.zeros(len(fields), dtype=torch.float64, device="cuda")
        same = torch.ones(1, dtype=torch.float64, device="cuda")
        for i, f in enumerate(fields):
            mine = g.get(f)
            ext = mine.shape
            want = ref[f][off[2]:off[2] + ext[0], off[1]:off[1] + ext[1], off[0]:off[0] + ext[2]].cpu().numpy()
            num[i] = float(((mine - want) ** 2).sum())
            den[i] = float((want ** 2).sum())
            if not np.array_equal(mine, want):
                same[0] = 0.0
        dist.all_reduce(num)
        dist.all_reduce(den)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rel = [float((a / b) ** 0.5) if b > 0 else (0.0 if a == 0 else float("inf")) for a, b in zip(num.tolist(), den.tolist())]
        out[name] = {"peer_mode": g.stats()["peer_mode"], "peer_overlap": g.stats()["peer_overlap"],
                     "cg_iterations": got_its, "iters_equal": got_its == want_its,
                     "iters_within_1": all(abs(a - b) <= 1 for a, b in zip(got_its, want_its)),
                     "max_rel_l2": max(rel), "rel_l2": dict(zip(["q", "u", "v", "w", "p"], rel)),
                     "bit_identical": bool(float(same[0]) == 1.0)}
        g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=512, help="cells per side per GPU (weak) / global (strong)")
    ap.add_argument("--iters", type=int, default=100, help="fixed CG iterations per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-timestep", action="store_true")
    ap.add_argument("--blocks", type=int, nargs=3, default=None, help="block grid bx by bz (default: split z, y, x)")
    ap.add_argument("--timestep-cells", type=int, default=128, help="grid of the extra full-timestep measurement")
    ap.add_argument("--cg-variant", type=int, default=-1, choices=[-1, 0, 1, 2, 3],
                    help="-1 = the library's own choice (default: the 64-byte form at this size), 1 = two-kernel iteration "
                         "(72 B/cell), 0 = three kernels (88 B), 2 = two kernels without a stored q (64 B), 3 = single-reduction form")
    ap.add_argument("--tune", action="append", default=[], help="key=value passed to cfb_set_tuning")
    ap.add_argument("--no-probe", action="store_true", help="skip the side measurement of the 64-byte CG form")
    ap.add_argument("--side", default=None, choices=["mg", "advect", "flat2d"], help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.side == "mg":
        return mg_side_measurements(args)
    if args.side == "advect":
        return probe_advect(args)
    if args.side == "flat2d":
        return probe_flat2d(args)

    # stdout carries exactly ONE JSON line: everything else that libraries print there (NCCL's version
    # banner, torch warnings) is sent to stderr at file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    from cajitafluids_b200 import Solver, config as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blocks = tuple(args.blocks) if args.blocks else block_grid(world)
    assert blocks[0] * blocks[1] * blocks[2] == world, "--blocks must multiply to --gpus"
    cfg, gcells = make_config(args, rank, world, blocks)
    cfg.device_id = local
    if world > 1:
        from cajitafluids_b200.distributed import attach_nccl
        attach_nccl(cfg, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    s = Solver(cfg)
    ncell_local = int(np.prod(s.owned_extent(K.QUANTITY)))
    ncell_global = int(np.prod(gcells))
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)
    s.set_tuning("cg_variant", args.cg_variant)  # (-1: the library's choice)
    for kv in args.tune:
        k, v = kv.split("=")
        s.set_tuning(k, int(v))

    def all_ranks_ok(ok):
        if dist is None:
            return ok
        t = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0]) == 0.0

    # warm-up; if the NVLink peer-memory exchange fails on ANY rank (a mapping that cannot be used, a
    # peer that never publishes), all ranks agree to fall back to the NCCL exchange and warm up again
    fallback = None
    err = None
    try:
        for _ in range(args.warmup):
            s.pcg_fixed(args.iters)
    except Exception as e:  # noqa: BLE001
        err = e
    if not all_ranks_ok(err is None):
        if world > 1 and s.stats()["peer_mode"]:
            fallback = f"peer-memory exchange failed during warm-up ({err}); NCCL send/recv used instead"
            s.set_tuning("peer_halo", 0)
            for _ in range(args.warmup):
                s.pcg_fixed(args.iters)
        else:
            raise err if err is not None else RuntimeError("another rank failed during warm-up")
    barrier()
    s.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, resid = s.pcg_fixed(args.iters)  # CUDA events on the library's stream around the whole solve
        dev_ms += ms
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.summary() if rank == 0 else None
    st = s.stats()
    if dist is not None:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0]), float(t[1])
    args.cg_variant_requested = args.cg_variant
    args.cg_variant = int(st["cg_variant"])  # the form that ran (the library's choice when none was asked for)
    total_iters = args.iters * args.steps
    global_its = total_iters / (dev_ms * 1e-3)
    # whole-job value: one unit = one CG iteration over one cells^3 block.  Weak scaling: every rank
    # owns such a block, so a global iteration processes `world` units; strong: the global grid is one unit.
    units = world if args.scaling == "weak" else 1
    value = global_its * units

    # roofline of the dominant kernel, from the per-kernel events recorded inside the timed region
    peak, peak_src = measured_peak()
    kt = max(1, st["k_timed_iters"])
    t_st = st["ms_k_stencil"] / kt
    v = args.cg_variant
    bdom, biter = BYTES_DOMINANT[v], BYTES_ITER[v]
    ach = ncell_local * bdom / (t_st * 1e-3) / 1e9 if t_st > 0 else 0.0
    it_ach = ncell_local * biter * total_iters / (dev_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": KERNEL_DOMINANT[v], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": bdom,
                "algorithmic_bytes_per_launch": ncell_local * bdom,
                "avg_launch_ms": t_st,
                "iteration": {"form": "two kernels" if v >= 1 else "three kernels", "bytes_per_cell": biter,
                              "achieved_gbs": it_ach, "frac": it_ach / peak}}
    if world > 1:
        # several GPUs: the events bracket each phase INCLUDING the exchange that follows its compute kernel; the
        # exposed exchange time is bracketed separately, so the kernel-only figure is stated next to it
        xa, xb = st["ms_k_exch_a"] / kt, st["ms_k_exch_b"] / kt
        k_only = max(t_st - xb, 1e-9)
        roofline.update({"avg_launch_ms": k_only, "achieved": ncell_local * bdom / (k_only * 1e-3) / 1e9,
                         "frac": ncell_local * bdom / (k_only * 1e-3) / 1e9 / peak,
                         "exchange": {"schedule": "overlapped: faces on the side stream, reductions in the compute "
                                                  "kernels' last blocks (their barrier wait is inside the kernel times)"
                                      if st["peer_overlap"] else
                                      ("one exchange kernel after each phase (NVLink peer stores + mailbox reduction)"
                                       if st["peer_mode"] else "NCCL send/recv before phase B, all-gather after each phase"),
                                      "exposed_after_phase_a_ms": xa, "exposed_after_phase_b_ms": xb,
                                      "phase_a_kernel_ms": st["ms_k_axpy"] / kt - xa, "phase_b_kernel_ms": k_only}})
    if v == 3:
        roofline["iteration"].update({"update_ms": st["ms_k_axpy"] / kt, "stencil_ms": t_st})
    elif v >= 1:
        roofline["iteration"].update({"rupdate_ms": st["ms_k_axpy"] / kt, "fused_ms": t_st})
    else:
        roofline["iteration"].update({"axpy_ms": st["ms_k_axpy"] / kt, "pupdate_ms": st["ms_k_pupdate"] / kt,
                                      "stencil_ms": t_st})
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(f"cg_variant{v}_dominant_{args.cells}")
            roofline["traffic_source"] = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                          "ncu --set full capture of this kernel at this size (a committed constant, "
                                          "not measured in this run)")
        except Exception:
            pass
    # the plain stencil7 + dot kernel (16 B/cell: CG kernel 4 on its own, used for the first q = A p
    # of every solve), timed back to back on the same vectors
    if world == 1:
        _, st_ms = s.stencil_dot(20)
        roofline["stencil7_dot_alone"] = {"avg_launch_ms": st_ms, "algorithmic_bytes_per_cell": BYTES_STENCIL,
                                          "achieved": ncell_local * BYTES_STENCIL / (st_ms * 1e-3) / 1e9,
                                          "frac": ncell_local * BYTES_STENCIL / (st_ms * 1e-3) / 1e9 / peak}
    launches = st["kernel_launches"]

    # e2e: the solver plug-in call with HOST vectors (pinned), H2D of b and D2H of x inside the timed region
    e2e = None
    rhs_norm = None
    if not args.no_e2e:
        shp = s.shape(K.RHS)
        b_host = torch.empty(shp, dtype=torch.float64).pin_memory().numpy()
        x_host = torch.empty(shp, dtype=torch.float64).pin_memory().numpy()
        b_host[...] = s.get(K.RHS)
        bsq = float(np.vdot(b_host.ravel(), b_host.ravel()))  # |b|^2 of this block (report |b| and the final |r|)
        if dist is not None:
            t = torch.tensor([bsq], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            bsq = float(t[0])
        rhs_norm = bsq ** 0.5
        s.pcg_solve_host(b_host, x_host)  # warm
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(args.steps, 3))
        for _ in range(e_steps):
            _, it, _ = s.pcg_solve_host(b_host, x_host)
        barrier()
        e_wall = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_wall = float(t[0])
        e2e = {"value": args.iters * e_steps / e_wall * units, "unit": UNIT, "h2d_bytes_per_step": int(b_host.nbytes) * world,
               "d2h_bytes_per_step": int(x_host.nbytes) * world, "steps": e_steps,
               "call": "cfb_pcg_solve_host (ReferenceConjugateGradient::solve(b, x) with host vectors)"}
        del b_host, x_host

    # extra: whole timesteps (advect + inputs + projection) of the default inflow problem, 1 GPU only
    extra = {"final_residual": resid, "rhs_norm": rhs_norm, "wall_s_timed_region": wall, "cells_local": ncell_local,
             "global_iterations_per_s": global_its,
             "value_unit": "CG iterations of one %d^3 block per second, summed over ranks" % args.cells}
    # whole timesteps (advect + inputs + projection) on the bench grid itself, at every N: the
    # projection runs the same fixed number of CG iterations as the headline (the reference's own
    # tol/max_iter cannot converge at 512^3)
    if world > 1 and not args.no_probe and not args.tune:
        multi_gpu_extras(args, s, extra, dist, torch, np, rank, world, local, blocks, units, resid, fallback, barrier,
                         all_ranks_ok)
    if not args.no_timestep:
        # BASELINE configs[4]: full timestep (advection + inputs + projection) with inflow source AND body force
        # (g = 9.8: src/BodyForce.hpp:43-60; the dt clamp then includes sqrt(g h), src/Solver.hpp:96-106), weak
        # scaled like the headline.  A context of its own: the force and the clamped dt are fixed at creation.
        s.close()
        s = None
        try:
            cfg_t, _ = make_config(args, rank, world, blocks, gravity=9.8)
            cfg_t.device_id = local
            if world > 1:
                from cajitafluids_b200.distributed import attach_nccl
                attach_nccl(cfg_t, dist)
            st_ = Solver(cfg_t)
            st_.set_tuning("cg_variant", args.cg_variant_requested)
            st_.setup()
            st_.step()
            barrier()
            st_.reset_stats()
            t0 = time.perf_counter()
            nst = 3
            for _ in range(nst):
                st_.step()
            barrier()
            dt_steps = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt_steps], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt_steps = float(t[0])
            st2 = st_.stats()
            extra["timesteps_per_s_bench_grid"] = {
                "global_cells": list(gcells), "value": nst / dt_steps, "cg_iters_per_step": args.iters,
                "interp_order": 3, "gravity": 9.8, "dt": st_.dt, "ms_advect": st2["ms_advect"] / nst,
                "ms_add_inputs": st2["ms_add_inputs"] / nst, "ms_build_rhs": st2["ms_build_rhs"] / nst,
                "ms_pcg": st2["ms_pcg"] / nst, "ms_apply_pressure": st2["ms_apply_pressure"] / nst,
                "note": "BASELINE configs[4]: inflow source and body force (0, -9.8, 0) on, dt clamped to "
                        "h / (|v_in| + sqrt(g h)); projection capped at the headline's fixed CG iteration count (the "
                        "reference's tol / max_iter cannot converge at this size, SURVEY F5); wall clock between "
                        "barriers, max over ranks"}
            st_.close()
        except Exception as e:  # noqa: BLE001
            extra["timesteps_per_s_bench_grid"] = {"error": repr(e)[:300]}
    if world == 1 and args.timestep_cells > 0:
        if s is not None:
            s.close()
            s = None
        from cajitafluids_b200 import default_config
        c2 = default_config(3, args.timestep_cells)
        s2 = Solver(c2)
        s2.setup()
        for _ in range(2):
            s2.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nst = 5
        it0 = s2.stats()["cg_iterations"]
        for _ in range(nst):
            s2.step()
        torch.cuda.synchronize()
        dt_steps = time.perf_counter() - t0
        extra["timesteps_per_s"] = {"cells": [args.timestep_cells] * 3, "value": nst / dt_steps,
                                    "cg_iters_per_step": (s2.stats()["cg_iterations"] - it0) / nst,
                                    "interp_order": 3}
        s2.close()

    # extra: BASELINE.json configs[2] — 256^3 PCG-only, fixed 200 iterations, synthetic divergence right-hand side,
    # HBM GB/s against the roofline (same kernels and CG form as the headline; one vector = 134 MB ~ the L2 size)
    if world == 1 and not args.no_timestep:
        try:
            from cajitafluids_b200 import default_config
            n2 = max(16, args.cells // 2)
            c5 = default_config(3, n2, box=n2 / 512.0)
            c5.cg_fixed_iters = 200
            c5.cg_print_level = 0
            s5 = Solver(c5)
            s5.set_tuning("cg_variant", args.cg_variant_requested)  # (-1: the library's choice at THIS size)
            v5 = int(s5.stats()["cg_variant"])
            s5.fill_synthetic_velocity(0)
            s5.build_rhs()
            for _ in range(3):
                s5.pcg_fixed(200)
            ms5 = 0.0
            for _ in range(3):
                m, r5 = s5.pcg_fixed(200)
                ms5 += m
            gbs5 = n2 ** 3 * BYTES_ITER[v5] * 600 / (ms5 * 1e-3) / 1e9
            extra["config2_pcg_only"] = {"cells": [n2] * 3, "cg_iters_per_step": 200, "steps": 3,
                                         "iterations_per_s": 600 / (ms5 * 1e-3), "achieved_gbs": gbs5,
                                         "frac_of_peak": gbs5 / peak, "bytes_per_cell": BYTES_ITER[v5], "cg_variant": v5,
                                         "final_residual": r5,
                                         "note": "at 256^3 a vector (134 MB) is about the size of the L2: not an HBM-only number"}
            try:  # ncu DRAM bytes of the two kernels at this size, committed constants like roofline.traffic
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                ka = "stencil7_dot_tma MODE 1 (phase A')" if v5 == 2 else "cg_rupdate_kernel"
                extra["config2_pcg_only"]["ncu_dram_bytes_per_launch"] = {
                    "cg_fused_kernel": tj.get(f"cg_variant{v5}_dominant_{n2}"), ka: tj.get(f"cg_variant{v5}_rupdate_{n2}"),
                    "algorithmic": {"cg_fused_kernel": BYTES_DOMINANT[v5] * n2 ** 3, ka: 24 * n2 ** 3},
                    "source": tj.get(f"_source_cg_variant{v5}_{n2}",
                                     "profiles/r2_launches_cg%d.csv (ncu, L2 flushed before every launch)" % n2)}
            except Exception:  # noqa: BLE001
                pass
            s5.close()
        except Exception as e:  # noqa: BLE001
            extra["config2_pcg_only"] = {"error": repr(e)[:300]}

    # extra: the same workload with the OTHER two-kernel form (72 B/cell when the 64-byte form is the one that ran, and
    # the other way round), in a child process.  Same fixed iteration count: its final residual must equal ours bit for bit.
    if world == 1 and args.cg_variant in (1, 2) and not args.no_probe and not args.tune:
        other = 3 - args.cg_variant
        key = "cg_variant%d" % other
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--steps", "3", "--warmup", "3",
                   "--cells", str(args.cells), "--iters", str(args.iters), "--cg-variant", str(other), "--no-cpu-baseline",
                   "--no-e2e", "--no-timestep", "--no-probe"]
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
            if p.returncode != 0 or not p.stdout.strip():
                raise RuntimeError("child exit %d: %s" % (p.returncode, p.stderr.strip()[-200:]))
            child = json.loads(p.stdout.strip().splitlines()[-1])
            extra[key] = {
                "value": child["value"], "unit": UNIT, "ms_per_step": child["ms_per_step"],
                "iteration": child["roofline"]["iteration"], "dominant_kernel_frac": child["roofline"]["frac"],
                "final_residual": child["extra"]["final_residual"],
                "same_residual_as_headline_form": child["extra"]["final_residual"] == resid,
                "note": "%d B/cell two-kernel form, child process, 3 steps; bit-identical to the headline form, which is the "
                        "faster of the two at this size" % BYTES_ITER[other]}
        except Exception as e:  # noqa: BLE001
            extra[key] = {"error": repr(e)[:300]}

    # extra: time to solution of ONE projection (Jacobi vs the opt-in multigrid preconditioner) and whole
    # timesteps with multigrid — mg_side_measurements() in a CHILD process with a time limit: the multigrid
    # kernels were written after the round's GPU budget was spent, and neither a failure nor a hang there may
    # cost the headline line.
    if world == 1 and not args.no_timestep and not args.no_probe:
        run_child(extra, "projection_time_to_solution",
                  ["--side", "mg", "--cells", str(args.cells), "--timestep-cells", str(args.timestep_cells)], 240)
        # opt-in kernel options, each in its own child and under its own key
        run_child(extra, "advect_tile_probe", ["--side", "advect", "--cells", str(args.cells)], 120)
        run_child(extra, "flat_2d_probe", ["--side", "flat2d", "--cells", str(args.cells)], 180)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)
        # the same whole timesteps as extra.timesteps_per_s (default inflow problem, reference tol / max_iter,
        # cubic interpolation) on the host cores: setup, then ONE timed step (bounded: ~600 CG iterations at 128^3)
        if args.timestep_cells > 0:
            try:
                from cajitafluids_b200 import default_config
                from oracle_api import Oracle
                import oracle_api
                oracle_api.set_num_threads(0)
                o = Oracle(default_config(3, args.timestep_cells))
                o.set_accumulation(False)  # plain double sums: the reference's arithmetic and cost
                o.setup()
                t0 = time.perf_counter()
                it0 = o.stats()["cg_iterations"]
                o.step()
                dt_cpu = time.perf_counter() - t0
                cpu["timesteps_per_s"] = {"cells": [args.timestep_cells] * 3, "value": 1.0 / dt_cpu,
                                          "cg_iters_per_step": o.stats()["cg_iterations"] - it0, "interp_order": 3,
                                          "sample": "one whole timestep after setup"}
                o.close()
            except Exception as e:  # noqa: BLE001
                cpu["timesteps_per_s"] = {"error": repr(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args),
                           "cells_per_gpu": [args.cells] * 3 if args.scaling == "weak" else None,
                           "global_cells": list(gcells), "blocks": list(blocks), "cg_iters_per_step": args.iters,
                           "l2": "inputs larger than L2 (each vector %.2f GB)" % (ncell_local * 8 / 1e9),
                           "timing": "CUDA events on the launching stream around each solve, max over ranks",
                           "cg_form": {0: "three kernels, 88 B/cell", 1: "two kernels, 72 B/cell",
                                       2: "two kernels, q not stored, 64 B/cell",
                                       3: "single-reduction (Chronopoulos-Gear), two kernels, 88 B/cell, opt-in"}[args.cg_variant],
                           "exchange": ("none (1 GPU)" if world == 1 else
                                        "NVLink peer stores (cudaIpc): faces on the side stream under interior work, CG "
                                        "sums through mailboxes in the compute kernels' last blocks" if st["peer_overlap"] else
                                        "NVLink peer stores (cudaIpc), ghosts + CG sums in one kernel per reduction point"
                                        if st["peer_mode"] else "NCCL send/recv + all-gather"),
                           **({"exchange_fallback": fallback} if fallback else {})},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "extra": extra}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
