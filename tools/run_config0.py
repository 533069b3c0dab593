"""BASELINE configs[0] as written: examples/advection at 64^3, single rank, local matrix-free Jacobi-PCG with the
reference's tolerance / max_iter, linear field interpolation, 100 RK3 steps (t_final = 0.5, dt = 0.005;
src/Solver.hpp:162-175) — the CUDA path through the C ABI against the CPU oracle, step by step.

    python tools/run_config0.py [n=64] [steps=100] > profiles/r2_config0_64cubed_100steps.json

Prints one JSON object: per-step CG iteration counts of both sides, the relative L2 difference of every field at
the end (and every 10 steps), bit-identity, wall times."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from cajitafluids_b200 import Solver, config as K, default_config
import oracle_api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = default_config(3, n, interp_order=1)
oracle_api.set_num_threads(0)
gpu, ora = Solver(cfg), oracle_api.Oracle(cfg)
fields = {"q": K.QUANTITY, "u": K.U, "v": K.V, "w": K.W, "p": K.PRESSURE}


def rel(a, b):
    nb = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / nb) if nb > 0 else float(np.abs(a).max() > 0)


def compare():
    return {k: rel(gpu.get(f), ora.get(f)) for k, f in fields.items()}, \
        all(np.array_equal(gpu.get(f), ora.get(f)) for f in fields.values())


t_gpu = t_cpu = 0.0
t0 = time.perf_counter()
gpu.setup()
t_gpu += time.perf_counter() - t0
t0 = time.perf_counter()
ora.setup()
t_cpu += time.perf_counter() - t0
its_g, its_o = [gpu.stats()["cg_iterations"]], [ora.stats()["cg_iterations"]]
checkpoints = []
for s in range(steps):
    t0 = time.perf_counter()
    gpu.step()
    t_gpu += time.perf_counter() - t0
    t0 = time.perf_counter()
    ora.step()
    t_cpu += time.perf_counter() - t0
    its_g.append(gpu.stats()["cg_iterations"])
    its_o.append(ora.stats()["cg_iterations"])
    if (s + 1) % 10 == 0 or s + 1 == steps:
        r, same = compare()
        checkpoints.append({"step": s + 1, "rel_l2": r, "bit_identical": same})
dg, do = np.diff([0] + its_g), np.diff([0] + its_o)
out = {"config": "BASELINE configs[0]: %d^3, single rank, Jacobi-PCG tol 1e-6 abs / max_iter 2000, linear field "
                 "interpolation, %d RK3 steps of dt = %g" % (n, steps, gpu.dt),
       "time_reached": gpu.time, "oracle_time_reached": ora.time,
       "cg_iterations_gpu": [int(v) for v in dg], "cg_iterations_oracle": [int(v) for v in do],
       "max_iteration_difference": int(np.abs(dg - do).max()), "iterations_equal": bool((dg == do).all()),
       "checkpoints": checkpoints, "final_rel_l2": checkpoints[-1]["rel_l2"],
       "max_rel_l2": max(max(c["rel_l2"].values()) for c in checkpoints),
       "bit_identical_at_every_checkpoint": all(c["bit_identical"] for c in checkpoints),
       "wall_s": {"gpu_path": t_gpu, "oracle_cpu": t_cpu, "oracle_cores": ora.num_threads()},
       "bar": "north_star: CG iterations within +-1, fields within 1e-10 relative L2"}
out["meets_bar"] = out["max_iteration_difference"] <= 1 and out["max_rel_l2"] < 1e-10
print(json.dumps(out))
