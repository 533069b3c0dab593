"""Small and mid grids (BASELINE configs[0] - [2]): the CG iteration with the launch-per-phase kernels and with the persistent
kernel ("cg_persist"), fixed iteration counts on the synthetic right-hand side, and whole timesteps of the default
problem.   python tools/profile_small.py [n ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

sizes = [int(a) for a in sys.argv[1:]] or [64, 96, 128, 160, 192, 256]
for n in sizes:
    for persist in (0, 1):
        cfg = default_config(3, n)
        s = Solver(cfg)
        s.set_tuning("cg_persist", persist)
        s.fill_synthetic_velocity(0)
        s.build_rhs()
        iters = 400
        for _ in range(3):
            s.pcg_fixed(iters)
        ms, res = s.pcg_fixed(iters)
        out = {"cells": n, "cg_persist": persist, "us_per_iteration": ms * 1e3 / iters, "iterations_per_s": iters / (ms * 1e-3),
               "achieved_gbs": n ** 3 * 72 * iters / (ms * 1e-3) / 1e9, "residual": res}
        s.close()
        if n <= 160:
            s = Solver(default_config(3, n))
            s.set_tuning("cg_persist", persist)
            s.setup()
            s.step()
            it0 = s.stats()["cg_iterations"]
            t0 = time.perf_counter()
            for _ in range(5):
                s.step()
            dt = time.perf_counter() - t0
            out.update({"timesteps_per_s": 5 / dt, "cg_iters_per_step": (s.stats()["cg_iterations"] - it0) / 5})
            s.close()
        print(json.dumps(out), flush=True)
