"""Small two-dimensional grids (the reference's own dimensionality and sizes): the CG iteration with the launch-per-phase
kernels and with the persistent kernel ("cg_persist"), fixed iteration counts on the synthetic right-hand side.
    python tools/profile_small_2d.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

for n in [int(a) for a in sys.argv[1:]] or [256, 512, 768, 1024, 1536, 2048, 3072]:
    for persist in (0, 1):
        s = Solver(default_config(2, n))
        s.set_tuning("cg_persist", persist)
        s.fill_synthetic_velocity(0)
        s.build_rhs()
        iters = 400
        for _ in range(3):
            s.pcg_fixed(iters)
        ms, res = s.pcg_fixed(iters)
        print(json.dumps({"dim": 2, "cells": n, "cg_persist": persist, "us_per_iteration": ms * 1e3 / iters,
                          "iterations_per_s": iters / (ms * 1e-3), "achieved_gbs": n ** 2 * 72 * iters / (ms * 1e-3) / 1e9,
                          "residual": res}), flush=True)
        s.close()
