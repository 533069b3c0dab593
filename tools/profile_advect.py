"""Advection kernel (TimeIntegrator::step: RK3 back-trace + cubic B-spline sample of q, u, v, w in one launch) with
the row mapping and with 32 x 2 x 2 entity tiles ("advect_tile"), n^3 cells, projection capped at a few CG iterations.

    python tools/profile_advect.py [n=512] [steps=3]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for tile in (0, 1):
    cfg = default_config(3, n, box=n / 512.0)
    cfg.cg_fixed_iters = 5
    cfg.cg_print_level = 0
    s = Solver(cfg)
    s.set_tuning("advect_tile", tile)
    s.setup()
    s.step()
    s.reset_stats()
    for _ in range(steps):
        s.step()
    st = s.stats()
    print(json.dumps({"cells": [n] * 3, "advect_tile": tile, "ms_advect": st["ms_advect"] / steps,
                      "entities_per_s": 4 * n ** 3 / (st["ms_advect"] / steps * 1e-3)}), flush=True)
    s.close()
