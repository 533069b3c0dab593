"""The 72-byte and the 64-byte two-kernel CG forms against each other by block size (where `cg_variant_auto` draws its line).
    python tools/profile_forms.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

sizes = [int(a) for a in sys.argv[1:]] or [192, 224, 256, 320, 384, 448, 512]
for n in sizes:
    row = {"cells": n}
    for v in (1, 2):
        cfg = default_config(3, n, box=n / 512.0)
        s = Solver(cfg)
        s.set_tuning("cg_variant", v)
        s.set_tuning("cg_persist", 0)
        s.fill_synthetic_velocity(0)
        s.build_rhs()
        iters = 100
        for _ in range(3):
            s.pcg_fixed(iters)
        ms = 0.0
        for _ in range(3):
            m, res = s.pcg_fixed(iters)
            ms += m
        row["variant%d_iterations_per_s" % v] = 3 * iters / (ms * 1e-3)
        row["variant%d_residual" % v] = res
        s.close()
    row["ratio_64_over_72"] = row["variant2_iterations_per_s"] / row["variant1_iterations_per_s"]
    print(json.dumps(row), flush=True)
