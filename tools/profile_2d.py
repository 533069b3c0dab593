"""Two-dimensional runs (the reference's own dimensionality): fixed CG iterations at n^2 with and without the
ghost-plane loads ("flat_2d"), every CG form.  Prints one JSON line per combination.

    python tools/profile_2d.py [n=8192] [iters=50]

Algorithmic bytes per cell are those of the 3-D forms (88 / 72 / 64); without flat_2d the TMA kernels also load the
two zero ghost planes of p (and r)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
BYTES = {0: 88, 1: 72, 2: 64}
ref = None
for flat in (0, 1):
    for variant in (1, 2, 0):
        cfg = default_config(2, n, box=n / 512.0)
        cfg.cg_print_level = 0
        s = Solver(cfg)
        s.set_tuning("flat_2d", flat)
        s.set_tuning("cg_variant", variant)
        s.fill_synthetic_velocity(0)
        s.build_rhs()
        for _ in range(3):
            s.pcg_fixed(iters)
        ms, res = s.pcg_fixed(iters)
        ref = res if ref is None else ref
        print(json.dumps({"cells": [n, n], "flat_2d": flat, "cg_variant": variant, "iters": iters, "ms": ms,
                          "iterations_per_s": iters / (ms * 1e-3), "gbs": n * n * BYTES[variant] * iters / (ms * 1e-3) / 1e9,
                          "residual": res, "same_residual": res == ref}), flush=True)
        s.close()
