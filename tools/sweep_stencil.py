"""Sweep stencil kernel variants / tilings and the CG streaming kernels on one GPU.
Usage: python tools/sweep_stencil.py [n ...]   -> prints GB/s (algorithmic bytes / CUDA-event time)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cajitafluids_b200 import Solver, config as K, default_config

PEAK = 6550.1
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
for n in sizes:
    cfg = default_config(3, n)
    cfg.cg_fixed_iters = 50
    s = Solver(cfg)
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    cells = n ** 3
    ms, res = s.pcg_fixed(50)
    print(f"n={n} pcg_fixed(50): {ms:.2f} ms -> {50e3 / ms:.1f} it/s, {cells * 88 * 50 / ms / 1e6:.0f} GB/s alg "
          f"({cells * 88 * 50 / ms / 1e6 / PEAK:.2%} of measured), resid {res:.3e}", flush=True)
    combos = [(0, 64, 16, 4, 64), (0, 64, 16, 4, 32), (0, 64, 16, 4, 128), (0, 64, 16, 4, n), (0, 64, 16, 6, 64),
              (0, 64, 8, 4, 64), (0, 64, 32, 3, 64), (0, 64, 32, 4, 64), (0, 128, 16, 3, 64), (0, 128, 16, 4, 64),
              (0, 128, 8, 4, 64), (0, 128, 32, 3, 64), (0, 128, 16, 4, 16), (0, 64, 32, 4, 16),
              (1, 64, 8, 0, 64), (1, 64, 8, 0, 16), (1, 64, 8, 0, n)]
    for var, tx, ty, st, zc in combos:
        s.set_tuning("stencil_variant", var)
        if var == 0:
            s.set_tuning("stencil_stages", st)
            s.set_tuning("stencil_tx", tx)
            s.set_tuning("stencil_ty", ty)
        s.set_tuning("stencil_zc", zc)
        dot, ms = s.stencil_dot(20)
        gbs = cells * 16 / ms / 1e6
        print(f"n={n} variant={var} tile={tx}x{ty} stages={st} zc={zc}: {ms * 1e3:.1f} us  {gbs:.0f} GB/s  "
              f"{gbs / PEAK:.2%} of measured  dot={dot:.6e}", flush=True)
    s.close()
