"""Multigrid-preconditioned projection of the default inflow problem with the fine level's smoothing sweeps on the TMA
z-march ("mg_tma" 1, the default in 3-D; "mg_tma_prolong": prolongation + first post-sweep too) and on the
one-thread-per-cell kernels (0): iterations, time per solve.
    python tools/profile_mg_tma.py [n ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

for n in [int(a) for a in sys.argv[1:]] or [512, 256, 128]:
    cfg = default_config(3, n)
    cfg.cg_max_iter = 20000
    s = Solver(cfg)
    s.set_preconditioner("mg")
    s.add_inputs()
    s.build_rhs()
    row = {"cells": n}
    for rep in range(2):
        for tma, prolong in ((1, 1), (1, 0), (0, 0)):
            s.set_tuning("mg_tma", tma)
            s.set_tuning("mg_tma_prolong", prolong)
            s.pcg_solve()
            s.reset_stats()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                it, res = s.pcg_solve()
            dt = (time.perf_counter() - t0) / reps
            st = s.stats()
            row[f"mg_tma{tma}_prolong{prolong}_rep{rep}"] = {"iterations": it, "residual": res, "ms_per_solve": dt * 1e3,
                                            "ms_per_iteration": dt * 1e3 / max(it, 1), "device_ms_pcg": st["ms_pcg"] / reps}
    print(json.dumps(row), flush=True)
    s.close()
