"""The multigrid projection at n^3 with the V-cycle replayed as a CUDA graph ("mg_graph") and the fine level on the TMA
march — the combination the bench's side measurement runs at 512^3.   python tools/check_mg_graph_512.py [n=512]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
out = {"cells": n}
for graph in (1, 0):
    cfg = default_config(3, n)
    cfg.cg_max_iter = 20000
    s = Solver(cfg)
    s.set_tuning("mg_graph", graph)
    s.set_preconditioner("mg")
    s.add_inputs()
    s.build_rhs()
    s.pcg_solve()
    t0 = time.perf_counter()
    it, res = s.pcg_solve()
    out[f"mg_graph{graph}"] = {"iterations": it, "residual": res, "ms_per_solve": (time.perf_counter() - t0) * 1e3}
    s.close()
print(json.dumps(out))
