"""Short target for ncu captures of the advection kernel: set-up + 2 steps at n^3, projection capped at 2 CG iterations.

    python tools/profile_advect_target.py [n=512] [advect_tile=0]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = default_config(3, n, box=n / 512.0)
cfg.cg_fixed_iters = 2
s = Solver(cfg)
s.set_tuning("advect_tile", tile)
s.setup()
for _ in range(2):
    s.step()
print(f"n={n} advect_tile={tile} ms_advect={s.stats()['ms_advect'] / 2:.3f}")
