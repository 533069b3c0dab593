"""Short target for ncu captures: a few fixed CG iterations at n^3 (default 512).

    python tools/profile_target.py [n=512] [iters=6] [cg_variant=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = default_config(3, n)
s = Solver(cfg)
s.set_tuning("cg_variant", variant)
s.fill_synthetic_velocity(0)
s.build_rhs()
ms, res = s.pcg_fixed(iters)
print(f"n={n} iters={iters} cg_variant={variant} ms={ms:.3f} resid={res:.6e}")
