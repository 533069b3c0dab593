"""Two-dimensional grids (the reference's own dimensionality): phase B tilings of the 72-byte form on one GPU.  In 2-D
there is ONE plane: a unit is a run of "fused_yc" tile rows marching along y through the TMA ring (fused_yc = 1 is the
round-2 form, one tile per block, which sat at 2.7 TB/s whatever the tile shape: profiles/r2_sweep_2d_tilings.log).
Usage: python tools/sweep_2d_tilings.py [n ...]   (CUDA-event times per phase)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cajitafluids_b200 import Solver, default_config

ITERS = 40
for n in [int(a) for a in sys.argv[1:]] or [8192, 4096, 2048, 1024]:
    s = Solver(default_config(2, n))
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)
    s.set_tuning("cg_variant", 1)
    s.set_tuning("cg_persist", 0)

    def run(label):
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st = s.stats()
        kt = max(1, st["k_timed_iters"])
        a, b = st["ms_k_axpy"] / kt, st["ms_k_stencil"] / kt
        cells = n * n
        print(f"2-D n={n} {label}: {ms / ITERS * 1e3:8.1f} us/it  A {a * 1e3:7.1f} us  B {b * 1e3:7.1f} us "
              f"({cells * 48 / b / 1e6:5.0f} GB/s)  resid {res:.6e}", flush=True)

    run("library's tiling and run length")
    for tx, ty, st in [(128, 16, 3), (64, 16, 3), (128, 8, 3), (64, 8, 4), (64, 32, 2)]:
        for k, v in (("fused_stages", st), ("fused_ty", ty), ("fused_tx", tx)):
            s.set_tuning(k, v)
        for yc in (1, 4, 8, 16, 32, 64):
            s.set_tuning("fused_yc", yc)
            run(f"tile={tx}x{ty} st={st} fused_yc={yc}")
    s.close()
