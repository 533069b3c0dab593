"""Which CG form and which z chunk of phase A' by block size, with r of phase A' staged by TMA:
72-byte form against the 64-byte form on one GPU, cubes and two-dimensional grids.
Usage: python tools/sweep_forms2.py [3:n | 2:n ...]   (CUDA-event times per phase)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cajitafluids_b200 import Solver, default_config

ITERS = 40
cases = sys.argv[1:] or ["3:192", "3:224", "3:256", "3:320", "3:384", "3:448", "3:512", "2:4096", "2:8192"]
for case in cases:
    dim, n = (int(v) for v in case.split(":"))
    s = Solver(default_config(dim, n))
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)
    s.set_tuning("cg_persist", 0)

    def run(label, variant):
        s.set_tuning("cg_variant", variant)
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st = s.stats()
        kt = max(1, st["k_timed_iters"])
        a, b = st["ms_k_axpy"] / kt, st["ms_k_stencil"] / kt
        print(f"{dim}-D n={n} {label}: {ms / ITERS * 1e3:8.1f} us/it  A {a * 1e3:7.1f} us  B {b * 1e3:7.1f} us  "
              f"resid {res:.6e}", flush=True)

    run("variant 1 (72 B)", 1)
    run("variant 2 (64 B) library's tiling and chunk", 2)
    for tx, ty, st in [(64, 16, 4), (128, 16, 3), (128, 8, 4)]:
        for zc in ([16, 24, 32, 48, 64] if dim == 3 else [64]):
            for k, v in (("stencil_stages", st), ("stencil_ty", ty), ("stencil_tx", tx), ("stencil_zc", zc)):
                s.set_tuning(k, v)
            run(f"variant 2 (64 B) A' tile={tx}x{ty} st={st} zc={zc}", 2)
    s.close()
