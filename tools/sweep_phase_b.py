"""Phase B options against the shipped choice on one GPU: sixteen warps per CTA ("fused_nt" 512) and the top-down
walk ("fused_reverse"), in both two-kernel CG forms.
Usage: python tools/sweep_phase_b.py [n ...]   (CUDA-event times per phase)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cajitafluids_b200 import Solver, default_config

ITERS = 40
for n in [int(a) for a in sys.argv[1:]] or [512, 384]:
    s = Solver(default_config(3, n))
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)

    def run(label):
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st = s.stats()
        kt = max(1, st["k_timed_iters"])
        a, b = st["ms_k_axpy"] / kt, st["ms_k_stencil"] / kt
        print(f"n={n} {label}: {ms / ITERS * 1e3:8.1f} us/it  A {a * 1e3:7.1f} us  B {b * 1e3:7.1f} us  resid {res:.6e}",
              flush=True)

    for variant in (2, 1):
        s.set_tuning("cg_variant", variant)
        run(f"variant {variant} library's tiling")
    for k, v in (("fused_stages", 3), ("fused_ty", 16), ("fused_tx", 128), ("fused_zc", 64)):
        s.set_tuning(k, v)
    for rep in range(2):
        for variant in (2, 1):
            for nt in (256, 512):
                for rev in (0, 1):
                    s.set_tuning("cg_variant", variant)
                    s.set_tuning("fused_nt", nt)
                    s.set_tuning("fused_reverse", rev)
                    run(f"variant {variant} 128x16x3 fused_nt={nt} reverse={rev}")
    s.close()
