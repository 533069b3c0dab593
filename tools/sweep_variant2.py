"""Sweep of the 64-byte CG iteration (cg_variant 2) on one GPU against the 72-byte default: phase A' tile
shape / stages / z chunk (the stencil kernel's tuning keys), phase B tiling as picked automatically.
Usage: python tools/sweep_variant2.py [n ...]   (CUDA-event times per phase, algorithmic GB/s)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cajitafluids_b200 import Solver, default_config

PEAK = 6550.1
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

ITERS = 40
sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
for n in sizes:
    s = Solver(default_config(3, n))
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)
    cells = n ** 3

    def run(label, variant, bpc, bpa, bpb):
        s.set_tuning("cg_variant", variant)
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st = s.stats()
        kt = max(1, st["k_timed_iters"])
        a, b = st["ms_k_axpy"] / kt, st["ms_k_stencil"] / kt
        print(f"n={n} {label}: {ms / ITERS * 1e3:8.1f} us/it {ITERS * 1e3 / ms:7.1f} it/s  "
              f"iter {cells * bpc * ITERS / ms / 1e6:5.0f} GB/s ({cells * bpc * ITERS / ms / 1e6 / PEAK:.1%})  "
              f"A {a * 1e3:7.1f} us -> {cells * bpa / a / 1e6:5.0f} GB/s ({cells * bpa / a / 1e6 / PEAK:.1%})  "
              f"B {b * 1e3:7.1f} us -> {cells * bpb / b / 1e6:5.0f} GB/s ({cells * bpb / b / 1e6 / PEAK:.1%})  "
              f"resid {res:.6e}", flush=True)

    run("variant 1 (72 B)", 1, 72, 24, 48)
    # phase A' with r streamed by 128-bit loads (the round-2 form) at its shipped tiling, then with r staged by TMA
    # ("stencil_rtma" existed for this measurement only — profiles/r2_sweep_rtma.log — and went with the 128-bit-load form)
    for rtma, tilings in ((1, [(64, 16, 4), (64, 8, 4), (128, 8, 4), (128, 16, 4), (128, 16, 3), (128, 32, 3), (64, 16, 6),
                               (64, 32, 4), (64, 32, 3)]),):
        for tx, ty, st in tilings:
            for zc in ([32, 64, 128] if n >= 256 else [8, 16, 32]):
                for k, v in (("stencil_stages", st), ("stencil_ty", ty), ("stencil_tx", tx), ("stencil_zc", zc)):
                    s.set_tuning(k, v)
                run(f"variant 2 (64 B) rtma={rtma} A' tile={tx}x{ty} st={st} zc={zc}", 2, 64, 24, 40)
    s.close()
