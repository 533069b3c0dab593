"""Sweep the tilings of the two-kernel CG iteration (kernels_fused.cu) on one GPU and compare with
the three-kernel form.  Usage: python tools/sweep_fused.py [n ...]
Prints per-kernel CUDA-event times and algorithmic GB/s (72 or 88 B/cell/iteration)."""
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
    cfg = default_config(3, n)
    s = Solver(cfg)
    s.fill_synthetic_velocity(0)
    s.build_rhs()
    s.set_tuning("time_kernels", 1)
    cells = n ** 3
    zcs = [29, 37, 64] if n == 256 else ([32, 64] if n > 256 else [8, 10, 16])
    combos = [(0, 0, 0, 0, 0)]
    # automatic tiling, with and without the top-down walk of phase B
    for rev in (1, 0):
        s.set_tuning("cg_variant", 1)
        s.set_tuning("fused_auto", 1)
        s.set_tuning("fused_reverse", rev)
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st_ = s.stats()
        kt = max(1, st_["k_timed_iters"])
        a, c = st_["ms_k_axpy"] / kt, st_["ms_k_stencil"] / kt
        print(f"n={n} auto tiling reverse={rev}: {ms / ITERS * 1e3:8.1f} us/it {ITERS * 1e3 / ms:7.1f} it/s  "
              f"iter {cells * 72 * ITERS / ms / 1e6:5.0f} GB/s ({cells * 72 * ITERS / ms / 1e6 / PEAK:.1%})  "
              f"A {a * 1e3:7.1f} us  B {c * 1e3:7.1f} us  resid {res:.6e}", flush=True)
    # phase A occupancy sweep at the automatic phase-B tiling
    for ctas in (2, 3, 4, 6):
        s.set_tuning("cg_variant", 1)
        s.set_tuning("rupdate_ctas", ctas)
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st_ = s.stats()
        kt = max(1, st_["k_timed_iters"])
        a = st_["ms_k_axpy"] / kt
        print(f"n={n} phase A ctas/SM={ctas}: {a * 1e3:7.1f} us -> {cells * 24 / a / 1e6:5.0f} GB/s "
              f"({cells * 24 / a / 1e6 / PEAK:.1%})   iteration {ms / ITERS * 1e3:8.1f} us", flush=True)
    s.set_tuning("rupdate_ctas", 3)
    for tx, ty, st in [(64, 16, 3), (64, 8, 4), (128, 16, 3)]:
        for zc in zcs:
            combos.append((1, tx, ty, st, zc))
    for var, tx, ty, st, zc in combos:
        s.set_tuning("cg_variant", var)
        if var == 1:
            s.set_tuning("fused_stages", st)
            s.set_tuning("fused_zc", zc)
            s.set_tuning("fused_tx", tx)
            s.set_tuning("fused_ty", ty)
        s.pcg_fixed(5)
        s.reset_stats()
        ms, res = s.pcg_fixed(ITERS)
        st_ = s.stats()
        kt = max(1, st_["k_timed_iters"])
        bpc = 72 if var == 1 else 88
        a, b, c = st_["ms_k_axpy"] / kt, st_["ms_k_pupdate"] / kt, st_["ms_k_stencil"] / kt
        dom = cells * (48 if var == 1 else 16) / c / 1e6
        print(f"n={n} form={'2k' if var else '3k'} tile={tx}x{ty} st={st} zc={zc}: {ms / ITERS * 1e3:8.1f} us/it "
              f"{ITERS * 1e3 / ms:7.1f} it/s  iter {cells * bpc * ITERS / ms / 1e6:5.0f} GB/s "
              f"({cells * bpc * ITERS / ms / 1e6 / PEAK:.1%})  A {a * 1e3:7.1f} us  mid {b * 1e3:6.1f} us  "
              f"B {c * 1e3:7.1f} us -> {dom:5.0f} GB/s ({dom / PEAK:.1%})  resid {res:.6e}", flush=True)
    s.close()
