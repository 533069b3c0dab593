"""bench.py end to end on the CPU: the script itself, unmodified, with the CUDA library replaced by the
host-emulated one (tests/emul) and the few torch.cuda calls it makes stubbed out, on a tiny grid.  Checks that
the one JSON line carries every key of the bench contract and that the numbers in it are self-consistent — not
the numbers themselves, which mean nothing here."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SHIM = r"""
import ctypes, sys, os
sys.path[:0] = [{root!r}, {tests!r}, {emul!r}]
import torch
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
import build_emul
import cajitafluids_b200
from cajitafluids_b200._capi import Context, Library
_lib = Library(build_emul.build(), "cfb_", mode=ctypes.RTLD_LOCAL)
class EmulSolver(Context):
    def __init__(self, cfg):
        super().__init__(_lib, cfg)
cajitafluids_b200.Solver = EmulSolver
sys.argv = ["bench.py"] + {argv!r}
import runpy
runpy.run_path(os.path.join({root!r}, "bench.py"), run_name="__main__")
"""


def run_bench(argv):
    code = SHIM.format(root=ROOT, tests=HERE, emul=os.path.join(HERE, "emul"), argv=argv)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def test_bench_line_carries_the_contract_keys():
    d = run_bench(["--cells", "32", "--iters", "6", "--steps", "2", "--warmup", "3", "--timestep-cells", "32", "--no-probe"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "extra"):
        assert k in d, k
    assert d["metric"] == "pcg_iterations_per_second" and d["unit"] == "iterations/s" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("pcg_32cubed_fixed6") and "l2" in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "algorithmic_bytes_per_cell",
              "algorithmic_bytes_per_launch", "avg_launch_ms", "iteration"):
        assert k in r, k
    # no CG form asked for: the library's choice, which is the 72-byte form at this size (phase B: 48 B/cell)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["algorithmic_bytes_per_cell"] == 48
    assert r["algorithmic_bytes_per_launch"] == 48 * 32 ** 3 and r["iteration"]["bytes_per_cell"] == 72
    assert "72 B/cell" in d["config"]["cg_form"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "iterations/s" and "sample" in c
    assert c["timesteps_per_s"]["value"] > 0 and c["timesteps_per_s"]["cells"] == [32, 32, 32]
    e = d["e2e"]
    assert e["unit"] == "iterations/s" and e["h2d_bytes_per_step"] == 8 * 32 ** 3 and e["d2h_bytes_per_step"] == 8 * 32 ** 3
    assert d["gpu_launches"] > 0
    x = d["extra"]
    assert x["cells_local"] == 32 ** 3 and "timesteps_per_s_bench_grid" in x and x["timesteps_per_s"]["cells"] == [32, 32, 32]
    assert x["config2_pcg_only"]["cells"] == [16, 16, 16] and x["config2_pcg_only"]["iterations_per_s"] > 0
    assert x["rhs_norm"] > 0
    ts = x["timesteps_per_s_bench_grid"]  # BASELINE configs[4]: body force on, dt clamp includes sqrt(g h)
    assert ts["gravity"] == 9.8 and ts["value"] > 0 and ts["dt"] < 1.0 / 512 and ts["ms_advect"] > 0
    assert "value_with_setup" in c and c["value_with_setup"] <= c["value"] * 1.05
    # the multigrid side measurements run in a child process of the real bench (skipped by --no-probe: a child
    # would not see the emulation shim); the same entry point is run here directly
    assert "projection_time_to_solution" not in x
    x.update(run_bench(["--side", "mg", "--cells", "32", "--timestep-cells", "32"]))
    x.update(run_bench(["--side", "advect", "--cells", "32"]))
    x.update(run_bench(["--side", "flat2d", "--cells", "32"]))
    assert x["advect_tile_probe"]["tile0_ms"] > 0 and x["advect_tile_probe"]["tile1_ms"] > 0
    fl = x["flat_2d_probe"]
    assert fl["flat0_variant1"]["residual"] == fl["flat1_variant1"]["residual"] == fl["flat1_variant2"]["residual"]
    tts = x["projection_time_to_solution"]
    for kind in ("jacobi", "mg", "mg_graph", "mg_graph_coarse"):
        assert "cg_iterations" in tts[kind], tts[kind]
    # (box edge = cells / 512 keeps h = 1/512 like the real run: at 32 cells the inflow lies outside the box and
    # the right-hand side is zero, so only the bookkeeping can be checked here)
    assert tts["mg"]["cg_iterations"] == tts["mg_graph"]["cg_iterations"] == tts["mg_graph_coarse"]["cg_iterations"]
    assert tts["mg"]["final_residual"] == tts["mg_graph_coarse"]["final_residual"]
    assert x["timesteps_per_s_mg"]["cg_iters_per_step"] < x["timesteps_per_s"]["cg_iters_per_step"] / 4


def test_bench_other_cg_forms_and_reference_arm():
    d2 = run_bench(["--cells", "32", "--iters", "6", "--steps", "1", "--warmup", "3", "--cg-variant", "2", "--no-cpu-baseline",
                    "--no-e2e", "--no-timestep", "--timestep-cells", "0"])
    d0 = run_bench(["--cells", "32", "--iters", "6", "--steps", "1", "--warmup", "3", "--cg-variant", "0", "--no-cpu-baseline",
                    "--no-e2e", "--no-timestep", "--timestep-cells", "0"])
    assert d2["roofline"]["iteration"]["bytes_per_cell"] == 64 and d2["roofline"]["algorithmic_bytes_per_cell"] == 40
    assert d0["roofline"]["iteration"]["bytes_per_cell"] == 88 and d0["roofline"]["algorithmic_bytes_per_cell"] == 16
    assert d2["extra"]["final_residual"] == d0["extra"]["final_residual"]  # the forms are bit-identical
    assert d2["cpu_baseline"] is None and d2["e2e"] is None
    # the reference arm needs neither GPU nor shim
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "32", "--iters", "4",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    assert r["impl"] == "reference" and r["metric"] == "pcg_iterations_per_second" and r["value"] > 0
    assert r["e2e"] == {"value": r["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["value"] == r["value"]
    # same workload label as the main arm, the honoured step count, every host core whatever the launcher exported
    assert r["config"]["workload"] == d2["config"]["workload"].replace("fixed6", "fixed4") and r["steps"] == 1
    assert r["config"]["cg_iters_per_step"] == 4 and 2 <= r["config"]["cg_iters_timed_per_step"] <= 4
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--cells", "32",
                        "--iters", "4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    r2 = json.loads(p.stdout.strip().splitlines()[-1])
    assert r2["steps"] == 2 and r2["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and r2["n_gpus"] == 2
    env["RANK"] = "1"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--cells", "32"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and not p.stdout.strip()  # the other ranks exit 0 without work
