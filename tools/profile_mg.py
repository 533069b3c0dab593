"""Target for timing / ncu captures of the opt-in multigrid path and the output stage (1 GPU).

    python tools/profile_mg.py [n=256] [mode=solve|cycle|output]

solve : one projection of the default inflow problem with Jacobi and with MG (iterations, wall time,
        device time per phase)
cycle : 3 fixed MG-PCG iterations (the launch list of one V(2,2) cycle for ncu)
output: extraction kernel + copy, blocking form, and the deferred write
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cajitafluids_b200 import Solver, default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "solve"


def problem(kind, **kw):
    cfg = default_config(3, n)  # unit box: the default inflow lies inside it at every n
    cfg.cg_max_iter = 20000
    for k, v in kw.items():
        setattr(cfg, k, v)
    s = Solver(cfg)
    s.set_preconditioner(kind)
    s.add_inputs()
    s.build_rhs()
    return s


if mode == "solve":
    for kind in ("jacobi", "mg", "mg+graph", "mg+coarse", "mg+graph+coarse"):
        s = problem(kind.split("+")[0])
        if "graph" in kind:
            s.set_tuning("mg_graph", 1)
        if "coarse" in kind:
            s.set_tuning("mg_coarse_kernel", 1)
        s.pcg_solve()
        s.reset_stats()
        t0 = time.perf_counter()
        it, res = s.pcg_solve()
        dt = time.perf_counter() - t0
        st = s.stats()
        print(f"n={n} {kind}: {it} iterations, |r|={res:.3e}, wall {dt * 1e3:.2f} ms, device pcg {st['ms_pcg']:.2f} ms, "
              f"{st['kernel_launches']} launches, {dt * 1e3 / max(it, 1):.3f} ms/iteration")
        s.close()
elif mode == "cycle":
    s = problem("mg", cg_fixed_iters=3)
    print(s.pcg_solve())
elif mode == "output":
    s = problem("jacobi", cg_fixed_iters=20)
    s.setup()
    s.output()
    t0 = time.perf_counter()
    q, vel, nodes = s.output()
    print(f"n={n} output(): {1e3 * (time.perf_counter() - t0):.2f} ms blocking (kernel + D2H + host copies), "
          f"{(q.nbytes + vel.nbytes) / 1e9:.2f} GB")
    d = tempfile.mkdtemp()
    t0 = time.perf_counter()
    s.write_output(d, 0)
    t1 = time.perf_counter()
    s.step()
    t2 = time.perf_counter()
    s.output_flush()
    t3 = time.perf_counter()
    print(f"write_output returned after {1e3 * (t1 - t0):.2f} ms; step {1e3 * (t2 - t1):.2f} ms; flush (files) {1e3 * (t3 - t2):.2f} ms")
