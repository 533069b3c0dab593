"""Generates tests/golden/refrun_*.npz FROM THE REFERENCE ITSELF (run from the repo root, in a
container that has /root/reference):   python tests/golden/make_golden_ref.py

Each fixture is the output of oracle/_ref/libcfref.so — the unmodified cajitafluids sources compiled
against the single-rank stand-ins of oracle/refshim/ (see its README) — for one case of
refrun_cases.py: setup + N steps of Solver<2>, then
    q, u, v, p, rhs          owned entities of the final state, [y, x]
    cg_iterations            per pressure solve (setup first)
    resid_last               CG residual history of the last solve
    dt, time, cell           Solver::_dt (clamped), Solver::_time, Mesh::cellSize
    out_q, out_vel, out_nodes_x/y   what SiloWriter::siloWrite hands to Silo for the final state
                             (src/SiloWriter.hpp:109-186, captured by the silo.h stand-in): owned quantity,
                             cell-centred velocity [component, y, x], node coordinates
    exact_*                  the same run with the stand-in CG in the oracle's arithmetic (fused
                             multiply-adds + exactly accumulated dot products): BIT-EXACT fixtures
    plain_*                  the same run in plain double arithmetic (serial sums, no fused
                             multiply-adds): TOLERANCE fixtures (CG iterations +-1, 1e-10 relative L2)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import reference_api  # noqa: E402
from refrun_cases import CASES, FIELD_NAMES  # noqa: E402


def run(cfg, steps, exact):
    r = reference_api.Reference(cfg, exact=exact)
    r.setup()
    its = [r.stats()["cg_iterations"]]
    for _ in range(steps):
        r.step()
        its.append(r.stats()["cg_iterations"])
    out = {FIELD_NAMES[f]: r.get(f)[0] for f in FIELD_NAMES}
    out["cg_iterations"] = np.diff([0] + its)
    out["resid_last"] = r.residual_history()
    cell, dt, time = r.scalars()
    out["dt"], out["time"], out["cell"] = dt, time, cell
    q, vel, nodes = r.output()
    out["out_q"], out["out_vel"], out["out_nodes_x"], out["out_nodes_y"] = q[0], vel[:, 0], nodes[0], nodes[1]
    r.close()
    return out


def main():
    if not reference_api.available():
        raise SystemExit("needs /root/reference (make -C oracle ref)")
    for name, (mk, steps) in CASES.items():
        data = {}
        for tag, exact in (("exact", True), ("plain", False)):
            for k, v in run(mk(), steps, exact).items():
                data[f"{tag}_{k}"] = v
        path = os.path.join(HERE, f"refrun_{name}.npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path), "bytes", list(data["exact_cg_iterations"]))


if __name__ == "__main__":
    main()
