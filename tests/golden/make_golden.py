"""Generates the committed fixtures of tests/golden/ (run from the repo root:
`python tests/golden/make_golden.py`).

The reference itself cannot be built or imported here (Kokkos/Cabana/MPI absent), so these
are NOT outputs of the reference — parity stays "unpinned".  They pin what can be pinned:

  ref2d_default_n32.npz   the independent numpy/scipy restatement of the reference's 2-D step
                          (tests/ref2d_numpy.py): default inflow problem, 32^2, setup + 4 steps.
                          Tolerance fixture (different summation order): 1e-9 relative L2, CG iterations +-1.
  oracle_3d_n16.npz       the C++ oracle's own output for the 3-D extension (16^3, setup + 3 steps,
                          cubic interpolation) and
  oracle_2d_n24_gravity.npz  for a 2-D case with gravity and mixed walls: bit-exact regression fixtures
                          (both sides accumulate dot products exactly, so the CUDA path must reproduce
                          them bit for bit as well).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cajitafluids_b200 import config as K  # noqa: E402
from helpers import fields_of, make_cfg  # noqa: E402
from oracle_api import Oracle  # noqa: E402
from ref2d_numpy import Ref2D  # noqa: E402


def main():
    m = Ref2D(32)
    m.setup()
    its = [m.last_iters]
    for _ in range(4):
        m.step()
        its.append(m.last_iters)
    np.savez_compressed(os.path.join(HERE, "ref2d_default_n32.npz"), q=m.owned("q"), u=m.owned("u"), v=m.owned("v"),
                        p=m.owned("p"), cg_iterations=np.array(its), dt=m.dt, time=m.time)

    def oracle_case(name, cfg, dim, steps):
        o = Oracle(cfg)
        o.setup()
        its = [o.stats()["cg_iterations"]]
        for _ in range(steps):
            o.step()
            its.append(o.stats()["cg_iterations"])
        out = {f"f{f}": o.get(f) for f in fields_of(dim) + [K.PRESSURE]}
        np.savez_compressed(os.path.join(HERE, name), cg_iterations=np.array(its), time=o.time, **out)

    oracle_case("oracle_3d_n16.npz", make_cfg(3, 16), 3, 3)
    oracle_case("oracle_2d_n24_gravity.npz",
                make_cfg(2, 24, boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE], body_force=(0.0, -9.8, 0.0)), 2, 3)


if __name__ == "__main__":
    main()
