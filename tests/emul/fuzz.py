"""tests/emul/fuzz.py — randomized differential test: the host-emulated CUDA path against the oracle, bit for bit.

    python tests/emul/fuzz.py loops|tma FIRST_SEED NUMBER_OF_CASES

Every case draws dimension, (ragged or power-of-two) cell counts, wall types, body force, interpolation order,
quirk switches, fixed / converged solves, the CG form (three kernels / two kernels / 64-byte), the preconditioner
(Jacobi, multigrid with random sweep counts, with and without the single-CTA coarse kernel and the graph replay),
the flat_2d and advect_tile options,
seeded fields, then runs setup + one step on both sides and compares every field, the iteration count, the output
stage — and the error code if a solve does not converge.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import build_emul  # noqa: E402
from cajitafluids_b200 import config as K  # noqa: E402
from cajitafluids_b200._capi import CfbError, Context, Library  # noqa: E402
from helpers import fields_of, make_cfg, smooth_velocity  # noqa: E402
from oracle_api import Oracle  # noqa: E402


def one_case(lib, which, seed):
    """-> (ok, description)"""
    rng = np.random.default_rng(seed)
    dim = int(rng.choice([2, 3]))
    hi = (70 if which == "loops" else 40) if dim == 3 else 150
    cells = tuple(int(rng.integers(3, hi)) for _ in range(dim))
    if rng.random() < 0.3:
        cells = tuple(int(2 ** rng.integers(2, 6)) for _ in range(dim))
    h = 2.0 ** -int(rng.integers(4, 9))  # extents exactly divisible by the cell size (src/Mesh.hpp:56-64)
    box = tuple(c * h for c in cells)
    bt = [int(rng.choice([K.SOLID, K.SOLID, K.FREE])) for _ in range(2 * dim)]
    kw = dict(boundary_type=bt, interp_order=int(rng.choice([1, 3])), quirks=(bool(rng.integers(2)), bool(rng.integers(2))),
              body_force=tuple(float(x) for x in rng.uniform(-3, 3, size=3)), fixed_iters=int(rng.choice([0, 0, 7, 23])))
    variant = int(rng.choice([0, 1, 2]))
    prec = str(rng.choice(["jacobi", "jacobi", "mg"]))
    if rng.random() < 0.7:
        nu = (int(rng.integers(1, 4)),) * 2 + (int(rng.integers(1, 9)),)
    else:
        nu = (int(rng.integers(1, 4)), int(rng.integers(0, 3)), int(rng.integers(1, 5)))
    desc = (f"seed {seed}: dim={dim} cells={cells} walls={bt} order={kw['interp_order']} quirks={kw['quirks']} "
            f"fixed={kw['fixed_iters']} cg_variant={variant} prec={prec} nu={nu}")
    cfg = make_cfg(dim, cells, box=box, **kw)
    g, o = Context(lib, cfg), Oracle(cfg)
    g.set_tuning("cg_variant", variant)
    # later options draw from their own stream, so that the cases of earlier runs stay what they were
    rng2 = np.random.default_rng(seed + 1000003)
    flat, tile = int(rng2.integers(2)), int(rng2.integers(2))
    g.set_tuning("flat_2d", flat)
    g.set_tuning("advect_tile", tile)
    desc += f" flat_2d={flat} advect_tile={tile}"
    # round 2, second half: run length of the 2-D march, phase B tiling, persistent kernel, sixteen-warp phase B
    rng3 = np.random.default_rng(seed + 2000003)
    yc, tiling, persist = int(rng3.choice([0, 1, 2, 3, 7])), int(rng3.integers(6)), int(rng3.choice([-1, 0, 1]))
    tilings = [None, (64, 16, 3), (64, 8, 4), (128, 16, 3), (64, 8, 3), (128, 8, 4)]
    if which == "tma":
        if tilings[tiling]:
            for k, v in zip(("fused_tx", "fused_ty", "fused_stages"), tilings[tiling]):
                g.set_tuning(k, v)
            if tilings[tiling] == (128, 16, 3) and variant == 2 and not (dim == 2 and flat):
                g.set_tuning("fused_nt", int(rng3.choice([0, 256, 512])))
        if yc:
            g.set_tuning("fused_yc", yc)
        if tilings[tiling] in (None, (64, 16, 3), (64, 8, 4), (64, 8, 3), (128, 16, 3)):
            g.set_tuning("cg_persist", persist)
        desc += f" fused_yc={yc} tiling={tilings[tiling]} cg_persist={persist}"
    if prec == "mg":
        om = 0.0 if nu[0] == nu[1] else 0.7  # unsymmetric cycles: a damping that keeps CG going
        g.set_preconditioner("mg", *nu, om)
        o.set_preconditioner("mg", *nu, om)
        g.set_tuning("mg_coarse_kernel", int(rng.integers(2)))
        g.set_tuning("mg_graph", int(rng.integers(2)))
    for f, a in smooth_velocity(o, rng, amp=float(rng.uniform(0.1, 0.9))).items():
        g.set(f, a)
        o.set(f, a)
    q = rng.uniform(-1, 1, size=o.shape(K.QUANTITY))
    g.set(K.QUANTITY, q)
    o.set(K.QUANTITY, q)
    err = []
    for s in (g, o):
        try:
            s.setup()
            s.step()
            err.append(None)
        except CfbError as e:
            err.append(e.code)
    ok = err[0] == err[1]
    if ok and err[0] is None:
        ok = all(np.array_equal(g.get(f), o.get(f)) for f in fields_of(dim) + [K.PRESSURE]) and \
            g.stats()["cg_iterations"] == o.stats()["cg_iterations"]
        if ok:
            a, b = g.output(), o.output()
            ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    desc += f" -> errors {err[0]}/{err[1]}, {o.stats()['cg_iterations']} CG iterations"
    g.close()
    o.close()
    return ok, desc


def run(which, seed0, ncase, verbose=True):
    lib = Library(build_emul.build(tma=which == "tma"), "cfb_", mode=ctypes.RTLD_LOCAL)
    bad = []
    for seed in range(seed0, seed0 + ncase):
        try:
            ok, desc = one_case(lib, which, seed)
        except Exception as e:  # noqa: BLE001
            ok, desc = False, f"seed {seed}: {e!r}"
        if verbose or not ok:
            print(("ok   " if ok else "BAD  ") + desc, flush=True)
        if not ok:
            bad.append(desc)
    return bad


if __name__ == "__main__":
    failures = run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    print("bad:", len(failures))
    sys.exit(1 if failures else 0)
