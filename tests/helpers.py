"""Shared helpers for the parity tests: seeded inputs, paired (CUDA, oracle) contexts, norms."""
import numpy as np

from cajitafluids_b200 import config as K


def rel_l2(a, b):
    """relative L2 difference ||a-b|| / ||b||  (0 when both are zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b.ravel())
    nd = np.linalg.norm((a - b).ravel())
    if nb == 0.0:
        return 0.0 if nd == 0.0 else np.inf
    return nd / nb


def eigen_tol(n, dim=3):
    """Bound on the relative L2 error of A p against lambda p for the smoothest non-constant Neumann eigenvector
    p = prod cos(2 pi (g + 1/2) / n) of the (2 dim + 1)-point operator in double arithmetic.  The row sum
    diag * p - sum of neighbours cancels to lambda / diag = dim (2 - 2 cos(2 pi / n)) / (2 dim) ~ (2 pi / n)^2 / 2
    of its terms, so the rounding error relative to the result grows like n^2 (4.6e-13 at 256^3, 2.6e-12 at 512^3
    measured): a fixed bound is wrong at full size.  8 eps (2 dim) / (dim (2 - 2 cos(2 pi / n))), i.e. 2.4e-11 at
    n = 512 — checked against plain numpy in tests/test_oracle_kat.py::test_eigenvector_bound_scales_with_n."""
    return 8.0 * np.finfo(np.float64).eps * (2 * dim) / (dim * (2.0 - 2.0 * np.cos(2.0 * np.pi / n)))


def fields_of(dim):
    return [K.QUANTITY, K.U, K.V] + ([K.W] if dim == 3 else [])


def make_cfg(dim, cells, **kw):
    bt = kw.pop("boundary_type", None)
    fixed = kw.pop("fixed_iters", 0)
    tol = kw.pop("tol", None)
    max_iter = kw.pop("max_iter", None)
    force = kw.pop("body_force", None)
    cfg = K.default_config(dim, cells, **kw)
    if bt is not None:
        for i, b in enumerate(bt):
            cfg.boundary_type[i] = b
    cfg.cg_fixed_iters = fixed
    if tol is not None:
        cfg.cg_tolerance = tol
    if max_iter is not None:
        cfg.cg_max_iter = max_iter
    if force is not None:
        for d, f in enumerate(force):
            cfg.body_force[d] = f
    return cfg


def smooth_velocity(ctx, rng, amp=1.0, extent=None):
    """A smooth random MAC velocity (owned faces) with |u| <= amp: low-order sin/cos modes with seeded random
    coefficients, wall-compatible (zero normal component on the walls) on the unit box — or on a box with
    edge lengths `extent` and its low corner at the origin when that is given.  Returns {field: array[z,y,x]}."""
    dim = ctx.dim
    h = ctx.cell_size
    off = ctx.global_offset()
    out = {}
    for d in range(dim):
        f = K.U + d
        shp = ctx.shape(f)  # (z, y, x)
        ext = shp[::-1]
        coords = []
        for e in range(dim):
            g = np.arange(ext[e]) + off[e]
            coords.append((g * h if e == d else (g + 0.5) * h) / (1.0 if extent is None else extent[e]))
        grids = np.meshgrid(*coords[::-1], indexing="ij")[::-1]  # x, y, (z) each shaped (z,y,x)/(y,x)
        val = np.zeros(grids[0].shape)
        for _ in range(3):
            k = rng.integers(1, 4, size=dim)
            ph = rng.uniform(0, 2 * np.pi, size=dim)
            term = rng.uniform(-1, 1)
            for e in range(dim):
                term = term * (np.sin(np.pi * k[e] * grids[e]) if e == d else np.cos(np.pi * k[e] * grids[e] + ph[e]))
            val = val + term
        val = amp * val / max(np.abs(val).max(), 1e-30)
        out[f] = val.reshape(shp)
    return out


def random_cells(ctx, rng, field=K.QUANTITY):
    return rng.uniform(-1.0, 1.0, size=ctx.shape(field))


def set_both(gpu, ora, field, arr, version=K.CURRENT, region=K.OWNED):
    gpu.set(field, arr, version, region)
    ora.set(field, arr, version, region)


TOL_FIELD = 1e-10  # north_star bar: fields within 1e-10 relative L2 in FP64


def assert_same(a, b, what=""):
    """The stated bar (1e-10 relative L2) AND the stronger property this implementation has: both
    sides compute every reduction as a correctly rounded exact sum and every element-wise update
    with the same rounding sequence, so whole runs are bit-identical."""
    e = rel_l2(a, b)
    assert e < TOL_FIELD, f"{what}: rel l2 {e}"
    assert np.array_equal(a, b), f"{what}: within tolerance (rel l2 {e}) but not bit-identical"
