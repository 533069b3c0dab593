"""ctypes binding of the C ABI declared in include/cfb.h.

`Library(path, prefix)` binds one shared object exporting `<prefix>create`, `<prefix>step`, ...
The product binds `lib/libcfb_b200.so` with prefix `cfb_` (see `load()`); it has no CPU
fallback: a missing library or a box without an sm_100 GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import config as K

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcfb_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class CfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cfb status {code}] {msg}")
        self.code = code


class Library:
    """One loaded shared object implementing the cfb ABI under `prefix`."""

    # name -> (argtypes)   (restype is always int unless listed in _RET)
    _SIG = {
        "create": [C.POINTER(K.Config), C.POINTER(C.c_void_p)],
        "destroy": [C.c_void_p],
        "get_scalars": [C.c_void_p, _dp, _dp, _dp],
        "owned_extent": [C.c_void_p, C.c_int, _ip],
        "global_offset": [C.c_void_p, _ip],
        "upload": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p],
        "download": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p],
        "advance": [C.c_void_p, C.c_int],
        "add_inputs": [C.c_void_p],
        "time_integrator_step": [C.c_void_p],
        "build_rhs": [C.c_void_p],
        "pcg_solve": [C.c_void_p, _ip, _dp],
        "apply_pressure": [C.c_void_p],
        "correct_velocity": [C.c_void_p, _ip, _dp],
        "setup": [C.c_void_p],
        "step": [C.c_void_p],
        "solve": [C.c_void_p, C.c_double, C.c_int, _ip],
        "stencil_dot": [C.c_void_p, C.c_int, _dp, _dp],
        "get_stats": [C.c_void_p, C.POINTER(K.Stats)],
        "reset_stats": [C.c_void_p],
        "residual_history": [C.c_void_p, _dp, C.c_int, _ip],
        "output_extract": [C.c_void_p, _dp, _dp, _dp, _dp, _dp],
    }
    # product-only entry points (absent from the checker library)
    _SIG_OPT = {
        "default_config": [C.POINTER(K.Config), C.c_int],
        "partition": [C.c_int, C.c_int, C.c_int, _ip, _ip],
        "nccl_unique_id": [C.c_void_p],
        "field_ptr": [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                      C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "gather": [C.c_void_p, C.c_int],
        "pcg_solve_host": [C.c_void_p, C.c_void_p, C.c_void_p, _ip, _dp],
        "pcg_fixed": [C.c_void_p, C.c_int, _dp, _dp],
        "fill_synthetic_velocity": [C.c_void_p, C.c_int, C.c_uint64],
        "set_tuning": [C.c_void_p, C.c_char_p, C.c_int],
        "write_output": [C.c_void_p, C.c_char_p, C.c_int],
        "output_flush": [C.c_void_p],
        "set_output_dir": [C.c_void_p, C.c_char_p],
        "write_npy": [C.c_char_p, _dp, C.c_int, C.POINTER(C.c_int64)],
        "set_preconditioner": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double],
        "mg_apply": [C.c_void_p, _dp, _dp],
        "set_mg_max_levels": [C.c_void_p, C.c_int],
        "mg_num_levels": [C.c_void_p, _ip],
        "abi_version": [],
    }

    def __init__(self, path, prefix, mode=C.RTLD_GLOBAL):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path, mode=mode)
        self.fn = {}
        for table, required in ((self._SIG, True), (self._SIG_OPT, False)):
            for name, args in table.items():
                try:
                    f = getattr(self.dll, prefix + name)
                except AttributeError:
                    if required:
                        raise
                    continue
                f.argtypes, f.restype = args, C.c_int
                self.fn[name] = f
        le = getattr(self.dll, prefix + "last_error")
        le.argtypes, le.restype = [C.c_void_p], C.c_char_p
        self._last_error = le

    def last_error(self, ctx=None):
        s = self._last_error(ctx)
        return s.decode() if s else ""

    def check(self, rc, ctx=None):
        if rc != K.OK:
            raise CfbError(rc, self.last_error(ctx))


class Context:
    """One rank's solver state (== the reference's Solver + ProblemManager + VelocityCorrector)."""

    def __init__(self, lib, cfg):
        self.lib, self.cfg = lib, K.copy_config(cfg)
        self.h = C.c_void_p()
        rc = lib.fn["create"](C.byref(self.cfg), C.byref(self.h))
        if rc != K.OK:
            msg = lib.last_error(self.h if self.h else None)
            if self.h:
                lib.fn["destroy"](self.h)
                self.h = C.c_void_p()
            raise CfbError(rc, msg)
        self.dim = cfg.dim
        self.halo = cfg.halo_cell_width

    # -- lifecycle
    def close(self):
        if self.h:
            self.lib.fn["destroy"](self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _call(self, name, *args):
        self.lib.check(self.lib.fn[name](self.h, *args), self.h)

    # -- scalars / geometry
    def scalars(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._call("get_scalars", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    @property
    def cell_size(self):
        return self.scalars()[0]

    @property
    def dt(self):
        return self.scalars()[1]

    @property
    def time(self):
        return self.scalars()[2]

    def owned_extent(self, field):
        e = (C.c_int * 3)()
        self._call("owned_extent", field, e)
        return tuple(e)

    def global_offset(self):
        e = (C.c_int * 3)()
        self._call("global_offset", e)
        return tuple(e)

    def shape(self, field, region=K.OWNED):
        """numpy shape (z, y, x) of a dense host copy of `region` (z dropped to 1 in 2-D)."""
        if region == K.OWNED:
            e = self.owned_extent(field)
        else:
            # Cajita Ghost index space: owned cells + 2*halo, +1 along the face normal on every block
            e = list(self.owned_extent(K.QUANTITY))
            for d in range(self.dim):
                e[d] += 2 * self.halo + (1 if field in (K.U, K.V, K.W) and field - K.U == d else 0)
        return (e[2], e[1], e[0])

    # -- ProblemManager::get as host copies
    def get(self, field, version=K.CURRENT, region=K.OWNED):
        out = np.empty(self.shape(field, region), dtype=np.float64)
        self._call("download", field, version, region, out.ctypes.data_as(C.c_void_p))
        return out

    def set(self, field, array, version=K.CURRENT, region=K.OWNED):
        a = np.ascontiguousarray(array, dtype=np.float64)
        if a.shape != self.shape(field, region):
            raise ValueError(f"shape {a.shape} != {self.shape(field, region)}")
        self._call("upload", field, version, region, a.ctypes.data_as(C.c_void_p))

    def advance(self, field):
        self._call("advance", field)

    def gather(self, version=K.CURRENT):
        self._call("gather", version)

    # -- hot path
    def add_inputs(self):
        self._call("add_inputs")

    def time_integrator_step(self):
        self._call("time_integrator_step")

    def build_rhs(self):
        self._call("build_rhs")

    def pcg_solve(self):
        it, res = C.c_int(), C.c_double()
        self._call("pcg_solve", C.byref(it), C.byref(res))
        return it.value, res.value

    def apply_pressure(self):
        self._call("apply_pressure")

    def correct_velocity(self):
        it, res = C.c_int(), C.c_double()
        self._call("correct_velocity", C.byref(it), C.byref(res))
        return it.value, res.value

    def setup(self):
        self._call("setup")

    def step(self):
        self._call("step")

    def solve(self, t_final, write_freq=0):
        n = C.c_int()
        self._call("solve", C.c_double(t_final), write_freq, C.byref(n))
        return n.value

    def pcg_solve_host(self, b, x=None):
        b = np.ascontiguousarray(b, dtype=np.float64)
        if x is None:
            x = np.empty_like(b)
        it, res = C.c_int(), C.c_double()
        self._call("pcg_solve_host", b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                   C.byref(it), C.byref(res))
        return x, it.value, res.value

    def set_preconditioner(self, kind, nu_pre=2, nu_post=2, nu_coarse=8, omega=0.0):
        """kind: "jacobi" (the reference's diagonal preconditioner, default) or "mg" (opt-in geometric
        multigrid V(nu_pre, nu_post) cycle; omega <= 0: default damping)."""
        k = {"jacobi": K.PRECOND_JACOBI, "mg": K.PRECOND_MG}.get(kind, kind)
        self._call("set_preconditioner", int(k), int(nu_pre), int(nu_post), int(nu_coarse), C.c_double(omega))

    def set_mg_max_levels(self, n):
        self._call("set_mg_max_levels", int(n))

    def mg_num_levels(self):
        n = C.c_int()
        self._call("mg_num_levels", C.byref(n))
        return n.value

    def mg_apply(self, r):
        """z = M^-1 r: one multigrid V-cycle on a dense owned-cell array."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        if r.shape != self.shape(K.PRESSURE):
            raise ValueError(f"shape {r.shape} != {self.shape(K.PRESSURE)}")
        z = np.empty_like(r)
        self._call("mg_apply", r.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
        return z

    # -- output stage (SiloWriter::writeFile, src/SiloWriter.hpp:56-197)
    def output(self):
        """(quantity[z,y,x], velocity[d,z,y,x] at the cell centres, [nodes_x, nodes_y(, nodes_z)]) of the
        owned cells: what the reference hands to Silo for this block."""
        nz, ny, nx = self.shape(K.QUANTITY)
        q = np.empty((nz, ny, nx))
        vel = np.empty((self.dim, nz, ny, nx))
        nodes = [np.empty(n + 1) for n in (nx, ny, nz)[: self.dim]]
        ptr = [a.ctypes.data_as(_dp) for a in nodes] + [None] * (3 - self.dim)
        self._call("output_extract", q.ctypes.data_as(_dp), vel.ctypes.data_as(_dp), *ptr)
        return q, vel, nodes

    def write_output(self, directory, time_step):
        """siloWrite (src/SiloWriter.hpp:355-417) re-designed: extraction kernel + asynchronous copy now,
        the files (.npy per variable and block + a .json master on rank 0) at the next write / flush."""
        self._call("write_output", os.fsencode(directory), int(time_step))

    def output_flush(self):
        self._call("output_flush")

    def set_output_dir(self, directory):
        """Makes solve() write at step 0 and every write_freq steps like src/Solver.hpp:156,170-173."""
        self._call("set_output_dir", os.fsencode(directory) if directory else None)

    # -- micro-benchmarks / introspection
    def stencil_dot(self, reps=1):
        dot, ms = C.c_double(), C.c_double()
        self._call("stencil_dot", reps, C.byref(dot), C.byref(ms))
        return dot.value, ms.value

    def pcg_fixed(self, iters):
        ms, res = C.c_double(), C.c_double()
        self._call("pcg_fixed", iters, C.byref(ms), C.byref(res))
        return ms.value, res.value

    def fill_synthetic_velocity(self, variant=0, seed=1234):
        self._call("fill_synthetic_velocity", variant, C.c_uint64(seed))

    def set_tuning(self, key, value):
        self._call("set_tuning", key.encode(), int(value))

    def stats(self):
        s = K.Stats()
        self._call("get_stats", C.byref(s))
        return {n: getattr(s, n) for n, _ in K.Stats._fields_}

    def reset_stats(self):
        self._call("reset_stats")

    def residual_history(self, n=4096):
        buf = (C.c_double * n)()
        cnt = C.c_int()
        self._call("residual_history", buf, n, C.byref(cnt))
        return np.array(buf[: min(n, cnt.value)])

    def field_ptr(self, field, version=K.CURRENT):
        p, o, sy, sz = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
        self._call("field_ptr", field, version, C.byref(p), C.byref(o), C.byref(sy), C.byref(sz))
        return p.value, o.value, sy.value, sz.value


_lib = None


def load():
    """Load the CUDA library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH, "cfb_")
    return _lib
