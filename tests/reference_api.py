"""Test-side binding of the reference run here: oracle/_ref/libcfref.so (prefix cfref_).

The library is the UNMODIFIED cajitafluids source tree (compiled from /root/reference by
`make -C oracle ref`) on top of the single-rank stand-ins for Kokkos / Cajita / MPI / Silo in
oracle/refshim/.  It exists only where /root/reference exists (this container); the fixtures it
produced travel instead (tests/golden/refrun_*.npz, tests/golden/make_golden_ref.py).
Test infrastructure only: nothing in the product imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cajitafluids_b200._capi import Context, Library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_LIB = os.path.join(REF_DIR, "libcfref.so")
REFERENCE_TREE = "/root/reference"

_lib = None
_dp = C.POINTER(C.c_double)


def have_tree():
    return os.path.isdir(os.path.join(REFERENCE_TREE, "src"))


def build():
    """(Re)build oracle/_ref from the reference tree; a no-op when the tree is absent."""
    if have_tree():
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def available():
    try:
        build()
    except Exception:
        pass
    return os.path.exists(REF_LIB)


def load():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{REF_LIB}: needs {REFERENCE_TREE} (make -C oracle ref)")
        _lib = Library(REF_LIB, "cfref_")
        d = _lib.dll
        d.cfref_set_cg_arithmetic.argtypes = [C.c_int]
        d.cfref_matrix.argtypes = [C.c_void_p, _dp, _dp]
        d.cfref_silo_last.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, C.POINTER(C.c_int), _dp, _dp, _dp,
                                      _dp, _dp]
        d.cfref_num_threads.restype = C.c_int
    return _lib


class Reference(Context):
    """The reference's Solver<2> behind the cfb-shaped entry points.

    exact=True: the Cajita stand-in's CG uses the oracle's arithmetic (fused multiply-adds,
    double-double dot products), so runs are bit-comparable with the oracle and the CUDA path;
    exact=False: plain double arithmetic, the reference's own cost and rounding."""

    def __init__(self, cfg, exact=True):
        lib = load()
        lib.dll.cfref_set_cg_arithmetic(1 if exact else 0)
        super().__init__(lib, cfg)

    def matrix(self):
        nz, ny, nx = self.shape(0)
        A = np.empty((ny, nx, 5))
        Mi = np.empty((ny, nx))
        self.lib.check(self.lib.dll.cfref_matrix(self.h, A.ctypes.data_as(_dp), Mi.ctypes.data_as(_dp)), self.h)
        return A, Mi

    def silo_last(self):
        """(writes, cycle, time, quantity[y,x], u_cc[y,x], v_cc[y,x], xnodes, ynodes) of the last Silo write."""
        w, cyc, t = C.c_int(), C.c_int(), C.c_double()
        dims = (C.c_int * 2)()
        d = self.lib.dll
        d.cfref_silo_last(C.byref(w), C.byref(cyc), C.byref(t), dims, None, None, None, None, None)
        nx, ny = dims[0], dims[1]
        q, u, v = (np.empty((ny, nx)) for _ in range(3))
        xn, yn = np.empty(nx + 1), np.empty(ny + 1)
        d.cfref_silo_last(None, None, None, None, q.ctypes.data_as(_dp), u.ctypes.data_as(_dp),
                          v.ctypes.data_as(_dp), xn.ctypes.data_as(_dp), yn.ctypes.data_as(_dp))
        return w.value, cyc.value, t.value, q, u, v, xn, yn

    @staticmethod
    def num_threads():
        return load().dll.cfref_num_threads()
