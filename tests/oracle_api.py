"""Test-side binding of the CPU oracle (oracle/libcfo_oracle.so, prefix cfo_).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cajitafluids_b200 import config as K
from cajitafluids_b200._capi import Context, Library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libcfo_oracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        try:
            build()  # make is a no-op when the .so is newer than cfo_oracle.cpp and include/cfb.h
        except Exception:
            if not os.path.exists(ORACLE_LIB):
                raise
        _lib = Library(ORACLE_LIB, "cfo_")
        d = _lib.dll
        d.cfo_matrix_ptr.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_double)),
                                     C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int)]
        d.cfo_spline_weights.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        d.cfo_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]
        d.cfo_rk3.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        d.cfo_coordinates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        d.cfo_field_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_double)),
                                    C.POINTER(C.c_int)]
        d.cfo_set_callbacks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        d.cfo_num_threads.restype = C.c_int
        d.cfo_set_num_threads.argtypes = [C.c_int]
        d.cfo_set_num_threads.restype = C.c_int
        d.cfo_set_fixed_iters.argtypes = [C.c_void_p, C.c_int]
        d.cfo_set_cg_algorithm.argtypes = [C.c_void_p, C.c_int]
        d.cfo_set_accumulation.argtypes = [C.c_void_p, C.c_int]
        d.cfo_set_mg_max_levels.argtypes = [C.c_void_p, C.c_int]
        d.cfo_mg_num_levels.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    return _lib


def set_num_threads(n=0):
    """OpenMP threads of the oracle from here on (0: every core the process may run on, whatever
    OMP_NUM_THREADS a launcher exported); returns the count in use."""
    return load().dll.cfo_set_num_threads(int(n))


GATHER_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int)
ALLREDUCE_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int)


class Oracle(Context):
    def __init__(self, cfg):
        super().__init__(load(), cfg)
        self._cbs = None

    def set_accumulation(self, exact):
        """exact=True (default): double-double sums; False: plain double like the reference."""
        self.lib.dll.cfo_set_accumulation(self.h, 1 if exact else 0)

    def set_cg_algorithm(self, algorithm):
        """0: Cajita's ReferenceConjugateGradient loop (default); 1: the single-reduction form (the product's
        cg_variant 3), the same iterates in exact arithmetic."""
        self.lib.check(self.lib.dll.cfo_set_cg_algorithm(self.h, int(algorithm)), self.h)

    def set_fixed_iters(self, iters):
        """the following solves run exactly `iters` iterations (0: back to the stopping test)."""
        self.lib.dll.cfo_set_fixed_iters(self.h, int(iters))

    def num_threads(self):
        return self.lib.dll.cfo_num_threads()

    def matrix(self):
        """(A, Minv) as numpy views in the ghosted cell layout: A[z,y,x,c], Minv[z,y,x]."""
        A, M, n = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.c_int()
        self.lib.dll.cfo_matrix_ptr(self.h, C.byref(A), C.byref(M), C.byref(n))
        shp = self.shape(K.PRESSURE, K.GHOSTED)
        a = np.ctypeslib.as_array(A, shape=shp + (n.value,))
        m = np.ctypeslib.as_array(M, shape=shp)
        return a, m

    def view(self, field, version=K.CURRENT):
        """Zero-copy numpy view [z,y,x] of the ghosted array (reference local indices)."""
        p, e = C.POINTER(C.c_double)(), (C.c_int * 3)()
        self.lib.check(self.lib.dll.cfo_field_ptr(self.h, field, version, C.byref(p), e), self.h)
        return np.ctypeslib.as_array(p, shape=(e[2], e[1], e[0]))

    def interpolate(self, field, order, loc):
        l = (C.c_double * 3)(*list(loc) + [0.0] * (3 - len(loc)))
        v = C.c_double()
        self.lib.check(self.lib.dll.cfo_interpolate(self.h, field, order, l, C.byref(v)), self.h)
        return v.value

    def rk3(self, x0):
        a = (C.c_double * 3)(*list(x0) + [0.0] * (3 - len(x0)))
        t = (C.c_double * 3)()
        self.lib.dll.cfo_rk3(self.h, a, t)
        return tuple(t)[: self.dim]

    def coordinates(self, field, idx):
        i = (C.c_int * 3)(*list(idx) + [0] * (3 - len(idx)))
        x = (C.c_double * 3)()
        self.lib.dll.cfo_coordinates(self.h, field, i, x)
        return tuple(x)[: self.dim]

    def set_callbacks(self, gather, allreduce):
        g, a = GATHER_CB(gather), ALLREDUCE_CB(allreduce)
        self._cbs = (g, a)
        self.lib.dll.cfo_set_callbacks(self.h, C.cast(g, C.c_void_p), C.cast(a, C.c_void_p), None)


def spline_weights(order, xl):
    lib = load()
    n = order + 1
    s, w = (C.c_int * n)(), (C.c_double * n)()
    lib.dll.cfo_spline_weights(order, xl, s, w)
    return list(s), list(w)
