"""cajitafluids_b200 — B200-native (sm_100a) pressure projection + semi-Lagrangian advection.

The product is the CUDA shared library `lib/libcfb_b200.so` (sources in `csrc/`, C ABI in
`include/cfb.h`).  This package is only the thin host-side mirror of the reference's
Solver / ProblemManager / VelocityCorrector interface used by tests and bench.py.
"""
from . import config
from .config import Config, default_config
from ._capi import CfbError, Context, Library, load
from .solver import Solver, create_solver

__all__ = ["config", "Config", "default_config", "CfbError", "Context", "Library", "load",
           "Solver", "create_solver"]
