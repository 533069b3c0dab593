"""tests/emul/run_with_emul.py — run one of the repo's GPU scripts (tools/*.py, bench.py) against the
host-emulated library, to catch Python-level mistakes before spending GPU minutes on them:

    python tests/emul/run_with_emul.py tools/profile_mg.py 32 solve

`cajitafluids_b200.Solver` becomes a Context on libcfb_emul.so and the torch.cuda calls the scripts make are
stubbed.  The numbers printed mean nothing."""
import ctypes
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import torch  # noqa: E402

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self

import build_emul  # noqa: E402
import cajitafluids_b200  # noqa: E402
from cajitafluids_b200._capi import Context, Library  # noqa: E402

_lib = Library(build_emul.build(tma=os.environ.get("CFB_EMUL_TMA") == "1"), "cfb_", mode=ctypes.RTLD_LOCAL)


class EmulSolver(Context):
    def __init__(self, cfg):
        super().__init__(_lib, cfg)


cajitafluids_b200.Solver = EmulSolver
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(os.path.join(ROOT, script) if not os.path.isabs(script) else script, run_name="__main__")
