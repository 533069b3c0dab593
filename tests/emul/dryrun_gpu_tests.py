"""tests/emul/dryrun_gpu_tests.py — the `-m gpu` tests themselves, run against the host-emulated library:

    python tests/emul/run_with_emul.py tests/emul/dryrun_gpu_tests.py [pytest arguments]
    CFB_EMUL_TMA=1 python tests/emul/run_with_emul.py tests/emul/dryrun_gpu_tests.py -k flat

Catches mistakes in the tests' own logic (an incompatible right-hand side, a cycle CG cannot converge with)
before they cost a GPU run.  Not part of the suite (≈ 5 min); the C++ driver test is left out because it links
the CUDA library by path."""
import os
import sys

import pytest

os.environ.setdefault("CFB_FULL_N", "48")  # the full-size tests at a size the emulation finishes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.exit(pytest.main([os.path.join(ROOT, "tests"), "-q", "-m", "gpu", "-p", "no:cacheprovider", "--deselect",
                      "tests/test_zz_cpp_layer.py::test_cpp_shims_and_driver_on_the_gpu"] + sys.argv[1:]))
