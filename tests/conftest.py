import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _cpu_only_run(config):
    """the CPU suite (`-m "not gpu"`), not a GPU run and not a worker of an already parallel run."""
    mark = getattr(config.option, "markexpr", "") or ""
    return "not gpu" in mark and not hasattr(config, "workerinput")


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite is dominated by the host-emulated kernels (tests/emul: one fiber per CUDA thread); its tests are
    independent, so it is spread over a few worker processes when pytest-xdist is there and the caller did not
    choose (`-n ...`, `-p no:xdist`).  Never for GPU runs: those share one device.  The emulated libraries are built
    once, here, before any worker starts (the workers would otherwise race on tests/emul/_build)."""
    if not _cpu_only_run(config):
        return None
    try:
        import build_emul
        build_emul.build()
        build_emul.build(tma=True)
    except Exception:  # noqa: BLE001  (the tests that need them report the build error themselves)
        pass
    if os.environ.get("CFB_TEST_WORKERS", "") == "0" or not config.pluginmanager.hasplugin("xdist"):
        return None
    if getattr(config.option, "numprocesses", None) in (None, 0) and getattr(config.option, "dist", "no") == "no":
        n = int(os.environ.get("CFB_TEST_WORKERS", "0") or 0) or max(1, min(4, (os.cpu_count() or 2) // 2))
        if n > 1:
            # the checker is an OpenMP code and the multi-rank emulation runs one thread per rank: share the cores
            # among the workers instead of letting every worker spin on all of them
            os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 2) // n)))
            os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
            config.option.numprocesses = n
            config.option.tx = ["popen"] * n
            config.option.dist = "load"
    return None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
