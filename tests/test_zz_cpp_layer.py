"""The C++ drop-in layer end to end on the CPU: examples/advection.cpp (the reference's driver rewritten against
include/cajitafluids_b200/CajitaFluids.hpp) linked against the host-emulated library (tests/emul), run with the
reference's own command line, and compared with

  * the reference's own CLI (oracle/_ref/advection_ref = the unmodified examples/advection.cpp on the stand-ins):
    the solver's progress lines — iteration counts and residuals — are identical, character for character;
  * the oracle: the dumped quantity and the files the SiloWriter mirror wrote, bit for bit.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emul"))

import build_emul  # noqa: E402
from cajitafluids_b200 import config as K  # noqa: E402
from helpers import make_cfg  # noqa: E402
from oracle_api import Oracle  # noqa: E402

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "advection_ref")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    lib = build_emul.build()
    exe = str(tmp_path_factory.mktemp("drv") / "advection_emul")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "advection.cpp"), "-o", exe, lib,
                           "-Wl,-rpath," + os.path.dirname(lib), "-ldl", "-lpthread"])
    return exe


def solver_lines(text):
    return [ln for ln in text.splitlines() if ln.startswith(("Cajita CG", "Step "))]


@pytest.mark.parametrize("args,dim,n,steps", [(["-n", "32", "-t", "0.02"], 2, 32, 4),
                                              (["-n", "24", "-t", "0.011", "-g", "9.8", "-w", "0.06", "-h", "0.2",
                                                "-u", "0.8", "-v", "0.3", "-q", "2.5"], 2, 24, 3)])
def test_cpp_driver_on_emulated_kernels(driver, tmp_path, args, dim, n, steps):
    dump = str(tmp_path / "q.bin")
    p = subprocess.run([driver] + args + ["-D", str(dim), "-o", dump], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    mine = solver_lines(p.stdout)
    assert len([ln for ln in mine if ln.startswith("Cajita CG")]) == steps + 1
    if os.path.exists(REF_CLI):
        r = subprocess.run([REF_CLI] + args + ["-m", "Reference", "-p", "serial"], cwd=tmp_path, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0
        assert solver_lines(r.stdout) == mine
    def opt(flag, default):
        return float(args[args.index(flag) + 1]) if flag in args else default

    cfg = make_cfg(dim, n, gravity=opt("-g", 0.0))
    cfg.inflow_size[0], cfg.inflow_size[1] = opt("-w", 0.02), opt("-h", 0.1)
    cfg.inflow_velocity[0], cfg.inflow_velocity[1] = opt("-u", 1.0), opt("-v", 0.0)
    cfg.inflow_quantity = opt("-q", 3.0)
    o = Oracle(cfg)
    assert o.solve(float(args[args.index("-t") + 1]), 0) == steps
    q = np.fromfile(dump).reshape(o.shape(K.QUANTITY))
    assert np.array_equal(q, o.get(K.QUANTITY))
    # Solver::solve wrote through the SiloWriter mirror into data/ like the reference does
    wf = 20  # examples/advection.cpp:175: no option for it in the reference's CLI either
    written = sorted(int(re.search(r"(\d{5})\.json", f).group(1)) for f in os.listdir(tmp_path / "data") if f.endswith(".json"))
    assert written == [t for t in range(steps) if t % wf == 0]
    last = written[-1]
    oq, ov, _ = o.output()
    if last == steps - 1:
        fq = np.load(tmp_path / "data" / "raw" / ("CajitaFluidsOutput%05d%05d.quantity.npy" % (0, last)))
        fv = np.load(tmp_path / "data" / "raw" / ("CajitaFluidsOutput%05d%05d.velocity.npy" % (0, last)))
        assert np.array_equal(fq, oq[0]) and np.array_equal(fv, ov[:, 0])


def test_cpp_driver_multigrid_flag_and_errors(driver, tmp_path):
    p = subprocess.run([driver, "-n", "32", "-t", "0.011", "-c", "MG", "--output-dir", "none"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    its = [int(m) for m in re.findall(r"Finished in (\d+) iterations", p.stdout)]
    assert len(its) == 4 and max(its) <= 12
    assert not os.path.exists(tmp_path / "data")
    p = subprocess.run([driver, "-n", "32", "-t", "0.01", "-m", "PCG"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=60)
    assert p.returncode != 0 and "Reference" in (p.stdout + p.stderr)


def build_shim_tests(lib, out_dir):
    exe = os.path.join(str(out_dir), "tst_shims")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "tst_shims.cpp"), "-o", exe, lib,
                           "-Wl,-rpath," + os.path.dirname(lib), "-ldl", "-lpthread"])
    return exe


def test_reference_unit_tests_re_expressed_on_the_cpp_shims(tmp_path):
    """tests/cpp/tst_shims.cpp: the reference's tstMesh / tstProblemManager / (the intent of) tstBoundaryConditions
    and the three plug-in seams of INTEGRATION.md, on the host-emulated library."""
    exe = build_shim_tests(build_emul.build(), tmp_path)
    p = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "all shim tests passed" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_cpp_shims_and_driver_on_the_gpu(tmp_path):
    """The same C++ programs against the CUDA library."""
    from cajitafluids_b200._capi import LIB_PATH
    exe = build_shim_tests(LIB_PATH, tmp_path)
    p = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "all shim tests passed" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
    drv = os.path.join(ROOT, "examples", "advection_b200")
    if not os.path.exists(drv):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")])
    q = str(tmp_path / "q.bin")
    p = subprocess.run([drv, "-n", "32", "-t", "0.02", "-D", "2", "-o", q], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    its = [int(m) for m in re.findall(r"Finished in (\d+) iterations", p.stdout)]
    assert its == [105, 117, 119, 119, 118]  # the reference's own CLI prints the same (checked on the CPU above)
    o = Oracle(make_cfg(2, 32))
    o.solve(0.02, 0)
    assert np.array_equal(np.fromfile(q).reshape(o.shape(K.QUANTITY)), o.get(K.QUANTITY))
