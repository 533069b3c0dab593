"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cfb.h declares, agrees with the Python config mirror, and FAILS LOUDLY without a GPU
(no compute call is made here)."""
import ctypes as C
import os
import re

import pytest

from cajitafluids_b200 import CfbError, Context, config as K, create_solver, default_config, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cfb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = load()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib.dll, n)]
    assert not missing, missing
    assert lib.dll.cfb_abi_version() == 1


def test_config_struct_matches_c_layout_and_defaults():
    lib = load()
    for dim in (2, 3):
        c_cfg = K.Config()
        assert lib.fn["default_config"](C.byref(c_cfg), dim) == K.OK
        assert c_cfg.struct_size == C.sizeof(K.Config)
        py = default_config(dim, 128)
        for name, _ in K.Config._fields_:
            a, b = getattr(c_cfg, name), getattr(py, name)
            if hasattr(a, "__len__"):
                assert list(a) == list(b), name
            else:
                assert a == b, name


def test_partition_helper_is_cajitas_block_split():
    lib = load()
    o, f = C.c_int(), C.c_int()
    got = []
    for b in range(3):
        assert lib.fn["partition"](100, 3, b, C.byref(o), C.byref(f)) == K.OK
        got.append((o.value, f.value))
    assert got == [(34, 0), (33, 34), (33, 67)]
    assert lib.fn["partition"](100, 3, 3, C.byref(o), C.byref(f)) == K.ERR_INVALID


def test_invalid_arguments_are_rejected_before_touching_the_device():
    lib = load()
    cfg = default_config(3, 32)
    cfg.struct_size = 12
    with pytest.raises(CfbError) as e:
        Context(lib, cfg)
    assert e.value.code == K.ERR_INVALID
    cfg = default_config(3, 32)
    cfg.global_bounding_box[4] = 2.0  # src/Mesh.hpp:56-64 -> logic_error
    with pytest.raises(CfbError) as e:
        Context(lib, cfg)
    assert e.value.code == K.ERR_MESH_EXTENT
    assert "Extent not evenly divisible" in str(e.value)


def test_backend_string_dispatch_like_create_solver():
    # src/Solver.hpp:345-349: unknown device string -> runtime_error("invalid backend")
    with pytest.raises(RuntimeError, match="invalid backend"):
        create_solver("openmp", default_config(2, 32))
    with pytest.raises(RuntimeError):
        create_solver("b200", default_config(2, 32), matrix_solver="PCG")


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(CfbError) as e:
        create_solver("b200", default_config(3, 32))
    assert e.value.code == K.ERR_NO_DEVICE


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under cajitafluids_b200/ or include/ may name it."""
    bad = []
    for base in ("cajitafluids_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or fn == "Makefile":
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if "libcfo" in txt or "cfo_" in txt or "oracle_api" in txt:
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
