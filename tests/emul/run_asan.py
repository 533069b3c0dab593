"""tests/emul/run_asan.py — the emulated library under AddressSanitizer (out-of-bounds indexing in the
streaming kernels, the pack / unpack kernels and the host orchestration shows up as a heap-buffer-overflow,
because every device array is its own malloc block here).  Run from the repo root:

    python tests/emul/run_asan.py

It rebuilds the emulated sources with -fsanitize=address into _build/libcfb_emul_asan.so and re-executes
itself with libasan preloaded.  (The peer-memory path is left out: its fibers would need ASan annotations.)
"""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]


def build():
    import build_emul
    build_emul.build(force=True)
    out = build_emul.OUT
    cpp = [os.path.join(out, f.replace(".cu", "_emul.cpp")) for f in build_emul.SOURCES] + \
          [os.path.join(out, "emul_glue.cpp"), os.path.join(out, "nccl_emul.cpp")]
    lib = os.path.join(out, "libcfb_emul_asan.so")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer",
           "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-pthread", "-Wl,-Bsymbolic",
           "-Ddlopen=cfb_emul_dlopen", "-Ddlsym=cfb_emul_dlsym", "-Ddlerror=cfb_emul_dlerror", "-I", out, "-o", lib] + cpp
    subprocess.check_call(cmd)
    return lib


def workload(lib_path):
    import numpy as np
    from cajitafluids_b200 import config as K
    from cajitafluids_b200._capi import Context, Library
    from helpers import make_cfg
    from multirank import run_ranks

    lib = Library(lib_path, "cfb_", mode=ctypes.RTLD_LOCAL)

    def run(ctx, steps):
        ctx.setup()
        for _ in range(steps):
            ctx.step()

    for dim, cells, prec in ((2, (37, 23), "jacobi"), (3, (20, 14, 11), "jacobi"), (2, (48, 40), "mg"),
                             (3, (24, 20, 16), "mg"), (3, 18, "mg"), (3, (6, 6, 6), "mg")):
        box = 1.0 if isinstance(cells, int) else tuple(c / cells[0] for c in cells)
        g = Context(lib, make_cfg(dim, cells, box=box))
        g.set_preconditioner(prec)
        g.set_tuning("advect_tile", 1 if prec == "jacobi" else 0)  # entity tiles: ragged edges of the tile grid
        run(g, 2)
        g.output()
        if prec == "mg":
            for nu in ((1, 0, 1), (1, 1, 2), (3, 2, 4)):
                g.set_preconditioner("mg", *nu)
                g.mg_apply(np.ones(g.shape(K.PRESSURE)))
        g.close()
        print("ok", dim, cells, prec, flush=True)

    def body(ctx, rank):
        ctx.set_preconditioner("mg")
        ctx.setup()
        ctx.step()
        ctx.output()
        ctx.set_preconditioner("jacobi")
        ctx.step()
        return 1

    print("8 ranks:", run_ranks(lib, make_cfg(3, (32, 32, 32)), 8, body))

    def body2(ctx, rank):
        ctx.setup()
        ctx.step()
        ctx.gather()
        return 1

    print("6 ranks (1x3x2, uneven):", run_ranks(lib, make_cfg(3, (24, 20, 18), box=(1, 20 / 24, 18 / 24)), 6, body2, (1, 3, 2)))


if __name__ == "__main__":
    if os.environ.get("CFB_ASAN_CHILD") == "1":
        workload(sys.argv[1])
        print("asan run clean")
    else:
        lib = build()
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, CFB_ASAN_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__), lib], env=env))
