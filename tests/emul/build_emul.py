"""tests/emul/build_emul.py — builds tests/emul/_build/libcfb_emul.so: the product's plain streaming
kernels and host orchestration compiled as ordinary C++ (see cuda_runtime.h in this directory).

The .cu sources are taken from cajitafluids_b200/csrc as they are; the only rewrite is the launch
syntax  kernel<<<grid, block, smem, stream>>>( args )  ->  cfb_emul::launch( grid, block, [=]{ kernel( args ); } )
(arguments captured by value, as a real launch copies them: recorded launches can be replayed as a graph).
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cajitafluids_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcfb_emul.so")
LIB_TMA = os.path.join(OUT, "libcfb_emul_tma.so")  # the TMA kernels themselves instead of plain-loop stand-ins
TMA_SOURCES = ["kernels_stencil.cu", "kernels_fused.cu"]
SOURCES = ["cfb_api.cu", "kernels_fields.cu", "kernels_cg.cu", "kernels_cg1.cu", "mg.cu", "output.cu", "halo.cu"]
HEADERS = ["cfb_internal.h", "device_geo.cuh", "device_peer.cuh", "device_cg1.cuh"]
# kernels whose threads meet at __syncthreads() for real: one fiber per CUDA thread
# (a name with its template arguments selects that instantiation only: phase A meets at barriers only when it
# runs the mailbox exchange itself)
COOP_KERNELS = {"cg_xchg_kernel", "cg_persistent_kernel", "cg_face_kernel", "stencil7_dot_tma", "cg_fused_kernel", "cg_rupdate_kernel<true>", "mg_coarse_cycle_kernel",
                "mg_xchg_kernel"}
STANDINS = ["cuda_runtime.h", "cuda.h", "nccl.h", "device_reduce.cuh", "device_tma.cuh", "emul_glue.cpp",
            "nccl_emul.cpp"]


def _match(s, i, open_c, close_c):
    """index just after the bracket that closes the one at s[i]."""
    depth = 0
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(src):
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        # kernel expression: identifier, optionally followed by a template argument list
        k = i
        if src[k - 1] == ">":
            depth, k = 0, k - 1
            while True:
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", src[:k])
        start = m.start()
        kernel = src[start:i]
        j = src.find(">>>", i)
        cfg = _split_top(src[i + 3:j])
        a0 = src.index("(", j)
        a1 = _match(src, a0, "(", ")")
        args = src[a0 + 1:a1 - 1]
        fn = "launch_coop" if kernel.strip().split("<")[0] in COOP_KERNELS or kernel.strip() in COOP_KERNELS else "launch"
        out += src[pos:start] + (f"cfb_emul::{fn}( dim3( {cfg[0]} ), dim3( {cfg[1]} ), [=]() {{ {kernel}( {args} ); }} )")
        pos = a1


def build(force=False, tma=False):
    """tma=False: libcfb_emul.so, the TMA kernels replaced by plain-loop stand-ins (fast);
    tma=True: libcfb_emul_tma.so, kernels_stencil.cu / kernels_fused.cu themselves, one fiber per CUDA thread."""
    os.makedirs(OUT, exist_ok=True)
    lib = LIB_TMA if tma else LIB
    sources = SOURCES + (TMA_SOURCES if tma else [])
    deps = [os.path.join(CSRC, f) for f in sources + HEADERS] + [os.path.join(HERE, f) for f in STANDINS] + \
           [os.path.join(ROOT, "include", "cfb.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    cpp = []
    for f in sources:
        src = rewrite_launches(open(os.path.join(CSRC, f)).read())
        # the one piece of inline PTX outside the TMA kernels: a volatile 64-bit load
        src = src.replace('asm volatile( "ld.volatile.global.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );',
                          "v = *reinterpret_cast<const volatile unsigned long long*>( p );")
        # dynamic shared memory: a thread-local buffer of the rank thread
        src = src.replace("extern __shared__ unsigned char smem_raw[];",
                          "unsigned char* const smem_raw = cfb_emul::dyn_smem();")
        dst = os.path.join(OUT, f.replace(".cu", "_emul.cpp"))
        open(dst, "w").write(src)
        cpp.append(dst)
    for f in HEADERS:
        txt = open(os.path.join(CSRC, f)).read().replace('#include "../../include/cfb.h"',
                                                         f'#include "{os.path.join(ROOT, "include", "cfb.h")}"')
        open(os.path.join(OUT, f), "w").write(txt)
    for f in STANDINS:
        open(os.path.join(OUT, f), "w").write(open(os.path.join(HERE, f)).read())
    cpp.append(os.path.join(OUT, "emul_glue.cpp"))
    cpp.append(os.path.join(OUT, "nccl_emul.cpp"))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # -ffp-contract=off: like nvcc -fmad=false, only the explicit fma() calls fuse
    cmd = [cxx, "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-pthread",
           # the real library may be loaded in the same process (RTLD_GLOBAL): bind our own symbols to ourselves
           "-Wl,-Bsymbolic",
           "-Ddlopen=cfb_emul_dlopen", "-Ddlsym=cfb_emul_dlsym", "-Ddlerror=cfb_emul_dlerror"] + \
          (["-DCFB_EMUL_REAL_TMA"] if tma else []) + ["-I", OUT, "-o", lib] + cpp
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("emulation build failed")
    return lib


if __name__ == "__main__":
    print(build(force=True))
    print(build(force=True, tma=True))
