#!/bin/bash
# Round 2, session 14: phase A' of the 64-byte form with r staged by TMA ("stencil_rtma") against the LDG form (the key was removed with that form after this measurement),
# every stencil tiling, at 512^3 / 384^3 / 256^3 (one GPU, CUDA-event times per phase)
set -u
O=gpurun_out/r2s14
mkdir -p "$O"
timeout 60 python -m pytest tests/test_zz_b_cg_variant2.py -x -q -m gpu > "$O/pytest_variant2.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 300 python tools/sweep_variant2.py 512 384 256 > "$O/sweep_rtma.log" 2> "$O/sweep.err"
echo "sweep rc=$?" >> "$O/summary.txt"
