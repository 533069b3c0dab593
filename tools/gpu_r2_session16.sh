#!/bin/bash
# Round 2, session 16: the library's new choices (form, chunk) by size; phase B with sixteen warps / top-down walk
set -u
O=gpurun_out/r2s16
mkdir -p "$O"
timeout 90 python -m pytest tests/test_zz_b_cg_variant2.py -x -q -m gpu > "$O/pytest_variant2.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 120 python tools/sweep_phase_b.py 512 384 > "$O/sweep_phase_b.log" 2> "$O/sweep_b.err"
echo "sweep b rc=$?" >> "$O/summary.txt"
timeout 120 python tools/profile_forms.py > "$O/forms.log" 2> "$O/forms.err"
echo "forms rc=$?" >> "$O/summary.txt"
