#!/bin/bash
# Round 2, session 18: two-dimensional grids — units marching along y (fused_unit_flat): GPU parity, then run lengths
set -u
O=gpurun_out/r2s18
mkdir -p "$O"
timeout 200 python -m pytest tests/test_zzz_late_options.py -x -q -m gpu > "$O/pytest_late_options.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 150 python tools/sweep_2d_tilings.py > "$O/sweep_2d_march.log" 2> "$O/sweep_2d.err"
echo "sweep 2d rc=$?" >> "$O/summary.txt"
