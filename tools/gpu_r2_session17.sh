#!/bin/bash
# Round 2, session 17: two-dimensional grids — phase B tilings; the library's choices at 512^3 / 384^3 after fused_nt
set -u
O=gpurun_out/r2s17
mkdir -p "$O"
timeout 150 python tools/sweep_2d_tilings.py > "$O/sweep_2d_tilings.log" 2> "$O/sweep_2d.err"
echo "sweep 2d rc=$?" > "$O/summary.txt"
timeout 100 python tools/profile_forms.py 512 384 > "$O/forms.log" 2> "$O/forms.err"
echo "forms rc=$?" >> "$O/summary.txt"
