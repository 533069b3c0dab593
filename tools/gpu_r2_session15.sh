#!/bin/bash
# Round 2, session 15: ncu evidence of phase A' with r staged by TMA (512^3), then CG form / z chunk by block size
set -u
O=gpurun_out/r2s15
mkdir -p "$O"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"cg_fused_kernel|stencil7_dot_tma" -s 8 -c 2 \
    -o "$O/cg512_rtma_full" python tools/profile_target.py 512 8 2 > "$O/ncu_rtma.log" 2>&1
echo "ncu rc=$?" > "$O/summary.txt"
timeout 200 python tools/sweep_forms2.py > "$O/sweep_forms2.log" 2> "$O/sweep.err"
echo "sweep rc=$?" >> "$O/summary.txt"
