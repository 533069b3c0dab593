#!/bin/bash
# Round 2, session 22 (last GPU seconds): the GPU suite with prolongation + first post-sweep of the multigrid on the TMA
# march as well (MODE 5), then the multigrid projection in the three configurations
set -u
O=gpurun_out/r2s22
mkdir -p "$O"
timeout 100 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 40 python tools/profile_mg_tma.py 512 256 > "$O/mg_tma.json" 2> "$O/mg_tma.err"
echo "mg_tma rc=$?" >> "$O/summary.txt"
