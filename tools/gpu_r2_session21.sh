#!/bin/bash
# Round 2, session 21 (last GPU call): the GPU suite with the multigrid's fine-level sweeps on the TMA march, then
# the multigrid projection with and without them, smoke
set -u
O=gpurun_out/r2s21
mkdir -p "$O"
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 60 python tools/profile_mg_tma.py 512 256 128 > "$O/mg_tma.json" 2> "$O/mg_tma.err"
echo "mg_tma rc=$?" >> "$O/summary.txt"
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
