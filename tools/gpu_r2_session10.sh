#!/bin/bash
# Round 2, closing 1-GPU call: suite + smoke + bench with the committed code
set -u
O=gpurun_out/r2s10
mkdir -p "$O"
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
timeout 900 python bench.py > "$O/bench_n1.json" 2> "$O/bench_n1.err"
echo "bench rc=$?" >> "$O/summary.txt"
ls -la "$O" > "$O/listing.txt"
