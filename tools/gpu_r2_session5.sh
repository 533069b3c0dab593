#!/bin/bash
# Round 2, fifth GPU call (1 GPU): the persistent CG kernel — parity first, then timing against the launch-per-phase form
set -u
O=gpurun_out/r2s5
mkdir -p "$O"
timeout 600 python -m pytest tests/test_zz_d_persistent_cg.py -m gpu -q -p no:cacheprovider > "$O/pytest_persistent.log" 2>&1
echo "pytest persistent rc=$?" >> "$O/summary.txt"
timeout 600 python tools/profile_small.py 64 96 128 160 192 256 > "$O/small_grids.json" 2> "$O/small_grids.err"
# the whole suite with the persistent form chosen automatically on the small grids
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest all rc=$?" >> "$O/summary.txt"
timeout 300 python tools/run_config0.py 64 100 > "$O/config0_64cubed_100steps.json" 2> "$O/config0.err"
# 512^3 headline unchanged by the refactoring of the phase kernels?
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-probe --no-e2e > "$O/bench_n1_quick.json" 2> "$O/bench_n1_quick.err"
ls -la "$O" > "$O/listing.txt"
