#!/bin/bash
# Round 2, final 1-GPU call: the suite, the bench and the ncu evidence with the library's own choice of CG form
set -u
O=gpurun_out/r2s9
mkdir -p "$O"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
timeout 900 python bench.py > "$O/bench_n1.json" 2> "$O/bench_n1.err"
echo "bench rc=$?" >> "$O/summary.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$O/launches_bench512.csv" \
    python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline --no-e2e --no-timestep --no-probe > "$O/ncu_bench.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_fused_kernel|stencil7_dot_tma" -s 6 -c 4 \
    -o "$O/cg512_variant2_full" python tools/profile_target.py 512 8 2 > "$O/ncu_variant2.log" 2>&1
timeout 300 python tools/profile_small.py 128 > "$O/small_128.json" 2> "$O/small_128.err"
ls -la "$O" > "$O/listing.txt"
