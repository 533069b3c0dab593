#!/bin/bash
# Round 2, session 19: validation of the session's code (phase A' with r by TMA, sixteen-warp phase B, 2-D march) —
# the GPU suite, the bench as the driver runs it, smoke, ncu evidence of the kernels as shipped
set -u
O=gpurun_out/r2s19
mkdir -p "$O"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 600 python bench.py > "$O/bench_n1.json" 2> "$O/bench_n1.err"
echo "bench rc=$?" >> "$O/summary.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$O/launches_bench512.csv" \
    python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline --no-e2e --no-timestep --no-probe > "$O/ncu_bench.log" 2>&1
echo "ncu list rc=$?" >> "$O/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cg_fused_kernel|stencil7_dot_tma" -s 8 -c 2 \
    -o "$O/cg512_final_full" python tools/profile_target.py 512 8 2 > "$O/ncu_final.log" 2>&1
echo "ncu full rc=$?" >> "$O/summary.txt"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:"cg_fused_kernel|stencil7_dot_tma" -s 20 -c 40 --csv --log-file "$O/launches_cg256_variant2.csv" \
    python tools/profile_target.py 256 40 2 > "$O/ncu_256.log" 2>&1
echo "ncu 256 rc=$?" >> "$O/summary.txt"
timeout 200 python tools/profile_small_2d.py > "$O/small_2d.json" 2> "$O/small_2d.err"
echo "small 2d rc=$?" >> "$O/summary.txt"
