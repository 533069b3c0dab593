#!/bin/bash
# One gpurun call that (re)establishes the measured state of the repo on a B200:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'
# Everything lands in gpurun_out/session/.  Order: cheapest and most important first.
set -u
O=gpurun_out/session
mkdir -p "$O"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > "$O/gpu.csv" 2>&1

# 1. the GPU tests, file by file, so that one failing file does not hide the others
for f in tests/test_capi.py tests/test_golden.py tests/test_golden_refrun.py tests/test_gpu_parity.py \
         tests/test_zz_a_output_stage.py tests/test_zz_b_cg_variant2.py tests/test_zz_multigrid.py tests/test_zz_cpp_layer.py tests/test_zzy_full_size.py tests/test_zzz_late_options.py; do
    timeout 900 python -m pytest "$f" -m gpu -q -x > "$O/pytest_$(basename "$f" .py).log" 2>&1
    echo "$f rc=$?" >> "$O/pytest_summary.txt"
done

# 2. multigrid / output stage numbers (never measured before round 2)
for n in 128 256 512; do
    timeout 600 python tools/profile_mg.py $n solve >> "$O/mg_solve.log" 2>&1
done
timeout 300 python tools/profile_mg.py 256 output >> "$O/output_stage.log" 2>&1
timeout 300 python tools/profile_mg.py 512 output >> "$O/output_stage.log" 2>&1

# 3. the headline bench
timeout 900 python bench.py --steps 5 --warmup 3 > "$O/bench_n1.json" 2> "$O/bench_n1.err"

# 3b. the 64-byte iteration (cg_variant 2), never measured before round 2
timeout 600 python bench.py --steps 5 --warmup 3 --cg-variant 2 --no-cpu-baseline --no-e2e --no-timestep \
    > "$O/bench_n1_variant2.json" 2> "$O/bench_n1_variant2.err"
for zc in 16 32 64; do
    timeout 300 python bench.py --steps 3 --warmup 3 --cg-variant 2 --no-cpu-baseline --no-e2e --no-timestep \
        --tune stencil_zc=$zc >> "$O/bench_n1_variant2_sweep.json" 2>> "$O/bench_n1_variant2.err"
done

# 3c. two-dimensional runs with / without the ghost-plane loads (flat_2d), never measured before round 2
timeout 300 python tools/profile_2d.py 8192 50 > "$O/flat2d.json" 2> "$O/flat2d.err"

# 3d. advection kernel, rows vs entity tiles (advect_tile), never measured before round 2
timeout 300 python tools/profile_advect.py 512 3 > "$O/advect_tile.json" 2> "$O/advect_tile.err"

# 4. launch lists (ncu, serialised; shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/launches_mg256.csv" \
    python tools/profile_mg.py 256 cycle > "$O/ncu_mg.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mg_smooth_kernel|mg_restrict|mg_prolong|output_extract" \
    -c 8 -o "$O/mg_full" python tools/profile_mg.py 256 cycle >> "$O/ncu_mg.log" 2>&1
# 5. the 64-byte iteration under ncu: launch list, then the full set for its two kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file "$O/launches_variant2_512.csv" \
    python tools/profile_target.py 512 20 2 > "$O/ncu_variant2.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil7_dot_tma|cg_fused_kernel" -s 4 -c 6 \
    -o "$O/variant2_full" python tools/profile_target.py 512 6 2 >> "$O/ncu_variant2.log" 2>&1
ls -la "$O" > "$O/listing.txt"
