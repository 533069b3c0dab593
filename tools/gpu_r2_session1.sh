#!/bin/bash
# Round 2, first GPU call (1 GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_r2_session1.sh'
set -u
O=gpurun_out/r2s1
mkdir -p "$O"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > "$O/gpu.csv" 2>&1
# 1. the whole GPU suite, no -x: one failure must not hide the rest
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
# 2. the headline bench as the driver runs it (defaults)
timeout 900 python bench.py > "$O/bench_n1.json" 2> "$O/bench_n1.err"
echo "bench rc=$?" >> "$O/summary.txt"
# 3. advection kernel: timing rows vs tiles, then ncu
timeout 300 python tools/profile_advect.py 512 3 > "$O/advect_512.json" 2> "$O/advect_512.err"
timeout 300 python tools/profile_advect.py 128 5 > "$O/advect_128.json" 2>> "$O/advect_512.err"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o "$O/advect512_full" \
    python tools/profile_advect_target.py 512 0 > "$O/ncu_advect.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o "$O/advect512_tile_full" \
    python tools/profile_advect_target.py 512 1 >> "$O/ncu_advect.log" 2>&1
# 4. small / mid grids: launch lists with DRAM bytes and L2 hit rates (128^3, 256^3)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed
for n in 128 256; do
    timeout 600 ncu --metrics $M --clock-control none -s 20 -c 60 --csv --log-file "$O/launches_cg${n}.csv" \
        python tools/profile_target.py $n 60 1 > "$O/ncu_cg${n}.log" 2>&1
done
timeout 300 python tools/sweep_fused.py 128 > "$O/sweep_fused_128.log" 2>&1
# 5. multigrid: launch list of the V-cycle at 512^3 with DRAM bytes, solve timings
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file "$O/launches_mg512.csv" \
    python tools/profile_mg.py 512 cycle > "$O/ncu_mg512.log" 2>&1
timeout 600 python tools/profile_mg.py 512 solve > "$O/mg_solve_512.log" 2>&1
timeout 300 python tools/profile_mg.py 128 solve > "$O/mg_solve_128.log" 2>&1
# 6. the reference arm as the driver runs it
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > "$O/bench_reference_n1.json" 2> "$O/bench_reference_n1.err"
ls -la "$O" > "$O/listing.txt"
