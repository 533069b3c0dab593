#!/bin/bash
# Round 2, second GPU call (2 GPUs):  gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_session2.sh'
set -u
O=gpurun_out/r2s2
mkdir -p "$O"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > "$O/gpu.csv" 2>&1
nvidia-smi topo -m > "$O/topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. multi-GPU parity at world 2: established tests, then overlap / single-reduction / multigrid / fused
timeout 900 python -m pytest tests/test_multigpu.py tests/test_zzz_multigpu_late.py -m gpu -q -p no:cacheprovider > "$O/pytest_multigpu_2gpu.log" 2>&1
echo "pytest multigpu rc=$?" >> "$O/summary.txt"
# 2. single-GPU news: single-reduction CG tests, advection timing with the new work mapping
timeout 600 python -m pytest tests/test_zz_c_single_reduction.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > "$O/pytest_1gpu_new.log" 2>&1
echo "pytest 1gpu rc=$?" >> "$O/summary.txt"
timeout 300 python tools/profile_advect.py 512 3 > "$O/advect_512.json" 2> "$O/advect_512.err"
timeout 300 python tools/profile_advect.py 128 5 >> "$O/advect_512.json" 2>> "$O/advect_512.err"
# 3. the bench at N = 2 as the driver runs it (exchange schedules, strong scaling, decomposition parity inside)
timeout 900 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench n2 rc=$?" >> "$O/summary.txt"
# 4. x split (strided faces): both schedules
for ov in 0 1; do
    timeout 600 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --blocks 2 1 1 --steps 3 --no-e2e --no-timestep --no-cpu-baseline \
        --tune peer_overlap=$ov > "$O/bench_n2_blocks211_overlap$ov.json" 2> "$O/bench_n2_blocks211_overlap$ov.err"
done
# 5. 1-GPU bench of the single-reduction form (88 B/cell, one reduction point)
timeout 600 python bench.py --steps 3 --cg-variant 3 --no-cpu-baseline --no-e2e --no-timestep --no-probe > "$O/bench_n1_variant3.json" 2> "$O/bench_n1_variant3.err"
# 6. BASELINE configs[0] as written: 64^3, 100 steps, GPU vs oracle
timeout 900 python tools/run_config0.py 64 100 > "$O/config0_64cubed_100steps.json" 2> "$O/config0.err"
ls -la "$O" > "$O/listing.txt"
