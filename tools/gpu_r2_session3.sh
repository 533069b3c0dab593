#!/bin/bash
# Round 2, third GPU call (2 GPUs): the overlapped exchange, second form (boundary units on the side stream)
set -u
O=gpurun_out/r2s3
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_zzz_multigpu_late.py -m gpu -q -p no:cacheprovider -k "overlapped or single_reduction" > "$O/pytest_overlap_2gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
timeout 600 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --no-e2e --no-timestep --no-cpu-baseline > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench rc=$?" >> "$O/summary.txt"
for ov in 0 1; do
    timeout 600 $TR --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --steps 3 --no-e2e --no-timestep --no-cpu-baseline --no-probe \
        --tune peer_overlap=$ov > "$O/bench_n2_overlap$ov.json" 2> "$O/bench_n2_overlap$ov.err"
    timeout 600 $TR --nproc-per-node 2 --master-port 29623 bench.py --gpus 2 --blocks 2 1 1 --steps 3 --no-e2e --no-timestep --no-cpu-baseline --no-probe \
        --tune peer_overlap=$ov > "$O/bench_n2_blocks211_overlap$ov.json" 2> "$O/bench_n2_blocks211_overlap$ov.err"
    timeout 600 $TR --nproc-per-node 2 --master-port 29624 bench.py --gpus 2 --cells 256 --steps 5 --no-e2e --no-timestep --no-cpu-baseline --no-probe \
        --tune peer_overlap=$ov > "$O/bench_n2_256_overlap$ov.json" 2> "$O/bench_n2_256_overlap$ov.err"
done
ls -la "$O" > "$O/listing.txt"
