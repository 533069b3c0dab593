#!/bin/bash
# Round 2, N-GPU call (N = 4 or 8):  gpurun --gpus N --timeout 600 -- 'bash tools/gpu_r2_session_multi.sh N'
# One rendezvous for every multi-GPU parity section (GPU time is charged per GPU), then the bench as the driver runs it.
set -u
N=${1:-8}
O=gpurun_out/r2m$N
mkdir -p "$O"
nvidia-smi topo -m > "$O/topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 420 $TR --master-port 29631 tests/mgpu_worker.py --cells 48 40 36 --steps 3 --everything > "$O/mgpu_worker_everything.log" 2>&1
echo "worker rc=$?" >> "$O/summary.txt"
timeout 420 $TR --master-port 29632 bench.py --gpus $N > "$O/bench_n$N.json" 2> "$O/bench_n$N.err"
echo "bench rc=$?" >> "$O/summary.txt"
ls -la "$O" > "$O/listing.txt"
