#!/bin/bash
# Round 2, session 20 (2 GPUs): the bench exactly as the driver runs it at N = 2, with this session's kernels
# (decomposition parity of the peer and NCCL paths, exchange schedules, strong scaling inside the line)
set -u
O=gpurun_out/r2s20
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 420 $TR --master-port 29671 bench.py --gpus 2 > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench rc=$?" >> "$O/summary.txt"
