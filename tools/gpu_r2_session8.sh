#!/bin/bash
# Round 2, last 8-GPU call: the bench exactly as the driver runs it at N = 8, with the round's final code
set -u
O=gpurun_out/r2m8b
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 500 $TR --master-port 29651 bench.py --gpus 8 > "$O/bench_n8.json" 2> "$O/bench_n8.err"
echo "bench rc=$?" >> "$O/summary.txt"
ls -la "$O" > "$O/listing.txt"
