#!/bin/bash
# Round 2, closing 2-GPU call: the bench exactly as the driver runs it at N = 2 (and the reference arm under torchrun)
set -u
O=gpurun_out/r2s12
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 600 $TR --master-port 29661 bench.py --gpus 2 > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench rc=$?" >> "$O/summary.txt"
