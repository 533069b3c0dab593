#!/bin/bash
# Round 2, sixth GPU call (2 GPUs): x faces stored straight into the neighbours' ghost columns (peer_xdirect), the new
# (record of what ran: the peer_xdirect key measured here lost and was removed afterwards)
# bench extras (multigrid timesteps weak-scaled)
set -u
O=gpurun_out/r2s6
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29641 tests/mgpu_worker.py --cells 128 24 20 --steps 3 --blocks 2 1 1 --quick > "$O/mgpu_worker_xsplit_quick.log" 2>&1
echo "worker rc=$?" >> "$O/summary.txt"
for xd in 0 1; do
    timeout 300 $TR --master-port 29642 bench.py --gpus 2 --blocks 2 1 1 --steps 5 --no-e2e --no-timestep --no-cpu-baseline --no-probe \
        --tune peer_xdirect=$xd > "$O/bench_n2_blocks211_xdirect$xd.json" 2> "$O/bench_n2_blocks211_xdirect$xd.err"
done
timeout 600 $TR --master-port 29643 bench.py --gpus 2 --no-cpu-baseline > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench rc=$?" >> "$O/summary.txt"
ls -la "$O" > "$O/listing.txt"
