#!/bin/bash
# One multi-GPU gpurun call:   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_session_multi.sh N'
# (N = 2, 4 or 8).  Everything lands in gpurun_out/session_nN/.  Never run under ncu.
set -u
N=${1:-2}
O=gpurun_out/session_n$N
mkdir -p "$O"
nvidia-smi -L > "$O/gpus.txt" 2>&1
nvidia-smi topo -m > "$O/topo.txt" 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"

# 1. parity on N GPUs: every exchange schedule, field halo, whole steps, the block-decomposed multigrid
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "$N or two_blocks" > "$O/pytest_multigpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
timeout 1200 python -m pytest tests/test_zzz_multigpu_late.py -m gpu -q -k "$N" > "$O/pytest_multigpu_late.log" 2>&1
echo "pytest (late) rc=$?" >> "$O/summary.txt"

# 2. the headline bench, weak scaling (default: NVLink peer-memory exchange), then the NCCL path, then strong scaling
timeout 600 $RUN --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > "$O/bench_weak_peer.json" 2> "$O/bench_weak_peer.err"
CFB_PEER=0 timeout 600 $RUN --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-timestep \
    > "$O/bench_weak_nccl.json" 2> "$O/bench_weak_nccl.err"
timeout 600 $RUN --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-e2e --no-timestep \
    > "$O/bench_strong_peer.json" 2> "$O/bench_strong_peer.err"
# 3. the 64-byte iteration on N GPUs
timeout 600 $RUN --master-port 29614 bench.py --gpus $N --steps 5 --warmup 3 --cg-variant 2 --no-e2e --no-timestep \
    > "$O/bench_weak_peer_variant2.json" 2> "$O/bench_weak_peer_variant2.err"
# 3b. ghost / reduction exchange inside the two kernels of the iteration (peer_fused), weak and strong
timeout 600 $RUN --master-port 29616 bench.py --gpus $N --steps 5 --warmup 3 --tune peer_fused=1 --no-e2e --no-timestep \
    > "$O/bench_weak_peer_fused.json" 2> "$O/bench_weak_peer_fused.err"
timeout 600 $RUN --master-port 29617 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --cells 256 --tune peer_fused=1 \
    --no-e2e --no-timestep > "$O/bench_strong256_peer_fused.json" 2> "$O/bench_strong256_peer_fused.err"
timeout 600 $RUN --master-port 29618 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --cells 256 \
    --no-e2e --no-timestep > "$O/bench_strong256_peer.json" 2> "$O/bench_strong256_peer.err"
timeout 600 $RUN --master-port 29619 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --tune peer_fused=1 \
    --no-e2e --no-timestep > "$O/bench_strong_peer_fused.json" 2> "$O/bench_strong_peer_fused.err"
# 4. the multigrid preconditioner at scale: Jacobi vs MG, peer vs NCCL ghost exchanges
timeout 900 $RUN --master-port 29615 tools/profile_mg_multi.py 512 > "$O/mg_multi.log" 2>&1
ls -la "$O" > "$O/listing.txt"
