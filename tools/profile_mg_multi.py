"""Multi-GPU timing of the opt-in multigrid preconditioner (weak scaling: n^3 cells per GPU), under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/profile_mg_multi.py [n=512]

One projection of the default inflow problem with Jacobi and with MG (ghost exchanges over NVLink peer memory,
then over NCCL): iterations, wall time (max over ranks), launches.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from cajitafluids_b200 import Solver, default_config
from cajitafluids_b200.distributed import attach_nccl, block_grid, decompose

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
blocks = block_grid(world)
gcells = tuple(n * b for b in blocks)

for kind, peer in (("jacobi", 1), ("mg", 1), ("mg", 0)):
    cfg = default_config(3, gcells, box=tuple(c / 512.0 for c in gcells))
    cfg.cg_max_iter = 40000
    cfg = decompose(cfg, rank, world, blocks)
    cfg.device_id = local
    attach_nccl(cfg, dist)
    s = Solver(cfg)
    s.set_tuning("peer_halo", peer)
    s.set_preconditioner(kind)
    s.add_inputs()
    s.build_rhs()
    s.pcg_solve()
    s.reset_stats()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it, res = s.pcg_solve()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} global={gcells} {kind} exchange={'peer' if peer else 'nccl'}: {it} iterations, |r|={res:.3e}, "
              f"{float(t[0]) * 1e3:.1f} ms, {s.stats()['kernel_launches']} launches", flush=True)
    s.close()
dist.destroy_process_group()
