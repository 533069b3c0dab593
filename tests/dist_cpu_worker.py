"""CPU worker for tests/test_distributed_cpu.py: world_size-2 (or 4) gloo run of the N>1 host logic.

Each rank owns one block, held by the CPU oracle in "one MPI rank of the reference" mode; the ghost
exchange and the CG allreduces the reference does through Cajita::Halo / MPI_Allreduce are done here
over torch.distributed (gloo) with EXACTLY the exchange scheme of cajitafluids_b200/csrc/halo.cu
(three axis sweeps, low ghosts <- last h layers, high ghosts <- first h+1 layers of the upper block;
face-neighbour width-1 exchange for the CG).  The result must match a single-block oracle run of the
same global problem: this validates the decomposition rules (face ownership, offsets, coordinates,
which layers travel) that the CUDA library implements with NCCL.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from cajitafluids_b200 import config as K
from cajitafluids_b200.distributed import attach_nccl, block_grid, decompose, neighbours
from helpers import fields_of, make_cfg, rel_l2
from oracle_api import Oracle

H = 3


def exchange(arrs, n_owned, nbr, d, layers_lo, layers_hi, tang):
    """One axis sweep over numpy views `arrs` (local ghosted indexing, [z,y,x]).
    tang(e, ext) -> slice of the tangential dim e."""
    reqs, recvs = [], []
    for a in arrs:
        ax = 2 - d  # numpy axis of spatial dim d
        ext = a.shape[ax]
        n = n_owned[d]
        nhi = min(layers_hi, ext - (n + H))  # h+1 only for faces normal to d
        sl = [tang(e, a.shape[2 - e]) for e in (2, 1, 0)]

        def box(lo, hi):
            s = list(sl)
            s[ax] = slice(lo, hi)
            return tuple(s)
        if nbr[2 * d] >= 0:  # low neighbour: send my first nhi owned layers, receive my low ghosts
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[box(H, H + nhi)])), nbr[2 * d]))
            buf = torch.empty(a[box(H - layers_lo, H)].shape, dtype=torch.float64)
            reqs.append(dist.irecv(buf, nbr[2 * d]))
            recvs.append((a, box(H - layers_lo, H), buf))
        if nbr[2 * d + 1] >= 0:  # high neighbour: send my last layers, receive [n, n + nhi)
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[box(H + n - layers_lo, H + n)])),
                                   nbr[2 * d + 1]))
            buf = torch.empty(a[box(H + n, H + n + nhi)].shape, dtype=torch.float64)
            reqs.append(dist.irecv(buf, nbr[2 * d + 1]))
            recvs.append((a, box(H + n, H + n + nhi), buf))
    for r in reqs:
        r.wait()
    for a, b, buf in recvs:
        a[b] = buf.numpy()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cells = (24, 20, 16) if dim == 3 else (48, 40)
    steps = 3
    blocks = block_grid(world, dim)
    box = tuple(c / cells[0] for c in cells)
    g = make_cfg(dim, cells, box=box)
    # the NCCL-id plumbing with a fake id source (no GPU here): every rank must see rank 0's bytes
    c = attach_nccl(decompose(g, rank, world, blocks), dist, get_id=lambda: bytes((7 * i + 1) % 256 for i in range(256)))
    assert bytes(c.nccl_id) == bytes((7 * i + 1) % 256 for i in range(256)) and c.use_nccl == 1
    assert (c.world_rank, c.world_size) == (rank, world)
    c.use_nccl = 0
    nbr = neighbours(rank, blocks, dim)
    blk = Oracle(c)
    n_owned = blk.owned_extent(K.QUANTITY)

    def owned_tang(d):
        # tangential coverage of sweep d: dims < d include their ghosts, dims >= d owned (+1)
        def t(e, ext):
            if e >= dim:
                return slice(0, 1)
            if e < d:
                return slice(0, ext)
            return slice(H, min(H + n_owned[e] + 1, ext))
        return t

    def gather(_user, what):
        if what == 0:  # ProblemManager::gather( Current ): width-h, all fields, three sweeps
            arrs = [blk.view(f) for f in fields_of(dim)]
            for d in range(dim):
                exchange(arrs, n_owned, nbr, d, H, H + 1, owned_tang(d))
        elif what in (1, 3, 4):  # CG x / p halo, pressure halo: width 1, face neighbours
            a = blk.view({1: K.PRESSURE, 3: K.CG_P, 4: K.PRESSURE}[what])
            for d in range(dim):
                def t(e, ext, d=d):
                    return slice(0, 1) if e >= dim else slice(H, H + n_owned[e])
                exchange([a], n_owned, nbr, d, 1, 1, t)

    def allreduce(_user, vals, n):
        t = torch.tensor([vals[i] for i in range(n)], dtype=torch.float64)
        dist.all_reduce(t)
        for i in range(n):
            vals[i] = float(t[i])

    blk.set_callbacks(gather, allreduce)
    ref = Oracle(g)  # the same global problem on one block
    off = blk.global_offset()

    def mine(field, arr):
        e = blk.owned_extent(field)
        return arr[off[2]:off[2] + e[2], off[1]:off[1] + e[1], off[0]:off[0] + e[0]]

    blk.setup()
    ref.setup()
    it_b, it_r = [blk.stats()["cg_iterations"]], [ref.stats()["cg_iterations"]]
    for _ in range(steps):
        blk.step()
        ref.step()
        it_b.append(blk.stats()["cg_iterations"])
        it_r.append(ref.stats()["cg_iterations"])
    bad = []
    if any(abs(a - b) > 1 + i for i, (a, b) in enumerate(zip(it_b, it_r))):
        bad.append(f"iterations {it_b} vs {it_r}")
    for f in fields_of(dim) + [K.PRESSURE]:
        full = ref.get(f)
        err = np.linalg.norm((blk.get(f) - mine(f, full)).ravel()) / max(np.linalg.norm(full.ravel()), 1e-300)
        if not err < 1e-10:
            bad.append(f"field {f}: {err}")
    # ghosts of the advected fields after a gather equal the neighbour's owned values
    blk.lib  # noqa
    gather(None, 0)
    for f in fields_of(dim):
        glob = ref.get(f, region=K.GHOSTED)
        m = blk.get(f, region=K.GHOSTED)
        want = glob[off[2]:off[2] + m.shape[0], off[1]:off[1] + m.shape[1], off[0]:off[0] + m.shape[2]]
        err = np.abs(m - want).max() / max(np.abs(glob).max(), 1e-300)
        if not err < 1e-10:
            bad.append(f"ghosts of field {f}: {err}")
    flag = torch.tensor([len(bad)])
    dist.all_reduce(flag)
    for b in bad:
        print(f"FAIL rank {rank}: {b}", flush=True)
    if rank == 0:
        print(f"dist_cpu_worker dim={dim} world={world} blocks={blocks}: {'OK' if int(flag) == 0 else 'FAILED'}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(flag) else 0)


if __name__ == "__main__":
    main()
