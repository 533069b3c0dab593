"""Host-side decomposition logic: block grid, rank <-> block mapping, NCCL bootstrap.

One process drives one GPU (torchrun); the 3-D grid is block-decomposed like the reference's
Cajita::DimBlockPartitioner does over MPI ranks (examples/advection.cpp:444, src/Mesh.hpp:89-94).
torch.distributed is used only as plumbing: to broadcast the two ncclUniqueIds that the C library
turns into its own communicators, and for barriers / max-over-ranks in bench.py.
"""
import ctypes as C

from . import config as K
from ._capi import load


def block_grid(n_ranks, dim=3):
    """1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2, 8 -> 2x2x2: split z first, then y, then x, so that the
    largest faces to pack are the contiguous ones.  Powers of two only."""
    if n_ranks < 1 or n_ranks & (n_ranks - 1):
        raise ValueError("number of ranks must be a power of two")
    dims = [1, 1, 1]
    d, k = dim - 1, n_ranks
    while k > 1:
        dims[d] *= 2
        k //= 2
        d = (d - 1) % dim
    return tuple(dims)


def block_of_rank(rank, blocks):
    """rank = (bz * py + by) * px + bx  (x fastest), the mapping cfb_create checks."""
    bz, r = divmod(rank, blocks[0] * blocks[1])
    by, bx = divmod(r, blocks[0])
    return bx, by, bz


def rank_of_block(b, blocks):
    return (b[2] * blocks[1] + b[1]) * blocks[0] + b[0]


def neighbours(rank, blocks, dim=3):
    """[-x, +x, -y, +y, -z, +z] neighbour ranks (-1 on physical walls; the mesh is non-periodic,
    src/Mesh.hpp:85-87)."""
    b = list(block_of_rank(rank, blocks))
    out = []
    for d in range(3):
        for s in (-1, 1):
            if d >= dim:
                out.append(-1)
                continue
            nb = list(b)
            nb[d] += s
            out.append(rank_of_block(nb, blocks) if 0 <= nb[d] < blocks[d] else -1)
    return out


def partition(n, nb, b):
    """Cajita's block split: the first n % nb blocks get one more cell.  -> (owned, offset)"""
    base, rem = divmod(n, nb)
    return base + (1 if b < rem else 0), b * base + min(b, rem)


def decompose(cfg, rank, world, blocks=None):
    """Fill the partitioner part of a config for `rank` of `world`."""
    blocks = blocks or block_grid(world, cfg.dim)
    if blocks[0] * blocks[1] * blocks[2] != world:
        raise ValueError("block grid does not match world size")
    out = K.copy_config(cfg)
    b = block_of_rank(rank, blocks)
    for d in range(3):
        out.ranks_per_dim[d] = blocks[d]
        out.block_id[d] = b[d]
    out.world_rank, out.world_size = rank, world
    return out


def attach_nccl(cfg, dist, get_id=None):
    """Rank 0 creates the two ncclUniqueIds, everybody receives them through torch.distributed
    (works over nccl and gloo), and the config is marked multi-GPU."""
    if get_id is None:
        def get_id():
            buf = (C.c_ubyte * (2 * K.NCCL_ID_BYTES))()
            lib = load()
            lib.check(lib.fn["nccl_unique_id"](buf))
            return bytes(buf)
    box = [get_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    raw = box[0]
    if len(raw) != 2 * K.NCCL_ID_BYTES:
        raise ValueError("bad NCCL id length")
    for i, v in enumerate(raw):
        cfg.nccl_id[i] = v
    cfg.use_nccl = 1
    return cfg
