"""tests/emul/multirank.py — several ranks of the EMULATED library in one process, one thread per rank
(ctypes releases the GIL inside the library; the NCCL stand-in of nccl_emul.cpp passes the messages)."""
import ctypes as C
import os
import threading

from cajitafluids_b200 import config as K
from cajitafluids_b200._capi import Context
from cajitafluids_b200.distributed import block_grid, decompose


def run_ranks(lib, global_cfg, world, fn, blocks=None, timeout=300, peer=False):
    """fn(ctx, rank) -> result, on `world` rank threads; returns the list of results (raises the first
    exception of any rank)."""
    # peer=True: the NVLink peer-memory exchange of the CG iterations (the ranks share one address space, the
    # cudaIpc stand-ins hand out plain pointers, the exchange kernel runs with one fiber per CUDA thread)
    os.environ["CFB_PEER"] = "1" if peer else "0"
    blocks = blocks or block_grid(world, global_cfg.dim)
    ids = (C.c_ubyte * (2 * K.NCCL_ID_BYTES))()
    lib.check(lib.fn["nccl_unique_id"](ids))
    results, errors = [None] * world, [None] * world

    def body(rank):
        try:
            cfg = decompose(global_cfg, rank, world, blocks)
            for i in range(2 * K.NCCL_ID_BYTES):
                cfg.nccl_id[i] = ids[i]
            cfg.use_nccl = 1
            ctx = Context(lib, cfg)
            try:
                results[rank] = fn(ctx, rank)
            finally:
                ctx.close()
        except BaseException as e:  # noqa: BLE001
            errors[rank] = e

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    if any(t.is_alive() for t in threads):
        raise TimeoutError("a rank thread is stuck (deadlock in the emulated exchange?)")
    for e in errors:
        if e is not None:
            raise e
    return results


def block_slices(ctx, field):
    """numpy slices (z, y, x) of this rank's owned entities inside the global owned array."""
    off = ctx.global_offset()
    ext = ctx.owned_extent(field)
    return tuple(slice(off[d], off[d] + ext[d]) for d in (2, 1, 0))
