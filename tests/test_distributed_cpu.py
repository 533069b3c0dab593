"""CPU coverage of the N>1 path (gloo, world_size 2 and 4): decomposition rules, rank/block mapping,
NCCL-id bootstrap plumbing, and decomposition invariance of the block-wise algorithm with the same
ghost-exchange scheme the CUDA library runs over NCCL (see tests/dist_cpu_worker.py)."""
import os
import subprocess
import sys

import pytest

from cajitafluids_b200.distributed import block_grid, block_of_rank, neighbours, partition, rank_of_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_grid_splits_z_then_y_then_x():
    assert [block_grid(n) for n in (1, 2, 4, 8, 16)] == [(1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2), (2, 2, 4)]
    assert [block_grid(n, 2) for n in (1, 2, 4, 8)] == [(1, 1, 1), (1, 2, 1), (2, 2, 1), (2, 4, 1)]
    with pytest.raises(ValueError):
        block_grid(6)


def test_rank_block_mapping_roundtrip_and_neighbours():
    blocks = (2, 2, 2)
    for r in range(8):
        assert rank_of_block(block_of_rank(r, blocks), blocks) == r
    # rank 0 = block (0,0,0): +x is rank 1, +y rank 2, +z rank 4, walls on the low sides
    assert neighbours(0, blocks) == [-1, 1, -1, 2, -1, 4]
    assert neighbours(7, blocks) == [6, -1, 5, -1, 3, -1]
    assert neighbours(1, (1, 1, 2)) == [-1, -1, -1, -1, 0, -1]


def test_partition_covers_the_domain():
    for n, nb in ((512, 2), (100, 3), (7, 4)):
        parts = [partition(n, nb, b) for b in range(nb)]
        assert sum(p[0] for p in parts) == n
        assert all(parts[b][1] == sum(p[0] for p in parts[:b]) for b in range(nb))


@pytest.mark.parametrize("world,dim", [(2, 3), (2, 2), (4, 3)])
def test_gloo_block_run_matches_single_block(world, dim):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + 10 * world + dim),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py"), str(dim)]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-3000:])
    sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0
    assert "OK" in p.stdout
