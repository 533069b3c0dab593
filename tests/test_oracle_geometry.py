"""Pins the oracle's mesh / index-space / state model against the reference's own unit tests
(the only assertions the reference has near this path):

  tests/tstMesh.cpp:16-30   MeshTest.BasicParameters
  tests/tstMesh.cpp:32-69   MeshTest.LocalGridSetup
  tests/tstProblemManager.cpp:23-59   ProblemManagerTest.StateArrayTest
  tests/tstProblemManager.cpp:61-98   ProblemManagerTest.HaloTest (index spaces of the shared region)

The fixture is the reference's: a 512^2 cell box of width 1.0 with halo 3 (tests/tstMesh.hpp:47-49).
"""
import ctypes as C

import numpy as np
import pytest

from cajitafluids_b200 import config as K
from cajitafluids_b200._capi import CfbError
from oracle_api import Oracle


def mesh_fixture(dim=2, blocks=(1, 1, 1), block=(0, 0, 0), n=512):
    cfg = K.default_config(dim, n)
    for d in range(dim):
        cfg.ranks_per_dim[d] = blocks[d]
        cfg.block_id[d] = block[d]
    cfg.world_size = int(np.prod(blocks[:dim]))
    return cfg


def test_basic_parameters():
    # tstMesh.cpp:19-26: cellSize == boxWidth / boxCells ; min index 0 ; max index cells-1
    o = Oracle(mesh_fixture())
    assert o.cell_size == 1.0 / 512
    A, _ = o.matrix()
    # min/max global cell index are what build_matrix keys on: the wall layers are exactly
    # global index 0 and 511 (coefficient toward the wall zeroed there and only there)
    h = 3
    assert (A[0, h:h + 512, h, 1] == 0).all() and (A[0, h:h + 512, h + 1, 1] != 0).all()
    assert (A[0, h:h + 512, h + 511, 2] == 0).all() and (A[0, h:h + 512, h + 510, 2] != 0).all()


@pytest.mark.parametrize("blocks", [(1, 1), (2, 1), (2, 2), (4, 2)])
def test_local_grid_setup(blocks):
    # tstMesh.cpp:47-53: owned cell extent == boxCells / dimNumBlock
    # tstMesh.cpp:61-68: ghosted Face<I> extents == n + 2*halo + 1 by n + 2*halo (halo also on walls)
    for bi in range(blocks[0]):
        for bj in range(blocks[1]):
            o = Oracle(mesh_fixture(2, blocks + (1,), (bi, bj, 0)))
            ex, ey, _ = o.owned_extent(K.QUANTITY)
            assert (ex, ey) == (512 // blocks[0], 512 // blocks[1])
            gz, gy, gx = o.shape(K.U, K.GHOSTED)
            assert (gx, gy) == (ex + 2 * 3 + 1, ey + 2 * 3)
            gz, gy, gx = o.view(K.U).shape
            assert (gx, gy) == (ex + 2 * 3 + 1, ey + 2 * 3)
            # faces shared between blocks belong to the upper block; only the block on the high wall
            # owns n + 1 faces along the normal
            fx, fy, _ = o.owned_extent(K.U)
            assert fx == ex + (1 if bi == blocks[0] - 1 else 0) and fy == ey
            assert o.global_offset()[:2] == (bi * ex, bj * ey)


def test_uneven_partition_matches_c_helper():
    # Cajita splits n = q*nb + r as r blocks of q+1 then nb-r blocks of q
    o_sizes = [Oracle(mesh_fixture(2, (3, 1, 1), (b, 0, 0), n=100)).owned_extent(K.QUANTITY)[0] for b in range(3)]
    assert o_sizes == [34, 33, 33]
    offs = [Oracle(mesh_fixture(2, (3, 1, 1), (b, 0, 0), n=100)).global_offset()[0] for b in range(3)]
    assert offs == [0, 34, 67]


def test_state_array_swap():
    # tstProblemManager.cpp:40-58: advance( Cell, Quantity ) makes Next the Current
    o = Oracle(mesh_fixture(n=64))
    ez, ey, ex = o.shape(K.QUANTITY)
    i, j = np.meshgrid(np.arange(ex) + 3, np.arange(ey) + 3)
    cur = (i * 100.0 + j * 10.0)[None]
    o.set(K.QUANTITY, cur, K.CURRENT)
    o.set(K.QUANTITY, cur + 5, K.NEXT)
    o.advance(K.QUANTITY)
    assert np.array_equal(o.get(K.QUANTITY, K.CURRENT), cur + 5)
    assert np.array_equal(o.get(K.QUANTITY, K.NEXT), cur)


def test_ghosts_on_walls_are_allocated_and_zero():
    # zero-filled once (ProblemManager.hpp:149-165), never refreshed
    o = Oracle(mesh_fixture(n=32))
    o.setup()
    for _ in range(2):
        o.step()
    for f in (K.QUANTITY, K.V):
        g = o.get(f, region=K.GHOSTED)[0]
        assert not g[:3].any() and not g[-3:].any() and not g[:, :3].any() and not g[:, -3:].any()


def test_mesh_rejects_non_uniform_extent():
    # src/Mesh.hpp:56-64 -> std::logic_error
    cfg = mesh_fixture(n=64)
    cfg.global_bounding_box[4] = 1.5
    with pytest.raises(CfbError) as e:
        Oracle(cfg)
    assert e.value.code == K.ERR_MESH_EXTENT


def test_coordinates_follow_cajita_local_mesh():
    o = Oracle(mesh_fixture(n=64))
    h = 1.0 / 64
    # first owned cell centre / I-face / J-face (local index 3 == global 0)
    assert o.coordinates(K.QUANTITY, (3, 3)) == (0.5 * h, 0.5 * h)
    assert o.coordinates(K.U, (3, 3)) == (0.0, 0.5 * h)
    assert o.coordinates(K.V, (3, 3)) == (0.5 * h, 0.0)
    # last I-face sits on the high wall
    assert o.coordinates(K.U, (3 + 64, 3))[0] == 1.0
