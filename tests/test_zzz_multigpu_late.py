"""Multi-GPU tests of the opt-in paths (see tests/test_multigpu.py for the default ones): the block-decomposed multigrid
preconditioner, the overlapped exchange schedule of the CG iterations (`peer_overlap`) and the single-reduction CG
(`cg_variant` 3).  All are bit-identical to the single-block oracle in the multi-rank emulation
(tests/test_emulated_multirank.py)."""
import pytest

from test_multigpu import _run_worker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_block_decomposed_multigrid_matches_single_block_oracle(world):
    """The opt-in multigrid preconditioner runs the same global V-cycle on every decomposition (one-layer
    face exchange per operator application): V-cycle and MG-PCG bit for bit against the single-block oracle.
    (Checked on the CPU by tests/test_emulated_multirank.py; first GPU run pending.)"""
    _run_worker(world, ["--mg"], cells=(64, 32, 32))


@pytest.mark.parametrize("world,blocks", [(2, None), (2, (2, 1, 1)), (2, (1, 2, 1)), (4, None), (8, None)])
def test_overlapped_exchange_matches_single_block_oracle(world, blocks):
    """`peer_overlap`: the faces travel on the side stream under the interior units of phase B (r) and under the next
    phase A (search direction), the reductions run in the last blocks of the two compute kernels, boundary units
    last.  PCG bit for bit and whole steps against the single-block oracle, with the schedule on and off.
    (tests/test_emulated_multirank.py::test_overlapped_exchange_on_every_block_grid on the CPU.)"""
    extra = ["--overlap"] + (["--blocks"] + [str(b) for b in blocks] if blocks else [])
    _run_worker(world, extra, cells=(128, 24, 20) if blocks == (2, 1, 1) else (48, 40, 36))


@pytest.mark.parametrize("world,blocks", [(2, None), (2, (2, 1, 1)), (4, None), (8, None)])
def test_single_reduction_cg_matches_single_block_oracle(world, blocks):
    """`cg_variant` 3 (opt-in): one reduction point (three sums) and one ghost exchange (faces of r) per iteration, over
    the NVLink mailboxes and over NCCL; bit for bit the single-block checker's statement of the same algorithm."""
    extra = ["--cg1"] + (["--blocks"] + [str(b) for b in blocks] if blocks else [])
    _run_worker(world, extra, cells=(128, 24, 20) if blocks == (2, 1, 1) else (48, 40, 36))
