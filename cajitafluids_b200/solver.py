"""Host-side mirror of the reference's `createSolver` / `SolverBase` (src/Solver.hpp:41-48,283-350).

`create_solver(device, cfg)` keeps the reference's device-string dispatch: the only backend is
"b200" (alias "cuda"); anything else raises RuntimeError("invalid backend") exactly like
src/Solver.hpp:345-349 — there is no CPU fallback.
"""
from . import config as K
from ._capi import Context, load


class Solver(Context):
    """SolverBase: setup() / step() / solve(t_final, write_freq) + ProblemManager-style get/set."""

    def __init__(self, cfg):
        super().__init__(load(), cfg)

    # VelocityCorrectorBase::correctVelocity  (src/VelocityCorrector.hpp:37-42)
    def correctVelocity(self):
        return self.correct_velocity()


def create_solver(device, cfg, matrix_solver="Reference", preconditioner="None"):
    if device not in ("b200", "cuda"):
        raise RuntimeError("invalid backend")
    if matrix_solver != "Reference":
        # HYPRE solvers are out of scope; the reference would dispatch to hypre here
        raise RuntimeError("only the 'Reference' (Jacobi-PCG) matrix solver is implemented")
    return Solver(cfg)
