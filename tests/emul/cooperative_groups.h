// cooperative_groups.h — HOST STAND-IN (tests/emul only).  The emulated cooperative launch runs ONE block (its threads
// are fibers with real barriers), so a grid barrier is the block barrier; kernels that use it must work for any grid
// size, which is what lets them be checked here.
#pragma once
#include <cuda_runtime.h>
namespace cooperative_groups
{
struct grid_group
{
    void sync() const { __syncthreads(); }
};
inline grid_group this_grid() { return grid_group(); }
} // namespace cooperative_groups
