// ibk_tma.h -- tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point) for the tile kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "ibk_device.cuh"

namespace ibk
{
struct alignas(64) TmaMapSet
{
    CUtensorMap m[IBK_MAX_COMP];
};
// fp64 array of one component, box = (bx, by, bz) elements (bz ignored in 2D); false if TMA cannot address it
// promo: L2 promotion of the loads, 0 none, 1 64 B, 2 128 B, 3 256 B
bool make_tensor_map(CUtensorMap* m, const CompGeom& cg, int ndim, unsigned bx, unsigned by, unsigned bz, int promo = 2);
} // namespace ibk
