// Geometry shared by the pyramid builders and the kernels that regenerate anchors in registers.
#pragma once

#include "rn_common.cuh"

namespace rn {

// The assign kernel works on rectangular patches of kTileH x kTileW cells so that the patch's
// bounding box is tight and most ground-truth boxes can be culled per patch.
#ifndef RN_ASSIGN_TILE_H
#define RN_ASSIGN_TILE_H 8
#endif
#ifndef RN_ASSIGN_TILE_W
#define RN_ASSIGN_TILE_W 8
#endif
constexpr int kTileH = RN_ASSIGN_TILE_H;   // (overridable for the sweep in profiles/r02_other_configs.md)
constexpr int kTileW = RN_ASSIGN_TILE_W;

#ifdef __CUDACC__
// tf.linspace element i: start + step * i, multiply and add rounded separately
// (dataset.py:19-20, utils.py:27-28).
__device__ __forceinline__ float cell_centre(float start, float step, int i) {
  return fadd(start, fmul(step, (float)i));
}
#endif

}  // namespace rn
