// Shared by the two RoIAlign kernels (roi_align.cu: exact operation order; roi_align_tc.cu: 3xTF32 tensor cores):
// the level table and torchvision's sample coordinates.
#pragma once
#include "hdy_common.cuh"

namespace hdy {

constexpr int kRoiMaxM = 16;         // pooled size
constexpr int kRoiMaxS = 4;          // sampling ratio

struct RoiLevels {
  const float* data[HDY_MAX_LEVELS];
  int h[HDY_MAX_LEVELS], w[HDY_MAX_LEVELS];
  float scale[HDY_MAX_LEVELS];
  int nl;
  int nhwc;   // 0: [bs][C][h][w] (the reference's layout); 1: channels-last [bs][h][w][C] (tf32x3 entry point only)
};

struct SampleTab {
  int low, high;
  float l, h;
};

// One axis of torchvision's pre_calc_for_bilinear_interpolate (roi_align_common.h): sample `i` of bin `p`.
__device__ __forceinline__ SampleTab roi_sample(float start, float bin, int p, int i, int grid, int size) {
  float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)grid));
  SampleTab t;
  if (!(v >= -1.0f && v <= (float)size)) {  // also NaN: the reference's `v < -1 || v > size` is false for NaN, but a
    t.low = -1;                             // NaN coordinate is outside the contract (finite boxes)
    t.high = -1;
    t.l = 0.f;
    t.h = 0.f;
    return t;
  }
  if (v <= 0.f) v = 0.f;
  int low = (int)v, high;
  if (low >= size - 1) {
    high = low = size - 1;
    v = (float)low;
  } else {
    high = low + 1;
  }
  t.low = low;
  t.high = high;
  t.l = __fsub_rn(v, (float)low);
  t.h = __fsub_rn(1.0f, t.l);
  return t;
}

// host side: validates and copies the level table (sets the error string)
int roi_levels_from_host(const hdy_feature_level_t* levels_host, int nl, RoiLevels* L);

}  // namespace hdy
