// Shared device helpers of the mask kernels (mask.cu, mask_regions.cu).
#pragma once
#include "hdy_common.cuh"

namespace hdy {

// ATen's bilinear (align_corners=False): scale = in/out (fp32); src = scale*(dst+0.5)-0.5, clamped at 0;
// i0 = (int)src; i1 = min(i0+1, in-1); l1 = src-i0; l0 = 1-l1; v = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11).
struct Lerp {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  Lerp L;
  L.i0 = min((int)src, in_size - 1);
  L.i1 = min(L.i0 + 1, in_size - 1);
  L.l1 = fminf(fmaxf(__fsub_rn(src, (float)L.i0), 0.f), 1.f);
  L.l0 = __fsub_rn(1.0f, L.l1);
  return L;
}

__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, const Lerp& X, const Lerp& Y) {
  const float top = __fadd_rn(__fmul_rn(X.l0, v00), __fmul_rn(X.l1, v01));
  const float bot = __fadd_rn(__fmul_rn(X.l0, v10), __fmul_rn(X.l1, v11));
  return __fadd_rn(__fmul_rn(Y.l0, top), __fmul_rn(Y.l1, bot));
}

// ------------------------------------------------------------------------------------- (B) process_mask
// Window of mask i in output pixels.  crop_mask keeps proto pixels with x1d <= col < x2d, y1d <= row < y2d
// (float compares against arange); with upsample the bilinear taps spread every kept pixel over its
// neighbours, so the output window is the pre-image of [p0-1, p1) under i0 = floor(src).
struct PMGeom {
  float x1d, y1d, x2d, y2d;  // down-scaled box (fp32, as the reference computes it)
  int px0, py0, px1, py1;    // kept proto pixel range [p0, p1)
  int x0, y0, w, h;          // output window
};

__device__ __forceinline__ int ceil_to_int_clamped(float v, int lo, int hi) {
  if (!(v > (float)lo)) return lo;  // also NaN
  if (v >= (float)hi) return hi;
  return (int)ceilf(v);
}

// Source coordinate of output pixel `dst` exactly as lerp_coord (ATen) computes it.
__device__ __forceinline__ float pm_src(int dst, float scale) {
  const float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  return src < 0.f ? 0.f : src;
}

// Shrinks the window [o0, o1) of one axis to the output pixels whose value can exceed 0.5.  Proto pixels outside
// [p0, p1) are cropped to zero and the kept ones are at most 1, so an output pixel left of the kept range is worth at
// most its weight on pixel p0: with i0 = p0 - 1 that weight is l1 = src - i0, i.e. the pixel is dead while
// src <= p0 - 0.5 (and it is exactly zero while i1 < p0).  Symmetrically on the right, where pixel p1 exists (p1 < in):
// dead once src >= p1 - 0.5 (l0 = 1 - l1 <= 0.5).  Rounding cannot lift such a value over 0.5: every product and sum
// in bilerp is monotone in its operands and 0.5 * (l0 + l1) rounds to 0.5.  src is monotone in the pixel index, so
// both ends are found from an estimate and a short walk.  At 4x a nucleus' window shrinks by 6 pixels per axis
// (36 -> 30: one 32-bit word per row instead of two).
__device__ __forceinline__ void pm_trim(int& o0, int& o1, int p0, int p1, int in_size, float scale) {
  if (o1 <= o0) return;
  if (p0 > 0) {
    const float tl = (float)p0 - 0.5f;
    int c = (int)floorf((float)p0 / scale - 0.5f) - 1;
    c = min(max(c, o0), o1);
    while (c < o1 && !(pm_src(c, scale) > tl)) ++c;
    while (c > o0 && pm_src(c - 1, scale) > tl) --c;
    o0 = c;
  }
  if (p1 < in_size && o1 > o0) {
    const float tr = (float)p1 - 0.5f;
    int c = (int)floorf((float)p1 / scale - 0.5f) - 1;
    c = min(max(c, o0), o1);
    while (c < o1 && !(pm_src(c, scale) >= tr)) ++c;
    while (c > o0 && pm_src(c - 1, scale) >= tr) --c;
    o1 = c;
  }
}

__device__ __forceinline__ PMGeom pm_geometry(const float4 b, int mh, int mw, int ih, int iw, int upsample,
                                              float rx, float ry) {
  PMGeom g;
  g.x1d = __fmul_rn(b.x, rx);  // downsampled_bboxes[:, 0] *= mw / iw
  g.x2d = __fmul_rn(b.z, rx);
  g.y1d = __fmul_rn(b.y, ry);
  g.y2d = __fmul_rn(b.w, ry);
  // r >= x1 & r < x2 over integers r: [ceil(x1), ceil(x2))
  g.px0 = ceil_to_int_clamped(g.x1d, 0, mw);
  g.px1 = ceil_to_int_clamped(g.x2d, 0, mw);
  g.py0 = ceil_to_int_clamped(g.y1d, 0, mh);
  g.py1 = ceil_to_int_clamped(g.y2d, 0, mh);
  // a NaN coordinate fails every crop comparison of the reference (r >= x1 ...): the mask is empty
  const bool nan = !(g.x1d == g.x1d) || !(g.x2d == g.x2d) || !(g.y1d == g.y1d) || !(g.y2d == g.y2d);
  if (nan) g.px1 = g.px0, g.py1 = g.py0;
  if (g.px1 <= g.px0 || g.py1 <= g.py0) {
    g.x0 = g.y0 = g.w = g.h = 0;
    return g;
  }
  if (!upsample) {
    g.x0 = g.px0;
    g.y0 = g.py0;
    g.w = g.px1 - g.px0;
    g.h = g.py1 - g.py0;
  } else {
    // conservative superset: src = s*(dst+0.5)-0.5 with s = mw/iw; i0 in [p0-1, p1-1]  <=>  src in [p0-1, p1)
    const float sx = (float)mw / (float)iw, sy = (float)mh / (float)ih;
    int ox0 = (int)floorf(((float)g.px0 - 0.5f) / sx - 0.5f) - 1, ox1 = (int)ceilf(((float)g.px1 + 0.5f) / sx - 0.5f) + 1;
    int oy0 = (int)floorf(((float)g.py0 - 0.5f) / sy - 0.5f) - 1, oy1 = (int)ceilf(((float)g.py1 + 0.5f) / sy - 0.5f) + 1;
    ox0 = max(ox0, 0);
    oy0 = max(oy0, 0);
    ox1 = min(ox1, iw);
    oy1 = min(oy1, ih);
    // ... trimmed to the pixels that CAN exceed 0.5 (exact, see pm_trim)
    pm_trim(ox0, ox1, g.px0, g.px1, mw, sx);
    pm_trim(oy0, oy1, g.py0, g.py1, mh, sy);
    g.x0 = ox0;
    g.y0 = oy0;
    g.w = max(ox1 - ox0, 0);
    g.h = max(oy1 - oy0, 0);
    if (g.w == 0 || g.h == 0) g.x0 = g.y0 = g.w = g.h = 0;
  }
  return g;
}


// mask_regions.cu: two-phase process_mask (TMA-staged prototypes -> sigmoid patches -> upsample + pack).  Returns 1
// when the shapes / workspace do not meet its requirements (the caller then uses the per-detection kernel in mask.cu).
constexpr int kPatchPitch = 16;  // sigmoid patches of up to 16 x 16 proto pixels go through the workspace
size_t process_mask_workspace_bytes(long long bs, long long max_det);
int launch_process_mask_regions(const void* protos, int proto_dtype, const float* coef, const float* boxes,
                                const int32_t* counts, int bs, int max_det, int nm, int mh, int mw, int ih, int iw, int upsample, float rx,
                                float ry, float* out_dense, const int32_t* geom, const int64_t* offsets, uint32_t* bits,
                                long long capacity_words, int32_t* status, void* workspace, size_t workspace_bytes,
                                cudaStream_t stream);
// mask.cu: the per-detection kernel over a device-side list of slots
int launch_process_mask_listed(const void* protos, int proto_dtype, const float* coef, const float* boxes,
                               const int32_t* counts, int max_det, int nm, int mh, int mw, int ih, int iw, int upsample, float rx, float ry,
                               float* out_dense, const int64_t* offsets, uint32_t* bits, long long capacity_words,
                               int32_t* status, const int32_t* slot_list, const int32_t* list_count, cudaStream_t st);

}  // namespace hdy
