// Shared device/host helpers for the hd_yolo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "hd_yolo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "hd_yolo_b200 kernels are written for sm_100a (B200) only"
#endif

namespace hdy {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define HDY_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      hdy::set_error(__VA_ARGS__);        \
      return HDY_ERR_INVALID;             \
    }                                     \
  } while (0)

// ---- exact fp32 arithmetic (the file is also built with -fmad=false) ---------
// torch's CUDA sigmoid is 1/(1+exp(-x)) in fp32 with the full-precision expf and an
// IEEE division; keep the same expression so device results track ATen's.
// 1/y rounded to nearest is one value whether it is produced by a division or by the correctly rounded reciprocal;
// __frcp_rn is the cheaper instruction sequence.
__device__ __forceinline__ float sigmoidf_ref(float x) { return __frcp_rn(__fadd_rn(1.0f, expf(-x))); }

// Order-preserving map float -> uint32 (ascending).  -0.0 is canonicalised to +0.0 so that
// ties compare equal exactly as torch's sort does.
__device__ __forceinline__ uint32_t orderable_u32(float f) {
  f = f + 0.0f;  // -0.0 -> +0.0
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable_u32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
// Ascending 64-bit key == descending score, ties by ascending index: the order
// torchvision.ops.nms visits boxes in (scores.sort(stable=True, descending=True)).
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  return ((uint64_t)(~orderable_u32(score)) << 32) | (uint64_t)idx;
}
__device__ __forceinline__ float key_score(uint64_t k) { return from_orderable_u32(~(uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return (uint32_t)k; }

// IoU > thr exactly as torchvision's kernels evaluate it:
//   area = (x2-x1)*(y2-y1); w = max(min(x2)-max(x1),0); inter = w*h; iou = inter/(Sa+Sb-inter)
__device__ __forceinline__ float box_area(const float4& b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
__device__ __forceinline__ bool iou_gt(const float4& a, const float4& b, float thr) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(__fsub_rn(xx2, xx1), 0.0f), h = fmaxf(__fsub_rn(yy2, yy1), 0.0f);
  float inter = __fmul_rn(w, h);
  if (!(inter > 0.0f)) return false;  // 0/x is 0 or NaN: never > thr (thr >= 0)
  float uni = __fsub_rn(__fadd_rn(box_area(a), box_area(b)), inter);
  // The verdict is fl(inter / uni) > thr.  Away from the threshold it can be read off one product: with
  // t = fl(thr * uni), inter > t * (1 + 1e-5) implies the rounded quotient exceeds thr, inter < t * (1 - 1e-5) implies
  // it does not (each rounding moves a value by at most 6e-8 relative); only the sliver in between pays for the IEEE
  // division.  NaN / inf operands fail both tests and take the division, as before.
  if (thr > 1.0e-3f && uni > 1.0e-30f) {  // products stay normal: the relative error bounds hold
    const float t = __fmul_rn(thr, uni);
    if (inter > __fmul_rn(t, 1.00001f)) return true;
    if (inter < __fmul_rn(t, 0.99999f)) return false;
  }
  return __fdiv_rn(inter, uni) > thr;
}

// Can IoU(a', b') exceed thr when every coordinate of a and b moves by at most e (the rounding of `box + tile
// origin` in slide coordinates, Detect.merge_outputs yolo_head.py:455) and the result is evaluated in fp32?
// Upper bound: the largest possible intersection over the smallest possible union.  Conservative by construction
// (false positives only cost work), never optimistic.
__device__ __forceinline__ bool iou_may_exceed(const float4& a, const float4& b, float thr, float e) {
  const float e2 = 2.0f * e;
  const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x) + e2, h = fminf(a.w, b.w) - fmaxf(a.y, b.y) + e2;
  if (!(w > 0.0f) || !(h > 0.0f)) return false;
  const float imax = w * h;
  const float wa = fmaxf(a.z - a.x - e2, 0.0f), ha = fmaxf(a.w - a.y - e2, 0.0f);
  const float wb = fmaxf(b.z - b.x - e2, 0.0f), hb = fmaxf(b.w - b.y - e2, 0.0f);
  const float umin = wa * ha + wb * hb - imax;
  if (!(umin > 0.0f)) return true;
  return imax * 1.0001f > thr * umin;  // 1e-4: fp32 rounding of this bound and of the IoU it bounds
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// ---- mbarrier / bulk-async (TMA) primitives ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Device-side copy of the level table (passed by value as a kernel parameter).
struct LevelDev {
  const float* ptr;  // fp16 levels (LevelTable::dtype == HDY_F16): the same address, read as __half
  int ny, nx;
  int rows;        // na*ny*nx rows per tile on this level
  int row_offset;  // first row of this level inside the concatenated [N] ordering
  int chunk_begin; // first chunk index of this level inside a tile
  float stride;
  float aw[HDY_MAX_ANCHORS], ah[HDY_MAX_ANCHORS];
};
struct LevelTable {
  LevelDev lv[HDY_MAX_LEVELS];
  int nl, na, no, nc, N, chunks_per_tile, layout, dtype;
};

// element i of a level's tensor, widened to fp32 (exact)
template <bool HALF>
__device__ __forceinline__ float level_elem(const float* base, size_t i) {
  if (HALF) return __half2float(reinterpret_cast<const __half*>(base)[i]);
  return base[i];
}

// nms_smem.cu: one CTA per tile, tiles with at most 4096 candidates (others return at once)
constexpr int kNmsSmemCap = 4096;
int launch_nms_tiles_smem(const uint64_t* cand_keys, const float4* cand_boxes, const float* cand_cls,
                          const int32_t* counts, int bs, int cap, float thr, float class_offset, int max_nms,
                          int max_det, int32_t* keep_idx, int32_t* keep_slot, float4* keep_box, float* keep_score,
                          float* keep_cls, int32_t* keep_counts, float gray_eps, uint8_t* keep_fragile,
                          unsigned long long* phase_cycles, cudaStream_t stream);

// filter_tma.cu: returns 1 when the layout does not meet the bulk-copy alignment rules (use the generic kernel)
int launch_filter_compact_tma(const hdy_level_t* levels_host, int nl, int bs, int na, int nc, int no, float conf_thres,
                              float min_size, int cap, uint64_t* cand_keys, float* cand_boxes, int32_t* counts,
                              int32_t* status, cudaStream_t stream);

int build_level_table(const hdy_level_t* levels_host, int nl, int na, int no, int layout, int rows_per_chunk,
                      LevelTable* out);

}  // namespace hdy
