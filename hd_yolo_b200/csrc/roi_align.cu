// Multi-scale RoIAlign: the gather in front of the reference's mask head.
//
// Reference: Detect.multiscale_roi_align, metayolo/models/yolo_head.py:279-299 -- for every pyramid level i,
// torchvision.ops.roi_align(features[i], boxes[levels == i], (M, M), spatial_scale = 1/stride, sampling_ratio = 2,
// aligned = ROI_ALIGN = False (:15)) scattered into a zero [K, C, M, M] result (M = mask_output_size // 2 = 14,
// C = dim_reduced = 256).  The arithmetic is torchvision's (third-party; CPU algorithm restated in
// oracle/roi_align_core.c): per output bin, S x S bilinear samples, each  w1*v1 + w2*v2 + w3*v3 + w4*v4  with
// w1 = hy*hx ... formed first, accumulated sample by sample (iy outer, ix inner), then divided by the count.  Every
// product and sum below is rounded separately in that same order, so results match the CPU op bit for bit on finite
// inputs.
//
// B200 mapping.  The op writes 200 KB per RoI (C*M*M fp32) and reads only a ~5x5 window of every channel, so it is
// bound by HBM writes -- if the arithmetic keeps up: 16 products + 4 weight products per sample, 4 samples per output.
// One launch covers all levels (no `torch.where(levels == i)` loop, no scatter: the level id routes the RoI to its
// feature map).  One CTA per (RoI, 32-channel chunk):
//   * warp 0 builds the RoI's sample tables (M*S entries per axis: low/high index, l/h weight; out-of-range samples get
//     zero weights) and the bounding window of all taps;
//   * the window of the chunk's 32 channels is staged in shared memory (nuclei: ~25 floats per channel), so every
//     feature element leaves L2 once per chunk; windows over kWinMax floats (huge boxes) are read in place;
//   * one thread per (channel pair, output row): it walks the x samples in order, keeps the 2S rows x (low, high)
//     columns of both channels in registers and reloads them only when the low column changes (a warp-uniform branch;
//     ~6 reloads per 28 samples), so the weight products are shared by two channels and LDS traffic is ~10% of the FP
//     work;
//   * results go through a shared staging tile and leave as coalesced 128-bit stores (the chunk's 32 x M x M floats
//     are contiguous in the output).
#include <limits.h>
#include "roi_align_common.cuh"

namespace hdy {

constexpr int kRoiChunk = 32;        // channels per CTA
constexpr int kRoiPairs = kRoiChunk / 2;
constexpr int kWinMax = 168;         // staged window floats per channel (e.g. 12 x 14)

struct RoiSmem {
  SampleTab ytab[kRoiMaxM * kRoiMaxS];
  SampleTab xtab[kRoiMaxM * kRoiMaxS];
  int ylo, yhi, xlo, xhi;  // window of all taps (valid samples only); ylo > yhi: no valid sample
  float win[kRoiChunk * kWinMax];
  float stage[kRoiChunk * kRoiMaxM * kRoiMaxM];
};

template <int S, bool STAGED>
__device__ __forceinline__ void roi_rows(const RoiSmem& Sm, const float* __restrict__ base, int plane, int pitch,
                                         int sx, int ylo, int xlo, int M, int ph, float count, float* __restrict__ st0,
                                         float* __restrict__ st1) {
  // this thread's 2S feature rows (two per y sample) and their weights
  int roff[2 * S];
  float ly[S], hy[S];
#pragma unroll
  for (int iy = 0; iy < S; ++iy) {
    const SampleTab Y = Sm.ytab[ph * S + iy];
    // an out-of-range sample has zero weights; point its taps at the window origin
    roff[2 * iy] = (Y.low < 0 ? 0 : Y.low - ylo) * pitch;
    roff[2 * iy + 1] = (Y.low < 0 ? 0 : Y.high - ylo) * pitch;
    ly[iy] = Y.l;
    hy[iy] = Y.h;
  }
  float v0[2 * S][2], v1[2 * S][2];  // [row][low/high column], channel 0 / 1
  int cur = -2;
  for (int pw = 0; pw < M; ++pw) {
    float s0[S][S], s1[S][S];
#pragma unroll
    for (int ix = 0; ix < S; ++ix) {
      const SampleTab X = Sm.xtab[pw * S + ix];
      if (X.low != cur) {  // warp-uniform: the table is the RoI's
        cur = X.low;
        const int cl = X.low < 0 ? 0 : X.low - xlo, chh = X.low < 0 ? 0 : X.high - xlo;
#pragma unroll
        for (int k = 0; k < 2 * S; ++k) {
          if (STAGED) {
            v0[k][0] = base[roff[k] + cl];
            v0[k][1] = base[roff[k] + chh];
            v1[k][0] = base[plane + roff[k] + cl];
            v1[k][1] = base[plane + roff[k] + chh];
          } else {
            v0[k][0] = __ldg(base + roff[k] + cl * sx);
            v0[k][1] = __ldg(base + roff[k] + chh * sx);
            v1[k][0] = __ldg(base + plane + roff[k] + cl * sx);
            v1[k][1] = __ldg(base + plane + roff[k] + chh * sx);
          }
        }
      }
#pragma unroll
      for (int iy = 0; iy < S; ++iy) {
        const float w1 = __fmul_rn(hy[iy], X.h), w2 = __fmul_rn(hy[iy], X.l);
        const float w3 = __fmul_rn(ly[iy], X.h), w4 = __fmul_rn(ly[iy], X.l);
        s0[iy][ix] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v0[2 * iy][0]), __fmul_rn(w2, v0[2 * iy][1])),
                                         __fmul_rn(w3, v0[2 * iy + 1][0])),
                               __fmul_rn(w4, v0[2 * iy + 1][1]));
        s1[iy][ix] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1[2 * iy][0]), __fmul_rn(w2, v1[2 * iy][1])),
                                         __fmul_rn(w3, v1[2 * iy + 1][0])),
                               __fmul_rn(w4, v1[2 * iy + 1][1]));
      }
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int iy = 0; iy < S; ++iy)
#pragma unroll
      for (int ix = 0; ix < S; ++ix) {
        a0 = __fadd_rn(a0, s0[iy][ix]);
        a1 = __fadd_rn(a1, s1[iy][ix]);
      }
    // S*S is a power of two for S = 1, 2, 4: dividing by it and multiplying by its (exact) reciprocal round alike
    if ((S & (S - 1)) == 0) {
      st0[pw] = __fmul_rn(a0, 1.0f / (float)(S * S));
      st1[pw] = __fmul_rn(a1, 1.0f / (float)(S * S));
    } else {
      st0[pw] = __fdiv_rn(a0, count);
      st1[pw] = __fdiv_rn(a1, count);
    }
  }
}

// one (RoI, 32-channel chunk); every thread of the CTA calls it
template <int S>
__device__ __forceinline__ void roi_align_one(RoiSmem& Sm, const RoiLevels& L, int bs, int C,
                                              const float* __restrict__ rois, const float* __restrict__ level_of,
                                              int M, int aligned, float* __restrict__ out, long long n, int cbase) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nch = min(kRoiChunk, C - cbase);
  const int MM = M * M;
  float* o = out + ((size_t)n * C + cbase) * MM;

  const float* r = rois + n * 5;
  const int lvl = level_of ? (int)level_of[n] : 0;
  const bool lvl_ok = level_of ? (level_of[n] == (float)lvl && lvl >= 0 && lvl < L.nl) : true;  // `levels == i`
  const int b = (int)r[0];
  bool live = lvl_ok && b >= 0 && b < bs;
  const int H = live ? L.h[lvl] : 1, W = live ? L.w[lvl] : 1;

  if (live && warp == 0) {
    const float scale = L.scale[lvl], off = aligned ? 0.5f : 0.0f;
    const float sw = __fsub_rn(__fmul_rn(r[1], scale), off), sh = __fsub_rn(__fmul_rn(r[2], scale), off);
    const float ew = __fsub_rn(__fmul_rn(r[3], scale), off), eh = __fsub_rn(__fmul_rn(r[4], scale), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!aligned) {
      rw = fmaxf(rw, 1.0f);
      rh = fmaxf(rh, 1.0f);
    }
    const float bin_h = __fdiv_rn(rh, (float)M), bin_w = __fdiv_rn(rw, (float)M);
    int ylo = INT_MAX, yhi = -1, xlo = INT_MAX, xhi = -1;
    for (int i = lane; i < M * S; i += 32) {
      const SampleTab Y = roi_sample(sh, bin_h, i / S, i % S, S, H);
      const SampleTab X = roi_sample(sw, bin_w, i / S, i % S, S, W);
      Sm.ytab[i] = Y;
      Sm.xtab[i] = X;
      if (Y.low >= 0) {
        ylo = min(ylo, Y.low);
        yhi = max(yhi, Y.high);
      }
      if (X.low >= 0) {
        xlo = min(xlo, X.low);
        xhi = max(xhi, X.high);
      }
    }
    ylo = __reduce_min_sync(0xffffffffu, ylo);
    xlo = __reduce_min_sync(0xffffffffu, xlo);
    yhi = __reduce_max_sync(0xffffffffu, yhi);
    xhi = __reduce_max_sync(0xffffffffu, xhi);
    if (lane == 0) {
      Sm.ylo = ylo;
      Sm.yhi = yhi;
      Sm.xlo = xlo;
      Sm.xhi = xhi;
    }
  }
  __syncthreads();
  if (live && (Sm.yhi < 0 || Sm.xhi < 0)) live = false;  // every sample out of range: all weights zero
  if (!live) {
    // `result = torch.zeros(...)` rows nobody fills (level outside [0, nl)), or all-zero weights
    for (int i = t; i < nch * MM; i += blockDim.x) o[i] = 0.f;
    return;
  }
  const int ylo = Sm.ylo, xlo = Sm.xlo;
  const int wh = Sm.yhi - ylo + 1, ww = Sm.xhi - xlo + 1;
  const bool staged = wh * ww <= kWinMax;
  // element (c, y, x) of image b: NCHW  b*C*H*W + c*H*W + y*W + x;  channels-last  b*H*W*C + (y*W + x)*C + c
  const int sc = L.nhwc ? 1 : H * W, sy = L.nhwc ? W * C : W, sx = L.nhwc ? C : 1;
  const float* feat = L.data[lvl] + (size_t)b * C * H * W + (size_t)cbase * sc;
  if (staged) {
    const int per = wh * ww;
    // lane = channel, warps stride over the window elements: the element index (and its row / column split) is
    // warp-uniform, so no per-thread division is left in this loop
    const int nwarp = blockDim.x >> 5;
    if (lane < nch) {
      const float* src = feat + (size_t)lane * sc + (size_t)ylo * sy + (size_t)xlo * sx;
      float* dstw = Sm.win + lane * per;
      for (int e = warp; e < per; e += nwarp) {
        const int y = e / ww, x = e - y * ww;
        dstw[e] = __ldg(src + (size_t)y * sy + (size_t)x * sx);
      }
    }
    // the odd channel of the last pair of a ragged chunk reads defined values
    if (nch & 1)
      for (int i = t; i < per; i += blockDim.x) Sm.win[nch * per + i] = 0.f;
    __syncthreads();
  }
  const int cp = t / M, ph = t - cp * M;
  if (cp < kRoiPairs && 2 * cp < nch) {
    const float count = (float)max(S * S, 1);
    float* st0 = &Sm.stage[(2 * cp) * MM + ph * M];
    float* st1 = st0 + MM;
    if (staged)
      roi_rows<S, true>(Sm, Sm.win + (2 * cp) * (wh * ww), wh * ww, ww, 1, ylo, xlo, M, ph, count, st0, st1);
    else  // the odd channel of a ragged last pair re-reads the even one (its results are not written)
      roi_rows<S, false>(Sm, feat + (size_t)(2 * cp) * sc, (2 * cp + 1 < nch) ? sc : 0, sy, sx, 0, 0, M, ph, count,
                         st0, st1);
  }
  __syncthreads();
  const int total = nch * MM;
  if (((MM & 3) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(Sm.stage);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int i = t; i < total / 4; i += blockDim.x) __stcs(o4 + i, s4[i]);
  } else {
    for (int i = t; i < total; i += blockDim.x) o[i] = Sm.stage[i];
  }
}

// only == NULL: CTA x = RoI.  Otherwise only[0] RoIs listed in only[1..] (what the tensor-core path of
// roi_align_tc.cu left: windows over 6 x 6 taps), strided by the CTAs.
template <int S>
__global__ void __launch_bounds__(kRoiPairs* kRoiMaxM, S <= 2 ? 3 : 1) roi_align_levels_kernel(
    const RoiLevels L, int bs, int C, const float* __restrict__ rois, const float* __restrict__ level_of, int M,
    int aligned, float* __restrict__ out, const int32_t* __restrict__ only) {
  extern __shared__ __align__(16) unsigned char roi_smem_raw[];
  RoiSmem& Sm = *reinterpret_cast<RoiSmem*>(roi_smem_raw);
  const int cbase = blockIdx.y * kRoiChunk;
  if (!only) {
    roi_align_one<S>(Sm, L, bs, C, rois, level_of, M, aligned, out, (long long)blockIdx.x, cbase);
    return;
  }
  const int count = only[0];
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    roi_align_one<S>(Sm, L, bs, C, rois, level_of, M, aligned, out, (long long)only[1 + i], cbase);
    __syncthreads();  // the tables and staging tiles are re-used
  }
}


int roi_levels_from_host(const hdy_feature_level_t* levels_host, int nl, RoiLevels* Lp) {
  HDY_REQUIRE(levels_host != nullptr, "roi_align: levels is NULL");
  HDY_REQUIRE(nl >= 1 && nl <= HDY_MAX_LEVELS, "roi_align: nl=%d out of range [1,%d]", nl, HDY_MAX_LEVELS);
  RoiLevels& L = *Lp;
  memset(&L, 0, sizeof(L));
  L.nl = nl;
  for (int i = 0; i < nl; ++i) {
    HDY_REQUIRE(levels_host[i].data != nullptr, "roi_align: levels[%d].data is NULL", i);
    HDY_REQUIRE(levels_host[i].h >= 1 && levels_host[i].w >= 1, "roi_align: levels[%d] is empty", i);
    HDY_REQUIRE((long long)levels_host[i].h * levels_host[i].w < (1ll << 30), "roi_align: levels[%d] too large", i);
    L.data[i] = levels_host[i].data;
    L.h[i] = levels_host[i].h;
    L.w[i] = levels_host[i].w;
    L.scale[i] = levels_host[i].spatial_scale;
  }
  return HDY_OK;
}

int roi_align_args_ok(int bs, int channels, const float* rois, const float* level_of, int nl, int64_t K, int pooled,
                      int sampling_ratio, const float* out) {
  HDY_REQUIRE(bs >= 1 && channels >= 1, "roi_align: bs=%d channels=%d", bs, channels);
  HDY_REQUIRE(pooled >= 1 && pooled <= kRoiMaxM, "roi_align: output size %d out of range [1,%d]", pooled, kRoiMaxM);
  HDY_REQUIRE(sampling_ratio >= 1 && sampling_ratio <= kRoiMaxS,
              "roi_align: sampling_ratio=%d out of range [1,%d] (the adaptive grid of sampling_ratio<=0 is not on the "
              "reference's path)", sampling_ratio, kRoiMaxS);
  HDY_REQUIRE(K >= 0 && K < (1ll << 31), "roi_align: K out of range");
  if (K == 0) return HDY_OK;
  HDY_REQUIRE(rois != nullptr && out != nullptr, "roi_align: NULL pointer");
  HDY_REQUIRE(nl == 1 || level_of != nullptr, "roi_align: level ids required with more than one level");
  return HDY_OK;
}

// the exact-order kernel over all RoIs (only == NULL) or over those with only[n] != 0
int launch_roi_align_exact(const RoiLevels& L, int bs, int channels, const float* rois, const float* level_of,
                           int64_t K, int pooled, int sampling_ratio, int aligned, float* out, const int32_t* only,
                           cudaStream_t st) {
  const int chunks = (channels + kRoiChunk - 1) / kRoiChunk;
  HDY_REQUIRE(chunks <= 65535, "roi_align: too many channels");
  const int threads = ((kRoiPairs * pooled + 31) / 32) * 32;
  const dim3 grid((unsigned)(only ? (K < 1184 ? K : 1184) : K), (unsigned)chunks);
  const size_t smem = sizeof(RoiSmem);
#define HDY_ROI(SS)                                                                                                  \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(roi_align_levels_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         (int)smem);                                                                 \
    if (e != cudaSuccess) {                                                                                          \
      set_error("roi_align setup: %s", cudaGetErrorString(e));                                                       \
      return HDY_ERR_CUDA;                                                                                           \
    }                                                                                                                \
    roi_align_levels_kernel<SS><<<grid, threads, smem, st>>>(L, bs, channels, rois, level_of, pooled, aligned, out,  \
                                                             only);                                                  \
  } while (0)
  switch (sampling_ratio) {
    case 1: HDY_ROI(1); break;
    case 2: HDY_ROI(2); break;
    case 3: HDY_ROI(3); break;
    default: HDY_ROI(4); break;
  }
#undef HDY_ROI
  return check_launch("hdy_multiscale_roi_align");
}

}  // namespace hdy

extern "C" int hdy_multiscale_roi_align(const hdy_feature_level_t* levels_host, int nl, int bs, int channels,
                                        const float* rois, const float* level_of, int64_t K, int pooled,
                                        int sampling_ratio, int aligned, float* out, hdy_stream_t stream) {
  using namespace hdy;
  HDY_REQUIRE(levels_host != nullptr, "roi_align: levels is NULL");
  HDY_REQUIRE(nl >= 1 && nl <= HDY_MAX_LEVELS, "roi_align: nl=%d out of range [1,%d]", nl, HDY_MAX_LEVELS);
  int rc = roi_align_args_ok(bs, channels, rois, level_of, nl, K, pooled, sampling_ratio, out);
  if (rc || K == 0) return rc;
  RoiLevels L;
  rc = roi_levels_from_host(levels_host, nl, &L);
  if (rc) return rc;
  return launch_roi_align_exact(L, bs, channels, rois, level_of, K, pooled, sampling_ratio, aligned, out, nullptr,
                                (cudaStream_t)stream);
}
