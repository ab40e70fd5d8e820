// Region-centric process_mask: the north-star mask kernel (prototype x coefficient contraction in fp32 FFMA, sigmoid,
// box crop, bilinear upsample, > 0.5, bit-packed), with the prototypes staged through TMA.
//
// Semantics: ultralytics/yolov5 v7 utils/segment/general.py::process_mask + crop_mask (restated in oracle/port.py;
// the reference repo has no such function), identical to process_mask_kernel in mask.cu.
//
// Why regions.  A tile's prototypes are [32, mh, mw] fp32 (8.4 MB at 1024 px) and every detection touches a ~10x10
// window of all 32 planes.  Letting each detection fetch its own window (mask.cu) moves ~13 KB per detection through
// L2 in 40-byte row fragments.  Here the proto plane is cut into regions of 20x23 pixels; one CTA owns (tile, region):
//   * ONE 4-D TMA tile load (cp.async.bulk.tensor.4d, SASS UTMALDG) brings the region's 24x24x32 box (one extra
//     column/row for the second bilinear tap; out-of-bounds is zero-filled by the TMA unit) into 72 KB of shared
//     memory, so every proto element is read from L2 (24/20)(24/23) = 1.25 times and from HBM once;
//   * while the load is in flight the CTA scans the tile's boxes and lists the detections whose taps reach into the
//     region (a nucleus is split over 1-4 regions; big boxes simply over more);
//   * one warp per listed detection: coefficients to registers, cropped sigmoid(coef . proto) for the part of the
//     box inside the region into a per-warp patch (32 LDS + 32 FFMA per pixel), then the output pixels whose FIRST
//     bilinear tap lies in the region are interpolated from the patch in ATen's operation order, thresholded and
//     ballot-packed; a 32-pixel word owned entirely by this piece is stored, a word shared with the neighbouring
//     region is OR-ed in atomically.
#include <stdlib.h>
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include "mask_common.cuh"

namespace hdy {

constexpr int kRegBox = 24;         // staged proto pixels per side
constexpr int kRegRY = kRegBox - 1; // first-tap rows owned per region (one more staged row for the second tap)
constexpr int kRegRX = 20;          // first-tap columns owned per region: the TMA tile origin must be 16-byte aligned
                                    // in the innermost dimension (measured: tools/micro/tma4d_test.cu), so the x pitch
                                    // is a multiple of 4 floats and 3 of the 24 staged columns are slack
constexpr int kRegNm = 32;
constexpr int kRegThreads = 256;
constexpr int kRegWarps = kRegThreads / 32;
constexpr int kRegList = 512;       // detections examined per pass
constexpr int kRegBatch = 64;       // listed detections prepared (records + coefficients) at a time
constexpr int kRowTab = 48;         // output rows interpolated per pass

struct RowTab {
  int off0, off1;  // patch offsets of the two source rows
  float l0, l1;
};

// Everything about one (detection, region) piece that is the same for all lanes, computed once by ONE thread of the
// prepare step instead of redundantly by the 32 lanes of the warp that processes the piece.
struct PieceRec {
  int d, gx0, gy0, gw;          // detection slot inside the tile; output window origin and width
  int px0, px1, py0, py1;       // kept proto pixels [p0, p1)
  float x1d, x2d, y1d, y2d;     // down-scaled box (crop test)
  int ox_lo, ox_hi, oy_lo, oy_hi;  // output pixels this region produces
  long long off;                // first word of the mask's bit plane
  int live, pad;
};

struct RegSmem {
  float proto[kRegNm][kRegBox][kRegBox];       // TMA destination (dense, x fastest)
  float patch[kRegWarps][kRegBox * kRegBox];   // cropped sigmoid values, staged coordinates
  RowTab rows[kRegWarps][kRowTab];
  PieceRec rec[kRegBatch];
  float coef[kRegBatch][kRegNm];
  uint16_t list[kRegList];
  int nlist;
  int pad;
  uint64_t bar;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

// smallest o in [lo, hi] whose first tap is >= target (hi if there is none); i0 is non-decreasing in o.
// Closed-form guess from src = scale*(o+0.5)-0.5, then exact fix-up against lerp_coord itself.
__device__ __forceinline__ int first_tap_ge(int target, int lo, int hi, float scale, int in_size) {
  if (target <= 0) return lo;
  int o = (int)ceilf(((float)target + 0.5f) / scale - 0.5f);
  o = max(lo, min(hi, o));
  while (o > lo && lerp_coord(o - 1, scale, in_size).i0 >= target) --o;
  while (o < hi && lerp_coord(o, scale, in_size).i0 < target) ++o;
  return o;
}

template <bool PACKED, bool UPSAMPLE>
__global__ void __launch_bounds__(kRegThreads, 2) process_mask_regions_kernel(
    const __grid_constant__ CUtensorMap tmap, const float* __restrict__ coef, const float4* __restrict__ boxes,
    const int32_t* __restrict__ counts, int max_det, int mh, int mw, int ih, int iw, int rxn, int ryn, float rx,
    float ry, float* __restrict__ out_dense, const int64_t* __restrict__ offsets, uint32_t* __restrict__ bits,
    long long capacity_words, int32_t* __restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RegSmem& S = *reinterpret_cast<RegSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int per_tile = rxn * ryn;
  const int tile = blockIdx.x / per_tile, reg = blockIdx.x - tile * per_tile;
  const int RY = reg / rxn, RX = reg - RY * rxn;
  const int X0 = RX * kRegRX, Y0 = RY * kRegRY;
  const int n = min(counts[tile], max_det);
  if (n <= 0) return;

  if (t == 0) {
    mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(&S.bar, (uint32_t)sizeof(S.proto));
    tma_load_4d(&S.proto[0][0][0], &tmap, X0, Y0, 0, tile, &S.bar);
  }
  const float sxs = (float)mw / (float)iw, sys = (float)mh / (float)ih;  // ATen: scale = in / out (fp32)
  const int oh = UPSAMPLE ? ih : mh, ow = UPSAMPLE ? iw : mw;
  constexpr int HALO = UPSAMPLE ? 1 : 0;
  float* patch = S.patch[warp];
  RowTab* rowtab = S.rows[warp];
  bool loaded = false;

  for (int base = 0; base < n; base += kRegList) {
    __syncthreads();
    if (t == 0) S.nlist = 0;
    __syncthreads();
    // ---- which detections reach into this region?
    const int lim = min(base + kRegList, n);
    for (int d = base + t; d < lim; d += kRegThreads) {
      const float4 b = boxes[(size_t)tile * max_det + d];
      const int px0 = ceil_to_int_clamped(__fmul_rn(b.x, rx), 0, mw), px1 = ceil_to_int_clamped(__fmul_rn(b.z, rx), 0, mw);
      const int py0 = ceil_to_int_clamped(__fmul_rn(b.y, ry), 0, mh), py1 = ceil_to_int_clamped(__fmul_rn(b.w, ry), 0, mh);
      if (px1 <= px0 || py1 <= py0) continue;
      // first taps that see a kept pixel: [p0 - HALO, p1 - 1]
      if (px1 - 1 < X0 || px0 - HALO >= X0 + kRegRX || py1 - 1 < Y0 || py0 - HALO >= Y0 + kRegRY) continue;
      S.list[atomicAdd(&S.nlist, 1)] = (uint16_t)(d - base);
    }
    __syncthreads();
    const int nl = S.nlist;

    for (int b0 = 0; b0 < nl; b0 += kRegBatch) {
      const int nb = min(kRegBatch, nl - b0);
      // ---- prepare: one thread per piece computes its record; everybody fetches the coefficients
      if (t < nb) {
        PieceRec R;
        R.d = base + S.list[b0 + t];
        const size_t slot = (size_t)tile * max_det + R.d;
        const PMGeom g = pm_geometry(boxes[slot], mh, mw, ih, iw, UPSAMPLE ? 1 : 0, rx, ry);
        R.gx0 = g.x0;
        R.gy0 = g.y0;
        R.gw = g.w;
        R.px0 = g.px0;
        R.px1 = g.px1;
        R.py0 = g.py0;
        R.py1 = g.py1;
        R.x1d = g.x1d;
        R.x2d = g.x2d;
        R.y1d = g.y1d;
        R.y2d = g.y2d;
        R.live = g.w > 0 && g.h > 0;
        R.off = 0;
        R.pad = 0;
        if (UPSAMPLE) {
          const int tx_lo = max(g.px0 - 1, X0), tx_hi = min(g.px1, X0 + kRegRX);  // first taps [tx_lo, tx_hi)
          const int ty_lo = max(g.py0 - 1, Y0), ty_hi = min(g.py1, Y0 + kRegRY);
          R.ox_lo = first_tap_ge(tx_lo, g.x0, g.x0 + g.w, sxs, mw);
          R.ox_hi = first_tap_ge(tx_hi, g.x0, g.x0 + g.w, sxs, mw);
          R.oy_lo = first_tap_ge(ty_lo, g.y0, g.y0 + g.h, sys, mh);
          R.oy_hi = first_tap_ge(ty_hi, g.y0, g.y0 + g.h, sys, mh);
        } else {  // output pixel == proto pixel; this region owns [X0, X0+RX) x [Y0, Y0+RY)
          R.ox_lo = max(g.px0, X0);
          R.ox_hi = min(g.px1, X0 + kRegRX);
          R.oy_lo = max(g.py0, Y0);
          R.oy_hi = min(g.py1, Y0 + kRegRY);
        }
        if (R.ox_hi <= R.ox_lo || R.oy_hi <= R.oy_lo) R.live = 0;
        if (PACKED && R.live) {
          R.off = offsets[slot];
          if (R.off + (long long)((g.w + 31) >> 5) * g.h > capacity_words) {
            atomicOr(status, HDY_STATUS_OVERFLOW);
            R.live = 0;
          }
        }
        S.rec[t] = R;
      }
      for (int i = t; i < nb * kRegNm; i += kRegThreads) {
        const int e = i / kRegNm, c = i - e * kRegNm;
        S.coef[e][c] = coef[((size_t)tile * max_det + base + S.list[b0 + e]) * kRegNm + c];
      }
      __syncthreads();
      if (!loaded) {
        while (!mbar_try_wait(&S.bar, 0)) {
        }
        loaded = true;
      }

      // ---- one warp per piece
      for (int e = warp; e < nb; e += kRegWarps) {
        const PieceRec& R = S.rec[e];
        if (!R.live) continue;
        const size_t slot = (size_t)tile * max_det + R.d;
        const int wpr = (R.gw + 31) >> 5;
        float cf[kRegNm];
#pragma unroll
        for (int c = 0; c < kRegNm; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&S.coef[e][c]);
          cf[c] = v.x;
          cf[c + 1] = v.y;
          cf[c + 2] = v.z;
          cf[c + 3] = v.w;
        }
        // ---- cropped sigmoid(coef . proto) on the part of the box (+ halo) inside the staged box; lanes cover
        //      floor(32 / pw) patch rows at a time
        const int wx0 = max(R.px0 - HALO, X0), wx1 = min(R.px1 + HALO, X0 + kRegRX + HALO);
        const int wy0 = max(R.py0 - HALO, Y0), wy1 = min(R.py1 + HALO, Y0 + kRegRY + HALO);
        const int pw = wx1 - wx0;
        if (pw <= 0 || wy1 <= wy0) continue;
        {
          const int rows_per = 32 / pw;  // pw <= 24
          const int ly = lane / pw, lx = lane - ly * pw;
          const int xx = wx0 + lx, sx = xx - X0;
          const bool col_ok = ly < rows_per;
          const bool x_in = xx < mw && (float)xx >= R.x1d && (float)xx < R.x2d;
          __syncwarp();
          for (int yy = wy0 + ly; yy < wy1; yy += rows_per) {
            if (!col_ok) break;
            const int sy = yy - Y0;
            float v = 0.f;
            if (x_in && yy < mh && (float)yy >= R.y1d && (float)yy < R.y2d) {
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < kRegNm; ++c) acc = fmaf(cf[c], S.proto[c][sy][sx], acc);
              v = sigmoidf_ref(acc);
            }
            patch[sy * kRegBox + sx] = v;
          }
          __syncwarp();
        }

        const int w_lo = (R.ox_lo - R.gx0) >> 5, w_hi = (R.ox_hi - 1 - R.gx0) >> 5;
        if (!UPSAMPLE) {
          for (int w = w_lo; w <= w_hi; ++w) {
            const int xx = R.gx0 + (w << 5) + lane;
            const bool valid = xx >= R.ox_lo && xx < R.ox_hi;
            const bool full = (R.gx0 + (w << 5) >= R.ox_lo) && (min(R.gx0 + (w << 5) + 32, R.gx0 + R.gw) <= R.ox_hi);
            for (int yy = R.oy_lo; yy < R.oy_hi; ++yy) {
              const bool bit = valid && patch[(yy - Y0) * kRegBox + (valid ? xx - X0 : 0)] > 0.5f;
              if (PACKED) {
                const unsigned word = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) {
                  uint32_t* dst = bits + R.off + (long long)(yy - R.gy0) * wpr + w;
                  if (full)
                    *dst = word;
                  else if (word)
                    atomicOr(dst, word);
                }
              } else if (valid) {
                out_dense[slot * oh * ow + (size_t)yy * ow + xx] = bit ? 1.f : 0.f;
              }
            }
          }
          continue;
        }

        // ---- upsample: output pixels whose first tap lies in this region and sees the box
        for (int r0 = R.oy_lo; r0 < R.oy_hi; r0 += kRowTab) {
          const int nr = min(kRowTab, R.oy_hi - r0);
          __syncwarp();
          for (int r = lane; r < nr; r += 32) {
            const Lerp Y = lerp_coord(r0 + r, sys, mh);
            RowTab T;
            T.off0 = (Y.i0 - Y0) * kRegBox;
            T.off1 = (Y.i1 - Y0) * kRegBox;
            T.l0 = Y.l0;
            T.l1 = Y.l1;
            rowtab[r] = T;
          }
          __syncwarp();
          for (int w = w_lo; w <= w_hi; ++w) {
            const int ox = R.gx0 + (w << 5) + lane;
            const bool valid = ox >= R.ox_lo && ox < R.ox_hi;
            const bool full = (R.gx0 + (w << 5) >= R.ox_lo) && (min(R.gx0 + (w << 5) + 32, R.gx0 + R.gw) <= R.ox_hi);
            const Lerp X = lerp_coord(valid ? ox : R.ox_lo, sxs, mw);
            const float* p0 = patch + (X.i0 - X0);
            const float* p1 = patch + (X.i1 - X0);
            uint32_t* dst = bits + R.off + (long long)(r0 - R.gy0) * wpr + w;
            float* dd = out_dense + slot * oh * ow + (size_t)r0 * ow + ox;
            // An output row needs top = l0x*v00 + l1x*v01 of source row i0 and bot of source row i1; several output
            // rows share a source row (4 at the usual 4x upsample), so both are kept until the row table moves on.
            int prev0 = -1, prev1 = -1;
            float top = 0.f, bot = 0.f;
            unsigned myword = 0;
            for (int r = 0; r < nr; ++r) {
              const float4 rt = *reinterpret_cast<const float4*>(&rowtab[r]);
              const int o0 = __float_as_int(rt.x), o1 = __float_as_int(rt.y);
              if (o0 != prev0) {  // warp-uniform
                top = __fadd_rn(__fmul_rn(X.l0, p0[o0]), __fmul_rn(X.l1, p1[o0]));
                prev0 = o0;
              }
              if (o1 != prev1) {
                bot = __fadd_rn(__fmul_rn(X.l0, p0[o1]), __fmul_rn(X.l1, p1[o1]));
                prev1 = o1;
              }
              const float v = __fadd_rn(__fmul_rn(rt.z, top), __fmul_rn(rt.w, bot));
              const bool bit = valid && v > 0.5f;
              if (PACKED) {
                const unsigned word = __ballot_sync(0xffffffffu, bit);
                if ((r & 31) == lane) myword = word;
                if ((r & 31) == 31 || r == nr - 1) {  // lanes write the words of up to 32 rows at once
                  const int rr = (r & ~31) + lane;
                  if (rr <= r) {
                    if (full)
                      dst[(long long)rr * wpr] = myword;
                    else if (myword)
                      atomicOr(dst + (long long)rr * wpr, myword);
                  }
                }
              } else if (valid) {
                dd[(size_t)r * ow] = bit ? 1.f : 0.f;
              }
            }
          }
        }
      }
      __syncthreads();  // records and coefficients are overwritten by the next batch
    }
  }
  if (!loaded) {  // never leave with the bulk copy still in flight
    while (!mbar_try_wait(&S.bar, 0)) {
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <bool PACKED, bool UPSAMPLE>
static int launch_regions(const CUtensorMap& map, const float* coef, const float* boxes, const int32_t* counts, int bs,
                          int max_det, int mh, int mw, int ih, int iw, int rxn, int ryn, float rx, float ry,
                          float* out_dense, const int64_t* offsets, uint32_t* bits, long long capacity_words,
                          int32_t* status, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(process_mask_regions_kernel<PACKED, UPSAMPLE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegSmem));
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(process_mask_regions_kernel): %s", cudaGetErrorString(e));
    return HDY_ERR_CUDA;
  }
  const unsigned grid = (unsigned)((long long)bs * rxn * ryn);
  process_mask_regions_kernel<PACKED, UPSAMPLE><<<grid, kRegThreads, sizeof(RegSmem), st>>>(
      map, coef, reinterpret_cast<const float4*>(boxes), counts, max_det, mh, mw, ih, iw, rxn, ryn, rx, ry, out_dense,
      offsets, bits, capacity_words, status);
  return check_launch("hdy_process_mask(regions)");
}

int launch_process_mask_regions(const float* protos, const float* coef, const float* boxes, const int32_t* counts,
                                int bs, int max_det, int nm, int mh, int mw, int ih, int iw, int upsample, float rx,
                                float ry, float* out_dense, const int64_t* offsets, uint32_t* bits,
                                long long capacity_words, int32_t* status, cudaStream_t stream) {
  if (nm != kRegNm || (mw & 3) != 0 || ((uintptr_t)protos & 15) != 0 || max_det > 65535) return 1;
  if (getenv("HDY_MASK_GENERIC")) return 1;  // debugging aid: force the per-detection kernel
  const int rxn = (mw + kRegRX - 1) / kRegRX, ryn = (mh + kRegRY - 1) / kRegRY;
  if ((long long)bs * rxn * ryn >= (1ll << 31)) return 1;
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return 1;
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)mw, (cuuint64_t)mh, (cuuint64_t)nm, (cuuint64_t)bs};
  const cuuint64_t strides[3] = {(cuuint64_t)mw * 4, (cuuint64_t)mw * mh * 4, (cuuint64_t)mw * mh * nm * 4};
  const cuuint32_t box[4] = {kRegBox, kRegBox, kRegNm, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(protos), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1;
  const bool packed = out_dense == nullptr;
  if (packed)
    return upsample ? launch_regions<true, true>(map, coef, boxes, counts, bs, max_det, mh, mw, ih, iw, rxn, ryn, rx, ry,
                                                 nullptr, offsets, bits, capacity_words, status, stream)
                    : launch_regions<true, false>(map, coef, boxes, counts, bs, max_det, mh, mw, ih, iw, rxn, ryn, rx,
                                                  ry, nullptr, offsets, bits, capacity_words, status, stream);
  return upsample ? launch_regions<false, true>(map, coef, boxes, counts, bs, max_det, mh, mw, ih, iw, rxn, ryn, rx, ry,
                                                out_dense, nullptr, nullptr, 0, nullptr, stream)
                  : launch_regions<false, false>(map, coef, boxes, counts, bs, max_det, mh, mw, ih, iw, rxn, ryn, rx, ry,
                                                 out_dense, nullptr, nullptr, 0, nullptr, stream);
}

}  // namespace hdy
