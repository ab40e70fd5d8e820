// Fused decode + confidence filter + stream compaction, TMA-staged (layout 0: [bs, na, ny, nx, no] rows).
//
// Same arithmetic and outputs as filter_compact_logits_kernel in decode.cu (reference: yolo_head.py:185-213,
// utils_general.py:121-128, :332, :336-337); this is the path taken whenever every chunk of rows starts on a
// 16-byte boundary (always, for the usual even grids).
//
// HBM -> SM:  a persistent CTA walks (tile, chunk) work items; a producer warp issues ONE bulk-async copy
//             (cp.async.bulk.shared::cluster.global, the 1-D TMA path: SASS UBLKCP) per chunk of ROWS rows
//             into a ring of shared-memory stages, each guarded by a full/empty mbarrier pair (the copy completes
//             `full` through complete_tx, the consumer warps release the stage through `empty`).  No thread
//             spends instructions on moving the 95 % of rows that are rejected.
// filter:     rows are rejected on the LOGIT: x < t_lo = logit(conf) - margin implies sigmoid(x) <= conf for
//             certain (margin = 1e-4 (1 + |logit|), ~1000x the error of expf + division); every other row
//             evaluates the reference expression sigmoid(x) > conf itself, so the verdict is identical to
//             comparing the fp32 sigmoid.  One LDS + one compare per rejected row.
// compact:    survivors decode their box (4 more LDS + sigmoids), are ranked inside their warp by ballot, and
//             the warp reserves its run in the tile's candidate list with one atomicAdd.
#include "hdy_common.cuh"

namespace hdy {

constexpr int kTmaThreads = 256;

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

struct TmaItem {
  const float* src;  // first float of the chunk
  int rows, row0, level, tile;
};

__device__ __forceinline__ TmaItem tma_item(const LevelTable& T, int w, int rows_per_chunk) {
  TmaItem it;
  it.tile = w / T.chunks_per_tile;
  const int chunk = w - it.tile * T.chunks_per_tile;
  int l = 0;
#pragma unroll 1
  for (int i = 1; i < T.nl; ++i)
    if (chunk >= T.lv[i].chunk_begin) l = i;
  const LevelDev& L = T.lv[l];
  it.level = l;
  it.row0 = (chunk - L.chunk_begin) * rows_per_chunk;
  it.rows = min(rows_per_chunk, L.rows - it.row0);
  it.src = L.ptr + ((size_t)it.tile * L.rows + it.row0) * T.no;
  return it;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kConsumerWarps = kTmaThreads / 32;        // 8 consumer warps
constexpr int kTmaBlock = kTmaThreads + 32;             // + 1 producer warp

// Warp roles: warp 8 is the PRODUCER (one elected lane re-arms a stage's `full` barrier and issues its bulk copy as
// soon as the 8 consumer warps have released it through the `empty` barrier); warps 0-7 are CONSUMERS, each owning
// 32*RPT rows of every chunk.  There is no block-wide barrier in the loop.  A consumer warp reserves its run in the
// tile's candidate list with one atomicAdd and writes the run one iteration later, so the atomic's round trip to L2
// overlaps the next chunk.
template <int RPT>  // rows per lane and chunk; a chunk is kTmaThreads * RPT rows
__global__ void __launch_bounds__(kTmaBlock) filter_compact_tma_kernel(
    const __grid_constant__ LevelTable T, int total_items, int stages, int stage_floats, float t_lo, float conf_thres,
    float min_size, int cap, uint64_t* __restrict__ cand_keys, float4* __restrict__ cand_boxes,
    int32_t* __restrict__ counts, int32_t* __restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int ROWS = kTmaThreads * RPT;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);            // [stages]
  uint64_t* empty = full + 8;                                        // [stages]
  float* data = reinterpret_cast<float*>(smem_raw + 128);            // stages x stage_floats
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int no = T.no;
  const int my_items = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (t == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------------------------------------ producer
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        const int s = it % stages;
        if (it >= stages) {
          const uint32_t parity = (uint32_t)(it / stages - 1) & 1u;
          while (!mbar_try_wait(&empty[s], parity)) {
          }
        }
        const TmaItem w = tma_item(T, (int)blockIdx.x + it * (int)gridDim.x, ROWS);
        const uint32_t bytes = ((uint32_t)(w.rows * no) * 4u) & ~15u;
        mbar_arrive_expect_tx(&full[s], bytes);
        if (bytes) bulk_copy_g2s(data + (size_t)s * stage_floats, w.src, bytes, &full[s]);
      }
    }
    return;
  }

  // -------------------------------------------------------------------------------------------- consumers
  // run of the previous iteration, written once its atomicAdd has returned
  bool p_cand[RPT];
  uint64_t p_key[RPT];
  float4 p_box[RPT];
  int p_tile = 0, p_base = 0, p_tot = 0;
  unsigned p_mask[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    p_cand[k] = false;
    p_key[k] = 0;
    p_box[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    p_mask[k] = 0;
  }
  auto flush = [&]() {
    if (p_tot == 0) return;
    int pos = __shfl_sync(0xffffffffu, p_base, 0);
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (p_cand[k]) {
        const int q = pos + __popc(p_mask[k] & ((1u << lane) - 1u));
        if (q < cap) {
          const size_t o = (size_t)p_tile * cap + q;
          cand_keys[o] = p_key[k];
          cand_boxes[o] = p_box[k];
        } else {
          atomicOr(status, HDY_STATUS_OVERFLOW);
        }
      }
      pos += __popc(p_mask[k]);
    }
    p_tot = 0;
  };

  for (int it = 0; it < my_items; ++it) {
    const TmaItem w = tma_item(T, (int)blockIdx.x + it * (int)gridDim.x, ROWS);
    const LevelDev& L = T.lv[w.level];
    const int s = it % stages;
    const uint32_t parity = (uint32_t)(it / stages) & 1u;
    const float* sm = data + (size_t)s * stage_floats;
    const int avail = ((w.rows * no) * 4 & ~15) >> 2;  // floats that arrive through the bulk copy
    while (!mbar_try_wait(&full[s], parity)) {
    }
    auto at = [&](int idx) -> float { return idx < avail ? sm[idx] : __ldg(w.src + idx); };

    bool cand[RPT];
    uint64_t key[RPT];
    float4 box[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rr = (warp * RPT + k) * 32 + lane;
      cand[k] = false;
      key[k] = 0;
      box[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr < w.rows) {
        const int base = rr * no;
        const float x = at(base + 4);
        if (x >= t_lo) {  // x < t_lo (or NaN): sigmoid(x) <= conf for certain
          const float p_obj = sigmoidf_ref(x);
          if (p_obj > conf_thres) {  // the reference's own comparison                  utils_general.py:336-337
            const int row = w.row0 + rr;
            const int plane = L.ny * L.nx;
            const int a = row / plane, p = row - a * plane;
            const int gy = p / L.nx, gx = p - gy * L.nx;
            // xy = (sigmoid*2 - 0.5 + grid) * stride ; wh = (sigmoid*2)^2 * anchor_grid     yolo_head.py:203-204
            const float sx = sigmoidf_ref(at(base)), sy = sigmoidf_ref(at(base + 1));
            const float sw = sigmoidf_ref(at(base + 2)), sh = sigmoidf_ref(at(base + 3));
            const float cx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.0f), 0.5f), (float)gx), L.stride);
            const float cy = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.0f), 0.5f), (float)gy), L.stride);
            const float tw = __fmul_rn(sw, 2.0f), th = __fmul_rn(sh, 2.0f);
            const float bw = __fmul_rn(__fmul_rn(tw, tw), L.aw[a]), bh = __fmul_rn(__fmul_rn(th, th), L.ah[a]);
            const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);  // xywh2xyxy utils_general.py:121-128
            box[k] = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
            // remove_small_boxes(min_size)                                             utils_general.py:332
            cand[k] = (__fsub_rn(box[k].z, box[k].x) >= min_size) && (__fsub_rn(box[k].w, box[k].y) >= min_size);
            key[k] = make_key(p_obj, (uint32_t)(L.row_offset + row));
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage

    flush();  // previous run: its base is in lane 0's register by now
    int tot = 0;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      p_mask[k] = __ballot_sync(0xffffffffu, cand[k]);
      p_cand[k] = cand[k];
      p_key[k] = key[k];
      p_box[k] = box[k];
      tot += __popc(p_mask[k]);
    }
    p_tot = tot;
    p_tile = w.tile;
    if (tot && lane == 0) p_base = atomicAdd(counts + w.tile, tot);  // consumed by the next flush()
  }
  flush();
}

// Host side: returns HDY_OK, an error, or 1 when the layout does not meet the bulk-copy alignment rules
// (the caller then uses the generic kernel).
int launch_filter_compact_tma(const hdy_level_t* levels_host, int nl, int bs, int na, int nc, int no, float conf_thres,
                              float min_size, int cap, uint64_t* cand_keys, float* cand_boxes, int32_t* counts,
                              int32_t* status, cudaStream_t stream) {
  const int rpt = no <= 16 ? 2 : 1;
  const int rows_per_chunk = kTmaThreads * rpt;
  const size_t stage_bytes = ((size_t)rows_per_chunk * no * 4 + 127) & ~(size_t)127;
  if (stage_bytes > 100 * 1024) return 1;
  if (!(conf_thres > 1e-6f && conf_thres < 1.0f - 1e-6f)) return 1;  // logit(conf) is not finite enough
  for (int l = 0; l < nl; ++l) {
    if (((uintptr_t)levels_host[l].logits & 15) != 0) return 1;
    const long long rows = (long long)na * levels_host[l].ny * levels_host[l].nx;
    if ((rows * no) % 4 != 0) return 1;  // tile bases must stay 16-byte aligned
  }
  LevelTable T;
  int rc = build_level_table(levels_host, nl, na, no, 0, rows_per_chunk, &T);
  if (rc) return rc;
  T.nc = nc;
  const long long total = (long long)bs * T.chunks_per_tile;
  if (total >= (1ll << 31)) return 1;
  int stages = (int)((96 * 1024) / stage_bytes);
  stages = stages < 2 ? 2 : (stages > 6 ? 6 : stages);
  const size_t smem = 128 + stages * stage_bytes;
  // sigmoid(x) > conf  <=>  x > logit(conf) up to rounding: decide far from the boundary on the logit
  const double lg = log((double)conf_thres / (1.0 - (double)conf_thres));
  const double margin = 1e-4 * (1.0 + fabs(lg));
  const float t_lo = (float)(lg - margin);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int per_sm = (int)((200 * 1024) / smem) < 1 ? 1 : (int)((200 * 1024) / smem);
  long long grid = (long long)sm_count * (per_sm > 4 ? 4 : per_sm);
  if (grid > total) grid = total;
  cudaError_t e;
  if (rpt == 2) {
    e = cudaFuncSetAttribute(filter_compact_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      filter_compact_tma_kernel<2><<<(unsigned)grid, kTmaBlock, smem, stream>>>(
          T, (int)total, stages, (int)(stage_bytes / 4), t_lo, conf_thres, min_size, cap, cand_keys,
          reinterpret_cast<float4*>(cand_boxes), counts, status);
  } else {
    e = cudaFuncSetAttribute(filter_compact_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      filter_compact_tma_kernel<1><<<(unsigned)grid, kTmaBlock, smem, stream>>>(
          T, (int)total, stages, (int)(stage_bytes / 4), t_lo, conf_thres, min_size, cap, cand_keys,
          reinterpret_cast<float4*>(cand_boxes), counts, status);
  }
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(filter_compact_tma_kernel): %s", cudaGetErrorString(e));
    return HDY_ERR_CUDA;
  }
  return check_launch("hdy_filter_compact_logits(tma)");
}

}  // namespace hdy
