// hnet multi-level heads (H1-H3 of SURVEY.md section 8a).  The reference's hnet/detection/mask_rcnn.py delegates all of
// this arithmetic to torchvision (mask_rcnn.py:67 BoxCoder.decode, :72 RegionProposalNetwork.filter_proposals, :192
// RoIHeads.postprocess_detections, :248 maskrcnn_inference), so the restatement below follows torchvision 0.26:
//
//   BoxCoder.decode_single (models/detection/_utils.py):
//       w = x2-x1; h = y2-y1; cx = x1 + 0.5*w; cy = y1 + 0.5*h
//       dx = d0/wx; dy = d1/wy; dw = min(d2/ww, clip); dh = min(d3/wh, clip)        clip = log(1000/16)
//       pcx = dx*w + cx; pcy = dy*h + cy; pw = exp(dw)*w; ph = exp(dh)*h
//       box = (pcx - 0.5*pw, pcy - 0.5*ph, pcx + 0.5*pw, pcy + 0.5*ph)
//   RPN.filter_proposals (models/detection/rpn.py): per level top pre_nms_top_n by objectness logit, sigmoid,
//       clip_boxes_to_image, remove_small_boxes(min_size), score >= score_thresh, batched_nms by level,
//       first post_nms_top_n
//   RoIHeads.postprocess_detections (models/detection/roi_heads.py): softmax, per-class decode, clip, drop the
//       background column, flatten to (box, class) rows, score > score_thresh, remove_small_boxes(1e-2),
//       batched_nms by class, first detections_per_img
//
// batched_nms is class-separated here ("vanilla" form: one candidate list per (image, class), no coordinate offsets;
// what torchvision itself runs above 4000 box coordinates on CPU / 100000 on CUDA); the coordinate-trick form is
// available through hdy_nms_tiles' class_offset.  All kernels are streaming gathers: HBM-bound, a few bytes per box.
#include "hdy_common.cuh"

namespace hdy {

constexpr int kRcnnThreads = 256;

static inline unsigned rcnn_blocks(long long n, int threads = kRcnnThreads, int max_blocks = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (unsigned)b;
}

// ---------------------------------------------------------------------------------------------- H1
__global__ void __launch_bounds__(kRcnnThreads) rcnn_decode_kernel(const float4* __restrict__ deltas,
                                                                   const float4* __restrict__ boxes, long long R,
                                                                   int C, long long boxes_rows, float wx, float wy,
                                                                   float ww, float wh, float clip,
                                                                   float4* __restrict__ out) {
  const long long n = R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const float4 b = boxes[r % boxes_rows];
    const float4 d = deltas[i];
    const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
    const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
    const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
    float dw = __fdiv_rn(d.z, ww), dh = __fdiv_rn(d.w, wh);
    dw = dw > clip ? clip : dw;  // torch.clamp(max=clip): NaN stays NaN
    dh = dh > clip ? clip : dh;
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
    out[i] = make_float4(__fsub_rn(pcx, hw), __fsub_rn(pcy, hh), __fadd_rn(pcx, hw), __fadd_rn(pcy, hh));
  }
}

// F.softmax(class_logits, -1): exp(x - max) / sum, one thread per row (C is a handful of classes)
__global__ void __launch_bounds__(kRcnnThreads) softmax_rows_kernel(const float* __restrict__ logits, long long R,
                                                                    int C, float* __restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
    const float* x = logits + r * C;
    float m = x[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(x[c], m)));
    for (int c = 0; c < C; ++c) out[r * C + c] = __fdiv_rn(expf(__fsub_rn(x[c], m)), s);
  }
}

__device__ __forceinline__ float4 clip_box(float4 b, float W, float H) {
  // clip_boxes_to_image: x.clamp(min=0, max=width), y.clamp(min=0, max=height)
  b.x = fminf(fmaxf(b.x, 0.f), W);
  b.z = fminf(fmaxf(b.z, 0.f), W);
  b.y = fminf(fmaxf(b.y, 0.f), H);
  b.w = fminf(fmaxf(b.w, 0.f), H);
  return b;
}

__device__ __forceinline__ void append_candidate(int tile, int cap, uint64_t key, const float4& box, float cls,
                                                 uint64_t* cand_keys, float4* cand_boxes, float* cand_cls,
                                                 int32_t* counts, int32_t* status) {
  const int q = atomicAdd(counts + tile, 1);
  if (q < cap) {
    const size_t o = (size_t)tile * cap + q;
    cand_keys[o] = key;
    cand_boxes[o] = box;
    if (cand_cls) cand_cls[o] = cls;
  } else {
    atomicOr(status, HDY_STATUS_OVERFLOW);
  }
}

// ---------------------------------------------------------------------------------------------- H3 front half
// one thread per (row, class >= 1); image of a row by binary search in row_offsets
__global__ void __launch_bounds__(kRcnnThreads) rcnn_filter_compact_kernel(
    const float4* __restrict__ pred_boxes, const float* __restrict__ scores, const int32_t* __restrict__ row_offsets,
    const float* __restrict__ img_wh, int bs, int C, float score_thresh, float min_size, int per_class_tiles, int cap,
    uint64_t* __restrict__ cand_keys, float4* __restrict__ cand_boxes, float* __restrict__ cand_cls,
    int32_t* __restrict__ counts, int32_t* __restrict__ status) {
  const long long R = row_offsets[bs];
  const long long n = R * (C - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C - 1);
    const int c = 1 + (int)(i - r * (C - 1));
    const float s = scores[r * C + c];
    if (!(s > score_thresh)) continue;  // inds = torch.where(scores > self.score_thresh)
    int lo = 0, hi = bs;                // image with row_offsets[img] <= r < row_offsets[img+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_offsets[mid] <= r)
        lo = mid;
      else
        hi = mid;
    }
    const int img = lo;
    const float4 b = clip_box(pred_boxes[r * C + c], img_wh[2 * img], img_wh[2 * img + 1]);
    if (!(__fsub_rn(b.z, b.x) >= min_size && __fsub_rn(b.w, b.y) >= min_size)) continue;  // remove_small_boxes
    const uint32_t idx = (uint32_t)((r - row_offsets[img]) * (C - 1) + (c - 1));  // index in the flattened arrays
    const int tile = per_class_tiles ? img * (C - 1) + (c - 1) : img;
    append_candidate(tile, cap, make_key(s, idx), b, (float)c, cand_keys, cand_boxes, cand_cls, counts, status);
  }
}

// ---------------------------------------------------------------------------------------------- H2
struct RpnLevels {
  int begin[HDY_MAX_LEVELS + 1];  // first anchor of every level, A in [nl]
  int rank0[HDY_MAX_LEVELS + 1];  // first position of every level in the image's top-k concatenation
  int nl;
};

// composite sort key: [segment = image*nl + level : 14 bits][~orderable(logit) : 32 bits][anchor in level : 18 bits]
constexpr int kRpnIdxBits = 18, kRpnSegShift = 50;

__global__ void __launch_bounds__(kRcnnThreads) rpn_level_keys_kernel(const float* __restrict__ objectness, int N,
                                                                      int A, const __grid_constant__ RpnLevels L,
                                                                      uint64_t* __restrict__ keys) {
  const long long n = (long long)N * A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / A), a = (int)(i - (long long)img * A);
    int l = 0;
    for (int k = 1; k < L.nl; ++k)
      if (a >= L.begin[k]) l = k;
    const uint64_t seg = (uint64_t)(img * L.nl + l);
    keys[i] = (seg << kRpnSegShift) | ((uint64_t)(~orderable_u32(objectness[i])) << kRpnIdxBits) |
              (uint64_t)(a - L.begin[l]);
  }
}

// after the sort: the first pre_nms_top_n keys of every segment are that level's top-k, best first
__global__ void __launch_bounds__(kRcnnThreads) rpn_topk_compact_kernel(
    const uint64_t* __restrict__ sorted_keys, const float4* __restrict__ proposals, int N, int A,
    const __grid_constant__ RpnLevels L, int pre_nms_top_n, const float* __restrict__ img_wh, float min_size,
    float score_thresh, int per_level_tiles, int cap, uint64_t* __restrict__ cand_keys,
    float4* __restrict__ cand_boxes, float* __restrict__ cand_cls, int32_t* __restrict__ counts,
    int32_t* __restrict__ status) {
  const int per_img = L.rank0[L.nl];  // sum over levels of min(pre_nms_top_n, level size)
  const long long n = (long long)N * per_img;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / per_img), j = (int)(i - (long long)img * per_img);
    int l = 0;
    for (int k = 1; k < L.nl; ++k)
      if (j >= L.rank0[k]) l = k;
    const int rank = j - L.rank0[l];
    const uint64_t key = sorted_keys[(long long)img * A + L.begin[l] + rank];
    const int a = L.begin[l] + (int)(key & ((1ull << kRpnIdxBits) - 1));
    const float logit = from_orderable_u32(~(uint32_t)(key >> kRpnIdxBits));
    const float prob = sigmoidf_ref(logit);  // objectness_prob = torch.sigmoid(objectness)
    const float4 b = clip_box(proposals[(long long)img * A + a], img_wh[2 * img], img_wh[2 * img + 1]);
    if (!(__fsub_rn(b.z, b.x) >= min_size && __fsub_rn(b.w, b.y) >= min_size)) continue;
    if (!(prob >= score_thresh)) continue;  // keep = torch.where(scores >= self.score_thresh)
    const int tile = per_level_tiles ? img * L.nl + l : img;
    append_candidate(tile, cap, make_key(prob, (uint32_t)j), b, (float)l, cand_keys, cand_boxes, cand_cls, counts,
                     status);
  }
}

// ---------------------------------------------------------------------------------------------- regroup
// Survivors of `group` consecutive class/level tiles -> one candidate list per image (for the final score-ordered
// cut: hdy_nms_tiles with iou_thres >= 1 suppresses nothing, it only sorts and caps).
__global__ void __launch_bounds__(kRcnnThreads) regroup_kept_kernel(
    const int32_t* __restrict__ keep_idx, const float4* __restrict__ keep_box, const float* __restrict__ keep_score,
    const float* __restrict__ keep_cls, const int32_t* __restrict__ keep_counts, int n_tiles, int group, int max_det,
    int cap, uint64_t* __restrict__ cand_keys, float4* __restrict__ cand_boxes, float* __restrict__ cand_cls,
    int32_t* __restrict__ counts, int32_t* __restrict__ status) {
  const long long n = (long long)n_tiles * max_det;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tile = (int)(i / max_det), d = (int)(i - (long long)tile * max_det);
    if (d >= keep_counts[tile]) continue;
    append_candidate(tile / group, cap, make_key(keep_score[i], (uint32_t)keep_idx[i]), keep_box[i],
                     keep_cls ? keep_cls[i] : 0.f, cand_keys, cand_boxes, cand_cls, counts, status);
  }
}

static int fill_rpn_levels(const int32_t* level_sizes_host, int nl, int A, int pre_nms_top_n, RpnLevels* L) {
  HDY_REQUIRE(level_sizes_host && nl >= 1 && nl <= HDY_MAX_LEVELS, "rpn: 1..%d levels", HDY_MAX_LEVELS);
  long long run = 0, rank = 0;
  for (int l = 0; l < nl; ++l) {
    HDY_REQUIRE(level_sizes_host[l] > 0 && level_sizes_host[l] <= (1 << kRpnIdxBits),
                "rpn: a level holds %d anchors (at most %d)", level_sizes_host[l], 1 << kRpnIdxBits);
    L->begin[l] = (int)run;
    L->rank0[l] = (int)rank;
    run += level_sizes_host[l];
    rank += level_sizes_host[l] < pre_nms_top_n ? level_sizes_host[l] : pre_nms_top_n;
  }
  HDY_REQUIRE(run == A, "rpn: level sizes sum to %lld, expected A=%d", run, A);
  for (int l = nl; l <= HDY_MAX_LEVELS; ++l) {
    L->begin[l] = (int)run;
    L->rank0[l] = (int)rank;
  }
  L->nl = nl;
  return HDY_OK;
}

}  // namespace hdy

using namespace hdy;

extern "C" {

int hdy_rcnn_decode(const float* deltas, const float* boxes, int64_t R, int C, int64_t boxes_rows, float wx, float wy,
                    float ww, float wh, float xform_clip, float* out, hdy_stream_t stream) {
  HDY_REQUIRE(R >= 0 && C >= 1 && boxes_rows >= 1, "hdy_rcnn_decode: bad sizes");
  if (R == 0) return HDY_OK;
  HDY_REQUIRE(deltas && boxes && out && (((uintptr_t)deltas | (uintptr_t)boxes | (uintptr_t)out) & 15) == 0,
              "hdy_rcnn_decode: NULL or misaligned pointer");
  rcnn_decode_kernel<<<rcnn_blocks(R * C), kRcnnThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(boxes), R, C, boxes_rows, wx, wy, ww, wh,
      xform_clip, reinterpret_cast<float4*>(out));
  return check_launch("hdy_rcnn_decode");
}

int hdy_softmax_rows(const float* logits, int64_t R, int C, float* out, hdy_stream_t stream) {
  HDY_REQUIRE(R >= 0 && C >= 1, "hdy_softmax_rows: bad sizes");
  if (R == 0) return HDY_OK;
  HDY_REQUIRE(logits && out, "hdy_softmax_rows: NULL pointer");
  softmax_rows_kernel<<<rcnn_blocks(R), kRcnnThreads, 0, (cudaStream_t)stream>>>(logits, R, C, out);
  return check_launch("hdy_softmax_rows");
}

int hdy_rcnn_filter_compact(const float* pred_boxes, const float* scores, const int32_t* row_offsets,
                            const float* img_wh, int bs, int64_t R, int C, float score_thresh, float min_size,
                            int per_class_tiles, int cap, uint64_t* cand_keys, float* cand_boxes, float* cand_cls,
                            int32_t* counts, int32_t* status, hdy_stream_t stream) {
  HDY_REQUIRE(bs >= 0 && R >= 0 && C >= 2 && cap >= 1, "hdy_rcnn_filter_compact: bad sizes");
  if (bs == 0 || R == 0) return HDY_OK;
  HDY_REQUIRE(pred_boxes && scores && row_offsets && img_wh && cand_keys && cand_boxes && counts && status &&
                  (((uintptr_t)pred_boxes | (uintptr_t)cand_boxes) & 15) == 0,
              "hdy_rcnn_filter_compact: NULL or misaligned pointer");
  rcnn_filter_compact_kernel<<<rcnn_blocks(R * (C - 1)), kRcnnThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(pred_boxes), scores, row_offsets, img_wh, bs, C, score_thresh, min_size,
      per_class_tiles, cap, cand_keys, reinterpret_cast<float4*>(cand_boxes), cand_cls, counts, status);
  return check_launch("hdy_rcnn_filter_compact");
}

int hdy_rpn_level_keys(const float* objectness, int N, int A, const int32_t* level_sizes_host, int nl, uint64_t* keys,
                       hdy_stream_t stream) {
  HDY_REQUIRE(N >= 0 && A >= 1, "hdy_rpn_level_keys: bad sizes");
  HDY_REQUIRE((long long)N * nl < (1 << 14), "hdy_rpn_level_keys: at most 16383 (image, level) segments per call");
  RpnLevels L;
  int rc = fill_rpn_levels(level_sizes_host, nl, A, 0, &L);
  if (rc) return rc;
  if (N == 0) return HDY_OK;
  HDY_REQUIRE(objectness && keys, "hdy_rpn_level_keys: NULL pointer");
  rpn_level_keys_kernel<<<rcnn_blocks((long long)N * A), kRcnnThreads, 0, (cudaStream_t)stream>>>(objectness, N, A, L,
                                                                                                 keys);
  return check_launch("hdy_rpn_level_keys");
}

int hdy_rpn_topk_compact(const uint64_t* sorted_keys, const float* proposals, int N, int A,
                         const int32_t* level_sizes_host, int nl, int pre_nms_top_n, const float* img_wh,
                         float min_size, float score_thresh, int per_level_tiles, int cap, uint64_t* cand_keys,
                         float* cand_boxes, float* cand_cls, int32_t* counts, int32_t* status, hdy_stream_t stream) {
  HDY_REQUIRE(N >= 0 && A >= 1 && pre_nms_top_n >= 1 && cap >= 1, "hdy_rpn_topk_compact: bad sizes");
  RpnLevels L;
  int rc = fill_rpn_levels(level_sizes_host, nl, A, pre_nms_top_n, &L);
  if (rc) return rc;
  if (N == 0) return HDY_OK;
  HDY_REQUIRE(sorted_keys && proposals && img_wh && cand_keys && cand_boxes && counts && status &&
                  (((uintptr_t)proposals | (uintptr_t)cand_boxes) & 15) == 0,
              "hdy_rpn_topk_compact: NULL or misaligned pointer");
  rpn_topk_compact_kernel<<<rcnn_blocks((long long)N * L.rank0[nl]), kRcnnThreads, 0, (cudaStream_t)stream>>>(
      sorted_keys, reinterpret_cast<const float4*>(proposals), N, A, L, pre_nms_top_n, img_wh, min_size, score_thresh,
      per_level_tiles, cap, cand_keys, reinterpret_cast<float4*>(cand_boxes), cand_cls, counts, status);
  return check_launch("hdy_rpn_topk_compact");
}

int hdy_regroup_kept(const int32_t* keep_idx, const float* keep_box, const float* keep_score, const float* keep_cls,
                     const int32_t* keep_counts, int n_tiles, int group, int max_det, int cap, uint64_t* cand_keys,
                     float* cand_boxes, float* cand_cls, int32_t* counts, int32_t* status, hdy_stream_t stream) {
  HDY_REQUIRE(n_tiles >= 0 && group >= 1 && max_det >= 1 && cap >= 1, "hdy_regroup_kept: bad sizes");
  if (n_tiles == 0) return HDY_OK;
  HDY_REQUIRE(keep_idx && keep_box && keep_score && keep_counts && cand_keys && cand_boxes && counts && status &&
                  (((uintptr_t)keep_box | (uintptr_t)cand_boxes) & 15) == 0,
              "hdy_regroup_kept: NULL or misaligned pointer");
  regroup_kept_kernel<<<rcnn_blocks((long long)n_tiles * max_det), kRcnnThreads, 0, (cudaStream_t)stream>>>(
      keep_idx, reinterpret_cast<const float4*>(keep_box), keep_score, keep_cls, keep_counts, n_tiles, group, max_det,
      cap, cand_keys, reinterpret_cast<float4*>(cand_boxes), cand_cls, counts, status);
  return check_launch("hdy_regroup_kept");
}

}  // extern "C"
