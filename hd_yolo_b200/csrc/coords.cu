// C1: scale_coords + clip_coords (metayolo/models/utils_general.py:161-190) and Detect.rescale_outputs
// (metayolo/models/yolo_head.py:465-471) as one in-place kernel over [n, row_len] rows whose first four columns
// are xyxy.  Every step is a separate fp32 operation against an fp32 scalar, in the reference's order:
//   coords[:, [0,2]] -= pad_x; coords[:, [1,3]] -= pad_y; coords[:, :4] /= gain      (:174-176)
//   coords[:, :4] *= scale                                                            (yolo_head.py:469)
//   clamp x to [0, clip_w], y to [0, clip_h]                                          (:184-187)
//   round to nearest even (evaluation.py:109 `scale_coords(...).round()`)
#include "hdy_common.cuh"

namespace hdy {

__global__ void affine_boxes_kernel(float* __restrict__ rows, long long n, int row_len, float pad_x, float pad_y,
                                    float gain, float scale, float clip_w, float clip_h, int flags) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * 4) return;
  const long long i = e >> 2;
  const int c = (int)(e & 3);
  float* p = rows + i * row_len + c;
  float v = *p;
  const bool is_x = (c & 1) == 0;
  if (flags & 1) {
    v = __fsub_rn(v, is_x ? pad_x : pad_y);
    v = __fdiv_rn(v, gain);
  }
  if (flags & 2) v = __fmul_rn(v, scale);
  if (flags & 4) {
    const float hi = is_x ? clip_w : clip_h;
    // Tensor.clamp_(0, hi): min(max(v, 0), hi); NaN propagates
    if (v == v) v = fminf(fmaxf(v, 0.0f), hi);
  }
  if (flags & 8) v = rintf(v);
  *p = v;
}

}  // namespace hdy

extern "C" int hdy_affine_boxes(float* rows, int64_t n, int row_len, float pad_x, float pad_y, float gain, float scale,
                                float clip_w, float clip_h, int flags, hdy_stream_t stream) {
  HDY_REQUIRE(n >= 0 && row_len >= 4 && flags >= 0 && flags < 16, "hdy_affine_boxes: bad arguments");
  if (n == 0 || flags == 0) return HDY_OK;
  HDY_REQUIRE(rows != nullptr, "hdy_affine_boxes: NULL pointer");
  HDY_REQUIRE(!(flags & 1) || gain != 0.0f, "hdy_affine_boxes: gain is 0");
  const long long e = n * 4;
  hdy::affine_boxes_kernel<<<(unsigned)((e + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rows, n, row_len, pad_x, pad_y, gain, scale, clip_w, clip_h, flags);
  return hdy::check_launch("hdy_affine_boxes");
}
