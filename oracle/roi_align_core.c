/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of torchvision.ops.roi_align (forward, fp32) as the reference reaches it from
 * Detect.multiscale_roi_align (metayolo/models/yolo_head.py:279-299: output (M, M) with M = mask_output_size // 2,
 * spatial_scale = 1 / stride, sampling_ratio = 2, aligned = ROI_ALIGN = False, :15).  torchvision is a third-party
 * dependency of the reference (0.26.0 installed, unpinned, source not on disk); this follows its published CPU
 * algorithm (roi_align_kernel.cpp / roi_align_common.h: pre_calc_for_bilinear_interpolate, then one weighted sum
 * per sample, divided by the sample count) and is cross-checked bit for bit against the installed
 * torchvision.ops.roi_align in tests/test_oracle.py.
 *
 * Compiled with -ffp-contract=off: every product and sum is rounded separately, as in torchvision's generic build.
 */
#include <stdint.h>
#include <math.h>
#include <stdlib.h>

typedef struct {
  int pos1, pos2, pos3, pos4;
  float w1, w2, w3, w4;
} precalc_t;

/* input [N, C, H, W]; rois [K, 5] = (batch index, x1, y1, x2, y2); out [K, C, PH, PW].  Returns 0, or -1 (malloc). */
int oracle_roi_align(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, const float* rois, int64_t K,
                     float spatial_scale, int PH, int PW, int sampling_ratio, int aligned, float* out) {
  (void)N;
  for (int64_t n = 0; n < K; ++n) {
    const float* r = rois + n * 5;
    const int64_t b = (int64_t)r[0];
    const float offset = aligned ? 0.5f : 0.0f;
    const float roi_start_w = r[1] * spatial_scale - offset;
    const float roi_start_h = r[2] * spatial_scale - offset;
    const float roi_end_w = r[3] * spatial_scale - offset;
    const float roi_end_h = r[4] * spatial_scale - offset;
    float roi_width = roi_end_w - roi_start_w;
    float roi_height = roi_end_h - roi_start_h;
    if (!aligned) {
      roi_width = fmaxf(roi_width, 1.0f);
      roi_height = fmaxf(roi_height, 1.0f);
    }
    const float bin_size_h = roi_height / (float)PH;
    const float bin_size_w = roi_width / (float)PW;
    const int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)PH);
    const int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)PW);
    const float count = (float)(grid_h * grid_w > 1 ? grid_h * grid_w : 1);
    const size_t npc = (size_t)(grid_h > 0 ? grid_h : 0) * (size_t)(grid_w > 0 ? grid_w : 0) * PH * PW;
    precalc_t* pc = (precalc_t*)malloc((npc ? npc : 1) * sizeof(precalc_t));
    if (!pc) return -1;
    size_t idx = 0;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw)
        for (int iy = 0; iy < grid_h; ++iy) {
          const float yy = roi_start_h + ph * bin_size_h + (float)(iy + .5f) * bin_size_h / (float)grid_h;
          for (int ix = 0; ix < grid_w; ++ix) {
            const float xx = roi_start_w + pw * bin_size_w + (float)(ix + .5f) * bin_size_w / (float)grid_w;
            float x = xx, y = yy;
            precalc_t p;
            if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
              p.pos1 = p.pos2 = p.pos3 = p.pos4 = 0;
              p.w1 = p.w2 = p.w3 = p.w4 = 0.f;
              pc[idx++] = p;
              continue;
            }
            if (y <= 0) y = 0;
            if (x <= 0) x = 0;
            int y_low = (int)y, x_low = (int)x, y_high, x_high;
            if (y_low >= H - 1) {
              y_high = y_low = (int)H - 1;
              y = (float)y_low;
            } else {
              y_high = y_low + 1;
            }
            if (x_low >= W - 1) {
              x_high = x_low = (int)W - 1;
              x = (float)x_low;
            } else {
              x_high = x_low + 1;
            }
            const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
            p.w1 = hy * hx;
            p.w2 = hy * lx;
            p.w3 = ly * hx;
            p.w4 = ly * lx;
            p.pos1 = y_low * (int)W + x_low;
            p.pos2 = y_low * (int)W + x_high;
            p.pos3 = y_high * (int)W + x_low;
            p.pos4 = y_high * (int)W + x_high;
            pc[idx++] = p;
          }
        }
    for (int64_t c = 0; c < C; ++c) {
      const float* in = input + (b * C + c) * H * W;
      float* o = out + (n * C + c) * PH * PW;
      size_t k = 0;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float v = 0.f;
          for (int iy = 0; iy < grid_h; ++iy)
            for (int ix = 0; ix < grid_w; ++ix) {
              const precalc_t p = pc[k++];
              v += p.w1 * in[p.pos1] + p.w2 * in[p.pos2] + p.w3 * in[p.pos3] + p.w4 * in[p.pos4];
            }
          v /= count;
          o[ph * PW + pw] = v;
        }
    }
    free(pc);
  }
  return 0;
}
