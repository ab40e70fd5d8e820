/* ORACLE -- test infrastructure, not product code.
 *
 * Plain-C restatement of the greedy NMS that the reference obtains from the third-party op
 * torchvision.ops.nms (torchvision 0.26, CPU kernel; called at metayolo/models/utils_general.py:342,
 * :507 and metayolo/models/yolo.py:195).  torchvision's source is not on disk in the build
 * container, so this follows its published algorithm:
 *   order = scores.sort(stable, descending)      (ties -> lower index first, NaN sorts first)
 *   area  = (x2-x1)*(y2-y1)                      fp32
 *   for i in order, unless suppressed: keep i; for every later j:
 *       w = max(0, min(x2)-max(x1)); h likewise; inter = w*h
 *       ovr = inter / (area_i + area_j - inter)  fp32
 *       suppressed[j] |= ovr > iou_threshold     (float promoted to double against a double threshold)
 * tests/test_oracle.py pins it against torchvision.ops.nms on random and adversarial inputs.
 * Built by oracle/Makefile into oracle/liboracle_nms.so (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int before(const float* s, int64_t a, int64_t b) { /* a strictly before b in descending order */
  float x = s[a], y = s[b];
  if (isnan(x)) return !isnan(y);
  if (isnan(y)) return 0;
  return x > y;
}

static void merge_sort(int64_t* idx, int64_t* tmp, int64_t n, const float* s) {
  for (int64_t width = 1; width < n; width *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * width) {
      int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
      int64_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) tmp[k++] = before(s, idx[j], idx[i]) ? idx[j++] : idx[i++]; /* stable */
      while (i < mid) tmp[k++] = idx[i++];
      while (j < hi) tmp[k++] = idx[j++];
    }
    memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
  }
}

/* boxes [n,4] xyxy fp32, scores [n] fp32 -> keep [<=n] int64 (score-descending); returns count, -1 on OOM */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold, int64_t* keep) {
  if (n <= 0) return 0;
  int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t* tmp = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  float* area = (float*)malloc((size_t)n * sizeof(float));
  unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
  if (!order || !tmp || !area || !sup) {
    free(order); free(tmp); free(area); free(sup);
    return -1;
  }
  for (int64_t i = 0; i < n; ++i) {
    order[i] = i;
    area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  merge_sort(order, tmp, n, scores);
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (sup[i]) continue;
    keep[nk++] = i;
    float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3], ia = area[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (sup[j]) continue;
      float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
      float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
      float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
      float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
      float w = xx2 - xx1, h = yy2 - yy1;
      w = w > 0.0f ? w : 0.0f;
      h = h > 0.0f ? h : 0.0f;
      float inter = w * h;
      float ovr = inter / (ia + area[j] - inter);
      if ((double)ovr > iou_threshold) sup[j] = 1;
    }
  }
  free(order); free(tmp); free(area); free(sup);
  return nk;
}
