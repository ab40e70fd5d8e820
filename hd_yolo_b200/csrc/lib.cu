// Library plumbing: version, error string, level table construction.
#include <stdarg.h>
#include "hdy_common.cuh"

namespace hdy {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HDY_ERR_CUDA;
  }
  return HDY_OK;
}

int build_level_table(const hdy_level_t* lv, int nl, int na, int no, int layout, int rows_per_chunk,
                      LevelTable* out) {
  HDY_REQUIRE(lv != nullptr, "levels is NULL");
  HDY_REQUIRE(nl >= 1 && nl <= HDY_MAX_LEVELS, "nl=%d out of range [1,%d]", nl, HDY_MAX_LEVELS);
  HDY_REQUIRE(na >= 1 && na <= HDY_MAX_ANCHORS, "na=%d out of range [1,%d]", na, HDY_MAX_ANCHORS);
  HDY_REQUIRE(no >= 5, "no=%d < 5", no);
  HDY_REQUIRE(layout == 0 || layout == 1, "layout must be 0 (bs,na,ny,nx,no) or 1 (bs,na*no,ny,nx)");
  memset(out, 0, sizeof(*out));
  long long row_off = 0;
  int chunk = 0;
  for (int l = 0; l < nl; ++l) {
    HDY_REQUIRE(lv[l].logits != nullptr, "levels[%d].logits is NULL", l);
    HDY_REQUIRE(lv[l].ny > 0 && lv[l].nx > 0, "levels[%d] has empty grid", l);
    HDY_REQUIRE(lv[l].dtype == HDY_F32 || lv[l].dtype == HDY_F16, "levels[%d].dtype must be HDY_F32 or HDY_F16", l);
    HDY_REQUIRE(lv[l].dtype == lv[0].dtype, "levels disagree on dtype");
    HDY_REQUIRE(lv[l].dtype == HDY_F32 || layout == 0, "fp16 logits are read in layout 0 ([bs,na,ny,nx,no]) only");
    HDY_REQUIRE(((uintptr_t)lv[l].logits & (lv[l].dtype == HDY_F16 ? 1 : 3)) == 0, "levels[%d].logits misaligned", l);
    LevelDev& d = out->lv[l];
    d.ptr = static_cast<const float*>(lv[l].logits);
    d.ny = lv[l].ny;
    d.nx = lv[l].nx;
    long long rows = (long long)na * lv[l].ny * lv[l].nx;
    HDY_REQUIRE(row_off + rows < (1ll << 31), "too many rows per tile");
    d.rows = (int)rows;
    d.row_offset = (int)row_off;
    d.chunk_begin = chunk;
    d.stride = lv[l].stride;
    for (int a = 0; a < na; ++a) {
      d.aw[a] = lv[l].anchor_w[a];
      d.ah[a] = lv[l].anchor_h[a];
    }
    row_off += rows;
    chunk += (int)((rows + rows_per_chunk - 1) / rows_per_chunk);
  }
  out->nl = nl;
  out->na = na;
  out->no = no;
  out->N = (int)row_off;
  out->chunks_per_tile = chunk;
  out->layout = layout;
  out->dtype = lv[0].dtype;
  return HDY_OK;
}

__global__ void zero_i32_kernel(int32_t* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

}  // namespace hdy

extern "C" {

const char* hdy_version(void) { return "hd_yolo_b200 0.1 (sm_100a)"; }
const char* hdy_last_error(void) { return hdy::g_err; }

int hdy_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

int hdy_zero_i32(int32_t* p, size_t n, hdy_stream_t stream) {
  if (n == 0) return HDY_OK;
  HDY_REQUIRE(p != nullptr, "hdy_zero_i32: NULL pointer");
  hdy::zero_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n);
  return hdy::check_launch("hdy_zero_i32");
}

}  // extern "C"
