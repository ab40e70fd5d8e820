// Probe: TMA tensor load of a channels-last window with the 128-byte swizzle (debugging aid for roi_align_tc.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c, int x, int y, int n, int bytes, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(su32(sm)),
        "l"(&tm), "r"(c), "r"(x), "r"(y), "r"(n), "r"(su32(&bar))
        : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(su32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
  if (threadIdx.x == 0) out[bytes / 4] = (float)(su32(sm) & 1023);
}
int main(int argc, char** argv) {
  int W = 40, H = 40, C = 128, N = 2;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<float> h((size_t)N * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, 6, 6, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  if (r) return 0;
  int bytes = 32 * 36 * 4;
  float* out; cudaMalloc(&out, bytes + 4);
  int c0 = 32, x0 = 5, y0 = 7, n0 = 1;
  probe<<<1, 128, 16 * 1024>>>(tm, c0, x0, y0, n0, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 0;
  std::vector<float> o(bytes / 4 + 1); cudaMemcpy(o.data(), out, bytes + 4, cudaMemcpyDeviceToHost);
  printf("smem base & 1023 = %d\n", (int)o[bytes / 4]);
  int bad_plain = 0, bad_swz = 0;
  for (int y = 0; y < 6; ++y) for (int x = 0; x < 6; ++x) for (int c = 0; c < 32; ++c) {
    float want = h[(((size_t)n0 * H + y0 + y) * W + x0 + x) * C + c0 + c];
    int row = y * 6 + x;
    if (o[row * 32 + c] != want) ++bad_plain;
    int chunk = (c >> 2) ^ (row & 7);
    if (o[row * 32 + chunk * 4 + (c & 3)] != want) ++bad_swz;
  }
  printf("mismatches: plain layout %d, swizzled (chunk ^ (row & 7)) %d\n", bad_plain, bad_swz);
  return 0;
}
