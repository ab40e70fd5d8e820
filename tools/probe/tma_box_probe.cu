// Probe: which TMA tensor-load boxes work on this GPU for a [N][C][H][W] fp32 tensor (debugging aid for roi_align_tc.cu).
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
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int c, int n, int bytes, float* out, int dyn_off) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned char* dst = sm + dyn_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(su32(dst)),
        "l"(&tm), "r"(x), "r"(y), "r"(c), "r"(n), "r"(su32(&bar))
        : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(su32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(dst)[i];
}
int main(int argc, char** argv) {
  int bx = argc > 1 ? atoi(argv[1]) : 8, by = argc > 2 ? atoi(argv[2]) : 6, bc = argc > 3 ? atoi(argv[3]) : 64;
  int W = argc > 4 ? atoi(argv[4]) : 80, dyn_off = argc > 5 ? atoi(argv[5]) : 0;
  int H = W, C = 64, N = 2;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<float> h((size_t)N * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bc, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %d %d %d W %d off %d x0 %s: encode %d ", bx, by, bc, W, dyn_off, argc > 6 ? argv[6] : "5", (int)r);
  if (r) { printf("\n"); return 0; }
  int bytes = bx * by * bc * 4;
  float* out; cudaMalloc(&out, bytes);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int x0 = argc > 6 ? atoi(argv[6]) : 5, y0 = 7, c0 = 0, n0 = 1;
  probe<<<1, 128, 100 * 1024>>>(tm, x0, y0, c0, n0, bytes, out, dyn_off);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run %s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(bytes / 4); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < bc; ++c) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
      float want = (x0 + x < W && y0 + y < H) ? h[(((size_t)n0 * C + c0 + c) * H + y0 + y) * W + x0 + x] : 0.f;
      if (o[(c * by + y) * bx + x] != want) ++bad;
    }
    printf("mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
