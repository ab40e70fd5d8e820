// Sanity check of a 4-D TMA tile load (cp.async.bulk.tensor.4d) with the tensor map (a) as a __grid_constant__
// kernel parameter and (b) in global memory.  Prints whether the loaded box matches the source (with OOB zero fill).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__constant__ int cBX, cBY, cBC;
static int BX = 24, BY = 24, BC = 32, RANK = 4;
template <bool PARAM>
__global__ void k(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int x0, int y0, int tile, float* out, int rank, int BX, int BY, int BC) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* dst = reinterpret_cast<float*>(raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(raw + BX * BY * BC * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BX * BY * BC * 4) : "memory");
    const CUtensorMap* m = PARAM ? &pmap : gmap;
    if (rank == 4)
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(x0), "r"(y0), "r"(0), "r"(tile), "r"(smem_u32(bar)) : "memory");
    else if (rank == 3)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(x0), "r"(y0), "r"(0), "r"(smem_u32(bar)) : "memory");
    else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(x0), "r"(y0), "r"(smem_u32(bar)) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < BX * BY * BC; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
  if (argc > 4) { RANK = atoi(argv[1]); BX = atoi(argv[2]); BY = atoi(argv[3]); BC = atoi(argv[4]); }
  const int X0 = argc > 5 ? atoi(argv[5]) : 138;
  if (RANK < 3) BC = 1;
  printf("rank %d box %d x %d x %d x0 %d\n", RANK, BX, BY, BC, X0);
  const int mw = 160, mh = 160, nm = 32, bs = 3;
  std::vector<float> h((size_t)bs * nm * mh * mw);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003) * 0.5f;
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, BX * BY * BC * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
  CUtensorMap map;
  const cuuint64_t dims[4] = {mw, mh, nm, bs};
  const cuuint64_t strides[3] = {(cuuint64_t)mw * 4, (cuuint64_t)mw * mh * 4, (cuuint64_t)mw * mh * nm * 4};
  const cuuint32_t box[4] = {(cuuint32_t)BX, (cuuint32_t)BY, (cuuint32_t)BC, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = ((Fn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, RANK, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CUtensorMap* gm; cudaMalloc(&gm, sizeof(map)); cudaMemcpy(gm, &map, sizeof(map), cudaMemcpyHostToDevice);
  const size_t smem = BX * BY * BC * 4 + 64;
  std::vector<float> o(BX * BY * BC);
  for (int variant = 0; variant < 2; ++variant) {
    const int x0 = X0, y0 = 23, tile = RANK == 4 ? 2 : 0;
    if (variant == 0) { cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<false><<<1, 128, smem>>>(map, gm, x0, y0, tile, out, RANK, BX, BY, BC); }
    else { cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<true><<<1, 128, smem>>>(map, gm, x0, y0, tile, out, RANK, BX, BY, BC); }
    e = cudaDeviceSynchronize();
    printf("variant %s: %s\n", variant ? "param" : "global", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int c = 0; c < BC; ++c) for (int y = 0; y < BY; ++y) for (int x = 0; x < BX; ++x) {
      const int gx = x0 + x, gy = y0 + y;
      const float want = (gx < mw && gy < mh) ? h[(((size_t)tile * nm + c) * mh + gy) * mw + gx] : 0.f;
      bad += o[(c * BY + y) * BX + x] != want;
    }
    printf("  mismatches: %ld\n", bad);
  }
  return 0;
}
