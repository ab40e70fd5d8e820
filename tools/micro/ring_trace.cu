// Trace of the bulk-copy ring: globaltimer stamps of producer issue and consumer ready/release for CTA 0.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int kConsumers = 8;
__global__ void __launch_bounds__(288) ring_kernel(const float* __restrict__ src, int total_chunks, int chunk_bytes,
                                                   int stages, float thr, unsigned long long* out, unsigned long long* trace) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  float* data = reinterpret_cast<float*>(smem + 256);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int chunk_floats = chunk_bytes / 4;
  const int my = (total_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (t == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool tr = blockIdx.x == 0;
  if (warp == kConsumers) {
    if (lane == 0) {
      for (int it = 0; it < my; ++it) {
        const int s = it % stages;
        if (it >= stages) { const uint32_t par = (uint32_t)(it / stages - 1) & 1u; while (!mbar_try_wait(&empty[s], par)) {} }
        const float* g = src + ((long long)blockIdx.x + (long long)it * gridDim.x) * chunk_floats;
        mbar_arrive_expect_tx(&full[s], (uint32_t)chunk_bytes);
        bulk_copy_g2s(data + (size_t)s * chunk_floats, g, chunk_bytes, &full[s]);
        if (tr && it < 64) trace[it * 4 + 0] = gtime();
      }
    }
    return;
  }
  unsigned long long cnt = 0;
  const int rows = chunk_floats / 9;
  for (int it = 0; it < my; ++it) {
    const int s = it % stages;
    const uint32_t par = (uint32_t)(it / stages) & 1u;
    while (!mbar_try_wait(&full[s], par)) {}
    if (tr && t == 0 && it < 64) trace[it * 4 + 1] = gtime();
    const float* sm = data + (size_t)s * chunk_floats;
    for (int r = warp * 32 + lane; r < rows; r += kConsumers * 32) cnt += sm[r * 9 + 4] >= thr;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (tr && t == 0 && it < 64) trace[it * 4 + 2] = gtime();
  }
  if (cnt) atomicAdd(out, cnt);
}
int main() {
  const size_t bytes = 600ull << 20;
  float* buf; unsigned long long *out, *trace;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 8); cudaMalloc(&trace, 64 * 4 * 8);
  cudaMemset(buf, 0, bytes); cudaMemset(out, 0, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int gridmul : {1}) for (int stages : {4, 8}) {
    const int chunk = 4608;
    const size_t smem = 256 + (size_t)stages * chunk;
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(trace, 0, 64 * 4 * 8);
      ring_kernel<<<sms * gridmul, 288, smem>>>(buf, (int)(bytes / chunk), chunk, stages, 1.0f, out, trace);
      cudaDeviceSynchronize();
    }
    unsigned long long h[256]; cudaMemcpy(h, trace, sizeof h, cudaMemcpyDeviceToHost);
    printf("stages=%d  (ns relative to first issue)  it: issue ready release\n", stages);
    for (int it = 0; it < 40; ++it) printf("  %2d: %7lld %7lld %7lld\n", it, (long long)(h[it*4]-h[0]), (long long)(h[it*4+1]-h[0]), (long long)(h[it*4+2]-h[0]));
  }
  return 0;
}
