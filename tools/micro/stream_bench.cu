// Micro-benchmark: how fast can one B200 stream a large fp32 array into SMs through
//   (a) a ring of 1-D bulk-async copies (cp.async.bulk, SASS UBLKCP) guarded by mbarriers, as filter_tma.cu does,
//   (b) plain 128-bit streaming loads.
// Consumers only touch one float per 36-byte row (what the filter's phase A does).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

constexpr int kConsumers = 8;

// split: each chunk is issued as `split` bulk copies of chunk_bytes/split
__global__ void __launch_bounds__(288) ring_kernel(const float* __restrict__ src, long long total_chunks, int chunk_bytes,
                                                   int stages, int split, float thr, unsigned long long* out, int poll) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  float* data = reinterpret_cast<float*>(smem + 256);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int chunk_floats = chunk_bytes / 4;
  const long long my = (total_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (t == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == kConsumers) {
    if (lane == 0) {
      for (long long it = 0; it < my; ++it) {
        const int s = (int)(it % stages);
        if (it >= stages) { const uint32_t par = (uint32_t)(it / stages - 1) & 1u; while (!(poll ? mbar_test_wait(&empty[s], par) : mbar_try_wait(&empty[s], par))) {} }
        const float* g = src + (blockIdx.x + it * gridDim.x) * (long long)chunk_floats;
        mbar_arrive_expect_tx(&full[s], (uint32_t)chunk_bytes);
        const int piece = chunk_bytes / split;
        for (int k = 0; k < split; ++k)
          bulk_copy_g2s((char*)(data + (size_t)s * chunk_floats) + k * piece, (const char*)g + k * piece, piece, &full[s]);
      }
    }
    return;
  }
  unsigned long long cnt = 0;
  const int rows = chunk_floats / 9;
  for (long long it = 0; it < my; ++it) {
    const int s = (int)(it % stages);
    const uint32_t par = (uint32_t)(it / stages) & 1u;
    while (!(poll ? mbar_test_wait(&full[s], par) : mbar_try_wait(&full[s], par))) {}
    const float* sm = data + (size_t)s * chunk_floats;
    for (int r = warp * 32 + lane; r < rows; r += kConsumers * 32) cnt += sm[r * 9 + 4] >= thr;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (cnt) atomicAdd(out, cnt);
}

__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ src, long long n4, float thr,
                                                  unsigned long long* out) {
  unsigned long long cnt = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    cnt += (a.x >= thr) + (b.y >= thr) + (c.z >= thr) + (d.w >= thr);
  }
  for (; i < n4; i += stride) cnt += __ldcs(src + i).x >= thr;
  if (cnt) atomicAdd(out, cnt);
}

// burst: no ring.  Each CTA loads S chunks at once (one lane each or all by lane 0), waits, reads, exits.
__global__ void __launch_bounds__(256) burst_kernel(const float* __restrict__ src, int chunk_bytes, int S, int multi,
                                                    float thr, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  float* data = reinterpret_cast<float*>(smem + 256);
  const int t = threadIdx.x;
  const int chunk_floats = chunk_bytes / 4;
  if (t == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float* g = src + (long long)blockIdx.x * S * chunk_floats;
  if (multi ? (t < S) : (t == 0)) {
    for (int s = multi ? t : 0; s < (multi ? t + 1 : S); ++s) {
      mbar_arrive_expect_tx(&full[s], (uint32_t)chunk_bytes);
      bulk_copy_g2s(data + (size_t)s * chunk_floats, g + (size_t)s * chunk_floats, chunk_bytes, &full[s]);
    }
  }
  unsigned long long cnt = 0;
  const int rows = chunk_floats / 9;
  for (int s = 0; s < S; ++s) {
    while (!mbar_try_wait(&full[s], 0)) {}
    const float* sm = data + (size_t)s * chunk_floats;
    for (int r = t; r < rows; r += 256) cnt += sm[r * 9 + 4] >= thr;
  }
  if (cnt) atomicAdd(out, cnt);
}

int main() {
  const size_t bytes = 600ull << 20;  // > L2
  float* buf; unsigned long long* out;
  cudaMalloc(&buf, 2 * bytes); cudaMalloc(&out, 8); cudaMemset(buf, 0, 2 * bytes); cudaMemset(out, 0, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto report = [&](const char* what, float ms) { printf("%-52s %8.3f ms  %8.1f GB/s\n", what, ms, bytes / ms / 1e6); };
  for (int rep = 0; rep < 2; ++rep) {
    for (int gridmul : {4, 8, 16}) {
      cudaEventRecord(e0);
      for (int i = 0; i < 4; ++i) ldg_kernel<<<sms * gridmul, 256>>>((const float4*)(buf + (i & 1) * (bytes / 4)), bytes / 16, 1.0f, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      char w[128]; snprintf(w, sizeof w, "ldg.128 grid=%dxSM", gridmul); if (rep) report(w, ms / 4);
    }
    for (int chunk : {4608, 18432}) for (int stages : {2, 4, 8}) for (int per_sm : {1, 2}) for (int split : {1}) for (int poll : {0, 1}) {
      const size_t smem = 256 + (size_t)stages * chunk;
      if (smem * per_sm > 200 * 1024) continue;
      cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const long long chunks = bytes / chunk;
      cudaEventRecord(e0);
      for (int i = 0; i < 4; ++i) ring_kernel<<<sms * per_sm, 288, smem>>>(buf + (i & 1) * (bytes / 4), chunks, chunk, stages, split, 1.0f, out, poll);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaError_t err = cudaGetLastError();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      char w[128]; snprintf(w, sizeof w, "bulk ring chunk=%5d stages=%d cta/sm=%d poll=%d %s", chunk, stages, per_sm, poll, err ? cudaGetErrorString(err) : "");
      if (rep) report(w, ms / 4);
    }
  }
  for (int chunk : {4608, 18432}) for (int S : {1, 2, 4, 8}) for (int multi : {0, 1}) {
    const size_t smem = 256 + (size_t)S * chunk;
    if (smem > 200 * 1024) continue;
    cudaFuncSetAttribute(burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long ctas = bytes / ((long long)chunk * S);
    cudaEventRecord(e0);
    for (int i = 0; i < 4; ++i) burst_kernel<<<(unsigned)ctas, 256, smem>>>(buf + (i & 1) * (bytes / 4), chunk, S, multi, 1.0f, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaError_t err = cudaGetLastError();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    char w[128]; snprintf(w, sizeof w, "burst chunk=%5d S=%d multi=%d %s", chunk, S, multi, err ? cudaGetErrorString(err) : "");
    report(w, ms / 4);
  }
  return 0;
}
