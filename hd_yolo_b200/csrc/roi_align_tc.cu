// Multi-scale RoIAlign on the 5th-generation tensor cores (tcgen05, accumulators in TMEM): the selectable fast form of
// roi_align.cu for callers that accept the north star's 1e-5 relative tolerance instead of torchvision's exact
// operation order.
//
// Reference: Detect.multiscale_roi_align, metayolo/models/yolo_head.py:279-299 (torchvision.ops.roi_align per level,
// sampling_ratio = 2, aligned = False, output 14 x 14, 256 channels).
//
// Why tensor cores here (and not in the mask kernels): per RoI the op writes C*M*M*4 = 200 KB and reads a ~6 x 6 window
// of every channel, so the bound is the HBM write -- but the exact-order kernel needs 77 warp instructions per output
// (ncu: issue slots 73 % busy, DRAM 15 %, profiles/r01_i_roi_align.md).  Bilinear sampling + averaging is linear in the
// window: out[bin][c] = sum_k W[bin][k] * F[k][c] with W = (mean over the S samples of the y weights) x (the same for
// x), a [M*M x 36] matrix per RoI that all 256 channels share.  That is a 196 x 36 x 256 GEMM per RoI.
//
// Arithmetic: 3xTF32.  Both operands are split  v = hi + lo  (hi = cvt.rna.tf32(v), lo = cvt.rna.tf32(v - hi), both
// exactly representable, so the tensor core's own input conversion changes nothing) and D = Alo*Bhi + Ahi*Blo + Ahi*Bhi
// accumulates in fp32: the dropped lo*lo term and the rounding of lo are each <= 2^-22 relative per product.
//
// B200 mapping, one CTA per RoI at a time (persistent grid, 2 CTAs per SM so one CTA's operand build overlaps the
// other's MMAs and epilogue):
//   * sample tables exactly as the exact kernel builds them (roi_align_common.cuh), reduced to per-axis weights
//     wy[M][6], wx[M][6] over the RoI's tap window (<= 6 x 6 feature pixels: nuclei; larger windows are appended to a
//     list and done by the exact kernel afterwards);
//   * A = W [196 (-> 2 x M128) x 40] and, per half of the channels, B = F [128 x 40], both hi and lo, written by the
//     threads straight into the no-swizzle K-major core-matrix layout the shared-memory descriptors describe
//     (8 rows x 16 B core matrices; SBO = 128 B between row groups, LBO = the K-chunk pitch);
//   * one thread issues 3 terms x 2 M tiles x 5 K steps of tcgen05.mma.kind::tf32 (128 x 128 x 8) and commits to an
//     mbarrier;
//   * epilogue: tcgen05.ld 32x32b.x32 (lane = output bin, column = channel) and coalesced streaming stores -- the bins
//     of a channel are contiguous in the output, so a warp writes 128 B per register.
#include <limits.h>
#include "roi_align_common.cuh"
#include "tma_common.cuh"

namespace hdy {

int roi_align_args_ok(int bs, int channels, const float* rois, const float* level_of, int nl, int64_t K, int pooled,
                      int sampling_ratio, const float* out);
int launch_roi_align_exact(const RoiLevels& L, int bs, int channels, const float* rois, const float* level_of,
                           int64_t K, int pooled, int sampling_ratio, int aligned, float* out, const int32_t* only,
                           cudaStream_t st);

constexpr int kTcThreads = 256;
constexpr int kTcSpan = 6;                      // tap window per axis on the tensor-core path
constexpr int kTcK = 40;                        // 36 window pixels, padded to whole K steps of 8
constexpr int kTcKc = kTcK / 4;                 // 16-byte K chunks
constexpr int kTcKSteps = kTcK / 8;
constexpr int kTcRows = 196;                    // A rows held (pooled <= 14); see the over-read note
constexpr int kTcN = 64;                        // channels per MMA (N): one slice of the RoI's channels
constexpr int kTcALbo = kTcRows * 16;           // bytes between the K chunks of A
constexpr int kTcBLbo = kTcN * 16 + 16;         // ... of B; +16: a channel's consecutive k land in consecutive banks
constexpr int kTcABytes = kTcKc * kTcALbo;      // 31 360
constexpr int kTcBBytes = kTcKc * kTcBLbo;      // 10 400
constexpr int kTcCols = 256;                    // TMEM columns: 2 accumulator buffers x 2 M tiles x 64 channels (fp32)
// Over-read: the second M tile's descriptor covers rows 128..255 but only rows < bins are written.  What the MMA reads
// past row 195 of a chunk is the next chunk / the next buffer (A_hi -> A_lo -> the TMA ring, all inside this CTA's shared
// memory): finite or not, it only reaches accumulator rows >= bins, which nobody loads.

// The TMA box: its global origin must be 16-byte aligned (an unaligned x origin is an illegal-instruction fault, probed
// with tools/probe/tma_box_probe.cu), so the load starts at xlo & ~3 and is 12 pixels wide: 3 + 6 window columns, 48 B.
constexpr int kTcRawX = 12;
constexpr int kTcRawBytes = kTcN * kTcSpan * kTcRawX * 4;   // one slice's window, [channel][6][12] fp32: 18 432

struct TcMaps {
  CUtensorMap m[HDY_MAX_LEVELS];   // per level; NCHW: box {12, 6, 64, 1}; channels-last: box {32 ch, 6, 6, 1}, 128B swizzle
};

// Channels-last features ([bs][h][w][C]; torch.channels_last): a window row is 6 x C contiguous floats, so a TMA box
// {32 channels, 6, 6} is 36 requests of 128 bytes instead of 384 of 48, and its origin needs no alignment slack.  With the
// 128-byte swizzle it lands as [k = 36 window pixels][32 channels] rows of 128 bytes, 16-byte chunk j of row k stored at
// chunk j ^ (k & 7) (checked with tools/probe/tma_nhwc_probe.cu); the split pass reads it back with lane = channel
// (conflict-free) and writes the same K-major (hi, lo) operand as the NCHW path, one 16-byte K chunk per store.
// (Feeding the swizzled rows to the MMA directly as an MN-major SWIZZLE_128B operand returned zeros; not pursued.)
constexpr int kTcNhBlk = 40 * 128;              // one 32-channel block of a ring buffer: 36 rows, 1 024-byte aligned pitch
constexpr int kTcNhRing = 2 * kTcNhBlk;         // one ring buffer: a slice's 64 channels
constexpr int kTcNhA0 = (2 * kTcABytes + 1023) / 1024 * 1024;   // swizzled buffers are 1 024-byte aligned

struct TcTables {
  SampleTab ytab[kRoiMaxM * kRoiMaxS];
  SampleTab xtab[kRoiMaxM * kRoiMaxS];
  float wy[kRoiMaxM][8];
  float wx[kRoiMaxM][8];
  unsigned long long mbar;
  unsigned long long ring_bar[2];
  int bnd[2][4];   // window of all taps (ylo, yhi, xlo, xhi), ping-pong over consecutive RoIs
  int nxt[5];      // the next RoI, worked out one RoI ahead by the idle warp: (fits, xlo, ylo, level, image)
  uint32_t tmem_base;
};
constexpr size_t kTcSmemNchw = 2 * kTcABytes + 2 * kTcBBytes + kTcRawBytes + sizeof(TcTables);
constexpr size_t kTcSmemNhwc = kTcNhA0 + 2 * kTcNhRing + 2 * kTcBBytes + sizeof(TcTables);
static_assert(kTcSmemNchw <= 113 * 1024 && kTcSmemNhwc <= 113 * 1024, "two CTAs per SM");
static_assert(60 * 16 <= kTcRawBytes, "over-read of the last A chunk stays inside smem");
static_assert((2 * kTcABytes) % 128 == 0 && kTcRawBytes % 128 == 0, "TMA destinations are 128-byte aligned");

// round to nearest tf32 (10 mantissa bits; ties away from zero, like cvt.rna.tf32.f32) with two integer-pipe
// instructions: the cvt runs on the quarter-rate conversion pipe and was 16 % of the kernel's stall samples
__device__ __forceinline__ float tf32_rna(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

// shared-memory matrix descriptor: K-major, no swizzle, version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}

// instruction descriptor: D fp32 (bits 4-5 = 1), A and B tf32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// MC: the pooled size when it is the reference's 14 (compile-time: the epilogue's 128 stores per warp then take
// immediate offsets -- with a run-time row pitch the address arithmetic was 8 of every 9 instructions of the kernel's
// hottest line), else 0 (run-time M)
template <int MC, bool NHWC>
__global__ void __launch_bounds__(kTcThreads, 2) roi_align_tc_kernel(
    const __grid_constant__ TcMaps maps, const RoiLevels L, int bs, int C, const float* __restrict__ rois, const float* __restrict__ level_of, long long K,
    int M_rt, int S, int aligned, float* __restrict__ out, int32_t* __restrict__ fallback, uint32_t level_mask) {
  const int M = MC ? MC : M_rt;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* const A_hi = tc_smem;
  unsigned char* const A_lo = A_hi + kTcABytes;
  // NCHW: [A_hi | A_lo | TMA window (128-byte aligned) | B_hi | B_lo | tables]
  // NHWC: [A_hi | A_lo | pad to 1 KB | TMA ring 0 | TMA ring 1 | B_hi | B_lo | tables]
  unsigned char* const raw = NHWC ? tc_smem + kTcNhA0 : A_lo + kTcABytes;
  unsigned char* const B_hi = raw + (NHWC ? 2 * kTcNhRing : kTcRawBytes);
  unsigned char* const B_lo = B_hi + kTcBBytes;
  TcTables& T = *reinterpret_cast<TcTables*>(B_lo + kTcBBytes);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int bins = MC ? MC * MC : M * M, n_mt = (bins + 127) >> 7, slices = C / kTcN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T.tmem_base)),
                 "r"(kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t < 8) T.bnd[t >> 2][t & 3] = (t & 1) ? -1 : INT_MAX;
  if (t == 0) {
    mbar_init(reinterpret_cast<uint64_t*>(&T.mbar), 1);
    mbar_init(reinterpret_cast<uint64_t*>(&T.ring_bar[0]), 1);
    mbar_init(reinterpret_cast<uint64_t*>(&T.ring_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the padding columns k = 36..39 of B are never written again: zero (their A weights are zero, but 0 * garbage
  // could be NaN)
  for (int i = t; i < kTcN; i += kTcThreads) {
    *reinterpret_cast<float4*>(B_hi + (kTcKc - 1) * kTcBLbo + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(B_lo + (kTcKc - 1) * kTcBLbo + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = T.tmem_base;
  uint32_t phase = 0, ring_phase = 0;   // (bit r of ring_phase: parity of TMA buffer r's next completion)
  const float inv_s = 1.0f / (float)S;

  // the RoI row (image, box) and its level id are read one RoI ahead: their global-load latency sat at the head of every
  // RoI's dependency chain
  float nr[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, nlv = 0.f;
  if ((long long)blockIdx.x < K) {
#pragma unroll
    for (int i = 0; i < 5; ++i) nr[i] = __ldg(rois + (long long)blockIdx.x * 5 + i);
    nlv = level_of ? __ldg(level_of + blockIdx.x) : 0.f;
  }
  int it = 0;
  uint32_t pre = 0;   // (thread 0) slices of the coming RoI whose loads were started during the previous one
  for (long long n = blockIdx.x; n < K; n += gridDim.x) {
    const float r[5] = {nr[0], nr[1], nr[2], nr[3], nr[4]};
    const float lv_f = nlv;
    if (n + gridDim.x < K) {
#pragma unroll
      for (int i = 0; i < 5; ++i) nr[i] = __ldg(rois + (n + gridDim.x) * 5 + i);
      nlv = level_of ? __ldg(level_of + n + gridDim.x) : 0.f;
    }
    const int lvl = level_of ? (int)lv_f : 0;
    const bool lvl_ok = level_of ? (lv_f == (float)lvl && lvl >= 0 && lvl < L.nl) : true;
    const int b = (int)r[0];
    const bool live = lvl_ok && b >= 0 && b < bs;
    const int H = live ? L.h[lvl] : 1, W = live ? L.w[lvl] : 1;
    int* const bnd = T.bnd[it & 1];   // (reset one RoI ago, after everybody had read it for the last time)
    // ---- sample tables: threads [0, M*S) the y axis, [128, 128 + M*S) the x axis
    if (live && (t & 127) < M * S) {
      const float scale = L.scale[lvl], off = aligned ? 0.5f : 0.0f;
      const bool isx = t >= 128;
      const float s0 = __fsub_rn(__fmul_rn(isx ? r[1] : r[2], scale), off);
      const float e0 = __fsub_rn(__fmul_rn(isx ? r[3] : r[4], scale), off);
      float len = __fsub_rn(e0, s0);
      if (!aligned) len = fmaxf(len, 1.0f);
      const float bin = __fdiv_rn(len, (float)M);
      const int i = t & 127;
      const SampleTab Sa = roi_sample(s0, bin, i / S, i % S, S, isx ? W : H);
      (isx ? T.xtab : T.ytab)[i] = Sa;
      if (Sa.low >= 0) {
        atomicMin(&bnd[isx ? 2 : 0], Sa.low);
        atomicMax(&bnd[isx ? 3 : 1], Sa.high);
      }
    }
    __syncthreads();
    const int ylo = bnd[0], xlo = bnd[2], yhi = bnd[1], xhi = bnd[3];
    const int span_h = yhi - ylo + 1, span_w = xhi - xlo + 1;
    if (t == 0) {   // the other slot, for the next RoI: its readers are all past the barrier above
      int* const o = T.bnd[(it & 1) ^ 1];
      o[0] = INT_MAX;
      o[1] = -1;
      o[2] = INT_MAX;
      o[3] = -1;
    }
    ++it;
    // dead rows (zero-filled by the reference), all-zero weights and windows over 6 x 6 go to the exact kernel
    // ... and so do the levels without a tensor map (row pitch not a multiple of 16 bytes)
    if (!(live && yhi >= 0 && xhi >= 0 && span_h <= kTcSpan && span_w <= kTcSpan && ((level_mask >> lvl) & 1))) {
      if (t == 0) fallback[1 + atomicAdd(fallback, 1)] = (int32_t)n;
      __syncthreads();  // (the reset of the other slot lands before the next RoI's atomics)
      continue;
    }
    // ---- per-axis weights over the window: the mean over the bin's S samples of the bilinear tap weights
    {
      const bool isx = t >= 128;
      const int p = (t & 127) >> 3, j = t & 7;
      if (p < M) {
        const SampleTab* tab = isx ? T.xtab : T.ytab;
        const int lo0 = isx ? xlo : ylo;
        float w = 0.f;
        for (int i = 0; i < S; ++i) {
          const SampleTab Sa = tab[p * S + i];
          if (Sa.low >= 0) {
            if (Sa.low - lo0 == j) w += Sa.h;
            if (Sa.high - lo0 == j) w += Sa.l;
          }
        }
        (isx ? T.wx : T.wy)[p][j] = w * inv_s;
      }
    }
    __syncthreads();
    // ---- B.  NCHW: the window of a 64-channel slice ([64][6][12] fp32, out-of-range zero-filled) comes by ONE TMA tensor
    // load, one slice ahead of the tensor cores -- no LDG, no L1: with per-thread loads the 24-byte rows cost ~7 cache
    // lines per warp instruction and the LSU, shared with the epilogue's stores, set the pace.  Thread t < 252 then moves
    // window pixel k = t % 36 (= jy * 6 + jx) of channels t / 36 + 7 i from the TMA buffer into the (hi, lo) core-matrix
    // layout: its tap validity and both shared-memory offsets are fixed for the RoI.
    // Channels-last: two TMA loads ({32 channels, 6, 6}, swizzled) per slice into a ring of two buffers, two slices
    // ahead; the split streams float4s (see kTcNhBlk).
    constexpr int kKK = kTcSpan * kTcSpan, kCg = kTcThreads / kKK, kEl = (kTcN + kCg - 1) / kCg;   // 36, 7, 10
    const int bk = t % kKK, bc = t / kKK;
    const int bjy = bk / kTcSpan, bjx = bk - bjy * kTcSpan;
    const bool b_on = bc < kCg, b_tap = b_on && bjy < span_h && bjx < span_w;
    const int b_slot = (bk >> 2) * kTcBLbo + bc * 16 + (bk & 3) * 4;
    const int b_raw = ((bc * kTcSpan + bjy) * kTcRawX + (xlo & 3) + bjx) * 4;
    // slice `sl` of the RoI (LV, X, Y, IMG) into TMA buffer sl mod depth
#define HDY_TC_LOAD_AT(sl, LV, X, Y, IMG)                                                                        \
  {                                                                                                              \
    if constexpr (NHWC) {                                                                                        \
      uint64_t* bar = reinterpret_cast<uint64_t*>(&T.ring_bar[(sl) & 1]);                                        \
      unsigned char* dst = raw + ((sl) & 1) * kTcNhRing;                                                         \
      mbar_arrive_expect_tx(bar, 2 * 36 * 128);                                                                  \
      tma_load_4d(dst, &maps.m[LV], (sl) * kTcN, X, Y, IMG, bar);                                                \
      tma_load_4d(dst + kTcNhBlk, &maps.m[LV], (sl) * kTcN + 32, X, Y, IMG, bar);                                \
    } else {                                                                                                     \
      uint64_t* bar = reinterpret_cast<uint64_t*>(&T.ring_bar[0]);                                               \
      mbar_arrive_expect_tx(bar, kTcRawBytes);                                                                   \
      tma_load_4d(raw, &maps.m[LV], (X) & ~3, Y, (sl) * kTcN, IMG, bar);                                         \
    }                                                                                                            \
  }
    // TMA buffer freed: it takes virtual slice v -- slice v of this RoI, or, past its last slice, slice v - slices of the
    // NEXT RoI (whose window the idle warp worked out during the A build), so that a RoI's first windows are already
    // in flight while the previous RoI's last slices are stored.  (Buffer parity carries over when depth divides slices.
    // Channels-last only: with the NCHW boxes -- 384 requests of 48 bytes each -- starting them earlier measured slower.)
#define HDY_TC_LOAD(v)                                                                                           \
  if (t == 0) {                                                                                                  \
    if ((v) < slices) {                                                                                          \
      HDY_TC_LOAD_AT(v, lvl, xlo, ylo, b)                                                                        \
    } else if (NHWC && (v) - slices < kDepth && slices % kDepth == 0 && T.nxt[0]) {                              \
      HDY_TC_LOAD_AT((v) - slices, T.nxt[3], T.nxt[1], T.nxt[2], T.nxt[4])                                       \
      pre |= 1u << ((v) - slices);                                                                               \
    }                                                                                                            \
  }
    constexpr int kDepth = NHWC ? 2 : 1;
    if (t == 0) {   // what the previous RoI did not start already
      for (int j = 0; j < kDepth && j < slices; ++j)
        if (!((pre >> j) & 1)) HDY_TC_LOAD_AT(j, lvl, xlo, ylo, b)
      pre = 0;
    }
    // ---- the next RoI's window, by warp 7 (bins 224..255: no A row, no epilogue work): same samples, same decision
    if (NHWC && warp == 7) {
      int fit2 = 0, x2 = 0, y2 = 0, l2 = 0, b2 = 0;
      if (n + gridDim.x < K) {
        l2 = level_of ? (int)nlv : 0;
        const bool ok2 = level_of ? (nlv == (float)l2 && l2 >= 0 && l2 < L.nl) : true;
        b2 = (int)nr[0];
        if (ok2 && b2 >= 0 && b2 < bs && ((level_mask >> l2) & 1)) {
          const float scale = L.scale[l2], off = aligned ? 0.5f : 0.0f;
          const float sw = __fsub_rn(__fmul_rn(nr[1], scale), off), sh = __fsub_rn(__fmul_rn(nr[2], scale), off);
          const float ew = __fsub_rn(__fmul_rn(nr[3], scale), off), eh = __fsub_rn(__fmul_rn(nr[4], scale), off);
          float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
          if (!aligned) {
            rw = fmaxf(rw, 1.0f);
            rh = fmaxf(rh, 1.0f);
          }
          const float bin_h = __fdiv_rn(rh, (float)M), bin_w = __fdiv_rn(rw, (float)M);
          int ylo2 = INT_MAX, yhi2 = -1, xlo2 = INT_MAX, xhi2 = -1;
          for (int i = lane; i < M * S; i += 32) {
            const SampleTab Y = roi_sample(sh, bin_h, i / S, i % S, S, L.h[l2]);
            const SampleTab X = roi_sample(sw, bin_w, i / S, i % S, S, L.w[l2]);
            if (Y.low >= 0) {
              ylo2 = min(ylo2, Y.low);
              yhi2 = max(yhi2, Y.high);
            }
            if (X.low >= 0) {
              xlo2 = min(xlo2, X.low);
              xhi2 = max(xhi2, X.high);
            }
          }
          ylo2 = __reduce_min_sync(0xffffffffu, ylo2);
          xlo2 = __reduce_min_sync(0xffffffffu, xlo2);
          yhi2 = __reduce_max_sync(0xffffffffu, yhi2);
          xhi2 = __reduce_max_sync(0xffffffffu, xhi2);
          fit2 = yhi2 >= 0 && xhi2 >= 0 && yhi2 - ylo2 + 1 <= kTcSpan && xhi2 - xlo2 + 1 <= kTcSpan;
          x2 = xlo2;
          y2 = ylo2;
        }
      }
      if (lane == 0) {
        T.nxt[0] = fit2;
        T.nxt[1] = x2;
        T.nxt[2] = y2;
        T.nxt[3] = l2;
        T.nxt[4] = b2;
      }
    }
    // ---- A: thread = output bin, k = jy * 6 + jx
    if (t < bins) {
      const int py = t / M, px = t - py * M;
      float wyv[kTcSpan], wxv[kTcSpan];
#pragma unroll
      for (int j = 0; j < kTcSpan; ++j) {
        wyv[j] = T.wy[py][j];
        wxv[j] = T.wx[px][j];
      }
#pragma unroll
      for (int kc = 0; kc < kTcKc; ++kc) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = kc * 4 + e;
          const float v = k < kTcSpan * kTcSpan ? wyv[k / kTcSpan] * wxv[k % kTcSpan] : 0.f;
          hi[e] = tf32_rna(v);
          lo[e] = tf32_rna(v - hi[e]);
        }
        *reinterpret_cast<float4*>(A_hi + kc * kTcALbo + t * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(A_lo + kc * kTcALbo + t * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    // split: registers -> B (hi, lo) in the core-matrix layout
#define HDY_TC_SPLIT(sl)                                                                                 \
  if constexpr (NHWC) {                                                                                  \
    const int rb = (sl) & 1;                                                                             \
    while (!mbar_try_wait(reinterpret_cast<uint64_t*>(&T.ring_bar[rb]), (ring_phase >> rb) & 1)) {       \
    }                                                                                                    \
    ring_phase ^= 1u << rb;                                                                              \
    _Pragma("unroll") for (int j = 0; j < 3; ++j) {                                                      \
      const int idx = t + j * kTcThreads; /* item = (K chunk of 4 window pixels, channel): 9 x 64 */     \
      if (idx < 9 * kTcN) {                                                                              \
        const int c = idx & (kTcN - 1), kc = idx >> 6;                                                   \
        const unsigned char* blk = raw + rb * kTcNhRing + (c >> 5) * kTcNhBlk + (c & 3) * 4;             \
        const int cj = (c & 31) >> 2;                                                                    \
        float v[4];                                                                                      \
        _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                  \
          const int k = kc * 4 + e, jy = k / kTcSpan, jx = k - jy * kTcSpan;                             \
          v[e] = (jy < span_h && jx < span_w)                                                            \
                     ? *reinterpret_cast<const float*>(blk + k * 128 + ((cj ^ (k & 7)) << 4))            \
                     : 0.f;                                                                              \
        }                                                                                                \
        float4 hi, lo;                                                                                   \
        hi.x = tf32_rna(v[0]), hi.y = tf32_rna(v[1]), hi.z = tf32_rna(v[2]), hi.w = tf32_rna(v[3]);      \
        lo.x = tf32_rna(v[0] - hi.x), lo.y = tf32_rna(v[1] - hi.y);                                      \
        lo.z = tf32_rna(v[2] - hi.z), lo.w = tf32_rna(v[3] - hi.w);                                      \
        *reinterpret_cast<float4*>(B_hi + kc * kTcBLbo + c * 16) = hi;                                   \
        *reinterpret_cast<float4*>(B_lo + kc * kTcBLbo + c * 16) = lo;                                   \
      }                                                                                                  \
    }                                                                                                    \
  } else {                                                                                               \
    while (!mbar_try_wait(reinterpret_cast<uint64_t*>(&T.ring_bar[0]), ring_phase & 1)) {                \
    }                                                                                                    \
    ring_phase ^= 1u;                                                                                    \
    if (b_on) {                                                                                          \
      const unsigned char* rp = raw + b_raw;                                                             \
      _Pragma("unroll") for (int i = 0; i < kEl; ++i) {                                                  \
        if (bc + kCg * i < kTcN) {                                                                       \
          const float v = b_tap ? *reinterpret_cast<const float*>(rp + i * (kCg * kTcSpan * kTcRawX * 4)) : 0.f; \
          const float hi = tf32_rna(v), lo = tf32_rna(v - hi);                                           \
          *reinterpret_cast<float*>(B_hi + b_slot + i * (kCg * 16)) = hi;                                \
          *reinterpret_cast<float*>(B_lo + b_slot + i * (kCg * 16)) = lo;                                \
        }                                                                                                \
      }                                                                                                  \
    }                                                                                                    \
  }
    // one thread: 3 terms x M tiles x 5 K steps into accumulator buffer `buf`, then commit to the mbarrier
#define HDY_TC_ISSUE(buf)                                                                                         \
  if (t == 0) {                                                                                                   \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                               \
    const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);    \
    for (int mt = 0; mt < n_mt; ++mt) {                                                                           \
      const uint32_t d = tmem + (uint32_t)((buf) * 2 * kTcN + mt * kTcN);                                         \
      uint32_t acc = 0;                                                                                           \
      _Pragma("unroll") for (int term = 0; term < 3; ++term) { /* the small terms first */                       \
        const uint32_t a0 = (term == 0 ? a_lo : a_hi) + (uint32_t)(mt * 128 * 16);                                \
        const uint32_t b0 = term == 1 ? b_lo : b_hi;                                                              \
        _Pragma("unroll") for (int ks = 0; ks < kTcKSteps; ++ks) {                                                \
          tc_mma(d, tc_desc(a0 + ks * 2 * kTcALbo, kTcALbo, 128), tc_desc(b0 + ks * 2 * kTcBLbo, kTcBLbo, 128),   \
                 kTcIdesc, acc);                                                                                  \
          acc = 1;                                                                                                \
        }                                                                                                         \
      }                                                                                                           \
    }                                                                                                             \
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(             \
                     smem_u32(&T.mbar))                                                                           \
                 : "memory");                                                                                     \
  }
    // Pipeline over the slices: the MMAs of slice sl + 1 (other accumulator buffer) run while the epilogue of slice sl
    // stores, and the loads of slice sl + 2 are in flight behind both.  B is single-buffered: it is rewritten only after
    // the mbarrier said the MMAs reading it are done.
    HDY_TC_SPLIT(0)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // (every warp is past the previous RoI's epilogue: both accumulator buffers are free)
    HDY_TC_ISSUE(0)
    HDY_TC_LOAD(kDepth)   // (TMA buffer 0 was read out before the barrier)
    // one slice: wait for its MMAs; hand slice sl + 1 to the tensor cores; start the load of slice sl + 2; then store
    // slice sl from accumulator buffer P
#define HDY_TC_STEP(sl, P)                                                                                    \
  {                                                                                                                \
    while (!mbar_try_wait(reinterpret_cast<uint64_t*>(&T.mbar), phase)) {                                          \
    }                                                                                                              \
    phase ^= 1;                                                                                                    \
    if ((sl) + 1 < slices) {                                                                                       \
      HDY_TC_SPLIT((sl) + 1)                                                                                       \
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                 \
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                             \
      __syncthreads(); /* every warp is past the epilogue of slice sl - 1, which read the buffer written next */   \
      HDY_TC_ISSUE(1 - (P))                                                                                        \
      HDY_TC_LOAD((sl) + 1 + kDepth)                                                                               \
    }                                                                                                              \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                                \
    /* epilogue: warp -> (M tile, lane quarter); lane = bin, registers = 32 consecutive channels */                \
    const int mt = warp >> 2, q = warp & 3;                                                                        \
    const int bin0 = mt * 128 + q * 32, bin = bin0 + lane;                                                         \
    float* o = out + ((size_t)n * C + (size_t)(sl) * kTcN) * bins;                                                 \
    if (mt < n_mt && bin0 < bins) {                                                                                \
      uint32_t v0[32], v1[32];                                                                                     \
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((P) * 2 * kTcN + mt * kTcN);              \
      tc_ld32(ta, v0);                                                                                             \
      tc_ld32(ta + 32, v1);                                                                                        \
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");                                                 \
      if (bin < bins) {                                                                                            \
        _Pragma("unroll") for (int j = 0; j < 32; ++j) __stcs(o + (size_t)j * bins + bin, __uint_as_float(v0[j])); \
        _Pragma("unroll") for (int j = 0; j < 32; ++j)                                                             \
            __stcs(o + (size_t)(32 + j) * bins + bin, __uint_as_float(v1[j]));                                     \
      }                                                                                                            \
    }                                                                                                              \
  }
    for (int sl = 0; sl < slices; sl += 2) {
      HDY_TC_STEP(sl, 0)
      if (sl + 1 < slices) HDY_TC_STEP(sl + 1, 1)
    }
#undef HDY_TC_STEP
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#undef HDY_TC_SPLIT
#undef HDY_TC_ISSUE
#undef HDY_TC_LOAD
#undef HDY_TC_LOAD_AT
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcCols) : "memory");
}

}  // namespace hdy

extern "C" int hdy_multiscale_roi_align_tf32x3(const hdy_feature_level_t* levels_host, int nl, int bs, int channels,
                                               int channels_last, const float* rois, const float* level_of, int64_t K,
                                               int pooled, int sampling_ratio, int aligned, float* out,
                                               int32_t* fallback, hdy_stream_t stream) {
  using namespace hdy;
  HDY_REQUIRE(levels_host != nullptr, "roi_align: levels is NULL");
  HDY_REQUIRE(nl >= 1 && nl <= HDY_MAX_LEVELS, "roi_align: nl=%d out of range [1,%d]", nl, HDY_MAX_LEVELS);
  int rc = roi_align_args_ok(bs, channels, rois, level_of, nl, K, pooled, sampling_ratio, out);
  if (rc || K == 0) return rc;
  RoiLevels L;
  rc = roi_levels_from_host(levels_host, nl, &L);
  if (rc) return rc;
  L.nhwc = channels_last ? 1 : 0;
  HDY_REQUIRE(channels % kTcN == 0, "roi_align (tf32x3): channels=%d must be a multiple of %d", channels, kTcN);
  HDY_REQUIRE(fallback != nullptr, "roi_align (tf32x3): fallback scratch is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (pooled * pooled > kTcRows)  // (the A operand holds 196 rows: pooled <= 14; 15 and 16 take the exact kernel)
    return launch_roi_align_exact(L, bs, channels, rois, level_of, K, pooled, sampling_ratio, aligned, out, nullptr, st);
  // One tensor map per level; TMA wants the base and every stride 16-byte granular.
  //   NCHW  [bs][C][H][W]: dims {W, H, C, bs}, box {12, 6, 64, 1}
  //   NHWC  [bs][H][W][C]: dims {C, W, H, bs}, box {32, 6, 6, 1}, SWIZZLE_128B (32 channels = one 128-byte row)
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  uint32_t level_mask = 0;
  EncodeTiledFn enc = tensor_map_encoder();
  for (int i = 0; enc && i < nl; ++i) {
    const cuuint64_t w = (cuuint64_t)L.w[i], h = (cuuint64_t)L.h[i], c = (cuuint64_t)channels, n = (cuuint64_t)bs;
    if (((uintptr_t)L.data[i] & 15) != 0) continue;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (L.nhwc) {
      const cuuint64_t dims[4] = {c, w, h, n};
      const cuuint64_t strides[3] = {c * 4, w * c * 4, h * w * c * 4};
      const cuuint32_t box[4] = {32, kTcSpan, kTcSpan, 1};
      r = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(L.data[i]), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      if ((w * 4 & 15) != 0) continue;
      const cuuint64_t dims[4] = {w, h, c, n};
      const cuuint64_t strides[3] = {w * 4, w * h * 4, w * h * c * 4};
      const cuuint32_t box[4] = {kTcRawX, kTcSpan, kTcN, 1};
      r = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(L.data[i]), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r == CUDA_SUCCESS) level_mask |= 1u << i;
  }
  if (!level_mask)
    return launch_roi_align_exact(L, bs, channels, rois, level_of, K, pooled, sampling_ratio, aligned, out, nullptr, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static bool attr_set[64] = {};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(roi_align_tc_kernel<14, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTcSmemNchw);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(roi_align_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kTcSmemNchw);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(roi_align_tc_kernel<14, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kTcSmemNhwc);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(roi_align_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kTcSmemNhwc);
    if (e != cudaSuccess) {
      set_error("roi_align (tf32x3) setup: %s", cudaGetErrorString(e));
      return HDY_ERR_CUDA;
    }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaError_t e = cudaMemsetAsync(fallback, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return HDY_ERR_CUDA;
  }
  const unsigned grid = (unsigned)(K < 2ll * sms ? K : 2ll * sms);
#define HDY_TC_LAUNCH(MC, NH, SMEM)                                                                                 \
  roi_align_tc_kernel<MC, NH><<<grid, kTcThreads, SMEM, st>>>(maps, L, bs, channels, rois, level_of, (long long)K,  \
                                                              pooled, sampling_ratio, aligned, out, fallback, level_mask)
  if (L.nhwc) {
    if (pooled == 14) HDY_TC_LAUNCH(14, true, kTcSmemNhwc); else HDY_TC_LAUNCH(0, true, kTcSmemNhwc);
  } else {
    if (pooled == 14) HDY_TC_LAUNCH(14, false, kTcSmemNchw); else HDY_TC_LAUNCH(0, false, kTcSmemNchw);
  }
#undef HDY_TC_LAUNCH
  rc = check_launch("hdy_multiscale_roi_align_tf32x3");
  if (rc) return rc;
  // the RoIs the tensor-core path left (windows over 6 x 6 taps, dead rows): exact kernel over the list
  return launch_roi_align_exact(L, bs, channels, rois, level_of, K, pooled, sampling_ratio, aligned, out, fallback, st);
}
