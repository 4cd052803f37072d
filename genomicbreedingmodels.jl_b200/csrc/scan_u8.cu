// Compact-dosage marker scan ("compact integer dosage encodings,
// 1 B/genotype => 8x fewer bytes, fused with the scan").
//
// Storage: one byte per genotype, code c in [0, 240], allele frequency a = c / 240 (240 is
// divisible by 1..6, 8, 10, 12, 15, 16 so every ploidy level of those ploidies is a code).
// A Float64 matrix is packed only if EVERY element satisfies fl(code/240) == a bit-exactly
// (pack kernel / host packer); otherwise the Float64 path is used.  The doctests' data
// (round.(af .* ploidy) ./ ploidy, /root/reference/src/gwas.jl:43-45) and the BASELINE
// dosage configurations pack.
//
// Kernel: same persistent producer/consumer structure as scan.cu.  The byte matrix is
// addressed through a UINT64 tensor map (8 codes per element), a stage is 1024 rows x C
// markers of codes (one TMA box) plus the matching slice of the side vectors (2048 rows with
// eight consumer warps).  A consumer
// lane owns 8 consecutive rows: one LDS.64 per marker brings its 8 codes; sum(c) and sum(c^2)
// are integer dot products (dp4a); the M dots sum(c*q) run in FP64 with the codes converted
// by the 2^52 trick (one DADD, no I2F).  Per-marker records are identical in meaning to the
// Float64 kernel's: [mean, SS, dots.. (, min nonzero)], so the finalisation kernel is shared.
// sum(c), sum(c^2) are exact integers: SS of a constant column is exactly 0.
//
// Algorithmic bytes: n per marker.
#include <math.h>

#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "reduce.cuh"

namespace gbm {

constexpr int kU8ConsumerWarps = 4;
constexpr int kU8Rows = kU8ConsumerWarps * 32 * 8;  // rows per stage: every consumer lane owns 8 rows
constexpr int kU8ConsumerThreads = kU8ConsumerWarps * 32;
constexpr int kU8Threads = kU8ConsumerThreads;   // no producer warp: lane 0 of warp 0 refills the ring
constexpr int kU8CtasPerSm = 2;                  // 4 warps x 2 CTAs = 2 warps per scheduler, <= 255 registers
constexpr int kU8SmemBudget = 100 * 1024;
constexpr double kLevels = 240.0;

template <int C, int M, bool MINNZ>
struct U8Cfg {
  static constexpr int NSUM = 2 + M;
  static constexpr int NV = C * NSUM;
  static constexpr int NVP = ((NV + 31) / 32) * 32;
  static constexpr int NS = NSUM + (MINNZ ? 1 : 0);
  static constexpr int A_BYTES = kU8Rows * C;          // bytes of codes
  static constexpr int Q_BYTES = kU8Rows * M * 8;
  static constexpr int STAGE_BYTES = A_BYTES + Q_BYTES;
  static constexpr int STAGES_RAW = kU8SmemBudget / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int RED_BYTES = 2 * kU8ConsumerWarps * NVP * 8 + 2 * kU8ConsumerWarps * C * 8 + 2 * C * 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RED_BYTES + 2 * STAGES * 8 + 128;
};

struct U8Params {
  int64_t n, p;
  int num_tiles, chunks;
  double inv_n;
  double* rec;
};

// byte k (0..3) of a 32-bit word as a double, two ways (the kernel mixes them to balance the
// conversion pipe against the FP64 pipe; U8_I2F_BYTES of every 8 codes take the first route):
//  - bfe + cvt: ptxas fuses the pair into ONE `I2F.F64.U8 Rd, Ra.Bk`
//  - PRMT into the low word of 2^52 (exactly representable up to 2^52 + 255) and one DADD
template <int K>
__device__ __forceinline__ double byte_to_double_i2f(uint32_t x) {
  uint32_t t;
  double d;
  asm("bfe.u32 %0, %1, %2, 8;" : "=r"(t) : "r"(x), "n"(8 * K));
  asm("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(t));
  return d;
}
template <int K>
__device__ __forceinline__ double byte_to_double_magic(uint32_t x) {
  const uint32_t code = __byte_perm(x, 0, 0x4440 | K);
  return __hiloint2double(0x43300000, static_cast<int>(code)) - 4503599627370496.0;
}
template <int K, int NI>  // K = 0..7: byte of the 64-bit word (lo, hi); NI bytes go through I2F
__device__ __forceinline__ double code_as_double(uint32_t lo, uint32_t hi) {
  const uint32_t x = K < 4 ? lo : hi;
  if constexpr ((K % 8) * NI / 8 != ((K % 8) + 1) * NI / 8)  // spreads the NI I2F bytes evenly over the 8
    return byte_to_double_i2f<(K & 3)>(x);
  else
    return byte_to_double_magic<(K & 3)>(x);
}
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// two codes in 16-bit lanes -> f(v) = (0x200 - v) & 0x1FF per lane (no borrow between lanes: v <= 255 < 0x200)
__device__ __forceinline__ uint32_t nzkey(uint32_t lanes) { return (0x02000200u - lanes) & 0x01FF01FFu; }

template <int K, int NI, int M_>
struct DotBytes {
  static __device__ __forceinline__ void run(uint32_t lo, uint32_t hi, const double (&q)[M_ > 0 ? M_ : 1][8],
                                             double* dots) {
    const double cd = code_as_double<K, NI>(lo, hi);
#pragma unroll
    for (int m = 0; m < M_; ++m) dots[m] = fma(cd, q[m][K], dots[m]);
    if constexpr (K < 7) DotBytes<K + 1, NI, M_>::run(lo, hi, q, dots);
  }
};

template <int C, int M, bool MINNZ, int NI>
__global__ void __launch_bounds__(kU8Threads, kU8CtasPerSm)
    scan_sums_u8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ,
                        const U8Params prm) {
  using Cfg = U8Cfg<C, M, MINNZ>;
  constexpr int NSUM = Cfg::NSUM, NV = Cfg::NV, NVP = Cfg::NVP, NS = Cfg::NS, STAGES = Cfg::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  double* red = reinterpret_cast<double*>(smem + STAGES * Cfg::STAGE_BYTES);
  double* redmin = red + 2 * kU8ConsumerWarps * NVP;
  double* shift_s = redmin + 2 * kU8ConsumerWarps * C;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(shift_s + 2 * C);
  uint64_t* empty_bar = full_bar + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kU8ConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // Flat iteration space of this CTA: (tile, chunk) pairs, tile = blockIdx.x + k * gridDim.x.
  const int my_tiles = (prm.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int total_iters = my_tiles * prm.chunks;
  auto issue = [&](int j) {  // called by thread 0 only: TMA loads of flat iteration j into its ring slot
    const int tile = static_cast<int>(blockIdx.x) + (j / prm.chunks) * static_cast<int>(gridDim.x);
    const int chunk = j % prm.chunks;
    const int st = j % STAGES;
    uint8_t* dst = ring + st * Cfg::STAGE_BYTES;
    mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
    tma_load_2d(dst, &tmA, chunk * (kU8Rows / 8), tile * C, &full_bar[st], kEvictFirst);
    if (M > 0) {
#pragma unroll
      for (int b = 0; b < kU8Rows / 256; ++b)  // Q boxes are 256 rows x M, laid out [b][m][256]
        tma_load_2d(dst + Cfg::A_BYTES + b * 256 * M * 8, &tmQ, chunk * kU8Rows + b * 256, 0, &full_bar[st],
                    kEvictLast);
    }
  };
  if (tid == 0) {
    prefetch_tensormap(&tmA);
    if (M > 0) prefetch_tensormap(&tmQ);
    for (int j = 0; j < STAGES && j < total_iters; ++j) issue(j);
  }

  int stage = 0;
  uint32_t phase = 0;
  int parity = 0;
  int it = 0;  // flat iteration index
  const int r = 8 * tid;  // this lane's 8 rows inside a 1024-row chunk
  for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, parity ^= 1) {
    uint32_t s1[C], s2[C];
    uint32_t mn[MINNZ ? C : 1];
    double dots[M > 0 ? C * M : 1];
#pragma unroll
    for (int c = 0; c < C; ++c) s1[c] = s2[c] = 0;
#pragma unroll
    for (int i = 0; i < C * M; ++i) dots[i] = 0.0;
    if (MINNZ) {
#pragma unroll
      for (int c = 0; c < C; ++c) mn[c] = 0u;  // two 16-bit lanes of max f(code); 0 = no non-zero code seen
    }

    for (int chunk = 0; chunk < prm.chunks; ++chunk) {
      mbar_wait(&full_bar[stage], phase);
      const uint64_t* sA = reinterpret_cast<const uint64_t*>(ring + stage * Cfg::STAGE_BYTES);
      const double* sQ = reinterpret_cast<const double*>(ring + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES);
      if (chunk == 0 && tid < C)
        shift_s[parity * C + tid] = static_cast<double>(static_cast<uint32_t>(sA[tid * (kU8Rows / 8)] & 0xFFull));
      const int64_t row0 = static_cast<int64_t>(chunk) * kU8Rows + r;
      const int valid = static_cast<int>(prm.n - row0 < 8 ? (prm.n - row0 < 0 ? 0 : prm.n - row0) : 8);
      // byte mask for the rows of this lane that exist (TMA zero-fills whole out-of-range
      // words, the tail bytes of the last word are zero in our own buffer; the mask keeps the
      // code independent of that)
      const uint64_t bmask = valid >= 8 ? ~0ull : ((1ull << (8 * valid)) - 1ull);
      double q[M > 0 ? M : 1][8];
      if (M > 0) {
        const int qb = r >> 8, qr = r & 255;  // box index and row inside the 256-row box
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const double* src = sQ + (qb * M + m) * 256 + qr;
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const double2 t2 = *reinterpret_cast<const double2*>(src + k);
            q[m][k] = t2.x;
            q[m][k + 1] = t2.y;
          }
        }
      }
      if (valid > 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const uint64_t w = sA[c * (kU8Rows / 8) + tid] & bmask;
          const uint32_t lo = static_cast<uint32_t>(w), hi = static_cast<uint32_t>(w >> 32);
          s1[c] = __dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, s1[c]));
          s2[c] = __dp4a(lo, lo, __dp4a(hi, hi, s2[c]));
          if (M > 0) DotBytes<0, NI, M>::run(lo, hi, q, &dots[c * M]);
          if (MINNZ) {
            // smallest non-zero code: bytes go to 16-bit lanes (PRMT), f(v) = (0x200 - v) & 0x1FF maps 0 -> 0 and
            // 1..240 -> 0x1FF..0x110 (decreasing), so a running packed maximum (VIMNMX.U16x2, native on sm_100a)
            // of f is the minimum over the non-zero codes: 2 instructions per genotype instead of ~4
            mn[c] = max_u16x2(mn[c], max_u16x2(max_u16x2(nzkey(__byte_perm(lo, 0, 0x4240)), nzkey(__byte_perm(lo, 0, 0x4341))),
                                               max_u16x2(nzkey(__byte_perm(hi, 0, 0x4240)), nzkey(__byte_perm(hi, 0, 0x4341)))));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (tid == 0 && it >= 1) {
        // refill one slot behind: the slot of iteration it-1 has been released by every warp by
        // now (or will be within a few cycles), so this wait does not hold warp 0 back
        const int j = it - 1 + STAGES;
        if (j < total_iters) {
          const int pst = (it - 1) % STAGES;
          mbar_wait(&empty_bar[pst], static_cast<uint32_t>(((it - 1) / STAGES) & 1));
          issue(j);
        }
      }
      ++it;
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }

    // epilogue.  Integer sums: hardware warp reduction (REDUX), sum(c^2) in two 16-bit halves so
    // the 32-lane total cannot overflow 32 bits whatever n is.  Dots: halving shuffle tree.
    {
      double* dst = red + (parity * kU8ConsumerWarps + warp) * NVP;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t t1 = __reduce_add_sync(0xffffffffu, s1[c]);
        const uint32_t t2l = __reduce_add_sync(0xffffffffu, s2[c] & 0xFFFFu);
        const uint32_t t2h = __reduce_add_sync(0xffffffffu, s2[c] >> 16);
        if (lane == c) {
          dst[c * NSUM + 0] = static_cast<double>(t1);
          dst[c * NSUM + 1] = static_cast<double>(t2l) + 65536.0 * static_cast<double>(t2h);
        }
      }
      if constexpr (M > 0) {
        constexpr int ND = C * M, NDP = ((ND + 31) / 32) * 32;
        double v[NDP];
#pragma unroll
        for (int i = 0; i < NDP; ++i) v[i] = i < ND ? dots[i < ND ? i : 0] : 0.0;  // ND <= NDP
        HalvingStep<NDP / 2, 16, NDP>::run(v, lane);
        const int base = halving_base<NDP>(lane);
#pragma unroll
        for (int i = 0; i < NDP / 32; ++i) {
          const int idx = base + i;
          if (idx < ND) dst[(idx / M) * NSUM + 2 + (idx % M)] = v[i];
        }
      }
    }
    if (MINNZ) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        uint32_t x = mn[c];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x = max_u16x2(x, __shfl_xor_sync(0xffffffffu, x, o));
        x = max(x & 0xFFFFu, x >> 16);                 // max f over the warp's rows
        x = x == 0u ? 255u : 0x200u - x;               // back to the code; 255 = none
        if (lane == 0) redmin[(parity * kU8ConsumerWarps + warp) * C + c] = static_cast<double>(x);
      }
    }
    named_bar_sync(1, kU8ConsumerThreads);
    if (tid < NV) {
      const int c = tid / NSUM, kk = tid - c * NSUM;
      const double* rp = red + parity * kU8ConsumerWarps * NVP;
      double tot = 0.0, S1 = 0.0, S2 = 0.0;
#pragma unroll
      for (int w = 0; w < kU8ConsumerWarps; ++w) {
        tot += rp[w * NVP + tid];
        S1 += rp[w * NVP + c * NSUM];
        S2 += rp[w * NVP + c * NSUM + 1];
      }
      double out;
      if (kk <= 1) {
        // shift by the first code in exact integer arithmetic: constant columns give SS == 0
        const long long c0 = static_cast<long long>(shift_s[parity * C + c]);
        const long long n = prm.n;
        const long long i1 = static_cast<long long>(S1) - n * c0;
        const long long i2 = static_cast<long long>(S2) - 2 * c0 * static_cast<long long>(S1) + n * c0 * c0;
        if (kk == 0) {
          out = (static_cast<double>(c0) + static_cast<double>(i1) * prm.inv_n) / kLevels;
        } else {
          const double d1 = static_cast<double>(i1);
          out = fmax(static_cast<double>(i2) - d1 * d1 * prm.inv_n, 0.0) / (kLevels * kLevels);
        }
      } else {
        out = tot / kLevels;
      }
      const int64_t col = static_cast<int64_t>(tile) * C + c;
      if (col < prm.p) prm.rec[col * NS + kk] = out;
    }
    if (MINNZ && tid >= kU8ConsumerThreads - C) {
      const int c = tid - (kU8ConsumerThreads - C);
      const double* mp = redmin + parity * kU8ConsumerWarps * C;
      double x = mp[c];
#pragma unroll
      for (int w = 1; w < kU8ConsumerWarps; ++w) x = fmin(x, mp[w * C + c]);
      const int64_t col = static_cast<int64_t>(tile) * C + c;
      if (col < prm.p) prm.rec[col * NS + NSUM] = (x >= 255.0) ? INFINITY : x / kLevels;
    }
  }
}

template <int C, int M, bool MINNZ, int NI = 4>
static void launch_u8_cfg(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, const double* Q, int64_t ldq,
                          double* rec, int sm_count, cudaStream_t stream) {
  using Cfg = U8Cfg<C, M, MINNZ>;
  static_assert(Cfg::STAGES >= 2, "ring too shallow");
  alignas(64) CUtensorMap tmA, tmQ;
  make_tensor_map_2d_u64(&tmA, A8, static_cast<uint64_t>((n + 7) / 8), static_cast<uint64_t>(p),
                         static_cast<uint64_t>(ld8), kU8Rows / 8, C);
  if (M > 0)
    make_tensor_map_2d_f64(&tmQ, Q, static_cast<uint64_t>(n), static_cast<uint64_t>(M), static_cast<uint64_t>(ldq),
                           256, M);
  else
    tmQ = tmA;
  U8Params prm;
  prm.n = n;
  prm.p = p;
  prm.num_tiles = static_cast<int>((p + C - 1) / C);
  prm.chunks = static_cast<int>((n + kU8Rows - 1) / kU8Rows);
  prm.inv_n = 1.0 / static_cast<double>(n);
  prm.rec = rec;
  auto kern = scan_sums_u8_kernel<C, M, MINNZ, NI>;
  GBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int max_grid = sm_count * kU8CtasPerSm;
  const int grid = prm.num_tiles < max_grid ? prm.num_tiles : max_grid;
  kern<<<grid, kU8Threads, Cfg::SMEM_BYTES, stream>>>(tmA, tmQ, prm);
  GBM_CUDA(cudaGetLastError());
}

// Mp: the padded side-vector count the Float64 dispatcher also uses (0,1,2,4,6,10,14)
void launch_scan_sums_u8(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, const double* Q, int Mp, int64_t ldq,
                         double* rec, int sm_count, cudaStream_t stream) {
  if (p <= 0 || n <= 0) return;
  switch (Mp) {
    case 0: launch_u8_cfg<16, 0, true>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream); break;
    case 1: launch_u8_cfg<16, 1, false>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream); break;
    case 2: {
      // tuning knob (bytes of every 8 converted by I2F.F64.U8 instead of PRMT + DADD)
      static const int ni = [] { const char* e = getenv("GBM_U8_I2F"); return e ? atoi(e) : 4; }();
      if (ni <= 0) launch_u8_cfg<16, 2, false, 0>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream);
      else if (ni >= 8) launch_u8_cfg<16, 2, false, 8>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream);
      else if (ni == 2) launch_u8_cfg<16, 2, false, 2>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream);
      else if (ni == 6) launch_u8_cfg<16, 2, false, 6>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream);
      else launch_u8_cfg<16, 2, false, 4>(A8, n, p, ld8, Q, ldq, rec, sm_count, stream);
      break;
    }
    // more side vectors (multi-trait): the caller decodes blocks and uses the Float64 kernel
    default: GBM_THROW(1, "scan(u8): unsupported side-vector count");
  }
}

// ------------------------------------------------------------------------------------
// pack (Float64 -> codes, with the exactness check) and decode (codes -> Float64)
// ------------------------------------------------------------------------------------
// Persistent CTAs walk the columns; within a column one thread per row: 8-byte loads and 1-byte stores, both
// fully coalesced (a warp reads 256 contiguous bytes and writes 32), four rows in flight per thread.  The earlier layout (one thread = 8 consecutive rows,
// one 64-bit store) read with a 64-byte lane stride and reached 2.4 TB/s; this one is HBM-bound.
__global__ void __launch_bounds__(256)
    pack_u8_kernel(const double* __restrict__ A, int64_t n, int64_t p, int64_t lda, uint8_t* __restrict__ out,
                   int64_t ld8, unsigned long long* __restrict__ inexact) {
  __shared__ double lut[241];
  for (int i = threadIdx.x; i < 241; i += blockDim.x) lut[i] = static_cast<double>(i) / kLevels;
  __syncthreads();
  unsigned bad = 0;
  auto encode = [&](double a) -> uint8_t {
    const double s = rint(a * kLevels);
    const bool in_range = s >= 0.0 && s <= 240.0;
    const int code = in_range ? static_cast<int>(s) : 0;
    if (!in_range || lut[code] != a) ++bad;
    return static_cast<uint8_t>(code);
  };
  for (int64_t j = blockIdx.x; j < p; j += gridDim.x) {  // persistent CTAs, one column at a time
    const double* col = A + j * lda;
    uint8_t* dst = out + j * ld8;
    int64_t i = threadIdx.x;
    for (; i + 3 * 256 < n; i += 4 * 256) {
      const double a0 = col[i], a1 = col[i + 256], a2 = col[i + 512], a3 = col[i + 768];
      dst[i] = encode(a0);
      dst[i + 256] = encode(a1);
      dst[i + 512] = encode(a2);
      dst[i + 768] = encode(a3);
    }
    for (; i < ld8; i += 256) dst[i] = i < n ? encode(col[i]) : static_cast<uint8_t>(0);
  }
  if (bad) atomicAdd(inexact, static_cast<unsigned long long>(bad));
}

void launch_pack_u8(const double* A, int64_t n, int64_t p, int64_t lda, uint8_t* out, int64_t ld8,
                    unsigned long long* inexact, cudaStream_t stream) {
  if (n <= 0 || p <= 0) return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p, static_cast<int64_t>(sms) * 8));
  pack_u8_kernel<<<grid, 256, 0, stream>>>(A, n, p, lda, out, ld8, inexact);
  GBM_CUDA(cudaGetLastError());
}

// persistent CTAs walk the columns (one tiny CTA per column slice spent its time on launch overhead)
__global__ void __launch_bounds__(256)
    decode_u8_kernel(const uint8_t* __restrict__ A8, int64_t n, int64_t p, int64_t ld8, double* __restrict__ out,
                     int64_t ldo) {
  __shared__ double lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = static_cast<double>(i) / kLevels;
  __syncthreads();
  for (int64_t j = blockIdx.x; j < p; j += gridDim.x) {
    const uint8_t* col = A8 + j * ld8;
    double* dst = out + j * ldo;
    for (int64_t i = threadIdx.x; i < ldo; i += 256) dst[i] = i < n ? lut[col[i]] : 0.0;
  }
}

void launch_decode_u8(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, double* out, int64_t ldo,
                      cudaStream_t stream) {
  if (n <= 0 || p <= 0) return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p, static_cast<int64_t>(sms) * 8));
  decode_u8_kernel<<<grid, 256, 0, stream>>>(A8, n, p, ld8, out, ldo);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
