// Multi-trait marker scan: the same per-marker sums as scan.cu -- mean, SS and the dots a'q_m against the
// M = k + T side vectors (covariates + residualised traits; /root/reference/src/gwas.jl:239-249 for every
// trait of a batch, BASELINE configs[4]) -- for M > 2, where the dots are a skinny GEMM
//     D (p x M) = X' Q,   X: n x p genotypes, Q: n x M,
// and belong on the FP64 tensor pipe: at T = 20 the FMA formulation needs 45 flop per 8 bytes and
// the CUDA-core kernel fell to 1.0 TB/s-equivalent in two passes; DMMA does the M <= 31 dots in ONE pass.
//
// Structure = gemm_tn.cu (persistent CTAs, one producer lane issuing TMA boxes into an mbarrier ring, eight
// mma.sync m8n8k4.f64 warps), specialised: a CTA tile is 128 markers x 8 NT side-vector columns (NT = 1..4),
// each warp owns 16 markers, a stage is 32 genotype rows (box of 36: the 288-byte pitch is 4 (mod 16)
// doubles, so the fragment loads are conflict-free without a swizzle).  Column 0 of Q is the vector of ones:
// with the operand shifted by the marker's first genotype, d = a - a[0], its dot is S1 = sum d; S2 = sum d^2
// is one DFMA per loaded fragment element.  mean = a[0] + S1/n and SS = S2 - S1^2/n exactly as in scan.cu
// (a constant marker gives SS = 0 exactly), the other columns are the dots (q_m is orthogonal to 1, so
// d'q_m = a'q_m).  Records [mean, SS, dot_1 .. dot_M] feed the shared finalisation kernel.
//
// One-byte dosage codes (scan_u8.cu's storage) run the same kernel natively (geometry V = 2 below).
//
// Algorithmic bytes: 8 n per marker (n for codes), read once.  FP64-pipe work per marker: 2 n 8 NT flops of DMMA.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace gbm {

namespace {

// Two geometries (chosen per NT by measurement):
//  V = 0: 128 markers x 32 rows per stage (36-row box), 8 DMMA warps, deep ring in one CTA per SM -- HBM-bound cases
//  V = 1: 256 markers x 16 rows per stage (20-row box), 16 DMMA warps (4 per scheduler), 4 stages -- DMMA-bound cases
//  V = 2: one-byte dosage codes (a = code / 240): 256 markers x 32 rows per stage, the code tile is a 48-byte box per
//         marker (32 used; the 12-word pitch puts the 8 markers of a fragment in 8 different banks), codes become
//         doubles by the 2^52 trick fused with the shift (one DADD, as in the Float64 variants), sums stay in code
//         units (S1, S2 exact integers) and are scaled by 1/240 resp. 1/240^2 in the epilogue
template <int NT, int V>
struct MtCfg {
  static constexpr bool U8 = V == 2;
  static constexpr int MARKERS = V == 0 ? 128 : 256;  // markers per CTA tile, 16 per warp
  static constexpr int K = V == 1 ? 16 : 32;          // genotype rows consumed per stage
  static constexpr int KBOX = K + 4;                  // side-vector rows fetched per stage (pitch = 4 mod 16 doubles)
  static constexpr int APITCH = U8 ? 48 : KBOX * 8;   // bytes per marker in the genotype tile
  static constexpr int WARPS = MARKERS / 16;          // DMMA warps
  static constexpr int THREADS = (WARPS + 1) * 32;
  static constexpr int A_BYTES = MARKERS * APITCH;
  static constexpr int Q_BYTES = NT * 8 * KBOX * 8;
  static constexpr int STAGE_BYTES = A_BYTES + Q_BYTES;
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
};

struct MtParams {
  int64_t n, p;
  int M;            // side vectors without the ones column
  int rec_stride;   // doubles per marker record (>= 2 + M)
  int num_tiles, ksteps;
  double inv_n;
  double* rec;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int NT, int V>
__global__ void __launch_bounds__(MtCfg<NT, V>::THREADS, 1)
    scan_sums_mt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ,
                        const MtParams prm) {
  using Cfg = MtCfg<NT, V>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int kMtMarkers = Cfg::MARKERS, kMtK = Cfg::K, kMtKBox = Cfg::KBOX, kMtWarps = Cfg::WARPS;
  constexpr int kMtABytes = Cfg::A_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMtWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kMtWarps) {  // producer
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      prefetch_tensormap(&tmQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        for (int s = 0; s < prm.ksteps; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(dst, &tmA, s * kMtK, tile * kMtMarkers, &full_bar[stage], kEvictNormal);
          tma_load_2d(dst + kMtABytes, &tmQ, s * kMtK, 0, &full_bar[stage], kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  const int g = lane >> 2, t = lane & 3;  // fragment coordinates: marker (or side vector) g, row t of the k4 step
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
    double acc[2][NT][2];
    double s2[2] = {0.0, 0.0}, shift[2] = {0.0, 0.0};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    for (int s = 0; s < prm.ksteps; ++s) {
      mbar_wait(&full_bar[stage], phase);
      const uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
      const double* sQ = reinterpret_cast<const double*>(sA + kMtABytes);
      const uint8_t* pa = sA + (warp * 16 + g) * Cfg::APITCH + t * (Cfg::U8 ? 1 : 8);
      const double* pq = sQ + g * kMtKBox + t;
      // genotype (mt, kk) of this lane as a double: the Float64 itself, or 2^52 + code (exact) for codes
      auto fetch = [&](const uint8_t* base) -> double {
        if constexpr (Cfg::U8)
          return __hiloint2double(0x43300000, static_cast<int>(*base));
        else
          return *reinterpret_cast<const double*>(base);
      };
      if (s == 0) {  // the marker's first genotype (row 0 of the first stage)
        shift[0] = fetch(sA + (warp * 16 + g) * Cfg::APITCH);
        shift[1] = fetch(sA + (warp * 16 + 8 + g) * Cfg::APITCH);
      }
      const int64_t row0 = static_cast<int64_t>(s) * kMtK + t;
      const bool full = static_cast<int64_t>(s + 1) * kMtK <= prm.n;
#pragma unroll
      for (int kk = 0; kk < kMtK / 4; ++kk) {
        double a[2], q[NT];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          a[mt] = fetch(pa + mt * 8 * Cfg::APITCH + kk * 4 * (Cfg::U8 ? 1 : 8)) - shift[mt];
          if (!full && row0 + kk * 4 >= prm.n) a[mt] = 0.0;  // rows past n are zero-filled by TMA: d must be 0 too
          s2[mt] = fma(a[mt], a[mt], s2[mt]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) q[nt] = pq[nt * 8 * kMtKBox + kk * 4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], q[nt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }

    // epilogue: lane (g, t) holds D[marker g][columns 2t, 2t+1] of every 8 x 8 block; S2 is spread over the
    // four t lanes of a marker (fixed-order butterfly)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      s2[mt] += __shfl_xor_sync(0xffffffffu, s2[mt], 1);
      s2[mt] += __shfl_xor_sync(0xffffffffu, s2[mt], 2);
      const double S1 = __shfl_sync(0xffffffffu, acc[mt][0][0], lane & ~3);  // column 0 (ones) lives in lane t = 0
      const int64_t marker = static_cast<int64_t>(tile) * kMtMarkers + warp * 16 + mt * 8 + g;
      if (marker >= prm.p) continue;
      double* rec = prm.rec + marker * prm.rec_stride;
      constexpr double denom = Cfg::U8 ? 240.0 : 1.0;  // codes -> allele frequencies (as scan_u8.cu: divisions)
      const double first = Cfg::U8 ? shift[mt] - 4503599627370496.0 : shift[mt];
      if (t == 0) {
        rec[0] = (first + S1 * prm.inv_n) / denom;
        rec[1] = fmax(s2[mt] - S1 * S1 * prm.inv_n, 0.0) / (denom * denom);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nt * 8 + 2 * t + e;  // column of [1 | Q]
          if (col >= 1 && col <= prm.M) rec[1 + col] = acc[mt][nt][e] / denom;
        }
    }
  }
}

template <int NT, int V>
void launch_mt(const void* A, int64_t n, int64_t p, int64_t lda, const double* Qx, int M, int64_t ldq, double* rec,
               int rec_stride, int sm_count, cudaStream_t stream) {
  using Cfg = MtCfg<NT, V>;
  static_assert(Cfg::STAGES >= 3, "ring too shallow");
  alignas(64) CUtensorMap tmA, tmQ;
  // tensor maps are declared with exactly n rows and M + 1 columns: TMA zero-fills whatever a box reads beyond
  if constexpr (Cfg::U8)
    make_tensor_map_2d_u8(&tmA, A, static_cast<uint64_t>(n), static_cast<uint64_t>(p), static_cast<uint64_t>(lda),
                          Cfg::APITCH, Cfg::MARKERS);
  else
    make_tensor_map_2d_f64(&tmA, static_cast<const double*>(A), static_cast<uint64_t>(n), static_cast<uint64_t>(p),
                           static_cast<uint64_t>(lda), Cfg::KBOX, Cfg::MARKERS);
  make_tensor_map_2d_f64(&tmQ, Qx, static_cast<uint64_t>(n), static_cast<uint64_t>(M + 1), static_cast<uint64_t>(ldq),
                         Cfg::KBOX, NT * 8);
  MtParams prm;
  prm.n = n;
  prm.p = p;
  prm.M = M;
  prm.rec_stride = rec_stride;
  prm.num_tiles = static_cast<int>((p + Cfg::MARKERS - 1) / Cfg::MARKERS);
  prm.ksteps = static_cast<int>((n + Cfg::K - 1) / Cfg::K);
  prm.inv_n = 1.0 / static_cast<double>(n);
  prm.rec = rec;
  auto kern = scan_sums_mt_kernel<NT, V>;
  GBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = prm.num_tiles < sm_count ? prm.num_tiles : sm_count;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmQ, prm);
  GBM_CUDA(cudaGetLastError());
}

static int mt_variant(int nt) {
  static const int forced = [] {
    const char* e = getenv("GBM_MT_VARIANT");  // measurement switch
    return e ? atoi(e) : -1;
  }();
  if (forced == 0 || forced == 1) return forced;
  // measured (n = 10,000, p = 400,000): T = 5: 4.64 (V0) / 4.70 ms (V1); T = 13: 5.10 / 5.00; T = 20: 7.43 / 6.76
  return nt >= 2 ? 1 : 0;
}

}  // namespace

int scan_mt_max_side_vectors() { return 31; }

void launch_scan_sums_mt(const double* A, int64_t n, int64_t p, int64_t lda, const double* Qx, int M, int64_t ldq,
                         double* rec, int rec_stride, int sm_count, cudaStream_t stream) {
  if (p <= 0 || n <= 0) return;
  if (M < 1 || M > 31 || rec_stride < 2 + M) GBM_THROW(1, "multi-trait scan: 1..31 side vectors per pass");
  const int nt = (M + 1 + 7) / 8;
  const int v = mt_variant(nt);
#define GBM_MT_CASE(NT_)                                                                                   \
  case NT_:                                                                                                \
    if (v == 0) launch_mt<NT_, 0>(A, n, p, lda, Qx, M, ldq, rec, rec_stride, sm_count, stream);            \
    else launch_mt<NT_, 1>(A, n, p, lda, Qx, M, ldq, rec, rec_stride, sm_count, stream);                   \
    break;
  switch (nt) {
    GBM_MT_CASE(1)
    GBM_MT_CASE(2)
    GBM_MT_CASE(3)
    default:
      if (v == 0) launch_mt<4, 0>(A, n, p, lda, Qx, M, ldq, rec, rec_stride, sm_count, stream);
      else launch_mt<4, 1>(A, n, p, lda, Qx, M, ldq, rec, rec_stride, sm_count, stream);
  }
#undef GBM_MT_CASE
}

// the same from one-byte dosage codes (a = code / 240): A8 is n x p bytes with column pitch ld8 (multiple of 16)
void launch_scan_sums_mt_u8(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, const double* Qx, int M, int64_t ldq,
                            double* rec, int rec_stride, int sm_count, cudaStream_t stream) {
  if (p <= 0 || n <= 0) return;
  if (M < 1 || M > 31 || rec_stride < 2 + M) GBM_THROW(1, "multi-trait scan: 1..31 side vectors per pass");
  switch ((M + 1 + 7) / 8) {
    case 1: launch_mt<1, 2>(A8, n, p, ld8, Qx, M, ldq, rec, rec_stride, sm_count, stream); break;
    case 2: launch_mt<2, 2>(A8, n, p, ld8, Qx, M, ldq, rec, rec_stride, sm_count, stream); break;
    case 3: launch_mt<3, 2>(A8, n, p, ld8, Qx, M, ldq, rec, rec_stride, sm_count, stream); break;
    default: launch_mt<4, 2>(A8, n, p, ld8, Qx, M, ldq, rec, rec_stride, sm_count, stream); break;
  }
}

}  // namespace gbm
