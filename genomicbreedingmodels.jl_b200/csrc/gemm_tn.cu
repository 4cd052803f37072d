// Eigen-rotation GEMM  C = A' * B  (A: K x M, B: K x N, both column-major, i.e. both operands
// contiguous along the contraction index; C: M x N column-major).  Used for U' * X of the
// GRM-covariance LMM scan (BASELINE.json north_star (b): "the eigen-rotation U'.X as FP64
// tensor-core (DMMA) tiled GEMM"; reference model: gwasreml, /root/reference/src/gwas.jl:
// 549-613, whose per-evaluation pinv(V) this rotation replaces).
//
// Same machinery as grm.cu: persistent CTAs, one producer lane issuing TMA boxes into an
// mbarrier ring, eight DMMA (mma.sync m8n8k4.f64) warps with 64x32 accumulator blocks.  Both
// operands are K-major here, so a tile is fetched as a box of 20 contraction rows x 128
// columns (16 used + 4 over-fetched): the 160-byte smem pitch is 4 (mod 16) doubles, which
// makes the fragment loads (lane -> column = lane>>2, k = lane&3) conflict-free without a
// swizzle.  TMA zero-fills out-of-range rows/columns, so ragged M, N, K need no padding.
//
// Algorithmic flops: 2 M N K.
#include "common.cuh"
#include "kernels.h"

namespace gbm {

constexpr int kGT = 128;       // CTA tile (M and N)
constexpr int kGK = 16;        // contraction rows consumed per stage
constexpr int kGKBox = 20;     // contraction rows fetched per stage (pitch = 4 mod 16)
constexpr int kGStages = 4;
constexpr int kGTileBytes = kGT * kGKBox * 8;          // 20480
constexpr int kGStageBytes = 2 * kGTileBytes;          // 40960
constexpr int kGConsumerWarps = 8;
constexpr int kGThreads = (kGConsumerWarps + 1) * 32;
constexpr int kGSmemBytes = kGStages * kGStageBytes + 2 * kGStages * 8 + 128;

struct GemmParams {
  int64_t M, N, K, ldc;
  int tiles_m, tiles_n, ksteps;
  double* C;
};

__device__ __forceinline__ void dmma884_tn(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kGThreads, 1)
    gemm_tn_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGStages * kGStageBytes);
  uint64_t* empty_bar = full_bar + kGStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kGConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int num_tiles = prm.tiles_m * prm.tiles_n;

  if (warp == kGConsumerWarps) {
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      prefetch_tensormap(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // consecutive tiles share the B (marker) block: column-block-major order
        const int tn = tile / prm.tiles_m, tm = tile - tn * prm.tiles_m;
        for (int s = 0; s < prm.ksteps; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * kGStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kGStageBytes);
          tma_load_2d(dst, &tmA, s * kGK, tm * kGT, &full_bar[stage], kEvictLast);
          tma_load_2d(dst + kGTileBytes, &tmB, s * kGK, tn * kGT, &full_bar[stage], kEvictNormal);
          if (++stage == kGStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tn = tile / prm.tiles_m, tm = tile - tn * prm.tiles_m;
    double acc[8][4][2];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    for (int s = 0; s < prm.ksteps; ++s) {
      mbar_wait(&full_bar[stage], phase);
      const double* sA = reinterpret_cast<const double*>(smem + stage * kGStageBytes);
      const double* sB = sA + kGT * kGKBox;
      const double* pa = sA + (wm * 64 + g) * kGKBox + t;
      const double* pb = sB + (wn * 32 + g) * kGKBox + t;
      // rows >= K of the last step are zero-filled by TMA, so no predicate is needed; the 4
      // over-fetched rows (16..19) are simply never read
#pragma unroll
      for (int kk = 0; kk < kGK / 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) a[mt] = pa[mt * 8 * kGKBox + kk * 4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = pb[nt * 8 * kGKBox + kk * 4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884_tn(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == kGStages) {
        stage = 0;
        phase ^= 1u;
      }
    }

    const int64_t row_base = static_cast<int64_t>(tm) * kGT + wm * 64 + g;
    const int64_t col_base = static_cast<int64_t>(tn) * kGT + wn * 32 + 2 * t;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      const int64_t row = row_base + mt * 8;
      if (row >= prm.M) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t col = col_base + nt * 8 + e;
          if (col < prm.N) prm.C[col * prm.ldc + row] = acc[mt][nt][e];
        }
      }
    }
  }
}

void launch_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                    int64_t N, int64_t K, int sm_count, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  // The K tail: a box that starts inside K but runs past it is zero-filled only beyond the
  // tensor's extent, so the tensor maps are declared with exactly K rows.
  alignas(64) CUtensorMap tmA, tmB;
  make_tensor_map_2d_f64(&tmA, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda),
                         kGKBox, kGT);
  make_tensor_map_2d_f64(&tmB, B, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(ldb),
                         kGKBox, kGT);
  GemmParams prm;
  prm.M = M;
  prm.N = N;
  prm.K = K;
  prm.ldc = ldc;
  prm.tiles_m = static_cast<int>((M + kGT - 1) / kGT);
  prm.tiles_n = static_cast<int>((N + kGT - 1) / kGT);
  prm.ksteps = static_cast<int>((K + kGK - 1) / kGK);
  prm.C = C;
  GBM_CUDA(cudaFuncSetAttribute(gemm_tn_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemBytes));
  const int64_t tiles = static_cast<int64_t>(prm.tiles_m) * prm.tiles_n;
  const int grid = static_cast<int>(tiles < sm_count ? tiles : sm_count);
  gemm_tn_dmma_kernel<<<grid, kGThreads, kGSmemBytes, stream>>>(tmA, tmB, prm);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
