// Marker-streaming scan: the B200 replacement for the per-marker loops of gwasols
// (/root/reference/src/gwas.jl:239-249) and gwaslmm (:363-389), for std(G, dims=1) and
// the fixed-locus filter of gwasprep (:112-115) and for the ploidy probe (:119).
//
// One pass over the n x p column-major matrix.  A persistent CTA per SM; one producer
// lane issues TMA box loads (256 rows x C markers of A, plus the matching 256 x M slice of
// the side vectors Q) into a shared-memory ring guarded by mbarriers; four consumer
// warps read the ring with 128-bit shared loads (two rows per lane), keep
// sum(d), sum(d^2), sum(d*q_m) per marker in FP64 registers (d = a - a[0]: shifting by
// the column's first element makes constant columns give exactly SS = 0, which is what
// the reference's two-pass std returns for them), reduce across the warp with a halving
// shuffle tree and across warps through shared memory in a fixed order (deterministic,
// independent of the grid size), and write one record per marker.
//
// Algorithmic bytes: 8*n per marker read once from HBM; 8*(2+M) written.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "pvalue.cuh"
#include "reduce.cuh"

namespace gbm {

constexpr int kRows = 256;           // rows per pipeline stage (one TMA box, boxDim <= 256)
constexpr int kConsumerWarps = 4;    // 128 lanes x 2 rows = 256 rows
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kScanThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kSmemBudget = 200 * 1024;

template <int C, int M, bool MINNZ>
struct ScanCfg {
  static constexpr int NSUM = 2 + M;                       // S1, S2, dots
  static constexpr int NV = C * NSUM;                      // reduced values per tile
  static constexpr int NVP = ((NV + 31) / 32) * 32;        // padded for the halving tree
  static constexpr int NS = NSUM + (MINNZ ? 1 : 0);        // record stride
  static constexpr int A_BYTES = kRows * C * 8;
  static constexpr int Q_BYTES = kRows * M * 8;
  static constexpr int STAGE_BYTES = A_BYTES + Q_BYTES;
  static constexpr int STAGES_RAW = kSmemBudget / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int RED_BYTES = 2 * kConsumerWarps * NVP * 8 + 2 * kConsumerWarps * C * 8 + 2 * C * 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RED_BYTES + 2 * STAGES * 8 + 128;
};

struct ScanParams {
  int64_t n, p;
  int num_tiles, chunks;
  double inv_n;
  double* rec;
};

template <int C, int M, bool MINNZ>
__global__ void __launch_bounds__(kScanThreads, 1)
    scan_sums_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ,
                     const ScanParams prm) {
  using Cfg = ScanCfg<C, M, MINNZ>;
  constexpr int NSUM = Cfg::NSUM, NV = Cfg::NV, NVP = Cfg::NVP, NS = Cfg::NS, STAGES = Cfg::STAGES;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  double* red = reinterpret_cast<double*>(smem + STAGES * Cfg::STAGE_BYTES);  // [2][warps][NVP]
  double* redmin = red + 2 * kConsumerWarps * NVP;                            // [2][warps][C]
  double* shift_s = redmin + 2 * kConsumerWarps * C;                          // [2][C]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(shift_s + 2 * C);
  uint64_t* empty_bar = full_bar + STAGES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      if (M > 0) prefetch_tensormap(&tmQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < prm.chunks; ++chunk) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = ring + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(dst, &tmA, chunk * kRows, tile * C, &full_bar[stage], kEvictFirst);
          if (M > 0) tma_load_2d(dst + Cfg::A_BYTES, &tmQ, chunk * kRows, 0, &full_bar[stage], kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // -------------------------------- consumers --------------------------------
  int stage = 0;
  uint32_t phase = 0;
  int parity = 0;
  const int r = 2 * tid;  // this lane's row pair inside a chunk
  for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, parity ^= 1) {
    double v[NVP];
    double mn[MINNZ ? C : 1];
    double sh[C];
#pragma unroll
    for (int i = 0; i < NVP; ++i) v[i] = 0.0;
    if (MINNZ) {
#pragma unroll
      for (int c = 0; c < C; ++c) mn[c] = INFINITY;
    }

    for (int chunk = 0; chunk < prm.chunks; ++chunk) {
      mbar_wait(&full_bar[stage], phase);
      const double* sA = reinterpret_cast<const double*>(ring + stage * Cfg::STAGE_BYTES);
      const double* sQ = sA + kRows * C;
      if (chunk == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) sh[c] = sA[c * kRows];
        if (tid < C) shift_s[parity * C + tid] = sA[tid * kRows];
      }
      double2 q[M > 0 ? M : 1];
#pragma unroll
      for (int m = 0; m < M; ++m) q[m] = *reinterpret_cast<const double2*>(sQ + m * kRows + r);
      const int64_t row0 = static_cast<int64_t>(chunk) * kRows + r;
      if (row0 + 1 < prm.n) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const double2 a = *reinterpret_cast<const double2*>(sA + c * kRows + r);
          const double d0 = a.x - sh[c], d1 = a.y - sh[c];
          v[c * NSUM + 0] += d0 + d1;
          v[c * NSUM + 1] = fma(d1, d1, fma(d0, d0, v[c * NSUM + 1]));
#pragma unroll
          for (int m = 0; m < M; ++m)
            v[c * NSUM + 2 + m] = fma(d1, q[m].y, fma(d0, q[m].x, v[c * NSUM + 2 + m]));
          if (MINNZ) {
            mn[c] = fmin(mn[c], a.x != 0.0 ? a.x : INFINITY);
            mn[c] = fmin(mn[c], a.y != 0.0 ? a.y : INFINITY);
          }
        }
      } else if (row0 < prm.n) {
        // last valid row of the matrix is this lane's first row; the second is padding
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const double ax = sA[c * kRows + r];
          const double d0 = ax - sh[c];
          v[c * NSUM + 0] += d0;
          v[c * NSUM + 1] = fma(d0, d0, v[c * NSUM + 1]);
#pragma unroll
          for (int m = 0; m < M; ++m) v[c * NSUM + 2 + m] = fma(d0, q[m].x, v[c * NSUM + 2 + m]);
          if (MINNZ) mn[c] = fmin(mn[c], ax != 0.0 ? ax : INFINITY);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }

    // ---- tile epilogue: warp tree, then a fixed-order sum over the four warps ----
    HalvingStep<NVP / 2, 16, NVP>::run(v, lane);
    {
      const int base = ((lane & 16) ? NVP / 2 : 0) + ((lane & 8) ? NVP / 4 : 0) + ((lane & 4) ? NVP / 8 : 0) +
                       ((lane & 2) ? NVP / 16 : 0) + ((lane & 1) ? NVP / 32 : 0);
      double* dst = red + (parity * kConsumerWarps + warp) * NVP + base;
#pragma unroll
      for (int i = 0; i < NVP / 32; ++i) dst[i] = v[i];
    }
    if (MINNZ) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        double x = mn[c];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
        if (lane == 0) redmin[(parity * kConsumerWarps + warp) * C + c] = x;
      }
    }
    named_bar_sync(1, kConsumerThreads);
    if (tid < NV) {
      const int c = tid / NSUM, kk = tid - c * NSUM;
      const double* rp = red + parity * kConsumerWarps * NVP;
      double tot = 0.0, s1 = 0.0;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        tot += rp[w * NVP + tid];
        s1 += rp[w * NVP + c * NSUM];
      }
      double out;
      if (kk == 0)
        out = shift_s[parity * C + c] + s1 * prm.inv_n;  // mean
      else if (kk == 1)
        out = fmax(tot - s1 * s1 * prm.inv_n, 0.0);  // centred sum of squares
      else
        out = tot;  // dot with a side vector (shift-invariant: q is orthogonal to 1)
      const int64_t col = static_cast<int64_t>(tile) * C + c;
      if (col < prm.p) prm.rec[col * NS + kk] = out;
    }
    if (MINNZ && tid >= kConsumerThreads - C) {
      const int c = tid - (kConsumerThreads - C);
      const double* mp = redmin + parity * kConsumerWarps * C;
      double x = fmin(fmin(mp[c], mp[C + c]), fmin(mp[2 * C + c], mp[3 * C + c]));
      const int64_t col = static_cast<int64_t>(tile) * C + c;
      if (col < prm.p) prm.rec[col * NS + NSUM] = x;
    }
    // red / shift buffers are double-buffered by tile parity: the barrier of the next tile
    // orders these reads before the writes of the tile after it.
  }
}

// ------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------
static int padded_M(int M) {
  if (M <= 0) return 0;
  if (M <= 1) return 1;
  if (M <= 2) return 2;
  if (M <= 4) return 4;
  if (M <= 6) return 6;
  if (M <= 10) return 10;
  if (M <= 14) return 14;
  return -1;
}
int scan_max_side_vectors() { return 14; }
int scan_record_stride(int M, bool minnz) {
  const int Mp = padded_M(M);
  return 2 + Mp + ((minnz || Mp == 0) ? 1 : 0);  // the M = 0 kernel always tracks min-nonzero
}

template <int C, int M, bool MINNZ>
static void launch_cfg(const double* A, int64_t n, int64_t p, int64_t lda, const double* Q, int64_t ldq,
                       double* rec, int sm_count, cudaStream_t stream) {
  using Cfg = ScanCfg<C, M, MINNZ>;
  static_assert(Cfg::STAGES >= 2, "ring too shallow");
  alignas(64) CUtensorMap tmA, tmQ;
  make_tensor_map_2d_f64(&tmA, A, static_cast<uint64_t>(n), static_cast<uint64_t>(p), static_cast<uint64_t>(lda),
                         kRows, C);
  if (M > 0)
    make_tensor_map_2d_f64(&tmQ, Q, static_cast<uint64_t>(n), static_cast<uint64_t>(M),
                           static_cast<uint64_t>(ldq), kRows, M);
  else
    tmQ = tmA;
  ScanParams prm;
  prm.n = n;
  prm.p = p;
  prm.num_tiles = static_cast<int>((p + C - 1) / C);
  prm.chunks = static_cast<int>((n + kRows - 1) / kRows);
  prm.inv_n = 1.0 / static_cast<double>(n);
  prm.rec = rec;
  auto kern = scan_sums_kernel<C, M, MINNZ>;
  GBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = prm.num_tiles < sm_count ? prm.num_tiles : sm_count;
  kern<<<grid, kScanThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmQ, prm);
  GBM_CUDA(cudaGetLastError());
}

void launch_scan_sums(const double* A, int64_t n, int64_t p, int64_t lda, const double* Q, int M, int64_t ldq,
                      bool minnz, double* rec, int sm_count, cudaStream_t stream) {
  if (p <= 0 || n <= 0) return;
  const int Mp = padded_M(M);
  if (Mp < 0) GBM_THROW(1, "scan: too many side vectors for one pass");
  if (minnz) {
    if (Mp != 0) GBM_THROW(1, "scan: min-nonzero tracking is only built for M = 0");
    launch_cfg<16, 0, true>(A, n, p, lda, Q, ldq, rec, sm_count, stream);
    return;
  }
  switch (Mp) {
    case 0: launch_cfg<16, 0, true>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 1: launch_cfg<16, 1, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 2: launch_cfg<16, 2, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 4: launch_cfg<8, 4, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 6: launch_cfg<8, 6, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 10: launch_cfg<4, 10, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    case 14: launch_cfg<4, 14, false>(A, n, p, lda, Q, ldq, rec, sm_count, stream); break;
    default: GBM_THROW(1, "scan: unsupported side-vector count");
  }
}

// ------------------------------------------------------------------------------------
// finalisation kernels: statistics + log-space p-values
// ------------------------------------------------------------------------------------
constexpr double kEps = 2.220446049250313e-16;

// one thread per (marker, trait): blockIdx.y is the trait, so a batch of traits fills the machine instead of
// looping serially in each marker's thread (the t-distribution tail is a continued fraction per value)
__global__ void __launch_bounds__(256) scan_finalize_kernel(const FinalizeParams prm) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= prm.p) return;
  const int t = blockIdx.y;
  const double* rec = prm.rec + j * prm.rec_stride;
  const double mean = rec[0], ss = rec[1];
  const double sd = sqrt(ss / static_cast<double>(prm.n - 1));
  // fixed-locus filter: v > eps && finite   (gwas.jl:113)
  const bool keep = (sd > kEps) && isfinite(sd);
  if (t == 0) {
    if (prm.mean) prm.mean[j] = mean;
    if (prm.sd) prm.sd[j] = sd;
    if (prm.keep) prm.keep[j] = keep ? 1 : 0;
  }
  double uu = 0.0;
  for (int i = 0; i < prm.k; ++i) uu = fma(rec[2 + i], rec[2 + i], uu);
  const double xMx = ss - uu;  // x'Mx on the raw scale, M = projector off [1, C]
  // A marker inside span[1, C] makes X'X singular.  The reference's `pinv(X' * X)` (gwas.jl:242) truncates singular
  // values below rtol * sigma_max with rtol = eps * min(size) = (k + 2) eps; sigma_max ~ n and the small one is
  // ~ xMx / (sd^2 n), so truncation happens below xMx / SS ~ (k + 2) eps n -- and its minimum-norm solution is finite.
  const double thr = fmax(static_cast<double>(prm.k + 2) * kEps * static_cast<double>(prm.n), 16.0 * kEps);
  const bool ok = keep && (xMx > thr * ss);
  const double dfres = static_cast<double>(prm.n - prm.k - 2);
  const int64_t o = static_cast<int64_t>(t) * prm.ld_out + j;
  double beta = NAN, se = NAN, stat = NAN, nlp = NAN;
  if (ok) {
    const double xMy = rec[2 + prm.k + t];
    const double s = xMy / sqrt(xMx);   // gwasols statistic (gwas.jl:245), SURVEY App. A.2
    beta = xMy * sd / xMx;              // coefficient of the standardised column
    const double se_ols = sd / sqrt(xMx);
    if (prm.model == 0) {
      stat = s;
      se = se_ols;
      nlp = -log_sf_t(s, static_cast<double>(prm.n - 1)) * 0.4342944819032518;
    } else {
      const double rss = prm.yMy[t] - s * s;
      const double sigma2 = rss / dfres;  // REML sigma^2 (1 + theta^2), SURVEY App. A.3
      stat = s / sqrt(sigma2);
      se = se_ols * sqrt(sigma2);
      nlp = -log_sf_normal(stat) * 0.4342944819032518;
    }
    if (prm.flags & 1) nlp -= 0.3010299956639812;  // two-sided
  } else if (keep && prm.model == 0 && prm.k > 0 && uu > 0.0) {
    // gwasols on a marker collinear with the covariates: the truncated pseudo-inverse's minimum-norm solution.  With
    // c = C'g (g the standardised column, C orthonormal) and h = C'y:  b_g = c'h / (1 + c'c),
    // Vinv_gg = c'c / (1 + c'c)^2, statistic c'h / |c|  (for the reference's X = [1, PC1, g]: sign(c) PC1'y).
    double uh = 0.0;
    for (int i = 0; i < prm.k; ++i) uh = fma(rec[2 + i], prm.wy[t * prm.k + i], uh);
    const double cc = uu / (sd * sd), ch = uh / sd;
    beta = ch / (1.0 + cc);
    se = sqrt(cc) / (1.0 + cc);
    stat = uh / sqrt(uu);
    nlp = -log_sf_t(stat, static_cast<double>(prm.n - 1)) * 0.4342944819032518;
    if (prm.flags & 1) nlp -= 0.3010299956639812;
  }
  // gwaslmm on such a marker: MixedModels' rank-deficient fit is not pinned (source absent); NaN here, the host
  // mirror writes the 0.0 a failed fit leaves (gwas.jl:367-382)
  if (prm.beta) prm.beta[o] = beta;
  if (prm.se) prm.se[o] = se;
  if (prm.stat) prm.stat[o] = stat;
  if (prm.nlp) prm.nlp[o] = nlp;
}

void launch_scan_finalize(const FinalizeParams& prm, cudaStream_t stream) {
  if (prm.p <= 0 || prm.T <= 0) return;
  const dim3 grid(static_cast<unsigned>((prm.p + 255) / 256), static_cast<unsigned>(prm.T));
  scan_finalize_kernel<<<grid, 256, 0, stream>>>(prm);
  GBM_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
    colstats_finalize_kernel(const double* __restrict__ rec, int rec_stride, int64_t n, int64_t p,
                             double* mean, double* sd, double* minnz, uint8_t* keep) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= p) return;
  const double* r = rec + j * rec_stride;
  const double s = sqrt(r[1] / static_cast<double>(n - 1));
  if (mean) mean[j] = r[0];
  if (sd) sd[j] = s;
  if (minnz) minnz[j] = r[rec_stride - 1];
  if (keep) keep[j] = ((s > kEps) && isfinite(s)) ? 1 : 0;
}

void launch_colstats_finalize(const double* rec, int rec_stride, int64_t n, int64_t p, double* mean, double* sd,
                              double* minnz, uint8_t* keep, cudaStream_t stream) {
  if (p <= 0) return;
  const unsigned grid = static_cast<unsigned>((p + 255) / 256);
  colstats_finalize_kernel<<<grid, 256, 0, stream>>>(rec, rec_stride, n, p, mean, sd, minnz, keep);
  GBM_CUDA(cudaGetLastError());
}

// idx_cols = findall(keep) 1-based ascending: single CTA, chunked ballot scan (p <= ~1e7 is
// a few hundred microseconds; this is bookkeeping, not the hot path).
__global__ void __launch_bounds__(1024)
    compact_keep_kernel(const uint8_t* __restrict__ keep, const double* __restrict__ minnz, int64_t p,
                        int64_t* idx_cols, int64_t* n_keep, double* min_kept) {
  __shared__ int warp_cnt[32];
  __shared__ int64_t base_s;
  __shared__ double wmin[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  double mymin = INFINITY;
  __syncthreads();
  for (int64_t j0 = 0; j0 < p; j0 += 1024) {
    const int64_t j = j0 + tid;
    const bool k = (j < p) && keep[j];
    if (k && minnz) mymin = fmin(mymin, minnz[j]);
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = warp_cnt[w];
      if (w < warp) off += c;
      tot += c;
    }
    const int64_t base = base_s;
    if (k && idx_cols) idx_cols[base + off + __popc(bal & ((1u << lane) - 1u))] = j + 1;
    __syncthreads();
    if (tid == 0) base_s = base + tot;
    __syncthreads();
  }
  for (int o = 16; o >= 1; o >>= 1) mymin = fmin(mymin, __shfl_xor_sync(0xffffffffu, mymin, o));
  if (lane == 0) wmin[warp] = mymin;
  __syncthreads();
  if (tid == 0) {
    double m = INFINITY;
    for (int w = 0; w < 32; ++w) m = fmin(m, wmin[w]);
    if (min_kept) *min_kept = m;
    if (n_keep) *n_keep = base_s;
  }
}

void launch_compact_keep(const uint8_t* keep, const double* minnz, int64_t p, int64_t* idx_cols, int64_t* n_keep,
                         double* min_kept, cudaStream_t stream) {
  compact_keep_kernel<<<1, 1024, 0, stream>>>(keep, minnz, p, idx_cols, n_keep, min_kept);
  GBM_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
    neglog10_sf_kernel(const double* __restrict__ stat, int64_t len, int dist, double df, double* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const double ln = dist == 0 ? log_sf_t(stat[i], df) : log_sf_normal(stat[i]);
  out[i] = -ln * 0.4342944819032518;
}

void launch_neglog10_sf(const double* stat, int64_t len, int dist, double df, double* out, cudaStream_t stream) {
  if (len <= 0) return;
  neglog10_sf_kernel<<<static_cast<unsigned>((len + 255) / 256), 256, 0, stream>>>(stat, len, dist, df, out);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
