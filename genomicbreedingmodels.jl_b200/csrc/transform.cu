// Transformation screens: the per-locus and pairwise OLS screens of
// transform1 / transform2 (/root/reference/src/transformation.jl:130-239, :319-466).
//
// The reference fits `ols(genomes = g, phenomes = p)` -- `[1 f(x)] \ y`, linear.jl:85 -- once per
// locus (transform1) or once per ordered pair of loci (transform2, l^2 fits) and keeps
// b_hat[2].  Here the slope comes from three sums per feature z = f(.):
//   s1 = sum d, s2 = sum d^2, sy = sum d * (y - ybar),  d = z - z[first row]
// (shifted so that a constant feature gives szz = 0 exactly), beta = sy / (s2 - s1^2 / n), with the
// rank test and minimum-norm solution of Julia's pivoted-QR `\` when [1 z] is rank deficient.
//
//   transform1_scan_kernel  one warp per locus, HBM-bound: 8 n bytes per locus, read once
//   transform_prep_kernel   X' = abs?(X + eps) once (and log X' for raise), so the pair kernel's inner loop
//                           is nothing but the regression sums
//   transform2_scan_kernel  64 x 64 tiles of pairs, rows streamed through shared memory (cp.async,
//                           two stages); FP64-pipe-bound: 4 FP64 instructions per (pair, row) for mult
//   transform{1,2}_apply    materialise the selected features T = f.(X[:, idx]) with the eps clean-up
//   transform_select        sortperm(abs.(beta), rev = true)[1:n_new] + abs(beta) > eps (stable radix sort)
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/gbm_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace gbm {

namespace {

constexpr double kEpsF64 = 2.220446049250313e-16;     // eps(Float64)
constexpr double kLog10Eps = -15.653559774527022;     // log10(eps(Float64))

// X .+= eps; X = abs.(X) (transformation.jl:158-162)
__device__ __forceinline__ double prep(double a, double eps, int use_abs) {
  a += eps;
  return use_abs ? fabs(a) : a;
}

// named endofunctions of one argument (transformation.jl:9, :18, :27); F = -1: identity (variance pass)
template <int F>
__device__ __forceinline__ double f1(double x) {
  if constexpr (F == GBM_F1_SQUARE) return __dmul_rn(x, x);
  else if constexpr (F == GBM_F1_INVONEPLUS) return 1.0 / (1.0 + x);
  else if constexpr (F == GBM_F1_LOG10EPS) return log10(x + kEpsF64) / kLog10Eps;
  else return x;
}
// ... of two arguments (transformation.jl:36, :45, :54)
template <int F>
__device__ __forceinline__ double f2(double x, double y) {
  if constexpr (F == GBM_F2_MULT) return __dmul_rn(x, y);
  else if constexpr (F == GBM_F2_ADDNORM) return __dadd_rn(x, y) / 2.0;
  else return pow(x, y);
}

// b_hat[2] of [1 z] \ y from the shifted sums.  Julia's `\` on a tall matrix is a pivoted QR with
// rank truncation at sigma_2 / sigma_1 < 2 eps and the minimum-norm solution below it; with
// lambda_1 lambda_2 = n szz and lambda_1 + lambda_2 = n + sum z^2 the test is n szz <= 4 eps^2 (n + sum z^2)^2,
// and the minimum-norm solution of a constant feature z = c is c ybar / (1 + c^2).
__device__ __forceinline__ double slope_from_sums(double n, double z0, double s1, double s2, double sy, double ybar) {
  const double szz = s2 - s1 * s1 / n;
  const double szsq = s2 + 2.0 * z0 * s1 + n * z0 * z0;
  const double tr = n + szsq;
  if (!(n * szz > 4.0 * kEpsF64 * kEpsF64 * tr * tr)) {
    const double c = z0 + s1 / n;
    return c * ybar / (1.0 + c * c);
  }
  return sy / szz;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---- transform1: one warp per locus ------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(256)
    transform1_scan_kernel(const double* __restrict__ A, int64_t n, int64_t p, int64_t lda,
                           const double* __restrict__ yc, double ybar, double eps, int use_abs, double var_thr,
                           double* __restrict__ beta, double* __restrict__ colvar) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  const int64_t n2 = n >> 1;
  const double2* __restrict__ y2 = reinterpret_cast<const double2*>(yc);
  for (int64_t j = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5; j < p; j += nwarps) {
    const double* col = A + j * lda;
    const double x0 = prep(col[0], eps, use_abs);
    const double z0 = f1<F>(x0);
    double v1 = 0, v2 = 0, s1 = 0, s2 = 0, sy = 0;
    auto acc = [&](double a, double y) {
      const double x = prep(a, eps, use_abs);
      const double dx = x - x0;
      v1 += dx;
      v2 = fma(dx, dx, v2);
      if constexpr (F >= 0) {
        const double d = f1<F>(x) - z0;
        s1 += d;
        s2 = fma(d, d, s2);
        sy = fma(d, y, sy);
      }
    };
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(col);
#pragma unroll 4
    for (int64_t i = lane; i < n2; i += 32) {
      const double2 a = c2[i];
      const double2 y = y2[i];
      acc(a.x, y.x);
      acc(a.y, y.y);
    }
    if ((n & 1) && lane == 0) acc(col[n - 1], yc[n - 1]);
    v1 = warp_sum(v1);
    v2 = warp_sum(v2);
    if constexpr (F >= 0) {
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      sy = warp_sum(sy);
    }
    if (lane == 0) {
      const double dn = static_cast<double>(n);
      const double var = (v2 - v1 * v1 / dn) / (dn - 1.0);  // var(x), corrected (transformation.jl:183)
      if (colvar) colvar[j] = var;
      if constexpr (F >= 0) beta[j] = (var < var_thr) ? 0.0 : slope_from_sums(dn, z0, s1, s2, sy, ybar);
    }
  }
}

// ---- transform2: 64 x 64 tiles of (i, j) pairs --------------------------------------------
constexpr int kTile = 64;    // loci per tile edge
constexpr int kRows = 32;    // rows per stage
constexpr int kPitch = 33;   // odd pitch: the 16 lanes of a half-warp hit 16 different bank pairs
struct T2Stage {
  double ai[kTile * kPitch];
  double aj[kTile * kPitch];
  double y[kRows];
};
constexpr int kRaiseFast = 3;  // raise through exp(x_j log x_i) on a precomputed log matrix (all x' > 0)

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// X' = prep(X) (and L = log X' when Lp != null); *nonpos is raised when some x' is not a positive finite number
__global__ void __launch_bounds__(256)
    transform_prep_kernel(const double* __restrict__ A, int64_t n, int64_t lda, double eps, int use_abs,
                          double* __restrict__ Xp, double* __restrict__ Lp, int64_t ldx, int* __restrict__ nonpos) {
  const int64_t j = blockIdx.y;
  bool bad = false;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < ldx; i += static_cast<int64_t>(gridDim.x) * 256) {
    const double x = i < n ? prep(A[j * lda + i], eps, use_abs) : 0.0;
    Xp[j * ldx + i] = x;
    if (Lp) {
      Lp[j * ldx + i] = i < n ? log(x) : 0.0;
      bad |= i < n && !(x > 0.0 && x < INFINITY);
    }
  }
  if (bad) atomicOr(nonpos, 1);
}

// d = f(x_i, x_j) - z0 from the staged operands: `ai` holds x'_i (log x'_i for kRaiseFast), `aj` holds x'_j
template <int F>
__device__ __forceinline__ double pair_feature(double ai, double aj) {
  if constexpr (F == GBM_F2_MULT) return __dmul_rn(ai, aj);
  else if constexpr (F == GBM_F2_ADDNORM) return __dadd_rn(ai, aj) * 0.5;
  else if constexpr (F == kRaiseFast) return exp(aj * ai);
  else return pow(ai, aj);
}
template <int F>
__device__ __forceinline__ double pair_shifted(double ai, double aj, double z0) {
  if constexpr (F == GBM_F2_MULT) return fma(ai, aj, -z0);  // one rounding instead of two
  else return pair_feature<F>(ai, aj) - z0;
}

// Xi: operand matrix of the i side (X', or log X' for kRaiseFast); Xj: X'.  Both n x l, pitch ldx.
// Register tile per thread: 4 (i) x TJ (j) pairs; 64 / TJ lanes run along j, so TJ = 2 is 512 threads
// (16 warps / SM, ~110 registers) and TJ = 4 is 256 threads (8 warps / SM, ~230 registers).
template <int F, int TJ>
__global__ void __launch_bounds__(1024 / TJ, 1)
    transform2_scan_kernel(const double* __restrict__ Xi, const double* __restrict__ Xj, int64_t n, int64_t l,
                           int64_t ldx, const double* __restrict__ yc, double ybar,
                           const double* __restrict__ colvar, double var_thr, int commutative,
                           int64_t row0, int64_t row1, double* __restrict__ beta) {
  constexpr int NJ = kTile / TJ;        // threads along j (16 or 32)
  constexpr int THREADS = 16 * NJ;      // 16 threads along i
  constexpr int LOADS = 2 * kTile * kRows / THREADS;  // cp.async per thread and stage (ai + aj)
  // rows [row0, row1) of the pair matrix (a marker shard of a multi-GPU screen); beta is that slab, (row1 - row0) x l
  const int bi = blockIdx.y + static_cast<int>(row0 / kTile), bj = blockIdx.x;
  if (commutative && bj < bi) return;  // every pair of the tile has j < i (transformation.jl:373)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2Stage* stage = reinterpret_cast<T2Stage*>(smem_raw);
  const int tid = threadIdx.x;
  const int tj = tid % NJ, ti = tid / NJ;  // lanes run along j: coalesced beta stores, conflict-free aj reads
  const int64_t i0 = static_cast<int64_t>(bi) * kTile, j0 = static_cast<int64_t>(bj) * kTile;
  // loci beyond l are clamped for loading (finite values, never stored)
  auto icol = [&](int c) { return min(i0 + c, l - 1); };
  auto jcol = [&](int c) { return min(j0 + c, l - 1); };

  auto issue = [&](int s, int64_t r0) {
    T2Stage& st = stage[s];
    const int row = tid & 31;
    if (r0 + row < n) {
#pragma unroll
      for (int k = 0; k < LOADS / 2; ++k) {
        const int c = (tid >> 5) + (THREADS / 32) * k;
        cp_async8(&st.ai[c * kPitch + row], Xi + icol(c) * ldx + r0 + row);
        cp_async8(&st.aj[c * kPitch + row], Xj + jcol(c) * ldx + r0 + row);
      }
      if (tid < 32) cp_async8(&st.y[row], yc + r0 + row);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double z0[4][TJ], s1[4][TJ], s2[4][TJ], sy[4][TJ];
  {
    double xi[4], xj[TJ];
#pragma unroll
    for (int a = 0; a < 4; ++a) xi[a] = Xi[icol(ti + 16 * a) * ldx];
#pragma unroll
    for (int b = 0; b < TJ; ++b) xj[b] = Xj[jcol(tj + NJ * b) * ldx];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < TJ; ++b) {
        z0[a][b] = pair_feature<F>(xi[a], xj[b]);
        s1[a][b] = s2[a][b] = sy[a][b] = 0.0;
      }
  }

  const int64_t nchunks = (n + kRows - 1) / kRows;
  issue(0, 0);
  for (int64_t c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue(static_cast<int>((c + 1) & 1), (c + 1) * kRows);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const T2Stage& st = stage[c & 1];
    const int rmax = static_cast<int>(min(static_cast<int64_t>(kRows), n - c * kRows));
    auto row_step = [&](int r) {
      double xi[4], xj[TJ];
      const double y = st.y[r];
#pragma unroll
      for (int a = 0; a < 4; ++a) xi[a] = st.ai[(ti + 16 * a) * kPitch + r];
#pragma unroll
      for (int b = 0; b < TJ; ++b) xj[b] = st.aj[(tj + NJ * b) * kPitch + r];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < TJ; ++b) {
          const double d = pair_shifted<F>(xi[a], xj[b], z0[a][b]);
          s1[a][b] += d;
          s2[a][b] = fma(d, d, s2[a][b]);
          sy[a][b] = fma(d, y, sy[a][b]);
        }
    };
    if (rmax == kRows && F <= GBM_F2_ADDNORM) {  // exp / pow bodies are long enough on their own
#pragma unroll 4
      for (int r = 0; r < kRows; ++r) row_step(r);
    } else {
      for (int r = 0; r < rmax; ++r) row_step(r);
    }
    __syncthreads();
  }

  const double dn = static_cast<double>(n);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t i = i0 + ti + 16 * a;
    if (i >= l || i < row0 || i >= row1 || colvar[i] < var_thr) continue;
#pragma unroll
    for (int b = 0; b < TJ; ++b) {
      const int64_t j = j0 + tj + NJ * b;
      if (j >= l || colvar[j] < var_thr || (commutative && j < i)) continue;
      beta[(i - row0) * l + j] = slope_from_sums(dn, z0[a][b], s1[a][b], s2[a][b], sy[a][b], ybar);
    }
  }
}

// ---- materialise selected features ---------------------------------------------------------
__device__ __forceinline__ double clean01(double t, double eps) {  // transformation.jl:223-227
  if (fabs(t) < eps) return 0.0;
  if (fabs(t - 1.0) < eps) return 1.0;
  return t;
}

template <int F>
__global__ void __launch_bounds__(256)
    transform1_apply_kernel(const double* __restrict__ A, int64_t n, int64_t lda, const int64_t* __restrict__ idx,
                            double eps, int use_abs, double* __restrict__ T, int64_t ldt) {
  const int64_t k = blockIdx.y;
  const double* col = A + (idx[k] - 1) * lda;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256)
    T[k * ldt + i] = clean01(f1<F>(prep(col[i], eps, use_abs)), eps);
}

template <int F>
__global__ void __launch_bounds__(256)
    transform2_apply_kernel(const double* __restrict__ A, int64_t n, int64_t l, int64_t lda,
                            const int64_t* __restrict__ counters, double eps, int use_abs, double* __restrict__ T,
                            int64_t ldt) {
  const int64_t k = blockIdx.y;
  const int64_t c0 = counters[k] - 1;  // counter = (i - 1) l + j, one-based (transformation.jl:445-446)
  const double* ci = A + (c0 / l) * lda;
  const double* cj = A + (c0 % l) * lda;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256)
    T[k * ldt + i] = clean01(f2<F>(prep(ci[i], eps, use_abs), prep(cj[i], eps, use_abs)), eps);
}

// ---- selection -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) abs_keys_kernel(const double* __restrict__ beta, int64_t len,
                                                       unsigned long long* __restrict__ keys,
                                                       long long* __restrict__ vals) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < len; i += static_cast<int64_t>(gridDim.x) * 256) {
    keys[i] = static_cast<unsigned long long>(__double_as_longlong(beta[i])) & 0x7FFFFFFFFFFFFFFFull;
    vals[i] = i + 1;
  }
}

template <typename T>
struct Scratch {
  T* p = nullptr;
  cudaStream_t s;
  Scratch(size_t count, cudaStream_t stream) : s(stream) {
    if (count) GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream));
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
};

int grid_for(int64_t work_items, int per_block, int sm_count) {
  const int64_t need = (work_items + per_block - 1) / per_block;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(need, static_cast<int64_t>(sm_count) * 8)));
}

}  // namespace

void launch_transform1_scan(int f, const double* A, int64_t n, int64_t p, int64_t lda, const double* yc, double ybar,
                            double eps, int use_abs, double var_thr, double* beta, double* colvar, int sm_count,
                            cudaStream_t stream) {
  if (p <= 0) return;
  const int grid = grid_for(p, 8, sm_count);
  switch (f) {
    case -1: transform1_scan_kernel<-1><<<grid, 256, 0, stream>>>(A, n, p, lda, yc, ybar, eps, use_abs, var_thr, beta, colvar); break;
    case GBM_F1_SQUARE: transform1_scan_kernel<GBM_F1_SQUARE><<<grid, 256, 0, stream>>>(A, n, p, lda, yc, ybar, eps, use_abs, var_thr, beta, colvar); break;
    case GBM_F1_INVONEPLUS: transform1_scan_kernel<GBM_F1_INVONEPLUS><<<grid, 256, 0, stream>>>(A, n, p, lda, yc, ybar, eps, use_abs, var_thr, beta, colvar); break;
    case GBM_F1_LOG10EPS: transform1_scan_kernel<GBM_F1_LOG10EPS><<<grid, 256, 0, stream>>>(A, n, p, lda, yc, ybar, eps, use_abs, var_thr, beta, colvar); break;
    default: GBM_THROW(GBM_ERR_ARGUMENT, "unknown one-argument transformation");
  }
  GBM_CUDA(cudaGetLastError());
}

template <int F, int TJ>
static void launch_t2_tj(const double* Xi, const double* Xj, int64_t n, int64_t l, int64_t ldx, const double* yc,
                         double ybar, const double* colvar, double var_thr, int commutative, int64_t row0,
                         int64_t row1, double* beta, cudaStream_t stream) {
  const size_t smem = 2 * sizeof(T2Stage);
  static bool configured = false;
  if (!configured) {
    GBM_CUDA(cudaFuncSetAttribute(transform2_scan_kernel<F, TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const unsigned nb = static_cast<unsigned>((l + kTile - 1) / kTile);
  const unsigned nby = static_cast<unsigned>((row1 + kTile - 1) / kTile - row0 / kTile);
  transform2_scan_kernel<F, TJ><<<dim3(nb, nby), 1024 / TJ, smem, stream>>>(Xi, Xj, n, l, ldx, yc, ybar, colvar, var_thr,
                                                                           commutative, row0, row1, beta);
}
template <int F>
static void launch_t2(const double* Xi, const double* Xj, int64_t n, int64_t l, int64_t ldx, const double* yc,
                      double ybar, const double* colvar, double var_thr, int commutative, int64_t row0, int64_t row1,
                      double* beta, cudaStream_t stream) {
  // measured (n = 10,000, l = 8,192): the 4 x 4 tile (256 threads) is as fast or faster for the short bodies
  // (mult 189 ms, addnorm 215 vs 218 ms), the 4 x 2 tile (512 threads, 16 warps / SM) for exp (975 vs 1093 ms)
  static const int tj = [] {
    const char* e = getenv("GBM_T2_TJ");  // measurement switch
    if (e && (atoi(e) == 2 || atoi(e) == 4)) return atoi(e);
    return F >= GBM_F2_RAISE ? 2 : 4;
  }();
  if (tj == 4)
    launch_t2_tj<F, 4>(Xi, Xj, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream);
  else
    launch_t2_tj<F, 2>(Xi, Xj, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream);
}

void launch_transform2_scan(int f, const double* A, int64_t n, int64_t l, int64_t lda, const double* yc, double ybar,
                            const double* colvar, double eps, int use_abs, double var_thr, int commutative,
                            int64_t row0, int64_t row1, double* beta, cudaStream_t stream) {
  if (l <= 0 || row1 <= row0) return;
  if ((l + kTile - 1) / kTile > 65535) GBM_THROW(GBM_ERR_ARGUMENT, "transform2: too many loci for one pairwise screen");
  if (f < GBM_F2_MULT || f > GBM_F2_RAISE) GBM_THROW(GBM_ERR_ARGUMENT, "unknown two-argument transformation");
  // X' (and log X' for raise) once, so that the pair kernel only accumulates
  const int64_t ldx = (n + 1) / 2 * 2;
  const bool want_log = f == GBM_F2_RAISE;
  Scratch<double> Xp(static_cast<size_t>(ldx) * l, stream), Lp(want_log ? static_cast<size_t>(ldx) * l : 0, stream);
  Scratch<int> nonpos(1, stream);
  GBM_CUDA(cudaMemsetAsync(nonpos.p, 0, sizeof(int), stream));
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>(32, (ldx + 255) / 256));
  for (int64_t c0 = 0; c0 < l; c0 += 65535) {
    const unsigned gy = static_cast<unsigned>(std::min<int64_t>(65535, l - c0));
    transform_prep_kernel<<<dim3(gx, gy), 256, 0, stream>>>(A + c0 * lda, n, lda, eps, use_abs, Xp.p + c0 * ldx,
                                                            want_log ? Lp.p + c0 * ldx : nullptr, ldx, nonpos.p);
  }
  GBM_CUDA(cudaGetLastError());
  int h_nonpos = 0;
  if (want_log) {
    GBM_CUDA(cudaMemcpyAsync(&h_nonpos, nonpos.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    GBM_CUDA(cudaStreamSynchronize(stream));
  }
  switch (f) {
    case GBM_F2_MULT: launch_t2<GBM_F2_MULT>(Xp.p, Xp.p, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream); break;
    case GBM_F2_ADDNORM: launch_t2<GBM_F2_ADDNORM>(Xp.p, Xp.p, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream); break;
    default:
      if (h_nonpos)  // zero / negative / non-finite bases: pow() itself decides (NaN where Julia throws DomainError)
        launch_t2<GBM_F2_RAISE>(Xp.p, Xp.p, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream);
      else
        launch_t2<kRaiseFast>(Lp.p, Xp.p, n, l, ldx, yc, ybar, colvar, var_thr, commutative, row0, row1, beta, stream);
  }
  GBM_CUDA(cudaGetLastError());
}

void launch_transform1_apply(int f, const double* A, int64_t n, int64_t lda, const int64_t* idx, int64_t count,
                             double eps, int use_abs, double* T, int64_t ldt, cudaStream_t stream) {
  if (count <= 0) return;
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>(64, (n + 255) / 256));
  for (int64_t k0 = 0; k0 < count; k0 += 65535) {
    const dim3 grid(gx, static_cast<unsigned>(std::min<int64_t>(65535, count - k0)));
    switch (f) {
      case GBM_F1_SQUARE: transform1_apply_kernel<GBM_F1_SQUARE><<<grid, 256, 0, stream>>>(A, n, lda, idx + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      case GBM_F1_INVONEPLUS: transform1_apply_kernel<GBM_F1_INVONEPLUS><<<grid, 256, 0, stream>>>(A, n, lda, idx + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      case GBM_F1_LOG10EPS: transform1_apply_kernel<GBM_F1_LOG10EPS><<<grid, 256, 0, stream>>>(A, n, lda, idx + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      default: GBM_THROW(GBM_ERR_ARGUMENT, "unknown one-argument transformation");
    }
  }
  GBM_CUDA(cudaGetLastError());
}

void launch_transform2_apply(int f, const double* A, int64_t n, int64_t l, int64_t lda, const int64_t* counters,
                             int64_t count, double eps, int use_abs, double* T, int64_t ldt, cudaStream_t stream) {
  if (count <= 0) return;
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>(64, (n + 255) / 256));
  for (int64_t k0 = 0; k0 < count; k0 += 65535) {
    const dim3 grid(gx, static_cast<unsigned>(std::min<int64_t>(65535, count - k0)));
    switch (f) {
      case GBM_F2_MULT: transform2_apply_kernel<GBM_F2_MULT><<<grid, 256, 0, stream>>>(A, n, l, lda, counters + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      case GBM_F2_ADDNORM: transform2_apply_kernel<GBM_F2_ADDNORM><<<grid, 256, 0, stream>>>(A, n, l, lda, counters + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      case GBM_F2_RAISE: transform2_apply_kernel<GBM_F2_RAISE><<<grid, 256, 0, stream>>>(A, n, l, lda, counters + k0, eps, use_abs, T + k0 * ldt, ldt); break;
      default: GBM_THROW(GBM_ERR_ARGUMENT, "unknown two-argument transformation");
    }
  }
  GBM_CUDA(cudaGetLastError());
}

// sortperm(abs.(beta), rev = true)[1:n_new] (stable: equal |beta| keep ascending index order, like
// Julia's default), then the leading entries with abs(beta) > eps.  idx_host gets 1-based positions.
int64_t transform_select(const double* beta_dev, int64_t len, int64_t n_new, double eps, int64_t* idx_host,
                         double* beta_host, bool* has_nan, int sm_count, cudaStream_t stream) {
  if (has_nan) *has_nan = false;
  if (len <= 0 || n_new <= 0) return 0;
  {  // the full stable sort needs four key / index arrays of `len` entries plus CUB's scratch (~ another two)
    size_t free_b = 0, total_b = 0;
    GBM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double need = 48.0 * static_cast<double>(len);
    if (need > 0.9 * static_cast<double>(free_b))
      GBM_THROW(1, "transform screen: selecting the top effects of " + std::to_string(len) + " needs about " +
                       std::to_string(static_cast<long long>(need / 1e9) + 1) + " GB of device scratch and " +
                       std::to_string(static_cast<long long>(free_b / 1e9)) + " GB are free; screen the pair matrix in row "
                       "blocks (gbm_transform2_screen_rows) or fewer loci at a time");
  }
  Scratch<unsigned long long> k_in(len, stream), k_out(len, stream);
  Scratch<long long> v_in(len, stream), v_out(len, stream);
  abs_keys_kernel<<<grid_for(len, 256, sm_count), 256, 0, stream>>>(beta_dev, len, k_in.p, v_in.p);
  GBM_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  GBM_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, len, 0, 63, stream));
  Scratch<unsigned char> tmp(tmp_bytes, stream);
  GBM_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, len, 0, 63, stream));
  const int64_t take = std::min(n_new, len);
  std::vector<unsigned long long> hk(take);
  std::vector<long long> hv(take);
  GBM_CUDA(cudaMemcpyAsync(hk.data(), k_out.p, sizeof(unsigned long long) * take, cudaMemcpyDeviceToHost, stream));
  GBM_CUDA(cudaMemcpyAsync(hv.data(), v_out.p, sizeof(long long) * take, cudaMemcpyDeviceToHost, stream));
  GBM_CUDA(cudaStreamSynchronize(stream));
  // "for j in idx_sorted: if abs(beta[j]) > eps: append" -- the order is descending, so this is a prefix,
  // except that NaN sorts first and fails the comparison: filter element-wise like the reference
  int64_t count = 0;
  for (int64_t k = 0; k < take; ++k) {
    double a;
    memcpy(&a, &hk[k], sizeof(double));
    if (a != a && has_nan) *has_nan = true;
    if (a > eps) idx_host[count++] = hv[k];
  }
  if (beta_host && count > 0) {
    // signed values of the selected entries
    Scratch<long long> d_idx(count, stream);
    Scratch<double> d_val(count, stream);
    GBM_CUDA(cudaMemcpyAsync(d_idx.p, idx_host, sizeof(long long) * count, cudaMemcpyHostToDevice, stream));
    launch_gather_values(beta_dev, d_idx.p, count, d_val.p, stream);
    GBM_CUDA(cudaMemcpyAsync(beta_host, d_val.p, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    GBM_CUDA(cudaStreamSynchronize(stream));
  }
  return count;
}

__global__ void gather_values_kernel(const double* __restrict__ src, const long long* __restrict__ idx1, int64_t count,
                                     double* __restrict__ dst) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < count) dst[k] = src[idx1[k] - 1];
}
void launch_gather_values(const double* src, const long long* idx1, int64_t count, double* dst, cudaStream_t stream) {
  if (count <= 0) return;
  gather_values_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, stream>>>(src, idx1, count, dst);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
