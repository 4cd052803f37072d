// Largest eigenpair of a symmetric positive semi-definite matrix by Lanczos with full
// reorthogonalisation -- the PC1 step of gwasols / gwaslmm (/root/reference/src/gwas.jl:234, :357:
// `fit(PCA, GRM; maxoutdim = 1)`, `E.proj[:, 1]`), which needs ONE eigenvector of B = Z Z', not the
// n x n decomposition cuSOLVER's syevdx computes on the way (tridiagonalisation of the whole matrix:
// 0.71 s at n = 10,000 against ~0.1 s here).  gbm_kstd_pc1 falls back to cuSOLVER when this does not
// converge, and GBM_PC1_SOLVER=cusolver forces it.
//
// Per step: w = B v_j (one pass over B, HBM-bound: 8 n^2 bytes), classical Gram-Schmidt twice against
// all previous vectors (alpha_j falls out of the projections), beta_j = ||w||, v_{j+1} = w / beta_j.
// alpha / beta stay on the device; every few steps the host solves the small tridiagonal problem
// (bisection + inverse iteration) and tests the residual estimate beta_j |s_j| <= tol theta.  All
// reductions have a fixed order: the result is a deterministic function of B.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/gbm_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace gbm {

namespace {

// the same for any block size up to 1024 threads (sh: 32 doubles)
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += sh[w];  // fixed order
  return t;  // valid in thread 0
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];  // fixed order
  return t;  // valid in thread 0
}

// y[j] = B[:, j] . v for j < ncols  (B: n rows, column-major, pitch ld; for a symmetric B this is B v): one warp
// per column
__global__ void __launch_bounds__(256) symv_kernel(const double* __restrict__ B, int64_t n, int64_t ncols, int64_t ld,
                                                   const double* __restrict__ v, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  const int64_t n2 = n >> 1;
  const double2* __restrict__ v2 = reinterpret_cast<const double2*>(v);
  for (int64_t j = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5; j < ncols; j += nwarps) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(B + j * ld);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
    for (int64_t i = lane; i < n2; i += 32) {
      const double2 a = c2[i], b = v2[i];
      s0 = fma(a.x, b.x, s0);
      s1 = fma(a.y, b.y, s1);
    }
    double s = s0 + s1;
    if ((n & 1) && lane == 0) s = fma(B[j * ld + n - 1], v[n - 1], s);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) y[j] = s;
  }
}

// The same with one CTA per column, for blocks with few columns (a rank's share of a sharded matrix): a single warp
// walking a whole 80 KB column is a ~40 us chain of dependent HBM latencies however few columns there are; 256
// threads cut it to a few rounds.  Fixed-order block reduction.
__global__ void __launch_bounds__(256) symv_cta_kernel(const double* __restrict__ B, int64_t n, int64_t ld,
                                                       const double* __restrict__ v, double* __restrict__ y) {
  __shared__ double sh[8];
  const int64_t j = blockIdx.x;
  const int64_t n2 = n >> 1;
  const double2* __restrict__ c2 = reinterpret_cast<const double2*>(B + j * ld);
  const double2* __restrict__ v2 = reinterpret_cast<const double2*>(v);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t i = threadIdx.x;
  for (; i + 256 < n2; i += 512) {
    const double2 a = c2[i], b = v2[i], c = c2[i + 256], d = v2[i + 256];
    s0 = fma(a.x, b.x, s0);
    s1 = fma(a.y, b.y, s1);
    s2 = fma(c.x, d.x, s2);
    s3 = fma(c.y, d.y, s3);
  }
  for (; i < n2; i += 256) {
    const double2 a = c2[i], b = v2[i];
    s0 = fma(a.x, b.x, s0);
    s1 = fma(a.y, b.y, s1);
  }
  double s = (s0 + s1) + (s2 + s3);
  if ((n & 1) && threadIdx.x == 0) s = fma(B[j * ld + n - 1], v[n - 1], s);
  const double t = block_sum_256(s, sh);
  if (threadIdx.x == 0) y[j] = t;
}

// partial[c][i] = sum over the c-th chunk of columns of Z[i, j] u[j]   (Z column-major, pitch ld): thread = row,
// so every load is coalesced; the chunks are summed in a fixed order by partial_reduce_kernel (deterministic)
constexpr int kGemvChunks = 64;
__global__ void __launch_bounds__(256) gemv_n_partial_kernel(const double* __restrict__ Z, int64_t n, int64_t ncols,
                                                             int64_t ld, const double* __restrict__ u,
                                                             double* __restrict__ partial) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t cw = (ncols + kGemvChunks - 1) / kGemvChunks;
  const int64_t j0 = min(ncols, static_cast<int64_t>(blockIdx.y) * cw), j1 = min(ncols, j0 + cw);
  if (i >= n) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int64_t j = j0;
  for (; j + 4 <= j1; j += 4) {
    a0 = fma(Z[(j + 0) * ld + i], u ? u[j + 0] : 1.0, a0);
    a1 = fma(Z[(j + 1) * ld + i], u ? u[j + 1] : 1.0, a1);
    a2 = fma(Z[(j + 2) * ld + i], u ? u[j + 2] : 1.0, a2);
    a3 = fma(Z[(j + 3) * ld + i], u ? u[j + 3] : 1.0, a3);
  }
  for (; j < j1; ++j) a0 = fma(Z[j * ld + i], u ? u[j] : 1.0, a0);
  partial[static_cast<int64_t>(blockIdx.y) * n + i] = (a0 + a1) + (a2 + a3);
}

// ---- fused gram step: partial[cta][:] = sum over the CTA's columns j of Z[:, j] (Z[:, j] . v) ----------------------
// One pass over Z per Lanczos step instead of two (u = Z'v, then w = Z u): a CTA takes whole columns; a column is
// brought into shared memory by one bulk copy (TMA, double-buffered, so the next column streams in while this one
// is used), the 512 threads take its dot product with v (each thread keeps ITS rows of v in registers for the whole
// kernel), and the column is added, scaled by that dot, to the thread's rows of the CTA's partial result (registers
// as well).  The partials of all CTAs are summed in a fixed order by partial_reduce_kernel: deterministic.
// Needs 2 n doubles of shared memory: n <= 12,288; larger n use the two-pass kernels.
constexpr int kFusedThreads = 512;
template <int RPT>  // rows per thread: n <= 512 * RPT
__global__ void __launch_bounds__(kFusedThreads, 1)
    gram_fused_kernel(const double* __restrict__ Z, int64_t n, int64_t ncols, int64_t ld, const double* __restrict__ v,
                      double* __restrict__ partial) {
  extern __shared__ __align__(128) uint8_t fsm[];
  const int64_t npad = (n + 1) / 2 * 2;
  double* buf0 = reinterpret_cast<double*>(fsm);
  double* buf1 = buf0 + npad;
  uint64_t* bar = reinterpret_cast<uint64_t*>(buf1 + npad);  // [2]
  double* red = reinterpret_cast<double*>(bar + 2);           // [16] warp partials
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  double vr[RPT], acc[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int64_t i = tid + static_cast<int64_t>(k) * kFusedThreads;
    vr[k] = i < n ? v[i] : 0.0;
    acc[k] = 0.0;
  }
  const uint32_t bytes = static_cast<uint32_t>(n * sizeof(double));
  int64_t j = blockIdx.x;
  if (tid == 0 && j < ncols) {
    mbar_arrive_expect_tx(&bar[0], bytes);
    tma_load_1d(buf0, Z + j * ld, bytes, &bar[0]);
  }
  uint32_t ph[2] = {0, 0};
  int b = 0;
  for (; j < ncols; j += gridDim.x, b ^= 1) {
    const int64_t jn = j + gridDim.x;
    if (tid == 0 && jn < ncols) {  // the other buffer was released by the __syncthreads that ended its column
      mbar_arrive_expect_tx(&bar[b ^ 1], bytes);
      tma_load_1d(b ? buf0 : buf1, Z + jn * ld, bytes, &bar[b ^ 1]);
    }
    mbar_wait(&bar[b], ph[b]);
    ph[b] ^= 1u;
    const double* col = b ? buf1 : buf0;
    double d0 = 0.0, d1 = 0.0;  // the column is read from shared memory twice (dot, then update): registers hold v and acc
#pragma unroll
    for (int k = 0; k < RPT; k += 2) {
      const int64_t i0 = tid + static_cast<int64_t>(k) * kFusedThreads, i1 = i0 + kFusedThreads;
      d0 = fma(i0 < n ? col[i0] : 0.0, vr[k], d0);
      d1 = fma(i1 < n ? col[i1] : 0.0, vr[k + 1], d1);
    }
    double d = d0 + d1;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
    if (lane == 0) red[warp] = d;
    __syncthreads();
    double u = 0.0;
#pragma unroll
    for (int w = 0; w < kFusedThreads / 32; ++w) u += red[w];  // fixed order, the same in every thread
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int64_t i = tid + static_cast<int64_t>(k) * kFusedThreads;
      acc[k] = fma(i < n ? col[i] : 0.0, u, acc[k]);
    }
    __syncthreads();  // red and this column's buffer may be reused
  }
  double* out = partial + static_cast<int64_t>(blockIdx.x) * n;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int64_t i = tid + static_cast<int64_t>(k) * kFusedThreads;
    if (i < n) out[i] = acc[k];
  }
}

// w[i] = sum_c partial[c][i] over `chunks` partial vectors.  A CTA owns 32 rows; its 8 warps take the chunks c = g,
// g + 8, ... (every load a coalesced 256-byte row segment, four independent chains per thread), the 8 group sums are
// added in a fixed order.  One thread per row walking all the chunks was a chain of up to 148 dependent L2 latencies
// (~25 us per Lanczos step); this is ~5.
__device__ __forceinline__ double chunk_group_sum(const double* __restrict__ partial, int64_t n, int chunks, int64_t i, int g) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int c = g;
  for (; c + 24 < chunks; c += 32) {
    s0 += partial[static_cast<int64_t>(c) * n + i];
    s1 += partial[static_cast<int64_t>(c + 8) * n + i];
    s2 += partial[static_cast<int64_t>(c + 16) * n + i];
    s3 += partial[static_cast<int64_t>(c + 24) * n + i];
  }
  for (; c < chunks; c += 8) s0 += partial[static_cast<int64_t>(c) * n + i];
  return (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256) partial_reduce_kernel(const double* __restrict__ partial, int64_t n, int chunks,
                                                             double* __restrict__ w) {
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  part[g][lane] = i < n ? chunk_group_sum(partial, n, chunks, i, g) : 0.0;
  __syncthreads();
  if (g == 0 && i < n) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][lane];
    w[i] = t;
  }
}
static void launch_partial_reduce(const double* partial, int64_t n, int chunks, double* w, cudaStream_t stream) {
  partial_reduce_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(partial, n, chunks, w);
}

// ---- partial-sum reduction fused with the all-reduce over peer memory (NVLink P2P stores) --------------------------
__device__ __forceinline__ void st_volatile_v2(double* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_volatile_v2(const double* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
// One CTA per 32 rows (the layout of partial_reduce_kernel, the same bits for the rank's own sum).  Every CTA waits
// inside the kernel for the other GPUs, so the whole grid must be resident at once: n <= kPeerSumMaxN.
constexpr int64_t kPeerSumMaxN = 32 * 8 * 100;  // 800 CTAs of 256 threads: fewer than the 8 per SM a B200 holds
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerMailbox mb, const double* __restrict__ partial, int chunks,
                                                       int64_t n, unsigned long long step, double* __restrict__ out) {
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  const int W = mb.world;
  const int64_t par = static_cast<int64_t>(step & 1ull) * W;
  const unsigned long long stamp = (step & 0xffffffffull) << 32;
  part[g][lane] = i < n ? chunk_group_sum(partial, n, chunks, i, g) : 0.0;
  __syncthreads();
  if (g != 0 || i >= n) return;
  double v = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) v += part[q][lane];
  {  // this rank's rows into slot `me` of every mailbox (q == me: the local one), data and stamp in the same words
    const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
    const unsigned long long lo = stamp | (u & 0xffffffffull), hi = stamp | (u >> 32);
    const int64_t at = 2 * ((par + mb.me) * mb.npad + i);
#pragma unroll
    for (int q = 0; q < PeerMailbox::kMaxRanks; ++q)
      if (q < W) st_volatile_v2(mb.slots[q] + at, lo, hi);
  }
  // every rank's contribution to THIS mailbox: poll until all W entries of the row carry the stamp (bounded: a dead
  // peer must not hang the GPU)
  const double* mine = mb.mine + 2 * (par * mb.npad + i);
  ulonglong2 x[PeerMailbox::kMaxRanks];
  volatile int* err = mb.error;
  const long long t0 = clock64();
  for (unsigned spin = 0;; ++spin) {
#pragma unroll
    for (int r = 0; r < PeerMailbox::kMaxRanks; ++r)  // all W loads in flight before any of them is examined
      if (r < W) x[r] = ld_volatile_v2(mine + 2 * r * mb.npad);
    bool all = true;
#pragma unroll
    for (int r = 0; r < PeerMailbox::kMaxRanks; ++r)
      if (r < W) all &= ((x[r].x & 0xffffffff00000000ull) == stamp) & ((x[r].y & 0xffffffff00000000ull) == stamp);
    if (all) break;
    if ((spin & 255u) == 255u) {
      if (*err) break;                        // an earlier step already gave up: do not wait 10 s per step
      if (clock64() - t0 > 20000000000ll) {  // ~10 s
        *err = 1;
        break;
      }
    }
  }
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < PeerMailbox::kMaxRanks; ++r)  // rank order
    if (r < W) t += __longlong_as_double(static_cast<long long>((x[r].x & 0xffffffffull) | (x[r].y << 32)));
  out[i] = t;
}

constexpr int64_t kFusedMaxN = kLanczosFusedMaxN;
static_assert(kFusedMaxN == 512 * 24, "gram_fused_kernel<24>: 24 rows per thread");
static_assert(kFusedMaxN <= kPeerSumMaxN, "the peer exchange must take every n the fused step takes");
static size_t gram_fused_smem(int64_t n) { return static_cast<size_t>((n + 1) / 2 * 2) * 16 + 16 + 16 * 8 + 128; }

// launches the fused step on `grid` CTAs; returns false when n is too large for it
static bool launch_gram_fused(const double* Z, int64_t n, int64_t ncols, int64_t ld, const double* v, double* partial,
                              int grid, cudaStream_t stream) {
  if (n > kFusedMaxN || (ld & 1) || (n & 1)) return false;  // bulk copies need 16-byte multiples
  const size_t smem = gram_fused_smem(n);
  auto go = [&](auto kern) {
    GBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kFusedThreads, smem, stream>>>(Z, n, ncols, ld, v, partial);
  };
  if (n <= 512 * 8) go(gram_fused_kernel<8>);
  else if (n <= 512 * 16) go(gram_fused_kernel<16>);
  else go(gram_fused_kernel<24>);
  return true;
}

// c[k] = V[:, k] . w for k < cols: one CTA per column, fixed-order reduction (four independent FMA chains per thread:
// the loop is latency-bound, not bandwidth-bound)
__global__ void __launch_bounds__(1024) dots_kernel(const double* __restrict__ V, int64_t n, int64_t ldv,
                                                    const double* __restrict__ w, double* __restrict__ c) {
  __shared__ double sh[32];
  const double* col = V + static_cast<int64_t>(blockIdx.x) * ldv;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t i = threadIdx.x;
  for (; i + 3072 < n; i += 4096) {
    s0 = fma(col[i], w[i], s0);
    s1 = fma(col[i + 1024], w[i + 1024], s1);
    s2 = fma(col[i + 2048], w[i + 2048], s2);
    s3 = fma(col[i + 3072], w[i + 3072], s3);
  }
  for (; i < n; i += 1024) s0 = fma(col[i], w[i], s0);
  const double t = block_sum((s0 + s1) + (s2 + s3), sh);
  if (threadIdx.x == 0) c[blockIdx.x] = t;
}

// w -= V[:, 0..cols) c ; alpha[j] (+)= c[j] (the projection on the current vector is the Lanczos alpha).
// A CTA owns 32 rows; its 8 warps split the columns (warp g takes k = g, g + 8, ...: every load is a coalesced
// 256-byte row segment), so a thread walks cols / 8 columns instead of all of them -- with n = 10,000 rows there are
// only 10,000 row-threads, and one thread per row walking 280 columns was a 40 us dependent-latency chain per call.
// The 8 partial sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256) project_out_kernel(const double* __restrict__ V, int64_t n, int64_t ldv, int cols,
                                                          const double* __restrict__ c, double* __restrict__ w,
                                                          double* __restrict__ alpha, int j, int accumulate) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (i < n) {
    int k = g;
    for (; k + 24 < cols; k += 32) {
      a0 = fma(V[(k + 0) * ldv + i], c[k + 0], a0);
      a1 = fma(V[(k + 8) * ldv + i], c[k + 8], a1);
      a2 = fma(V[(k + 16) * ldv + i], c[k + 16], a2);
      a3 = fma(V[(k + 24) * ldv + i], c[k + 24], a3);
    }
    for (; k < cols; k += 8) a0 = fma(V[k * ldv + i], c[k], a0);
  }
  part[g][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (g == 0 && i < n) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][lane];
    w[i] -= t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) alpha[j] = accumulate ? alpha[j] + c[j] : c[j];
}

// beta[j] = ||w||, V[:, j+1] = w / beta[j]   (single CTA)
__global__ void __launch_bounds__(1024) norm_next_kernel(const double* __restrict__ w, int64_t n, double* __restrict__ vnext,
                                                         double* __restrict__ beta, int j) {
  __shared__ double sh[32];
  __shared__ double inv;
  double s0 = 0.0, s1 = 0.0;
  int64_t i = threadIdx.x;
  for (; i + 1024 < n; i += 2048) {
    s0 = fma(w[i], w[i], s0);
    s1 = fma(w[i + 1024], w[i + 1024], s1);
  }
  for (; i < n; i += 1024) s0 = fma(w[i], w[i], s0);
  const double t = block_sum(s0 + s1, sh);
  if (threadIdx.x == 0) {
    const double b = sqrt(t);
    beta[j] = b;
    inv = b > 0.0 ? 1.0 / b : 0.0;
  }
  __syncthreads();
  for (int64_t k = threadIdx.x; k < n; k += 1024) vnext[k] = w[k] * inv;
}

// ---- the whole reorthogonalisation of one Lanczos step in ONE cooperative kernel ---------------------------------------
// Classical Gram-Schmidt twice, alpha, beta and the next basis vector were five launches (dots, project, dots, project,
// norm) of a few microseconds each -- with the matrix pass at ~0.02-0.12 ms they, and the gaps between them, were a
// third to a half of a step.  Here the five phases are separated by grid barriers instead of kernel boundaries:
//   c = V'w | w -= V c, alpha_j = c_j | c = V'w | w -= V c, alpha_j += c_j, row-group norms | beta_j, v_{j+1} = w / beta_j
// Work split: a CTA per column for the dots, a CTA per 32 rows (16 warps = 16 column groups) for the projections.
// Values produced inside the kernel are read back with ld.cg (L2), never through L1.  Every reduction has a fixed
// order, independent of the grid size: the result is the same function of (V, w) on every GPU.
namespace cg = cooperative_groups;
constexpr int kReorthThreads = 512;
__global__ void __launch_bounds__(kReorthThreads, 3) reorth_kernel(const double* __restrict__ V, int64_t n, int64_t ldv, int j,
                                                                double* __restrict__ w, double* __restrict__ c,
                                                                double* __restrict__ alpha, double* __restrict__ beta,
                                                                double* __restrict__ normpart, double* __restrict__ vnext) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[32];
  __shared__ double part[kReorthThreads / 32][33];
  __shared__ double inv_sh;
  const int cols = j + 1;
  const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;
  const int64_t ngroups = (n + 31) / 32;
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = blockIdx.x; k < cols; k += gridDim.x) {  // c[k] = V[:, k] . w
      const double* col = V + static_cast<int64_t>(k) * ldv;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int64_t i = tid;
      for (; i + 3 * kReorthThreads < n; i += 4 * kReorthThreads) {
        s0 = fma(col[i], __ldcg(w + i), s0);
        s1 = fma(col[i + kReorthThreads], __ldcg(w + i + kReorthThreads), s1);
        s2 = fma(col[i + 2 * kReorthThreads], __ldcg(w + i + 2 * kReorthThreads), s2);
        s3 = fma(col[i + 3 * kReorthThreads], __ldcg(w + i + 3 * kReorthThreads), s3);
      }
      for (; i < n; i += kReorthThreads) s0 = fma(col[i], __ldcg(w + i), s0);
      const double t = block_sum((s0 + s1) + (s2 + s3), sh);
      if (tid == 0) c[k] = t;
      __syncthreads();  // sh is reused by the next column
    }
    grid.sync();
    for (int64_t rg = blockIdx.x; rg < ngroups; rg += gridDim.x) {  // w -= V c on rows 32 rg .. 32 rg + 31
      const int64_t i = rg * 32 + lane;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (i < n) {
        constexpr int kW = kReorthThreads / 32;
        int k = wg;
        for (; k + 3 * kW < cols; k += 4 * kW) {
          a0 = fma(V[static_cast<int64_t>(k) * ldv + i], __ldcg(c + k), a0);
          a1 = fma(V[static_cast<int64_t>(k + kW) * ldv + i], __ldcg(c + k + kW), a1);
          a2 = fma(V[static_cast<int64_t>(k + 2 * kW) * ldv + i], __ldcg(c + k + 2 * kW), a2);
          a3 = fma(V[static_cast<int64_t>(k + 3 * kW) * ldv + i], __ldcg(c + k + 3 * kW), a3);
        }
        for (; k < cols; k += kW) a0 = fma(V[static_cast<int64_t>(k) * ldv + i], __ldcg(c + k), a0);
      }
      part[wg][lane] = (a0 + a1) + (a2 + a3);
      __syncthreads();
      if (wg == 0) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < kReorthThreads / 32; ++q) t += part[q][lane];  // fixed order
        double wn = 0.0;
        if (i < n) {
          wn = __ldcg(w + i) - t;
          w[i] = wn;
        }
        if (pass == 1) {
          double sq = wn * wn;
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, m);
          if (lane == 0) normpart[rg] = sq;
        }
      }
      __syncthreads();  // part is reused by the next row group
    }
    if (blockIdx.x == 0 && tid == 0) {
      const double cj = __ldcg(c + j);
      alpha[j] = pass ? alpha[j] + cj : cj;  // the projection on v_j is the Lanczos alpha
    }
    grid.sync();
  }
  // beta_j = ||w|| from the row-group sums of squares: the same order, hence the same bits, in every CTA
  double s = 0.0;
  for (int64_t g = tid; g < ngroups; g += kReorthThreads) s += __ldcg(normpart + g);
  const double t = block_sum(s, sh);
  if (tid == 0) {
    const double b = sqrt(t);
    if (blockIdx.x == 0) beta[j] = b;
    inv_sh = b > 0.0 ? 1.0 / b : 0.0;
  }
  __syncthreads();
  const double inv = inv_sh;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kReorthThreads + tid; i < n; i += static_cast<int64_t>(gridDim.x) * kReorthThreads)
    vnext[i] = __ldcg(w + i) * inv;
}

// CTAs of reorth_kernel that can be resident at once on the current device (0: no cooperative launch)
static int reorth_max_grid() {
  if (const char* e = getenv("GBM_PC1_NO_COOP"))
    if (atoi(e) != 0) return 0;
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop)
    return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reorth_kernel, kReorthThreads, 0) != cudaSuccess) return 0;
  return sms * std::min(per_sm, 3);
}

// deterministic start vector: splitmix64 of the row index, mapped to (-1, 1), normalised by norm_next_kernel
__global__ void __launch_bounds__(256) start_vector_kernel(double* __restrict__ w, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  uint64_t z = static_cast<uint64_t>(i) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  w[i] = static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// x = V[:, 0..cols) s
__global__ void __launch_bounds__(256) combine_kernel(const double* __restrict__ V, int64_t n, int64_t ldv, int cols,
                                                      const double* __restrict__ s, double* __restrict__ x) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  double a0 = 0.0, a1 = 0.0;
  int k = 0;
  for (; k + 2 <= cols; k += 2) {
    a0 = fma(V[k * ldv + i], s[k], a0);
    a1 = fma(V[(k + 1) * ldv + i], s[k + 1], a1);
  }
  if (k < cols) a0 = fma(V[k * ldv + i], s[k], a0);
  x[i] = a0 + a1;
}

// ---- host: largest eigenpair of the m x m symmetric tridiagonal (a, b) ----------------------------
int sturm_count_below(const std::vector<double>& a, const std::vector<double>& b, int m, double x) {
  int cnt = 0;
  double q = a[0] - x;
  if (q < 0) ++cnt;
  for (int i = 1; i < m; ++i) {
    if (q == 0.0) q = 1e-300;
    q = a[i] - x - b[i - 1] * b[i - 1] / q;
    if (q < 0) ++cnt;
  }
  return cnt;
}

void tridiag_top(const std::vector<double>& a, const std::vector<double>& b, int m, double* theta, std::vector<double>* s) {
  double lo = a[0], hi = a[0];
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i + 1 < m ? fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
  }
  // largest eigenvalue: the smallest x with count(x) == m is just above it
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count_below(a, b, m, mid) == m) hi = mid; else lo = mid;
  }
  *theta = 0.5 * (lo + hi);
  // inverse iteration with a pivoted tridiagonal solve
  s->assign(m, 1.0 / sqrt(static_cast<double>(m)));
  if (m == 1) {
    (*s)[0] = 1.0;
    return;
  }
  const double scale = std::max(fabs(lo), fabs(hi));
  const double shift = *theta + 4.0 * 2.220446049250313e-16 * (scale > 0 ? scale : 1.0);
  std::vector<double> d(m), du(m), du2(m), dl(m), rhs(m);
  for (int iter = 0; iter < 4; ++iter) {
    for (int i = 0; i < m; ++i) {
      d[i] = a[i] - shift;
      du[i] = i + 1 < m ? b[i] : 0.0;
      dl[i] = i + 1 < m ? b[i] : 0.0;  // dl[i] couples row i+1 to column i
      du2[i] = 0.0;
      rhs[i] = (*s)[i];
    }
    for (int i = 0; i + 1 < m; ++i) {  // elimination with partial pivoting (as LAPACK dgtsv)
      if (fabs(d[i]) >= fabs(dl[i])) {
        const double piv = d[i] != 0.0 ? d[i] : 1e-300;
        const double f = dl[i] / piv;
        d[i] = piv;
        d[i + 1] -= f * du[i];
        rhs[i + 1] -= f * rhs[i];
        du2[i] = 0.0;
      } else {
        const double f = d[i] / dl[i];
        d[i] = dl[i];
        const double t = d[i + 1];
        d[i + 1] = du[i] - f * t;
        du2[i] = i + 2 < m ? du[i + 1] : 0.0;
        if (i + 2 < m) du[i + 1] = -f * du2[i];
        du[i] = t;
        std::swap(rhs[i], rhs[i + 1]);
        rhs[i + 1] -= f * rhs[i];
      }
    }
    if (d[m - 1] == 0.0) d[m - 1] = 1e-300;
    rhs[m - 1] /= d[m - 1];
    if (m > 1) rhs[m - 2] = (rhs[m - 2] - du[m - 2] * rhs[m - 1]) / d[m - 2];
    for (int i = m - 3; i >= 0; --i) rhs[i] = (rhs[i] - du[i] * rhs[i + 1] - du2[i] * rhs[i + 2]) / d[i];
    double nrm = 0.0;
    for (int i = 0; i < m; ++i) nrm += rhs[i] * rhs[i];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0) || !std::isfinite(nrm)) break;
    for (int i = 0; i < m; ++i) (*s)[i] = rhs[i] / nrm;
  }
}

// Small pinned host buffers, reused across calls and threads (cudaMallocHost costs ~0.1 ms and synchronises): a lease
// takes one that is large enough from the free list or allocates it; buffers live until the process ends.
struct PinnedLease {
  double* p = nullptr;
  size_t count = 0;
  explicit PinnedLease(size_t n) {
    {
      std::lock_guard<std::mutex> lk(mutex());
      auto& fl = free_list();
      for (size_t i = 0; i < fl.size(); ++i)
        if (fl[i].second >= n) {
          p = fl[i].first;
          count = fl[i].second;
          fl.erase(fl.begin() + i);
          break;
        }
    }
    if (!p) {
      count = std::max<size_t>(n, 4 * 3001);
      GBM_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(double) * count, cudaHostAllocPortable));
    }
  }
  ~PinnedLease() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mutex());
    free_list().emplace_back(p, count);
  }
  PinnedLease(const PinnedLease&) = delete;
  PinnedLease& operator=(const PinnedLease&) = delete;
  static std::mutex& mutex() {
    static std::mutex m;
    return m;
  }
  static std::vector<std::pair<double*, size_t>>& free_list() {
    static std::vector<std::pair<double*, size_t>> v;
    return v;
  }
};

// Lanczos with full reorthogonalisation on the operator `apply(v, out)` (out = Op v, both device n-vectors, launched
// on `stream`); Op symmetric positive semi-definite.
template <typename Apply>
static bool lanczos_core(int64_t n, Apply&& apply, double tol, int max_iter, double* x_dev, double* theta_out,
                         int* iters_out, cudaStream_t stream) {
  const int m_max = static_cast<int>(std::min<int64_t>(max_iter, n - 1));
  if (m_max < 2) return false;
  const int64_t ldv = (n + 1) / 2 * 2;
  double *V = nullptr, *w = nullptr, *c = nullptr, *alpha = nullptr, *beta = nullptr, *sdev = nullptr, *normpart = nullptr;
  PinnedLease pin(static_cast<size_t>(4) * (m_max + 1));  // two slots of (alpha, beta) for the asynchronous checks
  cudaEvent_t ev[2] = {nullptr, nullptr};
  const int64_t ngroups = (n + 31) / 32;
  const int coop_grid = static_cast<int>(std::min<int64_t>(reorth_max_grid(), ngroups));
  auto release = [&] {
    for (double* p : {V, w, c, alpha, beta, sdev, normpart})
      if (p) cudaFreeAsync(p, stream);
    for (cudaEvent_t& e : ev)
      if (e) {
        cudaEventSynchronize(e);  // the pinned buffer goes back to the pool: no copy may still be in flight
        cudaEventDestroy(e);
        e = nullptr;
      }
  };
  try {
    for (cudaEvent_t& e : ev) GBM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&V), sizeof(double) * ldv * (m_max + 1), stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&w), sizeof(double) * ldv, stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c), sizeof(double) * (m_max + 1), stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&alpha), sizeof(double) * (m_max + 1), stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&beta), sizeof(double) * (m_max + 1), stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sdev), sizeof(double) * (m_max + 1), stream));
    GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&normpart), sizeof(double) * ngroups, stream));
    GBM_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * ldv, stream));
    const unsigned row_blocks = static_cast<unsigned>((n + 255) / 256);
    start_vector_kernel<<<row_blocks, 256, 0, stream>>>(w, n);
    norm_next_kernel<<<1, 1024, 0, stream>>>(w, n, V, beta, m_max);  // V[:, 0] = unit start vector (beta slot unused)

    std::vector<double> ha, hb, s;
    bool converged = false;
    int m = 0;
    double theta = 0.0;
    int next_check = 30;
    // Convergence checks do not drain the stream: alpha / beta go to a pinned buffer behind an event and are examined
    // ONE check later, while the GPU is already working on the following steps (a synchronous check idled the GPU for
    // the round trip + the host's tridiagonal solve every 6-10 steps).  The decision is taken on the same m as before,
    // so the result is unchanged; the steps enqueued past that m touch neither V[:, 0..m) nor alpha / beta[0..m).
    struct Pending {
      int m = 0;
      int slot = 0;
    } pending;
    int slot = 0;
    auto examine = [&](const Pending& c) {  // true: converged at c.m
      GBM_CUDA(cudaEventSynchronize(ev[c.slot]));
      const double* pa = pin.p + static_cast<size_t>(c.slot) * 2 * (m_max + 1);
      ha.assign(pa, pa + c.m);
      hb.assign(pa + (m_max + 1), pa + (m_max + 1) + c.m);
      tridiag_top(ha, hb, c.m, &theta, &s);
      const double resid = fabs(hb[c.m - 1] * s[c.m - 1]);
      return resid <= tol * fabs(theta) || !(hb[c.m - 1] > 1e-14 * fabs(theta));
    };
    for (int j = 0; j < m_max && !converged; ++j) {
      const double* vj = V + static_cast<int64_t>(j) * ldv;
      apply(vj, w);
      double* vnext = V + static_cast<int64_t>(j + 1) * ldv;
      if (coop_grid > 0) {  // Gram-Schmidt twice + alpha_j + beta_j + v_{j+1} in one cooperative launch
        int64_t n_arg = n, ldv_arg = ldv;
        int j_arg = j;
        void* args[] = {&V, &n_arg, &ldv_arg, &j_arg, &w, &c, &alpha, &beta, &normpart, &vnext};
        GBM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(reorth_kernel), dim3(coop_grid), dim3(kReorthThreads),
                                             args, 0, stream));
      } else {
        for (int pass = 0; pass < 2; ++pass) {  // classical Gram-Schmidt, twice
          dots_kernel<<<j + 1, 1024, 0, stream>>>(V, n, ldv, w, c);
          project_out_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(V, n, ldv, j + 1, c, w, alpha, j, pass);
        }
        norm_next_kernel<<<1, 1024, 0, stream>>>(w, n, vnext, beta, j);
      }
      if (j + 1 == next_check || j + 1 == m_max) {
        GBM_CUDA(cudaGetLastError());
        double* pa = pin.p + static_cast<size_t>(slot) * 2 * (m_max + 1);
        GBM_CUDA(cudaMemcpyAsync(pa, alpha, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, stream));
        GBM_CUDA(cudaMemcpyAsync(pa + (m_max + 1), beta, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, stream));
        GBM_CUDA(cudaEventRecord(ev[slot], stream));
        if (pending.m > 0 && examine(pending)) {
          converged = true;
          m = pending.m;
          break;
        }
        pending.m = j + 1;
        pending.slot = slot;
        slot ^= 1;
        next_check = j + 1 + (j + 1 < 100 ? 10 : 6);
      }
    }
    if (!converged && pending.m > 0) {  // the last check enqueued
      converged = examine(pending);
      m = pending.m;
    }
    if (converged) {
      GBM_CUDA(cudaMemcpyAsync(sdev, s.data(), sizeof(double) * m, cudaMemcpyHostToDevice, stream));
      combine_kernel<<<row_blocks, 256, 0, stream>>>(V, n, ldv, m, sdev, w);
      norm_next_kernel<<<1, 1024, 0, stream>>>(w, n, x_dev, beta, 0);  // unit norm
      // explicit residual ||Op x - theta x|| as the final word
      apply(x_dev, w);
      GBM_CUDA(cudaGetLastError());
      std::vector<double> hx(n), hy(n);
      GBM_CUDA(cudaMemcpyAsync(hx.data(), x_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
      GBM_CUDA(cudaMemcpyAsync(hy.data(), w, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
      GBM_CUDA(cudaStreamSynchronize(stream));
      long double rq = 0, r2 = 0;
      for (int64_t i = 0; i < n; ++i) rq += static_cast<long double>(hx[i]) * hy[i];
      for (int64_t i = 0; i < n; ++i) {
        const long double r = hy[i] - rq * hx[i];
        r2 += r * r;
      }
      theta = static_cast<double>(rq);
      converged = sqrt(static_cast<double>(r2)) <= std::max(5.0 * tol, 2e-13) * fabs(theta);
    }
    release();
    if (theta_out) *theta_out = theta;
    if (iters_out) *iters_out = m;
    return converged;
  } catch (...) {
    release();
    throw;
  }
}

// stream-ordered scratch vector
struct Scratch {
  double* p = nullptr;
  cudaStream_t s;
  Scratch(size_t count, cudaStream_t stream) : s(stream) {
    if (count) GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(double) * count, stream));
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
};

}  // namespace

bool lanczos_top_eigenpair(const double* B, int64_t n, int64_t ldb, bool gram, double tol, int max_iter, double* x_dev,
                           double* theta_out, int* iters_out, int sm_count, cudaStream_t stream) {
  if ((ldb & 1) != 0 || (reinterpret_cast<uintptr_t>(B) & 15u) != 0) return false;
  const int64_t ldv = (n + 1) / 2 * 2;
  // gram operator: u = B' v, w = B u through chunk partials
  const bool fused = gram && n <= kFusedMaxN && !(n & 1) && getenv("GBM_PC1_NO_FUSED") == nullptr;
  const int fgrid = static_cast<int>(std::min<int64_t>(n, sm_count));
  Scratch u(gram ? ldv : 0, stream),
      partial(gram ? static_cast<size_t>(n) * std::max<int64_t>(kGemvChunks, fused ? fgrid : 0) : 0, stream);
  const unsigned row_blocks = static_cast<unsigned>((n + 255) / 256);
  const unsigned symv_grid = static_cast<unsigned>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(sm_count) * 8));
  // out = Op v:  Op = B (symmetric), or Op = B B' for the gram operator (B = Z: the eigenvector of Z Z' without
  // ever forming it: two passes over Z per step instead of one pass over Z Z' plus an n^3 SYRK up front)
  auto apply = [&](const double* v, double* out) {
    if (!gram) {
      symv_kernel<<<symv_grid, 256, 0, stream>>>(B, n, n, ldb, v, out);
    } else if (fused && launch_gram_fused(B, n, n, ldb, v, partial.p, fgrid, stream)) {
      launch_partial_reduce(partial.p, n, fgrid, out, stream);  // one pass over B
    } else {
      symv_kernel<<<symv_grid, 256, 0, stream>>>(B, n, n, ldb, v, u.p);  // u[j] = B[:, j] . v
      gemv_n_partial_kernel<<<dim3(row_blocks, kGemvChunks), 256, 0, stream>>>(B, n, n, ldb, u.p, partial.p);
      launch_partial_reduce(partial.p, n, kGemvChunks, out, stream);
    }
  };
  return lanczos_core(n, apply, tol, max_iter, x_dev, theta_out, iters_out, stream);
}

void launch_peer_sum(const PeerMailbox& mb, const double* partial, int chunks, int64_t n, unsigned long long step,
                     double* out, cudaStream_t stream) {
  if (n > kPeerSumMaxN || mb.world < 1 || mb.world > PeerMailbox::kMaxRanks || n > mb.npad)
    GBM_THROW(GBM_ERR_ARGUMENT, "launch_peer_sum: the vector does not fit the mailboxes / a resident grid");
  peer_sum_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(mb, partial, chunks, n, step, out);
  GBM_CUDA(cudaGetLastError());
}

void lanczos_tridiag_top(const double* a, const double* b, int m, double* theta, double* s) {
  std::vector<double> va(a, a + m), vb(m, 0.0), vs;
  for (int i = 0; i + 1 < m; ++i) vb[i] = b[i];
  tridiag_top(va, vb, m, theta, &vs);
  for (int i = 0; i < m; ++i) s[i] = vs[i];
}

void ShardedAllReduce::sum_partials(const double* partial, int chunks, int64_t n, double* out, cudaStream_t stream) {
  launch_partial_reduce(partial, n, chunks, out, stream);
  sum(out, n);
}

void block_row_sums(const double* Zg, int64_t n, int64_t nc, int64_t ld, double* rowsum, cudaStream_t stream) {
  Scratch partial(static_cast<size_t>(n) * kGemvChunks, stream);
  const unsigned row_blocks = static_cast<unsigned>((n + 255) / 256);
  gemv_n_partial_kernel<<<dim3(row_blocks, kGemvChunks), 256, 0, stream>>>(Zg, n, nc, ld, nullptr, partial.p);
  launch_partial_reduce(partial.p, n, kGemvChunks, rowsum, stream);
  GBM_CUDA(cudaGetLastError());
}

bool lanczos_top_singular_sharded(const double* Zg, int64_t n, int64_t nc, int64_t ld, ShardedAllReduce* ar, double tol,
                                  int max_iter, double* x_dev, double* theta_out, int* iters_out, int sm_count,
                                  cudaStream_t stream) {
  if ((ld & 1) != 0 || (reinterpret_cast<uintptr_t>(Zg) & 15u) != 0) return false;
  const bool fused = nc > 0 && n <= kFusedMaxN && !(n & 1) && getenv("GBM_PC1_NO_FUSED") == nullptr;
  const int fgrid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nc, sm_count)));
  Scratch u(static_cast<size_t>(std::max<int64_t>(nc, 1)), stream),
      partial(static_cast<size_t>(n) * std::max<int64_t>(kGemvChunks, fused ? fgrid : 0), stream);
  const unsigned row_blocks = static_cast<unsigned>((n + 255) / 256);
  const unsigned symv_grid =
      static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((nc + 7) / 8, static_cast<int64_t>(sm_count) * 8)));
  auto apply = [&](const double* v, double* out) {
    if (fused && launch_gram_fused(Zg, n, nc, ld, v, partial.p, fgrid, stream)) {  // one pass over the block
      ar->sum_partials(partial.p, fgrid, n, out, stream);
      return;
    }
    if (nc > 0) {  // u = Zg' v
      if (nc < 8192)
        symv_cta_kernel<<<static_cast<unsigned>(nc), 256, 0, stream>>>(Zg, n, ld, v, u.p);
      else
        symv_kernel<<<symv_grid, 256, 0, stream>>>(Zg, n, nc, ld, v, u.p);
    }
    gemv_n_partial_kernel<<<dim3(row_blocks, kGemvChunks), 256, 0, stream>>>(Zg, n, nc, ld, u.p, partial.p);
    launch_partial_reduce(partial.p, n, kGemvChunks, out, stream);                   // this rank's Zg u
    ar->sum(out, n);                                                                   // sum over the ranks
  };
  return lanczos_core(n, apply, tol, max_iter, x_dev, theta_out, iters_out, stream);
}

}  // namespace gbm
