// Genomic relationship matrix: the dense contraction behind grmsimple(genomes) /
// grmploidyaware(genomes; ploidy) (call sites /root/reference/src/gwas.jl:120, :124;
// GenomicBreedingCore source is absent, definition = "centred X.X^T").
//
//   dK[i, i'] += sum_j (a_ij - mu_j)(a_i'j - mu_j)        lower-triangle 128x128 tiles
//
// FP64 tensor-core path of sm_100a: mma.sync m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no
// FP64 kind).  Persistent CTAs; a producer lane TMA-loads, per 16-marker step, the two
// 132-row x 16-marker boxes of A for the tile's row block and column block into a
// 5-slot mbarrier ring; each slot carries a small metadata word (tile, first/last step).  Boxes are 132 rows (not 128) so that the shared-memory pitch
// between consecutive markers is 1056 B = 4 (mod 16) doubles: the m8n8k4 fragment
// loads (lane -> k = lane&3, row = lane>>2) then hit 32 distinct banks per half-warp with
// no swizzle.  Eight consumer warps (2 x 4) own 64x32 accumulator blocks (64 FP64
// registers per lane), subtract the marker means on the fly (one DADD per fragment
// element, exact centring without a second copy of A) and issue 32 DMMA per k4 step.
// Out-of-range rows / markers are zero-filled by TMA.
//
// Work items are (tile, marker-slice) pairs fetched dynamically (atomic counter) by the
// producer lanes, full tiles before the ragged edge tiles; the slice count is chosen to keep
// the last wave full.  With more than one slice partial tiles
// are combined with FP64 atomics (RED.ADD.F64), otherwise by a plain read-modify-write.
//
// Algorithmic flops (SYRK convention): n (n+1) p.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace gbm {

constexpr int kTile = 128;
constexpr int kTilePad = 132;  // box rows; pitch = 4 (mod 16) doubles
constexpr int kKT = 16;        // markers per stage
constexpr int kGrmStages = 5;
constexpr int kGrmTileBytes = kTilePad * kKT * 8;               // 16896
constexpr int kGrmStageBytes = 2 * kGrmTileBytes + kKT * 8;     // + mu slice (128 B)
constexpr int kGrmConsumerWarps = 8;
constexpr int kGrmThreads = (kGrmConsumerWarps + 1) * 32;
constexpr int kGrmSmemBytes = kGrmStages * kGrmStageBytes + 2 * kGrmStages * 8 + kGrmStages * 16 + 128;

struct GrmParams {
  int64_t n, p;
  int num_tiles;    // lower-triangle tiles
  int num_full_tiles;  // the first num_full_tiles entries of tile_ij are interior (non-ragged) tiles
  int num_slices;   // split of the marker dimension
  int steps_total;  // ceil(p / 16)
  int steps_per_slice;
  const int2* tile_ij;  // (row block, col block), row block >= col block
  const double* mu;     // zero-padded to steps_total * 16
  double* dK;
  int* counter;         // dynamic work counter (zeroed before the launch)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kGrmThreads, 1)
    grm_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const GrmParams prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGrmStages * kGrmStageBytes);
  uint64_t* empty_bar = full_bar + kGrmStages;
  int4* meta = reinterpret_cast<int4*>(empty_bar + kGrmStages);  // per slot: {row block, col block, flags, -}

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kGrmStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kGrmConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int num_items = prm.num_tiles * prm.num_slices;

  if (warp == kGrmConsumerWarps) {
    // ---- producer lane: dynamic work fetch (atomic counter; items are ordered full tiles
    // first, ragged edge tiles last), TMA loads, per-slot metadata for the MMA warps ----
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      int stage = 0;
      uint32_t phase = 0;
      for (;;) {
        const int item = atomicAdd(prm.counter, 1);
        if (item >= num_items) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          meta[stage] = make_int4(0, 0, -1, 0);  // sentinel: no more work
          mbar_arrive(&full_bar[stage]);
          break;
        }
        // all (full tile, slice) items first, slice-major so neighbours share the marker range
        // in L2; then the (edge tile, slice) items
        int tile, slice;
        const int full_items = prm.num_full_tiles * prm.num_slices;
        if (item < full_items) {
          slice = item / prm.num_full_tiles;
          tile = item - slice * prm.num_full_tiles;
        } else {
          const int r = item - full_items, n_edge = prm.num_tiles - prm.num_full_tiles;
          slice = r / n_edge;
          tile = prm.num_full_tiles + (r - slice * n_edge);
        }
        const int2 ij = prm.tile_ij[tile];
        const int s0 = slice * prm.steps_per_slice;
        const int s1 = min(s0 + prm.steps_per_slice, prm.steps_total);
        for (int s = s0; s < s1; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * kGrmStageBytes;
          meta[stage] = make_int4(ij.x, ij.y, (s == s0 ? 1 : 0) | (s == s1 - 1 ? 2 : 0), 0);
          mbar_arrive_expect_tx(&full_bar[stage], kGrmStageBytes);
          tma_load_2d(dst, &tmA, ij.x * kTile, s * kKT, &full_bar[stage], kEvictNormal);
          tma_load_2d(dst + kGrmTileBytes, &tmA, ij.y * kTile, s * kKT, &full_bar[stage], kEvictNormal);
          tma_load_1d(dst + 2 * kGrmTileBytes, prm.mu + static_cast<int64_t>(s) * kKT, kKT * 8, &full_bar[stage]);
          if (++stage == kGrmStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------------------------- MMA warps ----------------------------------
  const int wm = warp >> 2, wn = warp & 3;    // 2 x 4 warps -> 64 x 32 blocks
  const int g = lane >> 2, t = lane & 3;      // fragment row / k index
  int stage = 0;
  uint32_t phase = 0;
  for (;;) {
    // ---- first step of a work item (or the sentinel) ----
    mbar_wait(&full_bar[stage], phase);
    int4 md = meta[stage];
    if (md.z < 0) break;
    const int bi = md.x, bj = md.y;
    double acc[8][4][2];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    // Edge tiles: 8-row groups of this warp's block that lie beyond n hold only TMA zero-fill;
    // skipping them (warp-uniform) makes a ragged last tile row/column cost what it uses.
    const int64_t rows_left = prm.n - (static_cast<int64_t>(bi) * kTile + wm * 64);
    const int64_t cols_left = prm.n - (static_cast<int64_t>(bj) * kTile + wn * 32);
    const int mt_valid = rows_left >= 64 ? 8 : (rows_left <= 0 ? 0 : static_cast<int>((rows_left + 7) / 8));
    const int nt_valid = cols_left >= 32 ? 4 : (cols_left <= 0 ? 0 : static_cast<int>((cols_left + 7) / 8));
    const bool full_tile = (mt_valid == 8) && (nt_valid == 4);

    for (;;) {  // steps of this item; the slot for the current step is already full
      const double* sI = reinterpret_cast<const double*>(smem + stage * kGrmStageBytes);
      const double* sJ = sI + kTilePad * kKT;
      const double* sMu = sJ + kTilePad * kKT;
      const double* pa = sI + t * kTilePad + wm * 64 + g;
      const double* pb = sJ + t * kTilePad + wn * 32 + g;
      if (full_tile) {
#pragma unroll
        for (int kk = 0; kk < kKT / 4; ++kk) {
          // Only the column-block operand is centred: sum_j a_ij (a_i'j - mu_j) = Kc[i,i'] + w_i'
          // with w = (A - 1 mu')mu, removed by grm_wcorrect_kernel.  w has the magnitude of the
          // centred entries themselves, so nothing cancels (unlike A A' - ...), and the MMA loop
          // carries 4 DADD instead of 12 per 32 DMMA.
          const double mu = sMu[kk * 4 + t];
          double a[8], b[4];
#pragma unroll
          for (int mt = 0; mt < 8; ++mt) a[mt] = pa[kk * 4 * kTilePad + mt * 8];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) b[nt] = pb[kk * 4 * kTilePad + nt * 8] - mu;
#pragma unroll
          for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
      } else if (mt_valid > 0 && nt_valid > 0) {
#pragma unroll
        for (int kk = 0; kk < kKT / 4; ++kk) {
          const double mu = sMu[kk * 4 + t];
          double b[4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) b[nt] = pb[kk * 4 * kTilePad + nt * 8] - mu;
#pragma unroll
          for (int mt = 0; mt < 8; ++mt) {
            if (mt < mt_valid) {
              const double a = pa[kk * 4 * kTilePad + mt * 8];
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                if (nt < nt_valid) dmma884(acc[mt][nt][0], acc[mt][nt][1], a, b[nt]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == kGrmStages) {
        stage = 0;
        phase ^= 1u;
      }
      if (md.z & 2) break;  // that was the item's last step
      mbar_wait(&full_bar[stage], phase);
      md = meta[stage];
    }

    // epilogue: accumulate the 64x32 block into dK (column-major, ld = n)
    const int64_t row_base = static_cast<int64_t>(bi) * kTile + wm * 64 + g;
    const int64_t col_base = static_cast<int64_t>(bj) * kTile + wn * 32 + 2 * t;
    const bool atomic = prm.num_slices > 1;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      const int64_t row = row_base + mt * 8;
      if (row >= prm.n) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t col = col_base + nt * 8 + e;
          if (col >= prm.n) continue;
          double* dst = prm.dK + col * prm.n + row;
          if (atomic)
            atomicAdd(dst, acc[mt][nt][e]);
          else
            *dst += acc[mt][nt][e];
        }
      }
    }
  }
}

// w_partial[slab][i] = sum_{j in slab} (a_ij - mu_j) mu_j : thread = row (coalesced down the
// column), the slab's means staged in shared memory.
constexpr int kWSlab = 1024;
__global__ void __launch_bounds__(256)
    grm_wpartial_kernel(const double* __restrict__ A, int64_t n, int64_t p, int64_t lda,
                        const double* __restrict__ mu, double* __restrict__ wpart) {
  __shared__ double smu[kWSlab];
  const int64_t j0 = static_cast<int64_t>(blockIdx.y) * kWSlab;
  const int cnt = static_cast<int>((p - j0) < kWSlab ? (p - j0) : kWSlab);
  for (int c = threadIdx.x; c < cnt; c += blockDim.x) smu[c] = mu[j0 + c];
  __syncthreads();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* col = A + j0 * lda + i;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  int c = 0;
  for (; c + 3 < cnt; c += 4) {
    acc0 = fma(col[(c + 0) * lda] - smu[c + 0], smu[c + 0], acc0);
    acc1 = fma(col[(c + 1) * lda] - smu[c + 1], smu[c + 1], acc1);
    acc2 = fma(col[(c + 2) * lda] - smu[c + 2], smu[c + 2], acc2);
    acc3 = fma(col[(c + 3) * lda] - smu[c + 3], smu[c + 3], acc3);
  }
  for (; c < cnt; ++c) acc0 = fma(col[c * lda] - smu[c], smu[c], acc0);
  wpart[static_cast<int64_t>(blockIdx.y) * n + i] = (acc0 + acc1) + (acc2 + acc3);
}

// w[i] = sum over slabs in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
    grm_wreduce_kernel(const double* __restrict__ wpart, int64_t n, int slabs, double* __restrict__ w) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < slabs; ++b) s += wpart[static_cast<int64_t>(b) * n + i];
  w[i] = s;
}

// dK[i, i'] -= w[i'] on the lower triangle (i >= i')
__global__ void __launch_bounds__(256)
    grm_wcorrect_kernel(double* __restrict__ K, int64_t ld, int64_t rows, const double* __restrict__ w) {
  // K points at the diagonal element of the first column handled by this launch
  const int64_t col = blockIdx.y;
  const double wc = w[col];
  for (int64_t i = col + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    K[col * ld + i] -= wc;
}

static void launch_grm_wcorrection(const double* A, int64_t n, int64_t p, int64_t lda, const double* mu, double* dK,
                                   cudaStream_t stream) {
  const int slabs = static_cast<int>((p + kWSlab - 1) / kWSlab);
  double *wpart = nullptr, *w = nullptr;
  GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&wpart), sizeof(double) * n * slabs, stream));
  GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&w), sizeof(double) * n, stream));
  const unsigned gx = static_cast<unsigned>((n + 255) / 256);
  for (int s0 = 0; s0 < slabs; s0 += 65535) {
    const int sc = slabs - s0 < 65535 ? slabs - s0 : 65535;
    grm_wpartial_kernel<<<dim3(gx, static_cast<unsigned>(sc)), 256, 0, stream>>>(
        A + static_cast<int64_t>(s0) * kWSlab * lda, n, p - static_cast<int64_t>(s0) * kWSlab, lda,
        mu + static_cast<int64_t>(s0) * kWSlab, wpart + static_cast<int64_t>(s0) * n);
  }
  grm_wreduce_kernel<<<gx, 256, 0, stream>>>(wpart, n, slabs, w);
  for (int64_t c0 = 0; c0 < n; c0 += 65535) {
    const unsigned gy = static_cast<unsigned>(n - c0 < 65535 ? n - c0 : 65535);
    grm_wcorrect_kernel<<<dim3(4, gy), 256, 0, stream>>>(dK + c0 * n + c0, n, n - c0, w + c0);
  }
  GBM_CUDA(cudaGetLastError());
  GBM_CUDA(cudaFreeAsync(wpart, stream));
  GBM_CUDA(cudaFreeAsync(w, stream));
}

// lower-triangle tile list (row block i >= column block j): pass 0 the full tiles, pass 1 the ragged edge row
__global__ void grm_tiles_kernel(int2* __restrict__ ij, int nb, int ragged) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int q = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j <= i; ++j) {
        const bool edge = ragged && (i == nb - 1);  // j <= i, so an edge column block implies an edge row block
        if ((pass == 0) != edge) ij[q++] = make_int2(i, j);
      }
}

void launch_grm_accumulate(const double* A, int64_t n, int64_t p, int64_t lda, const double* mu, double* dK,
                           int sm_count, cudaStream_t stream, bool centred) {
  if (n <= 0 || p <= 0) return;
  const int nb = static_cast<int>((n + kTile - 1) / kTile);
  const int num_tiles = nb * (nb + 1) / 2;
  const int steps_total = static_cast<int>((p + kKT - 1) / kKT);

  // slice count: fewest slices (<= 16) whose last wave is >= 95 % full, or the best seen
  // Slices of one tile are combined with FP64 atomics, whose order varies from run to run (last-bit differences in
  // the sums; inside the 1e-9 budget).  GBM_GRM_DETERMINISTIC=1 forbids slicing: every tile is accumulated by one
  // CTA in a fixed order, at the price of a less full last wave for small n (n = 10,000 is unsliced anyway).
  const char* det_env = getenv("GBM_GRM_DETERMINISTIC");
  const int max_s = (det_env && atoi(det_env) != 0) ? 1 : 16;
  int best_s = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_s; ++s) {
    if (s > steps_total) break;
    const int64_t items = static_cast<int64_t>(num_tiles) * s;
    const int64_t waves = (items + sm_count - 1) / sm_count;
    const double eff = static_cast<double>(items) / static_cast<double>(waves * sm_count);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_s = s;
    }
    if (eff >= 0.95) break;
  }
  const int steps_per_slice = (steps_total + best_s - 1) / best_s;
  const int num_slices = (steps_total + steps_per_slice - 1) / steps_per_slice;

  // tile table, built on the device (no pinned host table, no synchronisation per call): full tiles first, ragged
  // edge tiles (last row/column block) last -- with the dynamic fetch the cheap edge tiles fill the tail of the
  // schedule
  const bool ragged = (n % kTile) != 0;
  const int num_full_tiles = ragged ? num_tiles - nb : num_tiles;
  int2* d_ij = nullptr;
  int* d_counter = nullptr;
  GBM_CUDA(cudaMallocAsync(&d_ij, sizeof(int2) * num_tiles, stream));
  GBM_CUDA(cudaMallocAsync(&d_counter, sizeof(int), stream));
  GBM_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), stream));
  grm_tiles_kernel<<<1, 32, 0, stream>>>(d_ij, nb, ragged ? 1 : 0);

  alignas(64) CUtensorMap tmA;
  make_tensor_map_2d_f64(&tmA, A, static_cast<uint64_t>(n), static_cast<uint64_t>(p), static_cast<uint64_t>(lda),
                         kTilePad, kKT);
  GrmParams prm;
  prm.n = n;
  prm.p = p;
  prm.num_tiles = num_tiles;
  prm.num_full_tiles = num_full_tiles;
  prm.num_slices = num_slices;
  prm.steps_total = steps_total;
  prm.steps_per_slice = steps_per_slice;
  prm.tile_ij = d_ij;
  prm.mu = mu;
  prm.dK = dK;
  prm.counter = d_counter;
  const int64_t items = static_cast<int64_t>(num_tiles) * num_slices;
  const int grid = static_cast<int>(items < sm_count ? items : sm_count);
  GBM_CUDA(cudaFuncSetAttribute(grm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrmSmemBytes));
  grm_dmma_kernel<<<grid, kGrmThreads, kGrmSmemBytes, stream>>>(tmA, prm);
  GBM_CUDA(cudaGetLastError());
  if (centred) launch_grm_wcorrection(A, n, p, lda, mu, dK, stream);
  GBM_CUDA(cudaFreeAsync(d_ij, stream));
  GBM_CUDA(cudaFreeAsync(d_counter, stream));
}

// scale the lower triangle and mirror it: 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) grm_finalize_kernel(double* __restrict__ K, int64_t n, double scale) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = static_cast<int64_t>(bi) * 32 + tx, j = static_cast<int64_t>(bj) * 32 + r;
    double v = 0.0;
    if (i < n && j < n && i >= j) {
      v = K[j * n + i] * scale;
      K[j * n + i] = v;
    }
    tile[r][tx] = v;  // tile[jj][ii]
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // write K[jrow, icol] for jrow in block bj (fast index), icol in block bi
    const int64_t jrow = static_cast<int64_t>(bj) * 32 + tx, icol = static_cast<int64_t>(bi) * 32 + r;
    if (jrow < n && icol < n && icol > jrow) K[icol * n + jrow] = tile[tx][r];
  }
}

void launch_grm_finalize(double* dK, int64_t n, double scale, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned nb = static_cast<unsigned>((n + 31) / 32);
  grm_finalize_kernel<<<dim3(nb, nb), 256, 0, stream>>>(dK, n, scale);
  GBM_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) sum_q1mq_kernel(const double* __restrict__ mu, int64_t p, double* out) {
  __shared__ double ws[8];
  double acc = 0.0;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < p;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double q = mu[j];
    acc += q * (1.0 - q);
  }
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += ws[w];
    atomicAdd(out, s);
  }
}

void launch_sum_q1mq(const double* mu, int64_t p, double* out, cudaStream_t stream) {
  if (p <= 0) return;
  int grid = static_cast<int>((p + 255) / 256);
  if (grid > 1024) grid = 1024;
  sum_q1mq_kernel<<<grid, 256, 0, stream>>>(mu, p, out);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
