// Exact integer GRM for compact dosage matrices: the Blackwell tensor-core path.
//
// For a packed matrix (one byte per genotype, code c in [0,240], a = c/240; scan_u8.cu) the
// contraction behind grmsimple / grmploidyaware (call sites /root/reference/src/gwas.jl:120,
// :124) is an INTEGER product,  G[i,i'] = sum_j c_ij c_i'j , which the 5th-generation tensor
// cores compute exactly: tcgen05.mma kind::i8 (u8 x u8 -> s32 accumulators in TMEM).  FP64 has
// no tcgen05 kind (grm.cu uses DMMA for Float64 data); for dosage data this path is ~20x
// faster AND exact.  The centred GRM follows from exact integers,
//     Kc[i,i'] = ( G[i,i'] - (U_i + U_i')/n + M2/n^2 ) / 240^2 ,   S_j = sum_i c_ij ,
//     U_i = sum_j S_j c_ij ,   M2 = sum_j S_j^2 .
// G and U are integers below 2^53 (the API checks 57600 n p < 2^53 and otherwise takes the FP64 route), so their
// FP64 atomic accumulation is exact and order-independent; M2 may pass 2^53 and is a rounded sum, added in a fixed
// order (code_sums_final_kernel), so the whole GRM is a deterministic function of the codes; the roundings are
// M2's (1e-16 relative) and the final combination's.
//
// Kernel (one CTA per SM, 192 threads, warp-specialised):
//   warp 0  producer lane: dynamic work fetch, TMA loads (SWIZZLE_128B boxes of 128 rows x
//           128 markers; the matrix is column-major, i.e. MN-major operands, which kind::i8
//           supports) into a 4-slot ring: one box for the tile's 128 rows i, two for its 256 rows i';
//   warp 1  MMA lane: per slot four tcgen05.mma (M = 128, N = 256, K = 32) on smem descriptors,
//           tcgen05.commit frees the slot / publishes the accumulator; owns the TMEM allocation
//           (all 512 columns = 2 x 256: the next item accumulates while the previous one drains);
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns, convert to double, RED.ADD.F64 into
//           the n x n integer-valued matrix (exact, order-independent).
// A work item is (128 x 256 tile touching the lower triangle, slice of <= 32768 markers): 32768 * 240^2 < 2^31,
// so the s32 accumulators cannot overflow.  N = 256 instead of the first version's 128 x 128 tiles: an MMA reads
// (128 + 256) x 32 bytes of shared memory for 128 x 256 x 32 MACs, 25 % fewer bytes per MAC -- at N = 128 the
// tensor pipe waited on the 128 B/clk shared-memory port (40 % active, ncu); the bigger tile also halves the
// number of times a code slab is pulled through L2.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace gbm {

constexpr int kI8Tile = 128;                         // rows i of an output tile (M of the MMA)
constexpr int kI8TileN = 256;                        // rows i' of an output tile (N of the MMA)
constexpr int kI8KB = 128;                           // markers per ring slot
constexpr int kI8Stages = 4;
constexpr int kI8OperandBytes = kI8Tile * kI8KB;     // 16384: 128 markers x 128 rows of codes (one TMA box)
constexpr int kI8StageBytes = 3 * kI8OperandBytes;   // row block + two boxes of the column block
constexpr int kI8SliceStages = 256;                  // 256 * 128 = 32768 markers per accumulation
constexpr int kI8Threads = 192;
constexpr int kI8TmemCols = 512;                     // two 256-column accumulators: all of TMEM
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024 /*alignment slack*/ + 256;

struct I8Params {
  int64_t n, p;
  int num_tiles, num_slices, steps_total;
  const int2* tile_ij;
  double* dG;
  int* counter;
};

// Instruction descriptor: dense, no saturate, D = s32 (2), A = B = unsigned 8-bit (0), both
// MN-major (1), N >> 3 = 32, M >> 4 = 8.
constexpr uint32_t kI8Idesc = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kI8TileN >> 3) << 17) | ((kI8Tile >> 4) << 24);

// MN-major SWIZZLE_128B operand that is TWO 128-wide blocks along MN (the 256 rows i'): the blocks are the two TMA
// boxes, 16384 bytes apart (leading byte offset)
__device__ __forceinline__ uint64_t make_desc_mn_sw128_wide(uint32_t smem_addr) {
  return make_desc_mn_sw128(smem_addr) | (static_cast<uint64_t>(kI8OperandBytes >> 4) << 16);
}

__global__ void __launch_bounds__(kI8Threads, 1)
    grm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const I8Params prm) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kI8Stages * kI8StageBytes);
  uint64_t* empty_bar = full_bar + kI8Stages;
  uint64_t* tfull_bar = empty_bar + kI8Stages;   // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained
  int4* meta = reinterpret_cast<int4*>(tempty_bar + 2);  // [stages] {row block, col block, flags}
  int4* emeta = meta + kI8Stages;                        // [2] per accumulator buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emeta + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kI8Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);   // released by tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);   // tcgen05.commit (or the MMA lane, for the sentinel)
      mbar_init(&tempty_bar[b], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kI8TmemCols);  // whole warp (sync.aligned)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_items = prm.num_tiles * prm.num_slices;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      int stage = 0;
      uint32_t phase = 0;
      for (;;) {
        const int item = atomicAdd(prm.counter, 1);
        if (item >= num_items) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          meta[stage] = make_int4(0, 0, -1, 0);
          mbar_arrive(&full_bar[stage]);
          break;
        }
        const int slice = item / prm.num_tiles, tile = item - slice * prm.num_tiles;
        const int2 ij = prm.tile_ij[tile];
        const int s0 = slice * kI8SliceStages;
        const int s1 = min(s0 + kI8SliceStages, prm.steps_total);
        for (int s = s0; s < s1; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * kI8StageBytes;
          meta[stage] = make_int4(ij.x, ij.y, (s == s0 ? 1 : 0) | (s == s1 - 1 ? 2 : 0), 0);
          mbar_arrive_expect_tx(&full_bar[stage], kI8StageBytes);
          tma_load_2d(dst, &tmA, ij.x * kI8Tile, s * kI8KB, &full_bar[stage], kEvictNormal);
          tma_load_2d(dst + kI8OperandBytes, &tmA, ij.y * kI8TileN, s * kI8KB, &full_bar[stage], kEvictNormal);
          tma_load_2d(dst + 2 * kI8OperandBytes, &tmA, ij.y * kI8TileN + kI8Tile, s * kI8KB, &full_bar[stage], kEvictNormal);
          if (++stage == kI8Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase[2] = {0, 0};  // parity of the next tempty wait per buffer
      for (;;) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const int4 md = meta[stage];
        if (md.z < 0) {
          // forward the sentinel to the epilogue warps through the accumulator hand-off
          mbar_wait(&tempty_bar[buf], tphase[buf] ^ 1u);
          emeta[buf] = make_int4(0, 0, -1, 0);
          __threadfence_block();
          mbar_arrive(&tfull_bar[buf]);
          break;
        }
        if (md.z & 1) {
          // first step of an item: the accumulator buffer must have been drained
          mbar_wait(&tempty_bar[buf], tphase[buf] ^ 1u);
          tc_fence_after();
        }
        const uint32_t a_base = smem_u32(smem + stage * kI8StageBytes);
        const uint32_t b_base = a_base + kI8OperandBytes;
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * kI8TileN);
#pragma unroll
        for (int k4 = 0; k4 < kI8KB / 32; ++k4) {
          // K = 32 markers = four 8-marker swizzle atoms = 4096 bytes further into each box
          const uint64_t da = make_desc_mn_sw128(a_base + k4 * 4096);
          const uint64_t db = make_desc_mn_sw128_wide(b_base + k4 * 4096);
          mma_i8(d_addr, da, db, kI8Idesc, ((md.z & 1) && k4 == 0) ? 0u : 1u);
        }
        tc_commit(&empty_bar[stage]);  // slot is free once these MMAs have read it
        if (md.z & 2) {
          emeta[buf] = make_int4(md.x, md.y, 0, 0);
          __threadfence_block();       // the commit's arrive is asynchronous: publish emeta first
          tc_commit(&tfull_bar[buf]);  // accumulator complete -> epilogue
          tphase[buf] ^= 1u;
          buf ^= 1;
        }
        if (++stage == kI8Stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ---------------------------------- epilogue ----------------------------------
    const int q = warp & 3;  // TMEM lane quadrant this warp may access: lanes 32q .. 32q+31
    int buf = 0;
    uint32_t fphase[2] = {0, 0};
    for (;;) {
      mbar_wait(&tfull_bar[buf], fphase[buf]);
      fphase[buf] ^= 1u;
      tc_fence_after();
      const int4 em = emeta[buf];
      if (em.z < 0) break;
      const int64_t row = static_cast<int64_t>(em.x) * kI8Tile + q * 32 + lane;
      const int64_t col0 = static_cast<int64_t>(em.y) * kI8TileN;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf * kI8TileN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kI8TileN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        if (c == kI8TileN / 32 - 1) {
          // everything this warp needs is in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        if (row < prm.n) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int64_t col = col0 + c * 32 + i;
            if (col < prm.n && r[i] != 0u)
              atomicAdd(prm.dG + col * prm.n + row, static_cast<double>(static_cast<int>(r[i])));
          }
        }
      }
      buf ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kI8TmemCols);
}

// tile table on the device: row block bi (128 rows) x column block bj (256 rows), every pair that touches the
// lower triangle (bj <= bi / 2), in row-block order -- no host table, no pinned allocation, no synchronisation
__global__ void grm_i8_tiles_kernel(int2* __restrict__ ij, int nb) {
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= nb) return;
  // tiles before row block bi: sum_{b < bi} (b / 2 + 1) = h (h - 1) + (bi odd ? h : 0) + bi, h = bi / 2
  const int h = bi >> 1;
  int q = h * (h - 1) + ((bi & 1) ? h : 0) + bi;
  for (int bj = 0; bj <= h; ++bj) ij[q++] = make_int2(bi, bj);
}

void launch_grm_i8_accumulate(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, double* dG, int sm_count,
                              cudaStream_t stream) {
  if (n <= 0 || p <= 0) return;
  const int nb = static_cast<int>((n + kI8Tile - 1) / kI8Tile);
  int num_tiles = 0;
  for (int bi = 0; bi < nb; ++bi) num_tiles += bi / 2 + 1;
  const int steps_total = static_cast<int>((p + kI8KB - 1) / kI8KB);
  const int num_slices = (steps_total + kI8SliceStages - 1) / kI8SliceStages;
  int2* d_ij = nullptr;
  int* d_counter = nullptr;
  GBM_CUDA(cudaMallocAsync(&d_ij, sizeof(int2) * num_tiles, stream));
  GBM_CUDA(cudaMallocAsync(&d_counter, sizeof(int), stream));
  GBM_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), stream));
  grm_i8_tiles_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(d_ij, nb);

  alignas(64) CUtensorMap tmA;
  make_tensor_map_2d_u8_sw128(&tmA, A8, static_cast<uint64_t>(n), static_cast<uint64_t>(p),
                              static_cast<uint64_t>(ld8), kI8Tile, kI8KB);
  I8Params prm;
  prm.n = n;
  prm.p = p;
  prm.num_tiles = num_tiles;
  prm.num_slices = num_slices;
  prm.steps_total = steps_total;
  prm.tile_ij = d_ij;
  prm.dG = dG;
  prm.counter = d_counter;
  GBM_CUDA(cudaFuncSetAttribute(grm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes));
  const int64_t items = static_cast<int64_t>(num_tiles) * num_slices;
  const int grid = static_cast<int>(items < sm_count ? items : sm_count);
  grm_i8_kernel<<<grid, kI8Threads, kI8SmemBytes, stream>>>(tmA, prm);
  GBM_CUDA(cudaGetLastError());
  GBM_CUDA(cudaFreeAsync(d_ij, stream));
  GBM_CUDA(cudaFreeAsync(d_counter, stream));
}

// ---- U_i += sum_j S_j c_ij  (exact integers in FP64), one thread per 8 rows x column slab ----
constexpr int kUSlab = 512;
__global__ void __launch_bounds__(256)
    rowdot_u8_kernel(const uint8_t* __restrict__ A8, int64_t n, int64_t p, int64_t ld8, const double* __restrict__ S,
                     double* __restrict__ U) {
  __shared__ double ss[kUSlab];
  const int64_t j0 = static_cast<int64_t>(blockIdx.y) * kUSlab;
  const int cnt = static_cast<int>((p - j0) < kUSlab ? (p - j0) : kUSlab);
  for (int c = threadIdx.x; c < cnt; c += blockDim.x) ss[c] = S[j0 + c];
  __syncthreads();
  const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // 8-row word index
  if (w * 8 >= n) return;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint64_t* col = reinterpret_cast<const uint64_t*>(A8 + j0 * ld8) + w;
  for (int c = 0; c < cnt; ++c) {
    const uint64_t word = col[static_cast<int64_t>(c) * (ld8 / 8)];
    const double s = ss[c];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fma(static_cast<double>(static_cast<uint32_t>((word >> (8 * k)) & 0xFFu)), s, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t i = w * 8 + k;
    if (i < n && acc[k] != 0.0) atomicAdd(U + i, acc[k]);
  }
}

void launch_rowdot_u8(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, const double* S, double* U,
                      cudaStream_t stream) {
  if (n <= 0 || p <= 0) return;
  const int64_t words = (n + 7) / 8;
  const unsigned gx = static_cast<unsigned>((words + 255) / 256);
  const int64_t slabs = (p + kUSlab - 1) / kUSlab;
  for (int64_t s0 = 0; s0 < slabs; s0 += 65535) {
    const unsigned gy = static_cast<unsigned>(slabs - s0 < 65535 ? slabs - s0 : 65535);
    rowdot_u8_kernel<<<dim3(gx, gy), 256, 0, stream>>>(A8 + s0 * kUSlab * ld8, n, p - s0 * kUSlab, ld8,
                                                       S + s0 * kUSlab, U);
  }
  GBM_CUDA(cudaGetLastError());
}

// S_j = round(mean_j * n * 240) (exact code sums) and M2 = sum_j S_j^2.  M2 can pass 2^53 (S_j <= 240 n), so it is a
// rounded FP64 sum: block partials are written out and added in a FIXED order by one thread (deterministic).
constexpr int kCodeSumBlocks = 1024;
__global__ void __launch_bounds__(256)
    code_sums_kernel(const double* __restrict__ mean, int64_t p, double n_levels, double* __restrict__ S,
                     double* __restrict__ partial) {
  __shared__ double ws[8];
  double acc = 0.0;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < p;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double s = rint(mean[j] * n_levels);
    S[j] = s;
    acc = fma(s, s, acc);
  }
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void code_sums_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ M2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int b = 0; b < blocks; ++b) t += partial[b];
  M2[0] += t;
}

void launch_code_sums(const double* mean, int64_t p, int64_t n, double* S, double* M2, cudaStream_t stream) {
  if (p <= 0) return;
  int grid = static_cast<int>((p + 255) / 256);
  if (grid > kCodeSumBlocks) grid = kCodeSumBlocks;
  double* partial = nullptr;
  GBM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partial), sizeof(double) * grid, stream));
  code_sums_kernel<<<grid, 256, 0, stream>>>(mean, p, static_cast<double>(n) * 240.0, S, partial);
  code_sums_final_kernel<<<1, 32, 0, stream>>>(partial, grid, M2);
  GBM_CUDA(cudaGetLastError());
  GBM_CUDA(cudaFreeAsync(partial, stream));
}

// dK[i,i'] += ( G[i,i'] - centre*((U_i + U_i')/n - M2/n^2) ) / 240^2     on the lower triangle
__global__ void __launch_bounds__(256)
    grm_i8_combine_kernel(const double* __restrict__ G, int64_t n, int64_t col_offset,
                          const double* __restrict__ U, const double* __restrict__ M2, int centre,
                          double* __restrict__ dK) {
  const int64_t col = col_offset + blockIdx.y;
  const double inv_n = 1.0 / static_cast<double>(n);
  const double m2 = centre ? M2[0] * inv_n * inv_n : 0.0;
  const double uc = centre ? U[col] : 0.0;
  for (int64_t i = col + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double g = G[col * n + i];
    const double v = centre ? (g - (U[i] + uc) * inv_n + m2) : g;
    dK[col * n + i] += v * (1.0 / 57600.0);
  }
}

void launch_grm_i8_combine(const double* G, int64_t n, const double* U, const double* M2, int centre, double* dK,
                           cudaStream_t stream) {
  if (n <= 0) return;
  for (int64_t c0 = 0; c0 < n; c0 += 65535) {
    const unsigned gy = static_cast<unsigned>(n - c0 < 65535 ? n - c0 : 65535);
    grm_i8_combine_kernel<<<dim3(4, gy), 256, 0, stream>>>(G, n, c0, U, M2, centre, dK);
  }
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
