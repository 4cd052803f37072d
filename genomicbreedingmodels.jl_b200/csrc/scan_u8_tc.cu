// Marker scan on one-byte dosage codes with the per-marker dot products on the tcgen05 INT8 tensor cores.
//
// The streaming sums of the gwasols / gwaslmm marker loop (/root/reference/src/gwas.jl:239-249, :363-389; algebra in
// scan.cu) are, per marker j:  S1 = sum_i c_ij,  S2 = sum_i c_ij^2,  dot_m = sum_i c_ij q_im  for the <= 2 side
// vectors q (PC1 and the residualised trait).  On codes c in [0, 240] the CUDA-core kernel of scan_u8.cu needs three
// FP64-pipe slots per genotype (a conversion and two DFMAs) and is latency-bound at 0.43 of the HBM rate.  Here the
// dots are EXACT integer products instead: every side vector is written as a fixed-point number with seven balanced
// base-256 digits d_k in [-128, 127],
//     q_i = 2^(E - 55) sum_k 256^k d_ik          (E: max |q| < 2^(E-1); the error is below one ulp of the largest q_i)
// so  dot = 2^(E - 55) sum_k 256^k (sum_i c_i d_ik)  and each inner sum is a u8 x s8 -> s32 contraction over the rows:
// one tcgen05.mma kind::i8 with M = 128 markers, N = 16 columns [1 | 7 digits of q_1 | 7 digits of q_2 | 0], K = 32
// rows.  The column of ones gives S1.  The tensor cores run at a few per cent of their rate (16 of 256 columns), the
// FP64 pipe is idle, and the kernel is what the data path allows: HBM -> TMA -> shared memory, read once by the MMA
// and once by four CUDA-core warps that take S2 with dp4a (exact as well).  All sums are integers, so the result does
// not depend on tile order, grid or sharding, and SS of a constant column is exactly 0.
//
// Layout: the code matrix is column-major (a marker's rows are contiguous), i.e. a K-major A operand; a ring slot is
// one TMA box of 128 rows x 128 markers with the 128-byte swizzle (16 KB) plus the matching 128 rows x 16 digit
// columns of the side vectors (2 KB, K-major B operand).  Accumulators live in TMEM (two buffers of 16 columns);
// s32 cannot overflow within 65,536 rows (65,536 * 240 * 128 < 2^31), longer columns are accumulated in segments.
// Warps: 0 = TMA producer lane, 1 = MMA issuer lane (owns TMEM), 2-5 = S2 + epilogue (one thread per marker).
//
// Algorithmic bytes: n per marker.  Records are those of scan_u8.cu ([mean, SS, dot_1, dot_2] in allele-frequency
// units), so scan_finalize_kernel is shared.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace gbm {

namespace {

constexpr int kTcMarkers = 128;                          // M of the MMA: markers per tile
constexpr int kTcRows = 128;                             // rows per ring slot: one 128-byte swizzled line per marker
constexpr int kTcN = 16;                                 // digit columns (N of the MMA)
constexpr int kTcDigits = 7;
constexpr int kTcABytes = kTcMarkers * kTcRows;          // 16384
constexpr int kTcBBytes = kTcN * kTcRows;                // 2048
constexpr int kTcStageBytes = kTcABytes + kTcBBytes;     // 18432 = 18 * 1024: every tile stays 1024-byte aligned
constexpr int kTcStages = 10;
constexpr int kTcSegChunks = 512;                        // 65,536 rows per accumulation segment
constexpr int kTcThreads = 192;
constexpr int kTcTmemCols = 32;                          // two accumulators of 16 columns
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*alignment slack*/ + 512;
constexpr double kLevels = 240.0;

struct TcParams {
  int64_t n, p;
  int num_tiles, chunks;
  int ns;             // record stride (2 + padded side-vector count)
  int m;              // side vectors (0..2)
  double inv_n;
  double scale[2];    // 2^(E_m - 55)
  double* rec;
};

// dense, D = s32, A = unsigned 8-bit, B = signed 8-bit, both K-major, N = 16, M = 128
constexpr uint32_t kTcIdesc = (2u << 4) | (0u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((kTcN >> 3) << 17) |
                              ((kTcMarkers >> 4) << 24);

__global__ void __launch_bounds__(kTcThreads, 1)
    scan_sums_u8_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const TcParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* empty_bar = full_bar + kTcStages;
  uint64_t* tfull_bar = empty_bar + kTcStages;  // [2] accumulator segment complete
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 5);  // tcgen05.commit of the MMA lane + the four S2 warps
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_seg = (prm.chunks + kTcSegChunks - 1) / kTcSegChunks;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      prefetch_tensormap(&tmA);
      prefetch_tensormap(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < prm.chunks; ++chunk) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * kTcStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kTcStageBytes);
          tma_load_2d(dst, &tmA, chunk * kTcRows, tile * kTcMarkers, &full_bar[stage], kEvictFirst);
          tma_load_2d(dst + kTcABytes, &tmB, chunk * kTcRows, 0, &full_bar[stage], kEvictLast);
          if (++stage == kTcStages) {
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
      uint32_t tphase[2] = {0, 0};
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        for (int seg = 0; seg < num_seg; ++seg) {
          mbar_wait(&tempty_bar[buf], tphase[buf] ^ 1u);  // the previous user of this accumulator has read it out
          tc_fence_after();
          const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * kTcN);
          const int c0 = seg * kTcSegChunks, c1 = min(c0 + kTcSegChunks, prm.chunks);
          for (int chunk = c0; chunk < c1; ++chunk) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + stage * kTcStageBytes);
            const uint32_t b_base = a_base + kTcABytes;
#pragma unroll
            for (int k4 = 0; k4 < kTcRows / 32; ++k4)  // K = 32 rows = 32 bytes further inside the swizzled lines
              mma_i8(d_addr, make_desc_k_sw128(a_base + k4 * 32), make_desc_k_sw128(b_base + k4 * 32), kTcIdesc,
                     (chunk == c0 && k4 == 0) ? 0u : 1u);
            tc_commit(&empty_bar[stage]);
            if (++stage == kTcStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          tc_commit(&tfull_bar[buf]);
          tphase[buf] ^= 1u;
          buf ^= 1;
        }
      }
    }
  } else {
    // ------------------------- S2 (dp4a) + epilogue: one thread per marker -------------------------
    const int q = warp & 3;              // TMEM lane quadrant of this warp
    const int t = q * 32 + lane;         // marker inside the tile = TMEM lane = line of the A tile
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t fphase[2] = {0, 0};
    for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
      unsigned long long s2 = 0ull;
      long long s1 = 0;
      double dots[2] = {0.0, 0.0};
      uint32_t code0 = 0;
      for (int seg = 0; seg < num_seg; ++seg) {
        const int c0 = seg * kTcSegChunks, c1 = min(c0 + kTcSegChunks, prm.chunks);
        for (int chunk = c0; chunk < c1; ++chunk) {
          mbar_wait(&full_bar[stage], phase);
          // line t of the swizzled tile: its eight 16-byte pieces sit at piece index (k ^ (t & 7)); reading them in
          // that order makes the eight lanes of a quarter-warp hit eight different bank groups
          const uint4* line = reinterpret_cast<const uint4*>(smem + stage * kTcStageBytes + t * 128);
          uint32_t acc = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint4 w = line[k ^ (t & 7)];
            if (k == 0 && chunk == 0) code0 = w.x & 0xFFu;  // logical piece 0, byte 0: the marker's first code
            acc = __dp4a(w.x, w.x, acc);
            acc = __dp4a(w.y, w.y, acc);
            acc = __dp4a(w.z, w.z, acc);
            acc = __dp4a(w.w, w.w, acc);
          }
          s2 += acc;  // <= 128 * 240^2 per slot
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // this segment's accumulator: 16 s32 columns of TMEM lane t
        mbar_wait(&tfull_bar[buf], fphase[buf]);
        fphase[buf] ^= 1u;
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(tmem_base + static_cast<uint32_t>(buf * kTcN) + (static_cast<uint32_t>(q * 32) << 16), r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        buf ^= 1;
        s1 += static_cast<int>(r[0]);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          double v = 0.0;
#pragma unroll
          for (int k = kTcDigits - 1; k >= 0; --k) v = v * 256.0 + static_cast<double>(static_cast<int>(r[1 + m * kTcDigits + k]));
          dots[m] += v;  // Horner in FP64: every partial value is an integer below 2^53 * 2^14
        }
      }
      // record: shift by the first code in exact integer arithmetic (a constant column gives SS == 0)
      const int64_t col = static_cast<int64_t>(tile) * kTcMarkers + t;
      if (col < prm.p) {
        const long long cz = static_cast<long long>(code0), n = prm.n;
        const long long i1 = s1 - n * cz;
        const long long i2 = static_cast<long long>(s2) - 2 * cz * s1 + n * cz * cz;
        const double d1 = static_cast<double>(i1);
        double* out = prm.rec + col * prm.ns;
        out[0] = (static_cast<double>(cz) + d1 * prm.inv_n) / kLevels;
        out[1] = fmax(static_cast<double>(i2) - d1 * d1 * prm.inv_n, 0.0) / (kLevels * kLevels);
        for (int m = 0; m < prm.m; ++m) out[2 + m] = dots[m] * prm.scale[m] / kLevels;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTcTmemCols);
}

}  // namespace

int scan_u8_tc_digit_rows(int64_t n) { return static_cast<int>((n + 127) / 128 * 128); }

// Host: the fixed-point digits of the side vectors.  Q: n x M (ldq), M <= 2, every column orthogonal to 1.
// digits: 16 x ld (ld = scan_u8_tc_digit_rows(n)) int8, column 0 = ones, 1..7 / 8..14 = digits of q_1 / q_2 (least
// significant first), zero padded.  scale[m] = 2^(E_m - 55).
void scan_u8_tc_build_digits(const double* Q, int64_t n, int M, int64_t ldq, int8_t* digits, int64_t ld, double* scale) {
  for (int64_t i = 0; i < static_cast<int64_t>(kTcN) * ld; ++i) digits[i] = 0;
  for (int64_t i = 0; i < n; ++i) digits[i] = 1;
  for (int m = 0; m < M && m < 2; ++m) {
    const double* q = Q + static_cast<int64_t>(m) * ldq;
    double mx = 0.0;
    for (int64_t i = 0; i < n; ++i) mx = fmax(mx, fabs(q[i]));
    int E = 0;
    if (mx > 0.0) {
      frexp(mx, &E);  // mx = f * 2^E with f in [0.5, 1)
      E += 1;         // one bit of headroom: |q| 2^(55-E) < 2^54, so the balanced top digit stays within [-64, 64]
    }                 // (without it a q_i within 0.4 % of 2^E would carry into an eighth digit)
    const double up = ldexp(1.0, 55 - E);
    scale[m] = ldexp(1.0, E - 55);
    for (int64_t i = 0; i < n; ++i) {
      // balanced base-256 digits, least significant first
      long long v = llrint(q[i] * up);
      for (int k = 0; k < kTcDigits; ++k) {
        long long d = v & 0xFF;
        if (d >= 128) d -= 256;
        digits[static_cast<int64_t>(1 + m * kTcDigits + k) * ld + i] = static_cast<int8_t>(d);
        v = (v - d) >> 8;
      }
      // v == 0 here: |q| 2^(55-E) <= 2^54 = 64 * 256^6
    }
  }
}

void launch_scan_sums_u8_tc(const uint8_t* A8, int64_t n, int64_t p, int64_t ld8, const int8_t* digits, int64_t ldd,
                            const double* scale, int M, int rec_stride, double* rec, int sm_count, cudaStream_t stream) {
  if (p <= 0 || n <= 0) return;
  alignas(64) CUtensorMap tmA, tmB;
  make_tensor_map_2d_u8_sw128(&tmA, A8, static_cast<uint64_t>(n), static_cast<uint64_t>(p), static_cast<uint64_t>(ld8),
                              kTcRows, kTcMarkers);
  make_tensor_map_2d_u8_sw128(&tmB, digits, static_cast<uint64_t>(n), static_cast<uint64_t>(kTcN),
                              static_cast<uint64_t>(ldd), kTcRows, kTcN);
  TcParams prm;
  prm.n = n;
  prm.p = p;
  prm.num_tiles = static_cast<int>((p + kTcMarkers - 1) / kTcMarkers);
  prm.chunks = static_cast<int>((n + kTcRows - 1) / kTcRows);
  prm.ns = rec_stride;
  prm.m = M;
  prm.inv_n = 1.0 / static_cast<double>(n);
  prm.scale[0] = M > 0 ? scale[0] : 0.0;
  prm.scale[1] = M > 1 ? scale[1] : 0.0;
  prm.rec = rec;
  GBM_CUDA(cudaFuncSetAttribute(scan_sums_u8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  const int grid = prm.num_tiles < sm_count ? prm.num_tiles : sm_count;
  scan_sums_u8_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(tmA, tmB, prm);
  GBM_CUDA(cudaGetLastError());
}

}  // namespace gbm
