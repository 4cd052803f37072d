// On-device synthetic allele-frequency generator: the CUDA twin of oracle/synth.py
// (same 64-bit integer arithmetic, so any column block can be re-made on the CPU
// bit-exactly).  Stands in for GenomicBreedingCore.simulategenomes (doctest call sites
// /root/reference/src/gwas.jl:41-45), which is Julia code absent from this image.
#include "common.cuh"
#include "kernels.h"

namespace gbm {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr uint64_t kGold = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kColSalt = 0xD1B54A32D192ED03ull;
constexpr uint64_t kFixSalt = 0x8CB92BA72F3D8DD7ull;
constexpr uint64_t kFixedOneIn = 97;

// grid: (ceil(n / (256*4)), p_local); each thread makes 4 consecutive rows of one column.
__global__ void __launch_bounds__(256) generate_kernel(double* __restrict__ A, int64_t n, int64_t lda,
                                                       int64_t col0, uint64_t seed, int kind) {
  const int64_t jl = blockIdx.y;
  const uint64_t j = static_cast<uint64_t>(col0 + jl);
  const uint64_t hcol = mix64(seed * kGold + (j + 1) * kColSalt);
  const uint64_t u16 = hcol >> 48;
  const uint64_t thr = 3277ull + ((u16 * 29491ull) >> 16);
  const uint64_t hf = mix64(hcol ^ kFixSalt);
  const bool fixed = (hf % kFixedOneIn) == 0;
  const int level = static_cast<int>((hf >> 32) % 3);
  double fixed_val;
  if (level == 0)
    fixed_val = 0.0;
  else if (level == 1)
    fixed_val = 1.0;
  else
    fixed_val = (kind == 0) ? 0.5 : (kind == 1 ? 0.25 : 0.3);

  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  double* col = A + jl * lda;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + r;
    if (i >= n) break;
    const uint64_t h = mix64(hcol + static_cast<uint64_t>(i + 1) * kGold);
    const uint64_t f0 = h & 0xFFFF, f1 = (h >> 16) & 0xFFFF, f2 = (h >> 32) & 0xFFFF, f3 = (h >> 48) & 0xFFFF;
    double a;
    if (kind == 0) {
      a = 0.5 * static_cast<double>((f0 < thr) + (f1 < thr));
    } else if (kind == 1) {
      a = 0.25 * static_cast<double>((f0 < thr) + (f1 < thr) + (f2 < thr) + (f3 < thr));
    } else {
      long long base = static_cast<long long>(thr >> 4);
      long long noise = static_cast<long long>(f0 >> 6) + static_cast<long long>(f1 >> 6) - 1024;
      long long m = base + noise;
      m = m < 0 ? 0 : (m > 4096 ? 4096 : m);
      a = static_cast<double>(m) * (1.0 / 4096.0);
    }
    col[i] = fixed ? fixed_val : a;
  }
}

void launch_generate(double* A, int64_t n, int64_t p, int64_t lda, int64_t col0, uint64_t seed, int kind,
                     cudaStream_t stream) {
  const int64_t max_y = 65535;
  for (int64_t j0 = 0; j0 < p; j0 += max_y) {
    const int64_t pc = (p - j0 < max_y) ? (p - j0) : max_y;
    dim3 grid(static_cast<unsigned>((n + 1023) / 1024), static_cast<unsigned>(pc));
    generate_kernel<<<grid, 256, 0, stream>>>(A + j0 * lda, n, lda, col0 + j0, seed, kind);
  }
}

}  // namespace gbm
