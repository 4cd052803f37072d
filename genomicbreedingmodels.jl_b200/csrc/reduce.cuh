// Warp reduction helper shared by the streaming kernels.
#pragma once

namespace gbm {

// Halving butterfly: N values per lane in, N/32 fully reduced values per lane out.  All
// register indices are compile-time.
template <int CNT, int MASK, int N>
struct HalvingStep {
  static __device__ __forceinline__ void run(double (&v)[N], int lane) {
    const bool upper = (lane & MASK) != 0;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
      const double send = upper ? v[i] : v[i + CNT];
      const double keep = upper ? v[i + CNT] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
    if constexpr (MASK > 1) HalvingStep<CNT / 2, MASK / 2, N>::run(v, lane);
  }
};

// Lane -> first original index of the N/32 values it holds after HalvingStep<N/2,16,N>.
template <int N>
__device__ __forceinline__ int halving_base(int lane) {
  return ((lane & 16) ? N / 2 : 0) + ((lane & 8) ? N / 4 : 0) + ((lane & 4) ? N / 8 : 0) + ((lane & 2) ? N / 16 : 0) +
         ((lane & 1) ? N / 32 : 0);
}

}  // namespace gbm
