// Shared device/host helpers for libgbm_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

namespace gbm {

// ------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------
void set_error(const std::string& msg);
struct Error {
  int code;
  std::string msg;
};
#define GBM_THROW(code_, msg_) throw ::gbm::Error{(code_), std::string(msg_)}
#define GBM_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      throw ::gbm::Error{3, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                                ":" + std::to_string(__LINE__) + ")"};                              \
  } while (0)

// One State per driven GPU.  The process-wide default one belongs to gbm_init; a gbm_group (group.cu) owns
// one per local GPU and runs each on its own host thread.  state() resolves to the State bound to the calling
// thread (bind_state), the default one otherwise; entry points are serialised per State.
struct State {
  std::mutex api_mutex;
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // the stream every kernel of the library is launched on
  cudaStream_t copy_stream = nullptr;
  void* cusolver = nullptr;
  cudaEvent_t ev[8] = {};
  // gbm_scan_host staging, cached across calls.  Host lane: blocks packed to codes by the host cores
  // (pinned host_codes -> dev_codes).  Copy-engine lane: Float64 blocks DMA'd as they are (raw_f64),
  // packed on the device (raw_codes, raw_flag = count of non-code elements).
  static constexpr int kRawSlots = 3;
  static constexpr int kHostSlots = 3;
  void* host_codes[kHostSlots] = {};
  void* dev_codes[kHostSlots] = {};
  size_t code_bytes = 0;
  cudaEvent_t hl_copied[kHostSlots] = {}, hl_consumed[kHostSlots] = {};
  void* raw_f64[kRawSlots] = {};
  void* raw_codes[kRawSlots] = {};
  size_t raw_bytes = 0, raw_code_bytes = 0;
  unsigned long long* raw_flag_dev = nullptr;   // kRawSlots counters
  unsigned long long* raw_flag_host = nullptr;  // pinned mirror
  cudaEvent_t raw_copied[kRawSlots] = {}, raw_packed[kRawSlots] = {}, raw_consumed[kRawSlots] = {};
  cudaStream_t raw_stream = nullptr;
  // pinned Float64 staging ring for uploads from pageable host memory (filled by the host workers)
  void* up_stage[kHostSlots] = {};
  size_t up_bytes = 0;
  cudaEvent_t up_copied[kHostSlots] = {};
  int lane_choice = 0, lane_choice_threads = 0;  // gbm_scan_host: measured lane (1 host packer, 2 copy engine)
  double h2d_ms = 0, kernel_ms = 0, main_ms = 0, d2h_ms = 0;
  int64_t launches = 0;
  int64_t packed_blocks = 0, host_packed_blocks = 0, h2d_bytes = 0;
};
State& state();
void bind_state(State* s);  // nullptr: back to the process-wide default State
void init_state(State& st, int device);
void shutdown_state(State& st);
void require_ready();

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
void make_tensor_map_2d_f64(CUtensorMap* map, const double* base, uint64_t rows, uint64_t cols,
                            uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);
// byte matrix (codes) with the 128-byte swizzle, box = box_rows (<=128) x box_cols: the canonical MN-major
// SWIZZLE_128B operand layout of tcgen05.mma for 8-bit types
void make_tensor_map_2d_u8_sw128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_bytes,
                                 uint32_t box_rows, uint32_t box_cols);
// byte matrix, no swizzle: box_rows bytes (a multiple of 16) x box_cols columns
void make_tensor_map_2d_u8(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_bytes,
                           uint32_t box_rows, uint32_t box_cols);
// byte matrix viewed as uint64 words: rows64 words per column, column pitch ld_bytes (multiple of 16)
void make_tensor_map_2d_u64(CUtensorMap* map, const void* base, uint64_t rows64, uint64_t cols, uint64_t ld_bytes,
                            uint32_t box_rows64, uint32_t box_cols);

// ------------------------------------------------------------------------------------
// device-side PTX wrappers: mbarrier + TMA (cp.async.bulk[.tensor])
// ------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// L2 eviction-priority policies (createpolicy encodings as used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace gbm
