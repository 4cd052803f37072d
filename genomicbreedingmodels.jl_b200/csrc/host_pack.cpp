// Host-side packer: Float64 allele frequencies -> one-byte dosage codes (a = code / 240) with
// the exactness check fl(code / 240) == a, on all cores of the calling process.  Part of the
// end-to-end path (gbm_scan_host): packing a block on the host lets 8x fewer bytes cross PCIe;
// this is the host half of SURVEY.md 8f rank 3 (the `Matrix{Float64}(allele_frequencies[...])`
// conversion copy of /root/reference/src/prediction.jl:129 fused with a compact encoding).
//
// Plain C++ (no CUDA): AVX2 body selected at run time, scalar fallback, persistent workers.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#ifdef __linux__
#include <sched.h>
#endif

namespace gbm {

// ---- one column ---------------------------------------------------------------------
static inline bool pack_col_scalar(const double* col, int64_t n, uint8_t* dst, int64_t i0) {
  bool bad = false;
  for (int64_t i = i0; i < n; ++i) {
    const double a = col[i];
    const double s = a * 240.0;
    int code = (s >= -0.5 && s < 240.5) ? static_cast<int>(s + 0.5) : 0;
    bad |= (static_cast<double>(code) / 240.0 != a);
    dst[i] = static_cast<uint8_t>(code);
  }
  return bad;
}

__attribute__((target("avx2"))) static bool pack_col_avx2(const double* col, int64_t n, uint8_t* dst) {
  const __m256d k240 = _mm256_set1_pd(240.0);
  const __m128i hi = _mm_set1_epi32(240);
  const __m128i zero = _mm_setzero_si128();
  int badmask = 0;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256d a0 = _mm256_loadu_pd(col + i), a1 = _mm256_loadu_pd(col + i + 4);
    const __m128i c0 = _mm256_cvtpd_epi32(_mm256_mul_pd(a0, k240));  // round to nearest
    const __m128i c1 = _mm256_cvtpd_epi32(_mm256_mul_pd(a1, k240));
    const __m256d b0 = _mm256_div_pd(_mm256_cvtepi32_pd(c0), k240);
    const __m256d b1 = _mm256_div_pd(_mm256_cvtepi32_pd(c1), k240);
    badmask |= _mm256_movemask_pd(_mm256_cmp_pd(b0, a0, _CMP_NEQ_UQ)) | _mm256_movemask_pd(_mm256_cmp_pd(b1, a1, _CMP_NEQ_UQ));
    // range 0..240 (cvt of NaN / huge gives INT_MIN, caught here or by the compare above)
    const __m128i oob = _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(c0, hi), _mm_cmpgt_epi32(zero, c0)),
                                     _mm_or_si128(_mm_cmpgt_epi32(c1, hi), _mm_cmpgt_epi32(zero, c1)));
    badmask |= _mm_movemask_epi8(oob);
    const __m128i w16 = _mm_packus_epi32(c0, c1);
    const __m128i w8 = _mm_packus_epi16(w16, w16);
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), w8);
  }
  bool bad = badmask != 0;
  if (i < n) bad |= pack_col_scalar(col, n, dst, i);
  return bad;
}

static bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

// packs columns [c0, c1) of the block; stops early once `stop` is raised (another worker
// found an element that is not a code) and raises it itself in that case
static void pack_columns(const double* A, int64_t n, int64_t lda, int64_t c0, int64_t c1, uint8_t* out, int64_t ldo,
                         std::atomic<int>* stop) {
  const bool avx2 = have_avx2();
  for (int64_t j = c0; j < c1; ++j) {
    if (stop->load(std::memory_order_relaxed)) return;
    const double* col = A + j * lda;
    uint8_t* dst = out + j * ldo;
    const bool bad = avx2 ? pack_col_avx2(col, n, dst) : pack_col_scalar(col, n, dst, 0);
    for (int64_t i = n; i < ldo; ++i) dst[i] = 0;
    if (bad) {
      stop->store(1, std::memory_order_relaxed);
      return;
    }
  }
}

// ---- persistent workers ---------------------------------------------------------------
class Pool {
 public:
  explicit Pool(int n) : n_(n) {
    for (int t = 0; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> lk(m_);
    fn_ = &fn;
    pending_ = n_;
    ++epoch_;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (quit_) return;
        fn = fn_;
      }
      (*fn)(t);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool quit_ = false;
};

int host_threads() {
  int t = 0;
#ifdef __linux__
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) t = CPU_COUNT(&set);
#endif
  if (t <= 0) t = static_cast<int>(std::thread::hardware_concurrency());
  // one process per GPU (torchrun / MPI): share the host cores between the local ranks
  const char* lws = getenv("LOCAL_WORLD_SIZE");
  if (lws && atoi(lws) > 1) t /= atoi(lws);
  if (t < 1) t = 1;
  return t > 64 ? 64 : t;
}

static Pool& pool() {
  static Pool p(host_threads());
  return p;
}

// Packs the n x pc block at A (pitch lda) into out (pitch ldo bytes, rows n..ldo-1 zeroed).
// Returns true when every element is exactly a code; false as soon as one is not (out is then
// incomplete and must not be used).
bool pack_block_host(const double* A, int64_t n, int64_t lda, int64_t pc, uint8_t* out, int64_t ldo) {
  Pool& pl = pool();
  const int T = pl.size();
  std::atomic<int> stop(0);
  // interleaved chunks of 16 columns keep the workers' streams close together in memory
  const int64_t chunk = 16;
  const int64_t nchunks = (pc + chunk - 1) / chunk;
  std::atomic<int64_t> next(0);
  pl.run([&](int) {
    for (;;) {
      const int64_t c = next.fetch_add(1, std::memory_order_relaxed);
      if (c >= nchunks || stop.load(std::memory_order_relaxed)) return;
      const int64_t c0 = c * chunk, c1 = c0 + chunk < pc ? c0 + chunk : pc;
      pack_columns(A, n, lda, c0, c1, out, ldo, &stop);
    }
  });
  (void)T;
  return stop.load() == 0;
}

// count of inexact elements (no early exit) -- for gbm_pack_host's report
int64_t count_inexact_host(const double* A, int64_t n, int64_t lda, int64_t pc) {
  Pool& pl = pool();
  std::vector<int64_t> bad(pl.size(), 0);
  std::atomic<int64_t> next(0);
  pl.run([&](int t) {
    int64_t local = 0;
    for (;;) {
      const int64_t j = next.fetch_add(1, std::memory_order_relaxed);
      if (j >= pc) break;
      const double* col = A + j * lda;
      for (int64_t i = 0; i < n; ++i) {
        const double a = col[i], s = a * 240.0;
        const int code = (s >= -0.5 && s < 240.5) ? static_cast<int>(s + 0.5) : 0;
        local += (static_cast<double>(code) / 240.0 != a);
      }
    }
    bad[t] = local;
  });
  int64_t tot = 0;
  for (int64_t b : bad) tot += b;
  return tot;
}

}  // namespace gbm
