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
#include <string.h>

#include <atomic>
#include <chrono>
#include <exception>
#include <condition_variable>
#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
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

// Division-free exactness test used by the vector bodies.  For a in [0, 1] and c = rint(240 a):
//   fl(c / 240) == a   <=>   r == 0  or  |r| * 2^52 < 120 * 2^e(a),   r = fma(a, 240, -c),
// because 240 a - c is a small multiple of ulp(a) (so the fma returns it exactly), c / 240 is
// either exactly representable (15 | c, then r == 0) or strictly inside a binade at least
// 1/240 - 1/256 away from any power of two (so a and c / 240 share ulp(a) = 2^(e(a) - 52) and
// ties cannot occur), and "a is the double nearest to c / 240" is |a - c/240| < ulp(a) / 2.
// 2^e(a) is a with its mantissa bits cleared.  tests/test_abi_and_host.py checks the
// equivalence with the division on every code, its neighbours and random doubles.
__attribute__((target("avx2,fma"))) static bool pack_col_avx2(const double* col, int64_t n, uint8_t* dst) {
  const __m256d k240 = _mm256_set1_pd(240.0), k120 = _mm256_set1_pd(120.0), one = _mm256_set1_pd(1.0);
  const __m256d two52 = _mm256_set1_pd(4503599627370496.0), zero = _mm256_setzero_pd();
  const __m256d expmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7FF0000000000000ll));
  const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7FFFFFFFFFFFFFFFll));
  __m256d badv = zero;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256d a0 = _mm256_loadu_pd(col + i), a1 = _mm256_loadu_pd(col + i + 4);
    const __m128i c0 = _mm256_cvtpd_epi32(_mm256_mul_pd(a0, k240));  // round to nearest
    const __m128i c1 = _mm256_cvtpd_epi32(_mm256_mul_pd(a1, k240));
    const __m256d r0 = _mm256_and_pd(_mm256_fmsub_pd(a0, k240, _mm256_cvtepi32_pd(c0)), absmask);
    const __m256d r1 = _mm256_and_pd(_mm256_fmsub_pd(a1, k240, _mm256_cvtepi32_pd(c1)), absmask);
    // good = in [0, 1] (ordered compares: NaN fails) and (r == 0 or |r| 2^52 < 120 2^e)
    const __m256d in0 = _mm256_and_pd(_mm256_cmp_pd(a0, zero, _CMP_GE_OQ), _mm256_cmp_pd(a0, one, _CMP_LE_OQ));
    const __m256d in1 = _mm256_and_pd(_mm256_cmp_pd(a1, zero, _CMP_GE_OQ), _mm256_cmp_pd(a1, one, _CMP_LE_OQ));
    const __m256d ok0 = _mm256_or_pd(_mm256_cmp_pd(r0, zero, _CMP_EQ_OQ),
                                     _mm256_cmp_pd(_mm256_mul_pd(r0, two52), _mm256_mul_pd(_mm256_and_pd(a0, expmask), k120), _CMP_LT_OQ));
    const __m256d ok1 = _mm256_or_pd(_mm256_cmp_pd(r1, zero, _CMP_EQ_OQ),
                                     _mm256_cmp_pd(_mm256_mul_pd(r1, two52), _mm256_mul_pd(_mm256_and_pd(a1, expmask), k120), _CMP_LT_OQ));
    badv = _mm256_or_pd(badv, _mm256_or_pd(_mm256_andnot_pd(_mm256_and_pd(in0, ok0), one), _mm256_andnot_pd(_mm256_and_pd(in1, ok1), one)));
    const __m128i w16 = _mm_packus_epi32(c0, c1);
    const __m128i w8 = _mm_packus_epi16(w16, w16);
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), w8);
  }
  bool bad = _mm256_movemask_pd(_mm256_cmp_pd(badv, zero, _CMP_NEQ_UQ)) != 0;
  if (i < n) bad |= pack_col_scalar(col, n, dst, i);
  return bad;
}

__attribute__((target("avx512f,avx512vl,avx512dq"))) static bool pack_col_avx512(const double* col, int64_t n, uint8_t* dst, int pf) {
  const __m512d k240 = _mm512_set1_pd(240.0), k120 = _mm512_set1_pd(120.0), one = _mm512_set1_pd(1.0);
  const __m512d two52 = _mm512_set1_pd(4503599627370496.0), zero = _mm512_setzero_pd();
  const __m512i expmask = _mm512_set1_epi64(0x7FF0000000000000ll);
  __mmask8 good = 0xFF;
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    // software prefetch one page ahead: the hardware streamer stops at 4 KB boundaries (pinned host
    // memory is 4 KB-paged) and the early touch also starts the page walk
    _mm_prefetch(reinterpret_cast<const char*>(col + i) + pf, _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(col + i) + pf + 64, _MM_HINT_T0);
    const __m512d a0 = _mm512_loadu_pd(col + i), a1 = _mm512_loadu_pd(col + i + 8);
    const __m256i c0 = _mm512_cvtpd_epi32(_mm512_mul_pd(a0, k240));
    const __m256i c1 = _mm512_cvtpd_epi32(_mm512_mul_pd(a1, k240));
    const __m512d r0 = _mm512_abs_pd(_mm512_fmsub_pd(a0, k240, _mm512_cvtepi32_pd(c0)));
    const __m512d r1 = _mm512_abs_pd(_mm512_fmsub_pd(a1, k240, _mm512_cvtepi32_pd(c1)));
    const __m512d e0 = _mm512_castsi512_pd(_mm512_and_epi64(_mm512_castpd_si512(a0), expmask));
    const __m512d e1 = _mm512_castsi512_pd(_mm512_and_epi64(_mm512_castpd_si512(a1), expmask));
    const __mmask8 in0 = _mm512_cmp_pd_mask(a0, zero, _CMP_GE_OQ) & _mm512_cmp_pd_mask(a0, one, _CMP_LE_OQ);
    const __mmask8 in1 = _mm512_cmp_pd_mask(a1, zero, _CMP_GE_OQ) & _mm512_cmp_pd_mask(a1, one, _CMP_LE_OQ);
    const __mmask8 ok0 = _mm512_cmp_pd_mask(r0, zero, _CMP_EQ_OQ) |
                         _mm512_cmp_pd_mask(_mm512_mul_pd(r0, two52), _mm512_mul_pd(e0, k120), _CMP_LT_OQ);
    const __mmask8 ok1 = _mm512_cmp_pd_mask(r1, zero, _CMP_EQ_OQ) |
                         _mm512_cmp_pd_mask(_mm512_mul_pd(r1, two52), _mm512_mul_pd(e1, k120), _CMP_LT_OQ);
    good &= in0 & ok0 & in1 & ok1;
    const __m128i b0 = _mm256_cvtepi32_epi8(c0), b1 = _mm256_cvtepi32_epi8(c1);  // low 8 bytes each
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi64(b0, b1));
  }
  bool bad = good != 0xFF;
  if (i < n) bad |= pack_col_scalar(col, n, dst, i);
  return bad;
}

static int pack_prefetch_bytes() {
  static const int v = [] {
    const char* e = getenv("GBM_PACK_PREFETCH");
    return e ? atoi(e) : 4096;
  }();
  return v;
}

// 0 scalar, 1 avx2+fma, 2 avx512; GBM_PACK_ISA=scalar|avx2|avx512 caps it (testing)
static int pack_isa() {
  static const int v = [] {
    int best = 0;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) best = 1;
    if (best == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512dq"))
      best = 2;
    if (const char* e = getenv("GBM_PACK_ISA")) {
      const std::string s(e);
      const int cap = s == "scalar" ? 0 : s == "avx2" ? 1 : 2;
      if (cap < best) best = cap;
    }
    return best;
  }();
  return v;
}

// packs columns [c0, c1) of the block; returns true when an element is not a code
static bool pack_columns(const double* A, int64_t n, int64_t lda, int64_t c0, int64_t c1, uint8_t* out, int64_t ldo) {
  const int isa = pack_isa();
  const int pf = pack_prefetch_bytes();
  for (int64_t j = c0; j < c1; ++j) {
    const double* col = A + j * lda;
    uint8_t* dst = out + j * ldo;
    const bool bad = isa == 2   ? pack_col_avx512(col, n, dst, pf)
                     : isa == 1 ? pack_col_avx2(col, n, dst)
                                : pack_col_scalar(col, n, dst, 0);
    for (int64_t i = n; i < ldo; ++i) dst[i] = 0;
    if (bad) return true;
  }
  return false;
}

// copies columns [c0, c1) into a staging block of pitch ldo doubles (rows n..ldo-1 zeroed): how pageable
// host memory reaches pinned staging on all cores instead of through the driver's single-threaded bounce
static void copy_columns(const double* A, int64_t n, int64_t lda, int64_t c0, int64_t c1, double* out, int64_t ldo) {
  for (int64_t j = c0; j < c1; ++j) {
    memcpy(out + j * ldo, A + j * lda, sizeof(double) * n);
    for (int64_t i = n; i < ldo; ++i) out[j * ldo + i] = 0.0;
  }
}

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

// ---- persistent workers over a queue of block jobs -----------------------------------------
// A job is one column block cut into chunks of ~256 KB of Float64.  Workers take chunks from the
// oldest job that still has some, so several submitted blocks are packed back to back without a
// join between them (gbm_scan_host keeps a few blocks submitted ahead of the one it waits for).
struct PackJob {
  const double* A;
  int64_t n, lda, pc;
  uint8_t* out;   // codes (mode 0) or Float64 staging (mode 1)
  int64_t ldo;    // pitch of out in elements of its type
  int mode = 0;   // 0: pack to codes with the exactness check; 1: plain re-pitching copy
  int64_t chunk, nchunks;
  int64_t next = 0, done = 0;  // guarded by the queue mutex
  std::atomic<int> bad{0};
};

class PackQueue {
 public:
  explicit PackQueue(int n) {
    for (int t = 0; t < n; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~PackQueue() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    work_.notify_all();
    for (auto& w : workers_) w.join();
  }
  void submit(PackJob* job) {
    {
      std::lock_guard<std::mutex> lk(m_);
      jobs_.push_back(job);
    }
    work_.notify_all();
  }
  // Waits for the job, removes it from the queue.  While waiting the calling thread invokes *idle
  // about every 100 us; an exception from idle is re-thrown once the job has finished.
  void wait(PackJob* job, const std::function<void()>* idle) {
    std::unique_lock<std::mutex> lk(m_);
    std::exception_ptr err;
    auto finished = [job] { return job->done == job->nchunks; };
    while (!finished()) {
      if (!idle || err) {
        done_.wait(lk, finished);
        break;
      }
      if (done_.wait_for(lk, std::chrono::microseconds(100), finished)) break;
      lk.unlock();
      try {
        (*idle)();
      } catch (...) {
        err = std::current_exception();
      }
      lk.lock();
    }
    for (size_t i = 0; i < jobs_.size(); ++i)
      if (jobs_[i] == job) {
        jobs_.erase(jobs_.begin() + i);
        break;
      }
    if (err) std::rethrow_exception(err);
  }

 private:
  void loop() {
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      PackJob* job = nullptr;
      for (PackJob* j : jobs_)
        if (j->next < j->nchunks) {
          job = j;
          break;
        }
      if (!job) {
        if (quit_) return;
        work_.wait(lk);
        continue;
      }
      const int64_t c = job->next++;
      lk.unlock();
      if (!job->bad.load(std::memory_order_relaxed)) {
        const int64_t c0 = c * job->chunk, c1 = std::min(job->pc, c0 + job->chunk);
        if (job->mode == 1)
          copy_columns(job->A, job->n, job->lda, c0, c1, reinterpret_cast<double*>(job->out), job->ldo);
        else if (pack_columns(job->A, job->n, job->lda, c0, c1, job->out, job->ldo))
          job->bad.store(1, std::memory_order_relaxed);
      }
      lk.lock();
      if (++job->done == job->nchunks) done_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<PackJob*> jobs_;
  std::mutex m_;
  std::condition_variable work_, done_;
  bool quit_ = false;
};

static PackQueue& pack_queue() {
  static PackQueue q(host_threads());
  return q;
}

// Queues the n x pc block at A (pitch lda) for packing into out (pitch ldo bytes, rows n..ldo-1 zeroed).
PackJob* pack_submit(const double* A, int64_t n, int64_t lda, int64_t pc, uint8_t* out, int64_t ldo) {
  PackJob* job = new PackJob;
  job->A = A;
  job->n = n;
  job->lda = lda;
  job->pc = pc;
  job->out = out;
  job->ldo = ldo;
  job->chunk = std::max<int64_t>(1, (int64_t(256) << 10) / (8 * n));
  job->nchunks = (pc + job->chunk - 1) / job->chunk;
  pack_queue().submit(job);
  return job;
}

// Queues a plain copy of the n x pc block at A (pitch lda) into out (pitch ldo doubles, rows n..ldo-1 zeroed).
PackJob* copy_submit(const double* A, int64_t n, int64_t lda, int64_t pc, double* out, int64_t ldo) {
  PackJob* job = new PackJob;
  job->A = A;
  job->n = n;
  job->lda = lda;
  job->pc = pc;
  job->out = reinterpret_cast<uint8_t*>(out);
  job->ldo = ldo;
  job->mode = 1;
  job->chunk = std::max<int64_t>(1, (int64_t(256) << 10) / (8 * n));
  job->nchunks = (pc + job->chunk - 1) / job->chunk;
  pack_queue().submit(job);
  return job;
}

// Waits for a submitted block and frees the job.  True when every element is exactly a code; false as soon as
// one is not (out is then incomplete and must not be used).
bool pack_wait(PackJob* job, const std::function<void()>* idle) {
  std::unique_ptr<PackJob> own(job);
  pack_queue().wait(job, idle);
  return job->bad.load() == 0;
}

bool pack_block_host(const double* A, int64_t n, int64_t lda, int64_t pc, uint8_t* out, int64_t ldo,
                     const std::function<void()>* idle) {
  return pack_wait(pack_submit(A, n, lda, pc, out, ldo), idle);
}

// Testing hook: the single-column packer body of one ISA level (0 scalar with the division,
// 1 AVX2+FMA, 2 AVX-512) on every column; col_ok[j] = 1 when column j was accepted.
// Returns false when the CPU lacks that ISA.
bool pack_check_columns(const double* A, int64_t n, int64_t lda, int64_t pc, int isa, uint8_t* col_ok) {
  const bool has1 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  const bool has2 = has1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
                    __builtin_cpu_supports("avx512dq");
  if (isa < 0 || isa > 2 || (isa == 1 && !has1) || (isa == 2 && !has2)) return false;
  std::vector<uint8_t> tmp(static_cast<size_t>(n) + 64);
  for (int64_t j = 0; j < pc; ++j) {
    const double* col = A + j * lda;
    const bool bad = isa == 2   ? pack_col_avx512(col, n, tmp.data(), 4096)
                     : isa == 1 ? pack_col_avx2(col, n, tmp.data())
                                : pack_col_scalar(col, n, tmp.data(), 0);
    col_ok[j] = bad ? 0 : 1;
  }
  return true;
}

// count of inexact elements (no early exit) -- for gbm_pack_host's report; failure path only
int64_t count_inexact_host(const double* A, int64_t n, int64_t lda, int64_t pc) {
  const int T = host_threads();
  std::vector<int64_t> bad(T, 0);
  std::atomic<int64_t> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
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
  for (auto& x : th) x.join();
  int64_t tot = 0;
  for (int64_t b : bad) tot += b;
  return tot;
}

}  // namespace gbm
