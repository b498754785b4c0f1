// host_pack.cu — carel_bow_pack_ell_host: the dense BoW target -> ELL list ON THE HOST (no device code in this file).
//
// The reference hands the decoder target over as dense fp32 rows [N, V] built on the host
// (ECPEDataset._get_bow_representations, drl_classifier_ec_mmd_final_mul.py:100-119) and copies them to the GPU every
// step (:830): 983 MB per 8192-clause batch, 19 ms over PCIe Gen5 — 16x the whole train step.  A row holds ~24
// non-zeros, so a host that scans its own rows and ships the list instead moves 27 MB.  This is the same scan as
// carel_bow_pack_ell (bow.cu): `v != 0.f` (so -0.0 is skipped), the first `cap` non-zeros of a row are stored and
// nnz[r] counts all of them; the list is in ascending column order here (the GPU scan interleaves the four columns of a
// lane inside a 128-column group: the same set of (column, value) pairs, which is all the consumers rely on).  Rows are handed out in chunks to a persistent pool of
// `n_threads` - 1 workers plus the calling thread (below); the common all-zero stretch is skipped on the raw bits, 128 bytes at a time with AVX2 (runtime check), else 32.
#include <immintrin.h>

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>

#include "api_common.cuh"

namespace carel {

// 32 floats at p are all +0.0f (AVX2: four unaligned 32-byte loads, or, ptest)
__attribute__((target("avx2"))) static inline bool zero32_avx2(const float* p) {
  const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
  const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8));
  const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 16));
  const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 24));
  const __m256i o = _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d));
  return _mm256_testz_si256(o, o) != 0;
}

// index of the first 32-float block in [j, V - 32] (stepping 32) that is not all zero, or the first j with j + 32 > V
__attribute__((target("avx2"))) static int64_t skip_zero_blocks_avx2(const float* row, int64_t j, int64_t V) {
  while (j + 32 <= V && zero32_avx2(row + j)) j += 32;
  return j;
}

// AVX-512: index of the first 128-float (512-byte) block in [j, V - 128] that is not all zero, or the first j with
// j + 128 > V; the lines 4 KB ahead are requested while the block is tested (tools/host_scan_rate.cpp on the bench box,
// 16 threads: 178 GB/s for the AVX2 loop, 197 GB/s for this one — both close to what the socket's DRAM delivers).
__attribute__((target("avx512f"))) static int64_t skip_zero_blocks_avx512(const float* row, int64_t j, int64_t V) {
  for (; j + 128 <= V; j += 128) {
    const char* p = reinterpret_cast<const char*>(row + j);
    for (int k = 0; k < 512; k += 64) _mm_prefetch(p + 4096 + k, _MM_HINT_T0);
    __m512i o = _mm512_loadu_si512(p);
    for (int k = 64; k < 512; k += 64) o = _mm512_or_si512(o, _mm512_loadu_si512(p + k));
    if (_mm512_test_epi32_mask(o, o) != 0) break;
  }
  return j;
}

static void pack_rows_host(const float* dense, int64_t ld, int64_t r0, int64_t r1, int64_t V, int64_t cap, int32_t* cols,
                           float* vals, int32_t* nnz, int simd) {
  const bool avx2 = simd >= 1, avx512 = simd >= 2;
  for (int64_t r = r0; r < r1; ++r) {
    const float* row = dense + r * ld;
    int32_t* c = cols + r * cap;
    float* v = vals + r * cap;
    int64_t count = 0, j = 0;
    auto take = [&](int64_t jj) {
      const float x = row[jj];
      if (x != 0.f) {
        if (count < cap) { c[count] = int32_t(jj); v[count] = x; }
        ++count;
      }
    };
    if (avx2) {      // 512- / 128-byte blocks on the bits; the 128-byte block that holds a non-zero is taken element-wise
      for (;;) {
        if (avx512) j = skip_zero_blocks_avx512(row, j, V);
        j = skip_zero_blocks_avx2(row, j, V);
        if (j + 32 > V) break;
        for (int k = 0; k < 32; ++k) take(j + k);
        j += 32;
      }
    }
    // head up to an 8-byte boundary, then 32-byte blocks tested on the bits (all-zero bits <=> eight +0.0f)
    while (j < V && (reinterpret_cast<uintptr_t>(row + j) & 7u)) take(j++);
    for (; j + 8 <= V; j += 8) {
      uint64_t w[4];
      std::memcpy(w, row + j, 32);
      if ((w[0] | w[1] | w[2] | w[3]) == 0) continue;
      for (int k = 0; k < 8; ++k) take(j + k);
    }
    for (; j < V; ++j) take(j);
    nnz[r] = int32_t(std::min<int64_t>(count, INT32_MAX));
  }
}

// ---- persistent worker pool ----------------------------------------------------------------------------------------
// A stage of 8192 rows is a few milliseconds of scanning; creating and joining 16 std::threads per call costs 0.3-0.6 ms
// of that.  The workers are created once (detached, parked on a condition variable), the calling thread works too, and
// rows are handed out in chunks from an atomic counter, so a thread that is descheduled for a while does not hold the
// others up.  One job at a time (callers serialise on `call_mu`); a forked child starts a fresh pool (threads do not
// survive fork).
struct ScanJob {
  const float* dense; int64_t ld, N, V, cap; int32_t* cols; float* vals; int32_t* nnz; int simd;
  int64_t chunk;
};

class ScanPool {
 public:
  static ScanPool& get() {
    static std::mutex mu;
    static ScanPool* pool = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (pool == nullptr || owner != getpid()) {      // first use, or a forked child: the parent's workers are not here
      pool = new ScanPool();                          // leaked on purpose: detached workers may outlive static destruction
      owner = getpid();
    }
    return *pool;
  }

  void run(const ScanJob& j, int n_threads) {
    std::lock_guard<std::mutex> call(call_mu_);
    const int helpers = std::max(0, n_threads - 1);
    {
      std::unique_lock<std::mutex> lk(mu_);
      while (int(n_workers_) < helpers) {
        std::thread(&ScanPool::worker, this, n_workers_).detach();
        ++n_workers_;
      }
      job_ = j;
      next_.store(0, std::memory_order_relaxed);
      wanted_ = helpers;           // workers with index < wanted_ take part in this job
      pending_ = helpers;
      ++generation_;
    }
    cv_work_.notify_all();
    work(j);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  void work(const ScanJob& j) {
    for (;;) {
      const int64_t r0 = next_.fetch_add(j.chunk, std::memory_order_relaxed);
      if (r0 >= j.N) break;
      pack_rows_host(j.dense, j.ld, r0, std::min(j.N, r0 + j.chunk), j.V, j.cap, j.cols, j.vals, j.nnz, j.simd);
    }
  }
  void worker(size_t index) {
    uint64_t seen = 0;
    for (;;) {
      ScanJob j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (int(index) >= wanted_) continue;        // not part of this job
        j = job_;
      }
      work(j);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      cv_done_.notify_one();
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  ScanJob job_{};
  std::atomic<int64_t> next_{0};
  size_t n_workers_ = 0;
  int wanted_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
};

}  // namespace carel

extern "C" int carel_bow_pack_ell_host(const float* bow_dense, int64_t ld_bow, int64_t N, int64_t V, int64_t ell_cap,
                                       int32_t* ell_cols, float* ell_vals, int32_t* ell_nnz, int n_threads) {
  using namespace carel;
  CAREL_REQUIRE(bow_dense && ell_cols && ell_vals && ell_nnz, "carel_bow_pack_ell_host: null pointer");
  CAREL_REQUIRE(N >= 0 && V > 0 && ell_cap > 0 && ld_bow >= V, "carel_bow_pack_ell_host: bad shape");
  int nt = n_threads > 0 ? n_threads : int(std::thread::hardware_concurrency());
  nt = int(std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 256), (N + 15) / 16)));
  const int simd = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx2") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
  if (nt == 1) {
    pack_rows_host(bow_dense, ld_bow, 0, N, V, ell_cap, ell_cols, ell_vals, ell_nnz, simd);
    return 0;
  }
  ScanJob j{bow_dense, ld_bow, N, V, ell_cap, ell_cols, ell_vals, ell_nnz, simd, 0};
  // ~8 chunks per thread: fine enough to even out a descheduled thread, coarse enough that the counter is not contended
  j.chunk = std::max<int64_t>(4, (N + int64_t(nt) * 8 - 1) / (int64_t(nt) * 8));
  ScanPool::get().run(j, nt);
  return 0;
}
