// host_scan_rate.cpp — microbenchmark behind carel_bow_pack_ell_host's inner loop (csrc/host_pack.cu): how fast can the
// host cores find the non-zeros of a dense fp32 [N, V] BoW target?  Variants of the all-zero-block skip, thread counts,
// rows interleaved per thread.  Build: g++ -O2 -pthread -o tools/host_scan_rate tools/host_scan_rate.cpp
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
using namespace std;
static const int64_t N = 8192, V = 30000;

static inline int64_t slow_block(const float* p, int n) { int64_t c = 0; for (int k = 0; k < n; ++k) c += p[k] != 0.f; return c; }

// v0: the product loop (128-byte blocks, 4 x 32-byte loads, ptest)
__attribute__((target("avx2"))) static int64_t row_v0(const float* row) {
  int64_t j = 0, c = 0;
  for (; j + 32 <= V; j += 32) {
    __m256i a = _mm256_loadu_si256((const __m256i*)(row + j)), b = _mm256_loadu_si256((const __m256i*)(row + j + 8)),
            cc = _mm256_loadu_si256((const __m256i*)(row + j + 16)), d = _mm256_loadu_si256((const __m256i*)(row + j + 24));
    __m256i o = _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(cc, d));
    if (!_mm256_testz_si256(o, o)) c += slow_block(row + j, 32);
  }
  return c + slow_block(row + j, int(V - j));
}
// v1: v0 + software prefetch PF bytes ahead (one line per 128-byte block pair)
template <int PF> __attribute__((target("avx2"))) static int64_t row_v1(const float* row) {
  int64_t j = 0, c = 0;
  for (; j + 32 <= V; j += 32) {
    _mm_prefetch((const char*)(row + j) + PF, _MM_HINT_T0);
    _mm_prefetch((const char*)(row + j) + PF + 64, _MM_HINT_T0);
    __m256i a = _mm256_loadu_si256((const __m256i*)(row + j)), b = _mm256_loadu_si256((const __m256i*)(row + j + 8)),
            cc = _mm256_loadu_si256((const __m256i*)(row + j + 16)), d = _mm256_loadu_si256((const __m256i*)(row + j + 24));
    __m256i o = _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(cc, d));
    if (!_mm256_testz_si256(o, o)) c += slow_block(row + j, 32);
  }
  return c + slow_block(row + j, int(V - j));
}
// v2: avx512, 256-byte blocks
__attribute__((target("avx512f"))) static int64_t row_v2(const float* row) {
  int64_t j = 0, c = 0;
  for (; j + 64 <= V; j += 64) {
    __m512i o = _mm512_or_si512(_mm512_or_si512(_mm512_loadu_si512(row + j), _mm512_loadu_si512(row + j + 16)),
                                _mm512_or_si512(_mm512_loadu_si512(row + j + 32), _mm512_loadu_si512(row + j + 48)));
    if (_mm512_test_epi32_mask(o, o)) c += slow_block(row + j, 64);
  }
  return c + slow_block(row + j, int(V - j));
}
// v4: avx512, BLK floats per test (64 = 256 bytes, 128 = 512 bytes), software prefetch PF bytes ahead (0 = none)
template <int BLK, int PF> __attribute__((target("avx512f"))) static int64_t row_v4(const float* row) {
  int64_t j = 0, c = 0;
  for (; j + BLK <= V; j += BLK) {
    if (PF) for (int k = 0; k < BLK * 4; k += 64) _mm_prefetch((const char*)(row + j) + PF + k, _MM_HINT_T0);
    __m512i o = _mm512_loadu_si512(row + j);
    for (int k = 16; k < BLK; k += 16) o = _mm512_or_si512(o, _mm512_loadu_si512(row + j + k));
    if (_mm512_test_epi32_mask(o, o)) c += slow_block(row + j, BLK);
  }
  return c + slow_block(row + j, int(V - j));
}
// v3: R rows interleaved per thread (R independent streams per core), 128-byte blocks each
template <int R> __attribute__((target("avx2"))) static int64_t rows_v3(const float* base, int64_t ld) {
  int64_t c = 0, j = 0;
  for (; j + 32 <= V; j += 32) {
    for (int r = 0; r < R; ++r) {
      const float* row = base + r * ld;
      __m256i a = _mm256_loadu_si256((const __m256i*)(row + j)), b = _mm256_loadu_si256((const __m256i*)(row + j + 8)),
              cc = _mm256_loadu_si256((const __m256i*)(row + j + 16)), d = _mm256_loadu_si256((const __m256i*)(row + j + 24));
      __m256i o = _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(cc, d));
      if (!_mm256_testz_si256(o, o)) c += slow_block(row + j, 32);
    }
  }
  for (int r = 0; r < R; ++r) c += slow_block(base + r * ld + j, int(V - j));
  return c;
}

struct Variant { const char* name; int rows; int64_t (*one)(const float*); int64_t (*many)(const float*, int64_t); };

int main(int argc, char** argv) {
  float* d = (float*)aligned_alloc(4096, N * V * 4);
  {  // first touch spread over threads, like pinned allocations that are then written by a loader
    vector<thread> th; int nt = 16;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { memset(d + (N * t / nt) * V, 0, (N * (t + 1) / nt - N * t / nt) * V * 4); });
    for (auto& x : th) x.join();
  }
  srand(1);
  for (int64_t r = 0; r < N; ++r) { int k = 5 + rand() % 35; for (int i = 0; i < k; ++i) d[r * V + rand() % V] = 1.f / k; }
  Variant vs[] = {{"v0 product", 1, row_v0, nullptr}, {"v1 pf512", 1, row_v1<512>, nullptr}, {"v1 pf1024", 1, row_v1<1024>, nullptr},
                  {"v1 pf2048", 1, row_v1<2048>, nullptr}, {"v1 pf4096", 1, row_v1<4096>, nullptr}, {"v2 avx512", 1, row_v2, nullptr},
                  {"v4 256 pf2k", 1, row_v4<64, 2048>, nullptr}, {"v4 256 pf4k", 1, row_v4<64, 4096>, nullptr},
                  {"v4 256 pf8k", 1, row_v4<64, 8192>, nullptr}, {"v4 512", 1, row_v4<128, 0>, nullptr},
                  {"v4 512 pf4k", 1, row_v4<128, 4096>, nullptr}, {"v4 512 pf8k", 1, row_v4<128, 8192>, nullptr},
                  {"v3 2rows", 2, nullptr, rows_v3<2>}, {"v3 4rows", 4, nullptr, rows_v3<4>}, {"v3 8rows", 8, nullptr, rows_v3<8>}};
  printf("hardware_concurrency %u\n", thread::hardware_concurrency());
  for (int nt : {1, 12, 14, 16}) {
    for (auto& v : vs) {
      double best = 1e9; int64_t tot = 0;
      for (int rep = 0; rep < 5; ++rep) {
        atomic<int64_t> next{0}; vector<int64_t> cs(nt, 0); vector<thread> th;
        const int64_t chunk = 32;
        auto t0 = chrono::steady_clock::now();
        for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
          int64_t c = 0;
          for (;;) { int64_t r0 = next.fetch_add(chunk); if (r0 >= N) break;
            for (int64_t r = r0; r < r0 + chunk && r < N; r += v.rows) c += v.one ? v.one(d + r * V) : v.many(d + r * V, V); }
          cs[t] = c; });
        for (auto& x : th) x.join();
        double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
        if (s < best) best = s; tot = 0; for (auto c : cs) tot += c;
      }
      printf("%2d thr %-12s %7.2f ms %6.1f GB/s nnz %ld\n", nt, v.name, best * 1e3, N * V * 4 / best / 1e9, (long)tot);
    }
  }
  return 0;
}
