// api_common.cu — error string, version, launch counter, debug GEMM and dense pairwise entry points.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "api_common.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "pack.cuh"

#ifndef CAREL_SIDE_BRANCHES_DEFAULT
#define CAREL_SIDE_BRANCHES_DEFAULT 3      // measured on B200 (profiles/r02_side_branches_ab.txt): 815 -> 851 steps/s, suite green either way
#endif

namespace carel {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
// SMs the persistent kernels leave free (carel_set_reserved_sms): a concurrent NCCL collective gets them.
std::atomic<int> g_reserved_sms{0};
// SM count of the CURRENT device (cached per device index: a process may drive several GPUs), minus the reservation.
int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  const int r = g_reserved_sms.load(std::memory_order_relaxed);
  return (r > 0 && n - r >= 8) ? n - r : n;
}

// Side branches (api_common.cuh): one stream + three events per device, created on first use.
std::atomic<int> g_side_branches{[] {
  const char* e = getenv("CAREL_SIDE_BRANCHES");
  return e != nullptr ? (atoi(e) & 3) : CAREL_SIDE_BRANCHES_DEFAULT;
}()};
int side_branch_mask() { return g_side_branches.load(std::memory_order_relaxed); }
SideBranch* side_branch() {
  if (g_side_branches.load(std::memory_order_relaxed) == 0) return nullptr;
  static std::mutex mu;
  static SideBranch table[64];
  static int state[64] = {0};          // 0 not tried, 1 ready, -1 failed
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (state[dev] == 0) {
    SideBranch& b = table[dev];
    const bool ok = cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.ev_join[0], cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.ev_join[1], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) (void)cudaGetLastError();
    state[dev] = ok ? 1 : -1;
  }
  return state[dev] == 1 ? &table[dev] : nullptr;
}

namespace {
struct TimingEntry {
  const char* tag;
  cudaEvent_t e0, e1;
};
// The timing log is a profiling aid for eager passes; launches may come from the caller's thread and from autograd's
// backward thread, so the log is guarded and each thread closes the entry IT opened.
std::atomic<bool> g_timing{false};
std::mutex g_timing_mu;
std::vector<TimingEntry> g_entries;
thread_local long long t_open_entry = -1;
}  // namespace
bool timing_enabled() { return g_timing.load(); }
void timing_begin(const char* tag, cudaStream_t st) {
  if (!g_timing.load()) return;
  TimingEntry e;
  e.tag = tag;
  cudaEventCreate(&e.e0);
  cudaEventCreate(&e.e1);
  cudaEventRecord(e.e0, st);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_entries.push_back(e);
  t_open_entry = (long long)g_entries.size() - 1;
}
void timing_end(cudaStream_t st) {
  if (!g_timing.load()) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (t_open_entry < 0 || t_open_entry >= (long long)g_entries.size()) return;
  cudaEventRecord(g_entries[size_t(t_open_entry)].e1, st);
  t_open_entry = -1;
}

}  // namespace carel

using namespace carel;

extern "C" void carel_set_reserved_sms(int n) { g_reserved_sms.store(n < 0 ? 0 : n, std::memory_order_relaxed); }
extern "C" int carel_get_reserved_sms(void) { return g_reserved_sms.load(std::memory_order_relaxed); }

extern "C" void carel_set_side_branches(int mask) { g_side_branches.store(mask < 0 ? 0 : (mask & 3), std::memory_order_relaxed); }
extern "C" int carel_get_side_branches(void) { return g_side_branches.load(std::memory_order_relaxed); }

extern "C" void carel_timing_enable(int on) {
  g_timing.store(on != 0);
  if (!on) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (auto& e : g_entries) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
    g_entries.clear();
  }
}
// Writes "tag total_ms count\n" lines (one per kernel tag, in first-launch order) and clears the log.
extern "C" int carel_timing_report(char* buf, size_t n) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_timing_mu);
  std::vector<const char*> order;
  std::map<std::string, std::pair<double, int>> agg;
  for (auto& e : g_entries) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.e0, e.e1) != cudaSuccess) ms = 0.f;
    auto it = agg.find(e.tag);
    if (it == agg.end()) { agg[e.tag] = {ms, 1}; order.push_back(e.tag); }
    else { it->second.first += ms; it->second.second += 1; }
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  g_entries.clear();
  size_t off = 0;
  if (n > 0) buf[0] = 0;
  for (const char* t : order) {
    const int w = snprintf(buf + off, off < n ? n - off : 0, "%s %.6f %d\n", t, agg[t].first, agg[t].second);
    if (w < 0 || off + size_t(w) >= n) break;
    off += size_t(w);
  }
  return int(off);
}

extern "C" int carel_version(void) { return CAREL_VERSION; }
extern "C" const char* carel_last_error(void) { return error_buffer(); }
extern "C" int64_t carel_launch_count(void) { return launch_counter().load(); }

extern "C" int carel_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb,
                               int b_mn, int64_t M, int64_t N, int64_t K, float* c, int64_t ldc,
                               const float* bias, int ksplit, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(M > 0 && N > 0 && K > 0, "carel_gemm_bf16: empty problem");
  CAREL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "carel_gemm_bf16: lda/ldb must be multiples of 8");
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_bf16_sw128(&ta, a, M, K, lda, 128);
  else rc = make_tmap_bf16_sw128(&ta, a, K, M, lda, 64);
  CAREL_REQUIRE(rc == 0, "carel_gemm_bf16: tensor map A failed (%d)", rc);
  if (!b_mn) rc = make_tmap_bf16_sw128(&tb, b, N, K, ldb, 128);
  else rc = make_tmap_bf16_sw128(&tb, b, K, N, ldb, 64);
  CAREL_REQUIRE(rc == 0, "carel_gemm_bf16: tensor map B failed (%d)", rc);
  GemmSched s = make_sched(M, N, K, 256, ksplit, TILE_FULL);
  EpiStore<256>::Params ep;
  ep.out = c; ep.bias = bias; ep.out_bf16 = nullptr; ep.M = int(M); ep.N = int(N);
  ep.ld = int(ldc); ep.ld_bf16 = 0; ep.atomic = s.ksplit > 1; ep.scale = 1.f;
  if (!a_mn && !b_mn) return launch_gemm<EpiStore<256>, 256, 1, false, false>(ta, tb, ta, tb, s, ep, st);
  if (!a_mn && b_mn) return launch_gemm<EpiStore<256>, 256, 1, false, true>(ta, tb, ta, tb, s, ep, st);
  if (a_mn && !b_mn) return launch_gemm<EpiStore<256>, 256, 1, true, false>(ta, tb, ta, tb, s, ep, st);
  return launch_gemm<EpiStore<256>, 256, 1, true, true>(ta, tb, ta, tb, s, ep, st);
}

// Same product on CTA pairs (gemm_2sm.cuh, cta_group::2); tiles of 256 x 256, optional split-K.
extern "C" int carel_gemm_bf16_2sm(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb,
                                   int b_mn, int64_t M, int64_t N, int64_t K, float* c, int64_t ldc,
                                   const float* bias, int ksplit, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(M > 0 && N > 0 && K > 0, "carel_gemm_bf16_2sm: empty problem");
  CAREL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "carel_gemm_bf16_2sm: lda/ldb must be multiples of 8");
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_bf16_sw128(&ta, a, M, K, lda, 128);
  else rc = make_tmap_bf16_sw128(&ta, a, K, M, lda, 64);
  CAREL_REQUIRE(rc == 0, "carel_gemm_bf16_2sm: tensor map A failed (%d)", rc);
  if (!b_mn) rc = make_tmap_bf16_sw128(&tb, b, N, K, ldb, 128);
  else rc = make_tmap_bf16_sw128(&tb, b, K, N, ldb, 64);
  CAREL_REQUIRE(rc == 0, "carel_gemm_bf16_2sm: tensor map B failed (%d)", rc);
  PairSched s = make_pair_sched(M, N, K, 0, 0, int((M + 127) / 128), sm_count(), ksplit);
  EpiStore<256>::Params ep;
  ep.out = c; ep.bias = bias; ep.out_bf16 = nullptr; ep.M = int(M); ep.N = int(N);
  ep.ld = int(ldc); ep.ld_bf16 = 0; ep.atomic = s.ksplit > 1; ep.scale = 1.f;
  if (!a_mn && !b_mn) return launch_gemm_2sm<EpiStore<256>, false, false>(ta, tb, s, ep, st);
  if (!a_mn && b_mn) return launch_gemm_2sm<EpiStore<256>, false, true>(ta, tb, s, ep, st);
  if (a_mn && !b_mn) return launch_gemm_2sm<EpiStore<256>, true, false>(ta, tb, s, ep, st);
  return launch_gemm_2sm<EpiStore<256>, true, true>(ta, tb, s, ep, st);
}

// ------------------------------------------------------------------------------------------
// dense pairwise matrices (API completeness; materialises [n1, n2])
// ------------------------------------------------------------------------------------------
extern "C" size_t carel_pairwise_workspace_bytes(int64_t n1, int64_t n2, int64_t dim) {
  const int ldp = aug_width(int(dim));
  WsCarver w(nullptr);
  w.take<__nv_bfloat16>(size_t(n1) * ldp);
  w.take<__nv_bfloat16>(size_t(n1) * ldp);
  w.take<__nv_bfloat16>(size_t(n2) * ldp);
  w.take<__nv_bfloat16>(size_t(n2) * ldp);
  return w.bytes();
}

extern "C" int carel_pairwise(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t n1,
                              int64_t n2, int64_t dim, int mode, const float* alphas_host,
                              int n_alpha, float eps, float* out, int64_t ldo, void* ws,
                              size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(n1 > 0 && n2 > 0 && dim > 0, "carel_pairwise: empty input");
  CAREL_REQUIRE(n_alpha >= 0 && n_alpha <= kMaxAlphas, "carel_pairwise: at most %d alphas", kMaxAlphas);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_pairwise_workspace_bytes(n1, n2, dim), "carel_pairwise: workspace too small");
  const int ldp = aug_width(int(dim));
  WsCarver w(ws);
  __nv_bfloat16* aA = w.take<__nv_bfloat16>(size_t(n1) * ldp);
  __nv_bfloat16* aB = w.take<__nv_bfloat16>(size_t(n1) * ldp);
  __nv_bfloat16* bA = w.take<__nv_bfloat16>(size_t(n2) * ldp);
  __nv_bfloat16* bB = w.take<__nv_bfloat16>(size_t(n2) * ldp);
  pack_rows_kernel<<<int((n1 + 7) / 8), 256, 0, st>>>(a, int(n1), int(dim), int(lda), aA, aB, ldp, 0, 1, nullptr);
  CAREL_LAUNCHED();
  pack_rows_kernel<<<int((n2 + 7) / 8), 256, 0, st>>>(b, int(n2), int(dim), int(ldb), bA, bB, ldp, 0, 1, nullptr);
  CAREL_LAUNCHED();
  CUtensorMap ta, tb;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, aA, n1, ldp, ldp, 128) == 0, "carel_pairwise: tensor map A failed");
  CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, bB, n2, ldp, ldp, 128) == 0, "carel_pairwise: tensor map B failed");
  GemmSched s = make_sched(n1, n2, ldp, 256, 1, TILE_FULL);
  EpiPairwise<256>::Params ep;
  ep.out = out; ep.M = int(n1); ep.N = int(n2); ep.ld = int(ldo);
  ep.mode = mode; ep.n_alpha = n_alpha; ep.eps = eps;
  for (int i = 0; i < kMaxAlphas; ++i) ep.alpha[i] = (i < n_alpha) ? alphas_host[i] : 0.f;
  return launch_gemm<EpiPairwise<256>, 256, 1, false, false>(ta, tb, ta, tb, s, ep, st);
}
