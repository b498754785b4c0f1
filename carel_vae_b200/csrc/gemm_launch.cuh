// gemm_launch.cuh — host-side launcher of gemm_tc_kernel (persistent grid = min(items, #SM)).
#pragma once
#include "api_common.cuh"
#include "gemm_2sm.cuh"
#include "gemm_as.cuh"
#include "gemm_tc.cuh"

namespace carel {

template <class Epi, int BN, int NG, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                       const CUtensorMap& b1, const GemmSched& s, const typename Epi::Params& ep,
                       cudaStream_t st, const char* tag = "gemm_tc") {
  using Cfg = GemmCfg<BN, NG>;
  auto kern = gemm_tc_kernel<Epi, BN, NG, A_MN, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
    attr_done = true;
  }
  const int grid = s.num_items < sm_count() ? s.num_items : sm_count();
  if (grid <= 0) return 0;
  timing_begin(tag, st);
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(a0, b0, a1, b1, s, ep);
  timing_end(st);
  CAREL_LAUNCHED();
  return 0;
}


// Split-K factor without empty splits: `desired` is rounded so that every split owns >= 1 k-block.
inline void set_ksplit(GemmSched& s, int desired) {
  const int num_kb = (s.num_k16 + 3) / 4;
  if (desired < 1) desired = 1;
  if (desired > num_kb) desired = num_kb;
  const int per = (num_kb + desired - 1) / desired;
  s.ksplit = (num_kb + per - 1) / per;
  s.num_items = s.m_tiles * s.n_tiles * s.ksplit;
}

template <class Epi, bool F16 = false, int EPI_SMEM = 0>
int launch_gemm_as(const CUtensorMap& a, const CUtensorMap& b, const AsSched& s, const typename Epi::Params& ep,
                   cudaStream_t st, const char* tag = "gemm_as") {
  auto kern = gemm_as_kernel<Epi, F16, EPI_SMEM>;
  constexpr int kSmem = as_smem_bytes(EPI_SMEM);
  static bool attr_done = false;
  if (!attr_done) {
    CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_done = true;
  }
  CAREL_REQUIRE(s.num_k16 >= 1 && s.num_k16 <= 16, "A-stationary GEMM: K must be <= 256");
  const int grid = (s.total + s.tpc - 1) / s.tpc;
  if (grid <= 0) return 0;
  timing_begin(tag, st);
  kern<<<grid, kGemmThreads, kSmem, st>>>(a, b, s, ep);
  timing_end(st);
  CAREL_LAUNCHED();
  return 0;
}

template <class Epi, bool A_MN = false, bool B_MN = false>
int launch_gemm_2sm(const CUtensorMap& a, const CUtensorMap& b, const PairSched& s, const typename Epi::Params& ep,
                    cudaStream_t st, const char* tag = "gemm_2sm") {
  auto kern = gemm_2sm_kernel<Epi, A_MN, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    attr_done = true;
  }
  if (s.total <= 0) return 0;
  timing_begin(tag, st);
  kern<<<2 * s.clusters, kGemmThreads, kPairSmemBytes, st>>>(a, b, s, ep);     // __cluster_dims__(2, 1, 1)
  timing_end(st);
  CAREL_LAUNCHED();
  return 0;
}

// Split-K factor that fills whole waves of the persistent grid: minimise waves * (k-blocks per item + a fixed
// per-item cost of ~3 k-blocks for the accumulator write-out); ties go to the smaller split (fewer reductions).
inline void set_ksplit_auto(GemmSched& s, int max_split = 32) {
  const int num_kb = (s.num_k16 + 3) / 4;
  const int tiles = s.m_tiles * s.n_tiles, sms = sm_count();
  int best = 1;
  long long best_cost = -1;
  for (int ks = 1; ks <= max_split && ks <= num_kb; ++ks) {
    const long long items = 1LL * tiles * ks;
    const long long waves = (items + sms - 1) / sms;
    const long long cost = waves * ((num_kb + ks - 1) / ks + 3);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ks; }
  }
  set_ksplit(s, best);
}

inline GemmSched make_sched(int64_t M, int64_t N, int64_t K, int BN, int ksplit, int mode) {
  GemmSched s;
  s.m_tiles = int((M + 127) / 128);
  s.n_tiles = int((N + BN - 1) / BN);
  s.num_k16 = int((K + 15) / 16);
  s.mode = mode;
  set_ksplit(s, ksplit);
  s.m_tile0 = 0;
  s.tri_rows = s.m_tiles;
  s.n_cols = int(N);
  s.row_tile_flags = nullptr;
  return s;
}

}  // namespace carel
