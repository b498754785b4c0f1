// gemm_launch.cuh — host-side launcher of gemm_tc_kernel (persistent grid = min(items, #SM)).
#pragma once
#include "api_common.cuh"
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

inline GemmSched make_sched(int64_t M, int64_t N, int64_t K, int BN, int ksplit, int mode) {
  GemmSched s;
  s.m_tiles = int((M + 127) / 128);
  s.n_tiles = int((N + BN - 1) / BN);
  s.num_k16 = int((K + 15) / 16);
  s.mode = mode;
  set_ksplit(s, ksplit);
  s.m_tile0 = 0;
  s.n_cols = int(N);
  s.row_tile_flags = nullptr;
  return s;
}

}  // namespace carel
