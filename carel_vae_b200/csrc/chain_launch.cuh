// chain_launch.cuh — host-side launcher and work decomposition of chain_tc_kernel.
#pragma once
#include "api_common.cuh"
#include "chain_tc.cuh"

namespace carel {

// Equal contiguous unit ranges per CTA (see ChainIter).  `max_split` is kept for source compatibility.
inline ChainSched make_chain_sched(int owner0, int n_owner, int n_inner, int k16, int stages, int /*max_split*/) {
  ChainSched s;
  s.owner0 = owner0;
  s.n_owner = n_owner;
  s.n_inner = n_inner;
  s.k16 = k16;
  s.n2_16 = k16;
  s.stages = stages;
  s.total_units = n_owner * n_inner;
  const int sms = sm_count();
  int grid = s.total_units < sms ? s.total_units : sms;
  if (grid < 1) grid = 1;
  s.upc = (s.total_units + grid - 1) / grid;
  if (s.upc < 1) s.upc = 1;
  s.num_items = (s.total_units + s.upc - 1) / s.upc;
  s.nsplit = (n_inner > 0 && s.upc % n_inner == 0) ? 1 : 2;
  return s;
}

template <class Pol, int NG, int BN1, int NB = 2>
int launch_chain(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                 const CUtensorMap& b1, const ChainSched& s, const typename Pol::Params& pp,
                 cudaStream_t st, const char* tag = "chain_tc") {
  using Cfg = ChainCfg<NG, BN1, NB>;
  auto kern = chain_tc_kernel<Pol, NG, BN1, NB>;
  CAREL_REQUIRE(16 * s.n2_16 <= Cfg::kMaxN2, "chain kernel: operand width %d exceeds the TMEM budget (%d)", 16 * s.n2_16, Cfg::kMaxN2);
  CAREL_REQUIRE(s.stages >= NB, "chain kernel: operand too wide for the shared-memory ring (%d stages, %d needed)", s.stages, NB);
  const int smem_bytes = Cfg::smem_bytes(s.k16, s.stages);
  static int attr_bytes = 0;
  if (smem_bytes > attr_bytes) {
    CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
    attr_bytes = 227 * 1024;
  }
  const int grid = s.num_items < sm_count() ? s.num_items : sm_count();
  if (grid <= 0) return 0;
  timing_begin(tag, st);
  kern<<<grid, kChainThreads, smem_bytes, st>>>(a0, b0, a1, b1, s, pp);
  timing_end(st);
  CAREL_LAUNCHED();
  return 0;
}

}  // namespace carel
