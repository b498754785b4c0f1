// chain_launch.cuh — host-side launcher and work decomposition of chain_tc_kernel.
#pragma once
#include "api_common.cuh"
#include "chain_tc.cuh"

namespace carel {

// Split the inner range so that the number of work items fills whole waves of the persistent
// grid: pick nsplit in [1, max_split] maximising items / (ceil(items / SMs) * SMs); ties go to
// the smaller split (less atomic traffic).
inline ChainSched make_chain_sched(int owner0, int n_owner, int n_inner, int k16, int stages, int max_split) {
  ChainSched s;
  s.owner0 = owner0;
  s.n_owner = n_owner;
  s.n_inner = n_inner;
  s.k16 = k16;
  s.n2_16 = k16;
  s.stages = stages;
  const int sms = sm_count();
  int best = 1;
  double best_eff = -1.0;
  for (int sp = 1; sp <= max_split && sp <= (n_inner > 0 ? n_inner : 1); ++sp) {
    const int per = (n_inner + sp - 1) / sp;
    if (per * (sp - 1) >= n_inner && sp > 1) continue;  // would create an empty split
    const long long items = 1LL * n_owner * sp;
    const long long waves = (items + sms - 1) / sms;
    // every wave costs `per` inner tiles plus a fixed per-item overhead (~2 tiles)
    const double cost = double(waves) * (per + 2.0);
    const double eff = 1.0 / cost;
    if (eff > best_eff * 1.0001) { best_eff = eff; best = sp; }
  }
  s.nsplit = best;
  s.num_items = n_owner * best;
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
