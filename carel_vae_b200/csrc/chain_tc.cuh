// chain_tc.cuh — persistent two-GEMM chain on tcgen05 (flash-attention-style recompute):
//
//   for an owner row tile o (128 rows) and every inner tile t of BN1 rows:
//     acc1[g] = A_o[g] * B_t[g]^T                 (GEMM-1, NG products, fp32 in TMEM)
//     P       = policy.weights(acc1)  -> bf16     (epilogue warps, written back to TMEM)
//     acc2   += P * B_t[0]                        (GEMM-2: A = P from TMEM, B = the SAME smem tile
//                                                  of product 0 read MN-major; fp32 in TMEM)
//   policy.final(acc2)                            (once per work item)
//
// Used by the backward passes: MMD / HSIC (P = C (.) K', acc2 = sum_j W_ij s_j) and the BoW
// decoder (P = dlogits tile, acc2 = dZ).  The N x N (or N x V) matrix P never leaves the SM.
//
// 320 threads: warp 0 TMEM allocator + TMA producer lane, warp 1 MMA issuer lane, warps 2-9
// epilogue (warp w: TMEM lanes 32*(w%4).., column half (w-2)/4).
// The operand width is K16 16-wide steps (runtime, <= 16); GEMM-2 has N = 16*K16.
// TMEM columns: acc1 2 x NG x BN1 | P BN1/2 | acc2 16*K16   (<= 512).
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace carel {

struct ChainSched {
  int n_owner;     // owner tiles (128 rows) in this launch
  int owner0;      // first owner tile index (row-block sharding)
  int n_inner;     // inner tiles (BN1 rows each) in total
  int nsplit;      // inner range split factor; item = owner * nsplit + split
  int num_items;
  int k16;         // operand width in 16-wide steps (GEMM-1 K)
  int n2_16;       // GEMM-2 N in 16-column groups (<= k16; acc2 columns = 16 * n2_16)
  int stages;      // B ring depth (2..4), chosen by the host from the smem budget
  int p_fp16;      // 1: the policy writes P as fp16 pairs (GEMM-2 A operand is FP16, B stays BF16)
};

struct ChainItem {
  int owner, t0, t1;
};

struct ChainIter {
  const ChainSched& s;
  int item;
  CAREL_DEVINL explicit ChainIter(const ChainSched& s_) : s(s_), item(blockIdx.x) {}
  CAREL_DEVINL bool next(ChainItem& it) {
    if (item >= s.num_items) return false;
    const int o = item / s.nsplit;
    const int sp = item - o * s.nsplit;
    const int per = (s.n_inner + s.nsplit - 1) / s.nsplit;
    it.owner = s.owner0 + o;
    it.t0 = sp * per;
    it.t1 = min(it.t0 + per, s.n_inner);
    item += gridDim.x;
    return true;
  }
};

struct ChainCtx {
  int owner;      // owner tile index
  int row;        // global owner row of this thread
  int half;       // column half of this warp
  int lane_q;     // lane quarter (0..3)
  int epi_tid;    // 0..255
  int t0, t1;     // inner tile range of the item
  int k16, n2_16;
  uint32_t acc2_taddr;  // TMEM address (lane quarter applied) of acc2 column 0
};

template <int NG, int BN1>
struct ChainCfg {
  static_assert(BN1 == 64 || BN1 == 128, "BN1");
  static constexpr int kAcc1Cols = NG * BN1;       // per buffer
  static constexpr int kPCols = BN1 / 2;           // packed bf16 pairs
  static constexpr int kPBase = 2 * kAcc1Cols;
  static constexpr int kAcc2Base = kPBase + kPCols;
  static constexpr int kMaxN2 = (512 - kAcc2Base) < 256 ? (512 - kAcc2Base) : 256;
  static constexpr int kMaxStages = 4;
  static __host__ __device__ int kb(int k16) { return (k16 + 3) / 4; }
  static __host__ __device__ int a_bytes(int k16) { return NG * kb(k16) * 16384; }
  static __host__ __device__ int b_bytes(int k16) { return NG * kb(k16) * BN1 * 128; }
  static __host__ int stages_for(int k16) {
    const int budget = 227 * 1024 - 2048 - a_bytes(k16);
    int st = budget / b_bytes(k16);
    return st > kMaxStages ? kMaxStages : st;
  }
  static __host__ int smem_bytes(int k16, int stages) {
    return a_bytes(k16) + stages * b_bytes(k16) + 1024 + 512;
  }
};

template <class Pol, int NG, int BN1>
__global__ void __launch_bounds__(kChainThreads, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const __grid_constant__ ChainSched sched,
                const __grid_constant__ typename Pol::Params pp) {
  using Cfg = ChainCfg<NG, BN1>;
  constexpr int MAXST = Cfg::kMaxStages;
  const int K16 = sched.k16;
  const int KB = Cfg::kb(K16);
  const int STAGES = sched.stages;
  const int a_bytes = Cfg::a_bytes(K16);
  const int b_bytes = Cfg::b_bytes(K16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * b_bytes);
  uint64_t* b_full = bars;                   // [MAXST]
  uint64_t* b_empty = bars + MAXST;          // [MAXST]
  uint64_t* a_full = bars + 2 * MAXST;       // [1]
  uint64_t* a_empty = a_full + 1;            // [1]
  uint64_t* acc1_full = a_empty + 1;         // [2]
  uint64_t* acc1_empty = acc1_full + 2;      // [2]
  uint64_t* p_full = acc1_empty + 2;         // [1]
  uint64_t* p_empty = p_full + 1;            // [1]
  uint64_t* acc2_full = p_empty + 1;         // [1]
  uint64_t* acc2_empty = acc2_full + 1;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 32) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NG == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 8);
    }
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 8);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    ChainIter iter(sched);
    ChainItem it;
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    while (iter.next(it)) {
      mbar_wait(a_empty, aphase ^ 1);
      mbar_expect_tx(a_full, a_bytes);
#pragma unroll 1
      for (int g = 0; g < NG; ++g)
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sA + (g * KB + kb) * 16384, g ? &tmA1 : &tmA0, a_full, kb * 64, it.owner * 128);
      aphase ^= 1;
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&b_empty[stage], phase ^ 1);
        mbar_expect_tx(&b_full[stage], b_bytes);
        uint8_t* dst = sB + stage * b_bytes;
#pragma unroll 1
        for (int g = 0; g < NG; ++g)
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(dst + (g * KB + kb) * (BN1 * 128), g ? &tmB1 : &tmB0, &b_full[stage],
                        kb * 64, t * BN1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    ChainIter iter(sched);
    ChainItem it;
    int stage = 0;          // stage of the tile whose GEMM-1 is issued next
    uint32_t phase = 0;
    int g2_stage = 0;       // stage of the tile whose GEMM-2 is issued next
    uint32_t aphase = 0, acc2_phase = 0, pphase = 0;
    int buf = 0; uint32_t bphase = 0;      // acc1 buffer of the next GEMM-1
    constexpr uint32_t idesc1 = make_idesc_bf16(128, BN1, 0, 0);
    const uint32_t idesc2 = sched.p_fp16 ? make_idesc_f16a_bf16b(128, 16 * sched.n2_16, 0, 1)
                                         : make_idesc_bf16(128, 16 * sched.n2_16, 0, 1);
    const uint32_t a_base = smem_u32(sA);
    const uint32_t b_base0 = smem_u32(sB);
    const uint32_t acc2_addr = tmem_base + Cfg::kAcc2Base;
    const uint32_t p_addr = tmem_base + Cfg::kPBase;
    while (iter.next(it)) {
      mbar_wait(a_full, aphase);
      aphase ^= 1;
      bool first_g2 = true;
      auto issue_g2 = [&]() {
        mbar_wait(p_full, pphase);
        pphase ^= 1;
        if (first_g2) mbar_wait(acc2_empty, acc2_phase ^ 1);
        tc_fence_after();
        const uint32_t b_tile = b_base0 + g2_stage * b_bytes;   // product 0, MN-major view
#pragma unroll
        for (int k16 = 0; k16 < BN1 / 16; ++k16) {
          const uint64_t db = make_sdesc_sw128(b_tile + k16 * 2048, BN1 * 128, 1024);
          umma_ts(acc2_addr, p_addr + k16 * 8, db, idesc2, (first_g2 && k16 == 0) ? 0u : 1u);
        }
        first_g2 = false;
        umma_commit(&b_empty[g2_stage]);
        umma_commit(p_empty);
        if (++g2_stage == STAGES) g2_stage = 0;
      };
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&b_full[stage], phase);
        mbar_wait(&acc1_empty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t b_tile = b_base0 + stage * b_bytes;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
          const uint32_t d_addr = tmem_base + buf * Cfg::kAcc1Cols + g * BN1;
#pragma unroll 1
          for (int k = 0; k < K16; ++k) {
            const int kb = k >> 2, k16 = k & 3;
            const uint64_t da = make_sdesc_sw128(a_base + (g * KB + kb) * 16384 + k16 * 32, 16, 1024);
            const uint64_t db = make_sdesc_sw128(b_tile + (g * KB + kb) * (BN1 * 128) + k16 * 32, 16, 1024);
            umma_ss(d_addr, da, db, idesc1, k ? 1u : 0u);
          }
        }
        umma_commit(&acc1_full[buf]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
        if (t > it.t0) issue_g2();
      }
      if (it.t1 > it.t0) issue_g2();
      umma_commit(acc2_full);
      umma_commit(a_empty);
      acc2_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===================== epilogue =====================
    ChainCtx ctx;
    ctx.half = (warp - 2) >> 2;
    ctx.lane_q = warp & 3;
    ctx.epi_tid = threadIdx.x - 64;
    ctx.k16 = K16;
    ctx.n2_16 = sched.n2_16;
    Pol pol(pp, ctx);
    ChainIter iter(sched);
    ChainItem it;
    int buf = 0; uint32_t bphase = 0;
    uint32_t pphase = 0, acc2_phase = 0;
    const uint32_t lane_off = uint32_t(ctx.lane_q * 32) << 16;
    ctx.acc2_taddr = tmem_base + lane_off + Cfg::kAcc2Base;
    const uint32_t p_addr = tmem_base + lane_off + Cfg::kPBase + ctx.half * (Cfg::kPCols / 2);
    while (iter.next(it)) {
      ctx.owner = it.owner;
      ctx.row = it.owner * 128 + ctx.lane_q * 32 + lane;
      ctx.t0 = it.t0;
      ctx.t1 = it.t1;
      pol.begin_item(ctx);
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&acc1_full[buf], bphase);
        tc_fence_after();
        const uint32_t acc1_addr = tmem_base + lane_off + buf * Cfg::kAcc1Cols + ctx.half * (BN1 / 2);
        pol.template weights<BN1>(ctx, t, acc1_addr, p_addr, [&]() {
          mbar_wait(p_empty, pphase ^ 1);
          pphase ^= 1;
        });
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p_full);
          mbar_arrive(&acc1_empty[buf]);
        }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
      mbar_wait(acc2_full, acc2_phase);
      acc2_phase ^= 1;
      tc_fence_after();
      pol.final(ctx, it.t1 > it.t0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
    pol.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace carel
