// chain_tc.cuh — persistent two-GEMM chain on tcgen05 (flash-attention-style recompute):
//
//   for an owner row tile o (128 rows) and every inner tile t of BN1 rows:
//     acc1[g] = A_o[g] * B_t[g]^T                 (GEMM-1, NG products, fp32 in TMEM)
//     P       = policy.weights(acc1)  -> bf16     (epilogue warps, written back to TMEM)
//     acc2   += P * B_t[0]                        (GEMM-2: A from TMEM, B = the SAME smem tile of
//                                                  product 0 read MN-major; fp32 in TMEM)
//   policy.final(acc2)                            (once per work item)
//
// Used by the backward passes: MMD / HSIC (P = C (.) K', acc2 = sum_j W_ij s_j) and the BoW
// decoder (P = dlogits tile, acc2 = dZ).  The N x N (or N x V) matrix P never leaves the SM.
//
// 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// epilogue (warp w: TMEM lanes 32*(w%4).., column half (w-4)/4).
// TMEM columns: acc1 2 x NG x BN1 | P 2 x BN1/2 | acc2 KP   (<= 512).
#pragma once
#include "common.cuh"

namespace carel {

struct ChainSched {
  int n_owner;     // owner tiles (128 rows) in this launch
  int owner0;      // first owner tile index (row-block sharding)
  int n_inner;     // inner tiles (BN1 rows each) in total
  int nsplit;      // inner range split factor; item = owner * nsplit + split
  int num_items;
  int num_kb;      // KP / 64
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
};

template <int NG, int BN1, int KP>
struct ChainCfg {
  static_assert(KP % 64 == 0 && KP <= 256, "KP");
  static_assert(BN1 == 64 || BN1 == 128, "BN1");
  static constexpr int kKB = KP / 64;
  static constexpr int kATile = NG * kKB * 16384;          // owner operand, resident per item
  static constexpr int kBTile = NG * kKB * BN1 * 128;      // one inner tile (all products)
  static constexpr int kStages = 4;
  static constexpr int kAcc1Cols = NG * BN1;               // per buffer
  static constexpr int kPCols = BN1 / 2;                   // per buffer (packed bf16 pairs)
  static constexpr int kAcc1Base = 0;
  static constexpr int kPBase = 2 * kAcc1Cols;
  static constexpr int kAcc2Base = kPBase + 2 * kPCols;
  static_assert(kAcc2Base + KP <= 512, "TMEM budget exceeded");
  static constexpr int kSmemBytes = kATile + kStages * kBTile + 1024 + 512;
  static_assert(kSmemBytes <= 227 * 1024, "smem budget exceeded");
};

template <class Pol, int NG, int BN1, int KP>
__global__ void __launch_bounds__(384, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const ChainSched sched, const typename Pol::Params pp) {
  using Cfg = ChainCfg<NG, BN1, KP>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int KB = Cfg::kKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kATile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::kBTile);
  uint64_t* b_full = bars;                   // [STAGES]
  uint64_t* b_empty = bars + STAGES;         // [STAGES]
  uint64_t* a_full = bars + 2 * STAGES;      // [1]
  uint64_t* a_empty = a_full + 1;            // [1]
  uint64_t* acc1_full = a_empty + 1;         // [2]
  uint64_t* acc1_empty = acc1_full + 2;      // [2]
  uint64_t* p_full = acc1_empty + 2;         // [2]
  uint64_t* p_empty = p_full + 2;            // [2]
  uint64_t* acc2_full = p_empty + 2;         // [1]
  uint64_t* acc2_empty = acc2_full + 1;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NG == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
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
      mbar_expect_tx(a_full, Cfg::kATile);
#pragma unroll 1
      for (int g = 0; g < NG; ++g)
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sA + (g * KB + kb) * 16384, g ? &tmA1 : &tmA0, a_full, kb * 64, it.owner * 128);
      aphase ^= 1;
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&b_empty[stage], phase ^ 1);
        mbar_expect_tx(&b_full[stage], Cfg::kBTile);
        uint8_t* dst = sB + stage * Cfg::kBTile;
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
    uint32_t aphase = 0, acc2_phase = 0;
    int buf = 0; uint32_t bphase = 0;      // acc1 buffer of the next GEMM-1
    int pbuf = 0; uint32_t pphase = 0;     // P buffer of the next GEMM-2
    constexpr uint32_t idesc1 = make_idesc_bf16(128, BN1, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, KP, 0, 1);
    const uint32_t a_base = smem_u32(sA);
    const uint32_t b_base0 = smem_u32(sB);
    const uint32_t acc2_addr = tmem_base + Cfg::kAcc2Base;
    while (iter.next(it)) {
      mbar_wait(a_full, aphase);
      aphase ^= 1;
      bool first_g2 = true;
      auto issue_g2 = [&]() {
        mbar_wait(&p_full[pbuf], pphase);
        if (first_g2) {
          mbar_wait(acc2_empty, acc2_phase ^ 1);
        }
        tc_fence_after();
        const uint32_t b_tile = b_base0 + g2_stage * Cfg::kBTile;   // product 0, MN-major view
        const uint32_t p_addr = tmem_base + Cfg::kPBase + pbuf * Cfg::kPCols;
#pragma unroll
        for (int k16 = 0; k16 < BN1 / 16; ++k16) {
          const uint64_t db = make_sdesc_sw128(b_tile + k16 * 2048, BN1 * 128, 1024);
          umma_ts(acc2_addr, p_addr + k16 * 8, db, idesc2, (first_g2 && k16 == 0) ? 0u : 1u);
        }
        first_g2 = false;
        umma_commit(&b_empty[g2_stage]);
        umma_commit(&p_empty[pbuf]);
        if (++g2_stage == STAGES) g2_stage = 0;
        if (++pbuf == 2) { pbuf = 0; pphase ^= 1; }
      };
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&b_full[stage], phase);
        mbar_wait(&acc1_empty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t b_tile = b_base0 + stage * Cfg::kBTile;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const uint32_t d_addr = tmem_base + Cfg::kAcc1Base + buf * Cfg::kAcc1Cols + g * BN1;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t da = make_sdesc_sw128(a_base + (g * KB + kb) * 16384 + k16 * 32, 16, 1024);
              const uint64_t db = make_sdesc_sw128(b_tile + (g * KB + kb) * (BN1 * 128) + k16 * 32, 16, 1024);
              umma_ss(d_addr, da, db, idesc1, (kb | k16) ? 1u : 0u);
            }
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
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    ChainCtx ctx;
    ctx.half = (warp - 4) >> 2;
    ctx.lane_q = warp & 3;
    ctx.epi_tid = threadIdx.x - 128;
    Pol pol(pp, ctx);
    ChainIter iter(sched);
    ChainItem it;
    int buf = 0; uint32_t bphase = 0;
    int pbuf = 0; uint32_t pphase = 0;
    uint32_t acc2_phase = 0;
    const uint32_t lane_off = uint32_t(ctx.lane_q * 32) << 16;
    constexpr int CH1 = BN1 / 64;   // 32-column chunks per warp per inner tile
    constexpr int CH2 = KP / 64;    // 32-column chunks per warp of acc2
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
        const uint32_t acc1_addr = tmem_base + lane_off + Cfg::kAcc1Base + buf * Cfg::kAcc1Cols;
        const uint32_t p_addr = tmem_base + lane_off + Cfg::kPBase + pbuf * Cfg::kPCols +
                                ctx.half * (Cfg::kPCols / 2);
#pragma unroll 1
        for (int c = 0; c < CH1; ++c) {
          const int tcol = ctx.half * (BN1 / 2) + c * 32;   // column within the inner tile
          uint32_t packed[16];
          pol.template weights<NG, BN1>(ctx, t, tcol, acc1_addr + tcol, packed);
          if (c == 0) mbar_wait(&p_empty[pbuf], pphase ^ 1);
          tmem_st16(p_addr + c * 16, packed);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[pbuf]);
          mbar_arrive(&acc1_empty[buf]);
        }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
        if (++pbuf == 2) { pbuf = 0; pphase ^= 1; }
      }
      mbar_wait(acc2_full, acc2_phase);
      acc2_phase ^= 1;
      tc_fence_after();
      pol.pre_final(ctx);
#pragma unroll 1
      for (int c = 0; c < CH2; ++c) {
        const int fcol = ctx.half * (KP / 2) + c * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_off + Cfg::kAcc2Base + fcol, r);
        tmem_ld_wait();
        pol.final(ctx, fcol, r, it.t1 > it.t0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
    pol.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace carel
