// chain_tc.cuh — persistent two-GEMM chain on tcgen05 (flash-attention-style recompute):
//
//   for an owner row tile o (128 rows) and every inner tile t of BN1 rows:
//     acc1[g] = A_o[g] * B_t[g]^T                 (GEMM-1, NG products, fp32 in TMEM)
//     P       = policy weights(acc1) -> bf16      (epilogue warps; written IN PLACE over the acc1
//                                                  columns they were computed from, optionally as a
//                                                  hi + lo pair of bf16 = 16 mantissa bits)
//     acc2   += P * B_t[0]                        (GEMM-2: A = P from TMEM, B = the SAME smem tile of
//                                                  product 0 read MN-major; fp32 in TMEM)
//   policy.final(acc2)                            (once per work item)
//
// Used by the backward passes: MMD / HSIC (P = C (.) K', acc2 = sum_j W_ij s_j) and the BoW
// decoder (P = dlogits tile, acc2 = dZ).  The N x N (or N x V) matrix P never leaves the SM.
//
// 128*kChainSplit + 64 threads: 4*kChainSplit epilogue warps (warp w: TMEM lanes 32*(w%4).., column part w/4 of the
// inner tile), then the TMEM allocator + TMA producer warp and the MMA issuer warp (highest ids: scheduler priority).
// TMEM columns: acc1 NB buffers x NG x BN1 | acc2 16*n2_16   (<= 512).  Within an acc1 buffer the half
// part h of product 0 (columns [h R, (h+1) R), R = BN1/kChainSplit) is overwritten by its own thread with
// P_hi (R/2 columns of packed bf16 pairs) followed by P_lo (R/2 columns).
// Ordering: tcgen05.mma instructions of one thread execute in issue order, so GEMM-1 of tile t+2
// (which overwrites the buffer holding P of tile t) needs no barrier beyond being issued after GEMM-2(t).
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace carel {

struct ChainSched {
  int n_owner;     // owner tiles (128 rows) in this launch
  int owner0;      // first owner tile index (row-block sharding)
  int n_inner;     // inner tiles (BN1 rows each) in total
  int nsplit;      // 1: every owner tile is processed whole by one CTA (plain stores allowed); 2: pieces
  int num_items;   // persistent grid size (CTAs with work)
  int total_units; // n_owner * n_inner (owner tile, inner tile) units, owner-major
  int upc;         // units per CTA: CTA b owns units [b * upc, (b + 1) * upc)
  int k16;         // operand width in 16-wide steps (GEMM-1 K)
  int n2_16;       // GEMM-2 N in 16-column groups (<= k16; acc2 columns = 16 * n2_16)
  int stages;      // B ring depth, chosen by the host from the smem budget
};

struct ChainItem {
  int owner, t0, t1;
};

// Stream-K style decomposition: the (owner, inner) units are cut into equal contiguous ranges, one per CTA, so
// the load is balanced to one inner tile whatever the owner count, a CTA touches at most
// 1 + ceil(upc / n_inner) owners (that many accumulator write-outs) and the write-outs of different CTAs fall
// at different times instead of colliding in one burst of reductions.
struct ChainIter {
  const ChainSched& s;
  int pos, end;
  CAREL_DEVINL explicit ChainIter(const ChainSched& s_) : s(s_) {
    pos = blockIdx.x * s.upc;
    end = min(pos + s.upc, s.total_units);
  }
  CAREL_DEVINL bool next(ChainItem& it) {
    if (pos >= end) return false;
    const int o = pos / s.n_inner;
    it.owner = s.owner0 + o;
    it.t0 = pos - o * s.n_inner;
    it.t1 = min(s.n_inner, it.t0 + (end - pos));
    pos += it.t1 - it.t0;
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

template <int NG, int BN1, int NB = 2>
struct ChainCfg {
  static_assert(BN1 == 64 || BN1 == 128, "BN1");
  static_assert(NB >= 2 && NB <= 4, "acc1 buffers");
  static constexpr int kAcc1Cols = NG * BN1;       // per buffer
  static constexpr int kAcc2Base = NB * kAcc1Cols;
  static_assert(kAcc2Base < 512, "TMEM budget");
  static constexpr int kMaxN2 = (512 - kAcc2Base) < 256 ? (512 - kAcc2Base) : 256;
  static constexpr int kMaxStages = 6;
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

template <class Pol, int NG, int BN1, int NB = 2>
__global__ void __launch_bounds__(kChainThreads, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const __grid_constant__ ChainSched sched,
                const __grid_constant__ typename Pol::Params pp) {
  using Cfg = ChainCfg<NG, BN1, NB>;
  constexpr int MAXST = Cfg::kMaxStages;
  constexpr int R = BN1 / kChainSplit;       // columns of the inner tile per epilogue thread
  constexpr bool HILO = Pol::kHiLo;
  constexpr bool F16 = Pol::kF16;            // operands and the weight tile P in FP16 instead of BF16
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
  uint64_t* acc1_full = a_empty + 1;         // [NB]
  uint64_t* p_full = acc1_full + NB;         // [NB]
  uint64_t* acc2_full = p_full + NB;         // [1]
  uint64_t* acc2_empty = acc2_full + 1;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Roles: warps [0, kEpi) epilogue; warp kEpi = TMEM allocator + TMA producer lane; warp kEpi + 1 = MMA issuer
  // lane.  The two single-lane roles sit in the HIGHEST warp ids on purpose: the SMSP arbiter prefers higher
  // warp ids, and an MMA issuer that has to queue behind spinning / MUFU-heavy epilogue warps starves the
  // tensor pipe (measured: narrow MMAs took 109 clk per issue instead of 55).
  constexpr int kEpi = 4 * kChainSplit;

  if (threadIdx.x == (kEpi + 1) * 32) {
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
    for (int i = 0; i < NB; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&p_full[i], 4 * kChainSplit);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 4 * kChainSplit);
    fence_mbar_init();
  }
  if (warp == kEpi) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpi && elect_one()) {   // elect.sync: ptxas knows ONE lane runs this region, so the
    // uniform-register operands of UTMALDG / UTCHMMA need no per-instruction broadcast loop
    // ===================== TMA producer =====================
    ChainIter iter(sched);
    ChainItem it;
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    int trace_n = 0; (void)trace_n;
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
        CAREL_TRACE(0, trace_n); ++trace_n;
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
  } else if (warp == kEpi + 1 && elect_one()) {
    // ===================== MMA issuer =====================
    ChainIter iter(sched);
    ChainItem it;
    int g1_stage = 0; uint32_t g1_phase = 0;   // smem stage of the next GEMM-1
    int g2_stage = 0;                          // smem stage of the next GEMM-2
    int g1_buf = 0;                            // acc1 buffer of the next GEMM-1
    int g2_buf = 0; uint32_t p_phases = 0;     // bit b = phase of p_full[b]
    uint32_t aphase = 0, acc2_phase = 0;
    int trace_g1 = 0, trace_g2 = 0; (void)trace_g1; (void)trace_g2;
    constexpr uint32_t idesc1 = F16 ? make_idesc_f16(128, BN1, 0, 0) : make_idesc_bf16(128, BN1, 0, 0);
    const uint32_t idesc2 = F16 ? make_idesc_f16(128, 16 * sched.n2_16, 0, 1) : make_idesc_bf16(128, 16 * sched.n2_16, 0, 1);
    const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(sA), 16, 1024);
    const uint32_t b_base0 = smem_u32(sB);
    const uint32_t acc2_addr = tmem_base + Cfg::kAcc2Base;
    while (iter.next(it)) {
      mbar_wait(a_full, aphase);
      aphase ^= 1;
      bool first_g2 = true;
      auto issue_g1 = [&]() {
        mbar_wait(&b_full[g1_stage], g1_phase);
        CAREL_TRACE(1, trace_g1); ++trace_g1;
        tc_fence_after();
        const uint32_t b_tile = b_base0 + g1_stage * b_bytes;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
          const uint32_t d_addr = tmem_base + g1_buf * Cfg::kAcc1Cols + g * BN1;
          uint64_t da = sdesc_advance(a_desc0, g * KB * 16384);
          uint64_t db = make_sdesc_sw128(b_tile + g * KB * (BN1 * 128), 16, 1024);
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            const int nk = min(4, K16 - kb * 4);
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              if (k16 < nk) umma_ss(d_addr, sdesc_advance(da, k16 * 32), sdesc_advance(db, k16 * 32), idesc1, (kb | k16) ? 1u : 0u);
            da = sdesc_advance(da, 16384);
            db = sdesc_advance(db, BN1 * 128);
          }
        }
        CAREL_TRACE(6, trace_g1 - 1);
        umma_commit(&acc1_full[g1_buf]);
        if (++g1_stage == STAGES) { g1_stage = 0; g1_phase ^= 1; }
        if (++g1_buf == NB) g1_buf = 0;
      };
      auto issue_g2 = [&]() {
        mbar_wait(&p_full[g2_buf], (p_phases >> g2_buf) & 1u);
        p_phases ^= 1u << g2_buf;
        if (first_g2) mbar_wait(acc2_empty, acc2_phase ^ 1);
        CAREL_TRACE(2, trace_g2); ++trace_g2;
        tc_fence_after();
        const uint64_t db0 = make_sdesc_sw128(b_base0 + g2_stage * b_bytes, BN1 * 128, 1024);   // product 0, MN-major view
        const uint32_t p_base = tmem_base + g2_buf * Cfg::kAcc1Cols;
#pragma unroll
        for (int h = 0; h < kChainSplit; ++h) {
#pragma unroll
          for (int part = 0; part < (HILO ? 2 : 1); ++part) {
#pragma unroll
            for (int k16 = 0; k16 < R / 16; ++k16) {
              umma_ts(acc2_addr, p_base + h * R + part * (R / 2) + k16 * 8, sdesc_advance(db0, (h * R + k16 * 16) * 128),
                      idesc2, first_g2 ? 0u : 1u);
              first_g2 = false;
            }
          }
        }
        CAREL_TRACE(7, trace_g2 - 1);
        umma_commit(&b_empty[g2_stage]);
        if (++g2_stage == STAGES) g2_stage = 0;
        if (++g2_buf == NB) g2_buf = 0;
      };
      // GEMM-1 runs NB - 1 inner tiles ahead of GEMM-2 (needs STAGES >= NB, checked by the launcher)
      int next_g1 = it.t0;
#pragma unroll 1
      for (; next_g1 < it.t1 && next_g1 < it.t0 + NB - 1; ++next_g1) issue_g1();
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        if (next_g1 < it.t1) { issue_g1(); ++next_g1; }
        issue_g2();
      }
      umma_commit(acc2_full);
      umma_commit(a_empty);
      acc2_phase ^= 1;
    }
  } else if (warp < kEpi) {
    // ===================== epilogue =====================
    ChainCtx ctx;
    ctx.half = warp >> 2;
    ctx.lane_q = warp & 3;
    ctx.epi_tid = threadIdx.x;
    ctx.k16 = K16;
    ctx.n2_16 = sched.n2_16;
    Pol pol(pp, ctx);
    ChainIter iter(sched);
    ChainItem it;
    int buf = 0;
    uint32_t full_phases = 0;                  // bit b = phase of acc1_full[b]
    uint32_t acc2_phase = 0;
    int trace_e = 0; (void)trace_e;
    const bool tracer = (threadIdx.x == 0); (void)tracer;
    const uint32_t lane_off = uint32_t(ctx.lane_q * 32) << 16;
    ctx.acc2_taddr = tmem_base + lane_off + Cfg::kAcc2Base;
    while (iter.next(it)) {
      ctx.owner = it.owner;
      ctx.row = it.owner * 128 + ctx.lane_q * 32 + lane;
      ctx.t0 = it.t0;
      ctx.t1 = it.t1;
      pol.begin_item(ctx);
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        pol.prefetch(ctx, t * BN1 + ctx.half * R);       // loads that do not depend on acc1: issued before the wait
        mbar_wait(&acc1_full[buf], (full_phases >> buf) & 1u);
        full_phases ^= 1u << buf;
        if (tracer) CAREL_TRACE(3, trace_e);
        tc_fence_after();
        // this thread's region of product 0 (product 1 sits BN1 columns further)
        const uint32_t reg0 = tmem_base + lane_off + buf * Cfg::kAcc1Cols + ctx.half * R;
        uint32_t r0[R / 16][16];
        uint32_t r1[NG == 2 ? R / 16 : 1][16];
#pragma unroll
        for (int c = 0; c < R / 16; ++c) {
          tmem_ld16(reg0 + c * 16, r0[c]);
          if (NG == 2) tmem_ld16(reg0 + BN1 + c * 16, r1[NG == 2 ? c : 0]);
        }
        tmem_ld_wait();
        if (tracer) CAREL_TRACE(4, trace_e);
        const int j0 = t * BN1 + ctx.half * R;
        pol.begin_tile(ctx, j0, R);
#pragma unroll
        for (int c = 0; c < R / 16; ++c) {
          float w[16];
          pol.weights16(ctx, j0 + c * 16, r0[c], r1[NG == 2 ? c : 0], w);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (F16) {
              hi[i] = pack_f16x2_sat(w[2 * i], w[2 * i + 1]);
              if (HILO) {
                const float2 f = unpack_f16x2(hi[i]);
                lo[i] = pack_f16x2(w[2 * i] - f.x, w[2 * i + 1] - f.y);
              }
            } else {
              const __nv_bfloat162 b = __floats2bfloat162_rn(w[2 * i], w[2 * i + 1]);
              hi[i] = *reinterpret_cast<const uint32_t*>(&b);
              if (HILO) {
                const float2 f = __bfloat1622float2(b);
                lo[i] = pack_bf16x2(w[2 * i] - f.x, w[2 * i + 1] - f.y);
              }
            }
          }
          tmem_st8(reg0 + c * 8, hi);
          if (HILO) tmem_st8(reg0 + R / 2 + c * 8, lo);
          pol.store16(ctx, j0 + c * 16, hi);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
        if (tracer) { CAREL_TRACE(5, trace_e); ++trace_e; }
        if (++buf == NB) buf = 0;
      }
      mbar_wait(acc2_full, acc2_phase);
      acc2_phase ^= 1;
      tc_fence_after();
      // the policy calls release() once, as soon as it no longer reads acc2 (e.g. after moving it to registers)
      pol.final(ctx, it.t1 > it.t0, [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty);
      });
    }
    pol.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpi) tmem_dealloc(tmem_base, 512);
}

}  // namespace carel
