// gemm_2sm.cuh — persistent tcgen05 GEMM on CTA PAIRS (cluster of 2, cta_group::2) for long-K products:
//
//   D[256 x 256] (fp32; 128 rows in the TMEM of each CTA of the pair) = A_tile[256 x K] * B_tile[256 x K]^T
//
// One UMMA (M = 256, N = 256, K = 16) spans the two SMs of a TPC.  Each CTA loads ITS 128 rows of A and ITS 128 rows
// of B per 64-wide k-block (32 KB per stage instead of the 48 KB a 128 x 256 single-CTA tile needs), so the pair
// moves 64 KB per k-block for 2 x the flops: the L2 -> SM traffic per flop is 2/3 of gemm_tc's and the ring is 7 stages
// deep instead of 4.  (gemm_tc at K = 784: 602 KB of operands per tile, tensor pipe 60 % busy, ring too shallow for
// the TMA latency under load.)
// Roles per CTA as in gemm_tc (16 epilogue warps, TMA producer warp, MMA warp); only the LEADER CTA (rank 0) issues
// MMAs.  Barriers:
//   full[s]    leader only, 1 arrival (leader's expect_tx of the bytes of BOTH CTAs) — both producers' TMA loads
//              complete on it (cta_group::2 loads signal the barrier of rank 0);
//   empty[s]   one per CTA, 1 arrival: the leader's tcgen05.commit multicast to both CTAs;
//   tfull[b]   one per CTA, 1 arrival: commit multicast (accumulator buffer b complete in both TMEMs);
//   tempty[b]  leader only, 2 x 16 arrivals: every epilogue warp of both CTAs (the partner's arrive remotely).
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace carel {

struct PairSched {
  int p_tile0, p_tiles;   // pair-row tiles (256 rows) [p_tile0, p_tile0 + p_tiles) of this launch
  int n_tiles;            // 256-column tiles
  int tri;                // 1: only tiles on / above the diagonal (tn >= pair-row index)
  int num_k16;            // K in 16-wide steps
  int ksplit;             // split-K factor (items = pair tiles x ksplit; the epilogue then accumulates atomically)
  int total;              // scheduled items
  int clusters;           // persistent grid = 2 * clusters CTAs
};

constexpr int kPairStages = 7;
constexpr int kPairStageBytes = 2 * 16384;       // A: 128 x 64 bf16, B: 128 x 64 bf16 (this CTA's halves)
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + 1024 + 256;
constexpr int kPairBand = 4;                     // pair-row tiles per band (column tile outermost inside a band)

// Tile list in bands of kPairBand pair rows, column-major inside a band (L2 reuse of the column operand at large N);
// every role of both CTAs of a cluster walks the same items.
struct PairIter {
  const PairSched& s;
  int item, band, band_base;
  int pm, tn, kb0, kb1;
  CAREL_DEVINL explicit PairIter(const PairSched& s_) : s(s_), item(blockIdx.x >> 1), band(0), band_base(0) {}
  CAREL_DEVINL int first_tn(int p) const { return s.tri ? (s.p_tile0 + p) : 0; }
  CAREL_DEVINL int band_count(int b) const {
    const int r0 = b * kPairBand, r1 = min(r0 + kPairBand, s.p_tiles);
    int c = 0;
    for (int r = r0; r < r1; ++r) c += s.n_tiles - first_tn(r);
    return c;
  }
  CAREL_DEVINL bool next() {
    if (item >= s.total) return false;
    const int tile = item / s.ksplit, ks = item - tile * s.ksplit;
    int c = band_count(band);
    while (tile >= band_base + c) {
      band_base += c;
      ++band;
      c = band_count(band);
    }
    int l = tile - band_base;
    const int r0 = band * kPairBand, nr = min(kPairBand, s.p_tiles - r0), last = r0 + nr - 1;
    const int fl = first_tn(last), cmin = s.n_tiles - fl;
    if (l < nr * cmin) {
      const int q = l / nr;
      tn = fl + q;
      pm = r0 + (l - q * nr);
    } else {
      l -= nr * cmin;
      pm = last;
      tn = fl;
      for (int r = r0; r <= last; ++r) {
        const int cnt = fl - first_tn(r);
        if (l < cnt) { pm = r; tn = first_tn(r) + l; break; }
        l -= cnt;
      }
    }
    pm += s.p_tile0;
    const int num_kb = (s.num_k16 + 3) >> 2, per = (num_kb + s.ksplit - 1) / s.ksplit;
    kb0 = ks * per;
    kb1 = min(kb0 + per, num_kb);
    item += s.clusters;
    return true;
  }
};

template <class Epi, bool A_MN = false, bool B_MN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ PairSched sched, const __grid_constant__ typename Epi::Params ep) {
  constexpr int BN = 256, STAGES = kPairStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kPairStageBytes);
  uint64_t* full_bar = bars;                     // [STAGES]  (used in the leader)
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]       (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kEpi = 4 * kColSplit;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == (kEpi + 1) * 32) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * 4 * kColSplit);
    }
    fence_mbar_init();
  }
  cluster_sync_all();                              // the partner's barriers exist before anything signals them
  if (warp == kEpi) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpi && elect_one()) {
    // ===================== TMA producer (both CTAs: own 128 rows of A, own 128 rows of B) =====================
    PairIter it(sched);
    int stage = 0;
    uint32_t phase = 0;
    while (it.next()) {
#pragma unroll 1
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * kPairStageBytes;
        uint8_t* sB = sA + 16384;
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
        const int m0 = it.pm * 256 + int(rank) * 128, n0 = it.tn * BN + int(rank) * 128;
        if (!A_MN) {
          tma_load_2d_2sm(sA, &tmA, &full_bar[stage], kb * 64, m0);
        } else {                     // MN-major: two [64 k x 64 m] chunks
          tma_load_2d_2sm(sA, &tmA, &full_bar[stage], m0, kb * 64);
          tma_load_2d_2sm(sA + 8192, &tmA, &full_bar[stage], m0 + 64, kb * 64);
        }
        if (!B_MN) {
          tma_load_2d_2sm(sB, &tmB, &full_bar[stage], kb * 64, n0);
        } else {
          tma_load_2d_2sm(sB, &tmB, &full_bar[stage], n0, kb * 64);
          tma_load_2d_2sm(sB + 8192, &tmB, &full_bar[stage], n0 + 64, kb * 64);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kEpi + 1 && leader && elect_one()) {
    // ===================== MMA issuer (leader CTA only) =====================
    PairIter it(sched);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t bphase = 0;
    constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    while (it.next()) {
      mbar_wait(&tempty_bar[buf], bphase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * BN;
#pragma unroll 1
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * kPairStageBytes);
        const uint64_t da0 = A_MN ? make_sdesc_sw128(a_base, 8192, 1024) : make_sdesc_sw128(a_base, 16, 1024);
        const uint64_t db0 = B_MN ? make_sdesc_sw128(a_base + 16384, 8192, 1024) : make_sdesc_sw128(a_base + 16384, 16, 1024);
        const int nk = min(4, sched.num_k16 - kb * 4);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16)
          if (k16 < nk)
            umma_ss_2sm(d_addr, sdesc_advance(da0, k16 * (A_MN ? 2048 : 32)), sdesc_advance(db0, k16 * (B_MN ? 2048 : 32)),
                        idesc, (kb > it.kb0 || k16 > 0) ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage], 0b11);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(&tfull_bar[buf], 0b11);
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
  } else if (warp < kEpi) {
    // ===================== epilogue (each CTA: its own 128 rows) =====================
    const int q = warp & 3;
    TileCtx ctx;
    ctx.half = warp >> 2;
    ctx.epi_tid = threadIdx.x;
    Epi epi(ep, ctx);
    PairIter it(sched);
    int buf = 0;
    uint32_t bphase = 0;
    while (it.next()) {
      ctx.tm = it.pm * 2 + int(rank);
      ctx.tn = it.tn;
      ctx.row = ctx.tm * 128 + q * 32 + lane;
      ctx.col0 = it.tn * BN;
      ctx.kb0 = it.kb0;
      ctx.kb1 = it.kb1;
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      ctx.taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * BN;
      epi.tile(ctx);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[buf]);
        else mbar_arrive_cluster(&tempty_bar[buf], 0);
      }
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
    epi.finish();
  }

  tc_fence_before();
  cluster_sync_all();                              // nobody exits while the partner may still signal / read it
  if (warp == kEpi) tmem_dealloc_2sm(tmem_base, 512);
}

inline PairSched make_pair_sched(int64_t M, int64_t N, int64_t K, int tri, int row_tile0, int row_tile1, int sms,
                                 int ksplit = 1) {
  PairSched s;
  s.p_tile0 = row_tile0 / 2;
  s.p_tiles = (row_tile1 - row_tile0 + 1) / 2;
  s.n_tiles = int((N + 255) / 256);
  s.tri = tri;
  s.num_k16 = int((K + 15) / 16);
  const int num_kb = (s.num_k16 + 3) / 4;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > num_kb) ksplit = num_kb;
  const int per = (num_kb + ksplit - 1) / ksplit;
  s.ksplit = (num_kb + per - 1) / per;          // no empty splits
  s.total = 0;
  for (int p = 0; p < s.p_tiles; ++p) s.total += s.n_tiles - (tri ? (s.p_tile0 + p) : 0);
  s.total *= s.ksplit;
  int cl = sms / 2;
  if (s.total < cl) cl = s.total;
  if (cl < 1) cl = 1;
  s.clusters = cl;
  (void)M;
  return s;
}

}  // namespace carel
