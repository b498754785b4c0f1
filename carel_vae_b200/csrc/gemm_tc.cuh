// gemm_tc.cuh — persistent, warp-specialised tcgen05 GEMM core with pluggable epilogues.
//
//   D[128 x BN] (fp32, TMEM) = A_tile[128 x K] * B_tile[BN x K]^T       (bf16 operands)
//
// One CTA per SM, 320 threads:
//   warps 0..   epilogue: 4*kColSplit warps; warp w reads TMEM lanes 32*(w%4).. and column part w/4
//   next warp   TMEM allocator, then one lane = TMA producer
//   last warp   one lane = tcgen05.mma issuer   (highest warp ids win the SMSP arbitration)
//
// Three pipelines: smem ring (full/empty mbarriers, TMA <-> MMA), TMEM accumulator double
// buffer (tmem_full/tmem_empty, MMA <-> epilogue), and the static persistent tile schedule.
//
// K is counted in 16-wide MMA steps (num_k16), so "augmented" operands whose width is not a
// multiple of 64 (e.g. 128 features + 16 norm/ones columns) cost only the extra steps: TMA
// zero-fills the rest of the last 64-wide box and no MMA is issued for it.
//
// Operand layouts (all 128-byte swizzled, loaded by TMA):
//   K-major  : global row-major [rows, K]; smem tile = [rows][64 bf16] per 64-wide k-block.
//   MN-major : global row-major [K, rows]; smem tile = per 64-row chunk [64 k][64 bf16].
// NG = 1 or 2 independent products per tile (HSIC needs the x- and the y-Gram of the same
// tile); product g accumulates into TMEM columns [g*BN, (g+1)*BN).
#pragma once
#include "common.cuh"

namespace carel {

// Epilogue column split: kColSplit groups of 4 warps, each group owning BN / kColSplit columns of the
// accumulator tile (4 = 16 epilogue warps, 4 per SM sub-partition, to hide MUFU / TMEM-load latency).
#ifndef CAREL_GEMM_COLSPLIT
#define CAREL_GEMM_COLSPLIT 4
#endif
constexpr int kColSplit = CAREL_GEMM_COLSPLIT;
constexpr int kEpiThreads = 128 * kColSplit;
constexpr int kGemmThreads = 64 + kEpiThreads;
#ifndef CAREL_CHAIN_COLSPLIT
#define CAREL_CHAIN_COLSPLIT 4
#endif
constexpr int kChainSplit = CAREL_CHAIN_COLSPLIT;   // column parts of an inner tile (4 epilogue warps each)
constexpr int kChainThreads = 64 + 128 * kChainSplit;

enum TileMode : int { TILE_FULL = 0, TILE_TRI = 1 };

struct GemmSched {
  int m_tiles;    // tiles of 128 rows
  int n_tiles;    // tiles of BN cols
  int num_k16;    // 16-wide k steps (per product)
  int ksplit;     // split-K factor (>=1); item = tile * ksplit + ks
  int mode;       // TileMode; TILE_TRI requires BN == 256 and a square problem
  int num_items;  // total work items
  int m_tile0;    // first row tile owned by this launch (row-block sharding across GPUs)
  int tri_rows;   // number of row tiles of this launch (m_tiles unless a row block [m_tile0, m_tile0 + tri_rows) is set)
  int n_cols;     // if > 0: logical N; the last column tile issues a narrower MMA (multiple of 16)
  const int* row_tile_flags;  // optional (device): flags[0] = any work at all, flags[1 + tm] = process row tile tm
};

// Sequential decoder of the persistent schedule; every role of a CTA walks the same items.
// TILE_TRI orders the tiles in BANDS of kTriBand row tiles, column tile outermost inside a band: the CTAs that run
// concurrently then share ~148 / kTriBand column tiles across kTriBand row tiles, so a column-operand tile fetched
// from HBM is reused kTriBand times out of L2.  (Row-major order streams the whole column operand once per row
// tile: at N = 131072, d = 768 that is 2048 x 411 MB and the Gram kernel becomes HBM-bound.)
constexpr int kTriBand = 8;
struct TileIter {
  const GemmSched& s;
  int item;
  int band, band_base;
  int tm, tn, kb0, kb1;
  CAREL_DEVINL explicit TileIter(const GemmSched& s_)
      : s(s_), item(blockIdx.x), band(0), band_base(0) {
    if (s.row_tile_flags != nullptr && s.row_tile_flags[0] == 0) item = s.num_items;   // nothing flagged
  }
  CAREL_DEVINL int first_tn(int t) const { return (s.m_tile0 + t) >> 1; }
  CAREL_DEVINL int band_count(int b) const {
    const int r0 = b * kTriBand, r1 = min(r0 + kTriBand, s.tri_rows);
    int c = 0;
    for (int r = r0; r < r1; ++r) c += s.n_tiles - first_tn(r);
    return c;
  }
  CAREL_DEVINL bool next() {
    for (;;) {
      if (!decode()) return false;
      if (s.row_tile_flags == nullptr || s.row_tile_flags[1 + tm] != 0) return true;
    }
  }
  CAREL_DEVINL bool decode() {
    if (item >= s.num_items) return false;
    const int tile = item / s.ksplit;
    const int ks = item - tile * s.ksplit;
    if (s.mode == TILE_FULL) {
      // bands of kTriBand row tiles, column tile outermost inside a band (same L2 argument as below)
      const int band_tiles = kTriBand * s.n_tiles;
      const int b = tile / band_tiles, l = tile - b * band_tiles;
      const int r0 = b * kTriBand, nr = min(kTriBand, s.tri_rows - r0);
      tn = l / nr;
      tm = r0 + (l - tn * nr);
    } else {
      int c = band_count(band);
      while (tile >= band_base + c) {
        band_base += c;
        ++band;
        c = band_count(band);
      }
      int l = tile - band_base;
      const int r0 = band * kTriBand, nr = min(kTriBand, s.tri_rows - r0), last = r0 + nr - 1;
      const int fl = first_tn(last), cmin = s.n_tiles - fl;
      if (l < nr * cmin) {            // common columns of the band: column-major
        const int q = l / nr;
        tn = fl + q;
        tm = r0 + (l - q * nr);
      } else {                        // the few near-diagonal tiles a row has beyond the band's common columns
        l -= nr * cmin;
        tm = last;
        tn = fl;
        for (int r = r0; r <= last; ++r) {
          const int cnt = fl - first_tn(r);
          if (l < cnt) { tm = r; tn = first_tn(r) + l; break; }
          l -= cnt;
        }
      }
    }
    tm += s.m_tile0;
    const int num_kb = (s.num_k16 + 3) >> 2;
    const int per = (num_kb + s.ksplit - 1) / s.ksplit;
    kb0 = ks * per;
    kb1 = min(kb0 + per, num_kb);
    item += gridDim.x;
    return true;
  }
};

// Context handed to an epilogue for one accumulator tile.
struct TileCtx {
  uint32_t taddr;  // TMEM address of (this warp's lane quarter, column 0 of the buffer)
  int tm, tn;      // tile coordinates
  int row;         // global row index of this thread (tm*128 + lane quarter*32 + lane)
  int col0;        // global column of tile column 0 (tn*BN)
  int half;        // column part handled by this warp: columns [half*BN/kColSplit, (half+1)*BN/kColSplit)
  int kb0, kb1;    // k-block range that was accumulated (split-K)
  int epi_tid;     // 0..255 within the epilogue warps
  bool first_of_row, last_of_row;   // A-stationary kernel only: first / last tile this CTA runs on row tile tm
  uint8_t* epi_smem;               // A-stationary kernel only: EPI_SMEM bytes of shared memory for the epilogue (or unset)
};

template <int BN, int NG>
struct GemmCfg {
  static constexpr int kStageA = 128 * 128;  // 16 KB
  static constexpr int kStageB = BN * 128;   // 16/32 KB
  static constexpr int kStageBytes = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 512;
  static constexpr int kBufCols = NG * BN;   // columns per accumulator buffer
  static_assert(2 * kBufCols <= 512, "TMEM budget");
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <class Epi, int BN, int NG, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ GemmSched sched,
               const __grid_constant__ typename Epi::Params ep) {
  using Cfg = GemmCfg<BN, NG>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Roles: warps [0, kEpi) epilogue; warp kEpi = TMEM allocator + TMA producer lane; warp kEpi + 1 = MMA issuer
  // lane.  The two single-lane roles sit in the HIGHEST warp ids on purpose: the SMSP arbiter prefers higher
  // warp ids, and an MMA issuer that has to queue behind spinning / MUFU-heavy epilogue warps starves the
  // tensor pipe (measured: narrow MMAs took 109 clk per issue instead of 55).
  constexpr int kEpi = 4 * kColSplit;

  if (threadIdx.x == (kEpi + 1) * 32) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NG == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kColSplit);
    }
    fence_mbar_init();
  }
  if (warp == kEpi) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpi && elect_one()) {   // elect.sync: ptxas knows ONE lane runs this region, so the
    // uniform-register operands of UTMALDG / UTCHMMA need no per-instruction broadcast loop
    // ===================== TMA producer =====================
    TileIter it(sched);
    int stage = 0;
    uint32_t phase = 0;
    while (it.next()) {
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        const CUtensorMap* mA = g ? &tmA1 : &tmA0;
        const CUtensorMap* mB = g ? &tmB1 : &tmB0;
#pragma unroll 1
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kStageA;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sA, mA, &full_bar[stage], kb * 64, it.tm * 128);
          } else {
            tma_load_2d(sA, mA, &full_bar[stage], it.tm * 128, kb * 64);
            tma_load_2d(sA + 8192, mA, &full_bar[stage], it.tm * 128 + 64, kb * 64);
          }
          if (!B_MN) {
#pragma unroll
            for (int nb = 0; nb < BN / 128; ++nb)
              tma_load_2d(sB + nb * 16384, mB, &full_bar[stage], kb * 64, it.tn * BN + nb * 128);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sB + c * 8192, mB, &full_bar[stage], it.tn * BN + c * 64, kb * 64);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kEpi + 1 && elect_one()) {
    // ===================== MMA issuer =====================
    TileIter it(sched);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t bphase = 0;
    while (it.next()) {
      int n_eff = BN;
      if (sched.n_cols > 0) n_eff = min(BN, ((sched.n_cols - it.tn * BN + 15) >> 4) << 4);
      const uint32_t idesc = make_idesc_bf16(128, n_eff, A_MN ? 1 : 0, B_MN ? 1 : 0);
      mbar_wait(&tempty_bar[buf], bphase ^ 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        const uint32_t d_addr = tmem_base + buf * Cfg::kBufCols + g * BN;
#pragma unroll 1
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da0 = A_MN ? make_sdesc_sw128(a_base, 8192, 1024) : make_sdesc_sw128(a_base, 16, 1024);
          const uint64_t db0 = B_MN ? make_sdesc_sw128(a_base + Cfg::kStageA, 8192, 1024)
                                    : make_sdesc_sw128(a_base + Cfg::kStageA, 16, 1024);
          const int nk = min(4, sched.num_k16 - kb * 4);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            if (k16 < nk)
              umma_ss(d_addr, sdesc_advance(da0, k16 * (A_MN ? 2048 : 32)),
                      sdesc_advance(db0, k16 * (B_MN ? 2048 : 32)), idesc, (kb > it.kb0 || k16 > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(&tfull_bar[buf]);
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
  } else if (warp < kEpi) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    TileCtx ctx;
    ctx.half = warp >> 2;
    ctx.epi_tid = threadIdx.x;
    Epi epi(ep, ctx);
    TileIter it(sched);
    int buf = 0;
    uint32_t bphase = 0;
    while (it.next()) {
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      ctx.taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::kBufCols;
      ctx.tm = it.tm;
      ctx.tn = it.tn;
      ctx.row = it.tm * 128 + q * 32 + lane;
      ctx.col0 = it.tn * BN;
      ctx.kb0 = it.kb0;
      ctx.kb1 = it.kb1;
      epi.tile(ctx);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpi) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// Software-pipelined walk over an accumulator row segment in chunks of 16 fp32 columns: the
// tcgen05.ld of chunk c+1 is in flight while chunk c is processed.  The loop is kept ROLLED
// (body = two chunks, ping-pong registers) so that the hot loop stays inside the ~6 KB L0
// instruction cache.  f(c, regs16) is called once per chunk; n16 must be even.
template <class F>
CAREL_DEVINL void for_each_chunk16(uint32_t taddr, int n16, F&& f) {
  uint32_t ra[16], rb[16];
  tmem_ld16(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < n16; c += 2) {
    tmem_ld_wait();
    tmem_ld16(taddr + (c + 1) * 16, rb);
    f(c, ra);
    tmem_ld_wait();
    if (c + 2 < n16) tmem_ld16(taddr + (c + 2) * 16, ra);
    f(c + 1, rb);
  }
}

// Same walk in chunks of 32 columns (tcgen05.ld 32x32b.x32): half as many TMEM loads / waits per tile, for
// epilogues light enough to keep two 32-register buffers.  n32 must be even.
template <class F>
CAREL_DEVINL void for_each_chunk32(uint32_t taddr, int n32, F&& f) {
  uint32_t ra[32], rb[32];
  tmem_ld32(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < n32; c += 2) {
    tmem_ld_wait();
    tmem_ld32(taddr + (c + 1) * 32, rb);
    f(c, ra);
    tmem_ld_wait();
    if (c + 2 < n32) tmem_ld32(taddr + (c + 2) * 32, ra);
    f(c + 1, rb);
  }
}

}  // namespace carel
