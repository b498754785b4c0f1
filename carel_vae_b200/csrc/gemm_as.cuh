// gemm_as.cuh — A-stationary persistent tcgen05 GEMM for short-K problems (K <= 256):
//
//   D[128 x 256] (fp32, TMEM) = A_tile[128 x K] * B_tile[256 x K]^T        (bf16, both K-major)
//
// gemm_tc_kernel re-loads BOTH operand tiles for every output tile; at K = 256 that is 192 KB per
// 128 x 256 tile and the kernel is bound by the L2 -> SM fabric (~42 B/clk/SM), not by the tensor pipe.
// Here a CTA walks a CONTIGUOUS range of the row-major tile order, so consecutive tiles share the row tile:
// A (<= 64 KB) stays resident in shared memory and only B (<= 128 KB per tile, in 64-wide k-blocks through a
// ring) is streamed.  Same roles / TMEM double buffering / epilogue interface (TileCtx) as gemm_tc_kernel.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace carel {

struct AsSched {
  int m_tile0, m_tiles;   // row tiles [m_tile0, m_tile0 + m_tiles) of 128 rows (row-block sharding across GPUs)
  int n_tiles;            // 256-column tiles
  int tri;                // 1: only tiles on / above the diagonal of a square problem (tn >= tm / 2)
  int num_k16;            // K in 16-wide steps (<= 16)
  int total;              // number of scheduled tiles
  int tpc;                // tiles per CTA: CTA b owns tiles [b * tpc, (b + 1) * tpc) of the row-major tile list
  int n_cols;             // logical N (the last column tile issues a narrower MMA)
  const int* row_tile_flags;   // optional (device, full mode): flags[0] = any work at all, flags[1 + tm] per row tile
};

#ifndef CAREL_AS_STAGES
#define CAREL_AS_STAGES 5
#endif
constexpr int kAsStages = CAREL_AS_STAGES;
constexpr int kAsBytesA = 4 * 16384;                     // 4 k-blocks of [128 x 64] bf16
constexpr int kAsStageB = 256 * 128;                     // one k-block of [256 x 64] bf16
constexpr int kAsSmemBytes = kAsBytesA + kAsStages * kAsStageB + 1024 + 256;
// With an epilogue staging area (EPI_SMEM bytes, e.g. tiles that are TMA-stored from shared memory) the B ring
// gives up a stage.
__host__ __device__ constexpr int as_stages(int epi_smem) { return epi_smem > 0 ? kAsStages - 1 : kAsStages; }
__host__ __device__ constexpr int as_smem_bytes(int epi_smem) { return kAsBytesA + as_stages(epi_smem) * kAsStageB + epi_smem + 1024 + 256; }

// Walks the CTA's range of the tile list; every role uses the same iterator.  `new_row` is true on the first
// tile of a row tile (A has to be (re)loaded), `last_of_row` on the last one this CTA processes for it.
struct AsIter {
  const AsSched& s;
  int remaining, cur_tm, cur_tn, prev_tm;
  int tm, tn;
  bool new_row, last_of_row;
  CAREL_DEVINL int first_tn(int m) const { return s.tri ? (m >> 1) : 0; }
  CAREL_DEVINL explicit AsIter(const AsSched& s_) : s(s_), prev_tm(-1) {
    int t = blockIdx.x * s.tpc;
    remaining = min(s.tpc, s.total - t);
    if (s.row_tile_flags != nullptr && s.row_tile_flags[0] == 0) remaining = 0;
    cur_tm = s.m_tile0;
    cur_tn = first_tn(cur_tm);
    if (remaining > 0) {
      if (!s.tri) {
        cur_tm = s.m_tile0 + t / s.n_tiles;
        cur_tn = t % s.n_tiles;
      } else {
        while (t >= s.n_tiles - first_tn(cur_tm)) {       // at most m_tiles steps, once per CTA
          t -= s.n_tiles - first_tn(cur_tm);
          ++cur_tm;
        }
        cur_tn = first_tn(cur_tm) + t;
      }
    }
  }
  CAREL_DEVINL bool next() {
    if (s.row_tile_flags != nullptr) {                     // skip unflagged row tiles
      while (remaining > 0 && s.row_tile_flags[1 + cur_tm] == 0) {
        remaining -= s.n_tiles - cur_tn;
        ++cur_tm;
        cur_tn = first_tn(cur_tm);
      }
    }
    if (remaining <= 0) return false;
    tm = cur_tm;
    tn = cur_tn;
    new_row = tm != prev_tm;
    prev_tm = tm;
    --remaining;
    if (++cur_tn == s.n_tiles) {
      ++cur_tm;
      cur_tn = first_tn(cur_tm);
    }
    last_of_row = (remaining <= 0) || (cur_tm != tm);
    return true;
  }
};

// F16: operands are fp16 instead of bf16 (same MMA rate, 11 mantissa bits).  EPI_SMEM: bytes of shared memory handed
// to the epilogue (TileCtx::epi_smem, 1024-byte aligned).
template <class Epi, bool F16 = false, int EPI_SMEM = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_as_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ AsSched sched, const __grid_constant__ typename Epi::Params ep) {
  constexpr int BN = 256, STAGES = as_stages(EPI_SMEM);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kAsBytesA;
  uint8_t* sEpi = sB + STAGES * kAsStageB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + EPI_SMEM);
  uint64_t* b_full = bars;                       // [STAGES]
  uint64_t* b_empty = bars + STAGES;             // [STAGES]
  uint64_t* a_full = bars + 2 * STAGES;          // [1]
  uint64_t* a_empty = a_full + 1;                // [1]
  uint64_t* tfull_bar = a_empty + 1;             // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kEpi = 4 * kColSplit;
  const int KB = (sched.num_k16 + 3) >> 2;

  if (threadIdx.x == (kEpi + 1) * 32) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kColSplit);
    }
    fence_mbar_init();
  }
  if (warp == kEpi) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpi && elect_one()) {
    // ===================== TMA producer =====================
    AsIter it(sched);
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    int trace_n = 0; (void)trace_n;
    while (it.next()) {
      CAREL_TRACE(0, trace_n); ++trace_n;
      if (it.new_row) {
        mbar_wait(a_empty, aphase ^ 1);
        aphase ^= 1;
        mbar_expect_tx(a_full, KB * 16384);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * 16384, &tmA, a_full, kb * 64, it.tm * 128);
      }
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&b_empty[stage], phase ^ 1);
        uint8_t* dst = sB + stage * kAsStageB;
        mbar_expect_tx(&b_full[stage], kAsStageB);
        tma_load_2d(dst, &tmB, &b_full[stage], kb * 64, it.tn * BN);
        tma_load_2d(dst + 16384, &tmB, &b_full[stage], kb * 64, it.tn * BN + 128);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kEpi + 1 && elect_one()) {
    // ===================== MMA issuer =====================
    AsIter it(sched);
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    int buf = 0;
    uint32_t bphase = 0;
    const uint64_t da0 = make_sdesc_sw128(smem_u32(sA), 16, 1024);
    int trace_n = 0; (void)trace_n;
    while (it.next()) {
      const int n_eff = min(BN, ((sched.n_cols - it.tn * BN + 15) >> 4) << 4);
      const uint32_t idesc = F16 ? make_idesc_f16(128, n_eff, 0, 0) : make_idesc_bf16(128, n_eff, 0, 0);
      if (it.new_row) {
        mbar_wait(a_full, aphase);
        aphase ^= 1;
      }
      mbar_wait(&tempty_bar[buf], bphase ^ 1);
      CAREL_TRACE(1, trace_n);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * BN;
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&b_full[stage], phase);
        tc_fence_after();
        const uint64_t da = sdesc_advance(da0, kb * 16384);
        const uint64_t db = make_sdesc_sw128(smem_u32(sB + stage * kAsStageB), 16, 1024);
        const int nk = min(4, sched.num_k16 - kb * 4);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16)
          if (k16 < nk) umma_ss(d_addr, sdesc_advance(da, k16 * 32), sdesc_advance(db, k16 * 32), idesc, (kb | k16) ? 1u : 0u);
        umma_commit(&b_empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      CAREL_TRACE(2, trace_n); ++trace_n;
      umma_commit(&tfull_bar[buf]);
      if (it.last_of_row) umma_commit(a_empty);     // every MMA that reads this A tile has been issued
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
  } else if (warp < kEpi) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    TileCtx ctx;
    ctx.half = warp >> 2;
    ctx.epi_tid = threadIdx.x;
    ctx.epi_smem = sEpi;
    Epi epi(ep, ctx);
    AsIter it(sched);
    int buf = 0;
    uint32_t bphase = 0;
    int trace_n = 0; (void)trace_n;
    const bool tracer = threadIdx.x == 0; (void)tracer;
    while (it.next()) {
      ctx.tm = it.tm;
      ctx.tn = it.tn;
      ctx.row = it.tm * 128 + q * 32 + lane;
      ctx.col0 = it.tn * BN;
      ctx.kb0 = 0;
      ctx.kb1 = KB;
      ctx.first_of_row = it.new_row;
      ctx.last_of_row = it.last_of_row;
      epi.prefetch(ctx);                               // loads that do not depend on the accumulator
      mbar_wait(&tfull_bar[buf], bphase);
      if (tracer) CAREL_TRACE(3, trace_n);
      tc_fence_after();
      ctx.taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * BN;
      epi.tile(ctx);
      if (tracer) { CAREL_TRACE(4, trace_n); ++trace_n; }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpi) tmem_dealloc(tmem_base, 512);
}

inline AsSched make_as_sched(int64_t M, int64_t N, int64_t K, int sms, int tri = 0, int m_tile0 = 0, int m_tile1 = -1) {
  AsSched s;
  const int all_m = int((M + 127) / 128);
  if (m_tile1 < 0 || m_tile1 > all_m) m_tile1 = all_m;
  s.m_tile0 = m_tile0;
  s.m_tiles = m_tile1 > m_tile0 ? m_tile1 - m_tile0 : 0;
  s.n_tiles = int((N + 255) / 256);
  s.tri = tri;
  s.num_k16 = int((K + 15) / 16);
  s.total = 0;
  for (int m = m_tile0; m < m_tile0 + s.m_tiles; ++m) s.total += s.n_tiles - (tri ? (m >> 1) : 0);
  int grid = s.total < sms ? s.total : sms;
  if (grid < 1) grid = 1;
  s.tpc = (s.total + grid - 1) / grid;
  if (s.tpc < 1) s.tpc = 1;
  s.n_cols = int(N);
  s.row_tile_flags = nullptr;
  return s;
}

}  // namespace carel
