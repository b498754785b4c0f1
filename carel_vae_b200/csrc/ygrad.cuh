// ygrad.cuh — decoder backward from the STORED logits (one GEMM per gradient, no recompute):
//
//   Y [N, V] (16-bit fixed point, one scale per row) holds the centred log2-domain logits written once by the forward
//   pass (pass F).  Both gradients are  acc += A * B  with  A = the dlogit tile, formed IN PLACE in shared memory over
//   the TMA-loaded Y tile (same bytes, same 128-byte-swizzled layout: every epilogue thread owns one clause row and
//   32 consecutive words of it in both kernels), and read by the tensor core straight from shared memory:
//   TRANS = false (dZ): owner = 256 clauses (two M = 128 accumulators), step = 64 words:
//                       A = dl [clauses x words]  K-major,   B = W_d rows of the 64 words (MN-major)
//   TRANS = true  (dW): owner = 128 words, step = 128 clauses:
//                       A = dl^T: the same tile read MN-major (M = words), B = Z rows of the clauses (MN-major),
//                       db = a second MMA (N = 16) of the same A against a tile of ones
//   No TMEM round trip for the weights, no transposed shared-memory reads; TMEM holds accumulators only.
//
// Roles: 16 epilogue warps, TMA producer warp, MMA issuer warp (kChainThreads).  Rings: Y 3 x 32 KB (TMA -> epilogue
// in place -> MMA -> free), B pieces of 64 reduction rows (ldk / 64 x 8 KB each, up to 4).  Work split: stream-K units
// of (owner, step) pairs (ChainSched).
//
// Per-step constants (b_j log2 e of the step's words; dW: also the two rowc planes of its 128 clauses) ride in the Y
// ring as 1-D bulk copies on the same barrier, so the epilogue loop has no global load.
//
// dlogit (dense part; the O(nnz) sparse part is added by the gather kernels of bow.cu):
//   log2 p = (float(Y16) - bmagic_i) inv_s_i + rc_i + cc_j,   rc_i = -lse2_i,   cc_j = b_j log2(e)
//   (rowc[i] = {inv_s_i, bmagic_i, rc_i, -S_i 2^10}: the four per-clause constants in one 16-byte load)
//   d' = (u - p S_i) 2^10,      u = p (p - c) / max(p (1 - p), 1e-12)        (torch's BCELoss backward rule)
//   series (p in (e^-27, 2^-10)):  u = (p - c)(1 + p + p^2), i.e.  u - p S = -c + p (1 - c - S + p (1 - c + p)).
//   Rows the forward pass certified (rowc plane 1, .w = 1: every p of the row in (e^-27, 2^-8)) evaluate it as a polynomial in
//   p' = 2^(log2 p - rc_i) with per-row coefficients (a1, a2, a3) = 2^10 ((1 - c - S) P, (1 - c) P^2, P^3), P = 2^rc_i
//   (the cubic term, < 2^-20 of the linear one, is dropped): one FFMA2 for the exponent, 2 MUFU and 2 FFMA2 per PAIR
//   of elements, no range test.  A warp with an uncertified
//   row tests every 16-element chunk instead and falls back to the exact terms where needed.
// The 2^10 puts d' (typically +-3e-5 before scaling) into fp16's normal range; the write-out multiplies the
// accumulator by grad_out / (N V) 2^-10.
#pragma once
#include "chain_tc.cuh"

namespace carel {

constexpr float kYgScale = 1024.f;
constexpr int kYgMaxYStages = 5, kYgMaxBStages = 4;   // ring depths are chosen per launch (yg_stages)
constexpr int kYgYBytes = 32768;                // one Y piece: [256 clauses x 64 words] or 2 x [128 clauses x 64 words]
constexpr int kYgOnesBytes = 1024;              // [8 rows x 128 B] of fp16 1.0: B operand of the db MMA (N = 16: the
                                                // second 8-row group aliases the first, SBO = 0)
// per-stage constants: b log2 e of the 128 owner words (dW, 512 B) or of the step's 64 words (dZ, 256 B).  The two
// rowc planes of a dW step's clauses are prefetched into registers one step ahead instead (8 KB per stage would cost
// the fourth B stage).
static __host__ __device__ constexpr int yg_c_bytes(bool trans) { return trans ? 512 : 256; }

struct YGradParams {
  const float4* rowc;    // [Npad]  {inv_s_i, bmagic_i, -lse2_i, -S_i 2^10}  (zero beyond N)
  int Npad;              // rows of one rowc plane; plane 1 = {a1, a2, a3, fast} (see above)
  const float* bl2e;     // [Vpad]  b_j log2(e)
  const float* grad_out; // device scalar
  float* out;            // TRANS: dW [V, ldo]; else dZ [N, ldo]   (accumulated with red.global.add)
  float* gb;             // TRANS only: db [V]
  int64_t ldo;
  int N, V, K;           // K = feature columns written out (<= 16 * k16)
  float c, inv_nv;
  int y_stages;          // Y ring depth (the B ring depth travels in ChainSched::stages)
  int probe;             // timing probes only (results are garbage): 1 = no B loads, 2 = no Y loads (tools/yg_sweep.sh)
};

// one B piece: 64 reduction rows x ldk features = ldk / 64 boxes of [64 rows x 128 B]
static __host__ __device__ inline int yg_b_bytes(int k16) { return ((k16 + 3) / 4) * 8192; }
// Ring depths (Y, B) from the 227 KB budget.  The Y pieces come from HBM (measured ~3100 clk from issue to landing
// under load), the B pieces from L2: the Y ring gets the depth first.  CAREL_YG_STAGES="y,b" overrides (experiments).
static __host__ inline void yg_stages(int k16, bool trans, int* ys, int* bs) {
  const int fixed = kYgOnesBytes + 256;
  // measured at N = 8192, V = 30000, K = 256 (tools/yg_sweep.sh): dZ (Y, B) = (3,4) 158.6, (4,3) 149.0, (5,2) 159.4 us;
  // dW consumes two B pieces per step and needs four of them in flight: (3,4) 182.0, (3,3) 205.2, (4,2) 215.8 us
  int y = trans ? 3 : 4, b = trans ? 4 : 3;
  if (const char* e = getenv(trans ? "CAREL_YG_STAGES_DW" : "CAREL_YG_STAGES_DZ")) {
    int ey = 0, eb = 0;
    if (sscanf(e, "%d,%d", &ey, &eb) == 2 && ey >= 2 && ey <= kYgMaxYStages && eb >= 2 && eb <= kYgMaxBStages) { y = ey; b = eb; }
  }
  while (y > 2 && fixed + y * (kYgYBytes + yg_c_bytes(trans)) + 2 * yg_b_bytes(k16) > 227 * 1024) --y;
  const int left = 227 * 1024 - fixed - y * (kYgYBytes + yg_c_bytes(trans));
  int bmax = left / yg_b_bytes(k16);
  if (bmax > kYgMaxBStages) bmax = kYgMaxBStages;
  if (b > bmax) b = bmax;
  *ys = y;
  *bs = b;
}
static __host__ inline int yg_smem_bytes(int k16, bool trans, int ys, int bs) {
  return ys * (kYgYBytes + yg_c_bytes(trans)) + bs * yg_b_bytes(k16) + kYgOnesBytes + 256;
}

// exact per-element terms (chunks that leave the series' range)
static CAREL_DEVINL float yg_exact_u(float x, float c) {
  const float p = __expf(x);
  return p * (p - c) / fmaxf(p * (1.f - p), 1e-12f);
}

// Stored logit: 16-bit fixed point with a per-row scale, written by pass F as the low 16 bits of
//   fl(acc a_i + bmagic_i),  a_i = log2(e) s_i,  bmagic_i = 2^23 + 32768 - rm_i s_i  (a float in [2^23, 2^24): ulp 1).
// Decoding puts the 16 bits back under the exponent of 2^23 and subtracts bmagic_i (exact), then scales by 1 / s_i:
// no integer <-> float conversion instruction on either side.  lo / hi = the two logits of a packed pair.
static CAREL_DEVINL float y16_lo(uint32_t u) { return __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7610)); }
static CAREL_DEVINL float y16_hi(uint32_t u) { return __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7632)); }
static CAREL_DEVINL float y16_one(uint16_t q) { return __uint_as_float(0x4B000000u | uint32_t(q)); }

// d'[16] from t[16] = log2 p and Sn[16] = -S 2^10 (per element)
static CAREL_DEVINL void yg_dlogit16(const float (&t)[16], const float (&Sn)[16], float c, bool in_range, int n_valid,
                                     float (&d)[16]) {
  float mx = fmax3(t[0], t[1], t[2]), mn = fmin3(t[0], t[1], t[2]);
#pragma unroll
  for (int i = 3; i < 15; i += 2) { mx = fmax3(mx, t[i], t[i + 1]); mn = fmin3(mn, t[i], t[i + 1]); }
  mx = fmaxf(mx, t[15]);
  mn = fminf(mn, t[15]);
  if (in_range && mx < -10.f && mn > -27.f * kLog2e) {
    const float2 one = make_float2(1.f, 1.f);
    const float2 sc = make_float2(kYgScale, kYgScale), mcs = make_float2(-c * kYgScale, -c * kYgScale);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 pv = make_float2(ex2_approx(t[i]), ex2_approx(t[i + 1]));
      const float2 q = __ffma2_rn(pv, __fadd2_rn(pv, one), one);              // 1 + p + p^2
      const float2 a = __ffma2_rn(pv, sc, mcs);                               // (p - c) 2^10
      const float2 m = __fmul2_rn(pv, make_float2(Sn[i], Sn[i + 1]));         // -p S 2^10
      const float2 o = __ffma2_rn(a, q, m);
      d[i] = o.x;
      d[i + 1] = o.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = t[i] * kLn2;
      d[i] = (i < n_valid) ? fmaf(yg_exact_u(x, c), kYgScale, __expf(x) * Sn[i]) : 0.f;
    }
  }
}

// EP: epilogue timing probes, instantiated only under -DCAREL_YG_PROBES (results are garbage): 1 = no MUFU,
// 2 = no shared-memory store of the dlogits, 4 = no shared-memory load of the logits, 8 = no arithmetic at all
template <bool TRANS, int EP = 0>
__global__ void __launch_bounds__(kChainThreads, 1)
ygrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ ChainSched sched, const __grid_constant__ YGradParams pp) {
  static_assert(kChainSplit == 4, "16 epilogue warps: 512 threads x 32 elements = one 32 KB piece of 16-bit logits");
  const int K16 = sched.k16;
  const int KB = (K16 + 3) / 4;
  const int b_bytes = yg_b_bytes(K16);
  const int BST = sched.stages;                   // B ring depth
  const int YST = pp.y_stages;                    // Y ring depth
  constexpr int kCB = yg_c_bytes(TRANS);
  // the base of dynamic shared memory is 1024-byte aligned (no static shared memory in this kernel); checked below
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sY = smem;
  uint8_t* sB = smem + YST * kYgYBytes;
  uint8_t* sOnes = sB + BST * b_bytes;
  uint8_t* sC = sOnes + kYgOnesBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + YST * kCB);
  uint64_t* y_full = bars;                        // [5]  TMA bytes (Y piece + its constants)
  uint64_t* y_empty = y_full + kYgMaxYStages;     // [5]  tcgen05.commit (the MMAs that read the piece are done)
  uint64_t* p_full = y_empty + kYgMaxYStages;     // [5]  16 epilogue warps (dlogits written in place)
  uint64_t* b_full = p_full + kYgMaxYStages;      // [4]
  uint64_t* b_empty = b_full + kYgMaxBStages;     // [4]  tcgen05.commit
  uint64_t* acc_full = b_empty + kYgMaxBStages;   // [1]
  uint64_t* acc_empty = acc_full + 1;             // [1]  16 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kEpi = 16;

  if (threadIdx.x == (kEpi + 1) * 32) {
    if (smem_u32(smem) & 1023u) {
      printf("carel: ygrad_kernel needs 1024-byte aligned dynamic shared memory\n");
      __trap();
    }
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kYgMaxYStages; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); mbar_init(&p_full[i], kEpi); }
    for (int i = 0; i < kYgMaxBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpi);
    fence_mbar_init();
  }
  if (TRANS) {       // the tile of ones behind the db MMA (any layout: every element is 1.0)
    for (int i = threadIdx.x; i < kYgOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;
    fence_proxy_async_smem();
  }
  if (warp == kEpi) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kDbCol = 256;                // TRANS: TMEM columns [256, 272) = db (every column the same sum)

  if (warp == kEpi && elect_one()) {
    // ===================== TMA producer =====================
    ChainIter iter(sched);
    ChainItem it;
    int ys = 0, bs = 0;
    uint32_t yphase = 0, bphase = 0;
    auto load_b = [&](int row0) {
      mbar_wait(&b_empty[bs], bphase ^ 1);
      if (pp.probe & 1) {
        mbar_arrive(&b_full[bs]);
      } else {
        mbar_expect_tx(&b_full[bs], b_bytes);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + bs * b_bytes + kb * 8192, &tmB, &b_full[bs], kb * 64, row0);
      }
      if (++bs == BST) { bs = 0; bphase ^= 1; }
    };
    int tn = 0;
    while (iter.next(it)) {
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t, ++tn) {
        mbar_wait(&y_empty[ys], yphase ^ 1);
        CAREL_TRACE(0, tn);
        mbar_expect_tx(&y_full[ys], ((pp.probe & 2) ? 0 : kYgYBytes) + kCB);
        uint8_t* cst = sC + ys * kCB;
        if (pp.probe & 2) {
          bulk_load_1d(cst, pp.bl2e + (TRANS ? it.owner * 128 : t * 64), kCB, &y_full[ys]);
        } else if (TRANS) {          // two boxes [128 clauses x 64 words]; rowc planes of the clauses; b of the owner's words
          tma_load_2d(sY + ys * kYgYBytes, &tmY, &y_full[ys], it.owner * 128, t * 128);
          tma_load_2d(sY + ys * kYgYBytes + 16384, &tmY, &y_full[ys], it.owner * 128 + 64, t * 128);
          bulk_load_1d(cst, pp.bl2e + it.owner * 128, 512, &y_full[ys]);
        } else {              // one box [256 clauses x 64 words]; b of the step's words
          tma_load_2d(sY + ys * kYgYBytes, &tmY, &y_full[ys], t * 64, it.owner * 256);
          bulk_load_1d(cst, pp.bl2e + t * 64, 256, &y_full[ys]);
        }
        if (++ys == YST) { ys = 0; yphase ^= 1; }
        if (TRANS) {
          load_b(t * 128);
          load_b(t * 128 + 64);
        } else {
          load_b(t * 64);
        }
        CAREL_TRACE(1, tn);
      }
    }
  } else if (warp == kEpi + 1 && elect_one()) {
    // ===================== MMA issuer =====================
    ChainIter iter(sched);
    ChainItem it;
    int ys = 0, bs = 0, tn = 0;
    uint32_t yphase = 0, bphase = 0, acc_phase = 0;
    const uint32_t idesc = make_idesc_f16(128, 16 * K16, TRANS ? 1 : 0, 1);     // B MN-major (features contiguous)
    const uint32_t idesc_ones = make_idesc_f16(128, 16, 1, 0);
    const uint64_t d_ones = make_sdesc_sw128(smem_u32(sOnes), 16, 0);
    while (iter.next(it)) {
      mbar_wait(acc_empty, acc_phase ^ 1);
      acc_phase ^= 1;
      uint32_t accum = 0;
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t, ++tn) {
        mbar_wait(&p_full[ys], yphase);
        CAREL_TRACE(4, tn);
        tc_fence_after();
        const uint32_t y_addr = smem_u32(sY + ys * kYgYBytes);
        if (TRANS) {
          // A = dl^T, MN-major: two 64-word chunks 16 KB apart, 8-clause groups 1 KB apart, 16 clauses = 2 KB per MMA
          const uint64_t da0 = make_sdesc_sw128(y_addr, 16384, 1024);
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            mbar_wait(&b_full[bs], bphase);
            if (hb == 1) CAREL_TRACE(5, tn);
            tc_fence_after();
            const uint64_t db0 = make_sdesc_sw128(smem_u32(sB + bs * b_bytes), 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = sdesc_advance(da0, (hb * 4 + k) * 2048);
              umma_ss(tmem_base, da, sdesc_advance(db0, k * 2048), idesc, accum);
              umma_ss(tmem_base + kDbCol, da, d_ones, idesc_ones, accum);
              accum = 1;
            }
            umma_commit(&b_empty[bs]);
            if (++bs == BST) { bs = 0; bphase ^= 1; }
          }
        } else {
          // A = dl, K-major: [256 clauses x 64 words], rows 128 B, the second accumulator's rows 16 KB further
          const uint64_t da0 = make_sdesc_sw128(y_addr, 16, 1024);
          mbar_wait(&b_full[bs], bphase);
          CAREL_TRACE(5, tn);
          tc_fence_after();
          const uint64_t db0 = make_sdesc_sw128(smem_u32(sB + bs * b_bytes), 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = sdesc_advance(db0, k * 2048);
            umma_ss(tmem_base, sdesc_advance(da0, k * 32), db, idesc, accum);
            umma_ss(tmem_base + 256, sdesc_advance(da0, 16384 + k * 32), db, idesc, accum);
            accum = 1;
          }
          umma_commit(&b_empty[bs]);
          if (++bs == BST) { bs = 0; bphase ^= 1; }
        }
        umma_commit(&y_empty[ys]);
        CAREL_TRACE(6, tn);
        if (++ys == YST) { ys = 0; yphase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (warp < kEpi) {
    // ===================== epilogue =====================
    // this thread's clause row inside the piece and its 64-byte segment (32 words) of that row
    const int r = TRANS ? ((warp & 3) * 32 + lane) : ((warp & 7) * 32 + lane);
    const int seg = TRANS ? (warp >> 2) : (warp >> 3);            // TRANS: 0..3 (box = seg / 2), else 0..1
    const int tile_off = (TRANS ? (seg >> 1) * 16384 : 0) + r * 128;
    const int half = seg & 1;
    const int q = warp & 3, part = warp >> 2;                     // write-out: TMEM lane quarter, column part
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    ChainIter iter(sched);
    ChainItem it;
    int ys = 0, tn = 0;
    uint32_t yphase = 0, acc_phase = 0;
    const float c = pp.c;
    const float2 mc2 = make_float2(-c * kYgScale, -c * kYgScale);
    while (iter.next(it)) {
      float4 own = make_float4(0.f, 0.f, 0.f, 0.f), pol = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 own_n = own, pol_n = pol;
      if (!TRANS) {
        const int clause = it.owner * 256 + r;
        if (clause < pp.N) {
          own = __ldg(pp.rowc + clause);
          pol = __ldg(pp.rowc + pp.Npad + clause);
        }
      } else if (it.t0 < it.t1) {                 // rowc is padded to a multiple of 128 rows (zeros)
        own_n = __ldg(pp.rowc + it.t0 * 128 + r);
        pol_n = __ldg(pp.rowc + pp.Npad + it.t0 * 128 + r);
      }
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t, ++tn) {
        const int clause = TRANS ? t * 128 + r : it.owner * 256 + r;
        const int j0 = TRANS ? it.owner * 128 + seg * 32 : t * 64 + seg * 32;      // first word of this thread
        mbar_wait(&y_full[ys], yphase);
        if (threadIdx.x == 0) CAREL_TRACE(2, tn);
        uint8_t* rowp = sY + ys * kYgYBytes + tile_off;
        const uint8_t* cst = sC + ys * kCB;
        if (TRANS) {            // this step's clause constants were loaded a step ago; fetch the next step's
          own = own_n;
          pol = pol_n;
          const int tn1 = min(t + 1, it.t1 - 1) * 128 + r;
          own_n = __ldg(pp.rowc + tn1);
          pol_n = __ldg(pp.rowc + pp.Npad + tn1);
        }
        const float4* cc4 = reinterpret_cast<const float4*>(cst) + seg * 8;   // this thread's 32 words
        const float2 nb2 = make_float2(-own.y, -own.y), is2 = make_float2(own.x, own.x);
        if (__all_sync(0xffffffffu, pol.w != 0.f)) {
          // ---- every row of this warp is certified: cubic in p', no range test (words beyond V meet zero-filled
          //      operand rows or unwritten output rows; their values only have to be finite) ----
          // (certified rows have every p < 2^-8; the p'^3 term is dropped, see bow_combine_kernel)
          const float2 a1 = make_float2(pol.x, pol.x), a2 = make_float2(pol.y, pol.y);
          // The dlogits overwrite the logits they came from, so the compiler must keep every shared-memory load
          // behind the stores written before it: all four 16-byte pieces of the row segment are loaded up front
          // (and the word constants one block ahead) instead of one load -> ... -> store chain per block, whose
          // load latency was exposed four times per step (ncu: 25 % of the kernel's stall samples).
          uint4 vy[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if constexpr (EP & 4) vy[i] = make_uint4(0x12345678u + i, 0x23456789u, 0x3456789au, 0x456789abu + r);
            else vy[i] = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + i) ^ (r & 7)) << 4));
          }
          float4 ca = cc4[0], cb = cc4[1];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint8_t* cp = rowp + (((half * 4 + i) ^ (r & 7)) << 4);
            const uint4 v = vy[i];
            float4 ca_n = ca, cb_n = cb;
            if (i < 3) { ca_n = cc4[2 * i + 2]; cb_n = cc4[2 * i + 3]; }
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
            const float2 cc[4] = {make_float2(ca.x, ca.y), make_float2(ca.z, ca.w), make_float2(cb.x, cb.y),
                                  make_float2(cb.z, cb.w)};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (EP & 8) {
                o[e] = u[e] ^ __float_as_uint(cc[e].x);
              } else {
                const float2 g = __fadd2_rn(make_float2(y16_lo(u[e]), y16_hi(u[e])), nb2);
                const float2 tp = __ffma2_rn(g, is2, cc[e]);
                const float2 pv = (EP & 1) ? tp : make_float2(ex2_approx(tp.x), ex2_approx(tp.y));
                float2 hh = __ffma2_rn(pv, a2, a1);
                hh = __ffma2_rn(pv, hh, mc2);
                o[e] = pack_f16x2_sat(hh.x, hh.y);
              }
            }
            if constexpr (EP & 2) {
              if ((o[0] ^ o[1] ^ o[2] ^ o[3]) == 0x9e3779b9u) *reinterpret_cast<uint4*>(cp) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
              *reinterpret_cast<uint4*>(cp) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            ca = ca_n;
            cb = cb_n;
          }
        } else {
          // ---- checked path: range test per 16 elements, exact terms where the series does not hold ----
          const bool row_dead = clause >= pp.N;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t yh[8];
            float tt[16], Sn[16], d[16];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint4 v = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + h * 2 + i) ^ (r & 7)) << 4));
              yh[4 * i] = v.x; yh[4 * i + 1] = v.y; yh[4 * i + 2] = v.z; yh[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = cc4[h * 4 + i];
              const uint32_t u0 = yh[2 * i], u1 = yh[2 * i + 1];
              const float2 g0 = __fadd2_rn(make_float2(y16_lo(u0), y16_hi(u0)), nb2);     // exact: integers below 2^24
              const float2 g1 = __fadd2_rn(make_float2(y16_lo(u1), y16_hi(u1)), nb2);
              const float2 t0 = __ffma2_rn(g0, is2, make_float2(a4.x + own.z, a4.y + own.z));
              const float2 t1 = __ffma2_rn(g1, is2, make_float2(a4.z + own.z, a4.w + own.z));
              tt[4 * i] = t0.x; tt[4 * i + 1] = t0.y; tt[4 * i + 2] = t1.x; tt[4 * i + 3] = t1.y;
              Sn[4 * i] = own.w; Sn[4 * i + 1] = own.w; Sn[4 * i + 2] = own.w; Sn[4 * i + 3] = own.w;
            }
            const int jh = j0 + h * 16;
            if (row_dead) {
#pragma unroll
              for (int i = 0; i < 16; ++i) d[i] = 0.f;
            } else {
              yg_dlogit16(tt, Sn, c, jh + 16 <= pp.V, pp.V - jh, d);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
              *reinterpret_cast<uint4*>(rowp + (((half * 4 + h * 2 + i) ^ (r & 7)) << 4)) =
                  make_uint4(pack_f16x2_sat(d[8 * i], d[8 * i + 1]), pack_f16x2_sat(d[8 * i + 2], d[8 * i + 3]),
                             pack_f16x2_sat(d[8 * i + 4], d[8 * i + 5]), pack_f16x2_sat(d[8 * i + 6], d[8 * i + 7]));
          }
        }
        // ---- the dlogits lie over the logits they came from; make them visible to the tensor core (async proxy) ----
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[ys]);
        if (threadIdx.x == 0) CAREL_TRACE(3, tn);
        if (threadIdx.x == 15 * 32) CAREL_TRACE(7, tn);
        if (++ys == YST) { ys = 0; yphase ^= 1; }
      }
      // ---- write-out of this item's accumulator(s) ----
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      {
        const float oscale = __ldg(pp.grad_out) * pp.inv_nv * (1.f / kYgScale);
        const int n16 = (pp.K + 15) / 16;
        const int q0 = (n16 * part) / 4, q1 = (n16 * (part + 1)) / 4;
        const int lim_rows = TRANS ? pp.V : pp.N;
        const bool vec = ((pp.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(pp.out) & 15) == 0);
        constexpr int NACC = TRANS ? 1 : 2;
        float dbv = 0.f;
        if (TRANS && part == 0) dbv = __uint_as_float(tmem_ld1(tmem_base + lane_off + kDbCol));
#pragma unroll 1
        for (int mh = 0; mh < NACC; ++mh) {
          const int orow = TRANS ? it.owner * 128 + q * 32 + lane : it.owner * 256 + mh * 128 + q * 32 + lane;
          const bool row_ok = (it.t1 > it.t0) && orow < lim_rows;
          float* dst = pp.out + size_t(min(orow, lim_rows - 1)) * pp.ldo;
          const uint32_t acc_addr = tmem_base + lane_off + mh * 256;
#pragma unroll 1
          for (int qq = q0; qq < q1; qq += 2) {
            uint32_t ra[16], rb[16];
            const bool two = qq + 1 < q1;
            tmem_ld16(acc_addr + qq * 16, ra);
            if (two) tmem_ld16(acc_addr + (qq + 1) * 16, rb);
            tmem_ld_wait();
            if (!row_ok) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              const uint32_t (&rr)[16] = h ? rb : ra;
              float* o = dst + (qq + h) * 16;
              const int f0 = (qq + h) * 16;
              if (vec && f0 + 16 <= pp.K) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  red_add_v4(o + i, __uint_as_float(rr[i]) * oscale, __uint_as_float(rr[i + 1]) * oscale,
                             __uint_as_float(rr[i + 2]) * oscale, __uint_as_float(rr[i + 3]) * oscale);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (f0 + i < pp.K) atomicAdd(o + i, __uint_as_float(rr[i]) * oscale);
              }
            }
          }
          if (TRANS && part == 0 && row_ok && dbv != 0.f) atomicAdd(pp.gb + orow, dbv * oscale);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpi) tmem_dealloc(tmem_base, 512);
}

template <bool TRANS>
int launch_ygrad(const CUtensorMap& ty, const CUtensorMap& tb, const ChainSched& s, const YGradParams& pp,
                 cudaStream_t st, const char* tag) {
  auto kern = ygrad_kernel<TRANS>;
#ifdef CAREL_YG_PROBES
  switch (pp.probe >> 2) {
    case 1: kern = ygrad_kernel<TRANS, 1>; break;
    case 2: kern = ygrad_kernel<TRANS, 2>; break;
    case 4: kern = ygrad_kernel<TRANS, 4>; break;
    case 8: kern = ygrad_kernel<TRANS, 8>; break;
    case 14: kern = ygrad_kernel<TRANS, 14>; break;
    default: break;
  }
  CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#endif
  const int smem = yg_smem_bytes(s.k16, TRANS, pp.y_stages, s.stages);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    CAREL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  if (s.num_items <= 0) return 0;
  timing_begin(tag, st);
  kern<<<s.num_items, kChainThreads, smem, st>>>(ty, tb, s, pp);
  timing_end(st);
  CAREL_LAUNCHED();
  return 0;
}

}  // namespace carel
