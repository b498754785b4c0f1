// ygrad.cuh — decoder backward from the STORED logits (one GEMM per gradient, no recompute):
//
//   Y [N, V] (16-bit fixed point, one scale per row) holds the centred log2-domain logits written once by the forward
//   pass (pass F); for an owner tile of 128 output rows and every inner tile t of 128 reduction rows
//     P     = dlogit tile (from Y and the per-row / per-column constants) -> fp16, written to TMEM
//     acc  += P * B_t                      (A = P from TMEM, B = the streamed operand tile, MN-major)
//   TRANS = false (dZ): owner = 128 clauses (TMEM lane = clause), inner = 128 words,   B_t = W_d rows of the words
//   TRANS = true  (dW): owner = 128 words   (TMEM lane = word),   inner = 128 clauses, B_t = Z rows of the clauses,
//                       db = thread-local column sum of the dlogits
//   The Y tile is the same [128 clauses x 128 words] box in both cases; TRANS reads it column-wise from shared memory.
//
// Roles as in chain_tc.cuh: 16 epilogue warps (warp w: TMEM lanes 32 (w % 4).., inner part w / 4 = 32 reduction
// rows), TMA producer warp, MMA issuer warp.  Rings: Y 3 x 32 KB, B 2 x (ldk / 64) x 16 KB; TMEM: acc 256 columns,
// P 4 buffers x 64 columns (128 fp16 per lane).  Work split: stream-K units of (owner, inner) pairs (ChainSched).
//
// dlogit (dense part; the O(nnz) sparse part is added by the gather kernels of bow.cu):
//   log2 p = (float(Y16) - bmagic_i) inv_s_i + rc_i + cc_j,   rc_i = -lse2_i,   cc_j = b_j log2(e)
//   (rowc[i] = {inv_s_i, bmagic_i, rc_i, -S_i 2^10}: the four per-clause constants in one 16-byte load)
//   d' = (u - p S_i) 2^10,      u = p (p - c) / max(p (1 - p), 1e-12)        (torch's BCELoss backward rule)
//   fast path (all 16 p in (e^-27, 2^-10)):  u = (p - c)(1 + p + p^2)
// The 2^10 puts d' (typically +-3e-5 before scaling) into fp16's normal range; the write-out multiplies the
// accumulator by grad_out / (N V) 2^-10.
#pragma once
#include "chain_tc.cuh"

namespace carel {

constexpr float kYgScale = 1024.f;
constexpr int kYgPBufs = 4;
constexpr int kYgYStages = 3, kYgBStages = 2;
constexpr int kYgYBytes = 128 * 128 * 2;        // one Y tile: two 64-column boxes of [128 rows x 128 B]

struct YGradParams {
  const float4* rowc;    // [Npad]  {inv_s_i, bmagic_i, -lse2_i, -S_i 2^10}  (zero beyond N)
  const float* bl2e;     // [Vpad]  b_j log2(e)
  const float* grad_out; // device scalar
  float* out;            // TRANS: dW [V, ldo]; else dZ [N, ldo]   (accumulated with red.global.add)
  float* gb;             // TRANS only: db [V]
  int64_t ldo;
  int N, V, K;           // K = feature columns written out (<= 16 * k16)
  float c, inv_nv;
};

static __host__ __device__ inline int yg_b_bytes(int k16) { return ((k16 + 3) / 4) * 16384; }
static __host__ inline int yg_smem_bytes(int k16) {
  return kYgYStages * kYgYBytes + kYgBStages * yg_b_bytes(k16) + 1024 + 512;
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

template <bool TRANS>
__global__ void __launch_bounds__(kChainThreads, 1)
ygrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ ChainSched sched, const __grid_constant__ YGradParams pp) {
  static_assert(kChainSplit == 4, "the epilogue splits the 128 reduction rows of an inner tile into 4 parts of 32");
  const int K16 = sched.k16;
  const int KB = (K16 + 3) / 4;
  const int b_bytes = yg_b_bytes(K16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sY = smem;
  uint8_t* sB = smem + kYgYStages * kYgYBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kYgBStages * b_bytes);
  uint64_t* y_full = bars;                        // [3]
  uint64_t* y_empty = y_full + kYgYStages;        // [3]  16 epilogue warps
  uint64_t* b_full = y_empty + kYgYStages;        // [2]
  uint64_t* b_empty = b_full + kYgBStages;        // [2]  tcgen05.commit
  uint64_t* p_full = b_empty + kYgBStages;        // [4]  16 epilogue warps
  uint64_t* p_empty = p_full + kYgPBufs;          // [4]  tcgen05.commit
  uint64_t* acc_full = p_empty + kYgPBufs;        // [1]
  uint64_t* acc_empty = acc_full + 1;             // [1]  16 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kEpi = 16;

  if (threadIdx.x == (kEpi + 1) * 32) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kYgYStages; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], kEpi); }
    for (int i = 0; i < kYgBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < kYgPBufs; ++i) { mbar_init(&p_full[i], kEpi); mbar_init(&p_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpi);
    fence_mbar_init();
  }
  if (warp == kEpi) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kPBase = 256;                // TMEM column of P buffer 0 (acc: columns [0, 256))

  if (warp == kEpi && elect_one()) {
    // ===================== TMA producer =====================
    ChainIter iter(sched);
    ChainItem it;
    int ys = 0, bs = 0;
    uint32_t yphase = 0, bphase = 0;
    while (iter.next(it)) {
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        // Y tile: clauses x words, whichever of the two is the owner
        const int yrow = (TRANS ? t : it.owner) * 128, ycol = (TRANS ? it.owner : t) * 128;
        mbar_wait(&y_empty[ys], yphase ^ 1);
        mbar_expect_tx(&y_full[ys], kYgYBytes);
        tma_load_2d(sY + ys * kYgYBytes, &tmY, &y_full[ys], ycol, yrow);
        tma_load_2d(sY + ys * kYgYBytes + 16384, &tmY, &y_full[ys], ycol + 64, yrow);
        if (++ys == kYgYStages) { ys = 0; yphase ^= 1; }
        mbar_wait(&b_empty[bs], bphase ^ 1);
        mbar_expect_tx(&b_full[bs], b_bytes);
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + bs * b_bytes + kb * 16384, &tmB, &b_full[bs], kb * 64, t * 128);
        if (++bs == kYgBStages) { bs = 0; bphase ^= 1; }
      }
    }
  } else if (warp == kEpi + 1 && elect_one()) {
    // ===================== MMA issuer =====================
    ChainIter iter(sched);
    ChainItem it;
    int bs = 0, buf = 0;
    uint32_t bphase = 0, pphases = 0, acc_phase = 0;
    const uint32_t idesc = make_idesc_f16(128, 16 * K16, 0, 1);      // A = P (TMEM, K-major), B MN-major
    while (iter.next(it)) {
      mbar_wait(acc_empty, acc_phase ^ 1);
      acc_phase ^= 1;
      bool first = true;
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        mbar_wait(&b_full[bs], bphase);
        mbar_wait(&p_full[buf], (pphases >> buf) & 1u);
        pphases ^= 1u << buf;
        tc_fence_after();
        const uint64_t db0 = make_sdesc_sw128(smem_u32(sB + bs * b_bytes), 16384, 1024);   // 64-column chunks 16 KB apart
        const uint32_t p_addr = tmem_base + kPBase + buf * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          umma_ts(tmem_base, p_addr + k * 8, sdesc_advance(db0, k * 2048), idesc, first ? 0u : 1u);
          first = false;
        }
        umma_commit(&b_empty[bs]);
        umma_commit(&p_empty[buf]);
        if (++bs == kYgBStages) { bs = 0; bphase ^= 1; }
        if (++buf == kYgPBufs) buf = 0;
      }
      umma_commit(acc_full);
    }
  } else if (warp < kEpi) {
    // ===================== epilogue =====================
    const int q = warp & 3, part = warp >> 2;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    ChainIter iter(sched);
    ChainItem it;
    int ys = 0, buf = 0;
    uint32_t yphase = 0, pe_phases = 0, acc_phase = 0;
    const float c = pp.c;
    while (iter.next(it)) {
      const int orow = it.owner * 128 + q * 32 + lane;          // this thread's output row (clause or word)
      // per-thread constants: TRANS: the word's bias; else the clause's {inv_s, bmagic, rc, Sn}
      float own_cc = 0.f, dbsum = 0.f;
      float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
      if (TRANS) {
        own_cc = __ldg(pp.bl2e + min(orow, pp.V - 1));          // this thread's word: b_v log2(e)
      } else if (orow < pp.N) {
        own = __ldg(pp.rowc + orow);
      }
#pragma unroll 1
      for (int t = it.t0; t < it.t1; ++t) {
        const int i0 = t * 128 + part * 32;                     // first reduction index of this thread's 32
        // !TRANS: the per-word constants of the first 16 columns are issued before the barrier wait (their latency
        // hides behind it), those of the second 16 when the first half is done.  TRANS: the per-clause float4
        // constants are uniform across the warp (one L1 transaction each) and are loaded where they are used.
        float4 ca[4];
        auto load_consts = [&](int h) {
          if (!TRANS) {
            const float4* b4 = reinterpret_cast<const float4*>(pp.bl2e + i0 + h * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) ca[i] = __ldg(b4 + i);
          }
        };
        load_consts(0);
        mbar_wait(&y_full[ys], yphase);
        // ---- this thread's 32 logits out of the swizzled tile ----
        const uint8_t* tile = sY + ys * kYgYBytes;
        uint32_t yh[16];
        if (!TRANS) {
          const int r = q * 32 + lane;
          const uint8_t* rowp = tile + (part >> 1) * 16384 + r * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = ((part & 1) * 4 + i) ^ (r & 7);
            const uint4 v = *reinterpret_cast<const uint4*>(rowp + ch * 16);
            yh[4 * i] = v.x; yh[4 * i + 1] = v.y; yh[4 * i + 2] = v.z; yh[4 * i + 3] = v.w;
          }
        } else {
          const int w = q * 32 + lane, wc = w & 63;
          const uint8_t* colp = tile + (w >> 6) * 16384 + (wc & 7) * 2;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = part * 32 + 2 * i, c1 = c0 + 1;
            const uint32_t lo = *reinterpret_cast<const uint16_t*>(colp + c0 * 128 + (((wc >> 3) ^ (c0 & 7)) << 4));
            const uint32_t hi = *reinterpret_cast<const uint16_t*>(colp + c1 * 128 + (((wc >> 3) ^ (c1 & 7)) << 4));
            yh[i] = lo | (hi << 16);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[ys]);
        if (++ys == kYgYStages) { ys = 0; yphase ^= 1; }
        // ---- dlogits ----
        uint32_t ph[16];
        const int lim = TRANS ? pp.N : pp.V;
        const bool row_dead = TRANS ? (orow >= pp.V) : (orow >= pp.N);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1) load_consts(1);
          float tt[16], Sn[16], d[16];
          if (!TRANS) {
            const float2 nb2 = make_float2(-own.y, -own.y), is2 = make_float2(own.x, own.x);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = ca[i];
              const uint32_t u0 = yh[h * 8 + 2 * i], u1 = yh[h * 8 + 2 * i + 1];
              const float2 g0 = __fadd2_rn(make_float2(y16_lo(u0), y16_hi(u0)), nb2);     // exact: integers below 2^24
              const float2 g1 = __fadd2_rn(make_float2(y16_lo(u1), y16_hi(u1)), nb2);
              const float2 t0 = __ffma2_rn(g0, is2, make_float2(a4.x + own.z, a4.y + own.z));
              const float2 t1 = __ffma2_rn(g1, is2, make_float2(a4.z + own.z, a4.w + own.z));
              tt[4 * i] = t0.x; tt[4 * i + 1] = t0.y; tt[4 * i + 2] = t1.x; tt[4 * i + 3] = t1.y;
              Sn[4 * i] = own.w; Sn[4 * i + 1] = own.w; Sn[4 * i + 2] = own.w; Sn[4 * i + 3] = own.w;
            }
          } else {
            const float4* rc4 = pp.rowc + i0 + h * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r0 = __ldg(rc4 + 2 * i), r1 = __ldg(rc4 + 2 * i + 1);
              const uint32_t u = yh[h * 8 + i];
              tt[2 * i] = fmaf(y16_lo(u) - r0.y, r0.x, r0.z + own_cc);
              tt[2 * i + 1] = fmaf(y16_hi(u) - r1.y, r1.x, r1.z + own_cc);
              Sn[2 * i] = r0.w;
              Sn[2 * i + 1] = r1.w;
            }
          }
          const int j0 = i0 + h * 16;
          if (row_dead) {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = 0.f;
          } else {
            yg_dlogit16(tt, Sn, c, j0 + 16 <= lim, lim - j0, d);
          }
          if (TRANS) {
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; i += 2) s2 = __fadd2_rn(s2, make_float2(d[i], d[i + 1]));
            dbsum += s2.x + s2.y;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) ph[h * 8 + i] = pack_f16x2_sat(d[2 * i], d[2 * i + 1]);
        }
        // ---- P -> TMEM (the MMA that read this buffer four tiles ago has completed) ----
        mbar_wait(&p_empty[buf], ((pe_phases >> buf) & 1u) ^ 1u);
        pe_phases ^= 1u << buf;
        tc_fence_after();
        {
          const uint32_t dst = tmem_base + lane_off + kPBase + buf * 64 + part * 16;
          uint32_t lo16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) lo16[i] = ph[i];
          tmem_st16(dst, lo16);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
        if (++buf == kYgPBufs) buf = 0;
      }
      // ---- write-out of this item's accumulator ----
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      {
        const float oscale = __ldg(pp.grad_out) * pp.inv_nv * (1.f / kYgScale);
        const int n16 = (pp.K + 15) / 16;
        const int q0 = (n16 * part) / 4, q1 = (n16 * (part + 1)) / 4;
        const int lim_rows = TRANS ? pp.V : pp.N;
        const bool row_ok = (it.t1 > it.t0) && orow < lim_rows;
        float* dst = pp.out + size_t(min(orow, lim_rows - 1)) * pp.ldo;
        const bool vec = ((pp.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(pp.out) & 15) == 0);
        const uint32_t acc_addr = tmem_base + lane_off;
        bool released = false;
        auto release = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          released = true;
        };
#pragma unroll 1
        for (int qq = q0; qq < q1; qq += 2) {
          uint32_t ra[16], rb[16];
          const bool two = qq + 1 < q1;
          tmem_ld16(acc_addr + qq * 16, ra);
          if (two) tmem_ld16(acc_addr + (qq + 1) * 16, rb);
          tmem_ld_wait();
          if (qq + 2 >= q1) release();              // last read of the accumulator by this thread
          if (!row_ok) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const uint32_t (&r)[16] = h ? rb : ra;
            float* o = dst + (qq + h) * 16;
            const int f0 = (qq + h) * 16;
            if (vec && f0 + 16 <= pp.K) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                red_add_v4(o + i, __uint_as_float(r[i]) * oscale, __uint_as_float(r[i + 1]) * oscale,
                           __uint_as_float(r[i + 2]) * oscale, __uint_as_float(r[i + 3]) * oscale);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (f0 + i < pp.K) atomicAdd(o + i, __uint_as_float(r[i]) * oscale);
            }
          }
        }
        if (!released) release();
        if (TRANS && row_ok && dbsum != 0.f) atomicAdd(pp.gb + orow, dbsum * oscale);
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
  const int smem = yg_smem_bytes(s.k16);
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
