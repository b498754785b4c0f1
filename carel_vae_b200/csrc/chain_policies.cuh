// chain_policies.cuh — per-loss policies of chain_tc_kernel (what P is, what acc2 becomes).
#pragma once
#include "chain_tc.cuh"
#include "epilogues.cuh"

namespace carel {

// ---------------------------------------------------------------------------------------------
// MMD backward.  With C the symmetric coefficient matrix (a00 / a11 on the within-sample
// off-diagonal, a01 on both cross blocks, 0 on the diagonal) and
//   K'_ij = sum_alpha -alpha exp(-alpha (|d2_ij| + eps)) sign(d2_ij),   W = C (.) K'
// the gradient of the reference MMD (drl_classifier_ec_mmd_final_mul.py:547-569 through
// pdist :580-589) w.r.t. the stacked sample is
//   grad s_i = 4 [ (sum_j W_ij) s_i - sum_j W_ij s_j ].
// Operands are augmented (pack_rows_kernel aug=1): acc1 = -d2/2, so sign(d2) = -sign(acc1) and
//   W_ij = C_ij * (sum_alpha alpha e_alpha) * sign(acc1)           (one FMUL + one LOP3).
// P = bf16(W); acc2 = P * [S_j | 1 ...]: columns < dim give sum_j W_ij s_j and column `dim`
// (the ones column of the B operand) gives sum_j W_ij from the SAME rounded W.
// ---------------------------------------------------------------------------------------------
struct MmdBwdParams {
  const __nv_bfloat16* SA;  // [M, ldp] augmented row operand (features in columns < dim)
  const float* grad_out;    // device scalar
  float* grad_x;
  float* grad_y;
  int ldgx, ldgy, dim, ldp;
  int M, n1;
  float c00, c11, c01;
  int n_alpha;
  float a_scale[kMaxAlphas];  // -2 alpha log2 e
  float a_shift[kMaxAlphas];  // -alpha eps log2 e
  float a_pos[kMaxAlphas];    // alpha
  // FP16 weights: W is scaled by the power of two that puts its LARGEST possible entry, wmax_coef * kmax with
  // wmax_coef = max|c| max alpha and kmax = the forward pass's maximum off-diagonal kernel value, just below 2^10;
  // hi + lo then resolve weights down to 2^-34 of the largest one, whatever the absolute scale of the kernel values
  const unsigned* kmax_bits;  // device scalar written by the forward pass (EpiMMD), float bits
  float wmax_coef;
  int atomic;                 // 0: each gradient row is produced by exactly one work item -> plain stores
  // fp32 mode (3-term split operands, pack3_rows_kernel): acc2 holds W.s1 in columns [0, kp) and W.s2 in
  // [kp, 2 kp), the row sum in column 2 kp; s_i itself is read in fp32 from S32.
  int split, kp, ld32;
  const float* S32;
};

template <bool F16_, bool HILO_ = true>
struct PolMmdBwdT {
  using Params = MmdBwdParams;
  static constexpr bool kF16 = F16_;      // FP16 operands + FP16 (hi + lo) weights, scaled into range by w_scale
  // P = W as a hi + lo pair (the 2^-9 rounding of a single bf16 would dominate the gradient error of small samples).
  // A single FP16 weight (2^-12, independent over the columns a row sums: ~2^-12 / sqrt(M)) is enough once the stacked
  // sample has thousands of rows: the launcher drops the lo half there and GEMM-2 + the conversion work halve.
  static constexpr bool kHiLo = HILO_;
  const Params& p;
  bool fast;
  float cf, wscale, inv_wscale;
  CAREL_DEVINL PolMmdBwdT(const Params& p_, const ChainCtx&) : p(p_), fast(false), cf(0.f), wscale(1.f), inv_wscale(1.f) {
    if (kF16) {
      const float wmax = p.wmax_coef * __uint_as_float(__ldg(p.kmax_bits));
      if (wmax > 0.f) {
        int e;
        frexpf(wmax, &e);                       // wmax = m 2^e, m in [0.5, 1)
        e = max(-100, min(e, 100));
        wscale = ldexpf(1.f, 10 - e);
        inv_wscale = ldexpf(1.f, e - 10);
      }
    }
  }
  CAREL_DEVINL void begin_item(const ChainCtx&) {}
  CAREL_DEVINL void prefetch(const ChainCtx&, int) {}
  CAREL_DEVINL float ksum(float a) const {   // sum_alpha alpha * exp(-alpha (|d2| + eps))
    float k = 0.f;
    for (int t = 0; t < p.n_alpha; ++t)
      k = fmaf(p.a_pos[t], ex2_approx(fmaf(a, p.a_scale[t], p.a_shift[t])), k);
    return k;
  }
  CAREL_DEVINL float coef(int row, int col) const {
    const bool rx = row < p.n1, cx = col < p.n1;
    return rx ? (cx ? p.c00 : p.c01) : (cx ? p.c01 : p.c11);
  }
  static CAREL_DEVINL float with_sign_of(float mag, uint32_t acc_bits) {
    return __uint_as_float(__float_as_uint(mag) ^ (acc_bits & 0x80000000u));
  }
  // columns [j0, j0 + R) of this thread: uniform coefficient, no diagonal, fully in range?
  CAREL_DEVINL void begin_tile(const ChainCtx& ctx, int j0, int R) {
    const int r0 = ctx.owner * 128, r1 = r0 + 127, j1 = j0 + R - 1;
    fast = r1 < p.M && j1 < p.M && ((r1 < p.n1) || (r0 >= p.n1)) && ((j1 < p.n1) || (j0 >= p.n1)) &&
           (j0 > r1 || j1 < r0);
    cf = coef(r0, j0) * wscale;
  }
  CAREL_DEVINL void weights16(const ChainCtx& ctx, int jc, const uint32_t (&r)[16], const uint32_t (&)[16],
                              float (&w)[16]) {
    const bool single = p.n_alpha == 1;
    const float sc = p.a_scale[0], sh = p.a_shift[0], a0 = p.a_pos[0];
    if (fast) {
      const float ca = cf * a0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = fabsf(__uint_as_float(r[i]));
        const float k = single ? ca * ex2_approx(fmaf(a, sc, sh)) : cf * ksum(a);
        w[i] = with_sign_of(k, r[i]);
      }
    } else {
      // boundary / diagonal tiles: same arithmetic, branch-free per-element coefficient
      const int row = ctx.row;
      const bool row_ok = row < p.M, rx = row < p.n1;
      const float c_same = (rx ? p.c00 : p.c01) * wscale, c_other = (rx ? p.c01 : p.c11) * wscale;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = jc + i;
        const float a = fabsf(__uint_as_float(r[i]));
        const float k = single ? a0 * ex2_approx(fmaf(a, sc, sh)) : ksum(a);
        const bool ok = row_ok && col < p.M && col != row;
        const float c = ok ? ((col < p.n1) ? c_same : c_other) : 0.f;
        w[i] = with_sign_of(c * k, r[i]);
      }
    }
  }
  CAREL_DEVINL void store16(const ChainCtx&, int, const uint32_t (&)[8]) {}
  template <class Release>
  CAREL_DEVINL void final(const ChainCtx& ctx, bool valid, Release&& release) {
    // row sum of the rounded W: the ones column of the B operand
    const uint32_t rtot_bits = tmem_ld1(ctx.acc2_taddr + (p.split ? 2 * p.kp : p.dim));
    tmem_ld_wait();
    const float rtot = __uint_as_float(rtot_bits);
    const int nfeat16 = p.split ? (p.kp >> 4) : ctx.n2_16;
    const int q0 = (nfeat16 * ctx.half) / kChainSplit, q1 = (nfeat16 * (ctx.half + 1)) / kChainSplit;
    const float g4 = 4.f * __ldg(p.grad_out) * inv_wscale;
    const bool row_ok = valid && ctx.row < p.M;
    const int rowc = min(ctx.row, p.M - 1);
    const __nv_bfloat16* srow = p.SA + size_t(rowc) * p.ldp;
    float* dst = (rowc < p.n1) ? p.grad_x + size_t(rowc) * p.ldgx
                               : p.grad_y + size_t(rowc - p.n1) * p.ldgy;
#pragma unroll 1
    for (int q = q0; q < q1; ++q) {
      uint32_t r[16], r2[16];
      float sv[16];
      tmem_ld16(ctx.acc2_taddr + q * 16, r);
      if (p.split) {
        tmem_ld16(ctx.acc2_taddr + p.kp + q * 16, r2);
        const float* s32 = p.S32 + size_t(rowc) * p.ld32 + q * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) sv[i] = (q * 16 + i < p.dim) ? __ldg(s32 + i) : 0.f;
      } else {
        const uint4 sv0 = __ldg(reinterpret_cast<const uint4*>(srow + q * 16));
        const uint4 sv1 = __ldg(reinterpret_cast<const uint4*>(srow + q * 16) + 1);
        const uint32_t u[8] = {sv0.x, sv0.y, sv0.z, sv0.w, sv1.x, sv1.y, sv1.z, sv1.w};
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float2 s2 = kF16 ? unpack_f16x2(u[h]) : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[h]));
          sv[2 * h] = s2.x;
          sv[2 * h + 1] = s2.y;
        }
      }
      tmem_ld_wait();
      if (!row_ok) continue;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float g = __uint_as_float(r[i]);
        if (p.split) g += __uint_as_float(r2[i]);
        v[i] = g4 * (rtot * sv[i] - g);
      }
      const int f0 = q * 16;
      if (!p.atomic && f0 + 16 <= p.dim && ((reinterpret_cast<uintptr_t>(dst + f0) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<float4*>(dst + f0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else if (!p.atomic) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (f0 + i < p.dim) dst[f0 + i] = v[i];
      } else if (f0 + 16 <= p.dim && ((reinterpret_cast<uintptr_t>(dst + f0) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) red_add_v4(dst + f0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (f0 + i < p.dim) atomicAdd(dst + f0 + i, v[i]);
      }
    }
    release();
  }
  CAREL_DEVINL void finish() {}
};
using PolMmdBwd = PolMmdBwdT<false, true>;

}  // namespace carel
