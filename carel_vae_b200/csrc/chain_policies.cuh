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
// P = bf16(W) tile; acc2 = sum_j W_ij s_j; the row sum uses the same rounded W so the two terms
// stay consistent.
// ---------------------------------------------------------------------------------------------
struct MmdBwdParams {
  const float* norms;
  const __nv_bfloat16* S;   // [M, Kp]
  const float* grad_out;    // device scalar
  float* grad_x;
  float* grad_y;
  int ldgx, ldgy, dim, Kp;
  int M, n1;
  float c00, c11, c01;
  int n_alpha;
  float a_scale[kMaxAlphas];  // -alpha log2 e
  float a_shift[kMaxAlphas];  // -alpha eps log2 e
  float a_neg[kMaxAlphas];    // -alpha
};

template <int KP>
struct PolMmdBwd {
  using Params = MmdBwdParams;
  const Params& p;
  float rsum, rtot, ni;
  CAREL_DEVINL PolMmdBwd(const Params& p_, const ChainCtx&) : p(p_), rsum(0.f), rtot(0.f), ni(0.f) {}
  CAREL_DEVINL void begin_item(const ChainCtx& ctx) {
    rsum = 0.f;
    ni = __ldg(p.norms + min(ctx.row, p.M - 1));
  }
  CAREL_DEVINL float kprime(float d2) const {
    const float a = fabsf(d2);
    float k = 0.f;
    if (p.n_alpha == 1) {
      k = p.a_neg[0] * ex2_approx(fmaf(a, p.a_scale[0], p.a_shift[0]));
    } else {
      for (int t = 0; t < p.n_alpha; ++t)
        k = fmaf(p.a_neg[t], ex2_approx(fmaf(a, p.a_scale[t], p.a_shift[t])), k);
    }
    return d2 > 0.f ? k : (d2 < 0.f ? -k : 0.f);
  }
  CAREL_DEVINL float coef(int row, int col) const {
    const bool rx = row < p.n1, cx = col < p.n1;
    return rx ? (cx ? p.c00 : p.c01) : (cx ? p.c01 : p.c11);
  }
  template <int NG, int BN1>
  CAREL_DEVINL void weights(const ChainCtx& ctx, int t, int tcol, uint32_t taddr,
                            uint32_t (&packed)[16]) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    const int r0 = ctx.owner * 128, r1 = r0 + 127;
    const int j0 = t * BN1 + tcol, j1 = j0 + 31;
    const bool fast = r1 < p.M && j1 < p.M && ((r1 < p.n1) || (r0 >= p.n1)) &&
                      ((j1 < p.n1) || (j0 >= p.n1)) && (j0 > r1 || j1 < r0);
    float w[32];
    if (fast) {
      const float cf = coef(r0, j0);
      const float4* nj4 = reinterpret_cast<const float4*>(p.norms + j0);
      float nj[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(nj4 + i);
        nj[4 * i] = v.x; nj[4 * i + 1] = v.y; nj[4 * i + 2] = v.z; nj[4 * i + 3] = v.w;
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        w[i] = cf * kprime(fmaf(-2.f, __uint_as_float(r[i]), ni + nj[i]));
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = j0 + i;
        const bool ok = ctx.row < p.M && col < p.M && col != ctx.row;
        w[i] = ok ? coef(ctx.row, col) *
                        kprime(fmaf(-2.f, __uint_as_float(r[i]), ni + __ldg(p.norms + col)))
                  : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(w[2 * i], w[2 * i + 1]);
      const float2 f = __bfloat1622float2(b);
      rsum += f.x + f.y;
      packed[i] = *reinterpret_cast<const uint32_t*>(&b);
    }
  }
  CAREL_DEVINL void pre_final(const ChainCtx& ctx) {
    __shared__ float rs[2][128];
    const int rl = ctx.lane_q * 32 + (threadIdx.x & 31);
    rs[ctx.half][rl] = rsum;
    named_bar_sync(1, 256);
    rtot = rs[0][rl] + rs[1][rl];
    named_bar_sync(1, 256);
  }
  CAREL_DEVINL void final(const ChainCtx& ctx, int fcol, const uint32_t (&r)[32], bool valid) {
    if (!valid || ctx.row >= p.M || fcol >= p.dim) return;
    const float g4 = 4.f * __ldg(p.grad_out);
    const uint4* s4 = reinterpret_cast<const uint4*>(p.S + size_t(ctx.row) * p.Kp + fcol);
    float* dst = (ctx.row < p.n1) ? p.grad_x + size_t(ctx.row) * p.ldgx
                                  : p.grad_y + size_t(ctx.row - p.n1) * p.ldgy;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 sv = __ldg(s4 + q);
      const uint32_t u[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 s2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[h]));
        const int f = fcol + q * 8 + h * 2;
        if (f < p.dim) atomicAdd(dst + f, g4 * (rtot * s2.x - __uint_as_float(r[q * 8 + h * 2])));
        if (f + 1 < p.dim)
          atomicAdd(dst + f + 1, g4 * (rtot * s2.y - __uint_as_float(r[q * 8 + h * 2 + 1])));
      }
    }
  }
  CAREL_DEVINL void finish() {}
};

}  // namespace carel
