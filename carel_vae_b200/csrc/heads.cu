// heads.cu — carel_heads_{fwd,bwd}: the three small classifier heads of DrlClassifier in two kernels
// (drl_classifier_ec_mmd_final_mul.py get_emotion_mul_loss :461-476, get_cause_mul_loss :478-492,
//  get_pair_mul_loss :494-513; drl_classifier_ec_hsic.py :459-505 for the sigmoid/BCE emotion head):
//
//   emotion : x = W_e (z_e (.) m_e) + b_e          mode 0: CrossEntropy over C_e classes (int64 labels)
//                                                   mode 1: BCELoss(sigmoid(x), y (1-ls) + ls / n_e)      (C_e = 1)
//   cause   : x = w_c . (z_c (.) m_c) + b_c         BCELoss(sigmoid(x), y (1-ls) + ls / n_c)
//   pair    : x = w_p . (z   (.) m_p) + b_p         BCEWithLogits(x, y (1-ls) + ls / n_p, pos_weight = (N - sum y) / sum y)
//                                                   (+ the reference's "replace an infinite loss by 0" guard, :510-511)
// m_* are the dropout masks already scaled by 1/(1-p) (drawn by torch in the reference's call order, so the RNG stream is
// the reference's), or NULL.  The reference runs ~100 launch-bound ATen kernels for this; here one warp
// handles one clause row: its 2L latent values stay in registers for all 8 dot products, losses are reduced with warp
// shuffles + one shared-memory hop + one double atomic per block, and the backward kernel produces the complete
// d(loss)/dz row (all three heads) plus the weight gradients (per-warp register accumulators -> shared -> atomics).
#include <algorithm>

#include "api_common.cuh"

namespace carel {

constexpr int kHeadsMaxC = 8;       // emotion classes
constexpr int kHeadsMaxK = 8;       // 2L / 32 elements per lane  (L <= 128)

struct HeadsArgs {
  const float* z; int64_t ldz;
  const float* m_e; const float* m_c; const float* m_p;     // [N, L], [N, L], [N, 2L] or NULL
  const float* W_e; const float* b_e;                        // [C_e, L], [C_e]
  const float* W_c; const float* b_c;                        // [L], [1]
  const float* W_p; const float* b_p;                        // [2L], [1]
  const int64_t* emo_idx;                                    // mode 0 labels
  const float* emo_y;                                        // mode 1 labels
  const float* cau_y; const float* pair_y;
  int N, L, C_e, emo_mode, guard_inf;
  float ls, inv_ne, inv_nc, inv_np;                          // label smoothing, 1 / class counts used for smoothing
  float wgt[3];                                              // loss weights of the combined output (emotion, cause, pair)
  float* logits;                                             // [N, C_e + 2]  (emotion | cause | pair), saved for backward
  double* sums;                                              // [5]: emo, cause, pair S0, pair S1, sum pair_y
  // backward
  const float* g;                                            // upstream gradient of the combined loss (device scalar)
  float* dz; int64_t lddz;                                   // [N, 2L]
  float* dW_e; float* db_e; float* dW_c; float* db_c; float* dW_p; float* db_p;   // accumulated (zeroed by the host side)
};

static CAREL_DEVINL float softplus_neg(float x) { return log1pf(__expf(-fabsf(x))) + fmaxf(-x, 0.f); }   // log(1 + e^-x)

CAREL_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
CAREL_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
CAREL_DEVINL float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
CAREL_DEVINL float dot4(float4 a, float4 b, float acc) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc)))); }

// VEC: L % 4 == 0 and every pointer / pitch 16-byte aligned: lane l owns latent dims 4l .. 4l+3 of each half, so a
// row is six 128-bit loads per lane, all in flight at once (the scalar form issues 24 dependent-address 32-bit loads
// and needed 1160 instructions per row; ncu: long-scoreboard bound at 3 blocks per SM).
template <bool VEC>
static __global__ void __launch_bounds__(256, 4) heads_fwd_kernel(const HeadsArgs a) {
  __shared__ double sh[8][5];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int L = a.L, nk = (L + 31) / 32;          // per-lane elements of each half
  double s_e = 0.0, s_c = 0.0, s_p0 = 0.0, s_p1 = 0.0, s_y = 0.0;
  for (int r = blockIdx.x * wpb + wid; r < a.N; r += gridDim.x * wpb) {
    const float* zr = a.z + size_t(r) * a.ldz;
    float acc_e[kHeadsMaxC], acc_c = 0.f, acc_p = 0.f;
#pragma unroll
    for (int c = 0; c < kHeadsMaxC; ++c) acc_e[c] = 0.f;
    if (VEC) {
      const int k = lane * 4;
      if (k < L) {
        const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 ze = ld4(zr + k), zc = ld4(zr + L + k);
        const float4 me = a.m_e ? ld4(a.m_e + size_t(r) * L + k) : one;
        const float4 mc = a.m_c ? ld4(a.m_c + size_t(r) * L + k) : one;
        const float4 mpe = a.m_p ? ld4(a.m_p + size_t(r) * 2 * L + k) : one;
        const float4 mpc = a.m_p ? ld4(a.m_p + size_t(r) * 2 * L + L + k) : one;
        const float4 xe = mul4(ze, me), xc = mul4(zc, mc), pe = mul4(ze, mpe), pc = mul4(zc, mpc);
#pragma unroll
        for (int c = 0; c < kHeadsMaxC; ++c)
          if (c < a.C_e) acc_e[c] = dot4(xe, ldg4(a.W_e + size_t(c) * L + k), 0.f);
        acc_c = dot4(xc, ldg4(a.W_c + k), 0.f);
        acc_p = dot4(pe, ldg4(a.W_p + k), dot4(pc, ldg4(a.W_p + L + k), 0.f));
      }
    } else
#pragma unroll 4
    for (int q = 0; q < nk; ++q) {
      const int k = q * 32 + lane;
      if (k < L) {
        const float ze = zr[k], zc = zr[L + k];
        const float xe = a.m_e ? ze * a.m_e[size_t(r) * L + k] : ze;
        const float xc = a.m_c ? zc * a.m_c[size_t(r) * L + k] : zc;
        const float pe = a.m_p ? ze * a.m_p[size_t(r) * 2 * L + k] : ze;
        const float pc = a.m_p ? zc * a.m_p[size_t(r) * 2 * L + L + k] : zc;
#pragma unroll
        for (int c = 0; c < kHeadsMaxC; ++c)
          if (c < a.C_e) acc_e[c] = fmaf(xe, __ldg(a.W_e + size_t(c) * L + k), acc_e[c]);
        acc_c = fmaf(xc, __ldg(a.W_c + k), acc_c);
        acc_p = fmaf(pe, __ldg(a.W_p + k), fmaf(pc, __ldg(a.W_p + L + k), acc_p));
      }
    }
#pragma unroll
    for (int c = 0; c < kHeadsMaxC; ++c)
      if (c < a.C_e) acc_e[c] = warp_sum(acc_e[c]) + __ldg(a.b_e + c);
    acc_c = warp_sum(acc_c) + __ldg(a.b_c);
    acc_p = warp_sum(acc_p) + __ldg(a.b_p);
    if (lane == 0) {
      float* lg = a.logits + size_t(r) * (a.C_e + 2);
      // ---- emotion ----
      if (a.emo_mode == 0) {
        float mx = acc_e[0];
        for (int c = 1; c < a.C_e; ++c) mx = fmaxf(mx, acc_e[c]);
        float se = 0.f;
        for (int c = 0; c < a.C_e; ++c) se += __expf(acc_e[c] - mx);
        const int lab = int(a.emo_idx[r]);
        float xl = 0.f;
        for (int c = 0; c < a.C_e; ++c) {
          lg[c] = acc_e[c];
          if (c == lab) xl = acc_e[c];
        }
        s_e += double(mx + logf(se) - xl);
      } else {
        for (int c = 0; c < a.C_e; ++c) {
          lg[c] = acc_e[c];
          const float p = 1.f / (1.f + __expf(-acc_e[c]));
          const float t = a.emo_y[r] * (1.f - a.ls) + a.ls * a.inv_ne;
          s_e += double(-(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(log1pf(-p), -100.f)));
        }
      }
      // ---- cause: sigmoid + BCELoss (log clamped at -100 like torch) ----
      {
        lg[a.C_e] = acc_c;
        const float p = 1.f / (1.f + __expf(-acc_c));
        const float t = a.cau_y[r] * (1.f - a.ls) + a.ls * a.inv_nc;
        s_c += double(-(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(log1pf(-p), -100.f)));
      }
      // ---- pair: BCEWithLogits; pos_weight enters linearly: loss_i = [(1-t) x + sp] + (pw - 1) t sp ----
      {
        lg[a.C_e + 1] = acc_p;
        const float y = a.pair_y[r];
        const float t = y * (1.f - a.ls) + a.ls * a.inv_np;
        const float sp = softplus_neg(acc_p);
        s_p0 += double((1.f - t) * acc_p + sp);
        s_p1 += double(t * sp);
        s_y += double(y);
      }
    }
  }
  if (lane == 0) { sh[wid][0] = s_e; sh[wid][1] = s_c; sh[wid][2] = s_p0; sh[wid][3] = s_p1; sh[wid][4] = s_y; }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh[w][threadIdx.x];
    if (t != 0.0) atomicAdd(a.sums + threadIdx.x, t);
  }
}

// losses[0..2] = emotion, cause, pair; losses[3] = w_e emotion + w_c cause + w_p pair
static __global__ void heads_finalize_kernel(const double* __restrict__ sums, int N, int C_e, int emo_mode, int guard_inf,
                                             float w_e, float w_c, float w_p, float* __restrict__ losses) {
  if (threadIdx.x != 0) return;
  losses[0] = float(sums[0] / (emo_mode == 0 ? double(N) : double(N) * C_e));
  losses[1] = float(sums[1] / double(N));
  const float pw = float((double(N) - sums[4]) / sums[4]);            // (N - sum y) / sum y  (inf / nan like the reference)
  float lp = (float(sums[2]) + (pw - 1.f) * float(sums[3])) / float(N);
  if (guard_inf && isinf(lp)) lp = 0.f;
  losses[2] = lp;
  losses[3] = w_e * losses[0] + w_c * losses[1] + w_p * lp;
}

// Weight gradients accumulate in ONE shared-memory tile per block (float atomics, lanes on distinct addresses) instead
// of per-warp register accumulators + a per-warp shared slice: 168 -> <= 64 registers and 46 KB -> 6 KB of shared
// memory, so 4 blocks (32 warps) per SM instead of 1 block (8 warps); ncu had the old form long-scoreboard bound at
// 12 % of the warp slots.  Each block ends with one global atomic per non-zero weight-gradient element.
template <bool VEC>
static __global__ void __launch_bounds__(256, 4) heads_bwd_kernel(const HeadsArgs a) {
  constexpr int kRows = kHeadsMaxC + 3;                      // dW_e rows | dW_c | dW_p (emotion half) | dW_p (cause half)
  constexpr int kW = kHeadsMaxK * 32 / 2;                    // 128 = max L
  __shared__ float sh_w[kRows][kW];
  __shared__ float sh_b[kHeadsMaxC + 2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int L = a.L, nk = (L + 31) / 32;
  for (int i = threadIdx.x; i < kRows * kW; i += blockDim.x) (&sh_w[0][0])[i] = 0.f;
  if (threadIdx.x < kHeadsMaxC + 2) sh_b[threadIdx.x] = 0.f;
  __syncthreads();
  // shared-memory column of latent dim k: in the 128-bit form lane l owns dims 4l..4l+3, which would put the lanes of
  // one atomic instruction 4 words apart (4-way bank conflicts); (k & 3) * 32 + (k >> 2) puts them on 32 banks
  auto slot = [](int k) { return VEC ? ((k & 3) * 32 + (k >> 2)) : k; };
  const float invN = 1.f / float(a.N);
  const float g0 = __ldg(a.g);
  const float ge = g0 * a.wgt[0], gc = g0 * a.wgt[1], gp = g0 * a.wgt[2];
  const float pw = float((double(a.N) - a.sums[4]) / a.sums[4]);
  // the guard zeroes the pair loss when it is infinite: then it has no gradient either
  float lp_chk = (float(a.sums[2]) + (pw - 1.f) * float(a.sums[3])) * invN;
  const bool pair_dead = a.guard_inf && isinf(lp_chk);
  float gb[kHeadsMaxC + 2];
#pragma unroll
  for (int c = 0; c < kHeadsMaxC + 2; ++c) gb[c] = 0.f;

  for (int r = blockIdx.x * wpb + wid; r < a.N; r += gridDim.x * wpb) {
    const float* lg = a.logits + size_t(r) * (a.C_e + 2);
    float dl_e[kHeadsMaxC];
    if (a.emo_mode == 0) {
      float mx = lg[0];
      for (int c = 1; c < a.C_e; ++c) mx = fmaxf(mx, lg[c]);
      float se = 0.f;
      for (int c = 0; c < a.C_e; ++c) se += __expf(lg[c] - mx);
      const int lab = int(a.emo_idx[r]);
#pragma unroll
      for (int c = 0; c < kHeadsMaxC; ++c)
        dl_e[c] = (c < a.C_e) ? (__expf(lg[c] - mx) / se - (c == lab ? 1.f : 0.f)) * ge * invN : 0.f;
    } else {
      const float t = a.emo_y[r] * (1.f - a.ls) + a.ls * a.inv_ne;
#pragma unroll
      for (int c = 0; c < kHeadsMaxC; ++c) {
        if (c < a.C_e) {
          const float p = 1.f / (1.f + __expf(-lg[c]));
          dl_e[c] = (p - t) / fmaxf(p * (1.f - p), 1e-12f) * (p * (1.f - p)) * ge * invN / float(a.C_e);
        } else {
          dl_e[c] = 0.f;
        }
      }
    }
    float dl_c, dl_p;
    {
      const float p = 1.f / (1.f + __expf(-lg[a.C_e]));
      const float t = a.cau_y[r] * (1.f - a.ls) + a.ls * a.inv_nc;
      dl_c = (p - t) / fmaxf(p * (1.f - p), 1e-12f) * (p * (1.f - p)) * gc * invN;
    }
    {
      const float x = lg[a.C_e + 1];
      const float t = a.pair_y[r] * (1.f - a.ls) + a.ls * a.inv_np;
      const float sig_neg = 1.f / (1.f + __expf(x));                       // sigma(-x) = -d softplus(-x)/dx
      dl_p = pair_dead ? 0.f : ((1.f - t) - (1.f + (pw - 1.f) * t) * sig_neg) * gp * invN;
    }
#pragma unroll
    for (int c = 0; c < kHeadsMaxC; ++c) gb[c] += dl_e[c];
    gb[kHeadsMaxC] += dl_c;
    gb[kHeadsMaxC + 1] += dl_p;
    const float* zr = a.z + size_t(r) * a.ldz;
    float* dzr = a.dz + size_t(r) * a.lddz;
    // one latent dim k of both halves: its dz pair and its eleven weight-gradient contributions
    auto one = [&](int k, float ze, float zc, float me, float mc, float mpe, float mpc, float& dze, float& dzc) {
      const int sl = slot(k);
      float de = 0.f;
#pragma unroll
      for (int c = 0; c < kHeadsMaxC; ++c) {
        if (c < a.C_e) {
          de = fmaf(dl_e[c], __ldg(a.W_e + size_t(c) * L + k), de);
          atomicAdd(&sh_w[c][sl], dl_e[c] * (ze * me));
        }
      }
      atomicAdd(&sh_w[kHeadsMaxC][sl], dl_c * (zc * mc));
      atomicAdd(&sh_w[kHeadsMaxC + 1][sl], dl_p * (ze * mpe));
      atomicAdd(&sh_w[kHeadsMaxC + 2][sl], dl_p * (zc * mpc));
      dze = de * me + dl_p * __ldg(a.W_p + k) * mpe;
      dzc = dl_c * __ldg(a.W_c + k) * mc + dl_p * __ldg(a.W_p + L + k) * mpc;
    };
    if (VEC) {
      const int k = lane * 4;
      if (k < L) {
        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 ze = ld4(zr + k), zc = ld4(zr + L + k);
        const float4 me = a.m_e ? ld4(a.m_e + size_t(r) * L + k) : one4;
        const float4 mc = a.m_c ? ld4(a.m_c + size_t(r) * L + k) : one4;
        const float4 mpe = a.m_p ? ld4(a.m_p + size_t(r) * 2 * L + k) : one4;
        const float4 mpc = a.m_p ? ld4(a.m_p + size_t(r) * 2 * L + L + k) : one4;
        float4 de, dc;
        one(k, ze.x, zc.x, me.x, mc.x, mpe.x, mpc.x, de.x, dc.x);
        one(k + 1, ze.y, zc.y, me.y, mc.y, mpe.y, mpc.y, de.y, dc.y);
        one(k + 2, ze.z, zc.z, me.z, mc.z, mpe.z, mpc.z, de.z, dc.z);
        one(k + 3, ze.w, zc.w, me.w, mc.w, mpe.w, mpc.w, de.w, dc.w);
        *reinterpret_cast<float4*>(dzr + k) = de;
        *reinterpret_cast<float4*>(dzr + L + k) = dc;
      }
    } else {
#pragma unroll
      for (int q = 0; q < kHeadsMaxK / 2; ++q) {
        const int k = q * 32 + lane;
        if (q < nk && k < L) {
          const float me = a.m_e ? a.m_e[size_t(r) * L + k] : 1.f;
          const float mc = a.m_c ? a.m_c[size_t(r) * L + k] : 1.f;
          const float mpe = a.m_p ? a.m_p[size_t(r) * 2 * L + k] : 1.f;
          const float mpc = a.m_p ? a.m_p[size_t(r) * 2 * L + L + k] : 1.f;
          one(k, zr[k], zr[L + k], me, mc, mpe, mpc, dzr[k], dzr[L + k]);
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kHeadsMaxC + 2; ++c)
      if (gb[c] != 0.f) atomicAdd(&sh_b[c], gb[c]);
  }
  __syncthreads();
  // ---- block tile -> one global atomic per non-zero element ----
  for (int idx = threadIdx.x; idx < kRows * L; idx += blockDim.x) {
    const int row = idx / L, k = idx - row * L;
    const float t = sh_w[row][slot(k)];
    if (t == 0.f) continue;
    if (row < kHeadsMaxC) {
      if (row < a.C_e && a.dW_e) atomicAdd(a.dW_e + size_t(row) * L + k, t);
    } else if (row == kHeadsMaxC) {
      if (a.dW_c) atomicAdd(a.dW_c + k, t);
    } else if (a.dW_p) {
      atomicAdd(a.dW_p + (row - kHeadsMaxC - 1) * L + k, t);
    }
  }
  if (threadIdx.x < kHeadsMaxC + 2) {
    const float t = sh_b[threadIdx.x];
    const int c = threadIdx.x;
    if (c < kHeadsMaxC) {
      if (c < a.C_e && a.db_e && t != 0.f) atomicAdd(a.db_e + c, t);
    } else if (c == kHeadsMaxC) {
      if (a.db_c && t != 0.f) atomicAdd(a.db_c, t);
    } else if (a.db_p && t != 0.f) {
      atomicAdd(a.db_p, t);
    }
  }
}

}  // namespace carel

using namespace carel;

// 128-bit path: L a multiple of 4 and every row pointer 16-byte aligned
static bool heads_vec_ok(const HeadsArgs& a, const float* dz, int64_t lddz) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (a.L % 4 == 0) && (a.ldz % 4 == 0) && al(a.z) && al(a.m_e) && al(a.m_c) && al(a.m_p) && al(a.W_e) &&
         al(a.W_c) && al(a.W_p) && (dz == nullptr || (al(dz) && lddz % 4 == 0));
}

extern "C" size_t carel_heads_workspace_bytes(int64_t N, int64_t C_e) {
  return align_up(size_t(N) * size_t(C_e + 2) * 4, 256) + 256;
}

static int heads_fill(HeadsArgs& a, const float* z, int64_t ldz, int64_t N, int64_t L, int64_t C_e, int emo_mode,
                      const float* const* masks3, const float* const* weights6, const void* emo_labels,
                      const float* cau_labels, const float* pair_labels, float label_smoothing, float n_e, float n_c,
                      float n_p, int guard_inf, const float* loss_weights3_host, void* ws, size_t ws_bytes) {
  CAREL_REQUIRE(N > 0 && L > 0 && L <= 128, "carel_heads: ec_dim must be in 1..128, got %lld", (long long)L);
  CAREL_REQUIRE(C_e >= 1 && C_e <= kHeadsMaxC, "carel_heads: 1..%d emotion outputs supported", kHeadsMaxC);
  CAREL_REQUIRE(emo_mode == 0 || emo_mode == 1, "carel_heads: emo_mode must be 0 (cross-entropy) or 1 (sigmoid + BCE)");
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_heads_workspace_bytes(N, C_e), "carel_heads: workspace too small");
  a.z = z; a.ldz = ldz;
  a.m_e = masks3 ? masks3[0] : nullptr; a.m_c = masks3 ? masks3[1] : nullptr; a.m_p = masks3 ? masks3[2] : nullptr;
  a.W_e = weights6[0]; a.b_e = weights6[1]; a.W_c = weights6[2]; a.b_c = weights6[3]; a.W_p = weights6[4]; a.b_p = weights6[5];
  a.emo_idx = emo_mode == 0 ? static_cast<const int64_t*>(emo_labels) : nullptr;
  a.emo_y = emo_mode == 1 ? static_cast<const float*>(emo_labels) : nullptr;
  a.cau_y = cau_labels; a.pair_y = pair_labels;
  a.N = int(N); a.L = int(L); a.C_e = int(C_e); a.emo_mode = emo_mode; a.guard_inf = guard_inf;
  a.ls = label_smoothing; a.inv_ne = 1.f / n_e; a.inv_nc = 1.f / n_c; a.inv_np = 1.f / n_p;
  for (int i = 0; i < 3; ++i) a.wgt[i] = loss_weights3_host ? loss_weights3_host[i] : 1.f;
  a.logits = static_cast<float*>(ws);
  a.sums = reinterpret_cast<double*>(static_cast<uint8_t*>(ws) + align_up(size_t(N) * size_t(C_e + 2) * 4, 256));
  return 0;
}

extern "C" int carel_heads_fwd(const float* z, int64_t ldz, int64_t N, int64_t L, int64_t C_e, int emo_mode,
                               const float* const* masks3, const float* const* weights6, const void* emo_labels,
                               const float* cau_labels, const float* pair_labels, float label_smoothing,
                               float n_e, float n_c, float n_p, int guard_inf, const float* loss_weights3_host,
                               float* losses4, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HeadsArgs a{};
  if (int rc = heads_fill(a, z, ldz, N, L, C_e, emo_mode, masks3, weights6, emo_labels, cau_labels, pair_labels,
                          label_smoothing, n_e, n_c, n_p, guard_inf, loss_weights3_host, ws, ws_bytes)) return rc;
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.sums, 0, 5 * sizeof(double), st));
  if (heads_vec_ok(a, nullptr, 0))
    heads_fwd_kernel<true><<<int(std::min<int64_t>((N + 7) / 8, 148 * 8)), 256, 0, st>>>(a);
  else
    heads_fwd_kernel<false><<<int(std::min<int64_t>((N + 7) / 8, 148 * 8)), 256, 0, st>>>(a);
  CAREL_LAUNCHED();
  heads_finalize_kernel<<<1, 32, 0, st>>>(a.sums, int(N), int(C_e), emo_mode, guard_inf, a.wgt[0], a.wgt[1], a.wgt[2], losses4);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_heads_bwd(const float* grad_out, const float* z, int64_t ldz, int64_t N, int64_t L, int64_t C_e,
                               int emo_mode, const float* const* masks3, const float* const* weights6,
                               const void* emo_labels, const float* cau_labels, const float* pair_labels,
                               float label_smoothing, float n_e, float n_c, float n_p, int guard_inf,
                               const float* loss_weights3_host, float* grad_z, int64_t ldgz, float* const* grad_weights6, void* ws, size_t ws_bytes,
                               void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HeadsArgs a{};
  if (int rc = heads_fill(a, z, ldz, N, L, C_e, emo_mode, masks3, weights6, emo_labels, cau_labels, pair_labels,
                          label_smoothing, n_e, n_c, n_p, guard_inf, loss_weights3_host, ws, ws_bytes)) return rc;
  CAREL_REQUIRE(grad_z != nullptr, "carel_heads_bwd: grad_z is required");
  a.g = grad_out;
  a.dz = grad_z; a.lddz = ldgz;
  a.dW_e = grad_weights6[0]; a.db_e = grad_weights6[1]; a.dW_c = grad_weights6[2]; a.db_c = grad_weights6[3];
  a.dW_p = grad_weights6[4]; a.db_p = grad_weights6[5];
  // zero the accumulation targets: ONE memset when the caller carved them back to back out of one buffer (the Python
  // wrapper does), else one each
  {
    float* ptr[6] = {a.dW_e, a.db_e, a.dW_c, a.db_c, a.dW_p, a.db_p};
    const size_t cnt[6] = {size_t(C_e) * L, size_t(C_e), size_t(L), 1, size_t(2 * L), 1};
    float* first = nullptr;
    float* end = nullptr;
    bool packed = true;
    size_t total = 0;
    for (int i = 0; i < 6; ++i) {
      if (!ptr[i]) continue;
      if (first == nullptr) first = ptr[i];
      else if (ptr[i] != end) packed = false;
      end = ptr[i] + cnt[i];
      total += cnt[i];
    }
    if (first != nullptr && packed) {
      CAREL_CHECK_CUDA(cudaMemsetAsync(first, 0, total * 4, st));
    } else {
      for (int i = 0; i < 6; ++i)
        if (ptr[i]) CAREL_CHECK_CUDA(cudaMemsetAsync(ptr[i], 0, cnt[i] * 4, st));
    }
  }
  const int blocks = int(std::min<int64_t>((N + 7) / 8, 148 * 4));      // 4 blocks per SM: 592 atomics per weight element
  if (heads_vec_ok(a, grad_z, ldgz))
    heads_bwd_kernel<true><<<blocks, 256, 0, st>>>(a);
  else
    heads_bwd_kernel<false><<<blocks, 256, 0, st>>>(a);
  CAREL_LAUNCHED();
  return 0;
}

namespace carel {
// ---------------------------------------------------------------------------------------------
// Loss assembly: out = sum_i w_i t_i over device scalars (the reference's chain of python-level scalar ops,
// drl_classifier_ec_mmd_final_mul.py:256-261).  Products and sums are rounded separately and taken left to right,
// so the value is bit-identical to the torch chain it replaces; one launch instead of ~10 forward and ~10 backward.
// ---------------------------------------------------------------------------------------------
constexpr int kWsumMax = 16;
struct WsumArgs {
  const float* t[kWsumMax];
  float w[kWsumMax];
  int n;
};
static __global__ void wsum_fwd_kernel(const WsumArgs a, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  float acc = __fmul_rn(a.w[0], *a.t[0]);
  for (int i = 1; i < a.n; ++i) acc = __fadd_rn(acc, __fmul_rn(a.w[i], *a.t[i]));
  *out = acc;
}
static __global__ void wsum_bwd_kernel(const WsumArgs a, const float* __restrict__ g, float* __restrict__ grads) {
  const int i = threadIdx.x;
  if (i < a.n) grads[i] = __fmul_rn(a.w[i], *g);
}
}  // namespace carel

extern "C" int carel_wsum_fwd(const float* const* terms_host, const float* weights_host, int n, float* out,
                              void* stream) {
  CAREL_REQUIRE(n >= 1 && n <= kWsumMax, "carel_wsum_fwd: 1..%d terms supported, got %d", kWsumMax, n);
  CAREL_REQUIRE(terms_host != nullptr && weights_host != nullptr && out != nullptr, "carel_wsum_fwd: null argument");
  WsumArgs a{};
  a.n = n;
  for (int i = 0; i < n; ++i) {
    CAREL_REQUIRE(terms_host[i] != nullptr, "carel_wsum_fwd: term %d is null", i);
    a.t[i] = terms_host[i];
    a.w[i] = weights_host[i];
  }
  wsum_fwd_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, out);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_wsum_bwd(const float* grad_out, const float* weights_host, int n, float* grads, void* stream) {
  CAREL_REQUIRE(n >= 1 && n <= kWsumMax, "carel_wsum_bwd: 1..%d terms supported, got %d", kWsumMax, n);
  CAREL_REQUIRE(grad_out != nullptr && weights_host != nullptr && grads != nullptr, "carel_wsum_bwd: null argument");
  WsumArgs a{};
  a.n = n;
  for (int i = 0; i < n; ++i) a.w[i] = weights_host[i];
  wsum_bwd_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, grad_out, grads);
  CAREL_LAUNCHED();
  return 0;
}
