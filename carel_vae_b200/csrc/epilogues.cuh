// epilogues.cuh — register-level epilogues plugged into gemm_tc_kernel.
//
// Each epilogue thread owns ONE accumulator row (TMEM lane) and walks its warp's column half
// in chunks of 32 fp32 columns (tcgen05.ld 32x32b.x32).  Nothing N x N ever goes to HBM:
// the epilogues reduce the tile to scalars / row vectors in registers.
#pragma once
#include "gemm_tc.cuh"

namespace carel {

constexpr int kMaxAlphas = 8;

// ---------------------------------------------------------------------------------------------
// Plain store: C[row, col] (+)= acc (+ bias[col]);  fp32 output, optional atomic accumulate
// (split-K / partial products), optional bf16 copy.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct EpiStore {
  struct Params {
    float* out;            // [M, ld]
    const float* bias;     // [N] or nullptr
    __nv_bfloat16* out_bf16;  // optional [M, ld_bf16]
    int M, N, ld, ld_bf16;
    int atomic;            // 1: atomicAdd into out (bias added only by the kb0 == 0 split)
    float scale;
  };
  const Params& p;
  CAREL_DEVINL EpiStore(const Params& p_, const TileCtx&) : p(p_) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int cbase = ctx.col0 + ctx.half * HALF;
    const bool row_ok = ctx.row < p.M && ctx.kb1 > ctx.kb0;   // an empty split-K item holds no valid accumulator
#pragma unroll 1
    for (int c = 0; c < HALF / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(ctx.taddr + ctx.half * HALF + c * 32, r);
      tmem_ld_wait();
      const int col = cbase + c * 32;
      if (!row_ok || col >= p.N) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r[i]) * p.scale;
        if (p.bias != nullptr && ctx.kb0 == 0 && col + i < p.N) v[i] += __ldg(p.bias + col + i);
      }
      float* o = p.out + size_t(ctx.row) * p.ld + col;
      if (p.atomic) {
        if (col + 32 <= p.N && (p.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) red_add_v4(o + i, v[i], v[i + 1], v[i + 2], v[i + 3]);   // 16-byte reductions
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col + i < p.N) atomicAdd(o + i, v[i]);
        }
      } else if (col + 32 <= p.N && (p.ld & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col + i < p.N) o[i] = v[i];
      }
      if (p.out_bf16 != nullptr) {
        __nv_bfloat16* ob = p.out_bf16 + size_t(ctx.row) * p.ld_bf16 + col;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col + i < p.N) ob[i] = __float2bfloat16_rn(v[i]);
      }
    }
  }
  CAREL_DEVINL void finish() {}
};

// ---------------------------------------------------------------------------------------------
// Pairwise-distance store on AUGMENTED operands (pack_rows_kernel, aug = 1): acc = -d2/2.
//   mode 0: sum_alpha exp(-alpha (|d2| + eps))       (the MMD kernel matrix, ret_matrix=True)
//   mode 1: sqrt(eps + |d2|)                         (reference pdist, norm = 2)
//   mode 2: d2 = -2 a.b + n_i + n_j                  (reference pairwise_distances)
//   mode 3: exp(-d2 / sigma)  with alpha[0] = 1/sigma (reference GaussianKernelMatrix)
// ---------------------------------------------------------------------------------------------
template <int BN>
struct EpiPairwise {
  struct Params {
    float* out;
    int M, N, ld, mode, n_alpha;
    float eps;
    float alpha[kMaxAlphas];
  };
  const Params& p;
  CAREL_DEVINL EpiPairwise(const Params& p_, const TileCtx&) : p(p_) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int cbase = ctx.col0 + ctx.half * HALF;
    const bool row_ok = ctx.row < p.M;
#pragma unroll 1
    for (int c = 0; c < HALF / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(ctx.taddr + ctx.half * HALF + c * 32, r);
      tmem_ld_wait();
      const int col = cbase + c * 32;
      if (!row_ok || col >= p.N) continue;
      float* o = p.out + size_t(ctx.row) * p.ld + col;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (col + i >= p.N) continue;
        const float d2 = -2.f * __uint_as_float(r[i]);
        float v;
        if (p.mode == 0) {
          v = 0.f;
          for (int a = 0; a < p.n_alpha; ++a) v += __expf(-p.alpha[a] * (fabsf(d2) + p.eps));
        } else if (p.mode == 1) {
          v = sqrtf(p.eps + fabsf(d2));
        } else if (p.mode == 2) {
          v = d2;
        } else {
          v = __expf(-d2 * p.alpha[0]);
        }
        o[i] = v;
      }
    }
  }
  CAREL_DEVINL void finish() {}
};

// ---------------------------------------------------------------------------------------------
// MMD forward (reference MMDStatistic.__call__, drl_classifier_ec_mmd_final_mul.py:547-569).
// Operands are the augmented rows of S = cat(x, y), so the accumulator already holds -d2/2:
//     k = sum_alpha exp(-alpha (|d2| + eps)) = sum_alpha 2^(|acc| * a_scale + a_shift)
// Three block sums are accumulated:  acc[0] = sum_{xx, i!=j} k, acc[1] = sum_{yy, i!=j} k,
// acc[2] = sum_{xy} k + sum_{yx} k.   tri != 0: only tiles on/above the diagonal are scheduled and
// every counted pair (j > i) is weighted 2 (K is symmetric).
// Hot loop per element: FFMA(|acc|) + MUFU.EX2 + FADD.
// ---------------------------------------------------------------------------------------------
struct EpiMMDParams {
  double* accum;       // [3]
  unsigned* kmax_bits; // optional: running maximum (float bits) of the off-diagonal kernel values sum_alpha e_alpha;
                       // the FP16 backward scales its weight tile by it (chain_policies.cuh)
  int M, n1, tri, n_alpha;
  float a_scale[kMaxAlphas];  // -2 alpha log2(e)
  float a_shift[kMaxAlphas];  // -alpha eps log2(e)
};

template <int BN>
struct EpiMMD {
  using Params = EpiMMDParams;
  const Params& p;
  double acc0, acc1, acc2;
  float kmx;
  CAREL_DEVINL EpiMMD(const Params& p_, const TileCtx&) : p(p_), acc0(0.0), acc1(0.0), acc2(0.0), kmx(0.f) {}
  CAREL_DEVINL void prefetch(const TileCtx&) {}

  CAREL_DEVINL float kmulti(float a) const {
    float k = 0.f;
    for (int t = 0; t < p.n_alpha; ++t) k += ex2_approx(fmaf(a, p.a_scale[t], p.a_shift[t]));
    return k;
  }
  CAREL_DEVINL void add(int type, float v) {
    if (type == 0) acc0 += double(v);
    else if (type == 1) acc1 += double(v);
    else acc2 += double(v);
  }

  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int r0 = ctx.tm * 128, r1 = r0 + 127;
    const int c0 = ctx.col0 + ctx.half * HALF, c1 = c0 + HALF - 1;
    const bool skip = (r0 >= p.M) || (c0 >= p.M) || (p.tri && c1 <= r0);
    if (skip) return;  // warp-uniform: nothing to read, the MMA result is discarded
    const bool rows_one_block = (r1 < p.n1) || (r0 >= p.n1);
    const bool cols_one_block = (c1 < p.n1) || (c0 >= p.n1);
    const bool off_diag = p.tri ? (c0 > r1) : (c0 > r1 || c1 < r0);
    const bool fast = r1 < p.M && c1 < p.M && rows_one_block && cols_one_block && off_diag;
    const float w = p.tri ? 2.f : 1.f;
    const uint32_t taddr = ctx.taddr + ctx.half * HALF;
    if (fast) {
      float s = 0.f;
      if (p.n_alpha == 1) {
        const float sc = p.a_scale[0], sh = p.a_shift[0];
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four independent accumulation chains
        for_each_chunk32(taddr, HALF / 32, [&](int, const uint32_t (&r)[32]) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float e0 = ex2_approx(fmaf(fabsf(__uint_as_float(r[i])), sc, sh));
            const float e1 = ex2_approx(fmaf(fabsf(__uint_as_float(r[i + 1])), sc, sh));
            const float e2 = ex2_approx(fmaf(fabsf(__uint_as_float(r[i + 2])), sc, sh));
            const float e3 = ex2_approx(fmaf(fabsf(__uint_as_float(r[i + 3])), sc, sh));
            s += e0; s1 += e1; s2 += e2; s3 += e3;
            kmx = fmax3(fmax3(kmx, e0, e1), e2, e3);
          }
        });
        s += (s1 + s2) + s3;
      } else {
        for_each_chunk16(taddr, HALF / 16, [&](int, const uint32_t (&r)[16]) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float k = kmulti(fabsf(__uint_as_float(r[i])));
            s += k;
            kmx = fmaxf(kmx, k);
          }
        });
      }
      const int type = (r0 < p.n1) ? ((c0 < p.n1) ? 0 : 2) : ((c0 < p.n1) ? 2 : 1);
      add(type, s * w);
    } else {
      // boundary / diagonal tiles: same arithmetic, branch-free per-element masks
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      const int row = ctx.row;
      const bool row_ok = row < p.M, rx = row < p.n1;
      const bool single = p.n_alpha == 1;
      const float sc = p.a_scale[0], sh = p.a_shift[0];
      for_each_chunk16(taddr, HALF / 16, [&](int c, const uint32_t (&r)[16]) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = c0 + c * 16 + i;
          const float a = fabsf(__uint_as_float(r[i]));
          const float k = single ? ex2_approx(fmaf(a, sc, sh)) : kmulti(a);
          const bool ok = row_ok && (col < p.M) && (p.tri ? (col > row) : (col != row));
          const bool cx = col < p.n1;
          kmx = ok ? fmaxf(kmx, k) : kmx;
          s0 += (ok && rx && cx) ? k : 0.f;
          s1 += (ok && !rx && !cx) ? k : 0.f;
          s2 += (ok && (rx != cx)) ? k : 0.f;
        }
      });
      acc0 += double(s0 * w);
      acc1 += double(s1 * w);
      acc2 += double(s2 * w);
    }
  }
  CAREL_DEVINL void finish() {
    const double v0 = warp_sum(acc0), v1 = warp_sum(acc1), v2 = warp_sum(acc2);
    if ((threadIdx.x & 31) == 0) {
      if (v0 != 0.0) atomicAdd(p.accum + 0, v0);
      if (v1 != 0.0) atomicAdd(p.accum + 1, v1);
      if (v2 != 0.0) atomicAdd(p.accum + 2, v2);
    }
    if (p.kmax_bits != nullptr) {
      const float km = warp_max(kmx);        // kernel values are >= 0: their float bits order like unsigned integers
      if ((threadIdx.x & 31) == 0 && km > 0.f) atomicMax(p.kmax_bits, __float_as_uint(km));
    }
  }
};

}  // namespace carel
