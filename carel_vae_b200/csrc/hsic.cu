// hsic.cu — carel_hsic_{fwd,bwd}: the Hilbert-Schmidt independence criterion of
// drl_classifier_ec_hsic.py:529-547,
//     K = exp(-d2_x / s_x), L = exp(-d2_y / s_y)  (no abs, no eps),  H = I - 11^T/m,
//     HSIC = trace(L H K H) / (m-1)^2,
// evaluated with the O(m^2) identity (verified against the reference):
//     HSIC = [ sum K.L - (2/m) sum_i rK_i rL_i + (sum K)(sum L)/m^2 ] / (m-1)^2
// instead of the reference's three dense m x m matmuls.  Both Grams of a tile are produced by the
// tensor cores into TMEM (two products per tile) and reduced in registers.
// Backward (same identity differentiated):
//     grad x_i = 4/(m-1)^2 [ (sum_j Wx_ij) x_i - sum_j Wx_ij x_j ],  Wx = -(1/s_x) (HLH) (.) K,
//     (HLH)_ij = L_ij - rL_i/m - rL_j/m + sum L / m^2              (and symmetrically for y).
#include <algorithm>

#include "api_common.cuh"
#include "chain_launch.cuh"
#include "chain_policies.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "pack.cuh"

namespace carel {

struct HsicWs {
  __nv_bfloat16 *XA, *XB, *YA, *YB;   // [m, ldp] augmented operands
  double* rK;       // [mpad] row sums of K
  double* rL;       // [mpad]
  double* acc;      // [4]  sum K.L, sum K, sum L, sum_i rK_i rL_i
  float* cK;        // [mpad] -rK_j / m
  float* cL;        // [mpad] -rL_j / m
  float* mean;      // [2]  sum K / m^2, sum L / m^2
  size_t bytes;
  int ldp, mpad;
};

static HsicWs carve_hsic(void* ws, int64_t m, int64_t d) {
  WsCarver w(ws);
  HsicWs r;
  r.ldp = aug_width(int(d));
  r.mpad = int(round_up64(m, 256));
  r.XA = w.take<__nv_bfloat16>(size_t(m) * r.ldp);
  r.XB = w.take<__nv_bfloat16>(size_t(m) * r.ldp);
  r.YA = w.take<__nv_bfloat16>(size_t(m) * r.ldp);
  r.YB = w.take<__nv_bfloat16>(size_t(m) * r.ldp);
  r.rK = w.take<double>(size_t(r.mpad));
  r.rL = w.take<double>(size_t(r.mpad));
  r.acc = w.take<double>(4);
  r.cK = w.take<float>(size_t(r.mpad));
  r.cL = w.take<float>(size_t(r.mpad));
  r.mean = w.take<float>(2);
  r.bytes = w.bytes();
  return r;
}

// ---------------------------------------------------------------------------------------------
// forward epilogue: two accumulators per tile (x-Gram at columns [0,BN), y-Gram at [BN,2BN))
// ---------------------------------------------------------------------------------------------
struct EpiHsicParams {
  double* rK;
  double* rL;
  double* acc;
  int m;
  float cx, cy;    // 2 log2(e) / s_x, 2 log2(e) / s_y   (acc = -d2/2  =>  K = 2^(acc * cx))
};
template <int BN>
struct EpiHsic {
  using Params = EpiHsicParams;
  const Params& p;
  double sKL, sK, sL;
  CAREL_DEVINL EpiHsic(const Params& p_, const TileCtx&) : p(p_), sKL(0.0), sK(0.0), sL(0.0) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int c0 = ctx.col0 + ctx.half * HALF;
    if (c0 >= p.m || ctx.tm * 128 >= p.m) return;
    const bool full = (c0 + HALF <= p.m) && (ctx.tm * 128 + 128 <= p.m);
    const bool row_ok = ctx.row < p.m;
    float tKL = 0.f, tK = 0.f, tL = 0.f;
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t rx[16], ry[16];
      tmem_ld16(ctx.taddr + ctx.half * HALF + c * 16, rx);
      tmem_ld16(ctx.taddr + BN + ctx.half * HALF + c * 16, ry);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float K = ex2_approx(__uint_as_float(rx[i]) * p.cx);
        float L = ex2_approx(__uint_as_float(ry[i]) * p.cy);
        if (!full) {
          const bool ok = row_ok && (c0 + c * 16 + i < p.m);
          K = ok ? K : 0.f;
          L = ok ? L : 0.f;
        }
        tK += K;
        tL += L;
        tKL = fmaf(K, L, tKL);
      }
    }
    if (row_ok) {
      atomicAdd(p.rK + ctx.row, double(tK));
      atomicAdd(p.rL + ctx.row, double(tL));
      sKL += double(tKL);
      sK += double(tK);
      sL += double(tL);
    }
  }
  CAREL_DEVINL void finish() {
    const double a = warp_sum(sKL), b = warp_sum(sK), c = warp_sum(sL);
    if ((threadIdx.x & 31) == 0) {
      if (a != 0.0) atomicAdd(p.acc + 0, a);
      if (b != 0.0) atomicAdd(p.acc + 1, b);
      if (c != 0.0) atomicAdd(p.acc + 2, c);
    }
  }
};

// single block: t2 = sum_i rK_i rL_i, the final value, and the vectors the backward pass needs
static __global__ void hsic_finalize_kernel(const double* __restrict__ rK, const double* __restrict__ rL,
                                            double* __restrict__ acc, int m, float* __restrict__ cK,
                                            float* __restrict__ cL, float* __restrict__ mean,
                                            float* __restrict__ out) {
  __shared__ double red[32];
  double t = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    t += rK[i] * rL[i];
    cK[i] = float(-rK[i] / double(m));
    cL[i] = float(-rL[i] / double(m));
  }
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      acc[3] = v;
      const double dm = double(m);
      const double h = (acc[0] - 2.0 / dm * v + acc[1] * acc[2] / (dm * dm)) / ((dm - 1.0) * (dm - 1.0));
      if (out != nullptr) out[0] = float(h);
      mean[0] = float(acc[1] / (dm * dm));
      mean[1] = float(acc[2] / (dm * dm));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward policy.  Product 0 is the side being differentiated ("self"), product 1 the other.
//   w_ij = -(1/s_self) K_self_ij (K_other_ij + cO_i + cO_j + meanO),   cO_i = -rOther_i / m
// ---------------------------------------------------------------------------------------------
struct HsicBwdParams {
  const __nv_bfloat16* SA;   // augmented row operand of the self side
  const float* c_other;      // [mpad]
  const float* mean_other;   // scalar
  const float* grad_out;
  float* grad;
  int ldg, dim, ldp, m;
  float c_self, c_oth;       // 2 log2 e / s
  float neg_inv_s;           // -1 / s_self
  float norm;                // 4 / (m-1)^2
};
struct PolHsicBwd {
  using Params = HsicBwdParams;
  static constexpr bool kHiLo = true;
  static constexpr bool kF16 = false;
  const Params& p;
  float ci;
  CAREL_DEVINL PolHsicBwd(const Params& p_, const ChainCtx&) : p(p_), ci(0.f) {}
  CAREL_DEVINL void begin_item(const ChainCtx& ctx) {
    ci = __ldg(p.c_other + min(ctx.row, p.m - 1)) + __ldg(p.mean_other);
  }
  CAREL_DEVINL void begin_tile(const ChainCtx&, int, int) {}
  CAREL_DEVINL void prefetch(const ChainCtx&, int) {}
  CAREL_DEVINL void weights16(const ChainCtx& ctx, int jc, const uint32_t (&rs)[16], const uint32_t (&ro)[16],
                              float (&w)[16]) {
    const float4* cj4 = reinterpret_cast<const float4*>(p.c_other + jc);
    float cj[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(cj4 + i);
      cj[4 * i] = v.x; cj[4 * i + 1] = v.y; cj[4 * i + 2] = v.z; cj[4 * i + 3] = v.w;
    }
    const bool row_ok = ctx.row < p.m;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float Ks = ex2_approx(__uint_as_float(rs[e]) * p.c_self);
      const float Ko = ex2_approx(__uint_as_float(ro[e]) * p.c_oth);
      const bool ok = row_ok && (jc + e < p.m);
      w[e] = ok ? p.neg_inv_s * Ks * (Ko + ci + cj[e]) : 0.f;
    }
  }
  CAREL_DEVINL void store16(const ChainCtx&, int, const uint32_t (&)[8]) {}
  template <class Release>
  CAREL_DEVINL void final(const ChainCtx& ctx, bool valid, Release&& release) {
    const uint32_t rtot_bits = tmem_ld1(ctx.acc2_taddr + p.dim);   // ones column: sum_j w_ij
    tmem_ld_wait();
    const float rtot = __uint_as_float(rtot_bits);
    const int q0 = (ctx.n2_16 * ctx.half) / kChainSplit, q1 = (ctx.n2_16 * (ctx.half + 1)) / kChainSplit;
    const float g = p.norm * __ldg(p.grad_out);
    const bool row_ok = valid && ctx.row < p.m;
    const int rowc = min(ctx.row, p.m - 1);
    const __nv_bfloat16* srow = p.SA + size_t(rowc) * p.ldp;
    float* dst = p.grad + size_t(rowc) * p.ldg;
#pragma unroll 1
    for (int q = q0; q < q1; ++q) {
      uint32_t r[16];
      tmem_ld16(ctx.acc2_taddr + q * 16, r);
      const uint4 sv0 = __ldg(reinterpret_cast<const uint4*>(srow + q * 16));
      const uint4 sv1 = __ldg(reinterpret_cast<const uint4*>(srow + q * 16) + 1);
      tmem_ld_wait();
      if (!row_ok) continue;
      const uint32_t u[8] = {sv0.x, sv0.y, sv0.z, sv0.w, sv1.x, sv1.y, sv1.z, sv1.w};
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const float2 s2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[h]));
        const int f = q * 16 + 2 * h;
        if (f < p.dim) atomicAdd(dst + f, g * (rtot * s2.x - __uint_as_float(r[2 * h])));
        if (f + 1 < p.dim) atomicAdd(dst + f + 1, g * (rtot * s2.y - __uint_as_float(r[2 * h + 1])));
      }
    }
    release();
  }
  CAREL_DEVINL void finish() {}
};

}  // namespace carel

using namespace carel;

extern "C" size_t carel_hsic_workspace_bytes(int64_t m, int64_t dx, int64_t dy) {
  return carve_hsic(nullptr, m, std::max(dx, dy)).bytes;
}

// Byte offsets inside the workspace of the quantities a multi-GPU caller all-reduces between the
// partial forward and carel_hsic_finalize: rK [m] double, rL [m] double, acc [4] double.
extern "C" int carel_hsic_ws_layout(int64_t m, int64_t dx, int64_t dy, int64_t* offsets3) {
  const uintptr_t base = 1 << 20;   // any non-null, 256-aligned pseudo base: only the differences matter
  HsicWs w = carve_hsic(reinterpret_cast<void*>(base), m, std::max(dx, dy));
  offsets3[0] = int64_t(reinterpret_cast<uintptr_t>(w.rK) - base);
  offsets3[1] = int64_t(reinterpret_cast<uintptr_t>(w.rL) - base);
  offsets3[2] = int64_t(reinterpret_cast<uintptr_t>(w.acc) - base);
  return 0;
}

extern "C" int carel_hsic_finalize(int64_t m, int64_t dx, int64_t dy, float* out, void* ws, size_t ws_bytes,
                                   void* stream) {
  HsicWs w = carve_hsic(ws, m, std::max(dx, dy));
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_hsic_finalize: workspace too small");
  hsic_finalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(w.rK, w.rL, w.acc, int(m), w.cK, w.cL, w.mean, out);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_hsic_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m,
                              int64_t dx, int64_t dy, float s_x, float s_y, int64_t row_tile_begin,
                              int64_t row_tile_end, float* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(m >= 2 && dx > 0 && dy > 0, "carel_hsic_fwd: need m >= 2 and non-empty features");
  CAREL_REQUIRE(s_x != 0.f && s_y != 0.f, "carel_hsic_fwd: zero bandwidth");
  HsicWs w = carve_hsic(ws, m, std::max(dx, dy));
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_hsic_fwd: workspace too small");
  pack_rows_kernel<<<int((m + 7) / 8), 256, 0, st>>>(x, int(m), int(dx), int(ldx), w.XA, w.XB, w.ldp, 0, 1, nullptr);
  CAREL_LAUNCHED();
  pack_rows_kernel<<<int((m + 7) / 8), 256, 0, st>>>(y, int(m), int(dy), int(ldy), w.YA, w.YB, w.ldp, 0, 1, nullptr);
  CAREL_LAUNCHED();
  // NB: with dx != dy the augmentation columns sit at different offsets (dx / dy); both rows are
  // zero padded to the common pitch, so the dot products are unaffected.
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.rK, 0, size_t(w.mpad) * 8, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.rL, 0, size_t(w.mpad) * 8, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, 4 * 8, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.cK, 0, size_t(w.mpad) * 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.cL, 0, size_t(w.mpad) * 4, st));
  CUtensorMap xa, xb, ya, yb;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&xa, w.XA, m, w.ldp, w.ldp, 128) == 0 &&
                make_tmap_bf16_sw128(&xb, w.XB, m, w.ldp, w.ldp, 128) == 0 &&
                make_tmap_bf16_sw128(&ya, w.YA, m, w.ldp, w.ldp, 128) == 0 &&
                make_tmap_bf16_sw128(&yb, w.YB, m, w.ldp, w.ldp, 128) == 0, "carel_hsic_fwd: tensor map failed");
  GemmSched s = make_sched(m, m, w.ldp, 128, 1, TILE_FULL);
  const int t0 = int(row_tile_begin < 0 ? 0 : row_tile_begin);
  const int t1 = int((row_tile_end < 0 || row_tile_end > s.m_tiles) ? s.m_tiles : row_tile_end);
  CAREL_REQUIRE(t0 <= t1, "carel_hsic_fwd: bad row tile range");
  s.m_tile0 = t0;
  s.tri_rows = t1 - t0;
  s.num_items = (t1 - t0) * s.n_tiles;
  EpiHsicParams ep;
  ep.rK = w.rK; ep.rL = w.rL; ep.acc = w.acc; ep.m = int(m);
  ep.cx = 2.f * kLog2e / s_x; ep.cy = 2.f * kLog2e / s_y;
  if (int rc = launch_gemm<EpiHsic<128>, 128, 2, false, false>(xa, xb, ya, yb, s, ep, st, "hsic_gram_fwd")) return rc;
  // out == NULL: partial forward of a row block (multi-GPU); the caller all-reduces rK, rL, acc
  // (carel_hsic_ws_layout) and calls carel_hsic_finalize.
  if (out != nullptr) return carel_hsic_finalize(m, dx, dy, out, ws, ws_bytes, stream);
  return 0;
}

extern "C" int carel_hsic_bwd(const float* grad_out, int64_t m, int64_t dx, int64_t dy, float s_x,
                              float s_y, int64_t row_tile_begin, int64_t row_tile_end, float* grad_x,
                              int64_t ldgx, float* grad_y, int64_t ldgy, void* ws, size_t ws_bytes,
                              void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HsicWs w = carve_hsic(ws, m, std::max(dx, dy));
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_hsic_bwd: workspace too small");
  using Cfg = ChainCfg<2, 64>;
  const int k16 = w.ldp / 16;
  CAREL_REQUIRE(16 * k16 <= Cfg::kMaxN2, "carel_hsic_bwd: feature width > %d is not supported by the fused backward", Cfg::kMaxN2 - 16);
  CUtensorMap xa, xb, ya, yb;
  {
    const int r0 = make_tmap_bf16_sw128(&xa, w.XA, m, w.ldp, w.ldp, 128);
    const int r1 = make_tmap_bf16_sw128(&xb, w.XB, m, w.ldp, w.ldp, 64);
    const int r2 = make_tmap_bf16_sw128(&ya, w.YA, m, w.ldp, w.ldp, 128);
    const int r3 = make_tmap_bf16_sw128(&yb, w.YB, m, w.ldp, w.ldp, 64);
    CAREL_REQUIRE((r0 | r1 | r2 | r3) == 0, "carel_hsic_bwd: tensor map failed (%d %d %d %d; m=%lld ldp=%d)", r0, r1, r2, r3, (long long)m, w.ldp);
  }
  const int m_tiles = int((m + 127) / 128);
  const int t0 = int(row_tile_begin < 0 ? 0 : row_tile_begin);
  const int t1 = int((row_tile_end < 0 || row_tile_end > m_tiles) ? m_tiles : row_tile_end);
  CAREL_REQUIRE(t0 <= t1, "carel_hsic_bwd: bad row tile range");
  ChainSched cs = make_chain_sched(t0, t1 - t0, int((m + 63) / 64), k16, Cfg::stages_for(k16), 32);
  const float norm = float(4.0 / ((double(m) - 1.0) * (double(m) - 1.0)));
  for (int side = 0; side < 2; ++side) {
    float* g = side ? grad_y : grad_x;
    if (g == nullptr) continue;
    const int64_t ldg = side ? ldgy : ldgx;
    const int64_t dim = side ? dy : dx;
    CAREL_CHECK_CUDA(cudaMemset2DAsync(g, size_t(ldg) * 4, 0, size_t(dim) * 4, size_t(m), st));
    HsicBwdParams pp;
    pp.SA = side ? w.YA : w.XA;
    pp.c_other = side ? w.cK : w.cL;
    pp.mean_other = side ? w.mean + 0 : w.mean + 1;
    pp.grad_out = grad_out; pp.grad = g; pp.ldg = int(ldg); pp.dim = int(dim); pp.ldp = w.ldp; pp.m = int(m);
    pp.c_self = 2.f * kLog2e / (side ? s_y : s_x);
    pp.c_oth = 2.f * kLog2e / (side ? s_x : s_y);
    pp.neg_inv_s = -1.f / (side ? s_y : s_x);
    pp.norm = norm;
    int rc = side ? launch_chain<PolHsicBwd, 2, 64>(ya, yb, xa, xb, cs, pp, st, "hsic_gram_bwd")
                  : launch_chain<PolHsicBwd, 2, 64>(xa, xb, ya, yb, cs, pp, st, "hsic_gram_bwd");
    if (rc) return rc;
  }
  return 0;
}
