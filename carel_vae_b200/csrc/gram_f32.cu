// gram_f32.cu — fp32 precision tier of the Gram losses (sgemm.cuh): MMD (drl_classifier_ec_mmd_final_mul.py:547-596) and
// HSIC (drl_classifier_ec_hsic.py:529-547) evaluated the way the reference's fp32 ATen path evaluates them — the Gram
// of dot products by an fp32 GEMM, d2 = |a|^2 + |b|^2 - 2 a.b, kernel values per element — on MATERIALISED row blocks
// of the N x N matrices (bounded workspace), any feature width.  Used by precision = "fp32" (1e-5 parity) and as the
// gradient of the bf16 tier for feature widths its fused backward cannot hold (> 250 columns, e.g. the d = 768 Gram).
//
//   MMD:   value = sum_ij C_ij K_ij,            K_ij = sum_a exp(-a (|d2_ij| + eps)),   C = a00 / a11 / a01 blocks, 0 diagonal
//          grad s_i = 4 [ (sum_j W_ij) s_i - sum_j W_ij s_j ],   W_ij = C_ij sum_a -a exp(-a (|d2| + eps)) sign(d2)
//   HSIC:  K = exp(-d2_x / s_x), L = exp(-d2_y / s_y)  (no abs, no eps),
//          value = [ sum K.L - (2/m) sum_i rK_i rL_i + (sum K)(sum L) / m^2 ] / (m-1)^2      (= trace(L H K H) / (m-1)^2)
//          dV/dK_ij = [ L_ij - (rL_i + rL_j)/m + sL/m^2 ] / (m-1)^2 =: P_ij,    Wx_ij = -P_ij K_ij / s_x
//          grad x_i = 4 [ (sum_j Wx_ij) x_i - sum_j Wx_ij x_j ]                  (and the same with K <-> L for y)
#include <algorithm>

#include "epilogues.cuh"
#include "sgemm.cuh"

namespace carel {

// rows of a materialised [rows, M] fp32 block: at most 256 MB, a multiple of 128
static inline int64_t f32_row_block(int64_t M) {
  int64_t rb = (int64_t(1) << 26) / std::max<int64_t>(M, 1);
  rb = std::max<int64_t>(128, rb / 128 * 128);
  return std::min<int64_t>(rb, round_up64(M, 128));
}

struct MmdF32Ws {
  float* S;       // [M, dim] stacked samples
  float* nrm;     // [M]
  float* W;       // [rb, M]
  float* G;       // [rb, dim]
  float* rs;      // [rb]
  double* acc;    // [4]
  size_t bytes;
  int64_t rb;
};
static MmdF32Ws carve_mmd_f32(void* ws, int64_t M, int64_t dim) {
  WsCarver w(ws);
  MmdF32Ws r;
  r.rb = f32_row_block(M);
  r.S = w.take<float>(size_t(M) * dim);
  r.nrm = w.take<float>(size_t(M));
  r.W = w.take<float>(size_t(r.rb) * M);
  r.G = w.take<float>(size_t(r.rb) * dim);
  r.rs = w.take<float>(size_t(r.rb));
  r.acc = w.take<double>(4);
  r.bytes = w.bytes();
  return r;
}

struct MmdElemParams {
  int M, n1, n_alpha;
  float c00, c11, c01, eps;
  float alpha[kMaxAlphas];
};

// One warp per row of the block: G (dot products) -> kernel values / weights in place, row-wise reductions.
//   MODE 0 (forward):  block sums of C (.) K into acc[0..2] (xx, yy, xy + yx), nothing written back
//   MODE 1 (backward): G_ij <- W_ij, rs[i] = sum_j W_ij
template <int MODE>
static __global__ void __launch_bounds__(256)
mmd_f32_elem_kernel(float* __restrict__ G, int64_t ldg, int r0, int rows, const float* __restrict__ nrm,
                    const MmdElemParams p, float* __restrict__ rs, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int rl = blockIdx.x * wpb + (threadIdx.x >> 5); rl < rows; rl += gridDim.x * wpb) {
    const int i = r0 + rl;
    float* g = G + int64_t(rl) * ldg;
    const float ni = nrm[i];
    const bool ix = i < p.n1;
    float s_same = 0.f, s_cross = 0.f;
    for (int j = lane; j < p.M; j += 32) {
      const float d2 = ni + nrm[j] - 2.f * g[j];
      const float a = fabsf(d2) + p.eps;
      const bool jx = j < p.n1;
      if (MODE == 0) {
        float k = 0.f;
        for (int t = 0; t < p.n_alpha; ++t) k += expf(-p.alpha[t] * a);
        if (j == i) k = 0.f;
        if (ix == jx) s_same += k; else s_cross += k;
      } else {
        float kp = 0.f;
        for (int t = 0; t < p.n_alpha; ++t) kp -= p.alpha[t] * expf(-p.alpha[t] * a);
        const float sgn = d2 > 0.f ? 1.f : (d2 < 0.f ? -1.f : 0.f);
        const float c = (j == i) ? 0.f : (ix == jx ? (ix ? p.c00 : p.c11) : p.c01);
        const float w = c * kp * sgn;
        g[j] = w;
        s_same += w;
      }
    }
    if (MODE == 0) {
      s_same = warp_sum(s_same);
      s_cross = warp_sum(s_cross);
      if (ix) a0 += double(s_same); else a1 += double(s_same);
      a2 += double(s_cross);
    } else {
      s_same = warp_sum(s_same);
      if (lane == 0) rs[rl] = s_same;
    }
  }
  if (MODE == 0 && lane == 0) {
    if (a0 != 0.0) atomicAdd(acc + 0, a0);
    if (a1 != 0.0) atomicAdd(acc + 1, a1);
    if (a2 != 0.0) atomicAdd(acc + 2, a2);
  }
}

// grad[i, :] = scale * g * (rs[i] * S[i, :] - G[i, :]) for the rows of one block, scattered to grad_x / grad_y
static __global__ void gram_f32_final_kernel(const float* __restrict__ S, int64_t lds, const float* __restrict__ G, int64_t ldg,
                                             const float* __restrict__ rs, int r0, int rows, int dim, int n1,
                                             const float* __restrict__ grad_out, float scale, float* __restrict__ gx,
                                             int64_t ldgx, float* __restrict__ gy, int64_t ldgy) {
  const float g = scale * __ldg(grad_out);
  const long long total = 1LL * rows * dim;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int rl = int(t / dim), c = int(t % dim);
    const int i = r0 + rl;
    const float v = g * (rs[rl] * S[int64_t(i) * lds + c] - G[int64_t(rl) * ldg + c]);
    if (i < n1) gx[int64_t(i) * ldgx + c] = v;
    else gy[int64_t(i - n1) * ldgy + c] = v;
  }
}

static __global__ void mmd_f32_finalize_kernel(const double* __restrict__ acc, double a00, double a11, double a01,
                                               float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = float(a00 * acc[0] + a11 * acc[1] + a01 * acc[2]);
}

static int mmd_f32_prepare(const MmdF32Ws& w, const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1,
                           int64_t n2, int64_t dim, cudaStream_t st) {
  CAREL_CHECK_CUDA(cudaMemcpy2DAsync(w.S, size_t(dim) * 4, x, size_t(ldx) * 4, size_t(dim) * 4, size_t(n1), cudaMemcpyDeviceToDevice, st));
  CAREL_CHECK_CUDA(cudaMemcpy2DAsync(w.S + size_t(n1) * dim, size_t(dim) * 4, y, size_t(ldy) * 4, size_t(dim) * 4, size_t(n2),
                                     cudaMemcpyDeviceToDevice, st));
  f32_rownorm_kernel<<<int(std::min<int64_t>((n1 + n2 + 7) / 8, 148 * 16)), 256, 0, st>>>(w.S, dim, int(n1 + n2), int(dim), w.nrm);
  CAREL_LAUNCHED();
  return 0;
}

static int fill_mmd_params(MmdElemParams& p, int64_t n1, int64_t n2, const float* alphas, int n_alpha, float eps) {
  CAREL_REQUIRE(n1 >= 2 && n2 >= 2, "carel_mmd: n_1 and n_2 must be >= 2 (a00 = 1/(n(n-1)))");
  CAREL_REQUIRE(n_alpha >= 1 && n_alpha <= kMaxAlphas, "carel_mmd: 1..%d alphas supported", kMaxAlphas);
  p.M = int(n1 + n2); p.n1 = int(n1); p.n_alpha = n_alpha; p.eps = eps;
  p.c00 = float(1.0 / (double(n1) * double(n1 - 1)));
  p.c11 = float(1.0 / (double(n2) * double(n2 - 1)));
  p.c01 = float(-1.0 / (double(n1) * double(n2)));
  for (int i = 0; i < kMaxAlphas; ++i) p.alpha[i] = i < n_alpha ? alphas[i] : 0.f;
  return 0;
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_mmd_f32_workspace_bytes(int64_t n1, int64_t n2, int64_t dim) {
  return carve_mmd_f32(nullptr, n1 + n2, dim).bytes;
}

// Block sums of the rows of tiles [row_tile_begin, row_tile_end) against EVERY column (full Gram rows, diagonal
// excluded): partial_sums (3 doubles, caller all-reduces them) or, with out_mmd, the finished value.
extern "C" int carel_mmd_f32_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1, int64_t n2,
                                 int64_t dim, const float* alphas_host, int n_alpha, float eps, int64_t row_tile_begin,
                                 int64_t row_tile_end, double* partial_sums, float* out_mmd, void* ws, size_t ws_bytes,
                                 void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MmdElemParams p;
  if (int rc = fill_mmd_params(p, n1, n2, alphas_host, n_alpha, eps)) return rc;
  const int64_t M = n1 + n2;
  MmdF32Ws w = carve_mmd_f32(ws, M, dim);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_mmd_f32_fwd: workspace too small");
  if (int rc = mmd_f32_prepare(w, x, ldx, y, ldy, n1, n2, dim, st)) return rc;
  double* acc = partial_sums ? partial_sums : w.acc;
  CAREL_CHECK_CUDA(cudaMemsetAsync(acc, 0, 3 * sizeof(double), st));
  const int64_t m_tiles = (M + 127) / 128;
  const int64_t t0 = row_tile_begin < 0 ? 0 : row_tile_begin;
  const int64_t t1 = (row_tile_end < 0 || row_tile_end > m_tiles) ? m_tiles : row_tile_end;
  for (int64_t r0 = t0 * 128; r0 < std::min(M, t1 * 128); r0 += w.rb) {
    const int64_t rows = std::min<int64_t>(w.rb, std::min(M, t1 * 128) - r0);
    if (int rc = launch_sgemm(false, false, w.S + r0 * dim, dim, w.S, dim, w.W, M, rows, M, dim, 1.f, nullptr, 0, 1, st)) return rc;
    mmd_f32_elem_kernel<0><<<int(std::min<int64_t>((rows + 7) / 8, 148 * 8)), 256, 0, st>>>(w.W, M, int(r0), int(rows), w.nrm, p,
                                                                                      nullptr, acc);
    CAREL_LAUNCHED();
  }
  if (out_mmd != nullptr) {
    mmd_f32_finalize_kernel<<<1, 32, 0, st>>>(acc, 1.0 / (double(n1) * double(n1 - 1)), 1.0 / (double(n2) * double(n2 - 1)),
                                             -1.0 / (double(n1) * double(n2)), out_mmd);
    CAREL_LAUNCHED();
  }
  return 0;
}

extern "C" int carel_mmd_f32_bwd(const float* grad_out, const float* x, int64_t ldx, const float* y, int64_t ldy,
                                 int64_t n1, int64_t n2, int64_t dim, const float* alphas_host, int n_alpha, float eps,
                                 int64_t row_tile_begin, int64_t row_tile_end, float* grad_x, int64_t ldgx, float* grad_y,
                                 int64_t ldgy, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MmdElemParams p;
  if (int rc = fill_mmd_params(p, n1, n2, alphas_host, n_alpha, eps)) return rc;
  const int64_t M = n1 + n2;
  MmdF32Ws w = carve_mmd_f32(ws, M, dim);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_mmd_f32_bwd: workspace too small");
  CAREL_REQUIRE(grad_x != nullptr && grad_y != nullptr, "carel_mmd_f32_bwd: both gradients are required");
  if (int rc = mmd_f32_prepare(w, x, ldx, y, ldy, n1, n2, dim, st)) return rc;
  const int64_t m_tiles = (M + 127) / 128;
  const int64_t t0 = row_tile_begin < 0 ? 0 : row_tile_begin;
  const int64_t t1 = (row_tile_end < 0 || row_tile_end > m_tiles) ? m_tiles : row_tile_end;
  if (t0 > 0 || t1 < m_tiles) {      // rows of other row blocks read as zero
    CAREL_CHECK_CUDA(cudaMemset2DAsync(grad_x, size_t(ldgx) * 4, 0, size_t(dim) * 4, size_t(n1), st));
    CAREL_CHECK_CUDA(cudaMemset2DAsync(grad_y, size_t(ldgy) * 4, 0, size_t(dim) * 4, size_t(n2), st));
  }
  for (int64_t r0 = t0 * 128; r0 < std::min(M, t1 * 128); r0 += w.rb) {
    const int64_t rows = std::min<int64_t>(w.rb, std::min(M, t1 * 128) - r0);
    if (int rc = launch_sgemm(false, false, w.S + r0 * dim, dim, w.S, dim, w.W, M, rows, M, dim, 1.f, nullptr, 0, 1, st)) return rc;
    mmd_f32_elem_kernel<1><<<int(std::min<int64_t>((rows + 7) / 8, 148 * 8)), 256, 0, st>>>(w.W, M, int(r0), int(rows), w.nrm, p,
                                                                                      w.rs, nullptr);
    CAREL_LAUNCHED();
    if (int rc = launch_sgemm(false, true, w.W, M, w.S, dim, w.G, dim, rows, dim, M, 1.f, nullptr, 0, 1, st)) return rc;
    gram_f32_final_kernel<<<int(std::min<int64_t>((rows * dim + 255) / 256, 148 * 16)), 256, 0, st>>>(
        w.S, dim, w.G, dim, w.rs, int(r0), int(rows), int(dim), int(n1), grad_out, 4.f, grad_x, ldgx, grad_y, ldgy);
    CAREL_LAUNCHED();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// HSIC
// ---------------------------------------------------------------------------------------------
namespace carel {

struct HsicF32Ws {
  float* K;        // [m, m]  kernel matrix of x; overwritten by Wx in backward
  float* Lm;       // [m, m]  kernel matrix of y; overwritten by Wy in backward
  float* nx;       // [m]
  float* ny;       // [m]
  double* rK;      // [m] row sums
  double* rL;      // [m]
  double* acc;     // [4]  sum K.L, sum_i rK_i rL_i (finalize), sum K, sum L
  float* rsx;      // [m] row sums of Wx
  float* rsy;      // [m]
  float* G;        // [m, max(dx, dy)]
  size_t bytes;
};
static HsicF32Ws carve_hsic_f32(void* ws, int64_t m, int64_t dx, int64_t dy) {
  WsCarver w(ws);
  HsicF32Ws r;
  r.K = w.take<float>(size_t(m) * m);
  r.Lm = w.take<float>(size_t(m) * m);
  r.nx = w.take<float>(size_t(m));
  r.ny = w.take<float>(size_t(m));
  r.rK = w.take<double>(size_t(m));
  r.rL = w.take<double>(size_t(m));
  r.acc = w.take<double>(4);
  r.rsx = w.take<float>(size_t(m));
  r.rsy = w.take<float>(size_t(m));
  r.G = w.take<float>(size_t(m) * std::max(dx, dy));
  r.bytes = w.bytes();
  return r;
}

// G (dot products) -> K = exp(-(n_i + n_j - 2 G) / sigma) in place, one warp per row; row sums in double
static __global__ void __launch_bounds__(256)
hsic_f32_kernel_matrix(float* __restrict__ G, int m, const float* __restrict__ nrm, float inv_sigma, double* __restrict__ rsum) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < m; i += gridDim.x * wpb) {
    float* g = G + int64_t(i) * m;
    const float ni = nrm[i];
    double s = 0.0;
    for (int j = lane; j < m; j += 32) {
      const float k = expf(-(ni + nrm[j] - 2.f * g[j]) * inv_sigma);
      g[j] = k;
      s += double(k);
    }
    s = warp_sum(s);
    if (lane == 0) rsum[i] = s;
  }
}

// acc[0] += sum_ij K_ij L_ij;  acc[2] += sum_i rK_i;  acc[3] += sum_i rL_i;  acc[1] += sum_i rK_i rL_i
static __global__ void __launch_bounds__(256)
hsic_f32_reduce_kernel(const float* __restrict__ K, const float* __restrict__ Lm, int m, const double* __restrict__ rK,
                       const double* __restrict__ rL, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double kl = 0.0, rr = 0.0, sk = 0.0, sl = 0.0;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < m; i += gridDim.x * wpb) {
    const float* k = K + int64_t(i) * m;
    const float* l = Lm + int64_t(i) * m;
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s = fmaf(k[j], l[j], s);
    kl += double(warp_sum(s));
    if (lane == 0) { rr += rK[i] * rL[i]; sk += rK[i]; sl += rL[i]; }
  }
  if (lane == 0) {
    atomicAdd(acc + 0, kl);
    atomicAdd(acc + 1, rr);
    atomicAdd(acc + 2, sk);
    atomicAdd(acc + 3, sl);
  }
}

static __global__ void hsic_f32_finalize_kernel(const double* __restrict__ acc, int m, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double dm = double(m);
    out[0] = float((acc[0] - 2.0 / dm * acc[1] + acc[2] * acc[3] / (dm * dm)) / ((dm - 1.0) * (dm - 1.0)));
  }
}

// K <- Wx = -P^K (.) K / s_x,  L <- Wy = -P^L (.) L / s_y  with  P^K_ij = [L_ij - (rL_i + rL_j)/m + sL/m^2] / (m-1)^2
// (and K <-> L); row sums of both
static __global__ void __launch_bounds__(256)
hsic_f32_weights_kernel(float* __restrict__ K, float* __restrict__ Lm, int m, const double* __restrict__ rK,
                        const double* __restrict__ rL, const double* __restrict__ acc, float inv_sx, float inv_sy,
                        float* __restrict__ rsx, float* __restrict__ rsy) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const double dm = double(m), inv_m = 1.0 / dm, inv_m2 = 1.0 / (dm * dm), nrmz = 1.0 / ((dm - 1.0) * (dm - 1.0));
  const double sK = acc[2], sL = acc[3];
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < m; i += gridDim.x * wpb) {
    float* k = K + int64_t(i) * m;
    float* l = Lm + int64_t(i) * m;
    const double rKi = rK[i], rLi = rL[i];
    float sx = 0.f, sy = 0.f;
    for (int j = lane; j < m; j += 32) {
      const float kv = k[j], lv = l[j];
      const float pk = float((double(lv) - (rLi + rL[j]) * inv_m + sL * inv_m2) * nrmz);
      const float pl = float((double(kv) - (rKi + rK[j]) * inv_m + sK * inv_m2) * nrmz);
      // the diagonal contributes W_ii (x_i - x_i) = 0: dropping it keeps the O(1) entries W_ii out of the difference
      // rs_i x_i - (W x)_i, whose fp32 cancellation noise would otherwise swamp a small gradient
      const float wx = (j == i) ? 0.f : -pk * kv * inv_sx, wy = (j == i) ? 0.f : -pl * lv * inv_sy;
      k[j] = wx;
      l[j] = wy;
      sx += wx;
      sy += wy;
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if (lane == 0) { rsx[i] = sx; rsy[i] = sy; }
  }
}

}  // namespace carel

extern "C" size_t carel_hsic_f32_workspace_bytes(int64_t m, int64_t dx, int64_t dy) {
  return carve_hsic_f32(nullptr, m, dx, dy).bytes;
}

extern "C" int carel_hsic_f32_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m, int64_t dx, int64_t dy,
                                  float s_x, float s_y, float* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(m >= 2 && dx > 0 && dy > 0, "carel_hsic_f32_fwd: need m >= 2 paired samples");
  CAREL_REQUIRE(m <= 46340, "carel_hsic_f32_fwd: the fp32 tier materialises the two m x m kernel matrices (m <= 46340)");
  HsicF32Ws w = carve_hsic_f32(ws, m, dx, dy);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_hsic_f32_fwd: workspace too small");
  const int rgrid = int(std::min<int64_t>((m + 7) / 8, 148 * 8));
  f32_rownorm_kernel<<<rgrid, 256, 0, st>>>(x, ldx, int(m), int(dx), w.nx);
  CAREL_LAUNCHED();
  f32_rownorm_kernel<<<rgrid, 256, 0, st>>>(y, ldy, int(m), int(dy), w.ny);
  CAREL_LAUNCHED();
  if (int rc = launch_sgemm(false, false, x, ldx, x, ldx, w.K, m, m, m, dx, 1.f, nullptr, 0, 1, st)) return rc;
  if (int rc = launch_sgemm(false, false, y, ldy, y, ldy, w.Lm, m, m, m, dy, 1.f, nullptr, 0, 1, st)) return rc;
  hsic_f32_kernel_matrix<<<rgrid, 256, 0, st>>>(w.K, int(m), w.nx, 1.f / s_x, w.rK);
  CAREL_LAUNCHED();
  hsic_f32_kernel_matrix<<<rgrid, 256, 0, st>>>(w.Lm, int(m), w.ny, 1.f / s_y, w.rL);
  CAREL_LAUNCHED();
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, 4 * sizeof(double), st));
  hsic_f32_reduce_kernel<<<rgrid, 256, 0, st>>>(w.K, w.Lm, int(m), w.rK, w.rL, w.acc);
  CAREL_LAUNCHED();
  hsic_f32_finalize_kernel<<<1, 32, 0, st>>>(w.acc, int(m), out);
  CAREL_LAUNCHED();
  return 0;
}

// Must follow carel_hsic_f32_fwd on the same workspace (K, L, row sums are read from it and overwritten).
extern "C" int carel_hsic_f32_bwd(const float* grad_out, const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m,
                                  int64_t dx, int64_t dy, float s_x, float s_y, float* grad_x, int64_t ldgx, float* grad_y,
                                  int64_t ldgy, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HsicF32Ws w = carve_hsic_f32(ws, m, dx, dy);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_hsic_f32_bwd: workspace too small");
  const int rgrid = int(std::min<int64_t>((m + 7) / 8, 148 * 8));
  hsic_f32_weights_kernel<<<rgrid, 256, 0, st>>>(w.K, w.Lm, int(m), w.rK, w.rL, w.acc, 1.f / s_x, 1.f / s_y, w.rsx, w.rsy);
  CAREL_LAUNCHED();
  if (grad_x != nullptr) {
    if (int rc = launch_sgemm(false, true, w.K, m, x, ldx, w.G, dx, m, dx, m, 1.f, nullptr, 0, 1, st)) return rc;
    gram_f32_final_kernel<<<int(std::min<int64_t>((m * dx + 255) / 256, 148 * 16)), 256, 0, st>>>(
        x, ldx, w.G, dx, w.rsx, 0, int(m), int(dx), int(m), grad_out, 4.f, grad_x, ldgx, nullptr, 0);
    CAREL_LAUNCHED();
  }
  if (grad_y != nullptr) {
    if (int rc = launch_sgemm(false, true, w.Lm, m, y, ldy, w.G, dy, m, dy, m, 1.f, nullptr, 0, 1, st)) return rc;
    gram_f32_final_kernel<<<int(std::min<int64_t>((m * dy + 255) / 256, 148 * 16)), 256, 0, st>>>(
        y, ldy, w.G, dy, w.rsy, 0, int(m), int(dy), int(m), grad_out, 4.f, grad_y, ldgy, nullptr, 0);
    CAREL_LAUNCHED();
  }
  return 0;
}
