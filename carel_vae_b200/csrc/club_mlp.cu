// club_mlp.cu — carel_club_mlp_{fwd,bwd}: the CLUB approximation network q(e | c) of the VI trainer
// (drl_classifier_ec_vi_final.py:155-162 ec_mu / ec_log_var, get_ec_emb :342-347):
//
//   mu = W2 relu(W1 c + b1) + b2,      lv = tanh(W4 relu(W3 c + b3) + b4),      c = z_c [N, L],  W* [L, L],  L <= 128
//
// Both MLPs, both layers, in ONE kernel per direction: a CTA owns 32 clause rows; the four weight matrices pass through
// shared memory one after the other (k-major copies prepared once per call), activations stay on chip between the
// layers; each thread computes a 4 x 4 register block per layer (two 128-bit shared-memory loads per 16 FMAs).  fp32
// FFMA arithmetic: 4 L^2 N MACs = 0.5 GFLOP at N = 8192, L = 128, i.e. launch-latency territory — the reference spends
// 8 small cuBLAS GEMMs + ~10 elementwise launches forward and ~25 backward on it.
// Backward: the same tiling produces dh / d(pre) and d c in one kernel (bias gradients through shared-memory column
// sums); the four weight gradients are [L, N] x [N, L] reductions over the ROWS and go through the split-K fp32 GEMM.
#include <algorithm>

#include "sgemm.cuh"

namespace carel {

constexpr int kMlpRows = 32;          // clause rows per CTA
constexpr int kMlpMaxL = 128;

// dst[k * L + j] = src[j * L + k] for the four matrices (blockIdx.y): k-major copies for the forward products
struct Mlp4 {
  const float* w[4];
};
static __global__ void club_mlp_transpose_kernel(const Mlp4 src, int L, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const float* s = src.w[blockIdx.z];
  float* d = dst + size_t(blockIdx.z) * L * L;
  const int j0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int j = j0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (j < L && k < L) ? s[size_t(j) * L + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, j = j0 + threadIdx.x;
    if (k < L && j < L) d[size_t(k) * L + j] = tile[threadIdx.x][i];
  }
}

// out[r][j] (4 x 4 block of this thread) = sum_k At[k][r] * Wk[k][j]   (both operands k-major in shared memory)
static CAREL_DEVINL void mlp_block_product(const float* __restrict__ At, const float* __restrict__ Wk, int L, int r0, int j0,
                                           float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < L; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(At + k * kMlpRows + r0);
    const float4 b = *reinterpret_cast<const float4*>(Wk + k * kMlpMaxL + j0);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// weights (k-major, [L][L]) -> shared [L][128]
static CAREL_DEVINL void mlp_load_weights(const float* __restrict__ Wk_g, int L, float* __restrict__ Ws) {
  // L % 4 == 0 and 16-byte aligned parameter storage: 128-bit loads
  for (int t = threadIdx.x; t < L * (kMlpMaxL / 4); t += blockDim.x) {
    const int k = t / (kMlpMaxL / 4), j = (t % (kMlpMaxL / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < L) v = __ldg(reinterpret_cast<const float4*>(Wk_g + size_t(k) * L + j));
    *reinterpret_cast<float4*>(Ws + k * kMlpMaxL + j) = v;
  }
}

struct ClubMlpArgs {
  const float* c; int64_t ldc;      // [N, L] input (row pitch ldc)
  const float* wk;                  // forward: 4 k-major weight copies [4][L][L] (W1^T, W2^T, W3^T, W4^T as [k][j])
  const float* w[4];                // backward: the weights as stored, [j][k] = k-major for d(input) products
  const float* b[4];
  int N, L;
  float* h1; float* h2;             // [N, L] relu outputs (saved)
  float* mu; float* lv;             // [N, L] outputs
  // backward
  const float* dmu; const float* dlv;   // [N, L] upstream
  float* dh1; float* dh2;               // [N, L] d(relu input) of the two hidden layers
  float* dpre;                          // [N, L] d(tanh input) = dlv (1 - lv^2)
  float* dc; int64_t lddc;              // [N, L] d(input) or NULL
  float* db[4];                         // accumulated (zeroed by the host side)
};

static __global__ void __launch_bounds__(256, 2) club_mlp_fwd_kernel(const ClubMlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                                     // [L][128]
  float* xt = Ws + kMlpMaxL * kMlpMaxL;               // [L][32]  input tile, k-major
  float* ht = xt + kMlpMaxL * kMlpRows;               // [L][32]  hidden tile, k-major
  const int L = a.L, row0 = blockIdx.x * kMlpRows;
  const int rg = threadIdx.x >> 5, cg = threadIdx.x & 31;     // 8 row groups x 32 column groups of 4
  const int r0 = rg * 4, j0 = cg * 4;
  for (int t = threadIdx.x; t < L * kMlpRows; t += blockDim.x) {   // lanes along the rows: conflict-free transposed store
    const int r = t % kMlpRows, k = t / kMlpRows;
    xt[k * kMlpRows + r] = (row0 + r < a.N) ? a.c[size_t(row0 + r) * a.ldc + k] : 0.f;
  }
  for (int net = 0; net < 2; ++net) {
    float acc[4][4];
    // ---- layer 1: h = relu(c W^T + b) ----
    __syncthreads();
    mlp_load_weights(a.wk + size_t(2 * net) * L * L, L, Ws);
    __syncthreads();
    mlp_block_product(xt, Ws, L, r0, j0, acc);
    float* hg = net ? a.h2 : a.h1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = j0 + j;
      const float bj = col < L ? __ldg(a.b[2 * net] + col) : 0.f;
      float hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) hv[i] = fmaxf(acc[i][j] + bj, 0.f);
      if (col < L) {
        *reinterpret_cast<float4*>(ht + col * kMlpRows + r0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (row0 + r0 + i < a.N) hg[size_t(row0 + r0 + i) * L + col] = hv[i];
      }
    }
    // ---- layer 2: out = h W^T + b (tanh for the log-variance net) ----
    __syncthreads();
    mlp_load_weights(a.wk + size_t(2 * net + 1) * L * L, L, Ws);
    __syncthreads();
    mlp_block_product(ht, Ws, L, r0, j0, acc);
    float* og = net ? a.lv : a.mu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + r0 + i;
      if (row >= a.N) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + j;
        if (col >= L) continue;
        const float v = acc[i][j] + __ldg(a.b[2 * net + 1] + col);
        og[size_t(row) * L + col] = net ? tanhf(v) : v;
      }
    }
  }
}

// d(hidden) / d(pre) / d(input) of both nets for 32 rows; bias gradients; the weight gradients are GEMMs over the rows.
static __global__ void __launch_bounds__(256, 2) club_mlp_bwd_kernel(const ClubMlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                                     // [L][128]  weights as stored: Ws[j][k] = W[j][k]  (j = reduction index)
  float* gt = Ws + kMlpMaxL * kMlpMaxL;               // [L][32]  upstream tile, j-major
  float* colsum = gt + kMlpMaxL * kMlpRows;           // [4][128]
  const int L = a.L, row0 = blockIdx.x * kMlpRows;
  const int rg = threadIdx.x >> 5, cg = threadIdx.x & 31;
  const int r0 = rg * 4, k0 = cg * 4;
  for (int t = threadIdx.x; t < 4 * kMlpMaxL; t += blockDim.x) colsum[t] = 0.f;
  float dc_acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dc_acc[i][j] = 0.f;
  for (int net = 0; net < 2; ++net) {
    // ---- upstream of layer 2 (mu: dmu; lv: dlv (1 - lv^2)) -> gt[j][r], column sums = db of layer 2 ----
    __syncthreads();
    for (int t = threadIdx.x; t < L * kMlpRows; t += blockDim.x) {
      const int r = t % kMlpRows, j = t / kMlpRows;
      float g = 0.f;
      if (row0 + r < a.N) {
        const size_t o = size_t(row0 + r) * L + j;
        if (net == 0) {
          g = a.dmu ? a.dmu[o] : 0.f;
        } else {
          const float lv = a.lv[o];
          g = (a.dlv ? a.dlv[o] : 0.f) * (1.f - lv * lv);
          a.dpre[o] = g;
        }
      }
      gt[j * kMlpRows + r] = g;
    }
    mlp_load_weights(a.w[2 * net + 1], L, Ws);        // W[j][k]: dh[r][k] = sum_j g[r][j] W[j][k]
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < kMlpRows; ++r) s += gt[j * kMlpRows + r];
      colsum[(2 * net + 1) * kMlpMaxL + j] = s;
    }
    float acc[4][4];
    mlp_block_product(gt, Ws, L, r0, k0, acc);
    // ---- relu backward -> dh (global, for the layer-1 weight gradient) and as the next upstream tile ----
    const float* hg = net ? a.h2 : a.h1;
    float* dhg = net ? a.dh2 : a.dh1;
    float dh[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + r0 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = k0 + j;
        float v = 0.f;
        if (row < a.N && col < L) {
          v = hg[size_t(row) * L + col] > 0.f ? acc[i][j] : 0.f;
          dhg[size_t(row) * L + col] = v;
        }
        dh[i][j] = v;
      }
    }
    __syncthreads();                                   // everyone is done with gt / Ws
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k0 + j < L) *reinterpret_cast<float4*>(gt + (k0 + j) * kMlpRows + r0) = make_float4(dh[0][j], dh[1][j], dh[2][j], dh[3][j]);
    mlp_load_weights(a.w[2 * net], L, Ws);            // layer 1 weights: dc[r][k] += sum_j dh[r][j] W1[j][k]
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < kMlpRows; ++r) s += gt[j * kMlpRows + r];
      colsum[(2 * net) * kMlpMaxL + j] = s;
    }
    if (a.dc != nullptr) {
      mlp_block_product(gt, Ws, L, r0, k0, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dc_acc[i][j] += acc[i][j];
    }
  }
  if (a.dc != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + r0 + i;
      if (row >= a.N) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < L) a.dc[size_t(row) * a.lddc + k0 + j] = dc_acc[i][j];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * kMlpMaxL; t += blockDim.x) {
    const int m = t / kMlpMaxL, j = t % kMlpMaxL;
    if (j < L && a.db[m] != nullptr && colsum[t] != 0.f) atomicAdd(a.db[m] + j, colsum[t]);
  }
}

struct ClubMlpWs {
  float* wk;      // [4][L][L]
  float* h1; float* h2;
  float* dh1; float* dh2; float* dpre;
  size_t bytes;
};
static ClubMlpWs carve_club_mlp(void* ws, int64_t N, int64_t L) {
  WsCarver w(ws);
  ClubMlpWs r;
  r.wk = w.take<float>(size_t(4) * L * L);
  r.h1 = w.take<float>(size_t(N) * L);
  r.h2 = w.take<float>(size_t(N) * L);
  r.dh1 = w.take<float>(size_t(N) * L);
  r.dh2 = w.take<float>(size_t(N) * L);
  r.dpre = w.take<float>(size_t(N) * L);
  r.bytes = w.bytes();
  return r;
}
constexpr int kMlpFwdSmem = (kMlpMaxL * kMlpMaxL + 2 * kMlpMaxL * kMlpRows) * 4;
constexpr int kMlpBwdSmem = (kMlpMaxL * kMlpMaxL + kMlpMaxL * kMlpRows + 4 * kMlpMaxL) * 4;

}  // namespace carel

using namespace carel;

extern "C" size_t carel_club_mlp_workspace_bytes(int64_t N, int64_t L) { return carve_club_mlp(nullptr, N, L).bytes; }

// w8[8] = {W1, b1, W2, b2 (ec_mu.0, ec_mu.2), W3, b3, W4, b4 (ec_log_var.0, ec_log_var.2)}; mu, lv [N, L] contiguous.
extern "C" int carel_club_mlp_fwd(const float* c, int64_t ldc, int64_t N, int64_t L, const float* const* w8, float* mu, float* lv,
                                  void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && L > 0 && L <= kMlpMaxL && (L % 4) == 0, "carel_club_mlp: need ec_dim <= 128 and a multiple of 4");
  ClubMlpWs w = carve_club_mlp(ws, N, L);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_club_mlp_fwd: workspace too small");
  Mlp4 src;
  for (int i = 0; i < 4; ++i) src.w[i] = w8[2 * i];
  const int t32 = int((L + 31) / 32);
  club_mlp_transpose_kernel<<<dim3(t32, t32, 4), dim3(32, 8), 0, st>>>(src, int(L), w.wk);
  CAREL_LAUNCHED();
  ClubMlpArgs a{};
  a.c = c; a.ldc = ldc; a.wk = w.wk; a.N = int(N); a.L = int(L);
  for (int i = 0; i < 4; ++i) a.b[i] = w8[2 * i + 1];
  a.h1 = w.h1; a.h2 = w.h2; a.mu = mu; a.lv = lv;
  static const cudaError_t attr = cudaFuncSetAttribute(club_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpFwdSmem);
  CAREL_CHECK_CUDA(attr);
  club_mlp_fwd_kernel<<<int((N + kMlpRows - 1) / kMlpRows), 256, kMlpFwdSmem, st>>>(a);
  CAREL_LAUNCHED();
  return 0;
}

// dmu / dlv [N, L] (either may be NULL); dc [N, L] (pitch lddc) may be NULL (detached input);
// g8[8]: gradients of w8 (each may be NULL), ACCUMULATED: the caller zeroes them (one fill of one flat buffer).  `lv` = the forward's output; `ws` = the forward's workspace.
extern "C" int carel_club_mlp_bwd(const float* dmu, const float* dlv, const float* c, int64_t ldc, const float* lv, int64_t N,
                                  int64_t L, const float* const* w8, float* dc, int64_t lddc, float* const* g8, void* ws,
                                  size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ClubMlpWs w = carve_club_mlp(ws, N, L);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_club_mlp_bwd: workspace too small");
  ClubMlpArgs a{};
  a.c = c; a.ldc = ldc; a.N = int(N); a.L = int(L);
  for (int i = 0; i < 4; ++i) { a.w[i] = w8[2 * i]; a.b[i] = w8[2 * i + 1]; a.db[i] = g8[2 * i + 1]; }
  a.h1 = w.h1; a.h2 = w.h2; a.lv = const_cast<float*>(lv);
  a.dmu = dmu; a.dlv = dlv; a.dh1 = w.dh1; a.dh2 = w.dh2; a.dpre = w.dpre; a.dc = dc; a.lddc = lddc;
  static const cudaError_t attr = cudaFuncSetAttribute(club_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpBwdSmem);
  CAREL_CHECK_CUDA(attr);
  club_mlp_bwd_kernel<<<int((N + kMlpRows - 1) / kMlpRows), 256, kMlpBwdSmem, st>>>(a);
  CAREL_LAUNCHED();
  // weight gradients: dW[j][k] = sum_rows up[r][j] in[r][k]   ([L, N] x [N, L]): the four products in one launch,
  // the rows split over a quarter of the SMs each, accumulated into the caller-zeroed outputs
  SgemmBatch4 b{};
  const float* ups[4] = {w.dh1, dmu, w.dh2, w.dpre};
  const float* ins[4] = {c, w.h1, c, w.h2};
  const int64_t ldi[4] = {ldc, L, ldc, L};
  bool any = false;
  for (int i = 0; i < 4; ++i) {
    b.A[i] = ups[i]; b.lda[i] = L; b.B[i] = ins[i]; b.ldb[i] = ldi[i]; b.C[i] = g8[2 * i];
    any |= (g8[2 * i] != nullptr && ups[i] != nullptr);
  }
  if (any)
    if (int rc = launch_sgemm_batch4_tt(b, L, L, L, N, st)) return rc;
  return 0;
}
