// bow_f32.cu — fp32 precision tier of the BoW reconstruction decoder + loss (sgemm.cuh):
// drl_classifier_ec_mmd_final_mul.py:253-254 (decoder, Softmax(dim=1)) and :381-387 (label smoothing + BCELoss), evaluated
// like the reference's fp32 ATen path: the logits [N, V] are MATERIALISED by an fp32 GEMM, softmax / BCE and their
// gradients are computed per element from them (torch's rules: log clamped at -100, p (1 - p) clamped at 1e-12 in the
// BCE backward), and the two weight-gradient products are fp32 GEMMs again.  1e-5 parity tier; the product path is bow.cu.
#include <algorithm>

#include "sgemm.cuh"

namespace carel {

struct BowF32Ws {
  float* Y;        // [N, V]  logits (forward) -> dlogits (backward)
  float* T;        // [N, V]  dense target expanded from an ELL triple (only when no dense target is given)
  float* lse;      // [N]
  double* acc;     // [1]
  int* flags;      // [1] ELL overflow without a dense target
  size_t bytes;
};
static BowF32Ws carve_bow_f32(void* ws, int64_t N, int64_t V, int need_target) {
  WsCarver w(ws);
  BowF32Ws r;
  r.Y = w.take<float>(size_t(N) * V);
  r.T = need_target ? w.take<float>(size_t(N) * V) : nullptr;
  r.lse = w.take<float>(size_t(N));
  r.acc = w.take<double>(1);
  r.flags = w.take<int>(1);
  r.bytes = w.bytes();
  return r;
}

static __global__ void bow_f32_scatter_ell_kernel(const int* __restrict__ cols, const float* __restrict__ vals,
                                                  const int* __restrict__ nnz, int N, int V, int cap, float* __restrict__ T,
                                                  int* __restrict__ flags) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 1LL * N * cap; t += (long long)gridDim.x * blockDim.x) {
    const int r = int(t / cap), q = int(t % cap);
    const int n = nnz[r];
    if (n > cap && q == 0) *flags = 1;           // truncated list: reported as a NaN loss, like the product path
    if (q < min(n, cap)) {
      const int c = cols[t];
      if (c >= 0 && c < V) atomicAdd(T + int64_t(r) * V + c, vals[t]);
    }
  }
}

template <class T>
static __device__ __forceinline__ T block_sum_256(T v, T* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = sh[0];
  for (int i = 1; i < 8; ++i) r += sh[i];
  return r;
}
static __device__ __forceinline__ float block_max_256(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < 8; ++i) r = fmaxf(r, sh[i]);
  return r;
}

// One block per row.  BWD = false: lse_i and the row's loss terms.  BWD = true: dlogits over the logits.
template <bool BWD>
static __global__ void __launch_bounds__(256)
bow_f32_row_kernel(float* __restrict__ Y, const float* __restrict__ target, int64_t ldt, int N, int V, float ls,
                   float* __restrict__ lse_io, double* __restrict__ acc, const float* __restrict__ grad_out, float inv_nv) {
  __shared__ float shf[8];
  __shared__ double shd[8];
  const float c = ls / float(V), oml = 1.f - ls;
  double loss_blk = 0.0;
  for (int r = blockIdx.x; r < N; r += gridDim.x) {
    float* y = Y + int64_t(r) * V;
    const float* tr = target + int64_t(r) * ldt;
    if (!BWD) {
      float mx = -INFINITY;
      for (int j = threadIdx.x; j < V; j += 256) mx = fmaxf(mx, y[j]);
      mx = block_max_256(mx, shf);
      double s = 0.0;
      for (int j = threadIdx.x; j < V; j += 256) s += double(expf(y[j] - mx));
      s = block_sum_256<double>(s, shd);
      const float lse = mx + float(log(s));
      double f = 0.0;
      for (int j = threadIdx.x; j < V; j += 256) {
        const float p = expf(y[j] - lse);
        const float t = tr[j] * oml + c;
        f += double(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(log1pf(-p), -100.f));
      }
      f = block_sum_256<double>(f, shd);
      if (threadIdx.x == 0) { lse_io[r] = lse; loss_blk += f; }
    } else {
      const float lse = lse_io[r];
      double sd = 0.0;                       // s_i = sum_j g_ij p_ij,  g = (p - t) / max(p (1 - p), 1e-12)
      for (int j = threadIdx.x; j < V; j += 256) {
        const float p = expf(y[j] - lse);
        const float t = tr[j] * oml + c;
        sd += double((p - t) / fmaxf(p * (1.f - p), 1e-12f) * p);
      }
      const float si = float(block_sum_256<double>(sd, shd));
      const float go = __ldg(grad_out) * inv_nv;
      for (int j = threadIdx.x; j < V; j += 256) {
        const float p = expf(y[j] - lse);
        const float t = tr[j] * oml + c;
        const float g = (p - t) / fmaxf(p * (1.f - p), 1e-12f);
        y[j] = p * (g - si) * go;
      }
    }
  }
  if (!BWD && threadIdx.x == 0 && loss_blk != 0.0) atomicAdd(acc, loss_blk);
}

static __global__ void bow_f32_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ flags, double inv_nv,
                                               float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = flags[0] ? __int_as_float(0x7fc00000) : float(-acc[0] * inv_nv);
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_bow_f32_workspace_bytes(int64_t N, int64_t K, int64_t V, int need_dense_target) {
  (void)K;
  return carve_bow_f32(nullptr, N, V, need_dense_target).bytes;
}

extern "C" int carel_bow_f32_fwd(const float* z, int64_t ldz, const float* dec_w, const float* dec_b, const float* bow_dense,
                                 int64_t ld_bow, const int32_t* ell_cols, const float* ell_vals, const int32_t* ell_nnz,
                                 int64_t N, int64_t K, int64_t V, int64_t ell_cap, float label_smoothing, float* loss_out,
                                 void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && K > 0 && V > 0, "carel_bow_f32_fwd: empty problem");
  CAREL_REQUIRE(bow_dense != nullptr || (ell_cols && ell_vals && ell_nnz && ell_cap > 0), "carel_bow_f32_fwd: need a dense or an ELL target");
  BowF32Ws w = carve_bow_f32(ws, N, V, bow_dense == nullptr);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_f32_fwd: workspace too small");
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double), st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(int), st));
  if (int rc = launch_sgemm(false, false, z, ldz, dec_w, K, w.Y, V, N, V, K, 1.f, dec_b, 0, 1, st)) return rc;
  const float* target = bow_dense;
  int64_t ldt = ld_bow;
  if (bow_dense == nullptr) {
    CAREL_CHECK_CUDA(cudaMemsetAsync(w.T, 0, size_t(N) * V * sizeof(float), st));
    bow_f32_scatter_ell_kernel<<<int(std::min<int64_t>((N * ell_cap + 255) / 256, 148 * 16)), 256, 0, st>>>(
        ell_cols, ell_vals, ell_nnz, int(N), int(V), int(ell_cap), w.T, w.flags);
    CAREL_LAUNCHED();
    target = w.T;
    ldt = V;
  }
  bow_f32_row_kernel<false><<<int(std::min<int64_t>(N, 148 * 8)), 256, 0, st>>>(w.Y, target, ldt, int(N), int(V), label_smoothing,
                                                                        w.lse, w.acc, nullptr, 0.f);
  CAREL_LAUNCHED();
  bow_f32_finalize_kernel<<<1, 32, 0, st>>>(w.acc, w.flags, 1.0 / (double(N) * double(V)), loss_out);
  CAREL_LAUNCHED();
  return 0;
}

// Must follow carel_bow_f32_fwd on the same workspace (logits, lse and the expanded target are read from it).
extern "C" int carel_bow_f32_bwd(const float* grad_out, const float* z, int64_t ldz, const float* dec_w, const float* bow_dense,
                                 int64_t ld_bow, int64_t N, int64_t K, int64_t V, float label_smoothing, float* grad_z,
                                 int64_t ldgz, float* grad_w, float* grad_b, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BowF32Ws w = carve_bow_f32(ws, N, V, bow_dense == nullptr);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_f32_bwd: workspace too small");
  const float* target = bow_dense ? bow_dense : w.T;
  const int64_t ldt = bow_dense ? ld_bow : V;
  bow_f32_row_kernel<true><<<int(std::min<int64_t>(N, 148 * 8)), 256, 0, st>>>(w.Y, target, ldt, int(N), int(V), label_smoothing, w.lse,
                                                                       nullptr, grad_out, float(1.0 / (double(N) * double(V))));
  CAREL_LAUNCHED();
  if (grad_z != nullptr)       // dZ = dlogit W_d
    if (int rc = launch_sgemm(false, true, w.Y, V, dec_w, K, grad_z, ldgz, N, K, V, 1.f, nullptr, 0, 1, st)) return rc;
  if (grad_w != nullptr)       // dW = dlogit^T Z
    if (int rc = launch_sgemm(true, true, w.Y, V, z, ldz, grad_w, K, V, K, N, 1.f, nullptr, 0, 1, st)) return rc;
  if (grad_b != nullptr) {
    CAREL_CHECK_CUDA(cudaMemsetAsync(grad_b, 0, size_t(V) * sizeof(float), st));
    if (int rc = launch_f32_colsum(w.Y, V, N, V, 1.f, grad_b, st)) return rc;
  }
  return 0;
}

// Debug / test entry of the fp32-tier GEMM itself.
extern "C" int carel_f32_gemm(const float* a, int64_t lda, int ta, const float* b, int64_t ldb, int tb, int64_t M, int64_t N,
                              int64_t K, float* c, int64_t ldc, const float* bias, float alpha, int mode, int ksplit,
                              void* stream) {
  return launch_sgemm(ta != 0, tb != 0, a, lda, b, ldb, c, ldc, M, N, K, alpha, bias, mode, ksplit, static_cast<cudaStream_t>(stream));
}
