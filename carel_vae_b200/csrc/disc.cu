// disc.cu — carel_disc_{fwd,bwd}: the two adversarial discriminators of the GAN trainer in one kernel each
// (drl_classifier_ec_gan.py: get_ec_disc_preds / get_ce_disc_preds :430-447, get_ec_disc_loss :449-463,
//  get_entropy_loss :465-477, call sites in forward :222-233):
//
//   p_ec = sigmoid(w_ec . (z_c (.) m_ec) + b_ec),   p_ce = sigmoid(w_ce . (z_e (.) m_ce) + b_ce)       (z DETACHED)
//   ec_disc_loss = BCELoss(p_ec, y_emo (1 - ls) + ls / n_class),   ce_disc_loss = BCELoss(p_ce, y_cau (1 - ls) + ls / n_class)
//   entropy      = mean_i p_ec log(p_ec + eps) + mean_i p_ce log(p_ce + eps)
// One warp per clause row (two dot products of length L from registers), block-level double sums; the backward pass
// produces the gradients of the four discriminator parameters for any combination of the three upstream gradients
// (the trainer calls backward three times: gan:790-798).  The reference runs ~30 launch-bound ATen kernels for this.
#include <algorithm>

#include "api_common.cuh"

namespace carel {

struct DiscArgs {
  const float* z; int64_t ldz;            // [N, 2L] = [z_e | z_c]
  const float* m_ec; const float* m_ce;   // [N, L] scaled dropout masks or NULL
  const float* w_ec; const float* b_ec;   // [L], [1]   reads z_c
  const float* w_ce; const float* b_ce;   // [L], [1]   reads z_e
  const float* y_emo; const float* y_cau; // [N]
  int N, L;
  float ls, inv_nc, eps;
  float* probs;                           // [N, 2] saved for backward
  double* sums;                           // [4] bce_ec, bce_ce, ent_ec, ent_ce
  // backward
  const float* g_ec; const float* g_ce; const float* g_ent;     // device scalars or NULL (= 0)
  float* dw_ec; float* db_ec; float* dw_ce; float* db_ce;       // accumulated (zeroed by the host side)
};

static __global__ void __launch_bounds__(256) disc_fwd_kernel(const DiscArgs a) {
  __shared__ double sh[8][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = blockIdx.x * wpb + wid; r < a.N; r += gridDim.x * wpb) {
    const float* zr = a.z + size_t(r) * a.ldz;
    float x_ec = 0.f, x_ce = 0.f;
    for (int k = lane; k < a.L; k += 32) {
      const float zc = zr[a.L + k] * (a.m_ec ? a.m_ec[size_t(r) * a.L + k] : 1.f);
      const float ze = zr[k] * (a.m_ce ? a.m_ce[size_t(r) * a.L + k] : 1.f);
      x_ec = fmaf(zc, __ldg(a.w_ec + k), x_ec);
      x_ce = fmaf(ze, __ldg(a.w_ce + k), x_ce);
    }
    x_ec = warp_sum(x_ec) + __ldg(a.b_ec);
    x_ce = warp_sum(x_ce) + __ldg(a.b_ce);
    if (lane == 0) {
      const float p1 = 1.f / (1.f + expf(-x_ec)), p2 = 1.f / (1.f + expf(-x_ce));
      a.probs[2 * r] = p1;
      a.probs[2 * r + 1] = p2;
      const float t1 = a.y_emo[r] * (1.f - a.ls) + a.ls * a.inv_nc, t2 = a.y_cau[r] * (1.f - a.ls) + a.ls * a.inv_nc;
      // torch.nn.BCELoss clamps the logs at -100
      s[0] -= double(t1 * fmaxf(logf(p1), -100.f) + (1.f - t1) * fmaxf(log1pf(-p1), -100.f));
      s[1] -= double(t2 * fmaxf(logf(p2), -100.f) + (1.f - t2) * fmaxf(log1pf(-p2), -100.f));
      s[2] += double(p1 * logf(p1 + a.eps));
      s[3] += double(p2 * logf(p2 + a.eps));
    }
  }
  if (lane == 0)
    for (int i = 0; i < 4; ++i) sh[wid][i] = s[i];
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh[w][threadIdx.x];
    if (t != 0.0) atomicAdd(a.sums + threadIdx.x, t);
  }
}

// out[0] = ec_disc_loss, out[1] = ce_disc_loss, out[2] = entropy sum
static __global__ void disc_finalize_kernel(const double* __restrict__ sums, int N, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = float(sums[0] / N);
    out[1] = float(sums[1] / N);
    out[2] = float(sums[2] / N + sums[3] / N);
  }
}

static __global__ void __launch_bounds__(256) disc_bwd_kernel(const DiscArgs a) {
  __shared__ float sh_w[2][128 + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float inv_n = 1.f / float(a.N);
  const float g_ec = a.g_ec ? __ldg(a.g_ec) : 0.f, g_ce = a.g_ce ? __ldg(a.g_ce) : 0.f, g_ent = a.g_ent ? __ldg(a.g_ent) : 0.f;
  for (int i = threadIdx.x; i < 2 * 129; i += blockDim.x) (&sh_w[0][0])[i] = 0.f;
  __syncthreads();
  float acc_ec[4] = {0.f, 0.f, 0.f, 0.f}, acc_ce[4] = {0.f, 0.f, 0.f, 0.f};     // this lane's dims k = lane + 32 q
  float acc_b1 = 0.f, acc_b2 = 0.f;
  for (int r = blockIdx.x * wpb + wid; r < a.N; r += gridDim.x * wpb) {
    const float p1 = a.probs[2 * r], p2 = a.probs[2 * r + 1];
    const float t1 = a.y_emo[r] * (1.f - a.ls) + a.ls * a.inv_nc, t2 = a.y_cau[r] * (1.f - a.ls) + a.ls * a.inv_nc;
    // d/dx: BCELoss backward (p - t) / max(p (1 - p), 1e-12), sigmoid backward p (1 - p); entropy term
    const float sp1 = p1 * (1.f - p1), sp2 = p2 * (1.f - p2);
    const float dx1 = inv_n * sp1 * (g_ec * (p1 - t1) / fmaxf(sp1, 1e-12f) + g_ent * (logf(p1 + a.eps) + p1 / (p1 + a.eps)));
    const float dx2 = inv_n * sp2 * (g_ce * (p2 - t2) / fmaxf(sp2, 1e-12f) + g_ent * (logf(p2 + a.eps) + p2 / (p2 + a.eps)));
    const float* zr = a.z + size_t(r) * a.ldz;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = lane + 32 * q;
      if (k < a.L) {
        acc_ec[q] = fmaf(dx1, zr[a.L + k] * (a.m_ec ? a.m_ec[size_t(r) * a.L + k] : 1.f), acc_ec[q]);
        acc_ce[q] = fmaf(dx2, zr[k] * (a.m_ce ? a.m_ce[size_t(r) * a.L + k] : 1.f), acc_ce[q]);
      }
    }
    acc_b1 += dx1;
    acc_b2 += dx2;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = lane + 32 * q;
    if (k < a.L) {
      atomicAdd(&sh_w[0][k], acc_ec[q]);
      atomicAdd(&sh_w[1][k], acc_ce[q]);
    }
  }
  if (lane == 0) {
    atomicAdd(&sh_w[0][128], acc_b1);
    atomicAdd(&sh_w[1][128], acc_b2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.L; i += blockDim.x) {
    if (sh_w[0][i] != 0.f) atomicAdd(a.dw_ec + i, sh_w[0][i]);
    if (sh_w[1][i] != 0.f) atomicAdd(a.dw_ce + i, sh_w[1][i]);
  }
  if (threadIdx.x == 0) {
    if (sh_w[0][128] != 0.f) atomicAdd(a.db_ec, sh_w[0][128]);
    if (sh_w[1][128] != 0.f) atomicAdd(a.db_ce, sh_w[1][128]);
  }
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_disc_workspace_bytes(int64_t N) {
  return align_up(size_t(N) * 2 * sizeof(float), 256) + 256;
}

static int disc_args(DiscArgs& a, const float* z, int64_t ldz, int64_t N, int64_t L, const float* const* masks,
                     const float* const* params, const float* y_emo, const float* y_cau, float ls, float n_class, float eps,
                     void* ws, size_t ws_bytes) {
  CAREL_REQUIRE(N > 0 && L > 0 && L <= 128, "carel_disc: need 1 <= ec_dim <= 128");
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_disc_workspace_bytes(N), "carel_disc: workspace too small");
  a.z = z; a.ldz = ldz; a.m_ec = masks[0]; a.m_ce = masks[1];
  a.w_ec = params[0]; a.b_ec = params[1]; a.w_ce = params[2]; a.b_ce = params[3];
  a.y_emo = y_emo; a.y_cau = y_cau; a.N = int(N); a.L = int(L);
  a.ls = ls; a.inv_nc = 1.f / n_class; a.eps = eps;
  a.probs = static_cast<float*>(ws);
  a.sums = reinterpret_cast<double*>(static_cast<uint8_t*>(ws) + align_up(size_t(N) * 2 * sizeof(float), 256));
  return 0;
}

// masks[2] = {m_ec, m_ce} ([N, L] or NULL); params[4] = {ec_disc.weight [L], ec_disc.bias [1], ce_disc.weight, ce_disc.bias};
// out[3] = {ec_disc_loss, ce_disc_loss, ec_entropy + ce_entropy}
extern "C" int carel_disc_fwd(const float* z, int64_t ldz, int64_t N, int64_t L, const float* const* masks,
                              const float* const* params, const float* y_emo, const float* y_cau, float label_smoothing,
                              float n_class, float eps, float* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DiscArgs a{};
  if (int rc = disc_args(a, z, ldz, N, L, masks, params, y_emo, y_cau, label_smoothing, n_class, eps, ws, ws_bytes)) return rc;
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.sums, 0, 4 * sizeof(double), st));
  disc_fwd_kernel<<<int(std::min<int64_t>((N + 7) / 8, sm_count() * 8)), 256, 0, st>>>(a);
  CAREL_LAUNCHED();
  disc_finalize_kernel<<<1, 32, 0, st>>>(a.sums, int(N), out);
  CAREL_LAUNCHED();
  return 0;
}

// g[3]: device scalars (upstream gradients of out[0..2]) or NULL; grads[4] = {dW_ec [L], db_ec [1], dW_ce [L], db_ce [1]},
// fully written (zeroed here, then accumulated).
extern "C" int carel_disc_bwd(const float* const* g, const float* z, int64_t ldz, int64_t N, int64_t L,
                              const float* const* masks, const float* const* params, const float* y_emo, const float* y_cau,
                              float label_smoothing, float n_class, float eps, float* const* grads, void* ws, size_t ws_bytes,
                              void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DiscArgs a{};
  if (int rc = disc_args(a, z, ldz, N, L, masks, params, y_emo, y_cau, label_smoothing, n_class, eps, ws, ws_bytes)) return rc;
  a.g_ec = g[0]; a.g_ce = g[1]; a.g_ent = g[2];
  a.dw_ec = grads[0]; a.db_ec = grads[1]; a.dw_ce = grads[2]; a.db_ce = grads[3];
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.dw_ec, 0, size_t(L) * 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.db_ec, 0, 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.dw_ce, 0, size_t(L) * 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(a.db_ce, 0, 4, st));
  disc_bwd_kernel<<<int(std::min<int64_t>((N + 7) / 8, sm_count() * 4)), 256, 0, st>>>(a);
  CAREL_LAUNCHED();
  return 0;
}
