// club.cu — carel_club_{fwd,bwd}: the CLUB terms of the VI variant after the two small MLPs
// (drl_classifier_ec_vi_final.py get_ec_aprx_loss :421-426, get_ec_upper_loss :428-439):
//
//   aprx  (perm == NULL):  mean_i sum_d [ (mu_id - e_id)^2 / exp(lv_id) + lv_id ]                    ( = -log-likelihood )
//   upper (perm != NULL):  mean_i sum_d [ (mu_id - e_pi(i)d)^2 - (mu_id - e_id)^2 ] / exp(lv_id) / 2
//
// mu, lv [N, L] = ec_mu(z_c), ec_log_var(z_c); e = z_e (a column view of z: row pitch lde); pi = the reference's
// torch.randperm(N).  The reference runs ~20 elementwise ATen kernels forward and ~25 backward over [N, L] for the two
// terms; here each term is one HBM-bound pass (one warp per row, coalesced 128-byte rows, warp-shuffle + one double
// atomic per block) forward and one backward.  pi is a permutation, so the scatter of d/de_pi(i) is conflict-free.
#include <algorithm>

#include "api_common.cuh"

namespace carel {

template <bool UPPER>
static __global__ void __launch_bounds__(256)
club_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ e, int64_t lde,
                const int64_t* __restrict__ perm, int N, int L, double* __restrict__ acc) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double tot = 0.0;
  for (int r = blockIdx.x * wpb + warp; r < N; r += gridDim.x * wpb) {
    const float* mr = mu + size_t(r) * L;
    const float* lr = lv + size_t(r) * L;
    const float* er = e + size_t(r) * lde;
    const float* ep = UPPER ? e + size_t(perm[r]) * lde : nullptr;
    float s = 0.f;
    for (int d = lane; d < L; d += 32) {
      const float m = mr[d], l = lr[d], den = expf(l);
      const float dp = m - er[d];
      if (UPPER) {
        const float dn = m - ep[d];
        s += (dn - dp) * (dn + dp) / den;          // = (dn^2 - dp^2) / den without the cancellation
      } else {
        s += dp * dp / den + l;
      }
    }
    tot += double(warp_sum(s));
  }
  if (lane == 0) sh[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh[w];
    if (t != 0.0) atomicAdd(acc, t);
  }
}

static __global__ void club_finalize_kernel(const double* __restrict__ acc, double scale, float* __restrict__ out) {
  if (threadIdx.x == 0) out[0] = float(acc[0] * scale);
}

template <bool UPPER>
static __global__ void __launch_bounds__(256)
club_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ mu, const float* __restrict__ lv,
                const float* __restrict__ e, int64_t lde, const int64_t* __restrict__ perm, int N, int L, float c,
                float* __restrict__ g_mu, float* __restrict__ g_lv, float* __restrict__ g_e) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float g = __ldg(grad_out) * c;                      // c = 1/N (aprx) or 1/(2N) (upper)
  for (int r = blockIdx.x * wpb + warp; r < N; r += gridDim.x * wpb) {
    const size_t o = size_t(r) * L;
    const float* er = e + size_t(r) * lde;
    const int64_t pr = UPPER ? perm[r] : 0;
    const float* ep = UPPER ? e + size_t(pr) * lde : nullptr;
    for (int d = lane; d < L; d += 32) {
      const float m = mu[o + d], inv = 1.f / expf(lv[o + d]);
      const float dp = m - er[d];
      if (UPPER) {
        const float dn = m - ep[d];
        const float w = inv * g, w2 = 2.f * w;         // shared factors: a fixed point pi(i) = i cancels exactly
        if (g_mu) g_mu[o + d] = w2 * (dn - dp);
        if (g_lv) g_lv[o + d] = -w * ((dn - dp) * (dn + dp));
        if (g_e) {
          atomicAdd(g_e + o + d, w2 * dp);
          atomicAdd(g_e + size_t(pr) * L + d, -(w2 * dn));
        }
      } else {
        const float t = 2.f * inv * dp * g;
        if (g_mu) g_mu[o + d] = t;
        if (g_lv) g_lv[o + d] = (1.f - inv * dp * dp) * g;
        if (g_e) g_e[o + d] = -t;
      }
    }
  }
}

}  // namespace carel

using namespace carel;

extern "C" int carel_club_fwd(const float* mu_hat, const float* lv_hat, const float* e, int64_t lde,
                              const int64_t* perm, int64_t N, int64_t L, float* out, double* acc, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && L > 0 && lde >= L, "carel_club_fwd: bad shape (N %lld, L %lld, lde %lld)", (long long)N,
                (long long)L, (long long)lde);
  CAREL_REQUIRE(mu_hat && lv_hat && e && out && acc, "carel_club_fwd: null argument");
  CAREL_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const int blocks = int(std::min<int64_t>((N + 7) / 8, 148 * 8));
  if (perm != nullptr)
    club_fwd_kernel<true><<<blocks, 256, 0, st>>>(mu_hat, lv_hat, e, lde, perm, int(N), int(L), acc);
  else
    club_fwd_kernel<false><<<blocks, 256, 0, st>>>(mu_hat, lv_hat, e, lde, perm, int(N), int(L), acc);
  CAREL_LAUNCHED();
  club_finalize_kernel<<<1, 32, 0, st>>>(acc, (perm != nullptr ? 0.5 : 1.0) / double(N), out);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_club_bwd(const float* grad_out, const float* mu_hat, const float* lv_hat, const float* e,
                              int64_t lde, const int64_t* perm, int64_t N, int64_t L, float* grad_mu,
                              float* grad_lv, float* grad_e, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && L > 0 && lde >= L, "carel_club_bwd: bad shape");
  CAREL_REQUIRE(grad_out && mu_hat && lv_hat && e, "carel_club_bwd: null argument");
  const int blocks = int(std::min<int64_t>((N + 7) / 8, 148 * 8));
  if (perm != nullptr) {
    if (grad_e) CAREL_CHECK_CUDA(cudaMemsetAsync(grad_e, 0, size_t(N) * L * sizeof(float), st));
    club_bwd_kernel<true><<<blocks, 256, 0, st>>>(grad_out, mu_hat, lv_hat, e, lde, perm, int(N), int(L),
                                                  float(0.5 / double(N)), grad_mu, grad_lv, grad_e);
  } else {
    club_bwd_kernel<false><<<blocks, 256, 0, st>>>(grad_out, mu_hat, lv_hat, e, lde, perm, int(N), int(L),
                                                   float(1.0 / double(N)), grad_mu, grad_lv, grad_e);
  }
  CAREL_LAUNCHED();
  return 0;
}
