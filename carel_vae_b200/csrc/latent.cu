// latent.cu — carel_latent_{fwd,bwd}: the VAE latent layer of DrlClassifier
// (drl_classifier_ec_mmd_final_mul.py: get_emotion_emb / get_cause_emb :312-336, sample_prior
// :345-351, get_kl_loss :525-534).
//
//   pre = X [N,D] * W4^T + b4,   W4 = [emotion_mu; emotion_log_var; cause_mu; cause_log_var]  ([4L, D])
//   z_s = mu_s + eps_s (shared by all rows) * exp(log_var_s)                 s in {emotion, cause}
//   kl_s = mean_i( -1/2 sum_d (1 + lv - exp(lv) - mu^2) )
//
// The projection is ONE tcgen05 GEMM (four nn.Linear fused); reparameterisation + KL are one
// HBM-bound pass with 128-bit accesses and warp-shuffle reductions.  Backward: one elementwise
// pass builds d(pre), then dX = dpre * W4 and [dW4 | db4] = dpre^T * [X | 1] on the tensor cores.
#include <algorithm>

#include "api_common.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "pack.cuh"
#include "sgemm.cuh"

extern "C" size_t carel_latent_f32_workspace_bytes(int64_t N, int64_t D, int64_t L);

namespace carel {

struct LatentWs {
  __nv_bfloat16* Xp;    // [N, ldx]   X | 1 (ones column at D) | 0
  __nv_bfloat16* Wp;    // [2 ldd, ldw]  W4 stacked twice (rows [0,4L) and [ldd, ldd+4L)); forward uses the first copy
  __nv_bfloat16* dpre;  // [N, 2 ldd]    bf16 d(pre) as hi | lo (two-term split keeps the backward GEMMs ~fp32-accurate)
  float* dW;            // [2 ldd, ldx]  fp32 accumulation target of [hi | lo]^T [X | 1]
  double* klacc;        // [2]
  size_t bytes;
  int ldx, ldw, ldd;
};

static LatentWs carve_latent(void* ws, int64_t N, int64_t D, int64_t L) {
  WsCarver w(ws);
  LatentWs r;
  r.ldx = int(round_up64(D + 1, 16));
  r.ldw = int(round_up64(D, 16));
  r.ldd = int(round_up64(4 * L, 16));
  r.Xp = w.take<__nv_bfloat16>(size_t(N) * r.ldx);
  r.Wp = w.take<__nv_bfloat16>(size_t(2 * r.ldd) * r.ldw);
  r.dpre = w.take<__nv_bfloat16>(size_t(N) * 2 * r.ldd);
  r.dW = w.take<float>(size_t(2 * r.ldd) * r.ldx);
  r.klacc = w.take<double>(2);
  r.bytes = w.bytes();
  return r;
}

// X -> bf16 with a ones column at index D (bias gradients come out of the dW GEMM for free).
static __global__ void pack_x_ones_kernel(const float* __restrict__ x, int N, int D, int ld,
                                          __nv_bfloat16* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    const float* s = x + size_t(r) * ld;
    __nv_bfloat16* o = out + size_t(r) * ldo;
    if (((D & 3) == 0) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
      for (int k = lane * 4; k < D; k += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + k));
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(o + k) = pk;
      }
    } else {
      for (int k = lane; k < D; k += 32) o[k] = __float2bfloat16_rn(__ldg(s + k));
    }
    for (int k = D + lane; k < ldo; k += 32) o[k] = __float2bfloat16_rn(k == D ? 1.f : 0.f);
  }
}

// The four [L, D] projection weights -> bf16 rows of Wp, stacked twice (rows [0, 4L) and [ldd, ldd + 4L)), and
// the four [L] biases -> one contiguous [4L] vector: one launch instead of eight packs and four copies.
struct PackW4Args {
  const float* w[4];
  const float* b[4];
};
static __global__ void pack_w4_kernel(const PackW4Args a, int L, int D, __nv_bfloat16* __restrict__ Wp, int ldw,
                                      int ldd, float* __restrict__ bias4) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < 4 * L; r += gridDim.x * wpb) {
    const int i = r / L, rr = r - i * L;
    const float* s = a.w[i] + size_t(rr) * D;
    __nv_bfloat16* o0 = Wp + size_t(r) * ldw;
    __nv_bfloat16* o1 = Wp + size_t(ldd + r) * ldw;
    if (((D & 3) == 0) && ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.w[i]) & 15) == 0)) {
      for (int k = lane * 4; k < ldw; k += 128) {          // 128-bit loads, 64-bit stores
        const float4 v = (k < D) ? __ldg(reinterpret_cast<const float4*>(s + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(o0 + k) = pk;
        *reinterpret_cast<uint2*>(o1 + k) = pk;
      }
    } else {
      for (int k = lane; k < ldw; k += 32) {
        const __nv_bfloat16 v = __float2bfloat16_rn(k < D ? __ldg(s + k) : 0.f);
        o0[k] = v;
        o1[k] = v;
      }
    }
    if (lane == 0) bias4[r] = __ldg(a.b[i] + rr);
  }
}

// One thread per 4 latent dims of one row and one space.  pre layout: [mu_e | lv_e | mu_c | lv_c].
// Writes z = [z_e | z_c] and accumulates the two KL sums (double) with one atomic per warp.
static __global__ void reparam_kl_fwd_kernel(const float* __restrict__ pre, int ldp,
                                             const float* __restrict__ eps_e,
                                             const float* __restrict__ eps_c, float* __restrict__ z,
                                             int ldz, int N, int L, double* __restrict__ klacc) {
  const int L4 = (L + 3) / 4;
  const long long total = 1LL * N * 2 * L4;
  float kl_local[2] = {0.f, 0.f};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = int(t % L4);
    const int sp = int((t / L4) % 2);
    const int row = int(t / (2LL * L4));
    const float* mu = pre + size_t(row) * ldp + sp * 2 * L + c4 * 4;
    const float* lv = mu + L;
    const float* eps = (sp ? eps_c : eps_e) + c4 * 4;
    float* zo = z + size_t(row) * ldz + sp * L + c4 * 4;
    float s = 0.f;
    if ((L & 3) == 0 && (ldp & 3) == 0 && (ldz & 3) == 0) {
      const float4 m4 = *reinterpret_cast<const float4*>(mu);
      const float4 l4 = *reinterpret_cast<const float4*>(lv);
      const float4 e4 = __ldg(reinterpret_cast<const float4*>(eps));
      const float ex = __expf(l4.x), ey = __expf(l4.y), ez = __expf(l4.z), ew = __expf(l4.w);
      *reinterpret_cast<float4*>(zo) =
          make_float4(fmaf(e4.x, ex, m4.x), fmaf(e4.y, ey, m4.y), fmaf(e4.z, ez, m4.z), fmaf(e4.w, ew, m4.w));
      s = (1.f + l4.x - ex - m4.x * m4.x) + (1.f + l4.y - ey - m4.y * m4.y) +
          (1.f + l4.z - ez - m4.z * m4.z) + (1.f + l4.w - ew - m4.w * m4.w);
    } else {
      for (int i = 0; i < 4 && c4 * 4 + i < L; ++i) {
        const float e = __expf(lv[i]);
        zo[i] = fmaf(__ldg(eps + i), e, mu[i]);
        s += 1.f + lv[i] - e - mu[i] * mu[i];
      }
    }
    kl_local[sp] += s;
  }
  // warp shuffle, then one shared-memory hop per block: two double atomics per BLOCK (per warp they serialise on
  // the two addresses and cost more than the pass itself)
  __shared__ float sh_kl[2][32];
  const float k0 = warp_sum(kl_local[0]), k1 = warp_sum(kl_local[1]);
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh_kl[0][wid] = k0; sh_kl[1][wid] = k1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += double(sh_kl[threadIdx.x][i]);
    if (t != 0.0) atomicAdd(klacc + threadIdx.x, t);
  }
}

static __global__ void kl_finalize_kernel(const double* __restrict__ klacc, int N, float* __restrict__ kl) {
  if (threadIdx.x < 2) kl[threadIdx.x] = float(-0.5 * klacc[threadIdx.x] / double(N));
}

// d(pre) from d(z) and the upstream gradients of the two KL scalars:
//   dmu = dz + g_kl * mu / N ;  dlv = dz * eps * exp(lv) - g_kl * (1 - exp(lv)) / (2N)
static __global__ void reparam_kl_bwd_kernel(const float* __restrict__ pre, int ldp,
                                             const float* __restrict__ eps_e,
                                             const float* __restrict__ eps_c,
                                             const float* __restrict__ gz, int ldgz,
                                             const float* __restrict__ gkl, int N, int L,
                                             __nv_bfloat16* __restrict__ dpre, int ldd) {
  const float invN = 1.f / float(N);
  const float g0 = gkl ? __ldg(gkl) : 0.f, g1 = gkl ? __ldg(gkl + 1) : 0.f;
  // one element: gradient pair + its two-term bf16 split
  auto one = [&](float mu, float lv, float eps, float dz, float g, float& dmu, float& dlv) {
    const float e = __expf(lv);
    dmu = dz + g * mu * invN;
    dlv = dz * eps * e - 0.5f * g * (1.f - e) * invN;
  };
  const bool vec = ((L & 3) == 0) && ((ldp & 3) == 0) && ((ldd & 3) == 0) && (gz == nullptr || (ldgz & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(pre) & 15) == 0) && ((reinterpret_cast<uintptr_t>(gz) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(eps_e) & 15) == 0) && ((reinterpret_cast<uintptr_t>(eps_c) & 15) == 0);
  if (vec) {
    // four latent dims per thread: 128-bit loads, 64-bit bf16x4 stores, 32-bit index arithmetic
    const int L4 = L >> 2;
    const unsigned total = unsigned(N) * 2u * unsigned(L4);       // < 2^32 for every shape the GEMMs accept
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      const unsigned c4 = t % unsigned(L4), rs = t / unsigned(L4);
      const int sp = int(rs & 1u), row = int(rs >> 1), c = int(c4) * 4;
      const float* base = pre + size_t(row) * ldp + sp * 2 * L + c;
      const float4 m4 = *reinterpret_cast<const float4*>(base);
      const float4 l4 = *reinterpret_cast<const float4*>(base + L);
      const float4 e4 = __ldg(reinterpret_cast<const float4*>((sp ? eps_c : eps_e) + c));
      const float4 d4 = gz ? *reinterpret_cast<const float4*>(gz + size_t(row) * ldgz + sp * L + c)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      const float g = sp ? g1 : g0;
      float dm[4], dl[4];
      one(m4.x, l4.x, e4.x, d4.x, g, dm[0], dl[0]);
      one(m4.y, l4.y, e4.y, d4.y, g, dm[1], dl[1]);
      one(m4.z, l4.z, e4.z, d4.z, g, dm[2], dl[2]);
      one(m4.w, l4.w, e4.w, d4.w, g, dm[3], dl[3]);
      auto split4 = [](const float (&v)[4], uint2& hi, uint2& lo) {
        __nv_bfloat16 h[4];
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { h[i] = __float2bfloat16_rn(v[i]); r[i] = v[i] - __bfloat162float(h[i]); }
        hi.x = pack_bf16x2(__bfloat162float(h[0]), __bfloat162float(h[1]));
        hi.y = pack_bf16x2(__bfloat162float(h[2]), __bfloat162float(h[3]));
        lo.x = pack_bf16x2(r[0], r[1]);
        lo.y = pack_bf16x2(r[2], r[3]);
      };
      uint2 mh, ml, lh, ll;
      split4(dm, mh, ml);
      split4(dl, lh, ll);
      __nv_bfloat16* o = dpre + size_t(row) * 2 * ldd + sp * 2 * L + c;
      *reinterpret_cast<uint2*>(o) = mh;
      *reinterpret_cast<uint2*>(o + L) = lh;
      *reinterpret_cast<uint2*>(o + ldd) = ml;
      *reinterpret_cast<uint2*>(o + ldd + L) = ll;
    }
    return;
  }
  const long long total = 1LL * N * 2 * L;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = int(t % L);
    const int sp = int((t / L) % 2);
    const int row = int(t / (2LL * L));
    const float mu = pre[size_t(row) * ldp + sp * 2 * L + c];
    const float lv = pre[size_t(row) * ldp + sp * 2 * L + L + c];
    const float eps = __ldg((sp ? eps_c : eps_e) + c);
    const float dz = gz ? gz[size_t(row) * ldgz + sp * L + c] : 0.f;
    float dmu, dlv;
    one(mu, lv, eps, dz, sp ? g1 : g0, dmu, dlv);
    __nv_bfloat16* o = dpre + size_t(row) * 2 * ldd + sp * 2 * L;
    const __nv_bfloat16 mh = __float2bfloat16_rn(dmu), lh = __float2bfloat16_rn(dlv);
    o[c] = mh;
    o[L + c] = lh;
    o[ldd + c] = __float2bfloat16_rn(dmu - __bfloat162float(mh));
    o[ldd + L + c] = __float2bfloat16_rn(dlv - __bfloat162float(lh));
  }
}

// zero the padding columns of dpre (ldd > 4L only when 4L is not a multiple of 16)
static __global__ void zero_pad_cols_kernel(__nv_bfloat16* p, int rows, int c0, int ldd) {
  const int w = ldd - c0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 2LL * rows * w;
       t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / (2 * w);
    const int q = int(t % (2 * w));
    p[size_t(r) * 2 * ldd + (q / w) * ldd + c0 + (q % w)] = __float2bfloat16_rn(0.f);
  }
}

// scatter the fp32 [4L, ldx] accumulation buffer (hi rows + lo rows) into the eight nn.Linear gradients:
// one block per output row (the target pointers are block-uniform), coalesced columns
static __global__ void latent_scatter_grads_kernel(const float* __restrict__ dW, int ldx, int ldd, int L, int D,
                                                   float* gw0, float* gw1, float* gw2, float* gw3,
                                                   float* gb0, float* gb1, float* gb2, float* gb3) {
  const int r = blockIdx.x;                     // 0 .. 4L-1
  const int blk = r / L, rr = r - blk * L;
  float* gw = blk == 0 ? gw0 : blk == 1 ? gw1 : blk == 2 ? gw2 : gw3;
  float* gb = blk == 0 ? gb0 : blk == 1 ? gb1 : blk == 2 ? gb2 : gb3;
  const float* hi = dW + size_t(r) * ldx;
  const float* lo = dW + size_t(ldd + r) * ldx;
  if (gw != nullptr)
    for (int c = threadIdx.x; c < D; c += blockDim.x) gw[size_t(rr) * D + c] = hi[c] + lo[c];
  if (gb != nullptr && threadIdx.x == 0) gb[rr] = hi[D] + lo[D];
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_latent_workspace_bytes(int64_t N, int64_t D, int64_t L) {
  return carve_latent(nullptr, N, D, L).bytes;
}

extern "C" int carel_latent_fwd(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L,
                                const float* const* w4, const float* const* b4, const float* eps_e,
                                const float* eps_c, float* pre, float* z, float* kl, void* ws,
                                size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && D > 0 && L > 0, "carel_latent_fwd: empty problem");
  LatentWs w = carve_latent(ws, N, D, L);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_latent_fwd: workspace too small");
  pack_x_ones_kernel<<<int((N + 7) / 8), 256, 0, st>>>(x, int(N), int(D), int(ldx), w.Xp, w.ldx);
  CAREL_LAUNCHED();
  float* bias4 = w.dW;     // contiguous [4L] staging vector at the head of the dW buffer (free during forward)
  {
    PackW4Args pa;
    for (int i = 0; i < 4; ++i) { pa.w[i] = w4[i]; pa.b[i] = b4[i]; }
    pack_w4_kernel<<<int((4 * L + 1) / 2), 64, 0, st>>>(pa, int(L), int(D), w.Wp, w.ldw, w.ldd, bias4);
    CAREL_LAUNCHED();
  }
  // bias: the four [L] vectors are gathered into pre's first row region by the epilogue via a
  // contiguous [4L] staging vector living at the head of the dW buffer (free during forward).
  if (w.ldd > 4 * L) {   // zero the unused rows between / after the two stacked copies of W4
    CAREL_CHECK_CUDA(cudaMemsetAsync(w.Wp + size_t(4 * L) * w.ldw, 0, size_t(w.ldd - 4 * L) * w.ldw * 2, st));
    CAREL_CHECK_CUDA(cudaMemsetAsync(w.Wp + size_t(w.ldd + 4 * L) * w.ldw, 0, size_t(w.ldd - 4 * L) * w.ldw * 2, st));
  }
  CUtensorMap ta, tb;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.Xp, N, D, w.ldx, 128) == 0, "carel_latent_fwd: tensor map X failed");
  CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.Wp, 4 * L, D, w.ldw, 128) == 0, "carel_latent_fwd: tensor map W failed");
  GemmSched s = make_sched(N, 4 * L, D, 256, 1, TILE_FULL);
  EpiStore<256>::Params ep;
  ep.out = pre; ep.bias = bias4; ep.out_bf16 = nullptr; ep.M = int(N); ep.N = int(4 * L);
  ep.ld = int(4 * L); ep.ld_bf16 = 0; ep.atomic = 0; ep.scale = 1.f;
  if (int rc = launch_gemm<EpiStore<256>, 256, 1, false, false>(ta, tb, ta, tb, s, ep, st, "latent_proj_fwd")) return rc;
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.klacc, 0, 2 * sizeof(double), st));
  const long long work = N * 2 * ((L + 3) / 4);
  const int blocks = int(std::min<long long>((work + 255) / 256, 148LL * 16));
  reparam_kl_fwd_kernel<<<blocks, 256, 0, st>>>(pre, int(4 * L), eps_e, eps_c, z, int(2 * L), int(N), int(L), w.klacc);
  CAREL_LAUNCHED();
  kl_finalize_kernel<<<1, 32, 0, st>>>(w.klacc, int(N), kl);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_latent_bwd(const float* grad_z, int64_t ldgz, const float* grad_kl,
                                const float* pre, const float* eps_e, const float* eps_c, int64_t N,
                                int64_t D, int64_t L, float* grad_x, int64_t ldgx,
                                float* const* grad_w4, float* const* grad_b4, void* ws,
                                size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LatentWs w = carve_latent(ws, N, D, L);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_latent_bwd: workspace too small");
  const long long work = N * 2 * L;
  const int blocks = int(std::min<long long>((work + 255) / 256, 148LL * 16));
  reparam_kl_bwd_kernel<<<blocks, 256, 0, st>>>(pre, int(4 * L), eps_e, eps_c, grad_z, int(ldgz), grad_kl,
                                               int(N), int(L), w.dpre, w.ldd);
  CAREL_LAUNCHED();
  if (w.ldd > 4 * L) {
    zero_pad_cols_kernel<<<64, 256, 0, st>>>(w.dpre, int(N), int(4 * L), w.ldd);
    CAREL_LAUNCHED();
  }
  const int64_t K2 = 2 * w.ldd;   // hi | lo
  CUtensorMap ta, tb;
  if (grad_x != nullptr) {
    // dX[N, D] = dpre[N, 4L] * W4[4L, D]     (A K-major, B MN-major)
    CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.dpre, N, K2, K2, 128) == 0, "carel_latent_bwd: tensor map dpre failed");
    CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.Wp, K2, D, w.ldw, 64) == 0, "carel_latent_bwd: tensor map W failed");
    GemmSched s = make_sched(N, D, K2, 256, 1, TILE_FULL);
    EpiStore<256>::Params ep;
    ep.out = grad_x; ep.bias = nullptr; ep.out_bf16 = nullptr; ep.M = int(N); ep.N = int(D);
    ep.ld = int(ldgx); ep.ld_bf16 = 0; ep.atomic = 0; ep.scale = 1.f;
    if (int rc = launch_gemm<EpiStore<256>, 256, 1, false, true>(ta, tb, ta, tb, s, ep, st, "latent_dx")) return rc;
  }
  bool any_w = false;
  for (int i = 0; i < 4; ++i) any_w |= (grad_w4[i] != nullptr) || (grad_b4[i] != nullptr);
  if (any_w) {
    // [dW4 | db4][4L, D+1] = dpre^T[4L, N] * [X | 1][N, D+1]    (A MN-major, B MN-major, split-K)
    CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.dpre, N, K2, K2, 64) == 0, "carel_latent_bwd: tensor map dpre^T failed");
    CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.Xp, N, D + 1, w.ldx, 64) == 0, "carel_latent_bwd: tensor map X failed");
    GemmSched s = make_sched(K2, D + 1, N, 256, 1, TILE_FULL);
    set_ksplit_auto(s);
    CAREL_CHECK_CUDA(cudaMemsetAsync(w.dW, 0, size_t(K2) * w.ldx * 4, st));
    EpiStore<256>::Params ep;
    ep.out = w.dW; ep.bias = nullptr; ep.out_bf16 = nullptr; ep.M = int(K2); ep.N = int(D + 1);
    ep.ld = w.ldx; ep.ld_bf16 = 0; ep.atomic = 1; ep.scale = 1.f;
    if (int rc = launch_gemm<EpiStore<256>, 256, 1, true, true>(ta, tb, ta, tb, s, ep, st, "latent_dw")) return rc;
    latent_scatter_grads_kernel<<<int(4 * L), 256, 0, st>>>(w.dW, w.ldx, w.ldd, int(L), int(D), grad_w4[0], grad_w4[1],
                                                    grad_w4[2], grad_w4[3], grad_b4[0], grad_b4[1],
                                                    grad_b4[2], grad_b4[3]);
    CAREL_LAUNCHED();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Inference: get_pair_preds (drl_classifier_ec_mmd_final_mul.py:265-282) = projections -> sample_prior x2 -> pair_classifier
// -> sigmoid.  The training kernels would also produce the KL sums and materialise z; here one warp per row folds the
// reparameterisation, the [2L] dot product with the pair weights and the sigmoid: 4 launches (pack X, pack W4, the
// projection GEMM, this kernel), nothing but the [N] predictions written.
// ---------------------------------------------------------------------------------------------
namespace carel {
static __global__ void pair_pred_kernel(const float* __restrict__ pre, int L, const float* __restrict__ eps_e,
                                        const float* __restrict__ eps_c, const float* __restrict__ w,
                                        const float* __restrict__ b, int N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    const float* p = pre + size_t(r) * 4 * L;
    float s = 0.f;
    for (int k = lane; k < 2 * L; k += 32) {
      const int sp = k >= L, c = k - sp * L;
      const float mu = p[sp * 2 * L + c], lv = p[sp * 2 * L + L + c];
      const float z = fmaf(__ldg((sp ? eps_c : eps_e) + c), __expf(lv), mu);
      s = fmaf(z, __ldg(w + k), s);
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = 1.f / (1.f + __expf(-(s + __ldg(b))));
  }
}
}  // namespace carel

extern "C" size_t carel_pair_preds_workspace_bytes(int64_t N, int64_t D, int64_t L) {
  return std::max(carel_latent_workspace_bytes(N, D, L), carel_latent_f32_workspace_bytes(N, D, L)) +
         align_up(size_t(N) * 4 * L * sizeof(float), 256);
}

// out [N] = sigmoid(pair_classifier([z_e | z_c])); pair_w [2L], pair_b [1] (pair_num_class = 1).  precision 0: bf16
// tensor-core projection, 2: fp32 tier.
extern "C" int carel_pair_preds(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L, const float* const* w4,
                                const float* const* b4, const float* eps_e, const float* eps_c, const float* pair_w,
                                const float* pair_b, float* out, void* ws, size_t ws_bytes, int precision, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && D > 0 && L > 0, "carel_pair_preds: empty problem");
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_pair_preds_workspace_bytes(N, D, L), "carel_pair_preds: workspace too small");
  const size_t lat_bytes = std::max(carel_latent_workspace_bytes(N, D, L), carel_latent_f32_workspace_bytes(N, D, L));
  float* pre = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + lat_bytes);
  if (precision == 2) {
    for (int i = 0; i < 4; ++i)
      if (int rc = launch_sgemm(false, false, x, ldx, w4[i], D, pre + i * L, 4 * L, N, L, D, 1.f, b4[i], 0, 1, st)) return rc;
  } else {
    LatentWs w = carve_latent(ws, N, D, L);
    pack_x_ones_kernel<<<int((N + 7) / 8), 256, 0, st>>>(x, int(N), int(D), int(ldx), w.Xp, w.ldx);
    CAREL_LAUNCHED();
    float* bias4 = w.dW;
    PackW4Args pa;
    for (int i = 0; i < 4; ++i) { pa.w[i] = w4[i]; pa.b[i] = b4[i]; }
    pack_w4_kernel<<<int((4 * L + 1) / 2), 64, 0, st>>>(pa, int(L), int(D), w.Wp, w.ldw, w.ldd, bias4);
    CAREL_LAUNCHED();
    CUtensorMap ta, tb;
    CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.Xp, N, D, w.ldx, 128) == 0, "carel_pair_preds: tensor map X failed");
    CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.Wp, 4 * L, D, w.ldw, 128) == 0, "carel_pair_preds: tensor map W failed");
    GemmSched s = make_sched(N, 4 * L, D, 256, 1, TILE_FULL);
    EpiStore<256>::Params ep;
    ep.out = pre; ep.bias = bias4; ep.out_bf16 = nullptr; ep.M = int(N); ep.N = int(4 * L);
    ep.ld = int(4 * L); ep.ld_bf16 = 0; ep.atomic = 0; ep.scale = 1.f;
    if (int rc = launch_gemm<EpiStore<256>, 256, 1, false, false>(ta, tb, ta, tb, s, ep, st, "latent_proj_fwd")) return rc;
  }
  pair_pred_kernel<<<int(std::min<int64_t>((N + 7) / 8, 148 * 16)), 256, 0, st>>>(pre, int(L), eps_e, eps_c, pair_w, pair_b, int(N), out);
  CAREL_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fp32 precision tier (sgemm.cuh): the same layer with fp32 GEMMs on the CUDA cores
// ---------------------------------------------------------------------------------------------
namespace carel {

// fp32 d(pre) [N, 4L] (same formulas as reparam_kl_bwd_kernel, no bf16 split)
static __global__ void reparam_kl_bwd_f32_kernel(const float* __restrict__ pre, const float* __restrict__ eps_e,
                                                 const float* __restrict__ eps_c, const float* __restrict__ gz, int ldgz,
                                                 const float* __restrict__ gkl, int N, int L, float* __restrict__ dpre) {
  const float invN = 1.f / float(N);
  const float g0 = gkl ? __ldg(gkl) : 0.f, g1 = gkl ? __ldg(gkl + 1) : 0.f;
  const long long total = 1LL * N * 2 * L;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = int(t % L);
    const int sp = int((t / L) % 2);
    const int row = int(t / (2LL * L));
    const size_t o = size_t(row) * 4 * L + sp * 2 * L + c;
    const float mu = pre[o], lv = pre[o + L];
    const float eps = __ldg((sp ? eps_c : eps_e) + c);
    const float dz = gz ? gz[size_t(row) * ldgz + sp * L + c] : 0.f;
    const float g = sp ? g1 : g0;
    const float e = expf(lv);
    dpre[o] = dz + g * mu * invN;
    dpre[o + L] = dz * eps * e - 0.5f * g * (1.f - e) * invN;
  }
}

}  // namespace carel

extern "C" size_t carel_latent_f32_workspace_bytes(int64_t N, int64_t D, int64_t L) {
  (void)D;
  return align_up(size_t(N) * 4 * L * sizeof(float), 256) + 256;
}

extern "C" int carel_latent_f32_fwd(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L, const float* const* w4,
                                    const float* const* b4, const float* eps_e, const float* eps_c, float* pre, float* z,
                                    float* kl, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && D > 0 && L > 0, "carel_latent_f32_fwd: empty problem");
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_latent_f32_workspace_bytes(N, D, L), "carel_latent_f32_fwd: workspace too small");
  double* klacc = reinterpret_cast<double*>(static_cast<uint8_t*>(ws) + align_up(size_t(N) * 4 * L * sizeof(float), 256));
  for (int i = 0; i < 4; ++i)       // pre[:, iL:(i+1)L] = X W_i^T + b_i   (nn.Linear, mmd:312-336)
    if (int rc = launch_sgemm(false, false, x, ldx, w4[i], D, pre + i * L, 4 * L, N, L, D, 1.f, b4[i], 0, 1, st)) return rc;
  CAREL_CHECK_CUDA(cudaMemsetAsync(klacc, 0, 2 * sizeof(double), st));
  const long long work = N * 2 * ((L + 3) / 4);
  const int blocks = int(std::min<long long>((work + 255) / 256, 148LL * 16));
  reparam_kl_fwd_kernel<<<blocks, 256, 0, st>>>(pre, int(4 * L), eps_e, eps_c, z, int(2 * L), int(N), int(L), klacc);
  CAREL_LAUNCHED();
  kl_finalize_kernel<<<1, 32, 0, st>>>(klacc, int(N), kl);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_latent_f32_bwd(const float* grad_z, int64_t ldgz, const float* grad_kl, const float* x, int64_t ldx,
                                    const float* const* w4, const float* pre, const float* eps_e, const float* eps_c,
                                    int64_t N, int64_t D, int64_t L, float* grad_x, int64_t ldgx, float* const* grad_w4,
                                    float* const* grad_b4, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= carel_latent_f32_workspace_bytes(N, D, L), "carel_latent_f32_bwd: workspace too small");
  float* dpre = static_cast<float*>(ws);
  const long long work = N * 2 * L;
  reparam_kl_bwd_f32_kernel<<<int(std::min<long long>((work + 255) / 256, 148LL * 16)), 256, 0, st>>>(
      pre, eps_e, eps_c, grad_z, int(ldgz), grad_kl, int(N), int(L), dpre);
  CAREL_LAUNCHED();
  for (int i = 0; i < 4; ++i) {
    const float* dp = dpre + i * L;                       // [N, L] block of d(pre), pitch 4L
    if (grad_x != nullptr)                                // dX (+)= dpre_i W_i
      if (int rc = launch_sgemm(false, true, dp, 4 * L, w4[i], D, grad_x, ldgx, N, D, L, 1.f, nullptr, i ? 1 : 0, 1, st)) return rc;
    if (grad_w4[i] != nullptr) {                          // dW_i = dpre_i^T X   (reduction over the N rows: split-K)
      CAREL_CHECK_CUDA(cudaMemsetAsync(grad_w4[i], 0, size_t(L) * D * sizeof(float), st));
      if (int rc = launch_sgemm(true, true, dp, 4 * L, x, ldx, grad_w4[i], D, L, D, N, 1.f, nullptr, 2, 0, st)) return rc;
    }
    if (grad_b4[i] != nullptr) {
      CAREL_CHECK_CUDA(cudaMemsetAsync(grad_b4[i], 0, size_t(L) * sizeof(float), st));
      if (int rc = launch_f32_colsum(dp, 4 * L, N, L, 1.f, grad_b4[i], st)) return rc;
    }
  }
  return 0;
}
