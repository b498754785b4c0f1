// mmd.cu — carel_mmd_{fwd,finalize,bwd}: unbiased multi-bandwidth RBF MMD^2 of two samples,
// as the reference computes it (MMDStatistic.__call__ / pdist,
// drl_classifier_ec_mmd_final_mul.py:537-596), without materialising the Gram matrix.
#include "api_common.cuh"
#include "chain_launch.cuh"
#include "chain_policies.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "pack.cuh"

namespace carel {

struct MmdWs {
  __nv_bfloat16* SA;  // [M, ldp] augmented row operand of cat(x, y)
  __nv_bfloat16* SB;  // [M, ldp] augmented column operand
  double* accum;      // [4] block sums
  float* S32;         // [M, dim] fp32 copy of cat(x, y) (fp32 mode only)
  size_t bytes;
  int ldp;
};

// precision 0: bf16 operands [s | aug] (aug_width);  1: two-term split operands (split_width)
static MmdWs carve_mmd(void* ws, int64_t M, int64_t dim, int precision) {
  WsCarver w(ws);
  MmdWs r;
  r.ldp = precision ? split_width(int(dim)) : aug_width(int(dim));
  r.SA = w.take<__nv_bfloat16>(size_t(M) * r.ldp);
  r.SB = w.take<__nv_bfloat16>(size_t(M) * r.ldp);
  r.accum = w.take<double>(4);
  r.S32 = precision ? w.take<float>(size_t(M) * dim) : nullptr;
  r.bytes = w.bytes();
  return r;
}

// FP16 operands: bf16-mode, short-K samples evaluated over their FULL row range (the weight scale of the FP16
// backward is the forward pass's maximum kernel value, which a row-block shard does not see completely).
static inline bool mmd_uses_f16(int precision, int ldp, int t0, int t1, int m_tiles) {
  return !precision && ldp <= 256 && t0 == 0 && t1 == m_tiles;
}

static __global__ void mmd_finalize_kernel(const double* __restrict__ acc, double a00, double a11,
                                           double a01, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    out[0] = float(a00 * acc[0] + a11 * acc[1] + a01 * acc[2]);
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_mmd_workspace_bytes(int64_t n1, int64_t n2, int64_t dim, int precision) {
  return carve_mmd(nullptr, n1 + n2, dim, precision).bytes;
}

extern "C" int carel_mmd_finalize(const double* partial_sums, int64_t n1, int64_t n2,
                                  float* out_mmd, void* stream) {
  CAREL_REQUIRE(n1 >= 2 && n2 >= 2, "carel_mmd: n_1 and n_2 must be >= 2 (a00 = 1/(n(n-1)))");
  const double a00 = 1.0 / (double(n1) * double(n1 - 1));
  const double a11 = 1.0 / (double(n2) * double(n2 - 1));
  const double a01 = -1.0 / (double(n1) * double(n2));
  mmd_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partial_sums, a00, a11, a01, out_mmd);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_mmd_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1,
                             int64_t n2, int64_t dim, const float* alphas_host, int n_alpha,
                             float eps, int flags, int64_t row_tile_begin, int64_t row_tile_end,
                             double* partial_sums, float* out_mmd, void* ws, size_t ws_bytes,
                             int precision, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(n1 >= 2 && n2 >= 2, "carel_mmd: n_1 and n_2 must be >= 2 (a00 = 1/(n(n-1)))");
  CAREL_REQUIRE(dim > 0, "carel_mmd: dim must be positive");
  CAREL_REQUIRE(n_alpha >= 1 && n_alpha <= kMaxAlphas, "carel_mmd: 1..%d alphas supported", kMaxAlphas);
  const int64_t M = n1 + n2;
  MmdWs w = carve_mmd(ws, M, dim, precision);
  const int ldp = w.ldp;
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_mmd_fwd: workspace too small");
  double* acc = partial_sums ? partial_sums : w.accum;

  // short-K bf16-mode samples (the head's latent codes: an intermediate of bounded range) go to the tensor core as
  // FP16 -- same MMA rate, 11 mantissa bits; the wide Gram sweep and the split (fp32-mode) operands stay BF16
  const int m_tiles_all = int((M + 127) / 128);
  const bool f16 = mmd_uses_f16(precision, ldp, int(row_tile_begin < 0 ? 0 : row_tile_begin),
                                int((row_tile_end < 0 || row_tile_end > m_tiles_all) ? m_tiles_all : row_tile_end), m_tiles_all);
  if (!precision) {
    pack_rows_kernel<<<int((M + 7) / 8), 256, 0, st>>>(x, int(M), int(dim), int(ldx), w.SA, w.SB, ldp, 0, 1, nullptr,
                                                      nullptr, y, int(n1), int(ldy), f16 ? 1 : 0);
    CAREL_LAUNCHED();
  } else {
    const Pack3Out oa{w.SA, 0, 0}, ob{w.SB, 1, 1};
    pack3_rows_kernel<<<int((n1 + 7) / 8), 256, 0, st>>>(x, int(n1), int(dim), int(ldx), oa, ob, 0, 0, nullptr, w.S32, int(dim));
    CAREL_LAUNCHED();
    pack3_rows_kernel<<<int((n2 + 7) / 8), 256, 0, st>>>(y, int(n2), int(dim), int(ldy), oa, ob, int(n1), 0, nullptr, w.S32, int(dim));
    CAREL_LAUNCHED();
  }
  CAREL_CHECK_CUDA(cudaMemsetAsync(acc, 0, 3 * sizeof(double), st));
  unsigned* kmax_bits = reinterpret_cast<unsigned*>(w.accum + 3);
  CAREL_CHECK_CUDA(cudaMemsetAsync(kmax_bits, 0, sizeof(double), st));

  CUtensorMap ta, tb;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.SA, M, ldp, ldp, 128) == 0, "carel_mmd_fwd: tensor map failed");
  CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.SB, M, ldp, ldp, 128) == 0, "carel_mmd_fwd: tensor map failed");
  const int tri = (flags & CAREL_MMD_FULL_GRAM) ? 0 : 1;
  GemmSched s = make_sched(M, M, ldp, 256, 1, tri ? TILE_TRI : TILE_FULL);
  const int t0 = int(row_tile_begin < 0 ? 0 : row_tile_begin);
  const int t1 = int((row_tile_end < 0 || row_tile_end > s.m_tiles) ? s.m_tiles : row_tile_end);
  CAREL_REQUIRE(t0 <= t1, "carel_mmd_fwd: bad row tile range");
  s.m_tile0 = t0;
  s.tri_rows = t1 - t0;
  if (tri) {
    int items = 0;
    for (int t = t0; t < t1; ++t) items += s.n_tiles - (t >> 1);
    s.num_items = items;
  } else {
    s.num_items = (t1 - t0) * s.n_tiles;
  }
  EpiMMDParams ep;
  ep.accum = acc; ep.kmax_bits = kmax_bits; ep.M = int(M); ep.n1 = int(n1); ep.tri = tri; ep.n_alpha = n_alpha;
  for (int i = 0; i < kMaxAlphas; ++i) {
    const float a = (i < n_alpha) ? alphas_host[i] : 0.f;
    ep.a_scale[i] = -2.f * a * kLog2e;
    ep.a_shift[i] = -a * eps * kLog2e;
  }
  if (ldp <= 256) {
    // short K (latent-sized samples): A-stationary schedule, only the column operand is streamed
    AsSched as = make_as_sched(M, M, ldp, sm_count(), tri, t0, t1);
    if (f16) {
      if (int rc = launch_gemm_as<EpiMMD<256>, true>(ta, tb, as, ep, st, "mmd_gram_fwd")) return rc;
    } else if (int rc = launch_gemm_as<EpiMMD<256>>(ta, tb, as, ep, st, "mmd_gram_fwd")) {
      return rc;
    }
  } else if ((t0 & 1) == 0 && ((t1 & 1) == 0 || t1 == s.m_tiles) && !(flags & CAREL_MMD_NO_PAIRS)) {
    // long K (the d = 768 Gram sweep): CTA pairs, one M = 256 UMMA across two SMs (gemm_2sm.cuh)
    PairSched ps = make_pair_sched(M, M, ldp, tri, t0, t1, sm_count());
    if (int rc = launch_gemm_2sm<EpiMMD<256>>(ta, tb, ps, ep, st, "mmd_gram_fwd")) return rc;
  } else if (int rc = launch_gemm<EpiMMD<256>, 256, 1, false, false>(ta, tb, ta, tb, s, ep, st, "mmd_gram_fwd")) {
    return rc;
  }
  if (out_mmd != nullptr) return carel_mmd_finalize(acc, n1, n2, out_mmd, stream);
  return 0;
}

template <int BN1, int NB, bool F16 = false, bool HILO = true>
static int mmd_bwd_launch(const MmdWs& w, MmdBwdParams pp, int64_t M, int ldp, int n2_16, int t0, int t1,
                          float* grad_x, int64_t ldgx, float* grad_y, int64_t ldgy, int64_t n1, int64_t n2,
                          int64_t dim, cudaStream_t st) {
  using Cfg = ChainCfg<1, BN1, NB>;
  CUtensorMap ta, tb;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.SA, M, ldp, ldp, 128) == 0, "carel_mmd_bwd: tensor map A failed");
  CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.SB, M, ldp, ldp, BN1) == 0, "carel_mmd_bwd: tensor map B failed");
  const int k16 = ldp / 16;
  ChainSched s = make_chain_sched(t0, t1 - t0, int((M + BN1 - 1) / BN1), k16, Cfg::stages_for(k16), 32);
  s.n2_16 = n2_16;
  // rows outside [t0, t1) * 128 must read as zero; with one work item per row tile the rows inside are
  // written exactly once (plain stores), otherwise partial sums are accumulated atomically
  pp.atomic = s.nsplit > 1;
  if (pp.atomic || t0 > 0 || t1 * 128 < M) {
    CAREL_CHECK_CUDA(cudaMemset2DAsync(grad_x, size_t(ldgx) * 4, 0, size_t(dim) * 4, size_t(n1), st));
    CAREL_CHECK_CUDA(cudaMemset2DAsync(grad_y, size_t(ldgy) * 4, 0, size_t(dim) * 4, size_t(n2), st));
  }
  return launch_chain<PolMmdBwdT<F16, HILO>, 1, BN1, NB>(ta, tb, ta, tb, s, pp, st, "mmd_gram_bwd");
}

extern "C" int carel_mmd_bwd(const float* grad_out, int64_t n1, int64_t n2, int64_t dim,
                             const float* alphas_host, int n_alpha, float eps,
                             int64_t row_tile_begin, int64_t row_tile_end, float* grad_x,
                             int64_t ldgx, float* grad_y, int64_t ldgy, void* ws, size_t ws_bytes,
                             int precision, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(n1 >= 2 && n2 >= 2, "carel_mmd: n_1 and n_2 must be >= 2 (a00 = 1/(n(n-1)))");
  CAREL_REQUIRE(n_alpha >= 1 && n_alpha <= kMaxAlphas, "carel_mmd: 1..%d alphas supported", kMaxAlphas);
  const int64_t M = n1 + n2;
  MmdWs w = carve_mmd(ws, M, dim, precision);
  const int ldp = w.ldp;
  const int n2cols = precision ? 2 * split_kp(int(dim)) + 16 : ldp;      // GEMM-2 width
  CAREL_REQUIRE(n2cols <= 256, "carel_mmd_bwd: feature width %lld is not supported by the fused backward (max 250 in bf16 mode, 120 in fp32 mode)", (long long)dim);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_mmd_bwd: workspace too small");
  const int m_tiles = int((M + 127) / 128);
  const int t0 = int(row_tile_begin < 0 ? 0 : row_tile_begin);
  const int t1 = int((row_tile_end < 0 || row_tile_end > m_tiles) ? m_tiles : row_tile_end);
  CAREL_REQUIRE(t0 <= t1, "carel_mmd_bwd: bad row tile range");
  MmdBwdParams pp;
  pp.split = precision ? 1 : 0; pp.kp = split_kp(int(dim)); pp.ld32 = int(dim); pp.S32 = w.S32;
  pp.SA = w.SA; pp.grad_out = grad_out; pp.grad_x = grad_x; pp.grad_y = grad_y;
  pp.ldgx = int(ldgx); pp.ldgy = int(ldgy); pp.dim = int(dim); pp.ldp = ldp;
  pp.M = int(M); pp.n1 = int(n1);
  pp.c00 = float(1.0 / (double(n1) * double(n1 - 1)));
  pp.c11 = float(1.0 / (double(n2) * double(n2 - 1)));
  pp.c01 = float(-1.0 / (double(n1) * double(n2)));
  pp.n_alpha = n_alpha;
  float amax = 0.f;
  for (int i = 0; i < kMaxAlphas; ++i) {
    const float a = (i < n_alpha) ? alphas_host[i] : 0.f;
    pp.a_scale[i] = -2.f * a * kLog2e;
    pp.a_shift[i] = -a * eps * kLog2e;
    pp.a_pos[i] = a;
    amax = fmaxf(amax, fabsf(a));
  }
  // kind::f16 does not accept an FP16 A with a BF16 B (illegal instruction on sm_100a): the weight tile P has the
  // operands' format.  BF16 has fp32's exponent range; FP16 weights are scaled in the kernel (PolMmdBwdT).
  pp.kmax_bits = reinterpret_cast<const unsigned*>(w.accum + 3);
  pp.wmax_coef = fmaxf(fmaxf(pp.c00, pp.c11), fabsf(pp.c01)) * amax;
  if (mmd_uses_f16(precision, ldp, t0, t1, m_tiles)) {
    if (M >= 4096)      // single FP16 weights: rounding 2^-12 / sqrt(M) < 4e-6 of a gradient row
      return mmd_bwd_launch<128, 2, true, false>(w, pp, M, ldp, n2cols / 16, t0, t1, grad_x, ldgx, grad_y, ldgy, n1, n2, dim, st);
    return mmd_bwd_launch<128, 2, true, true>(w, pp, M, ldp, n2cols / 16, t0, t1, grad_x, ldgx, grad_y, ldgy, n1, n2, dim, st);
  }
  return mmd_bwd_launch<128, 2>(w, pp, M, ldp, n2cols / 16, t0, t1, grad_x, ldgx, grad_y, ldgy, n1, n2, dim, st);
}

#ifdef CAREL_CHAIN_TRACE
// debug: chain-kernel timeline of this translation unit (see chain_tc.cuh)
extern "C" int carel_debug_chain_trace_mmd(long long* host_dst) {
  CAREL_CHECK_CUDA(cudaMemcpyFromSymbol(host_dst, carel::g_chain_trace, sizeof(long long) * 8 * 256));
  return 0;
}
#endif
