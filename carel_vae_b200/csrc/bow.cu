// bow.cu — carel_bow_{fwd,bwd}: bag-of-words reconstruction decoder + loss
// (drl_classifier_ec_mmd_final_mul.py:253-254 decoder + Softmax, :381-387 get_reconstruct_loss):
//
//   p = softmax(Z W_d^T + b_d),  t = bow (1 - ls) + ls / V,
//   loss = -mean_{i,j} [ t max(log p, -100) + (1 - t) max(log(1 - p), -100) ]          (BCELoss)
//
// With c = ls / V the loss splits into a DENSE part that depends on p only and a SPARSE correction on the
// non-zeros of bow:
//   t A + (1-t) B = c A + (1-c) B + (1-ls) bow (A - B),      A = max(log p,-100), B = max(log1p(-p),-100)
// so the GEMM epilogues never read the 1 GB dense target; it is packed once into an ELL list (HBM-bound pass) and
// the corrections are O(nnz) warp-per-row work.  Same split in backward:
//   g = (p - t) / max(p (1-p), 1e-12) / (N V)   (torch's BCELoss backward),  s_i = sum_j g_ij p_ij,
//   dlogit = p (g - s_i)  =  [u - p S_i] / (N V)  +  sparse,    u = p (p - c) / max(p (1-p), 1e-12).
//
// Round 2 layout ("Y-staged"): operands are FP16 (ZH [N, ldk], WH [V, ldk]: 11 mantissa bits at the bf16 MMA rate;
// the latent sample is an intermediate, and its 8-bit rounding as a bf16 operand alone put the head-level dW
// gradient at 1.1e-3).  The one dense forward GEMM (pass F) also stores the logits ONCE, bias-free and centred per
// row, as 16-bit FIXED POINT with one scale per row:
//   Y16_ij = round( (log2(e) z_i.w_j - rm_i) s_i ) + 32768,   rm_i = log2(e) z_i.mean_j(w_j),
//   s_i = 32000 / (log2(e) |z_i| (max_j |w_j| + |mean_j w_j|))        (Cauchy-Schwarz: no saturation possible)
// so that log2 p has the same ABSOLUTE precision for every element of a row (p the same RELATIVE precision), however
// peaked the row is — fp16 would lose the large logits (spacing 2^-5 at |y| = 32).  At the default initialisation the
// step is 5e-4 log2 units (p to 1e-4 relative).  2 bytes per element through a swizzled shared-memory stage and TMA
// stores.  Everything downstream reads Y:
//   pass F    gemm_as + EpiBowFused  per-row running reference + power sums -> lse, dense loss, S_i (bow_combine)
//   exact     bow_exact_rows_kernel  rows outside the series' range (max p >= 2^-5 or clamped logs) are redone
//                                    exactly, PER ROW, from Y (recomputing lse from the rounded logits)
//   sparse    bow_sparse_fwd_kernel  corrections of loss and S_i on the non-zeros (one gather from Y each)
//   backward  sparse gathers (initialise grad_z / grad_w / grad_b), then two single-GEMM kernels (ygrad.cuh):
//             dW += dlogit^T Z (owner = 128 words) and dZ += dlogit W_d (owner = 128 clauses); the dlogit tile is
//             formed in registers from the Y tile and goes to TMEM as the A operand.  No logit is recomputed.
#include <algorithm>
#include <cstdlib>

#include "api_common.cuh"
#include "chain_launch.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "ygrad.cuh"

namespace carel {

struct BowWs {
  __half* ZH;           // [N, ldk]
  __half* WH;           // [V, ldk]
  uint16_t* Y;          // [N, ldy]  centred log2-domain logits, 16-bit fixed point (see above)
  float4* part4;        // [ntp, Npad] pass-F partials (m, T1, T2, T3)
  float2* part2;        // [ntp, Npad] pass-F partials (T4, min)
  float* wsum;          // [ldk + 4] column sums of WH; wsum[ldk] = sum_j b_j; wsum[ldk + 1] = max_j |w_j| (as bits)
  float* rm;            // [Npad]  row reference rm_i (log2 units; 0 beyond N)
  float2* rowenc;       // [Npad]  {a_i = log2(e) s_i, bmagic_i}: the encoder constants of pass F
  float4* rowc;         // [2, Npad]  plane 0: {1 / s_i, bmagic_i, -lse2_i, -S_i 2^10}: the decoder constants (0 beyond N);
                        //            plane 1: {a1, a2, a3, fast}: the backward fast-path polynomial (ygrad.cuh)
  int* row_flag;        // [Npad] 1: row needs the exact pass
  int* flag_list;       // [Npad] compact list of the flagged rows
  int* counters;        // [4]  [0] number of flagged rows, [1] ELL overflow without a dense target
  float* bl2e;          // [V rounded up to 128]  b log2(e) (0 beyond V)
  float* dWacc;         // [V, K] scratch target of dW when the caller does not want grad_w
  float* dbacc;         // [V]    scratch target of db when the caller does not want grad_b
  double* lossacc;      // [2] dense, sparse
  int* ell_cols;        // [N, cap]
  float* ell_vals;      // [N, cap]
  int* ell_nnz;         // [N]  (> cap: row must be scanned densely)
  float* ell_ev;        // [N, cap]  per non-zero: -(1-ls) t p / max(p(1-p),1e-12)  (forward -> backward)
  int* colptr;          // [V+1]  CSC view of the ELL list (built on the device every forward)
  int* fillpos;         // [V+1]
  int* csc_row;         // [N*cap]
  int* csc_slot;        // [N*cap]
  size_t bytes;
  int ldk, ldy, Npad, ntp;
};

static BowWs carve_bow(void* ws, int64_t N, int64_t K, int64_t V, int64_t cap) {
  WsCarver w(ws);
  BowWs r;
  r.ldk = int(round_up64(K, 16));
  r.ldy = int(round_up64(V, 64));
  r.Npad = int(round_up64(N, 128));
  r.ntp = int((V + 255) / 256) * kColSplit;
  r.ZH = w.take<__half>(size_t(N) * r.ldk);
  r.WH = w.take<__half>(size_t(V) * r.ldk);
  r.Y = w.take<uint16_t>(size_t(N) * r.ldy);
  r.part4 = w.take<float4>(size_t(r.ntp) * r.Npad);
  r.part2 = w.take<float2>(size_t(r.ntp) * r.Npad);
  r.wsum = w.take<float>(size_t(r.ldk) + 4);
  r.rm = w.take<float>(size_t(r.Npad));
  r.rowenc = w.take<float2>(size_t(r.Npad));
  r.rowc = w.take<float4>(size_t(2) * r.Npad);
  r.row_flag = w.take<int>(size_t(r.Npad));
  r.flag_list = w.take<int>(size_t(r.Npad));
  r.counters = w.take<int>(4);
  r.bl2e = w.take<float>(size_t(round_up64(V, 128)));
  r.dWacc = w.take<float>(size_t(V) * K);
  r.dbacc = w.take<float>(size_t(V));
  r.lossacc = w.take<double>(2);
  r.ell_cols = w.take<int>(size_t(N) * cap);
  r.ell_vals = w.take<float>(size_t(N) * cap);
  r.ell_nnz = w.take<int>(size_t(N));
  r.ell_ev = w.take<float>(size_t(N) * cap);
  r.colptr = w.take<int>(size_t(V) + 1);
  r.fillpos = w.take<int>(size_t(V) + 1);
  r.csc_row = w.take<int>(size_t(N) * cap);
  r.csc_slot = w.take<int>(size_t(N) * cap);
  r.bytes = w.bytes();
  return r;
}

// ---------------------------------------------------------------------------------------------
// dense BoW target -> ELL (one warp per row, 128-bit loads; HBM-bound)
// ---------------------------------------------------------------------------------------------
// Measured and rejected (a two-stream probe on B200): a small-footprint form (U = 12, 128 threads, 72 registers, one block
// per SM, max-shared carve-out) on a forked stream, meant to stream the target underneath the latent GEMM / MMD forward /
// pass F.  By the resource arithmetic it is co-resident with the 576-thread, 96-register, 225 KB tensor-core CTAs, but
// the two never share an SM in practice (MMD forward x3: 263 us alone, 543 us with the 380 us background pack).
template <int U, int THREADS, int MINB>
static __global__ void __launch_bounds__(THREADS, MINB)
bow_pack_ell_kernel(const float* __restrict__ bow, int64_t ld, int N, int V, int cap, int* __restrict__ cols,
                    float* __restrict__ vals, int* __restrict__ nnz) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    const float* row = bow + size_t(r) * ld;
    int count = 0;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(bow) & 15) == 0);
    const int V4 = vec ? (V & ~3) : 0;
    // U 128-bit loads per lane in flight (128 U floats of the row per iteration) and ONE ballot for the common
    // all-zero batch: 4 loads + a ballot per load measured 5.69 TB/s, U = 8 in this form 6.72 TB/s (tools/pack_rate.cu)
    for (int base = 0; base < V4; base += 128 * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = base + u * 128 + lane * 4;
        v[u] = (j < V4) ? __ldcs(reinterpret_cast<const float4*>(row + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      bool any = false;
#pragma unroll
      for (int u = 0; u < U; ++u) any |= (v[u].x != 0.f) || (v[u].y != 0.f) || (v[u].z != 0.f) || (v[u].w != 0.f);
      if (__ballot_sync(0xffffffffu, any) == 0u) continue;           // the common case: 1024 zeros
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = base + u * 128 + lane * 4;
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const bool any_nz = (e[0] != 0.f) || (e[1] != 0.f) || (e[2] != 0.f) || (e[3] != 0.f);
        if (__ballot_sync(0xffffffffu, any_nz) == 0u) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool nz = e[q] != 0.f;
          const unsigned m = __ballot_sync(0xffffffffu, nz);
          if (nz) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) { cols[size_t(r) * cap + pos] = j + q; vals[size_t(r) * cap + pos] = e[q]; }
          }
          count += __popc(m);
        }
      }
    }
    for (int base = V4; base < V; base += 32) {
      const int j = base + lane;
      const float v = (j < V) ? __ldg(row + j) : 0.f;
      const bool nz = v != 0.f;
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      if (nz) {
        const int pos = count + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) { cols[size_t(r) * cap + pos] = j; vals[size_t(r) * cap + pos] = v; }
      }
      count += __popc(m);
    }
    if (lane == 0) nnz[r] = count;
  }
}


// ---------------------------------------------------------------------------------------------
// operand packs: fp32 rows -> fp16 rows (saturating, zero padded to ldk)
// ---------------------------------------------------------------------------------------------
// One warp per row.
//   W pack (wsum == nullptr): also max_j |wh_j| into *rmax_bits (positive floats order like their bit patterns).
//   Z pack (wsum != nullptr): also the row reference rm_i and the fixed-point scale of the stored logits:
//     rm_i = log2(e) zh_i . wsum / V,   bound_i = log2(e) |zh_i| (max_j |wh_j| + |wsum / V|) >= |log2(e) zh_i.wh_j - rm_i|,
//     s_i = 32000 / bound_i,   rowenc[i] = {log2(e) s_i, 2^23 + 32768 - rm_i s_i},   rowc[i] = {1 / s_i, bmagic_i, 0, 0}.
static __global__ void bow_pack_f16_kernel(const float* __restrict__ src, int rows, int K, int64_t ld,
                                           __half* __restrict__ dst, int ldk, const float* __restrict__ wsum,
                                           float inv_v, float* __restrict__ rm, float2* __restrict__ rowenc,
                                           float4* __restrict__ rowc, unsigned* __restrict__ rmax_bits) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float wbar2 = 0.f, rmax = 0.f;
  if (wsum != nullptr) {
    for (int k = lane; k < ldk; k += 32) { const float v = __ldg(wsum + k) * inv_v; wbar2 = fmaf(v, v, wbar2); }
    wbar2 = warp_sum(wbar2);
    rmax = __uint_as_float(*rmax_bits);
  }
  float wmax2 = 0.f;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* s = src + size_t(r) * ld;
    __half* d = dst + size_t(r) * ldk;
    float dot = 0.f, nrm2 = 0.f;
    const bool vec = ((K & 3) == 0) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
      for (int k = lane * 4; k < ldk; k += 128) {
        const float4 v = (k < K) ? __ldg(reinterpret_cast<const float4*>(s + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 o;
        o.x = pack_f16x2_sat(v.x, v.y);
        o.y = pack_f16x2_sat(v.z, v.w);
        *reinterpret_cast<uint2*>(d + k) = o;
        const float2 a = unpack_f16x2(o.x), b = unpack_f16x2(o.y);
        nrm2 += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
        if (wsum != nullptr) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wsum + k));
          dot += a.x * w4.x + a.y * w4.y + b.x * w4.z + b.y * w4.w;
        }
      }
    } else {
      for (int k = lane * 2; k < ldk; k += 64) {
        const float v0 = (k < K) ? __ldg(s + k) : 0.f, v1 = (k + 1 < K) ? __ldg(s + k + 1) : 0.f;
        const uint32_t o = pack_f16x2_sat(v0, v1);
        *reinterpret_cast<uint32_t*>(d + k) = o;
        const float2 a = unpack_f16x2(o);
        nrm2 += a.x * a.x + a.y * a.y;
        if (wsum != nullptr) dot += a.x * __ldg(wsum + k) + a.y * __ldg(wsum + k + 1);
      }
    }
    nrm2 = warp_sum(nrm2);
    if (wsum == nullptr) {
      wmax2 = fmaxf(wmax2, nrm2);
    } else {
      dot = warp_sum(dot);
      if (lane == 0) {
        const float rmi = dot * inv_v * kLog2e;
        const float bound = kLog2e * sqrtf(nrm2) * (rmax + sqrtf(wbar2)) * 1.002f;
        const float sc = bound > 1e-20f ? 32000.f / bound : 1.f;
        const float bmagic = (32768.f - rmi * sc) + 8388608.f;
        rm[r] = rmi;
        rowenc[r] = make_float2(kLog2e * sc, bmagic);
        rowc[r] = make_float4(1.f / sc, bmagic, 0.f, 0.f);
      }
    }
  }
  if (wsum == nullptr) {      // one atomic per BLOCK (30000 per-warp atomics on one address serialise to ~30 us)
    __shared__ float sh_max[32];
    if (lane == 0) sh_max[threadIdx.x >> 5] = wmax2;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int i = 0; i < wpb; ++i) m = fmaxf(m, sh_max[i]);
      if (m > 0.f) atomicMax(rmax_bits, __float_as_uint(sqrtf(m) * 1.0001f));
    }
  }
}

// Column sums of the packed decoder operand WH [V, ldk] (fp16 -> fp32: sum_j z_i.w_j = ZH_i . wsum), wsum[ldk] =
// sum_j b_j, and bl2e[j] = b_j log2(e) (zero padded), in one launch.
// Block = (ldk / 8 chunks of 8 halves) x 16 row lanes, 128-bit loads; one atomicAdd per column and block.
static __global__ void bow_colsum_bias_kernel(const __half* __restrict__ WH, int V, int ldk,
                                              float* __restrict__ wsum, const float* __restrict__ b, int Vpad,
                                              float* __restrict__ bl2e) {
  extern __shared__ float sh_cs[];           // [blockDim.y][ldk]
  __shared__ float sh_b[32];
  const int nx = ldk >> 3;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  float bs = 0.f;
  for (int j = blockIdx.x * nthreads + tid; j < Vpad; j += gridDim.x * nthreads) {
    const float bj = (j < V) ? __ldg(b + j) : 0.f;
    bl2e[j] = bj * kLog2e;
    bs += bj;
  }
  bs = warp_sum(bs);
  if ((tid & 31) == 0) sh_b[tid >> 5] = bs;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (threadIdx.x < nx) {
    for (int r = blockIdx.x * blockDim.y + threadIdx.y; r < V; r += gridDim.x * blockDim.y) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(WH + size_t(r) * ldk) + threadIdx.x);
      const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 f = unpack_f16x2(u[h]);
        acc[2 * h] += f.x;
        acc[2 * h + 1] += f.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sh_cs[threadIdx.y * ldk + threadIdx.x * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int k = tid; k < ldk; k += nthreads) {
    float t = 0.f;
    for (int y = 0; y < blockDim.y; ++y) t += sh_cs[y * ldk + k];
    atomicAdd(wsum + k, t);
  }
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < (nthreads + 31) / 32; ++i) t += sh_b[i];
    if (t != 0.f) atomicAdd(wsum + ldk, t);
  }
}

// ---------------------------------------------------------------------------------------------
// pass F epilogue (the ONE dense forward pass): per row and per 64-column part, with y = log2-domain logit
// (acc log2 e + b_j log2 e) and m = a running reference (moved when a chunk maximum exceeds it by more than 2^8):
//   T_k = sum_j 2^(k (y_ij - m)), k = 1..4,   and   mn = min_j y_ij.
// From these the combine kernel gets lse_i, the power sums S_k = sum_j p_ij^k = T_k / T_1^k and with them
//   sum_j log1p(-p) = -(S1 + S2/2 + S3/3 + S4/4 + ...),   sum_j (p - c)/(1 - p) = (S1+S2+S3+S4+..) - c (V + S1+S2+S3+..)
// (S1 = 1), valid to 2^-20 whenever max_j p_ij < 2^-5 (certified by S4 < 2^-20); rows beyond that (or with a
// clamped log) are flagged and redone exactly, per row, by bow_exact_rows_kernel.  Hot loop per PAIR of elements:
// FFMA2 (y), FADD2 (y - m), 2 MUFU.EX2 issued before the chunk's max / min chains, FMUL2, 2 FADD2, 2 FFMA2,
// 2 x 1/2 FMNMX3 — plus, for the stored logits, FFMA2 (acc a_i + bmagic_i) and one PRMT that packs the two low halves.
//
// Stored logits: each epilogue warp stages its 32 rows x 32 columns of 16-bit values (2 KB, 64-byte swizzle: conflict-free
// 16-byte stores) and one lane issues a TMA store of the box; the stage is reused once the previous store has read it.
//
// The A-stationary kernel walks consecutive column tiles of ONE row tile, so a thread meets the same row again in
// the next tile: the running state (m, T1..T4, mn) stays in registers across the tiles of a row segment and is
// written once per segment (seg = this CTA's index among the CTAs that share the row tile, <= (n_tiles-1)/tpc + 2
// of them) instead of once per column tile: 472 -> 16 partials per row at V = 30000.
// ---------------------------------------------------------------------------------------------
constexpr int kBowEpiSmem = 16 * 2048;
struct EpiBowFusedParams {
  CUtensorMap tmY;     // store map of Y [N, V] (16-bit): box 32 columns x 32 rows, 64-byte swizzle
  float4* part4;   // [nseg * kColSplit][Npad]  (m, T1, T2, T3)     m, mn in log2 units; zeroed before the pass
  float2* part2;   // [nseg * kColSplit][Npad]  (T4, mn)
  const float* bl2e;   // [Vpad]  b_j log2(e)
  const float2* rowenc;  // [Npad]  {a_i, bmagic_i}: Y16 = low 16 bits of fl(acc a_i + bmagic_i)
  int N, V, Npad;
  int n_tiles, tpc;    // of the AsSched the pass runs under (m_tile0 = 0, full grid)
};
template <int BN>
struct EpiBowFused {
  using Params = EpiBowFusedParams;
  const Params& p;
  float m, mn;
  float2 enc;
  float2 t1, t2, t3, t4;
  CAREL_DEVINL EpiBowFused(const Params& p_, const TileCtx&) : p(p_), enc(make_float2(0.f, 8388608.f + 32768.f)) {}
  CAREL_DEVINL void prefetch(const TileCtx&) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    static_assert(HALF == 64, "the Y stage holds 2 x 32 columns per warp and tile");
    const int c0 = ctx.col0 + ctx.half * HALF;
    if (ctx.first_of_row) {
      m = -INFINITY; mn = INFINITY;
      t1 = make_float2(0.f, 0.f); t2 = t1; t3 = t1; t4 = t1;
      if (ctx.row < p.N) enc = __ldg(p.rowenc + ctx.row);
    }
    if (c0 < p.V) {
      const bool full = c0 + HALF <= p.V;
      const float2 l2e = make_float2(kLog2e, kLog2e);
      const float4* cc4 = reinterpret_cast<const float4*>(p.bl2e + c0);
      const int lane = ctx.epi_tid & 31, wq = (ctx.epi_tid >> 5) & 3;
      uint8_t* stg = ctx.epi_smem + (ctx.epi_tid >> 5) * 2048;
      uint8_t* stg_row = stg + lane * 64;
      const int sw = (lane >> 1) & 3;
      const float2 ea2 = make_float2(enc.x, enc.x), eb2 = make_float2(enc.y, enc.y);
      for_each_chunk16(ctx.taddr + ctx.half * HALF, HALF / 16, [&](int c, const uint32_t (&r)[16]) {
        // b_j log2(e) of this chunk's 16 words: issued before the store block below (whose fences the compiler does
        // not move loads across), so the global-load latency is covered by it (ncu: 15 % of the stall samples sat
        // on the first use when the loads followed the fence)
        float4 bc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bc[i] = __ldg(cc4 + c * 4 + i);
        // ---- stored logits: low 16 bits of fl(acc a_i + bmagic_i) -> swizzled stage -> TMA store per 32 columns ----
        {
          if ((c & 1) == 0) {                   // the previous store of this warp has read the stage
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          const int lc = (c & 1) * 2;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t hs[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 v = __ffma2_rn(make_float2(__uint_as_float(r[8 * hh + 2 * i]), __uint_as_float(r[8 * hh + 2 * i + 1])),
                                          ea2, eb2);
              hs[i] = __byte_perm(__float_as_uint(v.x), __float_as_uint(v.y), 0x5410);
            }
            *reinterpret_cast<uint4*>(stg_row + (((lc + hh) ^ sw) << 4)) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
          }
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            const int col = c0 + (c >> 1) * 32;
            if (lane == 0 && col < p.V) {
              tma_store_2d(&p.tmY, stg, col, ctx.tm * 128 + wq * 32);
              tma_store_commit();
            }
          }
        }
        // y = log2-domain logit: acc log2(e) + b_j log2(e)   (plain operands; the bias is a per-column constant)
        float y[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = bc[i];
          const float2 y0 = __ffma2_rn(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), l2e,
                                       make_float2(b4.x, b4.y));
          const float2 y1 = __ffma2_rn(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), l2e,
                                       make_float2(b4.z, b4.w));
          y[4 * i] = y0.x; y[4 * i + 1] = y0.y; y[4 * i + 2] = y1.x; y[4 * i + 3] = y1.y;
        }
        // Speculative exponentials against the CURRENT reference m (not yet the chunk maximum): they do not wait
        // for the max chain below.  The reference only moves when the chunk maximum exceeds it by more than 2^8
        // (then e <= 2^8, e^4 <= 2^32: no overflow), in which case the chunk is redone against the new one.
        const float2 nm2 = make_float2(-m, -m);
        float2 e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 t = __fadd2_rn(make_float2(y[2 * i], y[2 * i + 1]), nm2);
          e[i] = make_float2(ex2_approx(t.x), ex2_approx(t.y));
        }
        float cm, cn;
        if (full) {
          const float ma = fmax3(y[0], y[1], y[2]), mb = fmax3(y[3], y[4], y[5]);
          const float mc = fmax3(y[6], y[7], y[8]), md = fmax3(y[9], y[10], y[11]);
          const float na = fmin3(y[0], y[1], y[2]), nb = fmin3(y[3], y[4], y[5]);
          const float nc = fmin3(y[6], y[7], y[8]), nd = fmin3(y[9], y[10], y[11]);
          cm = fmax3(fmax3(ma, mb, mc), md, fmax3(y[12], y[13], fmaxf(y[14], y[15])));
          cn = fmin3(fmin3(na, nb, nc), nd, fmin3(y[12], y[13], fminf(y[14], y[15])));
        } else {
          cm = -INFINITY;
          cn = INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = c0 + c * 16 + i < p.V;
            cn = fminf(cn, ok ? y[i] : INFINITY);
            y[i] = ok ? y[i] : -INFINITY;
            cm = fmaxf(cm, y[i]);
          }
        }
        mn = fminf(mn, cn);
        if (cm > m + 8.f || !full) {          // move the reference (rare after the first chunk) and redo the chunk
          if (cm > m) {
            const float sc = ex2_approx(m - cm), sc2 = sc * sc;
            t1.x *= sc; t1.y *= sc;
            t2.x *= sc2; t2.y *= sc2;
            t3.x *= sc2 * sc; t3.y *= sc2 * sc;
            t4.x *= sc2 * sc2; t4.y *= sc2 * sc2;
            m = cm;
          }
          const float2 nm = make_float2(-m, -m);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 t = __fadd2_rn(make_float2(y[2 * i], y[2 * i + 1]), nm);
            e[i] = (m > -INFINITY) ? make_float2(ex2_approx(t.x), ex2_approx(t.y)) : make_float2(0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 e2 = __fmul2_rn(e[i], e[i]);
          t1 = __fadd2_rn(t1, e[i]);
          t2 = __fadd2_rn(t2, e2);
          t3 = __ffma2_rn(e2, e[i], t3);
          t4 = __ffma2_rn(e2, e2, t4);
        }
      });
    }
    if (ctx.last_of_row && ctx.row < p.Npad) {
      const int seg = int(blockIdx.x) - (ctx.tm * p.n_tiles) / p.tpc;
      const size_t o = size_t(seg * kColSplit + ctx.half) * p.Npad + ctx.row;
      p.part4[o] = make_float4(m, t1.x + t1.y, t2.x + t2.y, t3.x + t3.y);
      p.part2[o] = make_float2(t4.x + t4.y, mn);
    }
  }
  CAREL_DEVINL void finish() {
    if ((threadIdx.x & 31) == 0) tma_store_wait<0>();      // every TMA store of this warp has completed
  }
};

// Combine the pass-F partials of the rows: lse_i, power sums, dense loss term and S_i = sum_j u_ij; writes
// rowc[i].z = -lse2_i, rowc[i].w = -S_i 2^10.  Rows outside the series' range are flagged (row_flag, appended to flag_list) and contribute
// nothing here.
// Block = 32 rows x 16 partial groups: thread (r, g) folds partials g, g + 16, ... of row r online (consecutive
// threads read consecutive rows: coalesced), the groups meet in shared memory, then warp g finishes rows 2g, 2g + 1.
constexpr int kCombineRows = 32, kCombineGroups = 16;
static __global__ void __launch_bounds__(kCombineRows * kCombineGroups)
bow_combine_kernel(const float4* __restrict__ part4, const float2* __restrict__ part2, int ntp, int Npad, int N,
                   int V, int ldk, float c, const float* __restrict__ wsum, const float* __restrict__ rm,
                   float4* __restrict__ rowc, double* __restrict__ lossacc,
                   int* __restrict__ row_flag, int* __restrict__ flag_list, int* __restrict__ counters) {
  __shared__ float sh[kCombineGroups][6][kCombineRows];
  __shared__ double sh_loss[kCombineGroups];
  const int rl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int r = blockIdx.x * kCombineRows + rl;          // r < Npad always (Npad is a multiple of 128)
  {
    float M = -INFINITY, T1 = 0.f, T2 = 0.f, T3 = 0.f, T4 = 0.f, mn = INFINITY;
#pragma unroll 4
    for (int t = g; t < ntp; t += kCombineGroups) {
      const float4 a = part4[size_t(t) * Npad + r];
      const float2 b = part2[size_t(t) * Npad + r];
      if (a.y > 0.f) {                 // an unwritten (zeroed) slot or an all-padding part has T1 = 0
        mn = fminf(mn, b.y);
        if (a.x > M) {          // rescale the running sums to the new reference
          const float w = ex2_approx(M - a.x), w2 = w * w;
          T1 *= w; T2 *= w2; T3 *= w2 * w; T4 *= w2 * w2;
          M = a.x;
        }
        const float w = ex2_approx(a.x - M), w2 = w * w;
        T1 = fmaf(a.y, w, T1);
        T2 = fmaf(a.z, w2, T2);
        T3 = fmaf(a.w, w2 * w, T3);
        T4 = fmaf(b.x, w2 * w2, T4);
      }
    }
    sh[g][0][rl] = M; sh[g][1][rl] = T1; sh[g][2][rl] = T2; sh[g][3][rl] = T3; sh[g][4][rl] = T4; sh[g][5][rl] = mn;
  }
  __syncthreads();
  const int lane = rl;
  const float bsum = __ldg(wsum + ldk);
  double f = 0.0;
  for (int q = 0; q < kCombineRows / kCombineGroups; ++q) {
    const int row_l = g * (kCombineRows / kCombineGroups) + q;
    const int row = blockIdx.x * kCombineRows + row_l;
    if (row >= N) {
      if (lane == 0) {                       // padding rows read by the backward kernels
        rowc[row] = make_float4(0.f, 0.f, 0.f, 0.f);
        rowc[Npad + row] = make_float4(0.f, 0.f, 0.f, 0.f);
        row_flag[row] = 0;
      }
      continue;
    }
    // lanes 0 .. groups-1 hold the partial groups
    float M = -INFINITY, T1 = 0.f, T2 = 0.f, T3 = 0.f, T4 = 0.f, mn = INFINITY;
    if (lane < kCombineGroups) {
      M = sh[lane][0][row_l]; T1 = sh[lane][1][row_l]; T2 = sh[lane][2][row_l]; T3 = sh[lane][3][row_l];
      T4 = sh[lane][4][row_l]; mn = sh[lane][5][row_l];
    }
    const float Mx = warp_max(M);
    const float w = (T1 > 0.f) ? ex2_approx(M - Mx) : 0.f, w2 = w * w;
    T1 = warp_sum(T1 * w); T2 = warp_sum(T2 * w2); T3 = warp_sum(T3 * w2 * w); T4 = warp_sum(T4 * w2 * w2);
    mn = -warp_max(-mn);
    if (lane == 0) {
      const float rmi = __ldg(rm + row);
      const float lse2 = Mx + log2f(T1);                 // the partials carry log2-domain maxima / minima
      const float lse = lse2 * kLn2;
      // sum_j logit_ij = z_i . sum_j w_j + sum_j b_j = rm_i V ln 2 + bsum
      const double sx = double(rmi) * double(V) * double(kLn2) + double(bsum);
      const float i1 = 1.f / T1, i2 = i1 * i1;
      const float S2 = T2 * i2, S3 = T3 * i2 * i1, S4 = T4 * i2 * i2;
      // max_j p_ij <= S4^(1/4): S4 < 2^-20 guarantees every p < 2^-5 (the partial maxima are only references)
      const bool valid = (T1 > 0.f) && (S4 < 9.5367431640625e-07f) && (mn - lse2 > -27.f * kLog2e);
      float u = 0.f;
      if (valid) {
        f += double(c) * (sx - double(V) * double(lse)) -
             double(1.f - c) * (1.0 + double(0.5f * S2 + (1.f / 3.f) * S3 + 0.25f * S4));
        u = (1.f + S2 + S3 + S4) - c * (float(V) + 1.f + S2 + S3);
      } else {
        flag_list[atomicAdd(counters, 1)] = row;
      }
      row_flag[row] = valid ? 0 : 1;
      float* rc = reinterpret_cast<float*>(rowc + row);
      rc[2] = -lse2;
      rc[3] = -u * kYgScale;
      // S4 < 2^-32 certifies every p < 2^-8: the whole row takes the backward kernels' unchecked polynomial path
      // (quadratic in p: truncation p^2 / (1 - c - S) <= 1.5e-4 of an element's own dlogit at p = 2^-8, ~1e-8 typically)
      rowc[Npad + row] = make_float4(0.f, 0.f, 0.f, (valid && S4 < 2.3283064365386963e-10f) ? 1.f : 0.f);
    }
  }
  if (lane == 0) sh_loss[g] = f;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kCombineGroups; ++i) t += sh_loss[i];
    if (t != 0.0) atomicAdd(lossacc, t);
  }
}

// exact per-element terms (rows / chunks that leave the series' range)
static CAREL_DEVINL void bow_exact_terms(float x, float c, float& f, float& u) {
  const float p = __expf(x);
  const float A = fmaxf(x, -100.f);
  const float B = fmaxf(log1pf(-p), -100.f);
  f = c * A + (1.f - c) * B;
  u = p * (p - c) / fmaxf(p * (1.f - p), 1e-12f);
}

template <class T>
static CAREL_DEVINL T block_reduce_256(T v, T* sh, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (t > v ? t : v) : v + t;
  }
  __syncthreads();                 // sh may still be read from the previous reduction
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  T r = sh[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) r = is_max ? (sh[i] > r ? sh[i] : r) : r + sh[i];
  return r;
}

// Exact redo of the flagged rows, one block per row at a time, straight from the stored logits: the row's lse is
// recomputed from the ROUNDED logits (so that p_ij here and in the backward kernels are bit-identical and sum to 1:
// for p -> 1 the gradient u_j - p_j S_i cancels catastrophically and is only right if S_i and u_j come from the
// same p), then the dense loss term and S_i are summed per element.  Costs V * 2 bytes per flagged row.
static __global__ void __launch_bounds__(256)
bow_exact_rows_kernel(const uint16_t* __restrict__ Y, int ldy, int V, float c, const float* __restrict__ bl2e,
                      const int* __restrict__ flag_list, const int* __restrict__ counters,
                      float4* __restrict__ rowc, double* __restrict__ lossacc) {
  __shared__ float shf[8];
  __shared__ double shd[8];
  const int nflag = counters[0];
  double loss_blk = 0.0;
  for (int e = blockIdx.x; e < nflag; e += gridDim.x) {
    const int row = flag_list[e];
    const uint16_t* yr = Y + size_t(row) * ldy;
    const float4 rc4 = rowc[row];
    const float inv_s = rc4.x, bm = rc4.y;
    auto logit2 = [&](int j) { return fmaf(y16_one(yr[j]) - bm, inv_s, __ldg(bl2e + j)); };    // log2-domain logit - rm_i
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < V; j += 256) mx = fmaxf(mx, logit2(j));
    mx = block_reduce_256<float>(mx, shf, true);
    float s = 0.f;
    for (int j = threadIdx.x; j < V; j += 256) s += exp2f(logit2(j) - mx);
    s = block_reduce_256<float>(s, shf, false);
    const float rc = -(mx + log2f(s));                  // log2 p_ij = logit2_ij + rc
    float fs = 0.f, us = 0.f;
    for (int j = threadIdx.x; j < V; j += 256) {
      const float x = (fmaf(y16_one(yr[j]) - bm, inv_s, __ldg(bl2e + j) + rc)) * kLn2;   // same association as the backward kernels
      float f, u;
      bow_exact_terms(x, c, f, u);
      fs += f;
      us += u;
    }
    const double ft = block_reduce_256<double>(double(fs), shd, false);
    us = block_reduce_256<float>(us, shf, false);
    __syncthreads();                 // every thread has read rowc[row] before it is rewritten
    if (threadIdx.x == 0) {
      rowc[row] = make_float4(inv_s, bm, rc, -us * kYgScale);
      loss_blk += ft;
    }
  }
  if (threadIdx.x == 0 && loss_blk != 0.0) atomicAdd(lossacc, loss_blk);
}

// ---------------------------------------------------------------------------------------------
// sparse corrections (forward): one warp per row over its ELL entries (or a dense scan if it overflowed), one
// gather from Y per non-zero:   x = ((Y16_ij - bmagic_i) / s_i + bl2e_j - lse2_i) ln 2 = log p_ij
//   lossacc[1] += sum (1-ls) t (A - B) ;  S_i += sum -(1-ls) t p / max(p(1-p),1e-12) ;  ell_ev[slot] = that term
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
bow_sparse_fwd_kernel(const uint16_t* __restrict__ Y, int ldy, const float* __restrict__ bl2e,
                      float4* __restrict__ rowc, const int* __restrict__ ell_cols,
                      const float* __restrict__ ell_vals, const int* __restrict__ ell_nnz, float* __restrict__ ell_ev,
                      const float* __restrict__ bow_dense, int64_t ld_bow, int N, int V, int cap, float one_minus_ls,
                      float c, int Npad, double* __restrict__ lossacc, int* __restrict__ counters) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double loss_warp = 0.0;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    int nnz = ell_nnz[r];
    const uint16_t* yr = Y + size_t(r) * ldy;
    const float4 rc4 = rowc[r];
    float dloss = 0.f, du = 0.f;
    auto one = [&](int j, float t, float& ev) {
      const float x = fmaf(y16_one(yr[j]) - rc4.y, rc4.x, __ldg(bl2e + j) + rc4.z) * kLn2;
      const float p = __expf(x);
      ev = -one_minus_ls * t * p / fmaxf(p * (1.f - p), 1e-12f);
      const float A = fmaxf(x, -100.f), B = fmaxf(log1pf(-p), -100.f);
      dloss += one_minus_ls * t * (A - B);
      du += ev;
    };
    if (nnz > cap && bow_dense != nullptr) {
      const float* brow = bow_dense + size_t(r) * ld_bow;
      for (int j = lane; j < V; j += 32) {
        const float v = __ldg(brow + j);
        float ev;
        if (v != 0.f) one(j, v, ev);
      }
    } else {
      if (nnz > cap) {          // no dense target to fall back to: the list is truncated -> report it (loss = NaN)
        if (lane == 0) counters[1] = 1;
        nnz = cap;
      }
      for (int q = lane; q < nnz; q += 32) {
        const size_t slot = size_t(r) * cap + q;
        float ev;
        one(ell_cols[slot], ell_vals[slot], ev);
        ell_ev[slot] = ev;
      }
    }
    dloss = warp_sum(dloss);
    du = warp_sum(du);
    if (lane == 0) {
      loss_warp += double(dloss);
      // this warp owns rowc[r].w (= -S_i 2^10) here: the combine / exact kernels ran before
      const float sn = rc4.w - du * kYgScale;
      if (du != 0.f) reinterpret_cast<float*>(rowc + r)[3] = sn;
      // S_i is final now: the polynomial of the backward fast path (ygrad.cuh),
      //   d' = -c 2^10 + p' (a1 + p' (a2 + p' a3)),  p' = 2^(log2 p - rc),  P = 2^rc
      const float P = exp2f(rc4.z), S = -sn * (1.f / kYgScale);
      float4* pl = rowc + Npad + r;
      *pl = make_float4(kYgScale * (1.f - c - S) * P, kYgScale * (1.f - c) * P * P, kYgScale * P * P * P, pl->w);
    }
  }
  __shared__ double sh_l[8];
  if (lane == 0) sh_l[threadIdx.x >> 5] = loss_warp;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh_l[w];
    if (t != 0.0) atomicAdd(lossacc + 1, t);
  }
}

// ---------------------------------------------------------------------------------------------
// CSC view of the ELL list (counting sort by vocabulary index), so that the sparse part of dW is a
// gather per word instead of 256 fp32 atomics per non-zero.
// ---------------------------------------------------------------------------------------------
// Rows with nnz > cap are left to the overflow kernel when a dense target exists (`truncate` = 0); without one the
// first cap entries are what the forward used, and they are included here.
static __global__ void csc_count_kernel(const int* __restrict__ cols, const int* __restrict__ nnz, int N,
                                        int cap, int truncate, int* __restrict__ cnt) {
  // N * cap < 2^31 for every workspace this library carves (the ELL arrays are int-indexed)
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < unsigned(N) * unsigned(cap); t += gridDim.x * blockDim.x) {
    const int r = int(t / unsigned(cap)), q = int(t % unsigned(cap)), n = nnz[r];
    if ((n <= cap || truncate) && q < n) atomicAdd(cnt + cols[t], 1);
  }
}
// Exclusive scan of the per-word counts (one block; cnt may alias fillpos).  One pass per 32768 words (V = 30000: a
// single pass): coalesced loads into a padded shared-memory tile (index j + j/32: thread t then reads its 32
// consecutive counts from 32 different banks), thread-local scan in registers, the 1024 thread totals scanned with
// warp shuffles, coalesced write-out.  (Reading each thread's 32 counts straight from global memory made every warp
// load touch 32 different lines: 17 us on the one SM this kernel runs on.)
constexpr int kScanPer = 32, kScanChunk = 1024 * kScanPer;
constexpr int kScanSmemBytes = (kScanChunk + kScanChunk / 32) * int(sizeof(int));
static __global__ void __launch_bounds__(1024) csc_scan_kernel(const int* cnt, int V, int* colptr, int* fillpos) {
  constexpr int PER = kScanPer, CH = kScanChunk;
  extern __shared__ int scan_buf[];               // CH + CH / 32 ints
  __shared__ int wtot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto pad = [](int j) { return j + (j >> 5); };
  int carry = 0;                                  // every thread tracks the running total
  for (int base = 0; base < V; base += CH) {
    // eight independent 128-bit loads per thread (a plain strided loop waits for each load in turn: 32 dependent
    // round trips on the one SM this kernel runs on)
    const bool vec = (reinterpret_cast<uintptr_t>(cnt) & 15) == 0;
    int4 ld[PER / 4];
#pragma unroll
    for (int q = 0; q < PER / 4; ++q) {
      const int j = (q * 1024 + threadIdx.x) * 4;
      if (vec && base + j + 3 < V) {
        ld[q] = *reinterpret_cast<const int4*>(cnt + base + j);
      } else {
        ld[q].x = (base + j < V) ? cnt[base + j] : 0;
        ld[q].y = (base + j + 1 < V) ? cnt[base + j + 1] : 0;
        ld[q].z = (base + j + 2 < V) ? cnt[base + j + 2] : 0;
        ld[q].w = (base + j + 3 < V) ? cnt[base + j + 3] : 0;
      }
    }
#pragma unroll
    for (int q = 0; q < PER / 4; ++q) {
      const int j = (q * 1024 + threadIdx.x) * 4;          // j % 32 <= 28: the four words share one padding offset
      int* d = scan_buf + pad(j);
      d[0] = ld[q].x; d[1] = ld[q].y; d[2] = ld[q].z; d[3] = ld[q].w;
    }
    __syncthreads();
    int v[PER], s = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = s; s += scan_buf[pad(threadIdx.x * PER + i)]; }   // thread-local exclusive scan
    int incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int wv = wtot[lane];                          // every warp scans the 32 warp totals itself
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wv, off);
      if (lane >= off) wv += t;
    }
    const int before = __shfl_sync(0xffffffffu, wv, (warp + 31) & 31);
    const int run = carry + (warp > 0 ? before : 0) + incl - s;
    carry += __shfl_sync(0xffffffffu, wv, 31);
#pragma unroll
    for (int i = 0; i < PER; ++i) scan_buf[pad(threadIdx.x * PER + i)] = run + v[i];
    __syncthreads();
    for (int j = threadIdx.x; j < CH; j += 1024)
      if (base + j < V) { const int t = scan_buf[pad(j)]; colptr[base + j] = t; fillpos[base + j] = t; }
    __syncthreads();                              // the tile (and wtot) are reused by the next chunk
  }
  if (threadIdx.x == 0) colptr[V] = carry;
}
static __global__ void csc_fill_kernel(const int* __restrict__ cols, const int* __restrict__ nnz, int N,
                                       int cap, int truncate, int* __restrict__ fillpos, int* __restrict__ csc_row,
                                       int* __restrict__ csc_slot) {
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < unsigned(N) * unsigned(cap); t += gridDim.x * blockDim.x) {
    const int r = int(t / unsigned(cap)), q = int(t % unsigned(cap)), n = nnz[r];
    if ((n <= cap || truncate) && q < n) {
      const int pos = atomicAdd(fillpos + cols[t], 1);
      csc_row[pos] = r;
      csc_slot[pos] = int(t);
    }
  }
}


// sparse part of dZ: one warp per row, dz_r = scale * sum_q ev_q * w_{j_q}; writes grad_z[r, 0:K]
// (zeros for rows without non-zeros; rows that overflowed the list are written by bow_overflow_bwd_kernel when a dense
// target exists, truncated to the list otherwise — as the forward did).
static __global__ void bow_sparse_dz_kernel(const __half* __restrict__ WH, int ldk, int K,
                                            const int* __restrict__ cols, const int* __restrict__ nnz,
                                            const float* __restrict__ ev, int N, int cap, int truncate,
                                            const float* __restrict__ grad_out, float inv_nv,
                                            float* __restrict__ grad_z, int64_t ldgz) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float scale = __ldg(grad_out) * inv_nv;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    int n = nnz[r];
    if (n > cap) {
      if (!truncate) continue;
      n = cap;
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool act = lane * 8 < ldk;
    // the lanes fetch the (column, weight) pairs of the row once; the gather then keeps FOUR decoder rows in flight
    for (int base = 0; base < n; base += 32) {
      const int cnt = min(32, n - base);
      const int my_j = (lane < cnt) ? __ldg(cols + size_t(r) * cap + base + lane) : 0;
      const float my_e = (lane < cnt) ? __ldg(ev + size_t(r) * cap + base + lane) : 0.f;
      for (int q4 = 0; q4 < cnt; q4 += 4) {
        uint4 w[4];
        float e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int src = (q4 + u) & 31;
          const int j = __shfl_sync(0xffffffffu, my_j, src);
          e[u] = (q4 + u < cnt) ? __shfl_sync(0xffffffffu, my_e, src) : 0.f;
          w[u] = act ? __ldg(reinterpret_cast<const uint4*>(WH + size_t(j) * ldk) + lane) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t v[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = unpack_f16x2(v[h]);
            acc[2 * h] = fmaf(e[u], f.x, acc[2 * h]);
            acc[2 * h + 1] = fmaf(e[u], f.y, acc[2 * h + 1]);
          }
        }
      }
    }
    if (act) {
      float* gz = grad_z + size_t(r) * ldgz + lane * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (lane * 8 + i < K) gz[i] = acc[i] * scale;
    }
  }
}

// sparse part of [dW | db]: one warp per vocabulary word, gather over the rows that contain it.  This kernel
// INITIALISES gw[j, :] and gb[j] (plain stores, zeros for words nobody uses), so the dense passes that follow
// can accumulate into the caller's gradient tensors directly: no memset, no staging buffer, no scatter pass.
static __global__ void bow_sparse_dw_kernel(const __half* __restrict__ ZH, int ldk, int K,
                                            const int* __restrict__ colptr, const int* __restrict__ csc_row,
                                            const int* __restrict__ csc_slot, const float* __restrict__ ev,
                                            int V, const float* __restrict__ grad_out, float inv_nv,
                                            float* __restrict__ gw, float* __restrict__ gb) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float scale = __ldg(grad_out) * inv_nv;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < V; j += gridDim.x * wpb) {
    const int p0 = colptr[j], p1 = colptr[j + 1];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float esum = 0.f;
    const bool act = lane * 8 < ldk;
    for (int base = p0; base < p1; base += 32) {
      const int cnt = min(32, p1 - base);
      const int my_r = (lane < cnt) ? __ldg(csc_row + base + lane) : 0;
      const float my_e = (lane < cnt) ? __ldg(ev + __ldg(csc_slot + base + lane)) : 0.f;
      for (int q4 = 0; q4 < cnt; q4 += 4) {
        uint4 z[4];
        float e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int src = (q4 + u) & 31;
          const int r = __shfl_sync(0xffffffffu, my_r, src);
          e[u] = (q4 + u < cnt) ? __shfl_sync(0xffffffffu, my_e, src) : 0.f;
          z[u] = act ? __ldg(reinterpret_cast<const uint4*>(ZH + size_t(r) * ldk) + lane) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          esum += e[u];
          const uint32_t v[4] = {z[u].x, z[u].y, z[u].z, z[u].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = unpack_f16x2(v[h]);
            acc[2 * h] = fmaf(e[u], f.x, acc[2 * h]);
            acc[2 * h + 1] = fmaf(e[u], f.y, acc[2 * h + 1]);
          }
        }
      }
    }
    float* dw = gw + size_t(j) * K;
    if (act) {
      if (lane * 8 + 8 <= K && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) {
        reinterpret_cast<float4*>(dw + lane * 8)[0] = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
        reinterpret_cast<float4*>(dw + lane * 8)[1] = make_float4(acc[4] * scale, acc[5] * scale, acc[6] * scale, acc[7] * scale);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (lane * 8 + i < K) dw[lane * 8 + i] = acc[i] * scale;
      }
    }
    if (lane == 0) gb[j] = esum * scale;
  }
}

// Backward of the rows that overflowed the ELL list (dense target required): the row is scanned again, every
// non-zero handled by the whole warp (lane k owns feature columns k, k + 32, ...): dz accumulates in registers and
// is written once (this kernel owns grad_z[r, :] of those rows), dW / db go through atomics.  Rare by construction.
static __global__ void __launch_bounds__(256)
bow_overflow_bwd_kernel(const uint16_t* __restrict__ Y, int ldy, const float* __restrict__ bl2e,
                        const float4* __restrict__ rowc, const __half* __restrict__ ZH, const __half* __restrict__ WH,
                        int ldk, int K, const int* __restrict__ ell_nnz, const float* __restrict__ bow_dense,
                        int64_t ld_bow, int N, int V, int cap, float one_minus_ls, const float* __restrict__ grad_out,
                        float inv_nv, float* __restrict__ grad_z, int64_t ldgz, float* __restrict__ gw,
                        float* __restrict__ gb) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float scale = __ldg(grad_out) * inv_nv;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    if (ell_nnz[r] <= cap) continue;
    const uint16_t* yr = Y + size_t(r) * ldy;
    const __half* zr = ZH + size_t(r) * ldk;
    const float4 rc4 = __ldg(rowc + r);
    const float* brow = bow_dense + size_t(r) * ld_bow;
    float zreg[8], dz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      zreg[i] = (k < K) ? __half2float(zr[k]) : 0.f;
      dz[i] = 0.f;
    }
    for (int base = 0; base < V; base += 32) {
      const int jj = base + lane;
      const float v = (jj < V) ? __ldg(brow + jj) : 0.f;
      unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int j = base + src;
        const float t = __shfl_sync(0xffffffffu, v, src);
        const float x = fmaf(y16_one(yr[j]) - rc4.y, rc4.x, __ldg(bl2e + j) + rc4.z) * kLn2;
        const float p = __expf(x);
        const float e = -one_minus_ls * t * p / fmaxf(p * (1.f - p), 1e-12f) * scale;
        const __half* wr = WH + size_t(j) * ldk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = lane + 32 * i;
          if (k < K) {
            dz[i] = fmaf(e, __half2float(wr[k]), dz[i]);
            atomicAdd(gw + size_t(j) * K + k, e * zreg[i]);
          }
        }
        if (lane == 0) atomicAdd(gb + j, e);
      }
    }
    float* gz = grad_z + size_t(r) * ldgz;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      if (k < K) gz[k] = dz[i];
    }
  }
}

static __global__ void bow_loss_finalize_kernel(const double* __restrict__ lossacc, const int* __restrict__ counters,
                                                double inv_nv, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    // a truncated non-zero list (ELL-only target with a row longer than the list) has no meaningful loss
    out[0] = counters[1] ? __int_as_float(0x7fc00000) : float(-(lossacc[0] + lossacc[1]) * inv_nv);
  }
}

}  // namespace carel

using namespace carel;

extern "C" size_t carel_bow_workspace_bytes(int64_t N, int64_t K, int64_t V, int64_t ell_cap) {
  return carve_bow(nullptr, N, K, V, ell_cap).bytes;
}

// Dense [N, V] target -> ELL arrays (also done implicitly by carel_bow_fwd when it gets a dense target only).
extern "C" int carel_bow_pack_ell(const float* bow_dense, int64_t ld_bow, int64_t N, int64_t V, int64_t ell_cap,
                                  int32_t* ell_cols, float* ell_vals, int32_t* ell_nnz, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && V > 0 && ell_cap > 0, "carel_bow_pack_ell: empty problem");
  CAREL_REQUIRE(bow_dense && ell_cols && ell_vals && ell_nnz, "carel_bow_pack_ell: null argument");
  bow_pack_ell_kernel<8, 256, 1><<<int(std::min<int64_t>((N + 7) / 8, sm_count() * 32)), 256, 0, st>>>(
      bow_dense, ld_bow, int(N), int(V), int(ell_cap), ell_cols, ell_vals, ell_nnz);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_bow_fwd(const float* z, int64_t ldz, const float* dec_w, const float* dec_b,
                             const float* bow_dense, int64_t ld_bow, const int32_t* ell_cols_in,
                             const float* ell_vals_in, const int32_t* ell_nnz_in, int64_t N,
                             int64_t K, int64_t V, int64_t ell_cap, float label_smoothing,
                             float* loss_out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(N > 0 && K > 0 && V > 0 && ell_cap > 0, "carel_bow_fwd: empty problem");
  CAREL_REQUIRE(N * ell_cap < (int64_t(1) << 31), "carel_bow_fwd: N * ell_cap must stay below 2^31 (int-indexed ELL / CSC arrays)");
  CAREL_REQUIRE(K <= 256, "carel_bow_fwd: decoder input width (2 * ec_dim) must be <= 256, got %lld", (long long)K);
  CAREL_REQUIRE(bow_dense != nullptr || (ell_cols_in && ell_vals_in && ell_nnz_in), "carel_bow_fwd: need a dense or an ELL target");
  BowWs w = carve_bow(ws, N, K, V, ell_cap);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_fwd: workspace too small");
  const int cap = int(ell_cap), sms = sm_count();
  const int truncate = bow_dense == nullptr ? 1 : 0;
  // ---- sparse target, on the side branch when there is one (api_common.cuh): the dense -> ELL scan (or the copy of the
  //      caller's list) and the CSC view of the list for the backward gathers depend on the target only, so they run
  //      beside the operand packing and whatever the dense pass leaves free; the list is joined in front of the sparse
  //      corrections, the CSC view at the end of this call ----
  SideBranch* sb = (!timing_enabled() && (side_branch_mask() & 1)) ? side_branch() : nullptr;   // timing passes measure kernels alone
  cudaStream_t ss = sb != nullptr ? sb->s : st;
  std::unique_lock<std::mutex> sb_lock;      // two host threads driving the same device take turns on the branch
  if (sb != nullptr) sb_lock = std::unique_lock<std::mutex>(sb->mu);
  if (sb != nullptr) CAREL_CHECK_CUDA(sb->fork(st));
  if (ell_cols_in != nullptr) {
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_cols, ell_cols_in, size_t(N) * cap * 4, cudaMemcpyDeviceToDevice, ss));
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_vals, ell_vals_in, size_t(N) * cap * 4, cudaMemcpyDeviceToDevice, ss));
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_nnz, ell_nnz_in, size_t(N) * 4, cudaMemcpyDeviceToDevice, ss));
  } else {
    bow_pack_ell_kernel<8, 256, 1><<<int(std::min<int64_t>((N + 7) / 8, sms * 32)), 256, 0, ss>>>(bow_dense, ld_bow, int(N), int(V), cap, w.ell_cols, w.ell_vals, w.ell_nnz);
    CAREL_LAUNCHED();
  }
  auto csc_view = [&]() -> int {      // CSC view of the non-zero list for the backward gather (cheap: N*cap entries)
    CAREL_CHECK_CUDA(cudaMemsetAsync(w.fillpos, 0, (size_t(V) + 1) * 4, ss));
    csc_count_kernel<<<sms * 2, 256, 0, ss>>>(w.ell_cols, w.ell_nnz, int(N), cap, truncate, w.fillpos);
    CAREL_LAUNCHED();
    {
      static const cudaError_t attr = cudaFuncSetAttribute(csc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           kScanSmemBytes);
      CAREL_CHECK_CUDA(attr);
    }
    csc_scan_kernel<<<1, 1024, kScanSmemBytes, ss>>>(w.fillpos, int(V), w.colptr, w.fillpos);
    CAREL_LAUNCHED();
    csc_fill_kernel<<<sms * 2, 256, 0, ss>>>(w.ell_cols, w.ell_nnz, int(N), cap, truncate, w.fillpos, w.csc_row, w.csc_slot);
    CAREL_LAUNCHED();
    return 0;
  };
  if (sb != nullptr) {
    CAREL_CHECK_CUDA(cudaEventRecord(sb->ev_join[0], ss));       // the list is complete here
    if (int rc = csc_view()) return rc;
  }
  // ---- operands: W_d -> fp16, its column sums / bias constants, then Z -> fp16 with the row references ----
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.wsum, 0, (size_t(w.ldk) + 4) * 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.lossacc, 0, 2 * sizeof(double), st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 4 * sizeof(int), st));
  unsigned* rmax_bits = reinterpret_cast<unsigned*>(w.wsum + w.ldk + 1);
  bow_pack_f16_kernel<<<int(std::min<int64_t>((V + 7) / 8, sms * 8)), 256, 0, st>>>(dec_w, int(V), int(K), K, w.WH, w.ldk,
                                                                                   nullptr, 0.f, nullptr, nullptr, nullptr, rmax_bits);
  CAREL_LAUNCHED();
  {
    const dim3 blk((w.ldk / 8 + 7) / 8 * 8, 16);
    bow_colsum_bias_kernel<<<sms * 2, blk, size_t(blk.y) * w.ldk * sizeof(float), st>>>(
        w.WH, int(V), w.ldk, w.wsum, dec_b, int(round_up64(V, 128)), w.bl2e);
  }
  CAREL_LAUNCHED();
  bow_pack_f16_kernel<<<int((N + 7) / 8), 256, 0, st>>>(z, int(N), int(K), ldz, w.ZH, w.ldk, w.wsum, 1.f / float(V), w.rm,
                                                      w.rowenc, w.rowc, rmax_bits);
  CAREL_LAUNCHED();
  // ---- pass F: the one dense pass (online max, power sums, stored logits) + per-row combine ----
  CUtensorMap ta, tb;
  CAREL_REQUIRE(make_tmap_f16_sw128(&ta, w.ZH, N, w.ldk, w.ldk, 128) == 0, "carel_bow_fwd: tensor map Z failed");
  CAREL_REQUIRE(make_tmap_f16_sw128(&tb, w.WH, V, w.ldk, w.ldk, 128) == 0, "carel_bow_fwd: tensor map W failed");
  AsSched as = make_as_sched(N, V, w.ldk, sms);
  EpiBowFusedParams e1;
  CAREL_REQUIRE(make_tmap_2b(&e1.tmY, w.Y, N, V, w.ldy, 32, 32, true) == 0, "carel_bow_fwd: tensor map Y failed");
  e1.part4 = w.part4; e1.part2 = w.part2; e1.bl2e = w.bl2e; e1.rowenc = w.rowenc; e1.N = int(N); e1.V = int(V); e1.Npad = w.Npad;
  e1.n_tiles = as.n_tiles; e1.tpc = as.tpc;
  const int nparts = std::min(as.n_tiles, (as.n_tiles - 1) / as.tpc + 2) * kColSplit;      // <= w.ntp
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.part4, 0, size_t(nparts) * w.Npad * sizeof(float4), st));
  if (int rc = launch_gemm_as<EpiBowFused<256>, true, kBowEpiSmem>(ta, tb, as, e1, st, "bow_passF")) return rc;
  bow_combine_kernel<<<w.Npad / kCombineRows, kCombineRows * kCombineGroups, 0, st>>>(
      w.part4, w.part2, nparts, w.Npad, int(N), int(V), w.ldk, label_smoothing / float(V), w.wsum, w.rm, w.rowc,
      w.lossacc, w.row_flag, w.flag_list, w.counters);
  CAREL_LAUNCHED();
  // ---- exact redo of the flagged rows from the stored logits (a no-op grid when nothing is flagged) ----
  bow_exact_rows_kernel<<<sms * 4, 256, 0, st>>>(w.Y, w.ldy, int(V), label_smoothing / float(V), w.bl2e,
                                                w.flag_list, w.counters, w.rowc, w.lossacc);
  CAREL_LAUNCHED();
  // ---- sparse corrections (need the list) ----
  if (sb != nullptr) CAREL_CHECK_CUDA(cudaStreamWaitEvent(st, sb->ev_join[0], 0));
  bow_sparse_fwd_kernel<<<int(std::min<int64_t>((N + 7) / 8, sms * 16)), 256, 0, st>>>(
      w.Y, w.ldy, w.bl2e, w.rowc, w.ell_cols, w.ell_vals, w.ell_nnz, w.ell_ev, bow_dense, ld_bow, int(N), int(V), cap,
      1.f - label_smoothing, label_smoothing / float(V), w.Npad, w.lossacc, w.counters);
  CAREL_LAUNCHED();
  bow_loss_finalize_kernel<<<1, 32, 0, st>>>(w.lossacc, w.counters, 1.0 / (double(N) * double(V)), loss_out);
  CAREL_LAUNCHED();
  if (sb != nullptr) {
    CAREL_CHECK_CUDA(sb->join(st, 1));      // the CSC view (and with it everything on the side stream) is complete
  } else if (int rc = csc_view()) {
    return rc;
  }
  return 0;
}

// parts: 1 = everything that completes grad_w / grad_b (the O(nnz) gathers of both gradients, the overflow rows and the
// dense dW kernel), 2 = the dense dZ kernel, 3 = both.  A data-parallel caller runs part 1, starts the all-reduce of
// grad_w on another stream and runs part 2 underneath it (parallel.OverlappedGradReducer).
extern "C" int carel_bow_bwd_parts(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N,
                                   int64_t K, int64_t V, int64_t ell_cap, float label_smoothing,
                                   float* grad_z, int64_t ldgz, float* grad_w, float* grad_b, void* ws,
                                   size_t ws_bytes, int parts, void* stream);
extern "C" int carel_bow_bwd(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N,
                             int64_t K, int64_t V, int64_t ell_cap, float label_smoothing,
                             float* grad_z, int64_t ldgz, float* grad_w, float* grad_b, void* ws,
                             size_t ws_bytes, void* stream) {
  return carel_bow_bwd_parts(grad_out, bow_dense, ld_bow, N, K, V, ell_cap, label_smoothing, grad_z, ldgz, grad_w, grad_b,
                             ws, ws_bytes, 3, stream);
}
extern "C" int carel_bow_bwd_parts(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N,
                                   int64_t K, int64_t V, int64_t ell_cap, float label_smoothing,
                                   float* grad_z, int64_t ldgz, float* grad_w, float* grad_b, void* ws,
                                   size_t ws_bytes, int parts, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CAREL_REQUIRE(parts >= 1 && parts <= 3, "carel_bow_bwd_parts: parts must be 1, 2 or 3");
  BowWs w = carve_bow(ws, N, K, V, ell_cap);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_bwd: workspace too small");
  CAREL_REQUIRE(grad_z != nullptr, "carel_bow_bwd: grad_z is required");
  const bool need_w = grad_w != nullptr || grad_b != nullptr;
  const int sms = sm_count();
  const int truncate = bow_dense == nullptr ? 1 : 0;
  // grad_z[r, 0:K] and (if wanted) grad_w / grad_b are fully written (not accumulated) by the sparse gather
  // kernels; the dense passes then add to them.  A gradient the caller does not want lands in scratch.
  float* gw = grad_w != nullptr ? grad_w : w.dWacc;
  float* gb = grad_b != nullptr ? grad_b : w.dbacc;
  const float inv_nv = float(1.0 / (double(N) * double(V)));
  if (parts & 1) {
  // the two O(nnz) gathers write disjoint outputs (grad_z | grad_w, grad_b): side by side when side branches are on (bit 1)
  SideBranch* sb = (need_w && !timing_enabled() && (side_branch_mask() & 2)) ? side_branch() : nullptr;
  cudaStream_t ss = sb != nullptr ? sb->s : st;
  std::unique_lock<std::mutex> sb_lock;
  if (sb != nullptr) sb_lock = std::unique_lock<std::mutex>(sb->mu);
  if (sb != nullptr) CAREL_CHECK_CUDA(sb->fork(st));
  bow_sparse_dz_kernel<<<int(std::min<int64_t>((N + 7) / 8, sms * 16)), 256, 0, ss>>>(
      w.WH, w.ldk, int(K), w.ell_cols, w.ell_nnz, w.ell_ev, int(N), int(ell_cap), truncate, grad_out, inv_nv, grad_z, ldgz);
  CAREL_LAUNCHED();
  if (need_w) {
    bow_sparse_dw_kernel<<<sms * 16, 256, 0, st>>>(w.ZH, w.ldk, int(K), w.colptr, w.csc_row, w.csc_slot, w.ell_ev,
                                                  int(V), grad_out, inv_nv, gw, gb);
    CAREL_LAUNCHED();
  }
  if (sb != nullptr) {
    CAREL_CHECK_CUDA(sb->join(st, 0));
    sb_lock.unlock();
  }
  if (bow_dense != nullptr) {
    bow_overflow_bwd_kernel<<<int(std::min<int64_t>((N + 7) / 8, sms * 16)), 256, 0, st>>>(
        w.Y, w.ldy, w.bl2e, w.rowc, w.ZH, w.WH, w.ldk, int(K), w.ell_nnz, bow_dense, ld_bow, int(N), int(V),
        int(ell_cap), 1.f - label_smoothing, grad_out, inv_nv, grad_z, ldgz, gw, gb);
    CAREL_LAUNCHED();
  }
  }
  // ---- dense part: two single-GEMM kernels over the stored logits (ygrad.cuh) ----
  const int k16 = w.ldk / 16;
  YGradParams yp;
  yp.rowc = w.rowc; yp.bl2e = w.bl2e; yp.grad_out = grad_out;
  yp.N = int(N); yp.V = int(V); yp.K = int(K); yp.Npad = w.Npad;
  yp.c = label_smoothing / float(V); yp.inv_nv = inv_nv;
  yp.probe = 0;
  if (const char* e = getenv("CAREL_YG_PROBE")) yp.probe = atoi(e);      // timing probes (ygrad.cuh), never set in product runs
  CUtensorMap ty, tb;
  if (need_w && (parts & 1)) {      // owner = 128 words, step = 128 clauses: dW += dlogit^T Z, db
    CAREL_REQUIRE(make_tmap_f16_sw128(&ty, w.Y, N, V, w.ldy, 128) == 0, "carel_bow_bwd: tensor map Y failed");
    CAREL_REQUIRE(make_tmap_f16_sw128(&tb, w.ZH, N, w.ldk, w.ldk, 64) == 0, "carel_bow_bwd: tensor map Z failed");
    int ys, bs;
    yg_stages(k16, true, &ys, &bs);
    ChainSched cs = make_chain_sched(0, int((V + 127) / 128), int((N + 127) / 128), k16, bs, 0);
    yp.out = gw; yp.gb = gb; yp.ldo = K; yp.y_stages = ys;
    if (int rc = launch_ygrad<true>(ty, tb, cs, yp, st, "bow_dw")) return rc;
  }
#ifdef CAREL_CHAIN_TRACE
  if (getenv("CAREL_TRACE_SKIP_DZ") != nullptr) return 0;      // debug: keep the dW kernel's timeline
#endif
  if (parts & 2) {   // owner = 256 clauses, step = 64 words: dZ += dlogit W_d
    CAREL_REQUIRE(make_tmap_f16_sw128(&ty, w.Y, N, V, w.ldy, 256) == 0, "carel_bow_bwd: tensor map Y failed");
    CAREL_REQUIRE(make_tmap_f16_sw128(&tb, w.WH, V, w.ldk, w.ldk, 64) == 0, "carel_bow_bwd: tensor map W failed");
    int ys, bs;
    yg_stages(k16, false, &ys, &bs);
    ChainSched cs = make_chain_sched(0, int((N + 255) / 256), int((V + 63) / 64), k16, bs, 0);
    yp.out = grad_z; yp.gb = nullptr; yp.ldo = ldgz; yp.y_stages = ys;
    if (int rc = launch_ygrad<false>(ty, tb, cs, yp, st, "bow_dz")) return rc;
  }
  return 0;
}

#ifdef CAREL_CHAIN_TRACE
// debug: the clock64 timeline block 0 of the last traced kernel recorded (8 slots x 256 steps)
extern "C" int carel_debug_chain_trace(long long* host_dst) {
  CAREL_CHECK_CUDA(cudaMemcpyFromSymbol(host_dst, carel::g_chain_trace, sizeof(long long) * 8 * 256));
  return 0;
}
#endif
