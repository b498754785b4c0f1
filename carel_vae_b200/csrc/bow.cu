// bow.cu — carel_bow_{fwd,bwd}: bag-of-words reconstruction decoder + loss
// (drl_classifier_ec_mmd_final_mul.py:253-254 decoder + Softmax, :381-387 get_reconstruct_loss):
//
//   p = softmax(Z W_d^T + b_d),  t = bow (1 - ls) + ls / V,
//   loss = -mean_{i,j} [ t max(log p, -100) + (1 - t) max(log(1 - p), -100) ]          (BCELoss)
//
// The [N, V] logits / probabilities never go to HBM.  With c = ls / V the loss splits into a
// DENSE part that depends on p only and a SPARSE correction on the non-zeros of bow:
//   t A + (1-t) B = c A + (1-c) B + (1-ls) bow (A - B),      A = max(log p,-100), B = max(log1p(-p),-100)
// so the GEMM epilogues never read the 1 GB dense target; it is packed once into an ELL list
// (HBM-bound pass) and the corrections are O(nnz) warp-per-row work.  Same split in backward:
//   g = (p - t) / max(p (1-p), 1e-12) / (N V)   (torch's BCELoss backward),  s_i = sum_j g_ij p_ij,
//   dlogit = p (g - s_i)  =  [u - p S_i] / (N V)  +  sparse,    u = p (p - c) / max(p (1-p), 1e-12).
//
// Operands: ZA = [z, -lse split(3), 1 1 1], WB = [w, 1 1 1, b split(3)] (bf16, pack_rows_kernel).  The tensor-core
// passes use the PLAIN columns [0, K) and add b_j log2(e) / -lse_i log2(e) in the epilogue (K = 256 is four full
// 64-wide TMA boxes); the augmented columns serve the O(nnz) sparse kernels, whose dot product ZA_i . WB_j is log p_ij.
//   pass F  gemm_as + EpiBowFused  ONE dense pass: per-row running reference + power sums -> lse, dense loss, S_i
//           (bow_combine_kernel); rows outside the series' range are flagged and redone exactly by
//   pass 2  gemm_as + EpiBowLoss   (skips every row tile without a flagged row: normally all of them)
//   sparse  bow_sparse_kernel      corrections of loss and S_i (fwd); bow_sparse_dz / bow_sparse_dw gathers (bwd),
//                                  which also initialise grad_z / grad_w / grad_b
//   pass 3  chain_tc + PolBowBwdW  owner = vocabulary tile: dlogit tile -> bf16 (TMEM only), dW += dlogit^T Z, db
//   pass 4  chain_tc + PolBowBwdZ  owner = clause tile:     dlogit tile recomputed,          dZ += dlogit W_d
#include <algorithm>

#include "api_common.cuh"
#include "chain_launch.cuh"
#include "chain_policies.cuh"
#include "epilogues.cuh"
#include "gemm_launch.cuh"
#include "pack.cuh"

namespace carel {

constexpr float kTinyLogP = -6.9314718f;   // log(2^-10): below this the 3-term series are exact to 2^-30

struct BowWs {
  __nv_bfloat16* ZA;    // [N, ldk]
  __nv_bfloat16* WB;    // [V, ldk]
  float4* part4;        // [ntp, Npad] pass-F partials (m, T1, T2, T3)
  float2* part2;        // [ntp, Npad] pass-F partials (T4, min)
  float* wsum;          // [ldk] column sums of WB
  int* row_flag;        // [Npad] 1: row needs the exact pass
  int* tile_flag;       // [Npad/128]
  float* lse;           // [Npad]  holds -lse
  float* nlse2;         // [Npad]  -lse log2(e) (0 beyond N)
  float* bl2e;          // [V rounded up to 64]  b log2(e) (0 beyond V)
  float* srow;          // [Npad]  S_i = N V s_i
  float* dWacc;         // [V, ldk] scratch target of dW when the caller does not want grad_w
  float* dbacc;         // [V]      scratch target of db when the caller does not want grad_b
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
  int ldk, Np, Npad, ntp;
};

static BowWs carve_bow(void* ws, int64_t N, int64_t K, int64_t V, int64_t cap) {
  WsCarver w(ws);
  BowWs r;
  r.ldk = aug_width(int(K));
  r.Np = int(round_up64(N, 64));
  r.Npad = int(round_up64(N, 128));
  r.ntp = int((V + 255) / 256) * kColSplit;
  r.ZA = w.take<__nv_bfloat16>(size_t(N) * r.ldk);
  r.WB = w.take<__nv_bfloat16>(size_t(V) * r.ldk);
  r.part4 = w.take<float4>(size_t(r.ntp) * r.Npad);
  r.part2 = w.take<float2>(size_t(r.ntp) * r.Npad);
  r.wsum = w.take<float>(size_t(r.ldk));
  r.row_flag = w.take<int>(size_t(r.Npad));
  r.tile_flag = w.take<int>(size_t(r.Npad / 128) + 1);   // [0] = any, [1 + tile]
  r.lse = w.take<float>(size_t(r.Npad));
  r.nlse2 = w.take<float>(size_t(r.Npad));
  r.bl2e = w.take<float>(size_t(round_up64(V, 64)));
  r.srow = w.take<float>(size_t(r.Npad));
  r.dWacc = w.take<float>(size_t(V) * r.ldk);
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
// pass F epilogue (the ONE dense forward pass): per row and per 64-column part, with y = log2-domain logit
// (acc log2 e + b_j log2 e) and m = a running reference (moved when a chunk maximum exceeds it by more than 2^8):
//   T_k = sum_j 2^(k (y_ij - m)), k = 1..4,   and   mn = min_j y_ij.
// From these the combine kernel gets lse_i, the power sums S_k = sum_j p_ij^k = T_k / T_1^k and with them
//   sum_j log1p(-p) = -(S1 + S2/2 + S3/3 + S4/4 + ...),   sum_j (p - c)/(1 - p) = (S1+S2+S3+S4+..) - c (V + S1+S2+S3+..)
// (S1 = 1), valid to 2^-20 whenever max_j p_ij < 2^-5 (certified by S4 < 2^-20); rows beyond that (or with a
// clamped log) are flagged and redone exactly by the EpiBowLoss pass.  Hot loop per PAIR of elements: FFMA2 (y),
// FADD2 (y - m), 2 MUFU.EX2 issued before the chunk's max / min chains, FMUL2, 2 FADD2, 2 FFMA2, 2 x 1/2 FMNMX3.
// ---------------------------------------------------------------------------------------------
//
// The A-stationary kernel walks consecutive column tiles of ONE row tile, so a thread meets the same row again in
// the next tile: the running state (m, T1..T4, mn) stays in registers across the tiles of a row segment and is
// written once per segment (seg = this CTA's index among the CTAs that share the row tile, <= (n_tiles-1)/tpc + 2
// of them) instead of once per column tile: 472 -> 16 partials per row at V = 30000.
struct EpiBowFusedParams {
  float4* part4;   // [nseg * kColSplit][Npad]  (m, T1, T2, T3)     m, mn in log2 units; zeroed before the pass
  float2* part2;   // [nseg * kColSplit][Npad]  (T4, mn)
  const float* bl2e;   // [Vpad]  b_j log2(e)
  int N, V, Npad;
  int n_tiles, tpc;    // of the AsSched the pass runs under (m_tile0 = 0, full grid)
};
template <int BN>
struct EpiBowFused {
  using Params = EpiBowFusedParams;
  const Params& p;
  float m, mn;
  float2 t1, t2, t3, t4;
  CAREL_DEVINL EpiBowFused(const Params& p_, const TileCtx&) : p(p_) {}
  CAREL_DEVINL void prefetch(const TileCtx&) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int c0 = ctx.col0 + ctx.half * HALF;
    if (ctx.first_of_row) {
      m = -INFINITY; mn = INFINITY;
      t1 = make_float2(0.f, 0.f); t2 = t1; t3 = t1; t4 = t1;
    }
    if (c0 < p.V) {
      const bool full = c0 + HALF <= p.V;
      const float2 l2e = make_float2(kLog2e, kLog2e);
      const float4* cc4 = reinterpret_cast<const float4*>(p.bl2e + c0);
      for_each_chunk16(ctx.taddr + ctx.half * HALF, HALF / 16, [&](int c, const uint32_t (&r)[16]) {
        // y = log2-domain logit: acc log2(e) + b_j log2(e)   (plain operands; the bias is a per-column constant)
        float y[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = __ldg(cc4 + c * 4 + i);
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
  CAREL_DEVINL void finish() {}
};

// Column sums of the packed decoder operand WB [V, ldk] (bf16 -> fp32: sum_j x_ij = ZA_i . wsum) and
// bl2e[j] = b_j log2(e) (zero padded to a multiple of 64) for the plain-operand passes, in one launch.
// Block = (ldk / 8 chunks of 8 bf16) x 16 row lanes, 128-bit loads; one atomicAdd per column and block.
static __global__ void bow_colsum_bias_kernel(const __nv_bfloat16* __restrict__ WB, int V, int ldk,
                                              float* __restrict__ wsum, const float* __restrict__ b, int Vpad,
                                              float* __restrict__ bl2e) {
  extern __shared__ float sh_cs[];           // [blockDim.y][ldk]
  const int nx = ldk >> 3;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  for (int j = blockIdx.x * nthreads + tid; j < Vpad; j += gridDim.x * nthreads)
    bl2e[j] = (j < V) ? __ldg(b + j) * kLog2e : 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (threadIdx.x < nx) {
    for (int r = blockIdx.x * blockDim.y + threadIdx.y; r < V; r += gridDim.x * blockDim.y) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(WB + size_t(r) * ldk) + threadIdx.x);
      const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[h]));
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
}

static CAREL_DEVINL void split3(float h, __nv_bfloat16& h1, __nv_bfloat16& h2, __nv_bfloat16& h3) {
  h1 = __float2bfloat16_rn(h);
  const float r1 = h - __bfloat162float(h1);
  h2 = __float2bfloat16_rn(r1);
  h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
}

// Combine the pass-F partials of the rows: lse_i, power sums, dense loss term and S_i = sum_j u_ij; writes
// -lse (3-way bf16 split) into the h columns of ZA.  Rows outside the series' range are flagged (row_flag /
// tile_flag[1 + tile], tile_flag[0] = any) and contribute nothing here.
// Block = 32 rows x 16 partial groups: thread (r, g) folds partials g, g + 16, ... of row r online (consecutive
// threads read consecutive rows: coalesced), the 8 groups meet in shared memory, then warp w finishes rows
// 2w, 2w + 1 (sum_j x_ij = ZA_i . wsum needs a warp per row).
constexpr int kCombineRows = 32, kCombineGroups = 16;
static __global__ void __launch_bounds__(kCombineRows * kCombineGroups)
bow_combine_kernel(const float4* __restrict__ part4, const float2* __restrict__ part2, int ntp, int Npad, int N,
                   int V, int K, int ldk, float c, const float* __restrict__ wsum, __nv_bfloat16* __restrict__ ZA,
                   float* __restrict__ neg_lse, float* __restrict__ nlse2, float* __restrict__ srow,
                   double* __restrict__ lossacc, int* __restrict__ row_flag, int* __restrict__ tile_flag) {
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
  // warp g finishes rows g * (32 / groups) ..
  const int lane = rl;
  double f = 0.0;
  for (int q = 0; q < kCombineRows / kCombineGroups; ++q) {
    const int row_l = g * (kCombineRows / kCombineGroups) + q;
    const int row = blockIdx.x * kCombineRows + row_l;
    if (row >= N) {
      if (lane == 0) { nlse2[row] = 0.f; srow[row] = 0.f; }      // padding rows read by the backward chains
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
    // sum_j x_ij from the column sums of the packed operand (the h columns of ZA are still zero here)
    float sx = 0.f;
    const __nv_bfloat16* zrow = ZA + size_t(row) * ldk;
    for (int k = lane; k < ldk; k += 32) sx = fmaf(__bfloat162float(zrow[k]), __ldg(wsum + k), sx);
    sx = warp_sum(sx);
    if (lane == 0) {
      const float lse2 = Mx + log2f(T1);                 // the partials carry log2-domain maxima / minima
      const float lse = lse2 * kLn2;
      const float i1 = 1.f / T1, i2 = i1 * i1;
      const float S2 = T2 * i2, S3 = T3 * i2 * i1, S4 = T4 * i2 * i2;
      // max_j p_ij <= S4^(1/4): S4 < 2^-20 guarantees every p < 2^-5 (the partial maxima are only references)
      const bool valid = (T1 > 0.f) && (S4 < 9.5367431640625e-07f) && (mn - lse2 > -27.f * kLog2e);
      float u = 0.f;
      if (valid) {
        f += double(c) * (double(sx) - double(V) * double(lse)) -
             double(1.f - c) * (1.0 + double(0.5f * S2 + (1.f / 3.f) * S3 + 0.25f * S4));
        u = (1.f + S2 + S3 + S4) - c * (float(V) + 1.f + S2 + S3);
      } else {
        tile_flag[0] = 1;
        tile_flag[1 + (row >> 7)] = 1;
      }
      row_flag[row] = valid ? 0 : 1;
      srow[row] = u;
      const float nl = -lse;
      neg_lse[row] = nl;
      nlse2[row] = -lse2;
      __nv_bfloat16 h1, h2, h3;
      split3(nl, h1, h2, h3);
      ZA[size_t(row) * ldk + K] = h1;
      ZA[size_t(row) * ldk + K + 1] = h2;
      ZA[size_t(row) * ldk + K + 2] = h3;
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

// exact per-element terms (used for the rare chunks that leave the series' range)
static CAREL_DEVINL void bow_exact_terms(float x, float c, float& f, float& u) {
  const float p = __expf(x);
  const float A = fmaxf(x, -100.f);
  const float B = fmaxf(log1pf(-p), -100.f);
  f = c * A + (1.f - c) * B;
  u = p * (p - c) / fmaxf(p * (1.f - p), 1e-12f);
}

// ---------------------------------------------------------------------------------------------
// pass 2 epilogue: dense loss  sum_ij [c A + (1-c) B]  and row sums  S_i = sum_j u_ij
// Fast path (every p in the 16-chunk < 2^-10): power sums S1..S3 of p:
//   sum log1p(-p) = -(S1 + S2/2 + S3/3) ,  sum u = sum (p-c)/(1-p) = (S1+S2+S3) - c (n + S1 + S2)
// ---------------------------------------------------------------------------------------------
struct EpiBowLossParams {
  double* lossacc;
  float* srow;
  const int* row_flag;   // optional: only flagged rows contribute (exact redo of the rows pass F could not take)
  const float* bl2e;     // [Vpad]  b_j log2(e)
  const float* nlse2;    // [Npad]  -lse_i log2(e)
  int N, V;
  float c;
};
template <int BN>
struct EpiBowLoss {
  using Params = EpiBowLossParams;
  const Params& p;
  double loss;
  CAREL_DEVINL EpiBowLoss(const Params& p_, const TileCtx&) : p(p_), loss(0.0) {}
  CAREL_DEVINL void prefetch(const TileCtx&) {}
  CAREL_DEVINL void tile(const TileCtx& ctx) {
    constexpr int HALF = BN / kColSplit;
    const int c0 = ctx.col0 + ctx.half * HALF;
    if (c0 >= p.V || ctx.tm * 128 >= p.N) return;
    const bool full = c0 + HALF <= p.V;
    float sx = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, cnt = 0.f, fex = 0.f, uex = 0.f;
    const float c = p.c;
    // Plain operands: log p = (acc log2(e) + b_j log2(e) - lse_i log2(e)) ln 2, evaluated with EXACTLY the
    // operations of bow_dlogit16 (backward): for p -> 1 the gradient u_j - p_j S_i cancels catastrophically and
    // is only right if S_i (here) and u_j (there) come from bit-identical log p.
    const float rc = __ldg(p.nlse2 + min(ctx.row, p.N - 1));
    const float2 l2e = make_float2(kLog2e, kLog2e), rc2 = make_float2(rc, rc);
    for_each_chunk16(ctx.taddr + ctx.half * HALF, HALF / 16, [&](int ch, const uint32_t (&r)[16]) {
      float x[16];
      const float* ccp = p.bl2e + c0 + ch * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 t = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), l2e,
                                               make_float2(__ldg(ccp + i), __ldg(ccp + i + 1))), rc2);
        x[i] = t.x * kLn2;
        x[i + 1] = t.y * kLn2;
      }
      float mx = fmaxf(x[0], x[1]), mn = fminf(x[0], x[1]);
#pragma unroll
      for (int i = 2; i < 16; ++i) { mx = fmaxf(mx, x[i]); mn = fminf(mn, x[i]); }
      const bool chunk_full = full || (c0 + ch * 16 + 16 <= p.V);
      if (chunk_full && mx < kTinyLogP && mn > -27.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pv = ex2_approx(x[i] * kLog2e);
          const float p2 = pv * pv;
          sx += x[i];
          s1 += pv;
          s2 += p2;
          s3 = fmaf(p2, pv, s3);
        }
        cnt += 16.f;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c0 + ch * 16 + i < p.V) {
            float f, u;
            bow_exact_terms(x[i], c, f, u);
            fex += f;
            uex += u;
          }
        }
      }
    });
    if (ctx.row < p.N && (p.row_flag == nullptr || p.row_flag[ctx.row] != 0)) {
      const float f = c * sx - (1.f - c) * (s1 + 0.5f * s2 + (1.f / 3.f) * s3) + fex;
      const float u = (s1 + s2 + s3) - c * (cnt + s1 + s2) + uex;
      loss += double(f);
      atomicAdd(p.srow + ctx.row, u);
    }
  }
  CAREL_DEVINL void finish() {
    const double v = warp_sum(loss);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(p.lossacc, v);
  }
};

// ---------------------------------------------------------------------------------------------
// sparse corrections: one warp per row over its ELL entries (or a dense scan if it overflowed)
//   x = ZA[i] . WB[j]  (same rounded operands as the tensor-core passes)
//   fwd (mode 0):  lossacc[1] += sum (1-ls) t (A - B) ;  S_i += sum -(1-ls) t p / max(p(1-p),1e-12)
//   bwd (mode 1):  e = -(1-ls) t p / max(p(1-p),1e-12) * g/(N V);
//                  dZ[i,:] += e w_j ; dW[j,:] += e z_i ; db[j] += e
// ---------------------------------------------------------------------------------------------
struct BowSparseArgs {
  const __nv_bfloat16* ZA;
  const __nv_bfloat16* WB;
  const int* ell_cols;
  const float* ell_vals;
  const int* ell_nnz;
  float* ell_ev;         // mode 0: written; mode 1: unused
  int overflow_only;     // mode 1: handle only rows that overflowed the ELL list (the rest goes through CSC)
  const float* bow_dense;
  int64_t ld_bow;
  int N, V, K, ldk, cap, mode;
  float one_minus_ls;
  double* lossacc;
  float* srow;
  const float* grad_out;
  float* grad_z;
  int64_t ldgz;
  float* gw;             // [V, K] dW target (accumulated)
  float* gb;             // [V]    db target (accumulated)
};

// One warp per row; 4 non-zeros in flight (one per 8-lane group); the row's own operand (ZA[i], 8-element
// chunks c = (lane & 7) + 8 k) stays in registers for all of its non-zeros.
constexpr int kSparseChunks = 5;     // 8 lanes x 5 chunks x 8 elements = 320 >= ldk (<= 272)

// MODE is a template parameter: the forward instance does not carry the dz accumulators (160 -> <= 80 registers,
// 1 -> 3 blocks per SM; the kernel is latency-bound on the gathers, so occupancy is what it needs).
template <int MODE>
static __global__ void __launch_bounds__(256, MODE == 0 ? 3 : 1) bow_sparse_kernel(const BowSparseArgs a) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int grp = lane >> 3, gl = lane & 7;
  const int nchunk = a.ldk >> 3;
  const float scale = (MODE == 1) ? __ldg(a.grad_out) / (float(a.N) * float(a.V)) : 0.f;
  double loss_warp = 0.0;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < a.N; r += gridDim.x * wpb) {
    const int nnz = a.ell_nnz[r];
    const bool dense_scan = nnz > a.cap;
    if (MODE == 1 && a.overflow_only && !dense_scan) continue;
    const __nv_bfloat16* zrow = a.ZA + size_t(r) * a.ldk;
    uint4 zreg[kSparseChunks];
#pragma unroll
    for (int k = 0; k < kSparseChunks; ++k) {
      const int c = gl + 8 * k;
      zreg[k] = (c < nchunk) ? __ldg(reinterpret_cast<const uint4*>(zrow) + c) : make_uint4(0, 0, 0, 0);
    }
    float dloss = 0.f, du = 0.f;
    float dz[kSparseChunks][8];
    if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < kSparseChunks; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) dz[k][i] = 0.f;
    }
    // handle up to 4 (j, t) pairs at once: group g takes pair g (valid = pair exists)
    auto handle4 = [&](int j, float t, bool valid, int slot) {
      const __nv_bfloat16* wrow = a.WB + size_t(valid ? j : 0) * a.ldk;
      uint4 wreg[kSparseChunks];
#pragma unroll
      for (int k = 0; k < kSparseChunks; ++k) {
        const int c = gl + 8 * k;
        wreg[k] = (valid && c < nchunk) ? __ldg(reinterpret_cast<const uint4*>(wrow) + c) : make_uint4(0, 0, 0, 0);
      }
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kSparseChunks; ++k) {
        const uint32_t uz[4] = {zreg[k].x, zreg[k].y, zreg[k].z, zreg[k].w};
        const uint32_t uw[4] = {wreg[k].x, wreg[k].y, wreg[k].z, wreg[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uz[q]));
          const float2 fw = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[q]));
          acc = fmaf(fz.x, fw.x, acc);
          acc = fmaf(fz.y, fw.y, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (!valid) return;
      const float x = acc;
      const float p = __expf(x);
      const float den = fmaxf(p * (1.f - p), 1e-12f);
      const float ev = -a.one_minus_ls * t * p / den;
      if (MODE == 0) {
        if (gl == 0) {
          if (slot >= 0) a.ell_ev[slot] = ev;
          const float A = fmaxf(x, -100.f), B = fmaxf(log1pf(-p), -100.f);
          dloss += a.one_minus_ls * t * (A - B);
          du += ev;
        }
      } else {
        const float e = ev * scale;
        float* dwrow = a.gw + size_t(j) * a.K;
#pragma unroll
        for (int k = 0; k < kSparseChunks; ++k) {
          const int c = gl + 8 * k;
          if (c * 8 < a.K) {
            const uint32_t uz[4] = {zreg[k].x, zreg[k].y, zreg[k].z, zreg[k].w};
            const uint32_t uw[4] = {wreg[k].x, wreg[k].y, wreg[k].z, wreg[k].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 fz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uz[q]));
              const float2 fw = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[q]));
              dz[k][2 * q] = fmaf(e, fw.x, dz[k][2 * q]);
              dz[k][2 * q + 1] = fmaf(e, fw.y, dz[k][2 * q + 1]);
              atomicAdd(dwrow + c * 8 + 2 * q, e * fz.x);
              atomicAdd(dwrow + c * 8 + 2 * q + 1, e * fz.y);
            }
          }
        }
        if (gl == 0) atomicAdd(a.gb + j, e);
      }
    };
    if (!dense_scan) {
      for (int base = 0; base < nnz; base += 32) {
        const int q = base + lane;
        const int myj = (q < nnz) ? a.ell_cols[size_t(r) * a.cap + q] : 0;
        const float myt = (q < nnz) ? a.ell_vals[size_t(r) * a.cap + q] : 0.f;
        const int cnt = min(32, nnz - base);
        for (int s4 = 0; s4 < cnt; s4 += 4) {
          const int src = s4 + grp;
          const int j = __shfl_sync(0xffffffffu, myj, src & 31);
          const float t = __shfl_sync(0xffffffffu, myt, src & 31);
          handle4(j, t, src < cnt, int(size_t(r) * a.cap) + base + src);
        }
      }
    } else {
      const float* brow = a.bow_dense + size_t(r) * a.ld_bow;
      for (int base = 0; base < a.V; base += 32) {
        const int jj = base + lane;
        const float v = (jj < a.V) ? __ldg(brow + jj) : 0.f;
        unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
        while (m) {
          // pop up to four set bits; group g takes the g-th
          int srcs[4];
          int n = 0;
          unsigned mm = m;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            srcs[g] = mm ? (__ffs(mm) - 1) : 0;
            if (mm) { ++n; mm &= mm - 1; }
          }
          m = mm;
          const int src = srcs[grp];
          handle4(base + src, __shfl_sync(0xffffffffu, v, src), grp < n, -1);
        }
      }
    }
    if (MODE == 0) {
      // combine the four group leaders
      dloss += __shfl_xor_sync(0xffffffffu, dloss, 8);
      dloss += __shfl_xor_sync(0xffffffffu, dloss, 16);
      du += __shfl_xor_sync(0xffffffffu, du, 8);
      du += __shfl_xor_sync(0xffffffffu, du, 16);
      if (lane == 0) {
        loss_warp += double(dloss);          // reduced per block below (one double atomic per row serialises)
        if (du != 0.f) atomicAdd(a.srow + r, du);
      }
    } else {
      // sum the per-group dz partials and write the row (this kernel owns grad_z[r, :] at this point)
      float* gz = a.grad_z + size_t(r) * a.ldgz;
#pragma unroll
      for (int k = 0; k < kSparseChunks; ++k) {
        const int c = gl + 8 * k;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = dz[k][i];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (grp == 0 && c * 8 + i < a.K) gz[c * 8 + i] = v;
        }
      }
    }
  }
  if (MODE == 0) {          // block-level sum of the loss corrections: one double atomic per block
    __shared__ double sh_l[8];
    if (lane == 0) sh_l[threadIdx.x >> 5] = loss_warp;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < wpb; ++w) t += sh_l[w];
      if (t != 0.0) atomicAdd(a.lossacc + 1, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CSC view of the ELL list (counting sort by vocabulary index), so that the sparse part of dW is a
// gather per word instead of 256 fp32 atomics per non-zero.
// ---------------------------------------------------------------------------------------------
static __global__ void csc_count_kernel(const int* __restrict__ cols, const int* __restrict__ nnz, int N,
                                        int cap, int* __restrict__ cnt) {
  // N * cap < 2^31 for every workspace this library carves (the ELL arrays are int-indexed)
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < unsigned(N) * unsigned(cap); t += gridDim.x * blockDim.x) {
    const int r = int(t / unsigned(cap)), q = int(t % unsigned(cap)), n = nnz[r];
    if (n <= cap && q < n) atomicAdd(cnt + cols[t], 1);
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
                                       int cap, int* __restrict__ fillpos, int* __restrict__ csc_row,
                                       int* __restrict__ csc_slot) {
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < unsigned(N) * unsigned(cap); t += gridDim.x * blockDim.x) {
    const int r = int(t / unsigned(cap)), q = int(t % unsigned(cap)), n = nnz[r];
    if (n <= cap && q < n) {
      const int pos = atomicAdd(fillpos + cols[t], 1);
      csc_row[pos] = r;
      csc_slot[pos] = int(t);
    }
  }
}

// sparse part of dZ: one warp per row, dz_r = scale * sum_q ev_q * w_{j_q}; writes grad_z[r, 0:K]
// (zeros for rows without non-zeros; overflow rows are written by bow_sparse_kernel instead).
static __global__ void bow_sparse_dz_kernel(const __nv_bfloat16* __restrict__ WB, int ldk, int K,
                                            const int* __restrict__ cols, const int* __restrict__ nnz,
                                            const float* __restrict__ ev, int N, int cap,
                                            const float* __restrict__ grad_out, float inv_nv,
                                            float* __restrict__ grad_z, int64_t ldgz) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float scale = __ldg(grad_out) * inv_nv;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    const int n = nnz[r];
    if (n > cap) continue;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool act = lane * 8 < K;
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
          w[u] = act ? __ldg(reinterpret_cast<const uint4*>(WB + size_t(j) * ldk) + lane) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t v[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[h]));
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
static __global__ void bow_sparse_dw_kernel(const __nv_bfloat16* __restrict__ ZA, int ldk, int K,
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
    const bool act = lane * 8 < K;
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
          z[u] = act ? __ldg(reinterpret_cast<const uint4*>(ZA + size_t(r) * ldk) + lane) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          esum += e[u];
          const uint32_t v[4] = {z[u].x, z[u].y, z[u].z, z[u].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[h]));
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

static __global__ void bow_loss_finalize_kernel(const double* __restrict__ lossacc, double inv_nv,
                                                float* __restrict__ out) {
  if (threadIdx.x == 0) out[0] = float(-(lossacc[0] + lossacc[1]) * inv_nv);
}

// ---------------------------------------------------------------------------------------------
// Backward chain policies.  Both recompute acc1 = log p (ZA . WB incl. bias and -lse) and turn it into the
// dense part of dlogit,  d = [u - p S_i] g / (N V),  u = p (p - c) / max(p (1-p), 1e-12):
//   fast path (all 16 p < 2^-10, none clamped):  u = (p - c)(1 + p + p^2)  (relative truncation 2^-30),
//   written with packed fp32 pairs:  d = ((p - c) scale) q + p Sn,  q = 1 + p (1 + p),  Sn = -S_i scale.
//   any other chunk: exact per-element terms.
//   PolBowBwdW  owner = 128 vocabulary rows, inner = 64 clause rows:  acc2 = dW_v = sum_i d_iv z_i (GEMM-2 over
//               the SAME ZA tile read MN-major), db_v = thread-local sum;  S_i varies along the columns and is
//               prefetched before the accumulator barrier.
//   PolBowBwdZ  owner = 128 clause rows, inner = 64 vocabulary rows:   acc2 = dZ_i = sum_v d_iv w_v;  S_i is a
//               per-thread constant.
// Neither writes the [N, V] dlogits anywhere: the recompute costs one more GEMM-1 but no HBM round trip.
// ---------------------------------------------------------------------------------------------
struct BowBwdParams {
  const float* srow;     // [Npad]  S_i (zero beyond N)
  const float* nlse2;    // [Npad]  -lse_i log2(e)
  const float* bl2e;     // [Vpad]  b_j log2(e)
  const float* grad_out;
  float* gw;             // [V, K]  dW target (W policy; initialised by bow_sparse_dw_kernel)
  float* gb;             // [V]     db target
  float* grad_z;         // [N, ldgz]    (Z policy)
  int64_t ldgz;
  int N, V, K, ldk;
  float c, inv_nv;
};

// d[16] from 16 accumulators acc = z_i . w_j (plain operands, no augmentation columns):
//   log2 p = acc log2(e) + cc[col] + rc      (cc: per-column constant, rc: per-row constant)
// Sn[i] = -S * scale per element.
static CAREL_DEVINL void bow_dlogit16(const uint32_t (&r)[16], const float (&cc)[16], float rc, const float (&Sn)[16],
                                      float c, float scale, bool in_range, int n_valid, float (&d)[16]) {
  const float2 l2e = make_float2(kLog2e, kLog2e), rc2 = make_float2(rc, rc);
  float t[16];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 v = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), l2e,
                                           make_float2(cc[i], cc[i + 1])), rc2);
    t[i] = v.x;
    t[i + 1] = v.y;
  }
  float mx = fmax3(t[0], t[1], t[2]), mn = fmin3(t[0], t[1], t[2]);
#pragma unroll
  for (int i = 3; i < 15; i += 2) { mx = fmax3(mx, t[i], t[i + 1]); mn = fmin3(mn, t[i], t[i + 1]); }
  mx = fmaxf(mx, t[15]);
  mn = fminf(mn, t[15]);
  if (in_range && mx < -10.f && mn > -27.f * kLog2e) {          // all p in (e^-27, 2^-10)
    const float2 one = make_float2(1.f, 1.f);
    const float2 sc = make_float2(scale, scale), mcs = make_float2(-c * scale, -c * scale);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 pv = make_float2(ex2_approx(t[i]), ex2_approx(t[i + 1]));
      const float2 q = __ffma2_rn(pv, __fadd2_rn(pv, one), one);              // 1 + p + p^2
      const float2 a = __ffma2_rn(pv, sc, mcs);                               // (p - c) scale
      const float2 m = __fmul2_rn(pv, make_float2(Sn[i], Sn[i + 1]));         // -p S scale
      const float2 o = __ffma2_rn(a, q, m);
      d[i] = o.x;
      d[i + 1] = o.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = t[i] * kLn2;
      float f, u;
      bow_exact_terms(x, c, f, u);
      d[i] = (i < n_valid) ? fmaf(u, scale, __expf(x) * Sn[i]) : 0.f;
    }
  }
}

struct PolBowBwdW {
  using Params = BowBwdParams;
  static constexpr bool kHiLo = false;    // measured decoder gradients are 6e-5..3e-4 with a single bf16 P
  const Params& p;
  float scale, dbsum, rc;
  float4 sraw[4], lraw[4];        // prefetched S_i and -lse_i log2(e) of this thread's 16 columns (raw: no use before
                                  // the accumulator barrier, so the L2 latency of the loads hides behind the wait)
  CAREL_DEVINL PolBowBwdW(const Params& p_, const ChainCtx&) : p(p_), scale(0.f), dbsum(0.f), rc(0.f) {}
  CAREL_DEVINL void begin_item(const ChainCtx& ctx) {
    scale = __ldg(p.grad_out) * p.inv_nv;
    dbsum = 0.f;
    rc = __ldg(p.bl2e + min(ctx.row, p.V - 1));         // this thread's word: b_v log2(e)
  }
  CAREL_DEVINL void prefetch(const ChainCtx&, int i0) {
    const float4* s4 = reinterpret_cast<const float4*>(p.srow + i0);
    const float4* l4 = reinterpret_cast<const float4*>(p.nlse2 + i0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sraw[i] = __ldg(s4 + i);
      lraw[i] = __ldg(l4 + i);
    }
  }
  CAREL_DEVINL void begin_tile(const ChainCtx&, int, int) {}
  CAREL_DEVINL void weights16(const ChainCtx& ctx, int i0, const uint32_t (&r)[16], const uint32_t (&)[16],
                              float (&d)[16]) {
    if (ctx.row >= p.V) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = 0.f;
      return;
    }
    float Sn[16], cc[16];
    const float ns = -scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Sn[4 * i] = sraw[i].x * ns; Sn[4 * i + 1] = sraw[i].y * ns; Sn[4 * i + 2] = sraw[i].z * ns; Sn[4 * i + 3] = sraw[i].w * ns;
      cc[4 * i] = lraw[i].x; cc[4 * i + 1] = lraw[i].y; cc[4 * i + 2] = lraw[i].z; cc[4 * i + 3] = lraw[i].w;
    }
    bow_dlogit16(r, cc, rc, Sn, p.c, scale, i0 + 16 <= p.N, p.N - i0, d);
    float2 t = make_float2(0.f, 0.f);    // db_v: thread-local column sum (fp32, before the bf16 rounding)
#pragma unroll
    for (int i = 0; i < 16; i += 2) t = __fadd2_rn(t, make_float2(d[i], d[i + 1]));
    dbsum += t.x + t.y;
  }
  CAREL_DEVINL void store16(const ChainCtx&, int, const uint32_t (&)[8]) {}
  // acc2 -> registers (this thread's K / kChainSplit columns of its row), release the accumulator, THEN reduce
  // into global memory: GEMM-2 of the next item does not wait for the reductions.
  template <class Release>
  CAREL_DEVINL void final(const ChainCtx& ctx, bool valid, Release&& release) {
    constexpr int kMaxQ = 256 / 16 / kChainSplit;   // chunks per thread at K = 256
    const int n16 = p.K / 16;                       // feature columns only
    const int q0 = (n16 * ctx.half) / kChainSplit, q1 = (n16 * (ctx.half + 1)) / kChainSplit;
    const bool row_ok = valid && ctx.row < p.V;
    float* dst = p.gw + size_t(min(ctx.row, p.V - 1)) * p.K;
    (void)kMaxQ;
    // two chunks (32 registers) at a time: holding all of this thread's columns at once spills
#pragma unroll 1
    for (int q = q0; q < q1; q += 2) {
      uint32_t ra[16], rb[16];
      const bool two = q + 1 < q1;
      tmem_ld16(ctx.acc2_taddr + q * 16, ra);
      if (two) tmem_ld16(ctx.acc2_taddr + (q + 1) * 16, rb);
      tmem_ld_wait();
      if (q + 2 >= q1) release();                 // last read of acc2 by this thread
      if (!row_ok) continue;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        red_add_v4(dst + q * 16 + i, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]), __uint_as_float(ra[i + 2]),
                   __uint_as_float(ra[i + 3]));
      if (two) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dst + (q + 1) * 16 + i, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]),
                     __uint_as_float(rb[i + 2]), __uint_as_float(rb[i + 3]));
      }
    }
    if (q0 >= q1) release();
    if (row_ok && dbsum != 0.f) atomicAdd(p.gb + ctx.row, dbsum);    // every column part adds its share
  }
  CAREL_DEVINL void finish() {}
};

struct PolBowBwdZ {
  using Params = BowBwdParams;
  static constexpr bool kHiLo = false;
  const Params& p;
  float scale, rc, sn;
  float4 braw[4];                 // prefetched b_j log2(e) of this thread's 16 columns
  CAREL_DEVINL PolBowBwdZ(const Params& p_, const ChainCtx&) : p(p_), scale(0.f), rc(0.f), sn(0.f) {}
  CAREL_DEVINL void begin_item(const ChainCtx& ctx) {
    scale = __ldg(p.grad_out) * p.inv_nv;
    const bool ok = ctx.row < p.N;
    sn = ok ? -__ldg(p.srow + ctx.row) * scale : 0.f;
    rc = ok ? __ldg(p.nlse2 + ctx.row) : 0.f;           // this thread's clause: -lse_i log2(e)
  }
  CAREL_DEVINL void prefetch(const ChainCtx&, int v0) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bl2e + v0);
#pragma unroll
    for (int i = 0; i < 4; ++i) braw[i] = __ldg(b4 + i);
  }
  CAREL_DEVINL void begin_tile(const ChainCtx&, int, int) {}
  CAREL_DEVINL void weights16(const ChainCtx&, int v0, const uint32_t (&r)[16], const uint32_t (&)[16],
                              float (&d)[16]) {
    float Sn[16], cc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cc[4 * i] = braw[i].x; cc[4 * i + 1] = braw[i].y; cc[4 * i + 2] = braw[i].z; cc[4 * i + 3] = braw[i].w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) Sn[i] = sn;
    bow_dlogit16(r, cc, rc, Sn, p.c, scale, v0 + 16 <= p.V, p.V - v0, d);
  }
  CAREL_DEVINL void store16(const ChainCtx&, int, const uint32_t (&)[8]) {}
  template <class Release>
  CAREL_DEVINL void final(const ChainCtx& ctx, bool valid, Release&& release) {
    constexpr int kMaxQ = 256 / 16 / kChainSplit;
    const int n16 = p.K / 16;
    const int q0 = (n16 * ctx.half) / kChainSplit, q1 = (n16 * (ctx.half + 1)) / kChainSplit;
    const bool row_ok = valid && ctx.row < p.N;
    float* dst = p.grad_z + size_t(min(ctx.row, p.N - 1)) * p.ldgz;
    const bool vec = ((p.ldgz & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.grad_z) & 15) == 0);
    (void)kMaxQ;
#pragma unroll 1
    for (int q = q0; q < q1; q += 2) {
      uint32_t ra[16], rb[16];
      const bool two = q + 1 < q1;
      tmem_ld16(ctx.acc2_taddr + q * 16, ra);
      if (two) tmem_ld16(ctx.acc2_taddr + (q + 1) * 16, rb);
      tmem_ld_wait();
      if (q + 2 >= q1) release();                 // last read of acc2 by this thread
      if (!row_ok) continue;
      float* o = dst + q * 16;
      if (vec) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(o + i, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]), __uint_as_float(ra[i + 2]),
                     __uint_as_float(ra[i + 3]));
        if (two) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            red_add_v4(o + 16 + i, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]), __uint_as_float(rb[i + 2]),
                       __uint_as_float(rb[i + 3]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(o + i, __uint_as_float(ra[i]));
        if (two) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(o + 16 + i, __uint_as_float(rb[i]));
        }
      }
    }
    if (q0 >= q1) release();
  }
  CAREL_DEVINL void finish() {}
};

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
  bow_pack_ell_kernel<8, 256, 1><<<int(std::min<int64_t>((N + 7) / 8, 148 * 32)), 256, 0, st>>>(
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
  CAREL_REQUIRE(K % 16 == 0 && K <= 256, "carel_bow_fwd: decoder input width (2 * ec_dim) must be a multiple of 16 and <= 256, got %lld", (long long)K);
  CAREL_REQUIRE(bow_dense != nullptr || (ell_cols_in && ell_vals_in && ell_nnz_in), "carel_bow_fwd: need a dense or an ELL target");
  BowWs w = carve_bow(ws, N, K, V, ell_cap);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_fwd: workspace too small");
  const int cap = int(ell_cap);
  // ---- operands ----
  pack_rows_kernel<<<int((N + 7) / 8), 256, 0, st>>>(z, int(N), int(K), int(ldz), w.ZA, nullptr, w.ldk, 0, 2, nullptr);
  CAREL_LAUNCHED();
  pack_rows_kernel<<<int(std::min<int64_t>((V + 7) / 8, 148 * 32)), 256, 0, st>>>(dec_w, int(V), int(K), int(K), nullptr, w.WB, w.ldk, 0, 1, nullptr, dec_b);
  CAREL_LAUNCHED();
  // ---- sparse target ----
  if (ell_cols_in != nullptr) {
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_cols, ell_cols_in, size_t(N) * cap * 4, cudaMemcpyDeviceToDevice, st));
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_vals, ell_vals_in, size_t(N) * cap * 4, cudaMemcpyDeviceToDevice, st));
    CAREL_CHECK_CUDA(cudaMemcpyAsync(w.ell_nnz, ell_nnz_in, size_t(N) * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    bow_pack_ell_kernel<8, 256, 1><<<int(std::min<int64_t>((N + 7) / 8, 148 * 32)), 256, 0, st>>>(bow_dense, ld_bow, int(N), int(V), cap, w.ell_cols, w.ell_vals, w.ell_nnz);
    CAREL_LAUNCHED();
  }
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.lossacc, 0, 2 * sizeof(double), st));
  // ---- pass F: the one dense pass (online max, power sums) + per-row combine ----
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.wsum, 0, size_t(w.ldk) * 4, st));
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.tile_flag, 0, (size_t(w.Npad / 128) + 1) * 4, st));
  {
    const dim3 blk((w.ldk / 8 + 7) / 8 * 8, 16);
    bow_colsum_bias_kernel<<<148 * 2, blk, size_t(blk.y) * w.ldk * sizeof(float), st>>>(
        w.WB, int(V), w.ldk, w.wsum, dec_b, int(round_up64(V, 64)), w.bl2e);
  }
  CAREL_LAUNCHED();
  // plain operands (columns [0, K) of ZA / WB) through the A-stationary GEMM: bias is added in the epilogue
  CUtensorMap ta2, tb2;
  CAREL_REQUIRE(make_tmap_bf16_sw128(&ta2, w.ZA, N, K, w.ldk, 128) == 0, "carel_bow_fwd: tensor map Z failed");
  CAREL_REQUIRE(make_tmap_bf16_sw128(&tb2, w.WB, V, K, w.ldk, 128) == 0, "carel_bow_fwd: tensor map W failed");
  AsSched as = make_as_sched(N, V, K, sm_count());
  EpiBowFusedParams e1;
  e1.part4 = w.part4; e1.part2 = w.part2; e1.bl2e = w.bl2e; e1.N = int(N); e1.V = int(V); e1.Npad = w.Npad;
  e1.n_tiles = as.n_tiles; e1.tpc = as.tpc;
  const int nparts = std::min(as.n_tiles, (as.n_tiles - 1) / as.tpc + 2) * kColSplit;      // <= w.ntp
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.part4, 0, size_t(nparts) * w.Npad * sizeof(float4), st));
  if (int rc = launch_gemm_as<EpiBowFused<256>>(ta2, tb2, as, e1, st, "bow_passF")) return rc;
  bow_combine_kernel<<<w.Npad / kCombineRows, kCombineRows * kCombineGroups, 0, st>>>(
      w.part4, w.part2, nparts, w.Npad, int(N), int(V), int(K), w.ldk, label_smoothing / float(V), w.wsum, w.ZA,
      w.lse, w.nlse2, w.srow, w.lossacc, w.row_flag, w.tile_flag);
  CAREL_LAUNCHED();
  // ---- exact redo of the flagged rows (skips every row tile without a flag: normally all of them) ----
  EpiBowLossParams e2;
  e2.lossacc = w.lossacc; e2.srow = w.srow; e2.row_flag = w.row_flag; e2.bl2e = w.bl2e; e2.nlse2 = w.nlse2;
  e2.N = int(N); e2.V = int(V); e2.c = label_smoothing / float(V);
  as.row_tile_flags = w.tile_flag;      // [0] = any flagged row at all, [1 + tile] per row tile
  if (int rc = launch_gemm_as<EpiBowLoss<256>>(ta2, tb2, as, e2, st, "bow_pass2_exact")) return rc;
  // ---- sparse corrections ----
  BowSparseArgs sa;
  sa.ZA = w.ZA; sa.WB = w.WB; sa.ell_cols = w.ell_cols; sa.ell_vals = w.ell_vals; sa.ell_nnz = w.ell_nnz;
  sa.bow_dense = bow_dense; sa.ld_bow = ld_bow; sa.N = int(N); sa.V = int(V); sa.K = int(K); sa.ldk = w.ldk;
  sa.cap = cap; sa.mode = 0; sa.one_minus_ls = 1.f - label_smoothing; sa.lossacc = w.lossacc; sa.srow = w.srow;
  sa.ell_ev = w.ell_ev; sa.overflow_only = 0;
  sa.grad_out = nullptr; sa.grad_z = nullptr; sa.ldgz = 0; sa.gw = nullptr; sa.gb = nullptr;
  bow_sparse_kernel<0><<<int(std::min<int64_t>((N + 7) / 8, 148 * 16)), 256, 0, st>>>(sa);
  CAREL_LAUNCHED();
  bow_loss_finalize_kernel<<<1, 32, 0, st>>>(w.lossacc, 1.0 / (double(N) * double(V)), loss_out);
  CAREL_LAUNCHED();
  // CSC view of the non-zero list for the backward gather (cheap: N*cap entries)
  CAREL_CHECK_CUDA(cudaMemsetAsync(w.fillpos, 0, (size_t(V) + 1) * 4, st));
  csc_count_kernel<<<148 * 2, 256, 0, st>>>(w.ell_cols, w.ell_nnz, int(N), cap, w.fillpos);
  CAREL_LAUNCHED();
  {
    static const cudaError_t attr = cudaFuncSetAttribute(csc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         kScanSmemBytes);
    CAREL_CHECK_CUDA(attr);
  }
  csc_scan_kernel<<<1, 1024, kScanSmemBytes, st>>>(w.fillpos, int(V), w.colptr, w.fillpos);
  CAREL_LAUNCHED();
  csc_fill_kernel<<<148 * 2, 256, 0, st>>>(w.ell_cols, w.ell_nnz, int(N), cap, w.fillpos, w.csc_row, w.csc_slot);
  CAREL_LAUNCHED();
  return 0;
}

extern "C" int carel_bow_bwd(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N,
                             int64_t K, int64_t V, int64_t ell_cap, float label_smoothing,
                             float* grad_z, int64_t ldgz, float* grad_w, float* grad_b, void* ws,
                             size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BowWs w = carve_bow(ws, N, K, V, ell_cap);
  CAREL_REQUIRE(ws != nullptr && ws_bytes >= w.bytes, "carel_bow_bwd: workspace too small");
  CAREL_REQUIRE(grad_z != nullptr, "carel_bow_bwd: grad_z is required");
  const bool need_w = grad_w != nullptr || grad_b != nullptr;
  // grad_z[r, 0:K] and (if wanted) grad_w / grad_b are fully written (not accumulated) by the sparse gather
  // kernels; the dense passes then add to them.  A gradient the caller does not want lands in scratch.
  float* gw = grad_w != nullptr ? grad_w : w.dWacc;
  float* gb = grad_b != nullptr ? grad_b : w.dbacc;
  // ---- sparse part: gathers through the ELL (rows) and CSC (words) views; rows that overflowed the ELL
  // list fall back to the scan + atomics kernel ----
  const float inv_nv = float(1.0 / (double(N) * double(V)));
  bow_sparse_dz_kernel<<<int(std::min<int64_t>((N + 7) / 8, 148 * 16)), 256, 0, st>>>(
      w.WB, w.ldk, int(K), w.ell_cols, w.ell_nnz, w.ell_ev, int(N), int(ell_cap), grad_out, inv_nv, grad_z, ldgz);
  CAREL_LAUNCHED();
  if (need_w) {
    bow_sparse_dw_kernel<<<148 * 16, 256, 0, st>>>(w.ZA, w.ldk, int(K), w.colptr, w.csc_row, w.csc_slot, w.ell_ev,
                                                  int(V), grad_out, inv_nv, gw, gb);
    CAREL_LAUNCHED();
  }
  if (bow_dense != nullptr) {
    BowSparseArgs sa;
    sa.ZA = w.ZA; sa.WB = w.WB; sa.ell_cols = w.ell_cols; sa.ell_vals = w.ell_vals; sa.ell_nnz = w.ell_nnz;
    sa.ell_ev = nullptr; sa.overflow_only = 1;
    sa.bow_dense = bow_dense; sa.ld_bow = ld_bow; sa.N = int(N); sa.V = int(V); sa.K = int(K); sa.ldk = w.ldk;
    sa.cap = int(ell_cap); sa.mode = 1; sa.one_minus_ls = 1.f - label_smoothing; sa.lossacc = nullptr; sa.srow = nullptr;
    sa.grad_out = grad_out; sa.grad_z = grad_z; sa.ldgz = ldgz; sa.gw = gw; sa.gb = gb;
    bow_sparse_kernel<1><<<int(std::min<int64_t>((N + 7) / 8, 148 * 16)), 256, 0, st>>>(sa);
    CAREL_LAUNCHED();
  }
  // The chains use the PLAIN operands (columns [0, K) of ZA / WB; bias and -lse are added in the epilogue), so
  // K = 4 full 64-wide blocks: A 64 KB + 5 ring stages of 32 KB, and 4 accumulator buffers in flight.
  const int k16 = int(K / 16);
  BowBwdParams bp;
  bp.srow = w.srow; bp.nlse2 = w.nlse2; bp.bl2e = w.bl2e; bp.grad_out = grad_out; bp.gw = gw; bp.gb = gb;
  bp.grad_z = grad_z; bp.ldgz = ldgz;
  bp.N = int(N); bp.V = int(V); bp.K = int(K); bp.ldk = w.ldk;
  bp.c = label_smoothing / float(V); bp.inv_nv = inv_nv;
  constexpr int NB = 4;
  using Cfg = ChainCfg<1, 64, NB>;
  // ---- pass 3 (vocabulary-major): dense dlogits -> [dW | db] accumulated in TMEM per 128 words ----
  if (need_w) {
    CUtensorMap ta, tb;
    CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.WB, V, K, w.ldk, 128) == 0, "carel_bow_bwd: tensor map W failed");
    CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.ZA, N, K, w.ldk, 64) == 0, "carel_bow_bwd: tensor map Z failed");
    ChainSched cs = make_chain_sched(0, int((V + 127) / 128), int((N + 63) / 64), k16, Cfg::stages_for(k16), 16);
    cs.n2_16 = int(K / 16);
    if (int rc = launch_chain<PolBowBwdW, 1, 64, NB>(ta, tb, ta, tb, cs, bp, st, "bow_pass3_dw")) return rc;
  }
  // ---- pass 4 (clause-major): dense dlogits -> dZ accumulated in TMEM per 128 clauses ----
  {
    CUtensorMap ta, tb;
    CAREL_REQUIRE(make_tmap_bf16_sw128(&ta, w.ZA, N, K, w.ldk, 128) == 0, "carel_bow_bwd: tensor map Z failed");
    CAREL_REQUIRE(make_tmap_bf16_sw128(&tb, w.WB, V, K, w.ldk, 64) == 0, "carel_bow_bwd: tensor map W failed");
    ChainSched cs = make_chain_sched(0, int((N + 127) / 128), int((V + 63) / 64), k16, Cfg::stages_for(k16), 64);
    cs.n2_16 = int(K / 16);
    if (int rc = launch_chain<PolBowBwdZ, 1, 64, NB>(ta, tb, ta, tb, cs, bp, st, "bow_pass4_dz")) return rc;
  }
  return 0;
}

#ifdef CAREL_CHAIN_TRACE
// debug: copy the chain-kernel timeline of this translation unit (8 x 256 clock64 values) to the host
extern "C" int carel_debug_chain_trace(long long* host_dst) {
  CAREL_CHECK_CUDA(cudaMemcpyFromSymbol(host_dst, carel::g_chain_trace, sizeof(long long) * 8 * 256));
  return 0;
}
#endif
