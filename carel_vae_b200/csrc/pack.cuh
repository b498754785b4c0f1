// pack.cuh — HBM-bound staging kernels: fp32 -> bf16 operand packing (optionally "augmented"
// with norm / ones columns), small reductions.  Coalesced 128-bit accesses, warp-shuffle
// reductions.
#pragma once
#include "common.cuh"

namespace carel {

// Width of an augmented operand row: K features + 3 norm-split + 3 ones columns, padded to 16.
inline __host__ __device__ int aug_width(int K) { return (K + 6 + 15) / 16 * 16; }

// One warp per row.
//   aug == 0:  A[r, 0:ldp] = bf16(src[r, 0:K]) zero padded.                       (B unused)
//   aug == 1:  with s = bf16(src[r,:]), h = -|s|^2 / 2 split exactly into three bf16 (h1+h2+h3):
//              A[r] = [ s, h1, h2, h3, 1, 1, 1, 0.. ]   (row operand)
//              B[r] = [ s, 1, 1, 1, h1, h2, h3, 0.. ]   (column operand)
//              so that  A[i] . B[j] = s_i.s_j - |s_i|^2/2 - |s_j|^2/2 = -d2_ij / 2  accumulates
//              in the tensor core and column K of B is the constant 1 (row sums for free).
//   norms (optional) receives |s|^2.
//   aug == 2:  like aug == 1 with h = 0.
//   hvec (optional, aug == 1): h = hvec[r] instead of -|s|^2/2 (decoder: bias / -logsumexp).
static __global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int K, int ld,
                                        __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ B,
                                        int ldp, int dst_row0, int aug, float* __restrict__ norms,
                                        const float* __restrict__ hvec = nullptr,
                                        const float* __restrict__ src1 = nullptr, int rows0 = 0, int ld1 = 0) {
  // optional second source: rows [rows0, rows) come from src1 (stacked samples cat(x, y) in ONE launch)
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const bool second = src1 != nullptr && r >= rows0;
    const float* s = second ? src1 + size_t(r - rows0) * ld1 : src + size_t(r) * ld;
    __nv_bfloat16* a = (A != nullptr) ? A + size_t(dst_row0 + r) * ldp : nullptr;
    __nv_bfloat16* b = (B != nullptr) ? B + size_t(dst_row0 + r) * ldp : nullptr;
    float nrm = 0.f;
    const bool vec = ((K & 3) == 0) && (((second ? ld1 : ld) & 3) == 0) && ((ldp & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(second ? src1 : src) & 15) == 0);
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + k));
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        const float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);
        nrm += flo.x * flo.x + flo.y * flo.y + fhi.x * fhi.x + fhi.y * fhi.y;
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        if (a != nullptr) *reinterpret_cast<uint2*>(a + k) = o;
        if (b != nullptr) *reinterpret_cast<uint2*>(b + k) = o;
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(s + k));
        const float f = __bfloat162float(v);
        nrm += f * f;
        if (a != nullptr) a[k] = v;
        if (b != nullptr) b[k] = v;
      }
    }
    nrm = warp_sum(nrm);
    if (norms != nullptr && lane == 0) norms[dst_row0 + r] = nrm;
    // tail: augmentation columns and zero padding
    const float h = (aug == 2) ? 0.f : ((hvec != nullptr) ? __ldg(hvec + r) : -0.5f * nrm);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(h);
    const float r1 = h - __bfloat162float(h1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
    for (int k = K + lane; k < ldp; k += 32) {
      __nv_bfloat16 va = zero, vb = zero;
      if (aug) {
        const int t = k - K;
        if (t == 0) { va = h1; vb = one; }
        else if (t == 1) { va = h2; vb = one; }
        else if (t == 2) { va = h3; vb = one; }
        else if (t == 3) { va = one; vb = h1; }
        else if (t == 4) { va = one; vb = h2; }
        else if (t == 5) { va = one; vb = h3; }
      }
      if (a != nullptr) a[k] = va;
      if (b != nullptr) b[k] = vb;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// "fp32 mode" operands: every value is split into two bf16 terms v = v1 + v2 (16 mantissa bits) and the
// product a.b ~ a1 b1 + a2 b1 + a1 b2 is obtained from ONE bf16 GEMM by concatenating along K:
//     pattern P (0): [ v1 | v1 | aug | v2 ]      pattern Q (1): [ v1 | v2 | aug | v1 ]
// (each block Kp = roundup16(K) wide, aug = 16 columns), a P row against a Q row gives all three terms.
// The aug block is [h1 h2 h3 1 1 1 0..] (type A) or [1 1 1 h1 h2 h3 0..] (type B) as in pack_rows_kernel,
// with h = -|v|^2/2 (the TRUE fp32 norm: off-diagonal products are unbiased around v.w, so using the
// truncated self-product would shift every d2 by ~2^-18 |v|^2), hvec[r], or 0.
// The first 2 Kp columns of a Q row are v1 | v2, which is what GEMM-2 of the chain kernel multiplies with.
// ---------------------------------------------------------------------------------------------
inline __host__ __device__ int split_kp(int K) { return (K + 15) / 16 * 16; }
inline __host__ __device__ int split_width(int K) { return 3 * split_kp(K) + 16; }

struct Pack3Out {
  __nv_bfloat16* ptr;   // nullptr = not produced
  int pattern;          // 0 = P, 1 = Q
  int aug_type;         // 0 = A, 1 = B
};

static __global__ void pack3_rows_kernel(const float* __restrict__ src, int rows, int K, int ld, Pack3Out o0,
                                         Pack3Out o1, int dst_row0, int hmode /*0 norm, 1 hvec, 2 zero*/,
                                         const float* __restrict__ hvec, float* __restrict__ copy32, int ld32) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Kp = split_kp(K), W = split_width(K);
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const float* s = src + size_t(r) * ld;
    __nv_bfloat16* d0 = o0.ptr ? o0.ptr + size_t(dst_row0 + r) * W : nullptr;
    __nv_bfloat16* d1 = o1.ptr ? o1.ptr + size_t(dst_row0 + r) * W : nullptr;
    float nrm = 0.f;
    for (int k = lane; k < Kp; k += 32) {
      const float v = (k < K) ? __ldg(s + k) : 0.f;
      const __nv_bfloat16 v1 = __float2bfloat16_rn(v);
      const float f1 = __bfloat162float(v1);
      const __nv_bfloat16 v2 = __float2bfloat16_rn(v - f1);
      nrm = fmaf(v, v, nrm);
      if (d0) {
        d0[k] = v1;
        d0[Kp + k] = o0.pattern ? v2 : v1;
        d0[2 * Kp + 16 + k] = o0.pattern ? v1 : v2;
      }
      if (d1) {
        d1[k] = v1;
        d1[Kp + k] = o1.pattern ? v2 : v1;
        d1[2 * Kp + 16 + k] = o1.pattern ? v1 : v2;
      }
      if (copy32 != nullptr && k < K) copy32[size_t(dst_row0 + r) * ld32 + k] = v;
    }
    nrm = warp_sum(nrm);
    const float h = (hmode == 2) ? 0.f : ((hmode == 1) ? __ldg(hvec + r) : -0.5f * nrm);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(h);
    const float r1 = h - __bfloat162float(h1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
    if (lane < 16) {
      const __nv_bfloat16 hv[3] = {h1, h2, h3};
      for (int o = 0; o < 2; ++o) {
        __nv_bfloat16* d = o ? d1 : d0;
        if (!d) continue;
        const int type = o ? o1.aug_type : o0.aug_type;
        __nv_bfloat16 v = zero;
        if (lane < 3) v = type ? one : hv[lane];
        else if (lane < 6) v = type ? hv[lane - 3] : one;
        d[2 * Kp + lane] = v;
      }
    }
  }
}

}  // namespace carel
