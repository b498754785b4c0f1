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
//   f16 != 0: the same layout in FP16 (11 mantissa bits at the same MMA rate; values saturate at +-65504): used where
//              the rows are an INTERMEDIATE of bounded range (the latent sample in the head's MMD) whose 8-bit bf16
//              rounding is not shared with the reference and would dominate the gradient error of small batches.
// 16-bit conversions of both formats behind one flag (bit patterns; round-to-nearest-even)
struct Cvt16 {
  int f16;
  CAREL_DEVINL uint16_t to(float v) const {
    if (f16) {
      const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      return *reinterpret_cast<const uint16_t*>(&h);
    }
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<const uint16_t*>(&b);
  }
  CAREL_DEVINL float from(uint16_t u) const {
    if (f16) return __half2float(*reinterpret_cast<const __half*>(&u));
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&u));
  }
};

static __global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int K, int ld,
                                        __nv_bfloat16* __restrict__ A_, __nv_bfloat16* __restrict__ B_,
                                        int ldp, int dst_row0, int aug, float* __restrict__ norms,
                                        const float* __restrict__ hvec = nullptr,
                                        const float* __restrict__ src1 = nullptr, int rows0 = 0, int ld1 = 0,
                                        int f16 = 0) {
  uint16_t* A = reinterpret_cast<uint16_t*>(A_);
  uint16_t* B = reinterpret_cast<uint16_t*>(B_);
  const Cvt16 cv{f16};
  // optional second source: rows [rows0, rows) come from src1 (stacked samples cat(x, y) in ONE launch)
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const bool second = src1 != nullptr && r >= rows0;
    const float* s = second ? src1 + size_t(r - rows0) * ld1 : src + size_t(r) * ld;
    uint16_t* a = (A != nullptr) ? A + size_t(dst_row0 + r) * ldp : nullptr;
    uint16_t* b = (B != nullptr) ? B + size_t(dst_row0 + r) * ldp : nullptr;
    float nrm = 0.f;
    const bool vec = ((K & 3) == 0) && (((second ? ld1 : ld) & 3) == 0) && ((ldp & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(second ? src1 : src) & 15) == 0);
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + k));
        const uint16_t q0 = cv.to(v.x), q1 = cv.to(v.y), q2 = cv.to(v.z), q3 = cv.to(v.w);
        const float f0 = cv.from(q0), f1 = cv.from(q1), f2 = cv.from(q2), f3 = cv.from(q3);
        nrm += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
        uint2 o;
        o.x = uint32_t(q0) | (uint32_t(q1) << 16);
        o.y = uint32_t(q2) | (uint32_t(q3) << 16);
        if (a != nullptr) *reinterpret_cast<uint2*>(a + k) = o;
        if (b != nullptr) *reinterpret_cast<uint2*>(b + k) = o;
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const uint16_t v = cv.to(__ldg(s + k));
        const float f = cv.from(v);
        nrm += f * f;
        if (a != nullptr) a[k] = v;
        if (b != nullptr) b[k] = v;
      }
    }
    nrm = warp_sum(nrm);
    if (norms != nullptr && lane == 0) norms[dst_row0 + r] = nrm;
    // tail: augmentation columns and zero padding
    const float h = (aug == 2) ? 0.f : ((hvec != nullptr) ? __ldg(hvec + r) : -0.5f * nrm);
    const uint16_t h1 = cv.to(h);
    const float r1 = h - cv.from(h1);
    const uint16_t h2 = cv.to(r1);
    const uint16_t h3 = cv.to(r1 - cv.from(h2));
    const uint16_t one = cv.to(1.f), zero = 0;
    for (int k = K + lane; k < ldp; k += 32) {
      uint16_t va = zero, vb = zero;
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
