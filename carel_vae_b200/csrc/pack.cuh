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
                                        const float* __restrict__ hvec = nullptr) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const float* s = src + size_t(r) * ld;
    __nv_bfloat16* a = (A != nullptr) ? A + size_t(dst_row0 + r) * ldp : nullptr;
    __nv_bfloat16* b = (B != nullptr) ? B + size_t(dst_row0 + r) * ldp : nullptr;
    float nrm = 0.f;
    const bool vec = ((K & 3) == 0) && ((ld & 3) == 0) && ((ldp & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
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

}  // namespace carel
