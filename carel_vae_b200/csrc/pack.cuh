// pack.cuh — HBM-bound staging kernels: fp32 -> bf16 operand packing with row norms,
// transposes, small reductions.  Coalesced 128-bit accesses, warp-shuffle reductions.
#pragma once
#include "common.cuh"

namespace carel {

// One warp per row.  dst[dst_row0 + r, 0:Kp] = bf16(src[r, 0:K]) zero padded to Kp;
// norms[dst_row0 + r] = sum_k bf16(src[r,k])^2 (the norm of what the tensor core will see).
// If bias_col >= 0 nothing special; plain conversion only.
static __global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int K, int ld,
                                 __nv_bfloat16* __restrict__ dst, int Kp, int dst_row0,
                                 float* __restrict__ norms) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const float* s = src + size_t(r) * ld;
    __nv_bfloat16* d = dst + size_t(dst_row0 + r) * Kp;
    float nrm = 0.f;
    const bool vec = ((K & 3) == 0) && ((ld & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
      for (int k = lane * 4; k < Kp; k += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) v = __ldg(reinterpret_cast<const float4*>(s + k));
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        const float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);
        nrm += flo.x * flo.x + flo.y * flo.y + fhi.x * fhi.x + fhi.y * fhi.y;
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(d + k) = o;
      }
    } else {
      for (int k = lane; k < Kp; k += 32) {
        const float v = (k < K) ? __ldg(s + k) : 0.f;
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        const float f = __bfloat162float(b);
        nrm += f * f;
        d[k] = b;
      }
    }
    if (norms != nullptr) {
      nrm = warp_sum(nrm);
      if (lane == 0) norms[dst_row0 + r] = nrm;
    }
  }
}

// dst[c, r] = bf16(src[r, c]) : [rows, cols] fp32 -> [cols_pad, ld_dst] bf16 (transpose), 32x32 tiles.
static __global__ void pack_transpose_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                      __nv_bfloat16* __restrict__ dst, int ld_dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? __ldg(src + size_t(r) * ld + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < ld_dst) dst[size_t(c) * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

static __global__ void fill_zero_f32_kernel(float* p, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x)
    p[i] = 0.f;
}

}  // namespace carel
