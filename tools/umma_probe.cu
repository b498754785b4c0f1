// umma_probe.cu — standalone semantic probe for the sm_100a building blocks
// (TMA SW128 loads, K-major / MN-major UMMA smem descriptors, manual-swizzle
// st.shared A operand, A-from-TMEM).  Single CTA, no pipelining: it only checks
// that the layouts mean what csrc/common.cuh says they mean.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../carel_vae_b200/csrc/common.cuh"
#include "../carel_vae_b200/csrc/tmap.h"

using namespace carel;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

// C1[128, BN] = A[128,K] * B[BN,K]^T  (rows of B start at brow0)
// if chain: P = bf16(C1 * scale) (BN must be 128, K must be 128)
//   C2[128, K] = P * B[brow0.., :]   via smem A (manual swizzle) + MN-major B
//   C3[128, K] = same via A from TMEM
template <int BN>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             float* C1, float* C2, float* C3, int K, int brow0, int chain, float scale) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int KB = K / 64;
  uint8_t* sA = smem;                    // KB x [128 x 128B]
  uint8_t* sB = sA + KB * 16384;         // KB x [BN x 128B]
  uint8_t* sP = sB + KB * BN * 128;      // 2 x [128 x 128B]
  __shared__ uint64_t bar_load, bar_mma, bar_mma2;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_mma2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, KB * (16384 + BN * 128));
    for (int kb = 0; kb < KB; ++kb) {
      tma_load_2d(sA + kb * 16384, &tmA, &bar_load, kb * 64, 0);
      for (int nb = 0; nb < BN / 128; ++nb)
        tma_load_2d(sB + kb * BN * 128 + nb * 16384, &tmB, &bar_load, kb * 64, brow0 + nb * 128);
    }
  }
  mbar_wait(&bar_load, 0);
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    for (int kb = 0; kb < KB; ++kb)
      for (int k16 = 0; k16 < 4; ++k16) {
        uint64_t da = make_sdesc_sw128(smem_u32(sA + kb * 16384) + k16 * 32, 16, 1024);
        uint64_t db = make_sdesc_sw128(smem_u32(sB + kb * BN * 128) + k16 * 32, 16, 1024);
        umma_ss(tmem, da, db, idesc, (kb | k16) ? 1u : 0u);
      }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_addr + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) C1[row * BN + c * 32 + i] = __uint_as_float(r[i]);
    if (chain) {
      // P -> smem (K-major SW128: k-block of 64 per 16 KB; row pitch 128 B; 16 B chunks XOR row%8)
      const int kb = (c * 32) / 64;
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[q * 8 + 0]) * scale, __uint_as_float(r[q * 8 + 1]) * scale);
        v.y = pack_bf16x2(__uint_as_float(r[q * 8 + 2]) * scale, __uint_as_float(r[q * 8 + 3]) * scale);
        v.z = pack_bf16x2(__uint_as_float(r[q * 8 + 4]) * scale, __uint_as_float(r[q * 8 + 5]) * scale);
        v.w = pack_bf16x2(__uint_as_float(r[q * 8 + 6]) * scale, __uint_as_float(r[q * 8 + 7]) * scale);
        const int chunk = ((c * 32) % 64) / 8 + q;
        *reinterpret_cast<uint4*>(sP + kb * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
      }
      // P -> TMEM (packed bf16x2: 16 columns per 32 k)
      uint32_t pk[16];
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * scale, __uint_as_float(r[2 * i + 1]) * scale);
      tmem_st16(lane_addr + 256 + c * 16, pk);
    }
  }
  if (chain) {
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc2 = make_idesc_bf16(128, 128, 0, 1);  // B MN-major, N = K(feat) = 128
      for (int k16 = 0; k16 < 8; ++k16) {
        uint64_t da = make_sdesc_sw128(smem_u32(sP + (k16 / 4) * 16384) + (k16 % 4) * 32, 16, 1024);
        uint64_t db = make_sdesc_sw128(smem_u32(sB) + k16 * 2048, 16384, 1024);
        umma_ss(tmem + 128, da, db, idesc2, k16 ? 1u : 0u);
      }
      for (int k16 = 0; k16 < 8; ++k16) {
        uint64_t db = make_sdesc_sw128(smem_u32(sB) + k16 * 2048, 16384, 1024);
        umma_ts(tmem + 384, tmem + 256 + k16 * 8, db, idesc2, k16 ? 1u : 0u);
      }
      umma_commit(&bar_mma2);
    }
    mbar_wait(&bar_mma2, 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_addr + 128 + c * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) C2[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
      tmem_ld32(lane_addr + 384 + c * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) C3[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int BN>
int run_case(int rows_total, int K, int brow0, int chain) {
  std::vector<float> S((size_t)rows_total * K);
  std::vector<__nv_bfloat16> Sb(S.size());
  srand(1234 + rows_total + K);
  for (size_t i = 0; i < S.size(); ++i) {
    float v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    Sb[i] = __float2bfloat16(v);
    S[i] = __bfloat162float(Sb[i]);
  }
  __nv_bfloat16* dS;
  float *dC1, *dC2, *dC3;
  CK(cudaMalloc(&dS, Sb.size() * 2));
  CK(cudaMemcpy(dS, Sb.data(), Sb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dC1, 128 * BN * 4));
  CK(cudaMalloc(&dC2, 128 * 128 * 4));
  CK(cudaMalloc(&dC3, 128 * 128 * 4));
  CK(cudaMemset(dC1, 0, 128 * BN * 4));
  CK(cudaMemset(dC2, 0, 128 * 128 * 4));
  CK(cudaMemset(dC3, 0, 128 * 128 * 4));
  CUtensorMap tm;
  if (make_tmap_bf16_sw128(&tm, dS, rows_total, K, K, 128)) {
    printf("tensor map encode failed\n");
    return 1;
  }
  const float scale = 0.05f;
  size_t smem = (K / 64) * (16384 + BN * 128) + 2 * 16384 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<BN><<<1, 128, smem>>>(tm, tm, dC1, dC2, dC3, K, brow0, chain, scale);
  CK(cudaDeviceSynchronize());
  std::vector<float> C1(128 * BN), C2(128 * 128), C3(128 * 128);
  CK(cudaMemcpy(C1.data(), dC1, C1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(C2.data(), dC2, C2.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(C3.data(), dC3, C3.size() * 4, cudaMemcpyDeviceToHost));
  auto sval = [&](int r, int k) -> float { return r < rows_total ? S[(size_t)r * K + k] : 0.f; };
  double e1 = 0, e2 = 0, e3 = 0, m1 = 0, m2 = 0;
  std::vector<float> R1(128 * BN);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < BN; ++j) {
      float acc = 0;
      for (int k = 0; k < K; ++k) acc += sval(i, k) * sval(brow0 + j, k);
      R1[i * BN + j] = acc;
      e1 = fmax(e1, fabs(acc - C1[i * BN + j]));
      m1 = fmax(m1, fabs(acc));
    }
  if (chain) {
    for (int i = 0; i < 128; ++i)
      for (int f = 0; f < K; ++f) {
        float acc = 0;
        for (int j = 0; j < 128; ++j) acc += bf(C1[i * BN + j] * scale) * sval(brow0 + j, f);
        e2 = fmax(e2, fabs(acc - C2[i * 128 + f]));
        e3 = fmax(e3, fabs(acc - C3[i * 128 + f]));
        m2 = fmax(m2, fabs(acc));
      }
  }
  printf("case BN=%d rows=%d K=%d brow0=%d chain=%d : gemm1 maxerr %.3e (max %.2f) | "
         "smemA+MNmajorB maxerr %.3e | tmemA maxerr %.3e (max %.3f)\n",
         BN, rows_total, K, brow0, chain, e1, m1, e2, e3, m2);
  cudaFree(dS); cudaFree(dC1); cudaFree(dC2); cudaFree(dC3);
  int bad = (e1 > 1e-3 * fmax(m1, 1.0));
  if (chain) bad |= (e2 > 2e-3 * fmax(m2, 1.0)) << 1 | (e3 > 2e-3 * fmax(m2, 1.0)) << 2;
  return bad;
}

int main() {
  int bad = 0;
  bad |= run_case<128>(256, 64, 128, 0);
  bad |= run_case<128>(256, 128, 128, 0);
  bad |= run_case<256>(384, 256, 128, 0);
  bad |= run_case<256>(300, 128, 128, 0);  // partially out-of-bounds B box (zero fill)
  bad |= run_case<128>(256, 128, 128, 1);  // chain
  bad |= run_case<128>(200, 128, 128, 1);  // chain with OOB rows
  printf(bad ? "PROBE FAILED (mask %d)\n" : "PROBE OK\n", bad);
  return bad ? 1 : 0;
}
