// umma_rate.cu — single-CTA timing probe: how long do NM tcgen05.mma of a given shape take, issued
// back to back, when they accumulate into 1, 2 or 4 independent TMEM accumulators?
// (operands are whatever is in shared memory: only the timing matters)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu -lcuda
#include <stdio.h>
#include <stdlib.h>
#include "../carel_vae_b200/csrc/common.cuh"
using namespace carel;


template <int N, int NACC, int B_MN, int A_TMEM, int WALK = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(int nm, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 44 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, B_MN);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
    const uint64_t da = make_sdesc_sw128(a_base, 16, 1024);
    const uint64_t db = B_MN ? make_sdesc_sw128(b_base, 8192, 1024) : make_sdesc_sw128(b_base, 16, 1024);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < nm; i += 4) {
        const int blk = WALK ? ((i >> 2) % 3) : 0;
        const uint64_t da_ = sdesc_advance(da, blk * 16384), db_ = sdesc_advance(db, blk * (N * 128));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t d = tmem + (A_TMEM ? 256 : 0) + ((NACC > 1) ? (ks % NACC) * (512 / 4 / (A_TMEM ? 2 : 1)) : 0);
          const uint32_t accf = (i > 0 || (NACC == 1 && ks > 0) || (NACC == 2 && ks > 1)) ? 1u : 0u;
          if (A_TMEM) umma_ts(d, tmem + ks * 8, sdesc_advance(db, ks * (B_MN ? 2048 : 32)), idesc, accf);
          else umma_ss(d, sdesc_advance(da_, ks * 32), sdesc_advance(db_, ks * (B_MN ? 2048 : 32)), idesc, accf);
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int B_MN, int A_TMEM, int WALK = 0>
void run(int nm, long long* d, int grid = 1) {
  auto k = rate_kernel<N, NACC, B_MN, A_TMEM, WALK>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<grid, 128, 200 * 1024>>>(nm, d);
  long long h[8];
  cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("grid=%3d walk=%d N=%3d nacc=%d b_mn=%d a_tmem=%d nm=%2d : issue %6lld clk, done %6lld clk  (%.1f clk/MMA; floor %d)\n", grid, WALK, N,
         NACC, B_MN, A_TMEM, nm, h[4], h[5], double(h[5]) / nm, N / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<64, 1, 0, 0>(64, d);  run<64, 2, 0, 0>(64, d);  run<64, 4, 0, 0>(64, d);
  run<128, 1, 0, 0>(64, d); run<128, 2, 0, 0>(64, d);
  run<256, 1, 0, 0>(64, d); run<256, 1, 1, 0>(64, d);
  run<256, 1, 1, 1>(64, d); run<256, 1, 0, 1>(64, d); run<144, 1, 1, 1>(64, d); run<64, 1, 1, 1>(64, d);
  run<64, 1, 0, 0>(16, d);  run<64, 1, 0, 0>(20, d); run<128, 1, 0, 0>(20, d); run<32, 1, 0, 0>(64, d);
  run<192, 1, 0, 0>(64, d);
  run<64, 1, 0, 0>(64, d, 148); run<128, 1, 0, 0>(64, d, 148); run<256, 1, 0, 0>(64, d, 148); run<256, 1, 1, 1>(64, d, 148);
  run<64, 1, 0, 0>(64, d, 74); run<128, 1, 0, 0>(64, d, 74);
  run<64, 1, 0, 0, 1>(64, d); run<128, 1, 0, 0, 1>(64, d); run<256, 1, 0, 0, 1>(64, d);
  return 0;
}
