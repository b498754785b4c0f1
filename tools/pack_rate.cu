// pack_rate.cu — dense [N,V] fp32 BoW target -> ELL list: candidate kernels against the HBM read roofline
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../carel_vae_b200/csrc -o pack_rate pack_rate.cu
// V0 = the register kernel of bow.cu (4 x 128-bit loads per lane in flight), V1 = the same with U loads and one
// ballot per batch in the all-zero case, V2 = warp-private cp.async.bulk rings (each warp streams its rows through
// its own STAGES x CHUNK floats of shared memory).  All variants must produce identical cols / vals / nnz.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace carel;

template <int U>
__global__ void pack_reg(const float* __restrict__ bow, int64_t ld, int N, int V, int cap, int* __restrict__ cols,
                         float* __restrict__ vals, int* __restrict__ nnz) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += gridDim.x * wpb) {
    const float* row = bow + size_t(r) * ld;
    int count = 0;
    const int V4 = V & ~3;
    for (int base = 0; base < V4; base += 128 * U) {
      float4 v[U];
      bool any = false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = base + u * 128 + lane * 4;
        v[u] = (j < V4) ? __ldcs(reinterpret_cast<const float4*>(row + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) any |= (v[u].x != 0.f) || (v[u].y != 0.f) || (v[u].z != 0.f) || (v[u].w != 0.f);
      if (__ballot_sync(0xffffffffu, any) == 0u) continue;
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
    if (lane == 0) nnz[r] = count;
  }
}

CAREL_DEVINL void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int STAGES, int CHUNK>
__global__ void pack_tma(const float* __restrict__ bow, int64_t ld, int N, int V, int cap, int* __restrict__ cols,
                         float* __restrict__ vals, int* __restrict__ nnz) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* buf = reinterpret_cast<float*>(smem_raw) + size_t(warp) * STAGES * CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(wpb) * STAGES * CHUNK * 4) + warp * STAGES;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int gw = blockIdx.x * wpb + warp, tw = gridDim.x * wpb;
  if (gw >= N) return;
  const int nchunks = (V + CHUNK - 1) / CHUNK;
  const int nrows = (N - gw + tw - 1) / tw;
  const int total = nrows * nchunks;
  auto issue = [&](int i) {
    if (i < total && lane == 0) {
      const int r = gw + (i / nchunks) * tw, c = i % nchunks, s = i % STAGES;
      const uint32_t bytes = uint32_t(min(CHUNK, V - c * CHUNK)) * 4u;
      mbar_expect_tx(bars + s, bytes);
      bulk_load_1d(buf + s * CHUNK, bow + size_t(r) * ld + size_t(c) * CHUNK, bytes, bars + s);
    }
  };
  for (int i = 0; i < STAGES; ++i) issue(i);
  int count = 0;
  for (int i = 0; i < total; ++i) {
    const int r = gw + (i / nchunks) * tw, c = i % nchunks, s = i % STAGES;
    mbar_wait(bars + s, uint32_t(i / STAGES) & 1u);
    const float4* src = reinterpret_cast<const float4*>(buf + s * CHUNK);
    const int valid = min(CHUNK, V - c * CHUNK);
#pragma unroll 2
    for (int k = 0; k < CHUNK / 128; k += 4) {
      float4 v[4];
      bool any = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = (k + u) * 128 + lane * 4;
        v[u] = (jj < valid) ? src[(k + u) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        any |= (v[u].x != 0.f) || (v[u].y != 0.f) || (v[u].z != 0.f) || (v[u].w != 0.f);
      }
      if (__ballot_sync(0xffffffffu, any) == 0u) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = c * CHUNK + (k + u) * 128 + lane * 4;
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
    __syncwarp();
    issue(i + STAGES);
    if (c == nchunks - 1) {
      if (lane == 0) nnz[r] = count;
      count = 0;
    }
  }
}

static void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) { printf("CUDA error %s: %s\n", what, cudaGetErrorString(e)); exit(1); }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 8192, V = argc > 2 ? atoi(argv[2]) : 30000, cap = 32, reps = 20;
  std::vector<float> h(size_t(N) * V, 0.f);
  srand(1);
  for (int r = 0; r < N; ++r)
    for (int t = 0; t < 24; ++t) h[size_t(r) * V + rand() % V] += 1.f / 24.f;
  float* d; int *cols, *nnz; float* vals;
  check(cudaMalloc(&d, h.size() * 4), "malloc");
  check(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice), "h2d");
  cudaMalloc(&cols, size_t(N) * cap * 4); cudaMalloc(&vals, size_t(N) * cap * 4); cudaMalloc(&nnz, size_t(N) * 4);
  std::vector<int> ref_cols(size_t(N) * cap), ref_nnz(N), got_cols(size_t(N) * cap), got_nnz(N);
  std::vector<float> ref_vals(size_t(N) * cap), got_vals(size_t(N) * cap);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = double(N) * V * 4 / 1e9;
  bool have_ref = false;
  auto run = [&](const char* name, auto launch) {
    cudaMemset(cols, 0xff, size_t(N) * cap * 4); cudaMemset(vals, 0xff, size_t(N) * cap * 4); cudaMemset(nnz, 0xff, size_t(N) * 4);
    for (int i = 0; i < 3; ++i) launch();
    check(cudaDeviceSynchronize(), name);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    check(cudaDeviceSynchronize(), name);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(got_cols.data(), cols, got_cols.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(got_vals.data(), vals, got_vals.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(got_nnz.data(), nnz, got_nnz.size() * 4, cudaMemcpyDeviceToHost);
    const char* verdict = "reference";
    if (!have_ref) { ref_cols = got_cols; ref_vals = got_vals; ref_nnz = got_nnz; have_ref = true; }
    else {
      bool same = got_nnz == ref_nnz;
      for (int r = 0; r < N && same; ++r)
        for (int s = 0; s < ref_nnz[r] && s < cap; ++s)
          same = same && got_cols[size_t(r) * cap + s] == ref_cols[size_t(r) * cap + s] &&
                 memcmp(&got_vals[size_t(r) * cap + s], &ref_vals[size_t(r) * cap + s], 4) == 0;
      verdict = same ? "identical" : "MISMATCH";
    }
    printf("%-44s %8.1f us  %7.1f GB/s  %s\n", name, ms / reps * 1e3, gb / (ms / reps * 1e-3), verdict);
  };
  const int64_t ld = V;
  run("V0 reg U=4, 8 warps, grid 1024", [&] { pack_reg<4><<<1024, 256>>>(d, ld, N, V, cap, cols, vals, nnz); });
  run("V1 reg U=8, 8 warps, grid 1024", [&] { pack_reg<8><<<1024, 256>>>(d, ld, N, V, cap, cols, vals, nnz); });
  run("V1 reg U=8, 4 warps, grid 2048", [&] { pack_reg<8><<<2048, 128>>>(d, ld, N, V, cap, cols, vals, nnz); });
  run("V1 reg U=4, 4 warps, grid 2048", [&] { pack_reg<4><<<2048, 128>>>(d, ld, N, V, cap, cols, vals, nnz); });
  run("V1 reg U=8, 16 warps, grid 512", [&] { pack_reg<8><<<512, 512>>>(d, ld, N, V, cap, cols, vals, nnz); });
  run("V1 reg U=8, 8 warps, grid 148*4 (loop)", [&] { pack_reg<8><<<148 * 4, 256>>>(d, ld, N, V, cap, cols, vals, nnz); });
#define TMA_CASE(S, C, W, G)                                                                                       \
  {                                                                                                                \
    const size_t sm = size_t(W) * S * C * 4 + size_t(W) * S * 8;                                                   \
    check(cudaFuncSetAttribute(pack_tma<S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)), "attr");     \
    char nm[96];                                                                                                   \
    snprintf(nm, sizeof nm, "V2 tma stages=%d chunk=%d warps=%d grid=%d", S, C, W, G);                             \
    run(nm, [&] { pack_tma<S, C><<<G, W * 32, sm>>>(d, ld, N, V, cap, cols, vals, nnz); });                        \
  }
  TMA_CASE(4, 1024, 8, 148)
  TMA_CASE(3, 1024, 14, 148)
  TMA_CASE(3, 1024, 16, 148)
  TMA_CASE(4, 1024, 12, 148)
  TMA_CASE(3, 2048, 8, 148)
  TMA_CASE(2, 2048, 14, 148)
  TMA_CASE(4, 1024, 7, 296)
  TMA_CASE(3, 1024, 7, 296)
  TMA_CASE(6, 512, 14, 148)
  TMA_CASE(4, 512, 14, 148)
  TMA_CASE(4, 512, 7, 592)
  return 0;
}
