// pipe_overlap.cu — do MUFU.EX2, packed FMA and FMNMX3 overlap on one SM sub-partition?  Each variant issues, per loop
// iteration and thread, NM ex2 + NF ffma2 + NA fmnmx3 on independent registers; 16 warps per SM (4 per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_overlap pipe_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int NM, int NF, int NA>
__global__ void k(float* out, int iters, long long* clk) {
  float m[8], a[8];
  float2 f[8];
  for (int i = 0; i < 8; ++i) { m[i] = threadIdx.x * 1e-3f + i * 0.1f; a[i] = m[i] + 1.f; f[i] = make_float2(m[i], a[i]); }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NM) m[i] = ex2(m[i]);
      if (i < NF) f[i] = __ffma2_rn(f[i], f[(i + 1) & 7], f[(i + 2) & 7]);
      if (i < NA) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += m[i] + a[i] + f[i].x + f[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int NM, int NF, int NA>
void run(float* o, long long* c) {
  const int iters = 4096;
  k<NM, NF, NA><<<148, 512>>>(o, iters, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("per iteration and thread: %d ex2 + %d ffma2 + %d fmnmx3 : %.1f clk per iteration per SM sub-partition-warp-round (%lld clk total)\n",
         NM, NF, NA, double(h) / iters, h);
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 8);
  run<8, 0, 0>(o, c); run<0, 8, 0>(o, c); run<0, 0, 8>(o, c);
  run<8, 8, 0>(o, c); run<8, 0, 8>(o, c); run<0, 8, 8>(o, c); run<8, 8, 8>(o, c);
  run<4, 8, 4>(o, c);
  return 0;
}
