// mufu_rate.cu — MUFU.EX2 / FFMA / FFMA2 / FMNMX3 issue throughput per SM on this GPU
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  float2 b[4];
  for (int i = 0; i < 4; ++i) b[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (MODE == 4) a[i] = fmaf(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);
      if (MODE == 5) a[i] = a[i] + a[(i + 1) & 7];
      if (MODE == 8) a[i] = a[i] * a[(i + 1) & 7];
    }
    if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { b[i] = __fadd2_rn(b[i], b[(i + 1) & 3]); b[i] = __fadd2_rn(b[i], b[(i + 2) & 3]); }
    }
    if (MODE == 7) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { b[i] = __fmul2_rn(b[i], b[(i + 1) & 3]); b[i] = __fmul2_rn(b[i], b[(i + 2) & 3]); }
    }
    if (MODE == 9) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { b[i] = __ffma2_rn(b[i], make_float2(1.0001f, 1.0001f), b[(i + 1) & 3]); b[i] = __ffma2_rn(b[i], make_float2(1.0001f, 1.0001f), b[(i + 2) & 3]); }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { b[i] = __ffma2_rn(b[i], b[(i + 1) & 3], b[(i + 2) & 3]); b[i] = __ffma2_rn(b[i], b[(i + 1) & 3], b[(i + 2) & 3]); }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 4; ++i) s += b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
  const char* names[] = {"MUFU.EX2", "FFMA imm", "FFMA2 (3-reg)", "FMNMX3", "FFMA 3-reg", "FADD", "FADD2", "FMUL2", "FMUL", "FFMA2 imm"};
  for (int threads : {512}) {
    for (int mode = 0; mode < 10; ++mode) {
      const int iters = 4096;
      if (mode == 0) k<0><<<148, threads>>>(o, iters, c);
      if (mode == 1) k<1><<<148, threads>>>(o, iters, c);
      if (mode == 2) k<2><<<148, threads>>>(o, iters, c);
      if (mode == 3) k<3><<<148, threads>>>(o, iters, c);
      if (mode == 4) k<4><<<148, threads>>>(o, iters, c);
      if (mode == 5) k<5><<<148, threads>>>(o, iters, c);
      if (mode == 6) k<6><<<148, threads>>>(o, iters, c);
      if (mode == 7) k<7><<<148, threads>>>(o, iters, c);
      if (mode == 8) k<8><<<148, threads>>>(o, iters, c);
      if (mode == 9) k<9><<<148, threads>>>(o, iters, c);
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double warp_instr = double(iters) * 8 * (threads / 32);
      printf("%-14s threads=%3d: %8lld clk, %.2f clk per warp-instr per SM -> %.1f lane-ops/clk/SM%s\n", names[mode], threads, h,
             h / warp_instr, 32.0 * warp_instr / h * ((mode == 2 || mode == 6 || mode == 7 || mode == 9) ? 2 : 1), (mode == 2 || mode == 6 || mode == 7 || mode == 9) ? " (2 per lane)" : "");
    }
  }
  return 0;
}
