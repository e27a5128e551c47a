// Issue rate of packed fp32 (FFMA2) against scalar FFMA on sm_100a: nvcc -arch=sm_100a f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
  u64 p0 = threadIdx.x, p1 = 1, p2 = 2, p3 = 3, p4 = 4, p5 = 5, p6 = 6, p7 = 7;
  const float m = 1.0001f; const u64 pm = 0x3f8003473f800347ull;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0 = fmaf(a0, m, a1); a1 = fmaf(a1, m, a2); a2 = fmaf(a2, m, a3); a3 = fmaf(a3, m, a4);
        a4 = fmaf(a4, m, a5); a5 = fmaf(a5, m, a6); a6 = fmaf(a6, m, a7); a7 = fmaf(a7, m, a0);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p0 = fma2(p0, pm, p1); p1 = fma2(p1, pm, p2); p2 = fma2(p2, pm, p3); p3 = fma2(p3, pm, p4);
        p4 = fma2(p4, pm, p5); p5 = fma2(p5, pm, p6); p6 = fma2(p6, pm, p7); p7 = fma2(p7, pm, p0);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a3 + a7 + (float)(p0 ^ p3 ^ p7);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 512>>>(d, iters); else k<1><<<148 * 2, 512>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double inst = 148.0 * 2 * 512 / 32 * iters * 64;   // warp instructions
      printf("%s: %.3f ms, %.1f warp-instr/clk/SM at 1.9 GHz\n", mode ? "FFMA2" : "FFMA ", ms,
             inst / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
