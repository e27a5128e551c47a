// Micro-benchmark: does the z-marching tile access pattern of k_phaseA by itself
// cost HBM bandwidth on B200?  Pointwise work with phase A's byte mix
// (reads a4 b4 c1 x8, writes o4 p4 x8 = 33 B/cell) under different thread->cell maps.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tiles tiles.cu ; run: ./tiles
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void work(const float* a, const float* b, const uint8_t* c, double* x,
                                     float* o, float* p, long long idx) {
  const float4 av = *reinterpret_cast<const float4*>(a + idx);
  const float4 bv = *reinterpret_cast<const float4*>(b + idx);
  const unsigned cv = *reinterpret_cast<const unsigned*>(c + idx);
  double2 x0 = *reinterpret_cast<const double2*>(x + idx);
  double2 x1 = *reinterpret_cast<const double2*>(x + idx + 2);
  const float s = (cv & 1) ? 1.f : 0.5f;
  float4 ov = make_float4(av.x + s * bv.x, av.y + s * bv.y, av.z + s * bv.z, av.w + s * bv.w);
  float4 pv = make_float4(av.x - bv.x, av.y - bv.y, av.z - bv.z, av.w - bv.w);
  x0.x += av.x; x0.y += av.y; x1.x += av.z; x1.y += av.w;
  *reinterpret_cast<float4*>(o + idx) = ov;
  *reinterpret_cast<float4*>(p + idx) = pv;
  *reinterpret_cast<double2*>(x + idx) = x0;
  *reinterpret_cast<double2*>(x + idx + 2) = x1;
}

__global__ void __launch_bounds__(256) k_linear(const float* a, const float* b, const uint8_t* c,
                                                double* x, float* o, float* p, long long nq) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nq;
       t += (long long)gridDim.x * blockDim.x)
    work(a, b, c, x, o, p, t * 4);
}

// block (BX lanes-of-4-cells, BY rows); marches zc planes
template <int BX, int BY, bool SYNC>
__global__ void __launch_bounds__(BX * BY) k_tile(const float* a, const float* b, const uint8_t* c,
                                                 double* x, float* o, float* p, int nx, int ny,
                                                 int nz, int zc) {
  const int i0 = (blockIdx.x * BX + threadIdx.x) * 4;
  const int j = blockIdx.y * BY + threadIdx.y;
  const int kb = blockIdx.z * zc, ke = min(kb + zc, nz);
  const long long plane = (long long)nx * ny;
  long long idx = ((long long)kb * ny + j) * nx + i0;
  for (int k = kb; k < ke; ++k, idx += plane) {
    if (i0 < nx && j < ny) work(a, b, c, x, o, p, idx);
    if (SYNC) __syncthreads();
  }
}

int main() {
  const int n = 512;
  const long long N = (long long)n * n * n;
  float *a, *b, *o, *p; uint8_t* c; double* x;
  CK(cudaMalloc(&a, N * 4)); CK(cudaMalloc(&b, N * 4)); CK(cudaMalloc(&o, N * 4));
  CK(cudaMalloc(&p, N * 4)); CK(cudaMalloc(&c, N)); CK(cudaMalloc(&x, N * 8));
  CK(cudaMemset(a, 0, N * 4)); CK(cudaMemset(b, 0, N * 4)); CK(cudaMemset(c, 1, N));
  CK(cudaMemset(x, 0, N * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms) {
    printf("%-34s %7.3f ms  %7.1f GB/s\n", name, ms, N * 33.0 / (ms * 1e-3) / 1e9);
  };
#define TIME(name, launch)                                   \
  { for (int w = 0; w < 3; ++w) { launch; }                  \
    cudaEventRecord(e0); for (int r = 0; r < 10; ++r) { launch; } cudaEventRecord(e1); \
    CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); report(name, ms / 10); }
  TIME("linear grid-stride 1184x256", (k_linear<<<1184, 256>>>(a, b, c, x, o, p, N / 4)));
  TIME("linear grid-stride 2368x256", (k_linear<<<2368, 256>>>(a, b, c, x, o, p, N / 4)));
  for (int zc : {8, 27, 64, 512}) {
    char nm[64];
    dim3 g1(n / 128, n / 8, (n + zc - 1) / zc);
    snprintf(nm, 64, "tile 128x8 zc=%d", zc);
    TIME(nm, (k_tile<32, 8, false><<<g1, dim3(32, 8)>>>(a, b, c, x, o, p, n, n, n, zc)));
    snprintf(nm, 64, "tile 128x8 zc=%d +sync", zc);
    TIME(nm, (k_tile<32, 8, true><<<g1, dim3(32, 8)>>>(a, b, c, x, o, p, n, n, n, zc)));
    dim3 g2(n / 512, n / 2, (n + zc - 1) / zc);
    snprintf(nm, 64, "tile 512x2 zc=%d", zc);
    TIME(nm, (k_tile<128, 2, false><<<g2, dim3(128, 2)>>>(a, b, c, x, o, p, n, n, n, zc)));
    dim3 g3(n / 256, n / 4, (n + zc - 1) / zc);
    snprintf(nm, 64, "tile 256x4 zc=%d", zc);
    TIME(nm, (k_tile<64, 4, false><<<g3, dim3(64, 4)>>>(a, b, c, x, o, p, n, n, n, zc)));
    dim3 g4(n / 512, n / 4, (n + zc - 1) / zc);
    snprintf(nm, 64, "tile 512x4 (512 thr) zc=%d", zc);
    TIME(nm, (k_tile<128, 4, false><<<g4, dim3(128, 4)>>>(a, b, c, x, o, p, n, n, n, zc)));
  }
  return 0;
}
