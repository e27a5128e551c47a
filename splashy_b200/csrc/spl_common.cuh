// spl_common.cuh -- shared device-side definitions for libsplashy_cuda (sm_100a)
//
// Data layout in HBM: every field is a dense box
// nx * ny * (nz + 2*hz) with x fastest; kernels receive the pointer of plane
// k = 0, so idx = (k*ny + j)*nx + i is valid for k in [-hz, nz + hz).  The hz
// halo planes on either side hold the neighbouring z-slab's planes (world > 1)
// or stay zero / SPL_ABSENT at the global ends.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spl {

// Cell.fs:9 media codes (+ 255 = absent from the reference's sparse Dictionary)
constexpr uint8_t AIR = 0, FLUID = 1, SOLID = 2, ABSENT = 255;

// per-cell stencil code: bit d set <=> neighbour d exists and is not Solid
// (Pressure.fs:34-38,53); bit 6 <=> the cell is Fluid, i.e. a pressure unknown
// (Pressure.fs:42-43).  diag(A) = popc(code & 63).
constexpr int B_XM = 1, B_XP = 2, B_YM = 4, B_YP = 8, B_ZM = 16, B_ZP = 32,
              B_FLUID = 64;

constexpr int MAXW = 16;       // max ranks (z-slabs)
constexpr int TILE_Y = 8;      // rows per CTA in the marching stencil kernels
constexpr int MAX_PARTIALS = 1 << 16;
constexpr int MAX_SBUF = 16;   // ring of search-direction buffers (deferred x update)

struct Geo {
  int nx, ny, nz;      // local box
  int hz;              // halo planes on each z side
  int z0, nzg;         // first global plane of this slab, global plane count
  long long plane;     // nx * ny
  long long ncell;     // nx * ny * nz
};

// PCG scalars living in device memory.  Slots indexed by rank are filled
// locally by the last CTA of a kernel and, when world > 1, completed by an
// in-place ncclAllGather, so every consumer sums them in rank order
// (deterministic for a fixed launch geometry).
struct Scal {
  double gA[MAXW];        // s.q
  double gBM[2][MAXW][2]; // slot [it & 1]: {rho_it = r.z, max |r|} per rank
  double gD[MAXW];        // max |div| (check kernel)
  double bmax;            // max |b| of the current solve
  double tol;
  double rmax;            // max |r| at exit
  unsigned long long nonfinite;   // summed over ranks on the host
  unsigned long long escapes;     // advection departures outside halo
  int world, rank;
  double alpha_hist[MAX_SBUF];    // step length of iteration it at [it % m]
  int done, iters, count, breakdown, converged;
  int flushed;                    // iterations whose alpha s term is already in x
  unsigned ticketA, ticketB, ticketC, ticketX;
  unsigned wl_count;              // advection worklist length
};

template <typename T, int N>
struct alignas((sizeof(T) * N > 16) ? 16 : sizeof(T) * N) VecT {
  T v[N];
};

template <typename T, int N>
__device__ __forceinline__ VecT<T, N> vload(const T* __restrict__ p, long long idx) {
  return *reinterpret_cast<const VecT<T, N>*>(p + idx);
}
template <typename T, int N>
__device__ __forceinline__ void vstore(T* __restrict__ p, long long idx,
                                       const VecT<T, N>& v) {
  *reinterpret_cast<VecT<T, N>*>(p + idx) = v;
}
template <int N>
__device__ __forceinline__ VecT<uint8_t, N> cload(const uint8_t* __restrict__ p,
                                                  long long idx) {
  return *reinterpret_cast<const VecT<uint8_t, N>*>(p + idx);
}

// 1 / diag for the Jacobi preconditioner; 0 on non-unknowns.
template <typename T>
__device__ __forceinline__ T inv_diag(unsigned code) {
  const int n = __popc(code & 63u);
  return ((code & B_FLUID) && n > 0) ? T(1) / T(n) : T(0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum / max of one double per thread; result valid in thread 0.
// Fixed tree => deterministic.  `red` is a __shared__ double[32].
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int nthr = blockDim.x * blockDim.y;
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (tid < 32) {
    t = (tid < (nthr + 31) / 32) ? red[tid] : 0.0;
    t = warp_sum(t);
  }
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int nthr = blockDim.x * blockDim.y;
  v = warp_max(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (tid < 32) {
    t = (tid < (nthr + 31) / 32) ? red[tid] : 0.0;
    t = warp_max(t);
  }
  return t;
}

// "Last CTA done" hand-off: returns true (to every thread) in the CTA that
// arrived last; that CTA then reduces the per-CTA partials in index order.
__device__ __forceinline__ bool last_block(unsigned* ticket, unsigned nblocks) {
  __shared__ int s_last;
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == nblocks - 1u);
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__device__ __forceinline__ unsigned linear_block() {
  return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
}
__device__ __forceinline__ unsigned num_blocks() {
  return gridDim.x * gridDim.y * gridDim.z;
}

}  // namespace spl
