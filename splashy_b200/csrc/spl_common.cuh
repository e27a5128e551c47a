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
  // z-slabs over peer memory: stamp up to which the first CTA of a launch has completed gA /
  // gBM from the peers' boxes (device-local; the other CTAs of the launch poll these)
  unsigned long long fillA, fillB;
};

// ---- z-slabs: the PCG iteration over peer memory (NVLink, CUDA IPC) ---------------------------
// Every rank owns a FusedBox that its peers map.  The last CTA of phase A / phase B stores the
// rank's partial scalar(s) and then the iteration stamp into every peer's box (release:
// __threadfence_system between the two); the kernels that consume the scalars wait for the
// stamps in their OWN box (local memory) and copy the values into the local slot arrays, so
// every rank still sums the slots in rank order.  A rank never waits for a kernel of its own
// GPU, and a peer is at most one iteration ahead (it needs this rank's scalars twice per
// iteration), hence four value slots indexed by the stamp.  hseq: the neighbour below / above
// has stored its boundary plane of the new search direction into this rank's halo.
// A scalar travels as ONE 16-byte slot {value, tag} with tag = stamp XOR bits(value): the
// reader accepts a slot only when tag XOR bits(value) equals the stamp it waits for, so a torn
// or reordered pair can never be taken for the value of that stamp -- no fence on either side
// of the exchange, the latency is one NVLink store.
struct alignas(16) TagSlot {
  double v;
  unsigned long long tag;
};
struct FusedBox {
  TagSlot A[4][MAXW];
  TagSlot B[4][MAXW][2];
  unsigned long long hseq[2];
};
struct Fused {                    // kernel argument
  FusedBox* box[MAXW];            // box[rank] = this rank's own box
  unsigned long long stamp;       // this iteration (monotonic over the life of the context)
  unsigned long long wait_b;      // stamp of the phase B whose slots are still travelling (0: none)
  int on;                         // 0: the slots are completed by separate gathers
  int halo;                       // 1: the halo planes arrive as peer stores too (hseq stamps)
  int mode;                       // experiments (SPL_FUSED_MODE): 1 fence after the poll, 4 no chunk remap
};

// bounded wait for *f >= want: a missing peer must abort the launch, not hang the GPU.  The
// poll is an acquire load at system scope (a trailing __threadfence_system in every CTA of a
// 2000-CTA launch is measurably slower).
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const unsigned long long* f, unsigned long long want,
                                           int fence_mode = 0) {
  unsigned long long t0 = 0;
  for (unsigned spin = 1; ld_acquire_sys(f) < want; ++spin) {
    if ((spin & 1023u) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ull) __trap();
    }
  }
  if (fence_mode & 1) __threadfence_system();
}
__device__ __forceinline__ void tag_store(TagSlot* p, double v, unsigned long long stamp) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(bits), "l"(bits ^ stamp)
               : "memory");
}
// bounded wait for the slot of `stamp`; returns its value
__device__ __forceinline__ double tag_wait(const TagSlot* p, unsigned long long stamp) {
  unsigned long long t0 = 0;
  for (unsigned spin = 1;; ++spin) {
    unsigned long long bits, tag;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(p) : "memory");
    if ((bits ^ tag) == stamp) return __longlong_as_double((long long)bits);
    if ((spin & 1023u) == 0u) {      // a missing peer must abort the launch, not hang the GPU
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ull) __trap();
    }
  }
}
// one thread: complete gBM[slot][*] from the peers' phase B of iteration fz.wait_b
__device__ __forceinline__ void fused_fill_B(const Fused& fz, Scal* sc, int slot) {
  if (!fz.on || fz.wait_b == 0) return;
  const FusedBox* mine = fz.box[sc->rank];
  const int b = (int)(fz.wait_b & 3ull);
  for (int w = 0; w < sc->world; ++w) {
    if (w == sc->rank) continue;
    sc->gBM[slot][w][0] = tag_wait(&mine->B[b][w][0], fz.wait_b);
    sc->gBM[slot][w][1] = tag_wait(&mine->B[b][w][1], fz.wait_b);
  }
}
// one thread: complete gA[*] from the peers' phase A of this iteration
__device__ __forceinline__ void fused_fill_A(const Fused& fz, Scal* sc) {
  if (!fz.on) return;
  const FusedBox* mine = fz.box[sc->rank];
  const int b = (int)(fz.stamp & 3ull);
  for (int w = 0; w < sc->world; ++w)
    if (w != sc->rank) sc->gA[w] = tag_wait(&mine->A[b][w], fz.stamp);
}
// Two-level wait.  A system-scope acquire in every CTA of a 2000-CTA launch costs ~60 us per
// launch on a busy SM (it drains the SM's outstanding memory traffic), so only the leader --
// the first CTA of the launch -- talks to the peers' stamps; it completes the local slot arrays
// and publishes a device-local stamp that every other CTA polls with plain volatile loads
// (the slots themselves are read with ld.cg afterwards).
__device__ __forceinline__ void spin_local(const unsigned long long* f, unsigned long long want) {
  const volatile unsigned long long* vf = f;
  unsigned long long t0 = 0;
  for (unsigned spin = 1; *vf < want; ++spin) {
    if ((spin & 1023u) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fused_sync_B(const Fused& fz, Scal* sc, int slot, bool leader) {
  if (!fz.on || fz.wait_b == 0) return;
  if (leader) {
    fused_fill_B(fz, sc, slot);
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(&sc->fillB) = fz.wait_b;
  } else {
    spin_local(&sc->fillB, fz.wait_b);
  }
}
__device__ __forceinline__ void fused_sync_A(const Fused& fz, Scal* sc, bool leader) {
  if (!fz.on) return;
  if (leader) {
    fused_fill_A(fz, sc);
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(&sc->fillA) = fz.stamp;
  } else {
    spin_local(&sc->fillA, fz.stamp);
  }
}
// one thread (of the last CTA): this rank's scalar(s) of iteration fz.stamp -> every peer
__device__ __forceinline__ void fused_push_A(const Fused& fz, const Scal* sc, double v) {
  if (!fz.on) return;
  const int b = (int)(fz.stamp & 3ull), r = sc->rank;
  for (int w = 0; w < sc->world; ++w)
    if (w != r) tag_store(&fz.box[w]->A[b][r], v, fz.stamp);
}
__device__ __forceinline__ void fused_push_B(const Fused& fz, const Scal* sc, double v0, double v1) {
  if (!fz.on) return;
  const int b = (int)(fz.stamp & 3ull), r = sc->rank;
  for (int w = 0; w < sc->world; ++w)
    if (w != r) {
      tag_store(&fz.box[w]->B[b][r][0], v0, fz.stamp);
      tag_store(&fz.box[w]->B[b][r][1], v1, fz.stamp);
    }
}

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
