// spl_pcg.cuh -- matrix-free PCG on the fluid-cell Laplacian (sm_100a)
//
// Replaces Pressure.calculate's assembly + direct solve (Pressure.fs:42-69):
// A is never formed; one PCG iteration is two kernels separated by
// the two global reductions CG needs:
//
//   phase A(it): s' = z + beta s ; q = A s' ; partial s'.q      17 B/cell (fp32)
//   phase B(it): r -= alpha q ; partial r.z, max|r|            13 B/cell
//   x flush    : x += sum_i alpha_i s_i   every m iterations   (4m + 16)/m B/cell
//
// The search directions of the last m iterations stay in a ring of m buffers
// (phase A writes s' into the next one instead of ping-ponging two), so the
// pressure accumulator x is touched once per m iterations instead of every
// iteration: m = 4 -> 8 B/cell/iteration instead of 20.  x is always fp64:
// with fp32 storage its rounding (eps |p|) is what limits the divergence-free
// residual on 256^3+ grids.
//
// HBM-bound; no tensor cores (no dense contraction on this path).
#pragma once
#include "spl_common.cuh"

namespace spl {

enum { PC_NONE = 0, PC_JACOBI = 1, PC_RBIC0 = 2, PC_MG = 3 };
// preconditioners whose z = M^-1 r is a stored field produced by separate kernels
// (phase A then reads z instead of r, phase B leaves rho = r.z to those kernels)
__host__ __device__ constexpr bool pc_explicit_z(int pc) { return pc == PC_RBIC0 || pc == PC_MG; }

// ---- scalar plumbing -------------------------------------------------------
__device__ __forceinline__ double sum_slots(const double* g, int world, int stride = 1) {
  double t = 0.0;
  for (int w = 0; w < world; ++w) t += __ldcg(g + (size_t)w * stride);
  return t;
}
// NaN / Inf in any slot propagates (returned as NaN)
__device__ __forceinline__ double max_slots(const double* g, int world, int stride = 1) {
  double t = 0.0;
  bool bad = false;
  for (int w = 0; w < world; ++w) {
    const double v = __ldcg(g + (size_t)w * stride);
    if (!(fabs(v) <= 1.7976931348623157e308)) bad = true;
    t = fmax(t, v);
  }
  return bad ? __longlong_as_double(0x7ff8000000000000LL) : t;
}
__device__ __forceinline__ double rho_of(const Scal* sc, int slot) {
  return sum_slots(&sc->gBM[slot][0][0], sc->world, 2);
}
__device__ __forceinline__ double rmax_of(const Scal* sc, int slot) {
  return max_slots(&sc->gBM[slot][0][1], sc->world, 2);
}

// 1/diag lookup: lut[n] = 1/n (correctly rounded), lut[0] = 0; indexed with
// popc(code & 63) for Fluid cells and 0 otherwise.
template <typename T>
__device__ __forceinline__ void fill_inv_lut(T* lut) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  if (tid < 8) lut[tid] = (tid >= 1 && tid <= 6) ? T(1) / T(tid) : T(0);
}
// codes of non-Fluid cells are 0 (k_codes), so the neighbour count is the index
__device__ __forceinline__ int lut_index(unsigned code) { return __popc(code & 63u); }

// Pressure.fs:44-55: (A s)_i = N_i s_i - sum over Fluid neighbours.  s is 0 on
// every non-Fluid cell and on every out-of-box position the kernels read, so the
// sum runs over all six neighbours unmasked; the differences form keeps smooth
// directions accurate in fp32 and the (6 - N_i) s_i term removes the neighbours
// that are Solid / absent.
template <typename T>
__device__ __forceinline__ T lap_cell(unsigned code, T s0, T xm, T xp, T ym, T yp, T zm, T zp) {
  const T d = ((s0 - xm) + (s0 - xp)) + ((s0 - ym) + (s0 - yp)) + ((s0 - zm) + (s0 - zp));
  const T miss = T(6 - __popc(code & 63u));
  const T a = d - miss * s0;
  return (code & B_FLUID) ? a : T(0);
}

// ---- phase A ---------------------------------------------------------------
// CTA = 32 x TILE_Y threads; a thread owns VEC consecutive x cells of one row
// and marches along z keeping the new search direction of planes c-1, c, c+1
// in registers.  +-x neighbours come from warp shuffles, +-y neighbours from a
// triple-buffered shared-memory tile (one barrier per plane); only the two
// CTA-edge rows and the CTA-edge lanes recompute a halo value from global
// memory.  Software pipeline: the raw operands (r, s, codes) of plane c+2 are
// requested at the top of iteration c, after those of plane c+1 were consumed,
// so a plane of loads per thread is in flight across the barrier and the
// stencil.  Every array is pulled from HBM once.
template <typename T, int VEC>
struct RawV {
  VecT<T, VEC> a, old;
  VecT<uint8_t, VEC> code;
};
template <typename T>
struct RawS {
  T a, old;
  unsigned code;
};

template <typename T, int VEC, int PC>
struct DirLoader {
  const uint8_t* __restrict__ codes;
  const T* __restrict__ zr;     // z (PC_RBIC0), r (PC_NONE / PC_JACOBI)
  const T* __restrict__ sold;
  const T* __restrict__ snew;   // only its z-halo planes are read
  const T* lut;                 // shared: 1/diag
  int nz;
  T beta;

  // planes outside [0, nz) belong to the neighbouring slab: the halo of the s'
  // buffer already holds the new direction (zeros at the global ends)
  __device__ __forceinline__ bool remote(int k) const { return k < 0 || k >= nz; }

  __device__ __forceinline__ RawV<T, VEC> load(long long idx, int k) const {
    RawV<T, VEC> r;
    if (remote(k)) {
      r.a = vload<T, VEC>(snew, idx);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { r.old.v[e] = T(0); r.code.v[e] = 0; }
    } else {
      r.a = vload<T, VEC>(zr, idx);
      r.old = vload<T, VEC>(sold, idx);
      r.code = cload<VEC>(codes, idx);
    }
    return r;
  }
  __device__ __forceinline__ VecT<T, VEC> combine(const RawV<T, VEC>& r, int k) const {
    if (remote(k)) return r.a;
    VecT<T, VEC> out;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T zz = (PC == PC_JACOBI) ? r.a.v[e] * lut[lut_index(r.code.v[e])] : r.a.v[e];
      out.v[e] = zz + beta * r.old.v[e];
    }
    return out;
  }
  __device__ __forceinline__ RawS<T> load1(long long idx, int k) const {
    RawS<T> r;
    if (remote(k)) {
      r.a = snew[idx]; r.old = T(0); r.code = 0;
    } else {
      r.a = zr[idx]; r.old = sold[idx]; r.code = codes[idx];
    }
    return r;
  }
  __device__ __forceinline__ T combine1(const RawS<T>& r, int k) const {
    if (remote(k)) return r.a;
    const T zz = (PC == PC_JACOBI) ? r.a * lut[lut_index(r.code)] : r.a;
    return zz + beta * r.old;
  }
};

// beta, stop decision: shared by both phase-A kernels (thread 0 of every CTA)
__device__ __forceinline__ void phaseA_scalars(Scal* sc, int par, int* stop_out, double* beta_out) {
  int stop = sc->done;
  double beta = 0.0;
  if (!stop) {
    const double rho1 = rho_of(sc, par ^ 1);   // rho_{it-1}
    const double rho2 = rho_of(sc, par);       // rho_{it-2}
    const double rmax = rmax_of(sc, par ^ 1);  // max|r| after iteration it-1
    const double bmax = sc->bmax;
    const bool bad = !(rmax == rmax);
    if (bad || rmax <= sc->tol * bmax || rho1 == 0.0) {
      stop = 1;
      if (linear_block() == 0) {
        sc->iters = sc->count;
        sc->rmax = rmax;
        sc->converged = bad ? 0 : 1;
        sc->breakdown = bad ? 1 : 0;
        __threadfence();
        sc->done = 1;
      }
    } else {
      beta = rho1 / rho2;     // rho2 = +inf before the first iteration
    }
  }
  *stop_out = stop;
  *beta_out = beta;
}

#ifndef SPL_PHASEA_MINB
#define SPL_PHASEA_MINB 4     // resident CTAs per SM for fp32 (fp64: half)
#endif
// TY = rows per CTA: 8 for 3-D boxes; 1 for the row-marching view of a 2-D box
// (nz = 1 is presented as nx x 1 x ny, so the march runs along y and every warp is
// independent -- the halo rows are outside the view and read as 0).
template <typename T, int VEC, int PC, int TY>
__global__ void __launch_bounds__(32 * TY, (TY == 1) ? 16 : ((sizeof(T) == 8) ? SPL_PHASEA_MINB / 2 : SPL_PHASEA_MINB))
k_phaseA(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ zr,
         const T* __restrict__ sold, T* __restrict__ snew, T* __restrict__ q,
         Scal* __restrict__ sc, double* __restrict__ partials, int par, int zchunk) {
  __shared__ double red[32];
  __shared__ double s_beta;
  __shared__ int s_stop;
  __shared__ T lut[8];
  __shared__ VecT<T, VEC> tile[3][TY + 2][32];
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<T>(lut);
  if (tid == 0) phaseA_scalars(sc, par, &s_stop, &s_beta);
  __syncthreads();
  if (s_stop) return;

  const DirLoader<T, VEC, PC> dl{codes, zr, sold, snew, lut, g.nz, T(s_beta)};
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * VEC;
  const int j = blockIdx.y * TY + ty;
  const bool active = (i0 < g.nx) && (j < g.ny);
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const long long row = g.nx, plane = g.plane;
  // CTA-edge rows recompute the neighbouring row, CTA-edge lanes the x
  // neighbour owned by the adjacent CTA.  Rows / columns outside the box must
  // read as 0 (the Laplacian is unmasked, see lap_cell).
  const int hrow_j = (ty == 0) ? j - 1 : j + 1;
  const bool hrow_ok = (TY > 1) && active && (ty == 0 || ty == TY - 1) && hrow_j >= 0 && hrow_j < g.ny;
  const long long hoff = (ty == 0) ? -row : row;
  const int hslot = (ty == 0) ? 0 : TY + 1;
  const bool has_e = active && ((lane == 0 && i0 > 0) || (lane == 31 && i0 + VEC < g.nx));
  const long long eoff = (lane == 0) ? -1 : VEC;

  VecT<T, VEC> zero;
  VecT<uint8_t, VEC> czero;
#pragma unroll
  for (int e = 0; e < VEC; ++e) { zero.v[e] = T(0); czero.v[e] = 0; }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    if (!active) tile[b][ty + 1][lane] = zero;                            // own row
    if (ty == 0 && !hrow_ok) tile[b][0][lane] = zero;                     // row j-1
    if (ty == TY - 1 && !hrow_ok) tile[b][TY + 1][lane] = zero;           // row j+1
  }

  double acc = 0.0;
  long long idx = ((long long)kb * g.ny + j) * row + i0;
  VecT<T, VEC> sm = zero, scur = zero, sp = zero;
  VecT<uint8_t, VEC> codec = czero;
  RawV<T, VEC> raw{zero, zero, czero}, hraw{zero, zero, czero};
  RawS<T> eraw{T(0), T(0), 0u};
  T leftc = T(0), rightc = T(0);

  // ---- prologue: planes kb-1 and kb combined, plane kb+1 requested ----------
  if (active) {
    sm = dl.combine(dl.load(idx - plane, kb - 1), kb - 1);
    const RawV<T, VEC> r0 = dl.load(idx, kb);
    scur = dl.combine(r0, kb);
    codec = r0.code;
    tile[kb % 3][ty + 1][lane] = scur;
    if (hrow_ok) tile[kb % 3][hslot][lane] = dl.combine(dl.load(idx + hoff, kb), kb);
    if (has_e) {
      const T e = dl.combine1(dl.load1(idx + eoff, kb), kb);
      if (lane == 0) leftc = e; else rightc = e;
    }
    raw = dl.load(idx + plane, kb + 1);
    if (kb + 1 < ke) {
      if (hrow_ok) hraw = dl.load(idx + plane + hoff, kb + 1);
      if (has_e) eraw = dl.load1(idx + plane + eoff, kb + 1);
    }
  }

  for (int c = kb; c < ke; ++c, idx += plane) {
    const int tb = c % 3, tn = (c + 1) % 3;
    VecT<uint8_t, VEC> coden = czero;
    T leftn = T(0), rightn = T(0);
    if (active) {
      // plane c+1: operands were requested one iteration ago
      sp = dl.combine(raw, c + 1);
      coden = raw.code;
      if (c + 1 < ke) {
        tile[tn][ty + 1][lane] = sp;
        if (hrow_ok) tile[tn][hslot][lane] = dl.combine(hraw, c + 1);
        if (has_e) {
          const T e = dl.combine1(eraw, c + 1);
          if (lane == 0) leftn = e; else rightn = e;
        }
      }
      // request plane c+2
      if (c + 2 <= ke) raw = dl.load(idx + 2 * plane, c + 2);
      if (c + 2 < ke) {
        if (hrow_ok) hraw = dl.load(idx + 2 * plane + hoff, c + 2);
        if (has_e) eraw = dl.load1(idx + 2 * plane + eoff, c + 2);
      }
    }
    __syncthreads();   // tile[tb] (plane c) was completed before the previous barrier
    const T lsh = __shfl_up_sync(0xffffffffu, scur.v[VEC - 1], 1);
    const T rsh = __shfl_down_sync(0xffffffffu, scur.v[0], 1);
    const T left = (lane != 0) ? lsh : leftc;
    const T right = (lane != 31) ? rsh : rightc;
    if (active) {
      const VecT<T, VEC> sym = tile[tb][ty][lane];
      const VecT<T, VEC> syp = tile[tb][ty + 2][lane];
      VecT<T, VEC> qv;
      T part = T(0);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const T s0 = scur.v[e];
        const T xm = (e > 0) ? scur.v[e > 0 ? e - 1 : 0] : left;
        const T xp = (e < VEC - 1) ? scur.v[e < VEC - 1 ? e + 1 : 0] : right;
        const T a = lap_cell<T>(codec.v[e], s0, xm, xp, sym.v[e], syp.v[e], sm.v[e], sp.v[e]);
        qv.v[e] = a;
        part += s0 * a;
      }
      acc += (double)part;   // VEC products in working precision, accumulated in fp64
      vstore<T, VEC>(q, idx, qv);
      vstore<T, VEC>(snew, idx, scur);
    }
    sm = scur;
    scur = sp;
    codec = coden;
    leftc = leftn;
    rightc = rightn;
  }

  const double bs = block_sum(acc, red);
  const unsigned nb = num_blocks();
  if (tid == 0) partials[linear_block()] = bs;
  if (last_block(&sc->ticketA, nb)) {
    double t = 0.0;
    for (unsigned b = tid; b < nb; b += 32 * TY) t += __ldcg(partials + b);
    t = block_sum(t, red);
    if (tid == 0) sc->gA[sc->rank] = t;
  }
}

// ---- phase A, cp.async ring variant (fp32, VEC = 4) ---------------------------
// Same arithmetic as k_phaseA.  B200 wants on the order of 100 KB of loads in
// flight per SM to saturate HBM3e; every byte a register pipeline keeps in
// flight pins a register.  Here each thread streams the raw operands of the
// next RING_D - 1 planes (r, s, codes; plus the CTA-edge rows / lanes) into its
// own slots of a shared-memory ring with cp.async (LDGSTS) and only touches
// them after cp.async.wait_group, so registers hold just the three combined
// planes.  The slots are thread-private: no extra barrier is needed.
#ifndef SPL_RING_D
#define SPL_RING_D 4      // 4 stages x 3 CTAs/SM measured best (profiles/README.md)
#endif
constexpr int RING_D = SPL_RING_D;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- split-phase CTA barrier (mbarrier) -------------------------------------------
// The marching kernels hand the tile of plane c+1 to the neighbouring warps one plane
// before it is read, so the hand-over does not need a rendezvous: every thread ARRIVES
// after writing its tile rows and, one iteration later, WAITS for that phase.  Warps
// drift up to one plane apart instead of meeting at a __syncthreads every plane
// (barrier stalls were 10-20 % of the issue slots, profiles/).  Hazards: see k_phaseA_ring.
#ifndef SPL_SPLIT_BARRIER
#define SPL_SPLIT_BARRIER 1
#endif
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
}

struct alignas(16) RingStage {
  float4 a[256];          // z or r of the thread's 4 cells
  float4 old[256];        // previous search direction
  float4 ha[64];          // CTA-edge rows (warp 0: row j-1, warp 7: row j+8)
  float4 hold[64];
  unsigned code[256];
  unsigned hcode[64];
  float ea[16], eold[16]; // CTA-edge lanes: x neighbour owned by the adjacent CTA
  unsigned ecode[16];     // (one byte used)
};

struct RingSmem {
  RingStage st[RING_D];
  float4 tile[3][TILE_Y + 2][32];
  double red[32];
  float lut[8];
  double beta;
  unsigned long long bar;   // split-phase barrier of the tile hand-over
  int stop;
};

#ifndef SPL_RING_MINB
#define SPL_RING_MINB 3
#endif
template <int PC>
__global__ void __launch_bounds__(32 * TILE_Y, SPL_RING_MINB)
k_phaseA_ring(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ zr,
              const float* __restrict__ sold, float* __restrict__ snew, float* __restrict__ q,
              Scal* __restrict__ sc, double* __restrict__ partials, int par, int zchunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSmem& sm_ = *reinterpret_cast<RingSmem*>(smem_raw);
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<float>(sm_.lut);
  if (tid == 0) {
    phaseA_scalars(sc, par, &sm_.stop, &sm_.beta);
    mbar_init(&sm_.bar, 32 * TILE_Y);
  }
  __syncthreads();
  if (sm_.stop) return;

  const float beta = (float)sm_.beta;
  const float* lut = sm_.lut;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * 4;
  const int j = blockIdx.y * TILE_Y + ty;
  const bool active = (i0 < g.nx) && (j < g.ny);
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const int nz = g.nz;
  const long long row = g.nx, plane = g.plane;
  // CTA-edge rows; a row outside the box must read as 0 (unmasked Laplacian)
  const int hrow_j = (ty == 0) ? j - 1 : j + 1;
  const bool has_hrow = active && (ty == 0 || ty == TILE_Y - 1) && hrow_j >= 0 && hrow_j < g.ny;
  const long long hoff = (ty == 0) ? -row : row;
  const int hslot = (ty == 0) ? 0 : TILE_Y + 1;
  const int hidx = ((ty == 0) ? 0 : 32) + lane;
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (!active) sm_.tile[b][ty + 1][lane] = z4;
      if (ty == 0 && !has_hrow) sm_.tile[b][0][lane] = z4;
      if (ty == TILE_Y - 1 && !has_hrow) sm_.tile[b][TILE_Y + 1][lane] = z4;
    }
  }
  const bool has_e = active && ((lane == 0 && i0 > 0) || (lane == 31 && i0 + 4 < g.nx));
  const long long eoff = (lane == 0) ? -1 : 4;
  const int eidx = ty * 2 + ((lane == 0) ? 0 : 1);
  const long long idx0 = ((long long)kb * g.ny + j) * row + i0;   // plane kb

  // request every operand of plane p into ring slot `slot`; `off` = element
  // offset of the thread's cells on plane p (running values: no div / mod / 64-bit
  // multiplies in the loop)
  auto request = [&](int p, int slot, long long off) {
    if (active && p <= ke) {
      RingStage& st = sm_.st[slot];
      const bool remote = (p < 0) || (p >= nz);
      if (remote) {
        cp_async16(&st.a[tid], snew + off);          // halo of s' already holds the new direction
      } else {
        cp_async16(&st.a[tid], zr + off);
        cp_async16(&st.old[tid], sold + off);
        cp_async4(&st.code[tid], codes + off);
      }
      if (p < ke) {                                    // centre planes only
        if (has_hrow) {
          cp_async16(&st.ha[hidx], zr + off + hoff);
          cp_async16(&st.hold[hidx], sold + off + hoff);
          cp_async4(&st.hcode[hidx], codes + off + hoff);
        }
        if (has_e) {
          cp_async4(&st.ea[eidx], zr + off + eoff);
          cp_async4(&st.eold[eidx], sold + off + eoff);
          // 4-byte aligned word that contains the neighbour's code byte
          cp_async4(&st.ecode[eidx], codes + ((off + eoff) & ~3LL));
        }
      }
    }
    cp_async_commit();
  };
  auto dirf = [&](float a, float o, unsigned code) {
    const float zz = (PC == PC_JACOBI) ? a * lut[lut_index(code)] : a;
    return zz + beta * o;
  };
  auto dir4 = [&](float4 a, float4 o, unsigned c) {
    float4 r;
    r.x = dirf(a.x, o.x, c & 0xffu);
    r.y = dirf(a.y, o.y, (c >> 8) & 0xffu);
    r.z = dirf(a.z, o.z, (c >> 16) & 0xffu);
    r.w = dirf(a.w, o.w, c >> 24);
    return r;
  };
  // the neighbour's code byte inside the aligned word: lane 0 reads element
  // i0 - 1 (byte 3 of the previous word), lane 31 element i0 + 4 (byte 0)
  const int eshift = (lane == 0) ? 24 : 0;
  auto combine_own = [&](int p, int slot, unsigned& code_out) {
    const RingStage& st = sm_.st[slot];
    const float4 a = st.a[tid];
    code_out = 0u;
    if (p < 0 || p >= nz) return a;
    code_out = st.code[tid];
    return dir4(a, st.old[tid], code_out);
  };
  auto combine_hrow = [&](int slot) {
    const RingStage& st = sm_.st[slot];
    return dir4(st.ha[hidx], st.hold[hidx], st.hcode[hidx]);
  };
  auto combine_edge = [&](int slot) {
    const RingStage& st = sm_.st[slot];
    return dirf(st.ea[eidx], st.eold[eidx], (st.ecode[eidx] >> eshift) & 0xffu);
  };
  auto next_slot = [](int s) { return (s + 1 == RING_D) ? 0 : s + 1; };
  auto next_tile = [](int t) { return (t + 1 == 3) ? 0 : t + 1; };

  // ---- prologue ---------------------------------------------------------------
  float4 smv = make_float4(0.f, 0.f, 0.f, 0.f), scur = smv, sp = smv;
  float leftc = 0.f, rightc = 0.f;
  long long roff = idx0;                                    // offset of the next requested plane
  int rslot = 0;                                            // ring slot of the next requested plane
  for (int p = kb; p < kb + RING_D - 1; ++p) {              // RING_D - 1 groups
    request(p, rslot, roff);
    roff += plane;
    rslot = next_slot(rslot);
  }
  if (active) {                                             // plane kb-1: registers
    const long long idm = idx0 - plane;
    if (kb - 1 < 0) {
      smv = *reinterpret_cast<const float4*>(snew + idm);
    } else {
      smv = dir4(*reinterpret_cast<const float4*>(zr + idm),
                 *reinterpret_cast<const float4*>(sold + idm),
                 *reinterpret_cast<const unsigned*>(codes + idm));
    }
  }
  cp_async_wait<RING_D - 2>();                              // plane kb has landed
  unsigned codec = 0;
  int tcur = 0;                                             // tile buffer of plane c
  if (active) {
    scur = combine_own(kb, 0, codec);
    sm_.tile[0][ty + 1][lane] = scur;
    if (has_hrow) sm_.tile[0][hslot][lane] = combine_hrow(0);
    if (has_e) {
      const float e = combine_edge(0);
      if (lane == 0) leftc = e; else rightc = e;
    }
  }

  // Tile hand-over with the split barrier: phase n = "every thread has written its rows
  // of tile(kb + n)".  Iteration n waits for phase n (so tile(c) is complete, and every
  // thread has finished iteration n-2, whose tile buffer iteration n overwrites), writes
  // tile(c+1), arrives for phase n+1, then computes plane c.  A thread can run at most one
  // tile-write ahead of the slowest one, which the three tile buffers cover.
#if SPL_SPLIT_BARRIER
  mbar_arrive(&sm_.bar);                                    // phase 0: tile(kb)
#endif
  double acc = 0.0;
  long long idx = idx0;
  int nslot = 1;                                            // ring slot of plane c+1
  for (int c = kb; c < ke; ++c, idx += plane) {
    request(c + RING_D - 1, rslot, roff);   // into the slot of plane c-1 (fully consumed)
    roff += plane;
    rslot = next_slot(rslot);
    cp_async_wait<RING_D - 2>();            // plane c+1 has landed
#if SPL_SPLIT_BARRIER
    mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);
#endif
    const int tnext = next_tile(tcur);
    float leftn = 0.f, rightn = 0.f;
    unsigned coden = 0;
    if (active) {
      sp = combine_own(c + 1, nslot, coden);
      if (c + 1 < ke) {
        sm_.tile[tnext][ty + 1][lane] = sp;
        if (has_hrow) sm_.tile[tnext][hslot][lane] = combine_hrow(nslot);
        if (has_e) {
          const float e = combine_edge(nslot);
          if (lane == 0) leftn = e; else rightn = e;
        }
      }
    }
    nslot = next_slot(nslot);
#if SPL_SPLIT_BARRIER
    mbar_arrive(&sm_.bar);
#else
    __syncthreads();   // tile[tcur] (plane c) was completed before the previous barrier
#endif
    const float lsh = __shfl_up_sync(0xffffffffu, scur.w, 1);
    const float rsh = __shfl_down_sync(0xffffffffu, scur.x, 1);
    const float left = (lane != 0) ? lsh : leftc;
    const float right = (lane != 31) ? rsh : rightc;
    if (active) {
      const float4 sym = sm_.tile[tcur][ty][lane];
      const float4 syp = sm_.tile[tcur][ty + 2][lane];
      const float s[4] = {scur.x, scur.y, scur.z, scur.w};
      const float ym[4] = {sym.x, sym.y, sym.z, sym.w};
      const float yp[4] = {syp.x, syp.y, syp.z, syp.w};
      const float zm[4] = {smv.x, smv.y, smv.z, smv.w};
      const float zp[4] = {sp.x, sp.y, sp.z, sp.w};
      float qa[4];
      float part = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xm = (e > 0) ? s[e > 0 ? e - 1 : 0] : left;
        const float xp = (e < 3) ? s[e < 3 ? e + 1 : 0] : right;
        qa[e] = lap_cell<float>((codec >> (8 * e)) & 0xffu, s[e], xm, xp, ym[e], yp[e], zm[e],
                                zp[e]);
        part += s[e] * qa[e];
      }
      acc += (double)part;   // 4 products in fp32, accumulated in fp64
      *reinterpret_cast<float4*>(q + idx) = make_float4(qa[0], qa[1], qa[2], qa[3]);
      *reinterpret_cast<float4*>(snew + idx) = scur;
    }
    smv = scur;
    scur = sp;
    codec = coden;
    leftc = leftn;
    rightc = rightn;
    tcur = tnext;
  }
  cp_async_wait<0>();

  const double bs = block_sum(acc, sm_.red);
  const unsigned nb = num_blocks();
  if (tid == 0) partials[linear_block()] = bs;
  if (last_block(&sc->ticketA, nb)) {
    double t = 0.0;
    for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(partials + b);
    t = block_sum(t, sm_.red);
    if (tid == 0) sc->gA[sc->rank] = t;
  }
}

// ---- boundary planes of the new direction (world > 1) ----------------------
// Computes s' on local planes 0 and nz-1 so they can be sent to the z
// neighbours before phase A reads them through the halo.
template <typename T, int PC>
__global__ void k_dir_boundary(Geo g, const uint8_t* __restrict__ codes,
                               const T* __restrict__ zr, const T* __restrict__ sold,
                               T* __restrict__ snew, Scal* __restrict__ sc, int par) {
  if (sc->done) return;
  const double rho1 = rho_of(sc, par ^ 1);
  const double rho2 = rho_of(sc, par);
  const double rmax = rmax_of(sc, par ^ 1);
  if (!(rmax == rmax) || rmax <= sc->tol * sc->bmax || rho1 == 0.0) return;
  const T beta = T(rho1 / rho2);
  const long long n = g.plane;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 2 * n;
       t += (long long)gridDim.x * blockDim.x) {
    const int which = t >= n;
    if (which && g.nz == 1) continue;
    const long long idx = (which ? (long long)(g.nz - 1) * g.plane : 0) + (t - which * n);
    const T a = zr[idx];
    const T b = sold[idx];
    snew[idx] = (PC == PC_JACOBI) ? a * inv_diag<T>(codes[idx]) + beta * b
                                  : a + beta * b;
  }
}

// The same over peer memory (fp32, nx % 4 = 0): s' of the two boundary planes goes straight
// into the neighbours' halo planes of the s' buffer (peer-mapped stores over NVLink), then the
// iteration stamp into their boxes -- no local copy (phase A forms these planes itself) and no
// NCCL call.  lo_dst / hi_dst: the neighbour's halo plane (NULL at the global ends).
template <int PC>
__global__ void __launch_bounds__(256)
k_send_boundary(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ zr,
                const float* __restrict__ sold, float* __restrict__ lo_dst,
                float* __restrict__ hi_dst, Scal* __restrict__ sc, int par, Fused fz,
                unsigned* __restrict__ ticket) {
  __shared__ int s_stop;
  __shared__ float s_beta;
  if (threadIdx.x == 0) {
    int stop = sc->done;
    float beta = 0.f;
    if (!stop) {
      fused_sync_B(fz, sc, par ^ 1, blockIdx.x == 0);
      const double rho1 = rho_of(sc, par ^ 1);
      const double rho2 = rho_of(sc, par);
      const double rmax = rmax_of(sc, par ^ 1);
      if (!(rmax == rmax) || rmax <= sc->tol * sc->bmax || rho1 == 0.0) stop = 1;   // phase A stops too
      else beta = (float)(rho1 / rho2);
    }
    s_stop = stop;
    s_beta = beta;
  }
  __syncthreads();
  if (s_stop) return;
  const float beta = s_beta;
  const long long n4 = g.plane / 4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 2 * n4;
       t += (long long)gridDim.x * blockDim.x) {
    const int which = t >= n4;
    float* __restrict__ dst = which ? hi_dst : lo_dst;
    if (!dst) continue;
    const long long o = (t - which * n4) * 4;
    const long long idx = (which ? (long long)(g.nz - 1) * g.plane : 0) + o;
    const float4 a = *reinterpret_cast<const float4*>(zr + idx);
    const float4 b = *reinterpret_cast<const float4*>(sold + idx);
    const unsigned cd = *reinterpret_cast<const unsigned*>(codes + idx);
    float4 r;
    // the rounding of phase A's dirf: zz = a / diag (one multiply), then fma(beta, s, zz)
    if (PC == PC_JACOBI) {
      r.x = __fmaf_rn(beta, b.x, __fmul_rn(a.x, inv_diag<float>(cd & 0xffu)));
      r.y = __fmaf_rn(beta, b.y, __fmul_rn(a.y, inv_diag<float>((cd >> 8) & 0xffu)));
      r.z = __fmaf_rn(beta, b.z, __fmul_rn(a.z, inv_diag<float>((cd >> 16) & 0xffu)));
      r.w = __fmaf_rn(beta, b.w, __fmul_rn(a.w, inv_diag<float>(cd >> 24)));
    } else {
      r.x = __fmaf_rn(beta, b.x, a.x); r.y = __fmaf_rn(beta, b.y, a.y);
      r.z = __fmaf_rn(beta, b.z, a.z); r.w = __fmaf_rn(beta, b.w, a.w);
    }
    *reinterpret_cast<float4*>(dst + o) = r;
  }
  // ONE system-scope fence per CTA (a membar.sys in every thread stalls the memory pipeline of
  // the SMs this kernel shares with phase A for tens of microseconds): the barrier orders the
  // CTA's peer stores before thread 0, whose release is cumulative over them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1u) {
      *ticket = 0u;
      __threadfence_system();
      const int rank = sc->rank;
      if (lo_dst) st_release_sys(&fz.box[rank - 1]->hseq[1], fz.stamp);
      if (hi_dst) st_release_sys(&fz.box[rank + 1]->hseq[0], fz.stamp);
    }
  }
}

// slots of the last phase B of a solve, for the kernels that follow the loop
__global__ void k_fused_wait_B(Scal* sc, int slot, Fused fz) {
  if (sc->done) return;
  if (threadIdx.x == 0) fused_fill_B(fz, sc, slot);
}

// ---- phase B ---------------------------------------------------------------
// init != 0: alpha forced to 0 (used once per solve to produce rho_0, max|b|).
template <typename T, int VEC, int PC>
__global__ void __launch_bounds__(256)
k_phaseB(Geo g, const uint8_t* __restrict__ codes, T* __restrict__ r,
         const T* __restrict__ q, const T* __restrict__ z, Scal* __restrict__ sc,
         double* __restrict__ partials, int par, int init, int nsbuf, Fused fz) {
  __shared__ double red[32];
  __shared__ double s_alpha;
  __shared__ T lut[8];
  const int tid = threadIdx.x;
  if (sc->done) return;
  fill_inv_lut<T>(lut);
  if (tid == 0) {
    double alpha = 0.0;
    if (!init) {
      fused_sync_A(fz, sc, blockIdx.x == 0);   // z-slabs over peer memory: s.q of the other ranks
      const double rho1 = rho_of(sc, par ^ 1);
      const double sq = sum_slots(sc->gA, sc->world);
      alpha = rho1 / sq;
      // remembered for the deferred x update (iteration number = count + 1;
      // count only moves after every CTA of this launch has passed this point)
      if (blockIdx.x == 0) sc->alpha_hist[(sc->count + 1) % nsbuf] = alpha;
    }
    s_alpha = alpha;
  }
  __syncthreads();
  const T alpha = T(s_alpha);
  const long long nq = g.ncell / VEC;
  double rz = 0.0, mx = 0.0;
  int bad = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + tid; t < nq;
       t += (long long)gridDim.x * blockDim.x) {
    const long long idx = t * VEC;
    VecT<T, VEC> rv = vload<T, VEC>(r, idx);
    if (!init) {
      const VecT<T, VEC> qv = vload<T, VEC>(q, idx);
#pragma unroll
      for (int e = 0; e < VEC; ++e) rv.v[e] -= alpha * qv.v[e];
      vstore<T, VEC>(r, idx, rv);
    }
    VecT<uint8_t, VEC> c;
    if (PC == PC_JACOBI) c = cload<VEC>(codes, idx);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const double rr = (double)rv.v[e];
      if (PC == PC_JACOBI) rz += rr * (double)(rv.v[e] * lut[lut_index(c.v[e])]);
      if (PC == PC_NONE) rz += rr * rr;
      // PC_RBIC0: r.z is produced by the preconditioner sweep
      const double ar = fabs(rr);
      bad |= !(ar <= 1.7976931348623157e308);
      mx = fmax(mx, ar);
    }
  }
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  const double bs = block_sum(rz, red);
  const int anybad = __syncthreads_or(bad);
  const double bm = block_max(mx, red);
  const unsigned nb = gridDim.x;
  if (tid == 0) {
    partials[blockIdx.x] = bs;
    partials[MAX_PARTIALS + blockIdx.x] = anybad ? kNaN : bm;
  }
  if (last_block(&sc->ticketB, nb)) {
    double t = 0.0, m = 0.0;
    int nanf = 0;
    for (unsigned b = tid; b < nb; b += blockDim.x) {
      t += __ldcg(partials + b);
      const double v = __ldcg(partials + MAX_PARTIALS + b);
      if (v != v) nanf = 1; else m = fmax(m, v);
    }
    t = block_sum(t, red);
    nanf = __syncthreads_or(nanf);
    m = block_max(m, red);
    if (tid == 0) {
      const int rank = sc->rank;
      if (!pc_explicit_z(PC)) sc->gBM[par][rank][0] = t;
      sc->gBM[par][rank][1] = nanf ? kNaN : m;
      if (!init) sc->count = sc->count + 1;
      if (!init && !pc_explicit_z(PC)) fused_push_B(fz, sc, t, nanf ? kNaN : m);
    }
  }
}

// ---- red-black incomplete Cholesky preconditioner (SPL_PC_RBIC0) ------------------
// Substitution for the reference's exact solve (Pressure.fs:69): natural-order
// IC(0)/MIC(0) is sequential, so the factorisation is taken in red-black order
// (red = (i+j+k) even first): M = (E + L) E^-1 (E + L^T) with
//   E_red   = N_i,
//   E_black = N_i - sum_{red Fluid nbr j} (1 + tau (nfl_j - 1)) / N_j   (safety 0.25 N_i)
// (oracle/dense_ref.py rbic0_diag / apply_rbic0).  z = M^-1 r is two half sweeps:
//   black: z_B = (r_B + sum_{red nbr} r_n einv_n) einv_B
//   red  : z_R = (r_R + sum_{black nbr} z_n) einv_R            (+ r.z partials)
// Couplings across z-slabs are dropped (block-Jacobi over slabs), so
// no extra halo exchange is needed; iteration counts may rise with the slab count.
__device__ __forceinline__ int cell_parity(const Geo& g, long long idx, int& k_out) {
  const int i = (int)(idx % g.nx);
  const long long t = idx / g.nx;
  const int j = (int)(t % g.ny);
  const int k = (int)(t / g.ny);
  k_out = k;
  return (i + j + k + g.z0) & 1;
}

// step 1: contrib_j = (1 + tau (nfl_j - 1)) / N_j on red Fluid cells, 0 elsewhere
template <typename T>
__global__ void k_rb_contrib(Geo g, const uint8_t* __restrict__ codes,
                             const uint8_t* __restrict__ media, T* __restrict__ contrib,
                             double tau) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    int k;
    const int par = cell_parity(g, idx, k);
    const unsigned c = codes[idx];
    T out = T(0);
    const int N = __popc(c & 63u);
    if ((c & B_FLUID) && par == 0 && N > 0) {
      int nfl = 0;   // Fluid neighbours inside this slab
      if ((c & B_XM) && media[idx - 1] == FLUID) ++nfl;
      if ((c & B_XP) && media[idx + 1] == FLUID) ++nfl;
      if ((c & B_YM) && media[idx - g.nx] == FLUID) ++nfl;
      if ((c & B_YP) && media[idx + g.nx] == FLUID) ++nfl;
      if ((c & B_ZM) && k > 0 && media[idx - g.plane] == FLUID) ++nfl;
      if ((c & B_ZP) && k < g.nz - 1 && media[idx + g.plane] == FLUID) ++nfl;
      out = (T)((1.0 + tau * (double)(nfl - 1)) / (double)N);
    }
    contrib[idx] = out;
  }
}

// step 2: einv = 1 / E on Fluid cells, 0 elsewhere
template <typename T>
__global__ void k_rb_einv(Geo g, const uint8_t* __restrict__ codes,
                          const T* __restrict__ contrib, T* __restrict__ einv, double safety) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    int k;
    const int par = cell_parity(g, idx, k);
    const unsigned c = codes[idx];
    const int N = __popc(c & 63u);
    T out = T(0);
    if ((c & B_FLUID) && N > 0) {
      double E = (double)N;
      if (par == 1) {
        double acc = 0.0;
        if (c & B_XM) acc += (double)contrib[idx - 1];
        if (c & B_XP) acc += (double)contrib[idx + 1];
        if (c & B_YM) acc += (double)contrib[idx - g.nx];
        if (c & B_YP) acc += (double)contrib[idx + g.nx];
        if ((c & B_ZM) && k > 0) acc += (double)contrib[idx - g.plane];
        if ((c & B_ZP) && k < g.nz - 1) acc += (double)contrib[idx + g.plane];
        const double eb = (double)N - acc;
        E = (eb < safety * (double)N) ? (double)N : eb;
      }
      out = (T)(1.0 / E);
    }
    einv[idx] = out;
  }
}

// one half sweep; COLOR = 1 black (reads r of red neighbours), 0 red (reads z of
// black neighbours and reduces r.z over every cell)
template <typename T, int COLOR>
__global__ void __launch_bounds__(256)
k_rb_sweep(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ r,
           const T* __restrict__ einv, T* __restrict__ z, Scal* __restrict__ sc,
           double* __restrict__ partials, int par_slot) {
  __shared__ double red[32];
  if (sc->done) return;
  double acc = 0.0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    int k;
    const int par = cell_parity(g, idx, k);
    const unsigned c = codes[idx];
    if (par == COLOR) {
      T zz = T(0);
      if (c & B_FLUID) {
        const bool zm = (c & B_ZM) && k > 0, zp = (c & B_ZP) && k < g.nz - 1;
        T sum = T(0);
        if (COLOR == 1) {
          if (c & B_XM) sum += r[idx - 1] * einv[idx - 1];
          if (c & B_XP) sum += r[idx + 1] * einv[idx + 1];
          if (c & B_YM) sum += r[idx - g.nx] * einv[idx - g.nx];
          if (c & B_YP) sum += r[idx + g.nx] * einv[idx + g.nx];
          if (zm) sum += r[idx - g.plane] * einv[idx - g.plane];
          if (zp) sum += r[idx + g.plane] * einv[idx + g.plane];
        } else {
          if (c & B_XM) sum += z[idx - 1];
          if (c & B_XP) sum += z[idx + 1];
          if (c & B_YM) sum += z[idx - g.nx];
          if (c & B_YP) sum += z[idx + g.nx];
          if (zm) sum += z[idx - g.plane];
          if (zp) sum += z[idx + g.plane];
        }
        zz = (r[idx] + sum) * einv[idx];
      }
      z[idx] = zz;
      if (COLOR == 0) acc += (double)r[idx] * (double)zz;
    } else if (COLOR == 0) {
      acc += (double)r[idx] * (double)z[idx];   // black cells, written by the black sweep
    }
  }
  if (COLOR == 0) {
    const double bs = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    if (last_block(&sc->ticketC, gridDim.x)) {
      double t = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(partials + b);
      t = block_sum(t, red);
      if (threadIdx.x == 0) sc->gBM[par_slot][sc->rank][0] = t;
    }
  }
}

// ---- the two half sweeps, four x cells per thread (fp32, nx % 4 = 0) -----------------
// Same expressions, neighbour order and guards as k_rb_sweep, cell by cell, so z is bit
// for bit the one-cell kernel's; what changes is the data movement: a thread owns an
// aligned quad (two red and two black cells, colours alternate from the parity of
// j + k + z0), every array is read with 128-bit loads -- own row, the +-y rows and the
// +-z planes -- plus one scalar either side for the x neighbours outside the quad, and z is
// written as whole quads.  BLACK pass: z_B on black cells, 0 on red ones (13 B / cell);
// RED pass: z_R from the black z around it, black values carried through, r.z reduced
// over all four cells (17 B / cell).  No per-cell div / mod, no idle half-warps.
struct RbQuad {
  long long idx;
  int k;
  unsigned p0;            // colour of cell 0 of the quad (1 = black)
  unsigned c[4];          // stencil codes
};
__device__ __forceinline__ RbQuad rb_quad(const Geo& g, long long q, int qx,
                                          const uint8_t* __restrict__ codes) {
  RbQuad r;
  r.idx = q << 2;
  const long long t = q / qx;
  const int j = (int)(t % g.ny);
  r.k = (int)(t / g.ny);
  r.p0 = (unsigned)(j + r.k + g.z0) & 1u;
  const unsigned c4 = *reinterpret_cast<const unsigned*>(codes + r.idx);
  r.c[0] = c4 & 0xffu; r.c[1] = (c4 >> 8) & 0xffu; r.c[2] = (c4 >> 16) & 0xffu; r.c[3] = c4 >> 24;
  return r;
}
__device__ __forceinline__ float4 ld4(const float* __restrict__ p, long long idx) {
  return *reinterpret_cast<const float4*>(p + idx);
}

__global__ void __launch_bounds__(256)
k_rb_black4(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ r,
            const float* __restrict__ einv, float* __restrict__ z, Scal* __restrict__ sc) {
  if (sc->done) return;
  const long long nq = g.ncell >> 2;
  const int qx = g.nx >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const RbQuad Q = rb_quad(g, q, qx, codes);
    const long long idx = Q.idx;
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned any = (Q.c[0] | Q.c[1] | Q.c[2] | Q.c[3]) & B_FLUID;
    if (any) {
      const float4 r4 = ld4(r, idx), e4 = ld4(einv, idx);
      const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
      // r einv of the red neighbours; a row / plane is fetched only if some cell of the
      // quad has that neighbour (the guards of k_rb_sweep)
      const unsigned all = Q.c[0] | Q.c[1] | Q.c[2] | Q.c[3];
      const bool zm = Q.k > 0, zp = Q.k < g.nz - 1;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 rym = z4, eym = z4, ryp = z4, eyp = z4, rzm = z4, ezm = z4, rzp = z4, ezp = z4;
      if (all & B_YM) { rym = ld4(r, idx - g.nx); eym = ld4(einv, idx - g.nx); }
      if (all & B_YP) { ryp = ld4(r, idx + g.nx); eyp = ld4(einv, idx + g.nx); }
      if ((all & B_ZM) && zm) { rzm = ld4(r, idx - g.plane); ezm = ld4(einv, idx - g.plane); }
      if ((all & B_ZP) && zp) { rzp = ld4(r, idx + g.plane); ezp = ld4(einv, idx + g.plane); }
      const float* a_rym = reinterpret_cast<const float*>(&rym);
      const float* a_eym = reinterpret_cast<const float*>(&eym);
      const float* a_ryp = reinterpret_cast<const float*>(&ryp);
      const float* a_eyp = reinterpret_cast<const float*>(&eyp);
      const float* a_rzm = reinterpret_cast<const float*>(&rzm);
      const float* a_ezm = reinterpret_cast<const float*>(&ezm);
      const float* a_rzp = reinterpret_cast<const float*>(&rzp);
      const float* a_ezp = reinterpret_cast<const float*>(&ezp);
      float rx[6], ex[6];                           // r, einv at i-1 .. i+4
      rx[0] = ex[0] = rx[5] = ex[5] = 0.f;
      if (Q.p0 == 1u && (Q.c[0] & B_XM)) { rx[0] = r[idx - 1]; ex[0] = einv[idx - 1]; }
      if (Q.p0 == 0u && (Q.c[3] & B_XP)) { rx[5] = r[idx + 4]; ex[5] = einv[idx + 4]; }
#pragma unroll
      for (int e = 0; e < 4; ++e) { rx[e + 1] = rr[e]; ex[e + 1] = ee[e]; }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned c = Q.c[e];
        if ((((unsigned)e ^ Q.p0) & 1u) && (c & B_FLUID)) {   // black Fluid cell
          // one fused multiply-add per neighbour: what `sum += r * einv` compiles to in
          // k_rb_sweep
          float sum = 0.f;
          if (c & B_XM) sum = fmaf(rx[e], ex[e], sum);
          if (c & B_XP) sum = fmaf(rx[e + 2], ex[e + 2], sum);
          if (c & B_YM) sum = fmaf(a_rym[e], a_eym[e], sum);
          if (c & B_YP) sum = fmaf(a_ryp[e], a_eyp[e], sum);
          if ((c & B_ZM) && zm) sum = fmaf(a_rzm[e], a_ezm[e], sum);
          if ((c & B_ZP) && zp) sum = fmaf(a_rzp[e], a_ezp[e], sum);
          zz[e] = (rr[e] + sum) * ee[e];
        }
      }
    }
    *reinterpret_cast<float4*>(z + idx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
  }
}

__global__ void __launch_bounds__(256)
k_rb_red4(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ r,
          const float* __restrict__ einv, float* __restrict__ z, Scal* __restrict__ sc,
          double* __restrict__ partials, int par_slot) {
  __shared__ double red[32];
  if (sc->done) return;
  double acc = 0.0;
  const long long nq = g.ncell >> 2;
  const int qx = g.nx >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const RbQuad Q = rb_quad(g, q, qx, codes);
    const long long idx = Q.idx;
    const unsigned any = (Q.c[0] | Q.c[1] | Q.c[2] | Q.c[3]) & B_FLUID;
    if (!any) continue;            // z = 0 there already (black pass), r.z terms are 0
    const float4 r4 = ld4(r, idx), e4 = ld4(einv, idx), z4 = ld4(z, idx);
    const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
    float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    const unsigned all = Q.c[0] | Q.c[1] | Q.c[2] | Q.c[3];
    const bool zm = Q.k > 0, zp = Q.k < g.nz - 1;
    float4 aym = make_float4(0.f, 0.f, 0.f, 0.f), ayp = aym, azm = aym, azp = aym;
    if (all & B_YM) aym = ld4(z, idx - g.nx);
    if (all & B_YP) ayp = ld4(z, idx + g.nx);
    if ((all & B_ZM) && zm) azm = ld4(z, idx - g.plane);
    if ((all & B_ZP) && zp) azp = ld4(z, idx + g.plane);
    const float ym[4] = {aym.x, aym.y, aym.z, aym.w}, yp[4] = {ayp.x, ayp.y, ayp.z, ayp.w};
    const float zmv[4] = {azm.x, azm.y, azm.z, azm.w}, zpv[4] = {azp.x, azp.y, azp.z, azp.w};
    float zx[6];                                    // black z at i-1 .. i+4
    zx[0] = (Q.p0 == 0u && (Q.c[0] & B_XM)) ? z[idx - 1] : 0.f;
    zx[5] = (Q.p0 == 1u && (Q.c[3] & B_XP)) ? z[idx + 4] : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) zx[e + 1] = zz[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned c = Q.c[e];
      if (!(((unsigned)e ^ Q.p0) & 1u)) {           // red cell
        float v = 0.f;
        if (c & B_FLUID) {
          float sum = 0.f;
          if (c & B_XM) sum += zx[e];
          if (c & B_XP) sum += zx[e + 2];
          if (c & B_YM) sum += ym[e];
          if (c & B_YP) sum += yp[e];
          if ((c & B_ZM) && zm) sum += zmv[e];
          if ((c & B_ZP) && zp) sum += zpv[e];
          v = (rr[e] + sum) * ee[e];
        }
        zz[e] = v;
      }
      acc += (double)rr[e] * (double)zz[e];
    }
    *reinterpret_cast<float4*>(z + idx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
  }
  const double bs = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = bs;
  if (last_block(&sc->ticketC, gridDim.x)) {
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(partials + b);
    t = block_sum(t, red);
    if (threadIdx.x == 0) sc->gBM[par_slot][sc->rank][0] = t;
  }
}

// ---- the two half sweeps as z-marching tile kernels (fp32, nx % 4 = 0) --------------------
// k_rb_black4 / k_rb_red4 are grid-stride kernels: the quads of the +-z planes, 2 MiB away, are
// partly read from DRAM a second time (ncu: 1.53 x / 1.18 x the algorithmic bytes).  Here a CTA
// owns a 128 x 8 tile of the x-y plane and marches along z: the operand of the plane below and
// the plane itself are carried in registers, the plane above is this iteration's load; the +-y
// rows and the x neighbours outside the quad come from a double-buffered shared tile written one
// plane ahead (one barrier per plane, never waiting for a load of its own iteration); only the
// tile-edge rows / lanes read a halo from global memory.  RED = 0: z_B from r einv of the red
// neighbours (operands r, einv); RED = 1: z_R from the black z around it, r.z reduced.  Same
// expressions and neighbour order as k_rb_sweep -> the same z bit for bit.
constexpr int RBM_TY = 8;

template <int RED>
__global__ void __launch_bounds__(32 * RBM_TY, 4)
k_rb_march(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ r,
           const float* __restrict__ einv, float* __restrict__ z, Scal* __restrict__ sc,
           double* __restrict__ partials, int par_slot, int zchunk) {
  __shared__ float sa[2][RBM_TY + 2][128];               // r (black pass) / z (red pass)
  __shared__ float se[RED ? 1 : 2][RED ? 1 : RBM_TY + 2][RED ? 4 : 128];   // einv (black pass)
  __shared__ double red[32];
  if (sc->done) return;
  const float* __restrict__ A = RED ? z : r;             // the array whose neighbours are summed
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 128 + 4 * lane, j = blockIdx.y * RBM_TY + ty;
  const int kb = blockIdx.z * zchunk, ke = min(kb + zchunk, g.nz);
  const bool act = i < g.nx && j < g.ny;
  const bool has_xl = act && i > 0, has_xr = act && i + 4 < g.nx;
  const bool row_lo = (ty == 0), row_hi = (ty == RBM_TY - 1);
  const bool edge = row_lo || row_hi;
  const bool edge_in = row_lo ? (i < g.nx && j >= 1 && j - 1 < g.ny) : (i < g.nx && j + 1 < g.ny);
  const long long eoff = row_lo ? -(long long)g.nx : (long long)g.nx;
  const int erow = row_lo ? 0 : RBM_TY + 1;
  const float4 Z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  long long idx = ((long long)kb * g.ny + j) * g.nx + i;
  double acc = 0.0;

  float4 am = Z4, ac = Z4, em = Z4, ec = Z4;             // planes k-1 and k of A (and einv)
  if (act) {
    ac = ld4(A, idx);
    if (!RED) ec = ld4(einv, idx);
    if (kb > 0) {
      am = ld4(A, idx - g.plane);
      if (!RED) em = ld4(einv, idx - g.plane);
    }
  }
  *reinterpret_cast<float4*>(&sa[0][ty + 1][4 * lane]) = ac;
  if (!RED) *reinterpret_cast<float4*>(&se[0][ty + 1][4 * lane]) = ec;
  if (edge) {
    *reinterpret_cast<float4*>(&sa[0][erow][4 * lane]) = edge_in ? ld4(A, idx + eoff) : Z4;
    if (!RED) *reinterpret_cast<float4*>(&se[0][erow][4 * lane]) = edge_in ? ld4(einv, idx + eoff) : Z4;
  }

  for (int k = kb; k < ke; ++k, idx += g.plane) {
    const int b = (k - kb) & 1;
    const bool zm = k > 0, zp = k < g.nz - 1;            // couplings across slabs are dropped
    // ---- this plane's loads (none is needed before the barrier) -----------------------------
    unsigned c4 = 0;
    float4 ap = Z4, ep = Z4, r4 = Z4, e4 = Z4, ha = Z4, he = Z4;
    float axl = 0.f, axr = 0.f, exl = 0.f, exr = 0.f;
    if (act) {
      c4 = *reinterpret_cast<const unsigned*>(codes + idx);
      if (zp) {
        ap = ld4(A, idx + g.plane);
        if (!RED) ep = ld4(einv, idx + g.plane);
      }
      if (RED) { r4 = ld4(r, idx); e4 = ld4(einv, idx); }
      if (lane == 0 && has_xl) { axl = A[idx - 1]; if (!RED) exl = einv[idx - 1]; }
      if (lane == 31 && has_xr) { axr = A[idx + 4]; if (!RED) exr = einv[idx + 4]; }
    }
    if (edge && edge_in && k + 1 < ke) {                 // halo row of plane k+1
      ha = ld4(A, idx + g.plane + eoff);
      if (!RED) he = ld4(einv, idx + g.plane + eoff);
    }
    __syncthreads();
    const float4 aym = *reinterpret_cast<const float4*>(&sa[b][ty][4 * lane]);
    const float4 ayp = *reinterpret_cast<const float4*>(&sa[b][ty + 2][4 * lane]);
    float4 eym = Z4, eyp = Z4;
    if (!RED) {
      eym = *reinterpret_cast<const float4*>(&se[b][ty][4 * lane]);
      eyp = *reinterpret_cast<const float4*>(&se[b][ty + 2][4 * lane]);
    }
    if (lane > 0) { axl = sa[b][ty + 1][4 * lane - 1]; if (!RED) exl = se[b][ty + 1][4 * lane - 1]; }
    if (lane < 31) { axr = sa[b][ty + 1][4 * lane + 4]; if (!RED) exr = se[b][ty + 1][4 * lane + 4]; }
    const float ax[6] = {axl, ac.x, ac.y, ac.z, ac.w, axr};      // i-1 .. i+4
    const float ex[6] = {exl, ec.x, ec.y, ec.z, ec.w, exr};
    const float ym_[4] = {aym.x, aym.y, aym.z, aym.w}, yp_[4] = {ayp.x, ayp.y, ayp.z, ayp.w};
    const float eym_[4] = {eym.x, eym.y, eym.z, eym.w}, eyp_[4] = {eyp.x, eyp.y, eyp.z, eyp.w};
    const float zm_[4] = {am.x, am.y, am.z, am.w}, zp_[4] = {ap.x, ap.y, ap.z, ap.w};
    const float ezm_[4] = {em.x, em.y, em.z, em.w}, ezp_[4] = {ep.x, ep.y, ep.z, ep.w};
    const float rr[4] = {RED ? r4.x : ac.x, RED ? r4.y : ac.y, RED ? r4.z : ac.z, RED ? r4.w : ac.w};
    const float ee[4] = {RED ? e4.x : ec.x, RED ? e4.y : ec.y, RED ? e4.z : ec.z, RED ? e4.w : ec.w};
    const unsigned p0 = (unsigned)(j + k + g.z0) & 1u;           // colour of cell 0 (1 = black)
    float zz[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned c = (c4 >> (8 * e)) & 0xffu;
      const bool black = (((unsigned)e ^ p0) & 1u) != 0u;
      float v = RED ? ax[e + 1] : 0.f;                   // red pass: black z carried through
      if (black != (RED != 0)) {                         // the colour this pass computes
        v = 0.f;
        if (c & B_FLUID) {
          float sum = 0.f;
          if (RED) {
            if (c & B_XM) sum += ax[e];
            if (c & B_XP) sum += ax[e + 2];
            if (c & B_YM) sum += ym_[e];
            if (c & B_YP) sum += yp_[e];
            if ((c & B_ZM) && zm) sum += zm_[e];
            if ((c & B_ZP) && zp) sum += zp_[e];
          } else {
            if (c & B_XM) sum = fmaf(ax[e], ex[e], sum);
            if (c & B_XP) sum = fmaf(ax[e + 2], ex[e + 2], sum);
            if (c & B_YM) sum = fmaf(ym_[e], eym_[e], sum);
            if (c & B_YP) sum = fmaf(yp_[e], eyp_[e], sum);
            if ((c & B_ZM) && zm) sum = fmaf(zm_[e], ezm_[e], sum);
            if ((c & B_ZP) && zp) sum = fmaf(zp_[e], ezp_[e], sum);
          }
          v = (rr[e] + sum) * ee[e];
        }
      }
      zz[e] = v;
      if (RED) acc += (double)rr[e] * (double)v;
    }
    if (act) *reinterpret_cast<float4*>(z + idx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
    am = ac; ac = ap;
    if (!RED) { em = ec; ec = ep; }
    // the next plane's tile rows (read after the next barrier)
    *reinterpret_cast<float4*>(&sa[b ^ 1][ty + 1][4 * lane]) = ap;
    if (!RED) *reinterpret_cast<float4*>(&se[b ^ 1][ty + 1][4 * lane]) = ep;
    if (edge) {
      *reinterpret_cast<float4*>(&sa[b ^ 1][erow][4 * lane]) = ha;
      if (!RED) *reinterpret_cast<float4*>(&se[b ^ 1][erow][4 * lane]) = he;
    }
  }
  if (RED) {
    const double bs = block_sum(acc, red);
    const unsigned nb = num_blocks();
    const unsigned tid = lane + 32 * ty;
    if (tid == 0) partials[linear_block()] = bs;
    if (last_block(&sc->ticketC, nb)) {
      double t = 0.0;
      for (unsigned q = tid; q < nb; q += 32 * RBM_TY) t += __ldcg(partials + q);
      t = block_sum(t, red);
      if (tid == 0) sc->gBM[par_slot][sc->rank][0] = t;
    }
  }
}

// ---- deferred pressure update ----------------------------------------------------
// Iteration it leaves its search direction in ring buffer it % m and its step
// length in alpha_hist[it % m]; this kernel adds every pending term
//   x += sum_{it = flushed+1 .. count} alpha_it s_it        (fp64 accumulator)
// and is launched after every m-th iteration and once after the loop, so a
// buffer is always flushed before phase A overwrites it.
struct SBufs {
  const void* p[MAX_SBUF];
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
k_xflush(Geo g, double* __restrict__ x, SBufs bufs, Scal* __restrict__ sc, int m) {
  __shared__ double s_alpha[MAX_SBUF];
  __shared__ const T* s_ptr[MAX_SBUF];
  __shared__ int s_n;
  const int lo = sc->flushed, hi = sc->count;
  if (threadIdx.x == 0) {
    int n = hi - lo;
    if (n > m) n = m;
    if (n < 0) n = 0;
    for (int t = 0; t < n; ++t) {
      const int it = lo + 1 + t;
      const double a = sc->alpha_hist[it % m];
      s_alpha[t] = (fabs(a) <= 1.7976931348623157e308) ? a : 0.0;
      s_ptr[t] = static_cast<const T*>(bufs.p[it % m]);
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  if (n > 0) {
    const long long nq = g.ncell / VEC;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nq;
         t += (long long)gridDim.x * blockDim.x) {
      const long long idx = t * VEC;
      VecT<double, VEC> xv = vload<double, VEC>(x, idx);
      for (int u = 0; u < n; ++u) {
        const VecT<T, VEC> sv = vload<T, VEC>(s_ptr[u], idx);
        const double a = s_alpha[u];
#pragma unroll
        for (int e = 0; e < VEC; ++e) xv.v[e] += a * (double)sv.v[e];
      }
      vstore<double, VEC>(x, idx, xv);
    }
  }
  if (last_block(&sc->ticketX, gridDim.x)) {
    if (threadIdx.x == 0) sc->flushed = hi;
  }
}

}  // namespace spl
