// spl_pcg.cuh -- matrix-free PCG on the fluid-cell Laplacian (sm_100a)
//
// Replaces Pressure.calculate's assembly + direct solve (Pressure.fs:42-69):
// A is never formed; one PCG iteration is two kernels separated by
// the two global reductions CG needs:
//
//   phase A(it): s' = z + beta s ; q = A s' ; partial s'.q ;
//                x += alpha_{it-1} s            (the previous direction, which
//                                                this kernel reads anyway)
//   phase B(it): r -= alpha q ; partial r.z, max|r|
//
// The pressure accumulator x is always fp64: with fp32 storage its rounding
// (eps |p|) is what limits the divergence-free residual on 256^3+ grids.
// fp32 fields: A = r4 + s4 + code1 + x8 -> s'4 + q4 + x8 = 33 B/cell,
//              B = r4 + q4 + code1 -> r4 = 13 B/cell.
//
// HBM-bound; no tensor cores (no dense contraction on this path).
#pragma once
#include "spl_common.cuh"

namespace spl {

enum { PC_NONE = 0, PC_JACOBI = 1, PC_RBIC0 = 2 };

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
// memory.  Software pipeline: the raw operands (r, s, codes, x) of plane c+2
// are requested at the top of iteration c and first touched at the top of
// iteration c+1, so a whole plane of loads per thread is always in flight.
// Every array is pulled from HBM once.
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

#ifndef SPL_PHASEA_MINB
#define SPL_PHASEA_MINB 3     // resident CTAs per SM for fp32 (fp64: one less)
#endif
template <typename T, int VEC, int PC>
__global__ void __launch_bounds__(32 * TILE_Y, (sizeof(T) == 8) ? SPL_PHASEA_MINB - 1 : SPL_PHASEA_MINB)
k_phaseA(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ zr,
         const T* __restrict__ sold, T* __restrict__ snew, T* __restrict__ q,
         double* __restrict__ x, Scal* __restrict__ sc, double* __restrict__ partials,
         int par, int zchunk) {
  __shared__ double red[32];
  __shared__ double s_beta, s_alpha;
  __shared__ int s_stop;
  __shared__ T lut[8];
  __shared__ VecT<T, VEC> tile[3][TILE_Y + 2][32];
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<T>(lut);
  if (tid == 0) {
    int stop = sc->done;
    double beta = 0.0, alpha_prev = 0.0;
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
        // step length of the previous iteration (what phase B(it-1) used)
        if (sc->count > 0) alpha_prev = rho2 / sum_slots(sc->gA, sc->world);
      }
    }
    s_stop = stop;
    s_beta = beta;
    s_alpha = alpha_prev;
  }
  __syncthreads();
  if (s_stop) return;

  const DirLoader<T, VEC, PC> dl{codes, zr, sold, snew, lut, g.nz, T(s_beta)};
  const double alpha_prev = s_alpha;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * VEC;
  const int j = blockIdx.y * TILE_Y + ty;
  const bool active = (i0 < g.nx) && (j < g.ny);
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const long long row = g.nx, plane = g.plane;
  // CTA-edge rows recompute the neighbouring row, CTA-edge lanes the x
  // neighbour owned by the adjacent CTA (memory is always valid thanks to the z
  // halo planes; values outside the box are masked by the code bits)
  const bool has_hrow = active && (ty == 0 || ty == TILE_Y - 1);
  const long long hoff = (ty == 0) ? -row : row;
  const int hslot = (ty == 0) ? 0 : TILE_Y + 1;
  const bool has_e = active && ((lane == 0 && i0 > 0) || (lane == 31 && i0 + VEC < g.nx));
  const long long eoff = (lane == 0) ? -1 : VEC;

  VecT<T, VEC> zero;
  VecT<uint8_t, VEC> czero;
  VecT<double, VEC> dzero;
#pragma unroll
  for (int e = 0; e < VEC; ++e) { zero.v[e] = T(0); czero.v[e] = 0; dzero.v[e] = 0.0; }

  double acc = 0.0;
  long long idx = ((long long)kb * g.ny + j) * row + i0;
  VecT<T, VEC> sm = zero, scur = zero, sp = zero, oldc = zero;
  VecT<uint8_t, VEC> codec = czero;
  VecT<double, VEC> xc = dzero, xn = dzero;
  RawV<T, VEC> raw{zero, zero, czero}, hraw{zero, zero, czero};
  RawS<T> eraw{T(0), T(0), 0u};
  T leftc = T(0), rightc = T(0);

  // rows outside the box must read as 0 (the Laplacian is unmasked)
  const int hrow_j = (ty == 0) ? j - 1 : j + 1;
  const bool hrow_ok = has_hrow && hrow_j >= 0 && hrow_j < g.ny;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    if (!active) tile[b][ty + 1][lane] = zero;                       // own row
    if (ty == 0 && !hrow_ok) tile[b][0][lane] = zero;                // row j-1
    if (ty == TILE_Y - 1 && !hrow_ok) tile[b][TILE_Y + 1][lane] = zero;   // row j+1
  }
  // ---- prologue: planes kb-1 and kb combined, plane kb+1 requested ----------
  if (active) {
    sm = dl.combine(dl.load(idx - plane, kb - 1), kb - 1);
    const RawV<T, VEC> r0 = dl.load(idx, kb);
    scur = dl.combine(r0, kb);
    oldc = r0.old;
    codec = r0.code;
    tile[kb % 3][ty + 1][lane] = scur;
    if (hrow_ok) tile[kb % 3][hslot][lane] = dl.combine(dl.load(idx + hoff, kb), kb);
    if (has_e) {
      const T e = dl.combine1(dl.load1(idx + eoff, kb), kb);
      if (lane == 0) leftc = e; else rightc = e;
    }
    xc = vload<double, VEC>(x, idx);
    raw = dl.load(idx + plane, kb + 1);
    if (kb + 1 < ke) {
      if (hrow_ok) hraw = dl.load(idx + plane + hoff, kb + 1);
      if (has_e) eraw = dl.load1(idx + plane + eoff, kb + 1);
      xn = vload<double, VEC>(x, idx + plane);
    }
  }

  for (int c = kb; c < ke; ++c, idx += plane) {
    const int tb = c % 3, tn = (c + 1) % 3;
    VecT<T, VEC> oldn = zero;
    VecT<uint8_t, VEC> coden = czero;
    T leftn = T(0), rightn = T(0);
    if (active) {
      // plane c+1: operands were requested one iteration ago
      sp = dl.combine(raw, c + 1);
      oldn = raw.old;
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
      // x += alpha_{it-1} * s_{it-1}  (fp64 accumulator, requested >= 1 plane ago)
#pragma unroll
      for (int e = 0; e < VEC; ++e) xc.v[e] += alpha_prev * (double)oldc.v[e];
      vstore<double, VEC>(x, idx, xc);
    }
    sm = scur;
    scur = sp;
    oldc = oldn;
    codec = coden;
    leftc = leftn;
    rightc = rightn;
    xc = xn;
    if (active && c + 2 < ke) xn = vload<double, VEC>(x, idx + 2 * plane);
  }

  const double bs = block_sum(acc, red);
  const unsigned nb = num_blocks();
  if (tid == 0) partials[linear_block()] = bs;
  if (last_block(&sc->ticketA, nb)) {
    double t = 0.0;
    for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(partials + b);
    t = block_sum(t, red);
    if (tid == 0) sc->gA[sc->rank] = t;
  }
}

// ---- phase A, cp.async ring variant (fp32, VEC = 4) ---------------------------
// Same arithmetic as k_phaseA.  B200 needs ~100 KB of loads in flight per SM to
// saturate HBM3e; a register software pipeline cannot hold that (every byte in
// flight pins a register).  Here each thread streams the raw operands of the
// next RING_D - 1 planes (r, s, codes, fp64 x; plus the CTA-edge rows / lanes)
// into its own slots of a shared-memory ring with cp.async (LDGSTS) and only
// touches them after cp.async.wait_group, so registers hold just the three
// combined planes.  The slots are thread-private: no extra barrier is needed.
constexpr int RING_D = 4;

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

// TMA-engine bulk prefetch into L2 (UBLKPF): one lane moves a whole tile row
// from HBM towards the SM without holding registers or shared memory.
__device__ __forceinline__ void prefetch_l2(const void* gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}
#ifndef SPL_L2_AHEAD
#define SPL_L2_AHEAD 3     // planes between the L2 prefetch and the cp.async request
#endif

struct alignas(16) RingStage {
  float4 a[256];          // z or r of the thread's 4 cells
  float4 old[256];        // previous search direction
  float4 x[2][256];       // fp64 pressure accumulator (4 doubles = 2 x 16 B)
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
  double beta, alpha;
  int stop;
};

template <int PC>
__global__ void __launch_bounds__(32 * TILE_Y, 2)
k_phaseA_ring(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ zr,
              const float* __restrict__ sold, float* __restrict__ snew, float* __restrict__ q,
              double* __restrict__ x, Scal* __restrict__ sc, double* __restrict__ partials,
              int par, int zchunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSmem& sm_ = *reinterpret_cast<RingSmem*>(smem_raw);
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<float>(sm_.lut);
  if (tid == 0) {
    int stop = sc->done;
    double beta = 0.0, alpha_prev = 0.0;
    if (!stop) {
      const double rho1 = rho_of(sc, par ^ 1);
      const double rho2 = rho_of(sc, par);
      const double rmax = rmax_of(sc, par ^ 1);
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
        beta = rho1 / rho2;
        if (sc->count > 0) alpha_prev = rho2 / sum_slots(sc->gA, sc->world);
      }
    }
    sm_.stop = stop;
    sm_.beta = beta;
    sm_.alpha = alpha_prev;
  }
  __syncthreads();
  if (sm_.stop) return;

  const float beta = (float)sm_.beta;
  const double alpha_prev = sm_.alpha;
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

  // request every operand of plane p into ring slot p % RING_D
  auto request = [&](int p) {
    if (active && p <= ke) {
      RingStage& st = sm_.st[p % RING_D];
      const long long idx = idx0 + (long long)(p - kb) * plane;
      const bool remote = (p < 0) || (p >= nz);
      if (remote) {
        cp_async16(&st.a[tid], snew + idx);          // halo of s' already holds the new direction
      } else {
        cp_async16(&st.a[tid], zr + idx);
        cp_async16(&st.old[tid], sold + idx);
        cp_async4(&st.code[tid], codes + idx);
      }
      if (p < ke) {                                    // centre planes only
        cp_async16(&st.x[0][tid], x + idx);
        cp_async16(&st.x[1][tid], x + idx + 2);
        if (has_hrow) {
          cp_async16(&st.ha[hidx], zr + idx + hoff);
          cp_async16(&st.hold[hidx], sold + idx + hoff);
          cp_async4(&st.hcode[hidx], codes + idx + hoff);
        }
        if (has_e) {
          cp_async4(&st.ea[eidx], zr + idx + eoff);
          cp_async4(&st.eold[eidx], sold + idx + eoff);
          // 4-byte aligned word that contains the neighbour's code byte
          const long long cidx = idx + eoff;
          cp_async4(&st.ecode[eidx], codes + (cidx & ~3LL));
        }
      }
    }
    cp_async_commit();
  };
  auto dirf = [&](float a, float o, unsigned code) {
    const float zz = (PC == PC_JACOBI) ? a * lut[lut_index(code)] : a;
    return zz + beta * o;
  };
  // combined plane p (own cells) from its ring slot; also the tile halo row and
  // the CTA-edge x neighbour when p is a centre plane
  auto combine_own = [&](int p) {
    const RingStage& st = sm_.st[p % RING_D];
    const float4 a = st.a[tid];
    if (p < 0 || p >= nz) return a;
    const float4 o = st.old[tid];
    const unsigned c = st.code[tid];
    float4 r;
    r.x = dirf(a.x, o.x, c & 0xffu);
    r.y = dirf(a.y, o.y, (c >> 8) & 0xffu);
    r.z = dirf(a.z, o.z, (c >> 16) & 0xffu);
    r.w = dirf(a.w, o.w, c >> 24);
    return r;
  };
  auto combine_hrow = [&](int p) {
    const RingStage& st = sm_.st[p % RING_D];
    const float4 a = st.ha[hidx];
    const float4 o = st.hold[hidx];
    const unsigned c = st.hcode[hidx];
    float4 r;
    r.x = dirf(a.x, o.x, c & 0xffu);
    r.y = dirf(a.y, o.y, (c >> 8) & 0xffu);
    r.z = dirf(a.z, o.z, (c >> 16) & 0xffu);
    r.w = dirf(a.w, o.w, c >> 24);
    return r;
  };
  auto combine_edge = [&](int p) {
    const RingStage& st = sm_.st[p % RING_D];
    const long long cidx = idx0 + (long long)(p - kb) * plane + eoff;
    const unsigned c = (st.ecode[eidx] >> (8 * (int)(cidx & 3LL))) & 0xffu;
    return dirf(st.ea[eidx], st.eold[eidx], c);
  };

  // ---- prologue ---------------------------------------------------------------
  float4 smv = make_float4(0.f, 0.f, 0.f, 0.f), scur = smv, sp = smv;
  float leftc = 0.f, rightc = 0.f;
  for (int p = kb; p < kb + RING_D - 1; ++p) request(p);    // RING_D - 1 groups
  if (active) {                                             // plane kb-1: registers
    const long long idm = idx0 - plane;
    if (kb - 1 < 0) {
      smv = *reinterpret_cast<const float4*>(snew + idm);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(zr + idm);
      const float4 o = *reinterpret_cast<const float4*>(sold + idm);
      const unsigned c = *reinterpret_cast<const unsigned*>(codes + idm);
      smv.x = dirf(a.x, o.x, c & 0xffu);
      smv.y = dirf(a.y, o.y, (c >> 8) & 0xffu);
      smv.z = dirf(a.z, o.z, (c >> 16) & 0xffu);
      smv.w = dirf(a.w, o.w, c >> 24);
    }
  }
  cp_async_wait<RING_D - 2>();                              // plane kb has landed
  if (active) {
    scur = combine_own(kb);
    sm_.tile[kb % 3][ty + 1][lane] = scur;
    if (has_hrow) sm_.tile[kb % 3][hslot][lane] = combine_hrow(kb);
    if (has_e) {
      const float e = combine_edge(kb);
      if (lane == 0) leftc = e; else rightc = e;
    }
  }

  double acc = 0.0;
  long long idx = idx0;
  // HBM -> L2: lane 0 of every warp prefetches its tile row (r, s, codes, x) of
  // the plane SPL_L2_AHEAD beyond the newest cp.async request
  const bool pf_lane = (lane == 0) && (j < g.ny) && (blockIdx.x * 128 < g.nx);
  const int pf_cells = min(128, g.nx - (int)blockIdx.x * 128);
  auto prefetch_plane = [&](int p) {
    if (SPL_L2_AHEAD > 0 && pf_lane && p < ke && p < nz) {
      const long long id = idx0 + (long long)(p - kb) * plane;   // lane 0: tile row start
      prefetch_l2(zr + id, pf_cells * 4);
      prefetch_l2(sold + id, pf_cells * 4);
      prefetch_l2(x + id, pf_cells * 8);
      if ((pf_cells & 15) == 0) prefetch_l2(codes + id, pf_cells);
    }
  };
  for (int p = kb + RING_D - 1; p < kb + RING_D - 1 + SPL_L2_AHEAD; ++p) prefetch_plane(p);
  for (int c = kb; c < ke; ++c, idx += plane) {
    prefetch_plane(c + RING_D - 1 + SPL_L2_AHEAD);
    request(c + RING_D - 1);          // into the slot of plane c-1 (fully consumed)
    cp_async_wait<RING_D - 2>();      // plane c+1 has landed
    float leftn = 0.f, rightn = 0.f;
    if (active) {
      sp = combine_own(c + 1);
      if (c + 1 < ke) {
        sm_.tile[(c + 1) % 3][ty + 1][lane] = sp;
        if (has_hrow) sm_.tile[(c + 1) % 3][hslot][lane] = combine_hrow(c + 1);
        if (has_e) {
          const float e = combine_edge(c + 1);
          if (lane == 0) leftn = e; else rightn = e;
        }
      }
    }
    __syncthreads();   // tile[c % 3] was completed before the previous barrier
    const float lsh = __shfl_up_sync(0xffffffffu, scur.w, 1);
    const float rsh = __shfl_down_sync(0xffffffffu, scur.x, 1);
    const float left = (lane != 0) ? lsh : leftc;
    const float right = (lane != 31) ? rsh : rightc;
    if (active) {
      const RingStage& st = sm_.st[c % RING_D];
      const float4 sym = sm_.tile[c % 3][ty][lane];
      const float4 syp = sm_.tile[c % 3][ty + 2][lane];
      const unsigned code4 = st.code[tid];
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
        qa[e] = lap_cell<float>((code4 >> (8 * e)) & 0xffu, s[e], xm, xp, ym[e], yp[e], zm[e],
                                zp[e]);
        part += s[e] * qa[e];
      }
      acc += (double)part;   // 4 products in fp32, accumulated in fp64
      *reinterpret_cast<float4*>(q + idx) = make_float4(qa[0], qa[1], qa[2], qa[3]);
      *reinterpret_cast<float4*>(snew + idx) = scur;
      // x += alpha_{it-1} * s_{it-1}  (fp64 accumulator)
      const float4 o = st.old[tid];
      const float4 x01 = st.x[0][tid], x23 = st.x[1][tid];
      double2 xa = *reinterpret_cast<const double2*>(&x01);
      double2 xb = *reinterpret_cast<const double2*>(&x23);
      xa.x += alpha_prev * (double)o.x;
      xa.y += alpha_prev * (double)o.y;
      xb.x += alpha_prev * (double)o.z;
      xb.y += alpha_prev * (double)o.w;
      *reinterpret_cast<double2*>(x + idx) = xa;
      *reinterpret_cast<double2*>(x + idx + 2) = xb;
    }
    smv = scur;
    scur = sp;
    leftc = leftn;
    rightc = rightn;
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

// ---- phase B ---------------------------------------------------------------
// init != 0: alpha forced to 0 (used once per solve to produce rho_0, max|b|).
template <typename T, int VEC, int PC>
__global__ void __launch_bounds__(256)
k_phaseB(Geo g, const uint8_t* __restrict__ codes, T* __restrict__ r,
         const T* __restrict__ q, const T* __restrict__ z, Scal* __restrict__ sc,
         double* __restrict__ partials, int par, int init) {
  __shared__ double red[32];
  __shared__ double s_alpha;
  __shared__ T lut[8];
  const int tid = threadIdx.x;
  if (sc->done) return;
  fill_inv_lut<T>(lut);
  if (tid == 0) {
    double alpha = 0.0;
    if (!init) {
      const double rho1 = rho_of(sc, par ^ 1);
      const double sq = sum_slots(sc->gA, sc->world);
      alpha = rho1 / sq;
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
      if (PC != PC_RBIC0) sc->gBM[par][rank][0] = t;
      sc->gBM[par][rank][1] = nanf ? kNaN : m;
      if (!init) sc->count = sc->count + 1;
    }
  }
}

// ---- last pending x update ----------------------------------------------------
// Phase A(it) adds alpha_{it-1} s_{it-1}; when the loop ends after K iterations
// the last term alpha_K s_K is still pending.  s_K sits in buffer sb for odd K
// and sa for even K (ping-pong), alpha_K = rho_{K-1} / (s_K . q_K).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
k_xflush(Geo g, double* __restrict__ x, const T* __restrict__ sa, const T* __restrict__ sb,
         const Scal* __restrict__ sc) {
  const int K = sc->count;
  if (K <= 0) return;
  const T* __restrict__ s = (K & 1) ? sb : sa;
  const double alpha = rho_of(sc, (K - 1) & 1) / sum_slots(sc->gA, sc->world);
  if (!(fabs(alpha) <= 1.7976931348623157e308)) return;
  const long long nq = g.ncell / VEC;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nq;
       t += (long long)gridDim.x * blockDim.x) {
    const long long idx = t * VEC;
    const VecT<T, VEC> sv = vload<T, VEC>(s, idx);
    VecT<double, VEC> xv = vload<double, VEC>(x, idx);
#pragma unroll
    for (int e = 0; e < VEC; ++e) xv.v[e] += alpha * (double)sv.v[e];
    vstore<double, VEC>(x, idx, xv);
  }
}

}  // namespace spl
