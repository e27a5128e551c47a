// spl_pcg.cuh -- matrix-free PCG on the fluid-cell Laplacian (sm_100a)
//
// Replaces Pressure.calculate's assembly + direct solve (Pressure.fs:42-69):
// A is never formed; one PCG iteration is two kernels separated by
// the two global reductions CG needs:
//
//   phase A : s' = z + beta s ; q = A s' ; partial s'.q          (17 B/cell)
//   phase B : x += alpha s' ; r -= alpha q ; partial r.z, max|r| (24 B/cell Jacobi)
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

// ---- phase A ---------------------------------------------------------------
// CTA = 32 x TILE_Y threads; a thread owns VEC consecutive x cells of one row
// and marches along z keeping the new search direction of planes k-1, k, k+1
// in registers.  +-x neighbours come from warp shuffles, +-y neighbours are
// recomputed from L1-resident rows, so every array is pulled from HBM once.
template <typename T, int VEC, int PC>
struct PhaseA {
  const uint8_t* __restrict__ codes;
  const T* __restrict__ zr;     // z (PC_RBIC0), r (PC_NONE / PC_JACOBI)
  const T* __restrict__ sold;
  T* __restrict__ snew;
  Geo g;
  T beta;

  __device__ __forceinline__ VecT<T, VEC> dir_vec(long long idx, int k) const {
    VecT<T, VEC> out;
    if (k < 0 || k >= g.nz) {
      // neighbouring slab's plane: already the new direction (halo exchange),
      // zeros at the global ends
      return vload<T, VEC>(snew, idx);
    }
    const VecT<T, VEC> a = vload<T, VEC>(zr, idx);
    const VecT<T, VEC> b = vload<T, VEC>(sold, idx);
    if (PC == PC_JACOBI) {
      const VecT<uint8_t, VEC> c = cload<VEC>(codes, idx);
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        out.v[e] = a.v[e] * inv_diag<T>(c.v[e]) + beta * b.v[e];
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) out.v[e] = a.v[e] + beta * b.v[e];
    }
    return out;
  }
  __device__ __forceinline__ T dir_one(long long idx, int k) const {
    if (k < 0 || k >= g.nz) return snew[idx];
    const T a = zr[idx];
    const T b = sold[idx];
    if (PC == PC_JACOBI) return a * inv_diag<T>(codes[idx]) + beta * b;
    return a + beta * b;
  }
};

template <typename T, int VEC, int PC>
__global__ void __launch_bounds__(32 * TILE_Y)
k_phaseA(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ zr,
         const T* __restrict__ sold, T* __restrict__ snew, T* __restrict__ q,
         Scal* __restrict__ sc, double* __restrict__ partials, int par,
         int zchunk) {
  __shared__ double red[32];
  __shared__ double s_beta;
  __shared__ int s_stop;
  const int tid = threadIdx.x + threadIdx.y * 32;
  if (tid == 0) {
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
    s_stop = stop;
    s_beta = beta;
  }
  __syncthreads();
  if (s_stop) return;

  PhaseA<T, VEC, PC> pa{codes, zr, sold, snew, g, T(s_beta)};
  const int lane = threadIdx.x;
  const int i0 = (blockIdx.x * 32 + lane) * VEC;
  const int j = blockIdx.y * TILE_Y + threadIdx.y;
  const bool active = (i0 < g.nx) && (j < g.ny);
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const long long row = g.nx;

  VecT<T, VEC> zero;
#pragma unroll
  for (int e = 0; e < VEC; ++e) zero.v[e] = T(0);

  double acc = 0.0;
  long long idx = ((long long)kb * g.ny + j) * row + i0;
  VecT<T, VEC> sm = zero, scur = zero, sp = zero;
  if (active) {
    sm = pa.dir_vec(idx - g.plane, kb - 1);
    scur = pa.dir_vec(idx, kb);
  }
  for (int k = kb; k < ke; ++k, idx += g.plane) {
    VecT<T, VEC> sym = zero, syp = zero;
    VecT<uint8_t, VEC> code;
#pragma unroll
    for (int e = 0; e < VEC; ++e) code.v[e] = 0;
    T left = T(0), right = T(0);
    if (active) {
      sp = pa.dir_vec(idx + g.plane, k + 1);
      // rows j-1 / j+1: memory is always valid (z halo planes), values are
      // masked by the code bits when the neighbour is outside the box
      sym = pa.dir_vec(idx - row, k);
      syp = pa.dir_vec(idx + row, k);
      code = cload<VEC>(codes, idx);
      // CTA-edge lanes fetch the x neighbour owned by the adjacent CTA
      if (lane == 0 && i0 > 0) left = pa.dir_one(idx - 1, k);
      if (lane == 31 && i0 + VEC < g.nx) right = pa.dir_one(idx + VEC, k);
    }
    const T lsh = __shfl_up_sync(0xffffffffu, scur.v[VEC - 1], 1);
    const T rsh = __shfl_down_sync(0xffffffffu, scur.v[0], 1);
    if (lane != 0) left = lsh;
    if (lane != 31) right = rsh;
    if (active) {
      VecT<T, VEC> qv;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const unsigned c = code.v[e];
        const T s0 = scur.v[e];
        const T xm = (e > 0) ? scur.v[e > 0 ? e - 1 : 0] : left;
        const T xp = (e < VEC - 1) ? scur.v[e < VEC - 1 ? e + 1 : 0] : right;
        // Pressure.fs:44-55: N_i s_i - sum over Fluid neighbours (s = 0 on Air)
        T a = T(0);
        a += (c & B_XM) ? (s0 - xm) : T(0);
        a += (c & B_XP) ? (s0 - xp) : T(0);
        a += (c & B_YM) ? (s0 - sym.v[e]) : T(0);
        a += (c & B_YP) ? (s0 - syp.v[e]) : T(0);
        a += (c & B_ZM) ? (s0 - sm.v[e]) : T(0);
        a += (c & B_ZP) ? (s0 - sp.v[e]) : T(0);
        a = (c & B_FLUID) ? a : T(0);
        qv.v[e] = a;
        acc += (double)s0 * (double)a;
      }
      vstore<T, VEC>(q, idx, qv);
      vstore<T, VEC>(snew, idx, scur);
    }
    sm = scur;
    scur = sp;
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

// The right-most active lane of a ragged row shuffles a 0 from its inactive
// neighbour; the value is masked because that cell's B_XP bit is clear.

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
k_phaseB(Geo g, const uint8_t* __restrict__ codes, T* __restrict__ x,
         T* __restrict__ r, const T* __restrict__ s, const T* __restrict__ q,
         const T* __restrict__ z, Scal* __restrict__ sc,
         double* __restrict__ partials, int par, int init) {
  __shared__ double red[32];
  __shared__ double s_alpha;
  const int tid = threadIdx.x;
  if (sc->done) return;
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
      VecT<T, VEC> xv = vload<T, VEC>(x, idx);
      const VecT<T, VEC> sv = vload<T, VEC>(s, idx);
      const VecT<T, VEC> qv = vload<T, VEC>(q, idx);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        xv.v[e] += alpha * sv.v[e];
        rv.v[e] -= alpha * qv.v[e];
      }
      vstore<T, VEC>(x, idx, xv);
      vstore<T, VEC>(r, idx, rv);
    }
    VecT<uint8_t, VEC> c;
    if (PC == PC_JACOBI) c = cload<VEC>(codes, idx);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const double rr = (double)rv.v[e];
      if (PC == PC_JACOBI) rz += rr * (double)(rv.v[e] * inv_diag<T>(c.v[e]));
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

}  // namespace spl
