// spl_stencil.cuh -- label codes, divergence / RHS, gradient subtraction,
// validators and forces (sm_100a).  All HBM-bound, one thread per cell (or per
// VEC cells), coalesced along x.
#pragma once
#include "spl_common.cuh"

namespace spl {

__device__ __forceinline__ bool is_nonsolid(uint8_t m) { return m == AIR || m == FLUID; }

// ---- labels -> stencil codes -------------------------------------------------
// Pressure.fs:34-38,53 (number_neighbors with media_is_not_solid) and :42-43
// (markers = Fluid cells).  Out-of-box / absent neighbours do not count.
__global__ void k_codes(Geo g, const uint8_t* __restrict__ media,
                        uint8_t* __restrict__ codes) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const int kg = k + g.z0;
    unsigned c = 0;
    if (i > 0 && is_nonsolid(media[idx - 1])) c |= B_XM;
    if (i < g.nx - 1 && is_nonsolid(media[idx + 1])) c |= B_XP;
    if (j > 0 && is_nonsolid(media[idx - g.nx])) c |= B_YM;
    if (j < g.ny - 1 && is_nonsolid(media[idx + g.nx])) c |= B_YP;
    if (kg > 0 && is_nonsolid(media[idx - g.plane])) c |= B_ZM;
    if (kg < g.nzg - 1 && is_nonsolid(media[idx + g.plane])) c |= B_ZP;
    // only Fluid cells (the unknowns) carry a code: every consumer tests the
    // Fluid bit first, and code == 0 lets popc(code & 63) index the 1/diag table
    c = (media[idx] == FLUID) ? (c | B_FLUID) : 0u;
    codes[idx] = (uint8_t)c;
  }
}

// ---- divergence + RHS ----------------------------------------------------------
// Pressure.fs:16-32: outgoing = own +faces gated by the +neighbour being
// non-solid, incoming = -neighbours' +faces gated by the -neighbour being
// non-solid; (in - out) per component, then x + y + z; not divided by h.
template <typename T>
__device__ __forceinline__ T cell_divergence(const Geo& g, unsigned c, long long idx,
                                             const T* __restrict__ u,
                                             const T* __restrict__ v,
                                             const T* __restrict__ w) {
  const T ox = (c & B_XP) ? u[idx] : T(0);
  const T oy = (c & B_YP) ? v[idx] : T(0);
  const T oz = (c & B_ZP) ? w[idx] : T(0);
  const T ix = (c & B_XM) ? u[idx - 1] : T(0);
  const T iy = (c & B_YM) ? v[idx - g.nx] : T(0);
  const T iz = (c & B_ZM) ? w[idx - g.plane] : T(0);
  return ((ix - ox) + (iy - oy)) + (iz - oz);
}

// scale = h rho / dt (Pressure.fs:64): out = scale * divergence on Fluid, else 0.
// scale = 1 gives the plain divergence (spl_download's `div`).
template <typename T>
__global__ void __launch_bounds__(256)
k_div_rhs(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ u,
          const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ out,
          T scale) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const unsigned c = codes[idx];
    T d = T(0);
    if (c & B_FLUID) d = scale * cell_divergence<T>(g, c, idx, u, v, w);
    out[idx] = d;
  }
}

// ---- gradient subtraction --------------------------------------------------------
// Pressure.fs:96-129 with every face updated exactly once (SURVEY A.7
// intended): the face between c and c+e gets  vel -= (dt/(h rho)) (p[c+e]-p[c])
// when both cells are non-solid and at least one is Fluid (p = 0 on Air,
// Constants.fs:20); a face between Fluid and Solid/absent takes the solid
// velocity 0; every other face is left alone.
template <typename T>
__global__ void __launch_bounds__(256)
k_grad_sub(Geo g, const uint8_t* __restrict__ media, const double* __restrict__ p,
           T* __restrict__ u, T* __restrict__ v, T* __restrict__ w, double kappa) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const int kg = k + g.z0;
    const uint8_t m0 = media[idx];
    const bool ns0 = is_nonsolid(m0), fl0 = (m0 == FLUID);
    const double p0 = fl0 ? p[idx] : 0.0;
    const bool inb[3] = {i < g.nx - 1, j < g.ny - 1, kg < g.nzg - 1};
    const long long off[3] = {1, (long long)g.nx, g.plane};
    T* f[3] = {u, v, w};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const uint8_t m1 = inb[a] ? media[idx + off[a]] : ABSENT;
      const bool ns1 = is_nonsolid(m1), fl1 = (m1 == FLUID);
      if (ns0 && ns1 && (fl0 || fl1)) {
        // the pressure difference is formed in fp64 (p is the fp64 accumulator)
        const double p1 = fl1 ? p[idx + off[a]] : 0.0;
        f[a][idx] = (T)((double)f[a][idx] - kappa * (p1 - p0));
      } else if ((fl0 && !ns1) || (!ns0 && fl1)) {
        f[a][idx] = T(0);
      }
    }
  }
}

// ---- validators -----------------------------------------------------------------
// Pressure.fs:132-149 + Vector.fs:15-18: max |divergence| over Fluid cells and
// the number of non-finite velocity / pressure entries.
template <typename T>
__global__ void __launch_bounds__(256)
k_check(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ u,
        const T* __restrict__ v, const T* __restrict__ w, const double* __restrict__ p,
        Scal* __restrict__ sc, double* __restrict__ partials) {
  __shared__ double red[32];
  double mx = 0.0;
  unsigned long long bad = 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const unsigned c = codes[idx];
    const T a = u[idx], b = v[idx], d = w[idx];
    bad += !isfinite((double)a);
    bad += !isfinite((double)b);
    bad += !isfinite((double)d);
    if (c & B_FLUID) {
      bad += !isfinite(p[idx]);
      const double dv = fabs((double)cell_divergence<T>(g, c, idx, u, v, w));
      if (dv == dv) mx = fmax(mx, dv);
    }
  }
  const double bm = block_max(mx, red);
  const double bb = block_sum((double)bad, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bm;
    if (bb > 0.0) atomicAdd(&sc->nonfinite, (unsigned long long)bb);
  }
  if (last_block(&sc->ticketC, gridDim.x)) {
    double m = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      m = fmax(m, __ldcg(partials + b));
    m = block_max(m, red);
    if (threadIdx.x == 0) sc->gD[sc->rank] = m;
  }
}

// ---- forces ---------------------------------------------------------------------
// Forces.fs:6-12: v += gravity * dt on every Fluid cell.
template <typename T>
__global__ void k_forces(Geo g, const uint8_t* __restrict__ media, T* __restrict__ u,
                         T* __restrict__ v, T* __restrict__ w, T gx, T gy, T gz) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    if (media[idx] == FLUID) {
      u[idx] += gx;
      v[idx] += gy;
      w[idx] += gz;
    }
  }
}

}  // namespace spl
