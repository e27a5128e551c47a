// spl_mg.cuh -- geometric multigrid V-cycle as a PCG preconditioner (SPL_PC_MG)
//
// Opt-in extension beyond the preconditioner family BASELINE.json names (MIC(0)/IC(0)
// on a parallel ordering -> SPL_PC_RBIC0): the reference solves A p = b exactly
// (Pressure.fs:69), and what decides the wall time of a converged projection on
// 256^3+ grids is the iteration count -- Jacobi-PCG needs ~3000 iterations at 512^3,
// MG-PCG ~30.  Scheme (after McAdams, Sifakis, Teran 2010, simplified):
//   * levels by 2x coarsening of the media labels: a coarse cell is Air if any child
//     is Air, else Fluid if any child is Fluid, else Solid (children outside the box
//     count as Solid); operator on level l = 4^-l * (integer Laplacian of the level's
//     labels) -- rediscretisation, same matrix-free stencil codes as the fine level;
//   * smoother: damped Jacobi (omega = 2/3), nu pre- and post-sweeps;
//   * restriction = average of the 8 children, prolongation = piecewise constant
//     (P = 8 R^T), coarsest level: a fixed number of damped-Jacobi sweeps from zero.
// Symmetric (same pre/post smoother, R ~ P^T, zero initial guess) => valid SPD
// preconditioner.  Oracle: oracle/dense_ref.py MG / pcg(precond=PC_MG).
// All kernels are plain one-thread-per-cell stencils (HBM-bound, ~13 B/cell/sweep).
#pragma once
#include "spl_common.cuh"

namespace spl {

constexpr int MG_MAX_LEVELS = 12;

__device__ __forceinline__ void mg_ijk(const Geo& g, long long idx, int& i, int& j, int& k) {
  i = (int)(idx % g.nx);
  const long long t = idx / g.nx;
  j = (int)(t % g.ny);
  k = (int)(t / g.ny);
}

// labels of the next coarser level
__global__ void k_mg_coarsen(Geo gf, const uint8_t* __restrict__ fine, Geo gc,
                             uint8_t* __restrict__ coarse) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < gc.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    mg_ijk(gc, idx, i, j, k);
    bool any_air = false, any_fluid = false;
    for (int c = 0; c < 8; ++c) {
      const int fi = 2 * i + (c & 1), fj = 2 * j + ((c >> 1) & 1), fk = 2 * k + (c >> 2);
      if (fi < gf.nx && fj < gf.ny && fk < gf.nz) {
        const uint8_t m = fine[((long long)fk * gf.ny + fj) * gf.nx + fi];
        any_air |= (m == AIR);
        any_fluid |= (m == FLUID);
      }
    }
    coarse[idx] = any_air ? AIR : (any_fluid ? FLUID : SOLID);
  }
}

// L u at one cell: N u_i - sum over in-box non-solid neighbours (u = 0 on non-Fluid)
template <typename T>
__device__ __forceinline__ T mg_lap(const Geo& g, unsigned c, long long idx,
                                    const T* __restrict__ u, T u0) {
  T sum = T(0);
  if (c & B_XM) sum += u[idx - 1];
  if (c & B_XP) sum += u[idx + 1];
  if (c & B_YM) sum += u[idx - g.nx];
  if (c & B_YP) sum += u[idx + g.nx];
  if (c & B_ZM) sum += u[idx - g.plane];
  if (c & B_ZP) sum += u[idx + g.plane];
  return T(__popc(c & 63u)) * u0 - sum;
}

// damped Jacobi: u_out = u_in + omega D^-1 (b - A u_in), A = scale * L.
// first != 0: u_in = 0 (u_in is not read)
template <typename T>
__global__ void __launch_bounds__(256)
k_mg_smooth(Geo g, const uint8_t* __restrict__ codes, const T* __restrict__ uin,
            const T* __restrict__ b, T* __restrict__ uout, T omega, T inv_scale, int first,
            const Scal* __restrict__ sc) {
  if (sc->done) return;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const unsigned c = codes[idx];
    T out = T(0);
    const int n = __popc(c & 63u);
    if ((c & B_FLUID) && n > 0) {
      const T w = omega / T(n);
      if (first) {
        out = w * (b[idx] * inv_scale);
      } else {
        const T u0 = uin[idx];
        out = u0 + w * (b[idx] * inv_scale - mg_lap<T>(g, c, idx, uin, u0));
      }
    }
    uout[idx] = out;
  }
}

// b_coarse = average over the 8 children of (b - A u), 0 on non-Fluid coarse cells
template <typename T>
__global__ void __launch_bounds__(256)
k_mg_restrict(Geo gf, const uint8_t* __restrict__ codes_f, const T* __restrict__ u,
              const T* __restrict__ b, T scale, Geo gc, const uint8_t* __restrict__ codes_c,
              T* __restrict__ bc, const Scal* __restrict__ sc) {
  if (sc->done) return;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < gc.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    T acc = T(0);
    if (codes_c[idx] & B_FLUID) {
      int i, j, k;
      mg_ijk(gc, idx, i, j, k);
      for (int c = 0; c < 8; ++c) {
        const int fi = 2 * i + (c & 1), fj = 2 * j + ((c >> 1) & 1), fk = 2 * k + (c >> 2);
        if (fi < gf.nx && fj < gf.ny && fk < gf.nz) {
          const long long f = ((long long)fk * gf.ny + fj) * gf.nx + fi;
          const unsigned cf = codes_f[f];
          if (cf & B_FLUID) acc += b[f] - scale * mg_lap<T>(gf, cf, f, u, u[f]);
        }
      }
      acc *= T(0.125);
    }
    bc[idx] = acc;
  }
}

// u_fine += e_coarse[parent] on Fluid cells
template <typename T>
__global__ void __launch_bounds__(256)
k_mg_prolong(Geo gf, const uint8_t* __restrict__ codes_f, T* __restrict__ u, Geo gc,
             const T* __restrict__ ec, const Scal* __restrict__ sc) {
  if (sc->done) return;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < gf.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    if (codes_f[idx] & B_FLUID) {
      int i, j, k;
      mg_ijk(gf, idx, i, j, k);
      u[idx] += ec[((long long)(k >> 1) * gc.ny + (j >> 1)) * gc.nx + (i >> 1)];
    }
  }
}

// rho = r . z into the PCG slot (deterministic last-CTA reduction)
template <typename T>
__global__ void __launch_bounds__(256)
k_dot_rz(Geo g, const T* __restrict__ r, const T* __restrict__ z, Scal* __restrict__ sc,
         double* __restrict__ partials, int par_slot) {
  __shared__ double red[32];
  if (sc->done) return;
  double acc = 0.0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x)
    acc += (double)r[idx] * (double)z[idx];
  const double bs = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = bs;
  if (last_block(&sc->ticketC, gridDim.x)) {
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(partials + b);
    t = block_sum(t, red);
    if (threadIdx.x == 0) sc->gBM[par_slot][sc->rank][0] = t;
  }
}

}  // namespace spl
