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

// restriction = mean of the children: 1 / 2^(number of coarsened axes).  An axis of
// extent 1 (2-D boxes) is not coarsened, so its "second child" does not exist.
__host__ __device__ inline double mg_restrict_weight(const Geo& fine) {
  int axes = (fine.nx > 1) + (fine.ny > 1) + (fine.nz > 1);
  return axes == 3 ? 0.125 : (axes == 2 ? 0.25 : (axes == 1 ? 0.5 : 1.0));
}

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
      acc *= T(mg_restrict_weight(gf));
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

// ---- bottom of the V-cycle in one CTA ----------------------------------------------------
// Levels with at most MG_BOTTOM_CELLS cells (16^3 and coarser) live entirely in shared
// memory: the whole sub-cycle -- pre-smoothing, restriction, the coarsest-level sweeps,
// prolongation, post-smoothing -- is one launch instead of ~35 launch-latency-bound ones.
// Same arithmetic and order of operations as the per-level kernels above.
constexpr int MG_BOTTOM_CELLS = 4096;
constexpr int MG_BOTTOM_LEVELS = 4;
constexpr int MG_MAX_NU = 8;

template <typename T>
struct MgBottomArgs {
  int nlev, nu, coarse_sweeps;
  Geo g[MG_BOTTOM_LEVELS];
  const uint8_t* codes[MG_BOTTOM_LEVELS];
  T scale[MG_BOTTOM_LEVELS];
  T omega[MG_MAX_NU];
  const T* b0;       // right-hand side of the first bottom level (global)
  T* u0;             // its solution (global)
  const Scal* sc;
};

template <typename T>
__host__ __device__ inline size_t mg_bottom_smem(const Geo* g, int nlev) {
  size_t bytes = 0;
  for (int l = 0; l < nlev; ++l) {
    const size_t n = (size_t)g[l].ncell;
    bytes += 3 * ((n * sizeof(T) + 15) & ~size_t(15)) + ((n + 15) & ~size_t(15));
  }
  return bytes;
}

template <typename T>
__device__ __forceinline__ T mg_lap_s(const Geo& g, unsigned c, int idx, const T* u, T u0) {
  T sum = T(0);
  if (c & B_XM) sum += u[idx - 1];
  if (c & B_XP) sum += u[idx + 1];
  if (c & B_YM) sum += u[idx - g.nx];
  if (c & B_YP) sum += u[idx + g.nx];
  if (c & B_ZM) sum += u[idx - (int)g.plane];
  if (c & B_ZP) sum += u[idx + (int)g.plane];
  return T(__popc(c & 63u)) * u0 - sum;
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_mg_bottom(const MgBottomArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_bt[];
  if (a.sc->done) return;
  T* U[MG_BOTTOM_LEVELS];
  T* V[MG_BOTTOM_LEVELS];
  T* B[MG_BOTTOM_LEVELS];
  uint8_t* C[MG_BOTTOM_LEVELS];
  {
    unsigned char* p = smem_bt;
    for (int l = 0; l < a.nlev; ++l) {
      const size_t n = (size_t)a.g[l].ncell;
      const size_t fb = (n * sizeof(T) + 15) & ~size_t(15);
      U[l] = reinterpret_cast<T*>(p); p += fb;
      V[l] = reinterpret_cast<T*>(p); p += fb;
      B[l] = reinterpret_cast<T*>(p); p += fb;
      C[l] = p; p += (n + 15) & ~size_t(15);
    }
  }
  const int tid = threadIdx.x, nt = blockDim.x;
  // the sub-cycle is local to this z-slab: faces towards a neighbouring slab are closed
  // (the z bits of the first / last local plane are cleared; they are 0 at the global
  // ends anyway), so with world > 1 the bottom of the cycle is block-diagonal
  for (int l = 0; l < a.nlev; ++l) {
    const int pl = (int)a.g[l].plane, nc = (int)a.g[l].ncell;
    for (int i = tid; i < nc; i += nt) {
      unsigned cd = a.codes[l][i];
      if (i < pl) cd &= ~(unsigned)B_ZM;
      if (i >= nc - pl) cd &= ~(unsigned)B_ZP;
      C[l][i] = (uint8_t)cd;
    }
  }
  for (int i = tid; i < (int)a.g[0].ncell; i += nt) B[0][i] = a.b0[i];
  __syncthreads();

  // `sweeps` damped-Jacobi sweeps on level l; from zero when `first`; weights in
  // ascending order, or descending (`rev`) for the post-smoother.  Result in U[l].
  auto smooth = [&](int l, int sweeps, bool first, bool rev) {
    const Geo& g = a.g[l];
    const T inv_scale = T(1) / a.scale[l];
    for (int s = 0; s < sweeps; ++s) {
      const int k = s % a.nu;
      const T omega = a.omega[rev ? a.nu - 1 - k : k];
      const T* in = U[l];
      T* out = V[l];
      for (int i = tid; i < (int)g.ncell; i += nt) {
        const unsigned c = C[l][i];
        const int n = __popc(c & 63u);
        T o = T(0);
        if ((c & B_FLUID) && n > 0) {
          const T w = omega / T(n);
          if (first && s == 0) {
            o = w * (B[l][i] * inv_scale);
          } else {
            const T u0 = in[i];
            o = u0 + w * (B[l][i] * inv_scale - mg_lap_s<T>(g, c, i, in, u0));
          }
        }
        out[i] = o;
      }
      __syncthreads();
      T* t = U[l]; U[l] = V[l]; V[l] = t;
    }
  };

  for (int l = 0; l + 1 < a.nlev; ++l) {
    smooth(l, a.nu, true, false);
    const Geo& gf = a.g[l];
    const Geo& gc = a.g[l + 1];
    for (int idx = tid; idx < (int)gc.ncell; idx += nt) {
      T acc = T(0);
      if (C[l + 1][idx] & B_FLUID) {
        const int i = idx % gc.nx, t = idx / gc.nx, j = t % gc.ny, k = t / gc.ny;
        for (int ch = 0; ch < 8; ++ch) {
          const int fi = 2 * i + (ch & 1), fj = 2 * j + ((ch >> 1) & 1), fk = 2 * k + (ch >> 2);
          if (fi < gf.nx && fj < gf.ny && fk < gf.nz) {
            const int f = (fk * gf.ny + fj) * gf.nx + fi;
            const unsigned cf = C[l][f];
            if (cf & B_FLUID) acc += B[l][f] - a.scale[l] * mg_lap_s<T>(gf, cf, f, U[l], U[l][f]);
          }
        }
        acc *= T(mg_restrict_weight(gf));
      }
      B[l + 1][idx] = acc;
    }
    __syncthreads();
  }
  smooth(a.nlev - 1, a.coarse_sweeps, true, false);
  for (int l = a.nlev - 2; l >= 0; --l) {
    const Geo& gf = a.g[l];
    const Geo& gc = a.g[l + 1];
    for (int idx = tid; idx < (int)gf.ncell; idx += nt) {
      if (C[l][idx] & B_FLUID) {
        const int i = idx % gf.nx, t = idx / gf.nx, j = t % gf.ny, k = t / gf.ny;
        U[l][idx] += U[l + 1][((k >> 1) * gc.ny + (j >> 1)) * gc.nx + (i >> 1)];
      }
    }
    __syncthreads();
    smooth(l, a.nu, false, true);
  }
  for (int i = tid; i < (int)a.g[0].ncell; i += nt) a.u0[i] = U[0][i];
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
