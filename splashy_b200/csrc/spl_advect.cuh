// spl_advect.cuh -- semi-Lagrangian advection of the staggered MAC velocity
// (sm_100a).  Replaces Convection.apply + Grid.update_velocities
// (Convection.fs:50-66, Simulator.fs:82).
//
// Intended semantics: for every Fluid cell and each of its three
// +faces, RK2-midpoint backtrace from the face centre through the staggered
// velocity field, then trilinear sample of that component at the departure
// point.  Corners outside the box contribute 0 (Convection.fs:36-38 "None ->
// 0").  Jacobi update: reads (u,v,w), writes (u2,v2,w2).
//
// Positions are kept as (integer cell, small real offset) so fp32 does not
// lose the fractional part on 512^3 / 4096^2 grids.
#pragma once
#include "spl_common.cuh"

namespace spl {

template <typename T>
struct FieldView {
  const T* __restrict__ f;
  int nx, ny;
  int klo, khi;       // local planes that exist in the global box  [klo, khi)
  int hlo, hhi;       // local planes that exist in memory          [hlo, hhi)
  long long plane;
};

template <typename T>
__device__ __forceinline__ T fetch(const FieldView<T>& F, int i, int j, int k,
                                   unsigned& escaped) {
  if ((unsigned)i >= (unsigned)F.nx || (unsigned)j >= (unsigned)F.ny) return T(0);
  if (k < F.klo || k >= F.khi) return T(0);
  if (k < F.hlo || k >= F.hhi) { escaped = 1u; return T(0); }
  return F.f[(long long)k * F.plane + (long long)j * F.nx + i];
}

// Convection.fs:21-39 with the intended corner key: weights [1-f, f] per axis,
// sum nested x', y', z'.  Sample point = (bi + ox, bj + oy, bk + oz).
template <typename T>
__device__ __forceinline__ T sample(const FieldView<T>& F, int bi, int bj, int bk,
                                    T ox, T oy, T oz, unsigned& escaped) {
  const T fx = floor(ox), fy = floor(oy), fz = floor(oz);
  const T wx1 = ox - fx, wy1 = oy - fy, wz1 = oz - fz;
  const T wx[2] = {T(1) - wx1, wx1};
  const T wy[2] = {T(1) - wy1, wy1};
  const T wz[2] = {T(1) - wz1, wz1};
  const int i0 = bi + (int)fx, j0 = bj + (int)fy, k0 = bk + (int)fz;
  T total = T(0);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c)
        total += fetch(F, i0 + a, j0 + b, k0 + c, escaped) * wx[a] * wy[b] * wz[c];
  return total;
}

// ---- two-pass layout ------------------------------------------------------------
// k_advect (hot): cells whose whole sampling footprint is provably inside the
// allocation -- at least ADV_MARGIN cells from the x/y box edges, every RK2 offset
// within ADV_MAXOFF cells and its z part inside the memory-resident planes (halo
// planes at the global ends are zero = "absent cell contributes 0") -- are updated with unchecked 8-corner gathers at small
// immediate offsets from the cell's own address.  Every other Fluid cell is
// appended to a worklist (warp-aggregated atomic) and finished by k_advect_rest
// with the fully bounds-tested sampler (box shell, > 3-cell displacements, slab
// ends), one dense thread per listed cell.  The hot kernel is
// instruction-issue / L1-gather bound, not HBM bound: it is
// written for instruction count (32-bit offsets, lerp form, no calls, no stack).
// x / y cells from the box edge: with every offset (incl. the +-1/2 staggering shifts) inside
// [-2.99, 2.99], floor() is in [-3, 2] and the corners span [-3, +3]
constexpr int ADV_MARGIN = 3;
constexpr float ADV_MAXOFF = 2.99f;

template <typename T>
__device__ __forceinline__ T mix(T a, T b, T w) {   // (1 - w) a + w b
  return fma(w, b, fma(-w, a, a));
}

// unchecked trilinear sample at (cell + offset); p0 = address of the cell's own
// value, nx / pl = element strides (pl < 2^31 is guaranteed by the host)
// floor + float->int of a small offset.  -DSPL_ADV_MAGIC=1 avoids the conversion pipe (FRND /
// F2I issue at 16 lanes/clk/SM, the hot kernel needs 54 per cell) by adding 1.5 * 2^23, which
// leaves round-to-nearest(o - 1/2) in the low mantissa bits; measured SLOWER at 512^3 (4.79 vs
// 4.68 ms: 784 instead of 744 instructions per cell -- the kernel is issue-bound, the
// conversion pipe is not the limiter), so it is off.
#ifndef SPL_ADV_MAGIC
#define SPL_ADV_MAGIC 0
#endif
__device__ __forceinline__ void floor_split(float o, float& f, int& i) {
#if SPL_ADV_MAGIC
  const float t = (o - 0.5f) + 12582912.0f;
  i = __float_as_int(t) - 0x4B400000;
  f = t - 12582912.0f;
#else
  f = floorf(o);
  i = (int)f;
#endif
}
__device__ __forceinline__ void floor_split(double o, double& f, int& i) {
  f = floor(o);
  i = (int)f;
}

template <typename T>
__device__ __forceinline__ T sample_fast(const T* __restrict__ p0, int nx, int pl, T ox, T oy,
                                         T oz) {
  T fx, fy, fz;
  int ix, iy, iz;
  floor_split(ox, fx, ix);
  floor_split(oy, fy, iy);
  floor_split(oz, fz, iz);
  const T wx = ox - fx, wy = oy - fy, wz = oz - fz;
  const int off = (iz * pl + iy * nx) + ix;
  const T* __restrict__ pa = p0 + off;
  const T* __restrict__ pb = pa + pl;
  const T A = mix(mix(pa[0], pa[1], wx), mix(pa[nx], pa[nx + 1], wx), wy);
  const T B = mix(mix(pb[0], pb[1], wx), mix(pb[nx], pb[nx + 1], wx), wy);
  return mix(A, B, wz);
}

// |offset| <= lim in x / y and the z corners [floor(oz + zpad_lo), floor(oz + zpad_hi) + 1]
// inside the planes present in memory
template <typename T>
__device__ __forceinline__ bool within(T a, T b, T c, T lim, T zlo, T zhi, T zpad) {
  return (fmax(fabs(a), fabs(b)) <= lim) && (c - zpad >= zlo) && (c + zpad < zhi);
}

// 2-D thread blocks (32 x 8 cells of one plane) keep the gather footprint of a
// CTA compact in L1; grid = (x tiles, y tiles, planes).
template <typename T>
__global__ void __launch_bounds__(256)
k_advect(Geo g, const uint8_t* __restrict__ media, const T* __restrict__ u,
         const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ u2,
         T* __restrict__ v2, T* __restrict__ w2, T c, int* __restrict__ worklist,
         Scal* __restrict__ sc) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const bool xy_in = (i >= ADV_MARGIN) && (i < g.nx - ADV_MARGIN) && (j >= ADV_MARGIN) &&
                     (j < g.ny - ADV_MARGIN);
  const int nx = g.nx, pl = (int)g.plane;
  const T h = T(0.5), qd = T(0.25);
  const T lim1 = T(ADV_MAXOFF) - h;     // midpoint offsets get +-1/2 added
  const T lim2 = T(ADV_MAXOFF);
  for (int k = blockIdx.z; k < g.nz; k += gridDim.z) {
    const long long idx = ((long long)k * g.ny + j) * nx + i;
    const T* __restrict__ pu = u + idx;
    const T* __restrict__ pv = v + idx;
    const T* __restrict__ pw = w + idx;
    T nu = pu[0], nv = pv[0], nw = pw[0];
    if (media[idx] == FLUID) {
      // planes reachable from k without leaving the allocation: a sample at z
      // offset oz touches planes k + floor(oz) and k + floor(oz) + 1
      const T zlo = -(T)(k + g.hz), zhi = (T)(g.nz + g.hz - 1 - k);
      bool ok = xy_in;
      if (ok) {
        const T u00 = nu, v00 = nv, w00 = nw;
        // ---- U face at (i + 1/2, j, k): stage 0 = closed-form trilinear weights
        T v0 = qd * ((pv[-nx] + v00) + (pv[-nx + 1] + pv[1]));
        T w0 = qd * ((pw[-pl] + w00) + (pw[-pl + 1] + pw[1]));
        T mx = h * c * u00, my = h * c * v0, mz = h * c * w0;
        ok = ok && within(mx, my, mz, lim1, zlo, zhi, h);
        if (ok) {
          const T u1 = sample_fast(pu, nx, pl, mx, my, mz);
          const T v1 = sample_fast(pv, nx, pl, mx + h, my - h, mz);
          const T w1 = sample_fast(pw, nx, pl, mx + h, my, mz - h);
          const T dx = c * u1, dy = c * v1, dz = c * w1;
          ok = within(dx, dy, dz, lim2, zlo, zhi, T(0));
          if (ok) nu = sample_fast(pu, nx, pl, dx, dy, dz);
        }
        // ---- V face at (i, j + 1/2, k)
        if (ok) {
          const T u0 = qd * ((pu[-1] + pu[nx - 1]) + (u00 + pu[nx]));
          w0 = qd * ((pw[-pl] + w00) + (pw[-pl + nx] + pw[nx]));
          mx = h * c * u0; my = h * c * v00; mz = h * c * w0;
          ok = within(mx, my, mz, lim1, zlo, zhi, h);
        }
        if (ok) {
          const T u1 = sample_fast(pu, nx, pl, mx - h, my + h, mz);
          const T v1 = sample_fast(pv, nx, pl, mx, my, mz);
          const T w1 = sample_fast(pw, nx, pl, mx, my + h, mz - h);
          const T dx = c * u1, dy = c * v1, dz = c * w1;
          ok = within(dx, dy, dz, lim2, zlo, zhi, T(0));
          if (ok) nv = sample_fast(pv, nx, pl, dx, dy, dz);
        }
        // ---- W face at (i, j, k + 1/2)
        if (ok) {
          const T u0 = qd * ((pu[-1] + pu[pl - 1]) + (u00 + pu[pl]));
          v0 = qd * ((pv[-nx] + pv[pl - nx]) + (v00 + pv[pl]));
          mx = h * c * u0; my = h * c * v0; mz = h * c * w00;
          ok = within(mx, my, mz, lim1, zlo, zhi, h);
        }
        if (ok) {
          const T u1 = sample_fast(pu, nx, pl, mx - h, my, mz + h);
          const T v1 = sample_fast(pv, nx, pl, mx, my - h, mz + h);
          const T w1 = sample_fast(pw, nx, pl, mx, my, mz);
          const T dx = c * u1, dy = c * v1, dz = c * w1;
          ok = within(dx, dy, dz, lim2, zlo, zhi, T(0));
          if (ok) nw = sample_fast(pw, nx, pl, dx, dy, dz);
        }
      }
      if (!ok) worklist[atomicAdd(&sc->wl_count, 1u)] = (int)idx;   // -> k_advect_rest
    }
    u2[idx] = nu;
    v2[idx] = nv;
    w2[idx] = nw;
  }
}

// generic face update: every corner bounds-tested
template <typename T>
__device__ __forceinline__ T advect_face(const FieldView<T>& U, const FieldView<T>& V,
                                         const FieldView<T>& W, int comp, int i, int j,
                                         int k, T c /* sgn*dt/h */, unsigned& esc) {
  const T h = T(0.5);
  // offset of the face centre from the sample lattice of each component:
  // sampling component b at (face of a) + d uses offset d + e_a/2 - e_b/2
  const T ax = (comp == 0) ? h : T(0), ay = (comp == 1) ? h : T(0),
          az = (comp == 2) ? h : T(0);
  const T u0 = sample(U, i, j, k, ax - h, ay, az, esc);
  const T v0 = sample(V, i, j, k, ax, ay - h, az, esc);
  const T w0 = sample(W, i, j, k, ax, ay, az - h, esc);
  const T mx = h * c * u0, my = h * c * v0, mz = h * c * w0;
  const T u1 = sample(U, i, j, k, mx + ax - h, my + ay, mz + az, esc);
  const T v1 = sample(V, i, j, k, mx + ax, my + ay - h, mz + az, esc);
  const T w1 = sample(W, i, j, k, mx + ax, my + ay, mz + az - h, esc);
  const T dx = c * u1, dy = c * v1, dz = c * w1;
  const FieldView<T>& F = (comp == 0) ? U : (comp == 1) ? V : W;
  return sample(F, i, j, k, dx, dy, dz, esc);
}

// cell-centred scalar: the same RK2 trace from the cell centre (i, j, k), then a trilinear
// sample of the scalar itself at the departure point (corners outside the box contribute 0,
// like the velocity components: Convection.fs:36-38).  The reference's Cell carries no scalar
// (Cell.fs:11-16); this is north_star's "velocity and scalar fields" (SURVEY 8d cfg2 "+8 B").
template <typename T>
__device__ __forceinline__ T advect_centre(const FieldView<T>& U, const FieldView<T>& V,
                                           const FieldView<T>& W, const FieldView<T>& S, int i,
                                           int j, int k, T c /* sgn*dt/h */, unsigned& esc) {
  const T h = T(0.5), o = T(0);
  const T u0 = sample(U, i, j, k, -h, o, o, esc);
  const T v0 = sample(V, i, j, k, o, -h, o, esc);
  const T w0 = sample(W, i, j, k, o, o, -h, esc);
  const T mx = h * c * u0, my = h * c * v0, mz = h * c * w0;
  const T u1 = sample(U, i, j, k, mx - h, my, mz, esc);
  const T v1 = sample(V, i, j, k, mx, my - h, mz, esc);
  const T w1 = sample(W, i, j, k, mx, my, mz - h, esc);
  return sample(S, i, j, k, c * u1, c * v1, c * w1, esc);
}

// second pass: the cells the hot kernel listed (box shell, slab ends, large
// displacements), one thread per listed cell.  s / s2 (nullable): a scalar field advected
// with the same velocity; do_vel = 0 leaves u2, v2, w2 alone (scalar-only pass).
template <typename T>
__global__ void __launch_bounds__(256, 2)
k_advect_rest(Geo g, const T* __restrict__ u, const T* __restrict__ v,
              const T* __restrict__ w, T* __restrict__ u2, T* __restrict__ v2,
              T* __restrict__ w2, const T* __restrict__ s, T* __restrict__ s2, int do_vel, T c,
              const int* __restrict__ worklist, Scal* __restrict__ sc) {
  FieldView<T> U{u, g.nx, g.ny, -g.z0, g.nzg - g.z0, -g.hz, g.nz + g.hz, g.plane};
  FieldView<T> V = U, W = U, S = U;
  V.f = v;
  W.f = w;
  S.f = s;
  unsigned esc = 0;
  const unsigned n = sc->wl_count;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int idx = worklist[t];
    const int i = idx % g.nx;
    const int r = idx / g.nx;
    const int j = r % g.ny;
    const int k = r / g.ny;
    if (do_vel) {
      u2[idx] = advect_face<T>(U, V, W, 0, i, j, k, c, esc);
      v2[idx] = advect_face<T>(U, V, W, 1, i, j, k, c, esc);
      w2[idx] = advect_face<T>(U, V, W, 2, i, j, k, c, esc);
    }
    if (s) s2[idx] = advect_centre<T>(U, V, W, S, i, j, k, c, esc);
  }
  if (esc) atomicAdd(&sc->escapes, 1ull);
}

// generic scalar pass (fp64, ragged nx, quirk contexts): every corner bounds-tested
template <typename T>
__global__ void __launch_bounds__(256, 2)
k_advect_scalar(Geo g, const uint8_t* __restrict__ media, const T* __restrict__ u,
                const T* __restrict__ v, const T* __restrict__ w, const T* __restrict__ s,
                T* __restrict__ s2, T c, Scal* __restrict__ sc) {
  FieldView<T> U{u, g.nx, g.ny, -g.z0, g.nzg - g.z0, -g.hz, g.nz + g.hz, g.plane};
  FieldView<T> V = U, W = U, S = U;
  V.f = v;
  W.f = w;
  S.f = s;
  unsigned esc = 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    T val = s[idx];
    if (media[idx] == FLUID) {
      const int i = (int)(idx % g.nx);
      const long long t = idx / g.nx;
      val = advect_centre<T>(U, V, W, S, i, (int)(t % g.ny), (int)(t / g.ny), c, esc);
    }
    s2[idx] = val;
  }
  if (esc) atomicAdd(&sc->escapes, 1ull);
}

// ---- literal reference behaviour (quirk flags) ----------------------
// Coord.fs:49-57 with .NET remainder semantics
__device__ __forceinline__ int ref_nearest(int n) {
  const int h = 10;
  const int r = n % h;          // C++ % truncates like .NET
  if (r == 0) return n;
  if (r >= h / 2) return n + (h - r);
  return n - r;
}

struct QuirkCfg {
  unsigned quirks;
  int ox, oy, oz;     // world cell index of local cell (0,0,0) (oz includes z0)
  double h;
};

// One thread per Fluid cell, fp64 arithmetic regardless of the storage type.
// With all three flags set this is Convection.fs:21-66 verbatim except for the
// commit order (Jacobi instead of the reference's in-place Seq.iter).
template <typename T>
__device__ __forceinline__ double q_fetch(const QuirkCfg& q, const FieldView<T>& F,
                                          int wi, int wj, int wk, bool& present,
                                          const uint8_t* __restrict__ media,
                                          unsigned& esc) {
  // world cell index -> box index
  const int i = wi - q.ox, j = wj - q.oy, k = wk - q.oz;
  present = false;
  if ((unsigned)i >= (unsigned)F.nx || (unsigned)j >= (unsigned)F.ny) return 0.0;
  if (k < F.klo || k >= F.khi) return 0.0;
  if (k < F.hlo || k >= F.hhi) { esc = 1u; return 0.0; }
  const long long idx = (long long)k * F.plane + (long long)j * F.nx + i;
  if (media[idx] == ABSENT) return 0.0;
  present = true;
  return (double)F.f[idx];
}

template <typename T>
__device__ __forceinline__ double q_interp(const QuirkCfg& q, const FieldView<T>& F,
                                           const uint8_t* __restrict__ media, double x,
                                           double y, double z, unsigned& esc) {
  // Convection.fs:21-39; x, y, z in index units (metres / h)
  const double i = floor(x), j = floor(y), k = floor(z);
  const int ii = (int)i, jj = (int)j, kk = (int)k;
  const double id[2] = {i - x + 1.0, x - i};
  const double jd[2] = {j - y + 1.0, y - j};
  const double kd[2] = {k - z + 1.0, z - k};
  double total = 0.0;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < 2; ++c) {
        int wi, wj, wk;
        if (q.quirks & 1u) {
          // Q3: int overload of Coord.construct applied to *index* values, the
          // result read as metres: world cell index = nearest(n) / 10
          wi = ref_nearest(ii + a) / 10;
          wj = ref_nearest(jj + b) / 10;
          wk = ref_nearest(kk + c) / 10;
        } else {
          wi = ii + a; wj = jj + b; wk = kk + c;
        }
        bool present;
        const double val = q_fetch(q, F, wi, wj, wk, present, media, esc);
        total += present ? val * id[a] * jd[b] * kd[c] : 0.0;
      }
  return total;
}

template <typename T>
__device__ __forceinline__ void q_velocity(const QuirkCfg& q, const FieldView<T>& U,
                                           const FieldView<T>& V, const FieldView<T>& W,
                                           const uint8_t* __restrict__ media, double X,
                                           double Y, double Z, double out[3],
                                           unsigned& esc) {
  // Convection.fs:41-48 (metres in)
  const double xh = X / q.h, yh = Y / q.h, zh = Z / q.h;
  if (q.quirks & 1u) {
    out[0] = q_interp(q, U, media, xh, yh - 0.5, zh - 0.5, esc);
    out[1] = q_interp(q, V, media, xh - 0.5, yh, zh - 0.5, esc);
    out[2] = q_interp(q, W, media, xh - 0.5, yh - 0.5, zh, esc);
  } else {
    out[0] = q_interp(q, U, media, xh - 0.5, yh, zh, esc);
    out[1] = q_interp(q, V, media, xh, yh - 0.5, zh, esc);
    out[2] = q_interp(q, W, media, xh, yh, zh - 0.5, esc);
  }
}

__device__ __forceinline__ int ref_round_nearest(double x) {
  // Coord.fs:65-67: round (banker's) -> int -> nearest
  return ref_nearest((int)rint(x));
}

template <typename T>
__global__ void k_advect_quirk(Geo g, QuirkCfg q, const uint8_t* __restrict__ media,
                               const T* __restrict__ u, const T* __restrict__ v,
                               const T* __restrict__ w, T* __restrict__ u2,
                               T* __restrict__ v2, T* __restrict__ w2, double dt,
                               Scal* __restrict__ sc, int* __restrict__ too_far) {
  FieldView<T> U{u, g.nx, g.ny, -g.z0, g.nzg - g.z0, -g.hz, g.nz + g.hz, g.plane};
  FieldView<T> V = U, W = U;
  V.f = v;
  W.f = w;
  unsigned esc = 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    T nu = u[idx], nv = v[idx], nw = w[idx];
    if (media[idx] == FLUID) {
      const int i = (int)(idx % g.nx);
      const long long t = idx / g.nx;
      const int j = (int)(t % g.ny);
      const int k = (int)(t / g.ny);
      const double sgn = (q.quirks & 2u) ? 1.0 : -1.0;       // Q4
      const double cx = (i + q.ox) * q.h, cy = (j + q.oy) * q.h,
                   cz = (k + q.oz) * q.h;
      if (q.quirks & 4u) {
        // Q5: one trace from the cell coordinate, whole-vector nearest copy
        double v0[3], v1[3];
        q_velocity(q, U, V, W, media, cx, cy, cz, v0, esc);
        q_velocity(q, U, V, W, media, cx + sgn * v0[0] * (dt / 2.0),
                   cy + sgn * v0[1] * (dt / 2.0), cz + sgn * v0[2] * (dt / 2.0), v1, esc);
        const int px = ref_round_nearest(cx + sgn * v1[0] * dt) / 10;
        const int py = ref_round_nearest(cy + sgn * v1[1] * dt) / 10;
        const int pz = ref_round_nearest(cz + sgn * v1[2] * dt) / 10;
        bool p0, p1, p2;
        const double a = q_fetch(q, U, px, py, pz, p0, media, esc);
        const double b = q_fetch(q, V, px, py, pz, p1, media, esc);
        const double c = q_fetch(q, W, px, py, pz, p2, media, esc);
        if (!p0) atomicExch(too_far, 1);   // Convection.fs:65
        nu = (T)a; nv = (T)b; nw = (T)c;
      } else {
        // per-face trace with the selected sampler / sign
        double res[3];
        for (int comp = 0; comp < 3; ++comp) {
          const double fx = cx + (comp == 0 ? 0.5 * q.h : 0.0);
          const double fy = cy + (comp == 1 ? 0.5 * q.h : 0.0);
          const double fz = cz + (comp == 2 ? 0.5 * q.h : 0.0);
          double v0[3], v1[3], v2d[3];
          q_velocity(q, U, V, W, media, fx, fy, fz, v0, esc);
          q_velocity(q, U, V, W, media, fx + sgn * v0[0] * (dt / 2.0),
                     fy + sgn * v0[1] * (dt / 2.0), fz + sgn * v0[2] * (dt / 2.0), v1, esc);
          q_velocity(q, U, V, W, media, fx + sgn * v1[0] * dt, fy + sgn * v1[1] * dt,
                     fz + sgn * v1[2] * dt, v2d, esc);
          res[comp] = v2d[comp];
        }
        nu = (T)res[0]; nv = (T)res[1]; nw = (T)res[2];
      }
    }
    u2[idx] = nu;
    v2[idx] = nv;
    w2[idx] = nw;
  }
  if (esc) atomicAdd(&sc->escapes, 1ull);
}

}  // namespace spl
