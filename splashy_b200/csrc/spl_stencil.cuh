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

// ---- 4-cells-per-thread variants (fp32, nx % 4 = 0) -----------------------------------
// Same arithmetic, cell by cell, as k_div_rhs / k_grad_sub / k_check; a thread owns 4
// consecutive x cells: 128-bit accesses for the own cells and the -y / -z (+y / +z)
// neighbours, one scalar for the x neighbour outside the quad, one div / mod per quad.
// Every neighbour address stays inside the allocation (z-halo planes exist), validity
// comes from the code bits / the labels.
__device__ __forceinline__ float quad_div(unsigned c, float u0, float um, float v0, float vm,
                                          float w0, float wm) {
  const float ox = (c & B_XP) ? u0 : 0.f;
  const float oy = (c & B_YP) ? v0 : 0.f;
  const float oz = (c & B_ZP) ? w0 : 0.f;
  const float ix = (c & B_XM) ? um : 0.f;
  const float iy = (c & B_YM) ? vm : 0.f;
  const float iz = (c & B_ZM) ? wm : 0.f;
  return ((ix - ox) + (iy - oy)) + (iz - oz);
}

__global__ void __launch_bounds__(256)
k_div_rhs4(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ u,
           const float* __restrict__ v, const float* __restrict__ w, float* __restrict__ out,
           float scale) {
  const long long nq = g.ncell >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const long long idx = q << 2;
    const unsigned c4 = *reinterpret_cast<const unsigned*>(codes + idx);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 & 0x40404040u) {
      const float4 u4 = *reinterpret_cast<const float4*>(u + idx);
      const float4 v4 = *reinterpret_cast<const float4*>(v + idx);
      const float4 w4 = *reinterpret_cast<const float4*>(w + idx);
      const float4 vm = *reinterpret_cast<const float4*>(v + idx - g.nx);
      const float4 wm = *reinterpret_cast<const float4*>(w + idx - g.plane);
      const float ul = (c4 & B_XM) ? u[idx - 1] : 0.f;
      const unsigned c0 = c4 & 0xffu, c1 = (c4 >> 8) & 0xffu, c2 = (c4 >> 16) & 0xffu, c3 = c4 >> 24;
      if (c0 & B_FLUID) o.x = scale * quad_div(c0, u4.x, ul, v4.x, vm.x, w4.x, wm.x);
      if (c1 & B_FLUID) o.y = scale * quad_div(c1, u4.y, u4.x, v4.y, vm.y, w4.y, wm.y);
      if (c2 & B_FLUID) o.z = scale * quad_div(c2, u4.z, u4.y, v4.z, vm.z, w4.z, wm.z);
      if (c3 & B_FLUID) o.w = scale * quad_div(c3, u4.w, u4.z, v4.w, vm.w, w4.w, wm.w);
    }
    *reinterpret_cast<float4*>(out + idx) = o;
  }
}

// one face: returns the new velocity of the face between a cell (m0, p0) and its
// + neighbour (m1, p1); see k_grad_sub
__device__ __forceinline__ float grad_face(uint8_t m0, uint8_t m1, double p0, double p1, float f,
                                           double kappa) {
  const bool ns0 = is_nonsolid(m0), fl0 = (m0 == FLUID);
  const bool ns1 = is_nonsolid(m1), fl1 = (m1 == FLUID);
  if (ns0 && ns1 && (fl0 || fl1))
    return (float)((double)f - kappa * ((fl1 ? p1 : 0.0) - (fl0 ? p0 : 0.0)));
  if ((fl0 && !ns1) || (!ns0 && fl1)) return 0.f;
  return f;
}

__global__ void __launch_bounds__(256)
k_grad_sub4(Geo g, const uint8_t* __restrict__ media, const double* __restrict__ p,
            float* __restrict__ u, float* __restrict__ v, float* __restrict__ w, double kappa) {
  const long long nq = g.ncell >> 2;
  const int qx = g.nx >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const long long idx = q << 2;
    const int i = (int)(q % qx) << 2;
    const long long t = q / qx;
    const int j = (int)(t % g.ny);
    const unsigned m4 = *reinterpret_cast<const unsigned*>(media + idx);
    const bool last_x = (i + 4 >= g.nx), last_y = (j == g.ny - 1);
    // labels of the + neighbours (ABSENT outside the box; the z-halo holds the
    // neighbouring slab's labels or ABSENT at the global end)
    const unsigned mxr = last_x ? (unsigned)ABSENT : (unsigned)media[idx + 4];
    const unsigned my4 = last_y ? 0xffffffffu : *reinterpret_cast<const unsigned*>(media + idx + g.nx);
    const unsigned mz4 = *reinterpret_cast<const unsigned*>(media + idx + g.plane);
    // any Fluid cell among the quad and its + neighbours?  otherwise nothing changes
    auto has_fluid = [](unsigned x) {
      const unsigned y = x ^ 0x01010101u;          // a byte is 0 where the label is FLUID (1)
      return ((y - 0x01010101u) & ~y & 0x80808080u) != 0u;
    };
    if (!(has_fluid(m4) || has_fluid(my4) || has_fluid(mz4) || mxr == (unsigned)FLUID)) continue;
    const uint8_t m[5] = {(uint8_t)(m4 & 0xffu), (uint8_t)((m4 >> 8) & 0xffu),
                          (uint8_t)((m4 >> 16) & 0xffu), (uint8_t)(m4 >> 24), (uint8_t)mxr};
    const double2 pa = *reinterpret_cast<const double2*>(p + idx);
    const double2 pb = *reinterpret_cast<const double2*>(p + idx + 2);
    const double px[5] = {pa.x, pa.y, pb.x, pb.y, (mxr == (unsigned)FLUID) ? p[idx + 4] : 0.0};
    double py[4] = {0.0, 0.0, 0.0, 0.0}, pz[4];
    if (!last_y) {
      const double2 a = *reinterpret_cast<const double2*>(p + idx + g.nx);
      const double2 b = *reinterpret_cast<const double2*>(p + idx + g.nx + 2);
      py[0] = a.x; py[1] = a.y; py[2] = b.x; py[3] = b.y;
    }
    {
      const double2 a = *reinterpret_cast<const double2*>(p + idx + g.plane);
      const double2 b = *reinterpret_cast<const double2*>(p + idx + g.plane + 2);
      pz[0] = a.x; pz[1] = a.y; pz[2] = b.x; pz[3] = b.y;
    }
    float4 u4 = *reinterpret_cast<const float4*>(u + idx);
    float4 v4 = *reinterpret_cast<const float4*>(v + idx);
    float4 w4 = *reinterpret_cast<const float4*>(w + idx);
    float* uu = reinterpret_cast<float*>(&u4);
    float* vv = reinterpret_cast<float*>(&v4);
    float* ww = reinterpret_cast<float*>(&w4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uu[e] = grad_face(m[e], m[e + 1], px[e], px[e + 1], uu[e], kappa);
      vv[e] = grad_face(m[e], (uint8_t)((my4 >> (8 * e)) & 0xffu), px[e], py[e], vv[e], kappa);
      ww[e] = grad_face(m[e], (uint8_t)((mz4 >> (8 * e)) & 0xffu), px[e], pz[e], ww[e], kappa);
    }
    *reinterpret_cast<float4*>(u + idx) = u4;
    *reinterpret_cast<float4*>(v + idx) = v4;
    *reinterpret_cast<float4*>(w + idx) = w4;
  }
}

__global__ void __launch_bounds__(256)
k_check4(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ u,
         const float* __restrict__ v, const float* __restrict__ w, const double* __restrict__ p,
         Scal* __restrict__ sc, double* __restrict__ partials) {
  __shared__ double red[32];
  float mx = 0.f;
  unsigned bad = 0;
  const long long nq = g.ncell >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const long long idx = q << 2;
    const unsigned c4 = *reinterpret_cast<const unsigned*>(codes + idx);
    const float4 u4 = *reinterpret_cast<const float4*>(u + idx);
    const float4 v4 = *reinterpret_cast<const float4*>(v + idx);
    const float4 w4 = *reinterpret_cast<const float4*>(w + idx);
    bad += !isfinite(u4.x) + !isfinite(u4.y) + !isfinite(u4.z) + !isfinite(u4.w);
    bad += !isfinite(v4.x) + !isfinite(v4.y) + !isfinite(v4.z) + !isfinite(v4.w);
    bad += !isfinite(w4.x) + !isfinite(w4.y) + !isfinite(w4.z) + !isfinite(w4.w);
    if (c4 & 0x40404040u) {
      const float4 vm = *reinterpret_cast<const float4*>(v + idx - g.nx);
      const float4 wm = *reinterpret_cast<const float4*>(w + idx - g.plane);
      const float ul = (c4 & B_XM) ? u[idx - 1] : 0.f;
      const double2 pa = *reinterpret_cast<const double2*>(p + idx);
      const double2 pb = *reinterpret_cast<const double2*>(p + idx + 2);
      const unsigned c0 = c4 & 0xffu, c1 = (c4 >> 8) & 0xffu, c2 = (c4 >> 16) & 0xffu, c3 = c4 >> 24;
      float d;
      if (c0 & B_FLUID) {
        bad += !isfinite(pa.x);
        d = fabsf(quad_div(c0, u4.x, ul, v4.x, vm.x, w4.x, wm.x));
        if (d == d) mx = fmaxf(mx, d);
      }
      if (c1 & B_FLUID) {
        bad += !isfinite(pa.y);
        d = fabsf(quad_div(c1, u4.y, u4.x, v4.y, vm.y, w4.y, wm.y));
        if (d == d) mx = fmaxf(mx, d);
      }
      if (c2 & B_FLUID) {
        bad += !isfinite(pb.x);
        d = fabsf(quad_div(c2, u4.z, u4.y, v4.z, vm.z, w4.z, wm.z));
        if (d == d) mx = fmaxf(mx, d);
      }
      if (c3 & B_FLUID) {
        bad += !isfinite(pb.y);
        d = fabsf(quad_div(c3, u4.w, u4.z, v4.w, vm.w, w4.w, wm.w));
        if (d == d) mx = fmaxf(mx, d);
      }
    }
  }
  const double bm = block_max((double)mx, red);
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

// ---- gradient subtraction + validators in one pass (fp32, nx % 4 = 0) -------------------
// Pressure.apply followed by the checks (Pressure.fs:96-149): the same arithmetic, cell by
// cell, as k_grad_sub4 then k_check4, with every array pulled from HBM once:
//   labels 1 + codes 1 + p 8 + u, v, w 12 in, 12 out = 34 B / cell   (33 + 21 as two passes)
// A CTA owns a 128 x GC_TY tile of the x-y plane (a thread = one aligned quad) and marches
// along z.  The pressure and labels of plane k+1 (the +z neighbours) are this plane's loads
// and the next plane's own values, carried in registers, and so is the new w of plane k-1
// (the incoming z face of the divergence).  The +-y neighbours' p and labels go through a
// double-buffered shared-memory tile (one barrier per plane); the incoming y face v'(j-1) is
// recomputed from them and the old v of row j-1 (an L1 / L2 hit: the row above loads the same
// line), the incoming x face u'(i-1) comes from the lane on the left.  Only the tile-edge rows /
// lanes read their halo from global memory.
// Out of place (u, v, w -> u2, v2, w2), so a halo is always an OLD value whichever CTA runs first.
// Packed label logic of grad_face for four cells at once (one label per byte, any byte value):
// bit 0 of byte e of `ns` = "label e is Air or Fluid" ((label & 0xfe) == 0), of `fl` = "Fluid".
__device__ __forceinline__ unsigned gc_ns4(unsigned m) {
  const unsigned t = m & 0xfefefefeu;
  return (~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u) >> 7;
}
struct GcFace {
  unsigned both, zero;   // bit 8e: face e takes the pressure gradient / the solid velocity 0
};
__device__ __forceinline__ GcFace gc_face4(unsigned ns0, unsigned fl0, unsigned ns1, unsigned fl1) {
  GcFace f;
  f.both = ns0 & ns1 & (fl0 | fl1);
  f.zero = (fl0 & ~ns1) | (~ns0 & fl1);
  return f;
}
// p0, p1 are already 0 on non-Fluid cells (sanitised when they enter the shared tile), so the
// value is grad_face's (double)f - kappa ((fl1 ? p1 : 0) - (fl0 ? p0 : 0)), computed
// unconditionally and selected by the masks: no branches.
__device__ __forceinline__ float gc_apply(const GcFace& fm, int e, double p0, double p1, float f,
                                          double kappa) {
  const float val = (float)((double)f - kappa * (p1 - p0));
  const float keep = ((fm.zero >> (8 * e)) & 1u) ? 0.f : f;
  return ((fm.both >> (8 * e)) & 1u) ? val : keep;
}
__device__ __forceinline__ double2 gc_sel2(unsigned fl, int e, double2 v) {
  return make_double2(((fl >> (8 * e)) & 1u) ? v.x : 0.0, ((fl >> (8 * e + 8)) & 1u) ? v.y : 0.0);
}

constexpr int GC_TY = 8;
struct GcSmem {
  double p[2][GC_TY + 2][128];      // rows j0-1 .. j0+GC_TY
  unsigned m[2][GC_TY + 2][32];     // labels, one quad per word
};

__global__ void __launch_bounds__(32 * GC_TY, 3)
k_grad_check4(Geo g, const uint8_t* __restrict__ media, const uint8_t* __restrict__ codes,
              const double* __restrict__ p, const float* __restrict__ u,
              const float* __restrict__ v, const float* __restrict__ w, float* __restrict__ u2,
              float* __restrict__ v2, float* __restrict__ w2, double kappa, Scal* __restrict__ sc,
              double* __restrict__ partials, int zchunk) {
  __shared__ GcSmem sm;
  __shared__ double red[32];
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 128 + 4 * lane, j = blockIdx.y * GC_TY + ty;
  const int kb = blockIdx.z * zchunk, ke = min(kb + zchunk, g.nz);
  const bool act = i < g.nx && j < g.ny;
  const bool has_xl = act && i > 0, has_xr = act && i + 4 < g.nx;
  const bool row_lo = (ty == 0), row_hi = (ty == GC_TY - 1);
  const bool lo_in = i < g.nx && j >= 1 && j - 1 < g.ny;        // row j-1 is a row of the box
  const bool hi_in = i < g.nx && j + 1 < g.ny;
  const unsigned ABS4 = 0xffffffffu;                            // four ABSENT labels
  const double2 Z2 = make_double2(0.0, 0.0);
  const float4 Z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  long long idx = ((long long)kb * g.ny + j) * g.nx + i;
  float mx = 0.f;
  unsigned bad = 0;

  // The shared tile of plane k+1 is written during iteration k (own row: this iteration's +z
  // loads; halo rows: loaded by the CTA's edge rows one plane ahead), so the barrier at the top
  // of an iteration never waits for a load issued in that iteration, and nothing but the new w
  // of plane k-1 is carried in registers.  Pressures enter the tile as 0 on non-Fluid cells.
  const bool edge = row_lo || row_hi;                 // GC_TY > 1: never both
  const bool edge_in = row_lo ? lo_in : hi_in;
  const long long eoff = row_lo ? -(long long)g.nx : (long long)g.nx;
  const int erow = row_lo ? 0 : GC_TY + 1;
  float wprev[4] = {0.f, 0.f, 0.f, 0.f};
  {
    unsigned m_c = ABS4;
    double2 pc0 = Z2, pc1 = Z2;
    if (act) {
      m_c = *reinterpret_cast<const unsigned*>(media + idx);
      pc0 = *reinterpret_cast<const double2*>(p + idx);
      pc1 = *reinterpret_cast<const double2*>(p + idx + 2);
      const unsigned mb = *reinterpret_cast<const unsigned*>(media + idx - g.plane);
      double2 b0 = *reinterpret_cast<const double2*>(p + idx - g.plane);
      double2 b1 = *reinterpret_cast<const double2*>(p + idx - g.plane + 2);
      const float4 wb = *reinterpret_cast<const float4*>(w + idx - g.plane);
      const unsigned nsb = gc_ns4(mb), flb = mb & nsb, nsc = gc_ns4(m_c), flc = m_c & nsc;
      b0 = gc_sel2(flb, 0, b0); b1 = gc_sel2(flb, 2, b1);
      pc0 = gc_sel2(flc, 0, pc0); pc1 = gc_sel2(flc, 2, pc1);
      const GcFace fz = gc_face4(nsb, flb, nsc, flc);
      wprev[0] = gc_apply(fz, 0, b0.x, pc0.x, wb.x, kappa);
      wprev[1] = gc_apply(fz, 1, b0.y, pc0.y, wb.y, kappa);
      wprev[2] = gc_apply(fz, 2, b1.x, pc1.x, wb.z, kappa);
      wprev[3] = gc_apply(fz, 3, b1.y, pc1.y, wb.w, kappa);
    }
    sm.m[0][ty + 1][lane] = m_c;
    *reinterpret_cast<double2*>(&sm.p[0][ty + 1][4 * lane]) = pc0;
    *reinterpret_cast<double2*>(&sm.p[0][ty + 1][4 * lane + 2]) = pc1;
    if (edge) {
      unsigned mm = ABS4;
      double2 a0 = Z2, a1 = Z2;
      if (edge_in) {
        mm = *reinterpret_cast<const unsigned*>(media + idx + eoff);
        a0 = *reinterpret_cast<const double2*>(p + idx + eoff);
        a1 = *reinterpret_cast<const double2*>(p + idx + eoff + 2);
        const unsigned f = mm & gc_ns4(mm);
        a0 = gc_sel2(f, 0, a0); a1 = gc_sel2(f, 2, a1);
      }
      sm.m[0][erow][lane] = mm;
      *reinterpret_cast<double2*>(&sm.p[0][erow][4 * lane]) = a0;
      *reinterpret_cast<double2*>(&sm.p[0][erow][4 * lane + 2]) = a1;
    }
  }

  for (int k = kb; k < ke; ++k, idx += g.plane) {
    const int b = (k - kb) & 1;
    // ---- this plane's loads (none is needed before the barrier) -----------------------------------
    unsigned c4 = 0, m_n = ABS4, mxl = ABSENT, mxr = ABSENT, hm = ABS4;
    float4 u4 = Z4, v4 = Z4, w4 = Z4, vq = Z4;
    double2 pn0 = Z2, pn1 = Z2, h0 = Z2, h1 = Z2;
    double pxl = 0.0, pxr = 0.0;
    float uxl = 0.f;
    if (act) {
      c4 = *reinterpret_cast<const unsigned*>(codes + idx);
      u4 = *reinterpret_cast<const float4*>(u + idx);
      v4 = *reinterpret_cast<const float4*>(v + idx);
      w4 = *reinterpret_cast<const float4*>(w + idx);
      m_n = *reinterpret_cast<const unsigned*>(media + idx + g.plane);
      pn0 = *reinterpret_cast<const double2*>(p + idx + g.plane);
      pn1 = *reinterpret_cast<const double2*>(p + idx + g.plane + 2);
      if (j > 0) vq = *reinterpret_cast<const float4*>(v + idx - g.nx);   // old v of row j-1
      if (lane == 0 && has_xl) { mxl = media[idx - 1]; pxl = p[idx - 1]; uxl = u[idx - 1]; }
      if (lane == 31 && has_xr) { mxr = media[idx + 4]; pxr = p[idx + 4]; }
    }
    if (edge && edge_in) {                            // halo row of plane k+1
      hm = *reinterpret_cast<const unsigned*>(media + idx + g.plane + eoff);
      h0 = *reinterpret_cast<const double2*>(p + idx + g.plane + eoff);
      h1 = *reinterpret_cast<const double2*>(p + idx + g.plane + eoff + 2);
    }
    __syncthreads();
    // ---- label masks of the quad and of its +x, +y, +z, -y neighbours -----------------------------
    const unsigned m_c = sm.m[b][ty + 1][lane];
    if (lane < 31) {
      mxr = sm.m[b][ty + 1][lane + 1] & 0xffu;
      pxr = sm.p[b][ty + 1][4 * lane + 4];
    } else {
      pxr = (mxr == (unsigned)FLUID) ? pxr : 0.0;
    }
    const unsigned ns0 = gc_ns4(m_c), fl0 = m_c & ns0;
    const unsigned m_x = (m_c >> 8) | (mxr << 24);            // labels of cells 1 .. 4
    const unsigned nsx = gc_ns4(m_x), flx = m_x & nsx;
    const unsigned nsz = gc_ns4(m_n), flz = m_n & nsz;
    const GcFace fx = gc_face4(ns0, fl0, nsx, flx), fzf = gc_face4(ns0, fl0, nsz, flz);
    pn0 = gc_sel2(flz, 0, pn0); pn1 = gc_sel2(flz, 2, pn1);
    double px[5];
    {
      const double2 a0 = *reinterpret_cast<const double2*>(&sm.p[b][ty + 1][4 * lane]);
      const double2 a1 = *reinterpret_cast<const double2*>(&sm.p[b][ty + 1][4 * lane + 2]);
      px[0] = a0.x; px[1] = a0.y; px[2] = a1.x; px[3] = a1.y; px[4] = pxr;
    }
    float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float ww[4] = {w4.x, w4.y, w4.z, w4.w}, vm[4];
    const unsigned cc[4] = {c4 & 0xffu, (c4 >> 8) & 0xffu, (c4 >> 16) & 0xffu, c4 >> 24};
    {
      const double pz[4] = {pn0.x, pn0.y, pn1.x, pn1.y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uu[e] = gc_apply(fx, e, px[e], px[e + 1], uu[e], kappa);
        ww[e] = gc_apply(fzf, e, px[e], pz[e], ww[e], kappa);
      }
    }
    {
      const unsigned my4 = sm.m[b][ty + 2][lane];
      const unsigned nsy = gc_ns4(my4), fly = my4 & nsy;
      const GcFace fy = gc_face4(ns0, fl0, nsy, fly);
      const double2 y0 = *reinterpret_cast<const double2*>(&sm.p[b][ty + 2][4 * lane]);
      const double2 y1 = *reinterpret_cast<const double2*>(&sm.p[b][ty + 2][4 * lane + 2]);
      const double py[4] = {y0.x, y0.y, y1.x, y1.y};
#pragma unroll
      for (int e = 0; e < 4; ++e) vv[e] = gc_apply(fy, e, px[e], py[e], vv[e], kappa);
    }
    {   // incoming y face: the new v of row j-1 (quad_div reads it only under B_YM)
      const unsigned mq4 = sm.m[b][ty][lane];
      const unsigned nsq = gc_ns4(mq4), flq = mq4 & nsq;
      const GcFace fq = gc_face4(nsq, flq, ns0, fl0);
      const double2 q0 = *reinterpret_cast<const double2*>(&sm.p[b][ty][4 * lane]);
      const double2 q1 = *reinterpret_cast<const double2*>(&sm.p[b][ty][4 * lane + 2]);
      const double pq[4] = {q0.x, q0.y, q1.x, q1.y};
      const float vqa[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) vm[e] = gc_apply(fq, e, pq[e], px[e], vqa[e], kappa);
    }
    // incoming x face of cell 0: the new u of the cell on the left
    float ul = __shfl_up_sync(0xffffffffu, uu[3], 1);
    if (lane == 0)
      ul = has_xl ? grad_face((uint8_t)mxl, (uint8_t)(m_c & 0xffu), pxl, px[0], uxl, kappa) : 0.f;
    if (act) {
      *reinterpret_cast<float4*>(u2 + idx) = make_float4(uu[0], uu[1], uu[2], uu[3]);
      *reinterpret_cast<float4*>(v2 + idx) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      *reinterpret_cast<float4*>(w2 + idx) = make_float4(ww[0], ww[1], ww[2], ww[3]);
      // ---- validators (k_check4) on the new faces ---------------------------------------------
#pragma unroll
      for (int e = 0; e < 4; ++e)
        bad += !isfinite(uu[e]) + !isfinite(vv[e]) + !isfinite(ww[e]);
      if (c4 & 0x40404040u) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (cc[e] & B_FLUID) {
            bad += !isfinite(px[e]);
            const float d = fabsf(quad_div(cc[e], uu[e], e ? uu[e - 1] : ul, vv[e], vm[e], ww[e],
                                           wprev[e]));
            if (d == d) mx = fmaxf(mx, d);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) wprev[e] = ww[e];
    // the next plane's own row (read after the next barrier; buffer b ^ 1 was last read before
    // this iteration's barrier)
    sm.m[b ^ 1][ty + 1][lane] = m_n;
    *reinterpret_cast<double2*>(&sm.p[b ^ 1][ty + 1][4 * lane]) = pn0;
    *reinterpret_cast<double2*>(&sm.p[b ^ 1][ty + 1][4 * lane + 2]) = pn1;
    if (edge) {
      const unsigned f = hm & gc_ns4(hm);
      sm.m[b ^ 1][erow][lane] = hm;
      *reinterpret_cast<double2*>(&sm.p[b ^ 1][erow][4 * lane]) = gc_sel2(f, 0, h0);
      *reinterpret_cast<double2*>(&sm.p[b ^ 1][erow][4 * lane + 2]) = gc_sel2(f, 2, h1);
    }
  }
  const double bm = block_max((double)mx, red);
  const double bb = block_sum((double)bad, red);
  const unsigned nb = num_blocks();
  if (lane == 0 && ty == 0) {
    partials[linear_block()] = bm;
    if (bb > 0.0) atomicAdd(&sc->nonfinite, (unsigned long long)bb);
  }
  if (last_block(&sc->ticketC, nb)) {
    const unsigned tid = lane + 32 * ty;
    double t = 0.0;
    for (unsigned q = tid; q < nb; q += 32 * GC_TY) t = fmax(t, __ldcg(partials + q));
    t = block_max(t, red);
    if (tid == 0) sc->gD[sc->rank] = t;
  }
}

// ---- true residual (iterative refinement of the fp32 solve) -----------------------------
// r = b - A x with x the fp64 pressure accumulator and the stencil evaluated in fp64:
// the recursive residual of PCG drifts from the true one by ~eps32 |A| |x| (the search
// directions are fp32), which is what limits the divergence-free residual of a solve that
// converges in a few large steps (multigrid).  One restart from this residual removes it.
// In place: r holds b on entry.
template <typename T>
__global__ void __launch_bounds__(256)
k_true_residual(Geo g, const uint8_t* __restrict__ codes, const double* __restrict__ x,
                T* __restrict__ r) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const unsigned c = codes[idx];
    T out = T(0);
    if (c & B_FLUID) {
      const double x0 = x[idx];
      double sum = 0.0;
      // Air neighbours carry p = 0 (Constants.fs:20) and x is 0 on every non-Fluid cell
      if (c & B_XM) sum += x[idx - 1];
      if (c & B_XP) sum += x[idx + 1];
      if (c & B_YM) sum += x[idx - g.nx];
      if (c & B_YP) sum += x[idx + g.nx];
      if (c & B_ZM) sum += x[idx - g.plane];
      if (c & B_ZP) sum += x[idx + g.plane];
      out = (T)((double)r[idx] - ((double)__popc(c & 63u) * x0 - sum));
    }
    r[idx] = out;
  }
}

// 4 cells per thread (fp32 residual, nx % 4 = 0): same sums in the same order as
// k_true_residual; neighbour loads are unconditional (the allocation has z-halo planes) and
// selected by the code bits.
__global__ void __launch_bounds__(256)
k_true_residual4(Geo g, const uint8_t* __restrict__ codes, const double* __restrict__ x,
                 float* __restrict__ r) {
  const long long nq = g.ncell >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const long long idx = q << 2;
    const unsigned c4 = *reinterpret_cast<const unsigned*>(codes + idx);
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 & 0x40404040u) {
      const float4 b4 = *reinterpret_cast<const float4*>(r + idx);
      const double2 xa = *reinterpret_cast<const double2*>(x + idx);
      const double2 xb = *reinterpret_cast<const double2*>(x + idx + 2);
      const double xs[6] = {(c4 & B_XM) ? x[idx - 1] : 0.0, xa.x, xa.y, xb.x, xb.y,
                            ((c4 >> 24) & B_XP) ? x[idx + 4] : 0.0};
      const double2 ya = *reinterpret_cast<const double2*>(x + idx - g.nx);
      const double2 yb = *reinterpret_cast<const double2*>(x + idx - g.nx + 2);
      const double2 yc = *reinterpret_cast<const double2*>(x + idx + g.nx);
      const double2 yd = *reinterpret_cast<const double2*>(x + idx + g.nx + 2);
      const double2 za = *reinterpret_cast<const double2*>(x + idx - g.plane);
      const double2 zb = *reinterpret_cast<const double2*>(x + idx - g.plane + 2);
      const double2 zc = *reinterpret_cast<const double2*>(x + idx + g.plane);
      const double2 zd = *reinterpret_cast<const double2*>(x + idx + g.plane + 2);
      const double ym[4] = {ya.x, ya.y, yb.x, yb.y}, yp[4] = {yc.x, yc.y, yd.x, yd.y};
      const double zm[4] = {za.x, za.y, zb.x, zb.y}, zp[4] = {zc.x, zc.y, zd.x, zd.y};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned c = (c4 >> (8 * e)) & 0xffu;
        o[e] = 0.f;
        if (c & B_FLUID) {
          double sum = 0.0;
          if (c & B_XM) sum += xs[e];
          if (c & B_XP) sum += xs[e + 2];
          if (c & B_YM) sum += ym[e];
          if (c & B_YP) sum += yp[e];
          if (c & B_ZM) sum += zm[e];
          if (c & B_ZP) sum += zp[e];
          o[e] = (float)((double)bb[e] - ((double)__popc(c & 63u) * xs[e + 1] - sum));
        }
      }
      out = make_float4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<float4*>(r + idx) = out;
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

// ---- viscosity -------------------------------------------------------
// Viscosity.fs:12-36 as a Jacobi update (quirk Q10 kept: no 1/h^2, always -6 v; a
// neighbour contributes only its d-axis component, only if it is Fluid and that
// component has the sign of d).  Reads (u,v,w), writes (u2,v2,w2).
template <typename T>
__global__ void __launch_bounds__(256)
k_viscosity(Geo g, const uint8_t* __restrict__ media, const T* __restrict__ u,
            const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ u2,
            T* __restrict__ v2, T* __restrict__ w2, T k) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int kk = (int)(t / g.ny);
    const int kg = kk + g.z0;
    const T* f[3] = {u, v, w};
    T* o[3] = {u2, v2, w2};
    const bool fl = media[idx] == FLUID;
    const bool inm[3] = {i > 0, j > 0, kg > 0};
    const bool inp[3] = {i < g.nx - 1, j < g.ny - 1, kg < g.nzg - 1};
    const long long off[3] = {1, (long long)g.nx, g.plane};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const T own = f[a][idx];
      T val = own;
      if (fl) {
        T shared = T(0);
        if (inp[a] && media[idx + off[a]] == FLUID) {
          const T n = f[a][idx + off[a]];
          if (n > T(0)) shared += n;          // Coord.is_bordering PosX: v.x > 0
        }
        if (inm[a] && media[idx - off[a]] == FLUID) {
          const T n = f[a][idx - off[a]];
          if (n < T(0)) shared += n;          // NegX: v.x < 0
        }
        val = own + (shared - T(6) * own) * k;
      }
      o[a][idx] = val;
    }
  }
}

// ---- post-projection cleanup ---------------------------------------------
// BFS layers from the Fluid cells (Build.fs:79-102 numbering): 0 Fluid, 1..2 buffer,
// 255 elsewhere.  Step i marks the unlabelled existing cells next to layer i-1.
__global__ void k_layer_init(Geo g, const uint8_t* __restrict__ media,
                             uint8_t* __restrict__ layer) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x)
    layer[idx] = (media[idx] == FLUID) ? 0 : 255;
}

__global__ void k_layer_step(Geo g, const uint8_t* __restrict__ media,
                             uint8_t* __restrict__ layer, int step) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    if (layer[idx] != 255 || media[idx] == ABSENT) continue;
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const int kg = k + g.z0;
    const uint8_t want = (uint8_t)(step - 1);
    bool near = false;
    near |= (i > 0) && layer[idx - 1] == want;
    near |= (i < g.nx - 1) && layer[idx + 1] == want;
    near |= (j > 0) && layer[idx - g.nx] == want;
    near |= (j < g.ny - 1) && layer[idx + g.nx] == want;
    near |= (kg > 0) && layer[idx - g.plane] == want;
    near |= (kg < g.nzg - 1) && layer[idx + g.plane] == want;
    if (near) layer[idx] = (uint8_t)step;
  }
}

// Build.propagate_velocities (order-free reading, see oracle/dense_ref.py) followed
// by Build.zero_solid_velocities and the "solid velocities are zero" rule
// (Build.fs:105-160, Simulator.fs:100-110).  Reads (u,v,w), writes (u2,v2,w2).
template <typename T>
__global__ void __launch_bounds__(256)
k_cleanup(Geo g, const uint8_t* __restrict__ media, const uint8_t* __restrict__ layer,
          const T* __restrict__ u, const T* __restrict__ v, const T* __restrict__ w,
          T* __restrict__ u2, T* __restrict__ v2, T* __restrict__ w2) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const int kg = k + g.z0;
    const T* f[3] = {u, v, w};
    T* o[3] = {u2, v2, w2};
    const uint8_t m0 = media[idx];
    const uint8_t l0 = layer[idx];
    const bool inm[3] = {i > 0, j > 0, kg > 0};
    const bool inp[3] = {i < g.nx - 1, j < g.ny - 1, kg < g.nzg - 1};
    const long long off[3] = {1, (long long)g.nx, g.plane};
    bool touched = false;
    uint8_t lp[3], lm[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lp[a] = inp[a] ? layer[idx + off[a]] : 255;
      lm[a] = inm[a] ? layer[idx - off[a]] : 255;
    }
    const uint8_t prev = (uint8_t)(l0 - 1);
    if (l0 >= 1 && l0 != 255) {
#pragma unroll
      for (int a = 0; a < 3; ++a) touched |= (lp[a] == prev) || (lm[a] == prev);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      T val = f[a][idx];
      const uint8_t mp = inp[a] ? media[idx + off[a]] : ABSENT;
      if (touched) {
        const uint8_t mm = inm[a] ? media[idx - off[a]] : ABSENT;
        T avg = T(0);
        if (lp[a] == prev) avg += val;
        if (lm[a] == prev) avg += f[a][idx - off[a]];
        const uint8_t md = (val < T(0)) ? mm : mp;       // Build.fs:123-125
        if (md != ABSENT && md != FLUID) val = avg;
      }
      // Build.fs:144-160: a component pointing into a Solid +neighbour is zeroed
      if (is_nonsolid(m0) && mp == SOLID && val > T(0)) val = T(0);
      if (m0 == SOLID) val = T(0);                       // Simulator.fs:106-109
      o[a][idx] = val;
    }
  }
}

}  // namespace spl
