// spl_markers.cuh -- SURVEY 8f row N3: marker motion and cell relabelling on the device
//
// The step right before the hot path (Simulator.fs:60-79): Simulator.move_markers
// (:38-52), Grid.move_cells (Grid.fs:42-47) and the Build.setup block (Build.fs:37-102:
// add_new_markers, set_fluid_layers, create_air_buffer, delete_unused).  The reference
// keeps the grid in a Dictionary and edits it cell by cell; here the labels live in the
// dense box, markers are an array of fp64 locations, and every stage is one data-parallel
// pass (the BFS of the air buffer: max_distance = 2 passes).  Oracle:
// oracle/dense_ref.py move_markers / move_cells / relabel.
//
// z-slabs (world > 1): every rank holds ALL marker locations.  A marker is moved by the rank
// whose slab holds its cell (markers outside the global box: rank 0), the other ranks write
// zeros, and one all-reduce (sum: exact, a single non-zero term) hands every rank the new
// locations -- no migration, no variable-size messages.  The relabelling passes read the
// neighbouring slab's boundary plane of the labels / layers through the z halo, exchanged
// between the passes; bounds and the world test use the global plane index.
#pragma once
#include "spl_common.cuh"

namespace spl {

struct MarkerStats {
  unsigned long long moved;      // markers whose cell changed (Simulator.fs:45-50)
  unsigned long long outside;    // markers outside the dense box (raw_get would throw)
  unsigned long long escaped;    // markers outside the world (Build.check_containment)
};

// Coord.construct(float) (Coord.fs:49-67) as a world cell index: F# `round` is banker's
// rounding (= llrint in the default rounding mode), then the snap to a multiple of h with
// the .NET remainder (sign of the dividend -- as C++ %).
__device__ __forceinline__ long long nearest_cell(double x, int h) {
  const long long n = llrint(x);
  const long long r = n % h;
  const long long s = (r == 0) ? n : ((r >= h / 2) ? n + (h - r) : n - r);
  return s / h;
}

// local cell index of a location; MC_OUTSIDE = outside the (global) dense box, MC_REMOTE =
// inside it but in another rank's slab.  (ox, oy, oz) = origin of the global box.
constexpr long long MC_OUTSIDE = -1, MC_REMOTE = -2;
__device__ __forceinline__ long long marker_cell(const Geo& g, const double* loc, int h,
                                                 int ox, int oy, int oz) {
  const long long i = nearest_cell(loc[0], h) - ox;
  const long long j = nearest_cell(loc[1], h) - oy;
  const long long kg = nearest_cell(loc[2], h) - oz;
  if (i < 0 || j < 0 || kg < 0 || i >= g.nx || j >= g.ny || kg >= g.nzg) return MC_OUTSIDE;
  const long long k = kg - g.z0;
  if (k < 0 || k >= g.nz) return MC_REMOTE;
  return (k * g.ny + j) * g.nx + i;
}

// Simulator.move_markers: loc += dt * get_velocity(cell(loc)); get_velocity (:24-35) =
// the cell's +face velocity plus the -neighbours' components on the shared faces (an
// absent neighbour contributes 0).  sum_faces = the reference's literal sum (quirk);
// otherwise the mean of the two faces = the cell-centre velocity.
template <typename T>
__global__ void k_move_markers(Geo g, const uint8_t* __restrict__ media,
                               const T* __restrict__ u, const T* __restrict__ v,
                               const T* __restrict__ w, double* __restrict__ loc,
                               long long n, double dt, int h, int ox, int oy, int oz,
                               int sum_faces, long long* __restrict__ old_cell,
                               long long* __restrict__ new_cell, MarkerStats* st, int rank,
                               int world) {
  unsigned n_out = 0, n_moved = 0;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n;
       m += (long long)gridDim.x * blockDim.x) {
    double l[3] = {loc[3 * m], loc[3 * m + 1], loc[3 * m + 2]};
    const long long idx = marker_cell(g, l, h, ox, oy, oz);
    old_cell[m] = idx;
    if (idx < 0) {
      // outside the box: stays where it is (rank 0 keeps and counts it); in another slab: its
      // owner moves it.  Ranks that do not own a marker contribute zeros to the all-reduce.
      new_cell[m] = idx;
      if (idx == MC_OUTSIDE && rank == 0) ++n_out;
      if (world > 1 && !(idx == MC_OUTSIDE && rank == 0))
        loc[3 * m] = loc[3 * m + 1] = loc[3 * m + 2] = 0.0;
      continue;
    }
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const T* f[3] = {u, v, w};
    const long long off[3] = {1, (long long)g.nx, g.plane};
    const bool inm[3] = {i > 0, j > 0, k + g.z0 > 0};   // plane -1 of a slab: the z halo
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double vel = (double)f[a][idx];
      if (inm[a] && media[idx - off[a]] != ABSENT) vel += (double)f[a][idx - off[a]];
      if (!sum_faces) vel *= 0.5;
      l[a] = __dadd_rn(l[a], __dmul_rn(vel, dt));   // no FMA: l .+ (v .* dt), Simulator.fs:42
      loc[3 * m + a] = l[a];
    }
    const long long nidx = marker_cell(g, l, h, ox, oy, oz);
    new_cell[m] = nidx;
    n_out += (nidx == MC_OUTSIDE);
    n_moved += (nidx != MC_OUTSIDE && nidx != idx);
  }
  // one atomic per warp (millions of markers move every step)
  n_out = __reduce_add_sync(0xffffffffu, n_out);
  n_moved = __reduce_add_sync(0xffffffffu, n_moved);
  if ((threadIdx.x & 31) == 0) {
    if (n_out) atomicAdd(&st->outside, (unsigned long long)n_out);
    if (n_moved) atomicAdd(&st->moved, (unsigned long long)n_moved);
  }
}

// cells of the current marker locations (no motion): used after spl_set_markers
// (count = 0: cells only -- the slabs' pass after the all-reduce of the locations)
__global__ void k_marker_cells(Geo g, const double* __restrict__ loc, long long n, int h,
                               int ox, int oy, int oz, long long* __restrict__ cell,
                               MarkerStats* st, int count) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n;
       m += (long long)gridDim.x * blockDim.x) {
    const double l[3] = {loc[3 * m], loc[3 * m + 1], loc[3 * m + 2]};
    const long long idx = marker_cell(g, l, h, ox, oy, oz);
    cell[m] = idx;
    if (count && idx == MC_OUTSIDE) atomicAdd(&st->outside, 1ull);
  }
}

// Grid.move_cells, the reference's literal rule (SPL_Q_MOVE_CELLS): the cell record --
// media, velocity, pressure -- travels with its marker and the old coordinate is deleted.
// One thread per moved marker; exact when the moves do not interact (no marker moves into
// a cell another moving marker leaves or enters), which is every state the reference
// reaches with isolated markers.
template <typename T>
__global__ void k_move_records(Geo g, uint8_t* __restrict__ media, T* __restrict__ u,
                               T* __restrict__ v, T* __restrict__ w, double* __restrict__ p,
                               const long long* __restrict__ old_cell,
                               const long long* __restrict__ new_cell, long long n) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n;
       m += (long long)gridDim.x * blockDim.x) {
    const long long o = old_cell[m], d = new_cell[m];
    if (o < 0 || d < 0 || o == d) continue;
    media[d] = media[o];
    u[d] = u[o];
    v[d] = v[o];
    w[d] = w[o];
    p[d] = p[o];
    media[o] = ABSENT;
    u[o] = T(0);
    v[o] = T(0);
    w[o] = T(0);
    p[o] = 0.0;
  }
}

__global__ void k_mark_cells(const long long* __restrict__ cell, long long n,
                             uint8_t* __restrict__ mark) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n;
       m += (long long)gridDim.x * blockDim.x)
    if (cell[m] >= 0) mark[cell[m]] = 1;
}

struct WorldBox {
  int lo[3], hi[3];     // inclusive box-index bounds of World.contains (Aabb.fs:13-17)
};
__device__ __forceinline__ bool in_world(const WorldBox& wb, int i, int j, int k) {
  return i >= wb.lo[0] && i <= wb.hi[0] && j >= wb.lo[1] && j <= wb.hi[1] && k >= wb.lo[2] &&
         k <= wb.hi[2];
}

// add_new_markers + set_fluid_layers: Fluid = in-world, non-Solid cells holding a marker;
// a Fluid cell that lost its marker turns into Air.  mark[] comes in as "holds a marker"
// and leaves as "existed before this relabelling"; escaped counts markers that sit
// outside the world (Build.check_containment).
__global__ void k_relabel_init(Geo g, WorldBox wb, uint8_t* __restrict__ media,
                               uint8_t* __restrict__ layer, uint8_t* __restrict__ mark,
                               MarkerStats* st) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const uint8_t m0 = media[idx];
    const bool has = mark[idx] != 0;
    const bool inw = in_world(wb, i, j, k + g.z0);
    if (has && !inw) atomicAdd(&st->escaped, 1ull);
    const bool fluid = has && inw && m0 != SOLID;
    const uint8_t lab = fluid ? FLUID : (m0 == FLUID ? AIR : m0);
    media[idx] = lab;
    layer[idx] = fluid ? 0 : 255;
    mark[idx] = (lab != ABSENT) ? 1 : 0;
  }
}

// create_air_buffer, ring `step`: an unlabelled non-Fluid cell next to a non-solid cell of
// layer step-1 joins layer `step`; if it did not exist it is created as Air (inside the
// world) or Solid (outside).  lazy = only cells that existed before this relabelling act
// as sources (the reference reads the grid before its additions are committed,
// Build.fs:83-86: SPL_Q_LAZY_BUFFER).
__global__ void k_relabel_step(Geo g, WorldBox wb, uint8_t* __restrict__ media,
                               uint8_t* __restrict__ layer, const uint8_t* __restrict__ mark,
                               int step, int lazy) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    if (layer[idx] != 255) continue;
    const uint8_t m0 = media[idx];
    if (m0 == FLUID) continue;
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const uint8_t want = (uint8_t)(step - 1);
    const long long off[6] = {-1, 1, -(long long)g.nx, (long long)g.nx, -g.plane, g.plane};
    const int kg = k + g.z0;                     // planes -1 / nz of a slab: the z halo
    const bool inb[6] = {i > 0, i < g.nx - 1, j > 0, j < g.ny - 1, kg > 0, kg < g.nzg - 1};
    bool near = false;
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      if (!inb[d]) continue;
      const long long n = idx + off[d];
      // a source keeps its layer and label for the rest of this pass; cells labelled by
      // this pass carry `step` (or still 255), never step-1
      if (layer[n] == want) {
        const uint8_t mn = media[n];
        if (mn != SOLID && mn != ABSENT && (!lazy || mark[n])) near = true;
      }
    }
    if (near) {
      if (m0 == ABSENT) media[idx] = in_world(wb, i, j, kg) ? AIR : SOLID;
      layer[idx] = (uint8_t)step;
    }
  }
}

// delete_unused + the zero state of created cells: cells without a layer are deleted;
// deleted and newly created cells carry zero velocity and pressure, Air pressure is
// Constants.atmospheric_pressure = 0 (Constants.fs:20, Build.fs:92-94).
template <typename T>
__global__ void k_relabel_finish(Geo g, uint8_t* __restrict__ media,
                                 const uint8_t* __restrict__ layer,
                                 const uint8_t* __restrict__ mark, T* __restrict__ u,
                                 T* __restrict__ v, T* __restrict__ w, double* __restrict__ p) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.ncell;
       idx += (long long)gridDim.x * blockDim.x) {
    const bool gone = layer[idx] == 255;
    if (gone) media[idx] = ABSENT;
    const bool fresh = gone || !mark[idx];
    if (fresh) {
      u[idx] = T(0);
      v[idx] = T(0);
      w[idx] = T(0);
    }
    if (fresh || media[idx] == AIR) p[idx] = 0.0;
  }
}

// Build.check_surroundings (Build.fs:168-174): every marker cell has all six neighbours
__global__ void k_check_surroundings(Geo g, const uint8_t* __restrict__ media,
                                     const long long* __restrict__ cell, long long n,
                                     unsigned long long* __restrict__ missing) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < n;
       m += (long long)gridDim.x * blockDim.x) {
    const long long idx = cell[m];
    if (idx < 0) continue;
    const int i = (int)(idx % g.nx);
    const long long t = idx / g.nx;
    const int j = (int)(t % g.ny);
    const int k = (int)(t / g.ny);
    const long long off[6] = {-1, 1, -(long long)g.nx, (long long)g.nx, -g.plane, g.plane};
    const int kg = k + g.z0;
    const bool inb[6] = {i > 0, i < g.nx - 1, j > 0, j < g.ny - 1, kg > 0, kg < g.nzg - 1};
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 6; ++d) ok &= inb[d] && media[idx + off[d]] != ABSENT;
    if (!ok) atomicAdd(missing, 1ull);
  }
}

}  // namespace spl
