// splashy_host.cpp -- see splashy_host.hpp.  Also exports a small C interface
// (splh_*) so the pytest suite can drive this C++ host the way the reference's
// NUnit fixture drives the F# modules (tests/splashy.Tests/Tests.fs:39-52).
#include "splashy_host.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <tuple>

#include "../../include/splashy_cuda.h"

namespace splashy {

// ---- Vector.fs ----------------------------------------------------------------
Vector3d::Vector3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {
  if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z)))   // Vector.fs:15-18
    throw Failure("Vector3d has an invalid quantity: either NaN or ±Infinity.");
}

// ---- Coord.fs -------------------------------------------------------------------
int Coord::nearest(int n) {                       // Coord.fs:49-57 (.NET % truncates like C++)
  const int h = Constants::h_int;
  const int r = n % h;
  if (r == 0) return n;
  if (r >= h / 2) return n + (h - r);
  return n - r;
}
Coord Coord::construct(int x, int y, int z) { return {nearest(x), nearest(y), nearest(z)}; }
Coord Coord::construct(double x, double y, double z) {   // round (banker's) -> int -> nearest
  auto f = [](double v) { return nearest((int)std::nearbyint(v)); };
  return {f(x), f(y), f(z)};
}
std::vector<std::pair<CoordDirection, Coord>> Coord::backward_neighbors() const {
  const int h = Constants::h_int;
  return {{CoordDirection::NegX, {x - h, y, z}},
          {CoordDirection::NegY, {x, y - h, z}},
          {CoordDirection::NegZ, {x, y, z - h}}};
}
std::vector<std::pair<CoordDirection, Coord>> Coord::forward_neighbors() const {
  const int h = Constants::h_int;
  return {{CoordDirection::PosX, {x + h, y, z}},
          {CoordDirection::PosY, {x, y + h, z}},
          {CoordDirection::PosZ, {x, y, z + h}}};
}
std::vector<std::pair<CoordDirection, Coord>> Coord::neighbors() const {
  auto f = forward_neighbors();
  auto b = backward_neighbors();
  f.insert(f.end(), b.begin(), b.end());
  return f;
}
Coord Coord::get_neighbor(CoordDirection d) const {
  const int h = Constants::h_int;
  switch (d) {
    case CoordDirection::NegX: return {x - h, y, z};
    case CoordDirection::PosX: return {x + h, y, z};
    case CoordDirection::NegY: return {x, y - h, z};
    case CoordDirection::PosY: return {x, y + h, z};
    case CoordDirection::NegZ: return {x, y, z - h};
    default: return {x, y, z + h};
  }
}

// ---- Grid.fs ----------------------------------------------------------------------
std::optional<Cell> Grid::get(const Coord& where) const {
  auto it = cells_.find(where);
  if (it == cells_.end()) return std::nullopt;
  return it->second;
}
const Cell& Grid::raw_get(const Coord& where) const {
  auto it = cells_.find(where);
  if (it == cells_.end()) throw Failure("Option.get: no cell at this coordinate.");
  return it->second;
}
std::vector<Coord> Grid::filter(const std::function<bool(const Cell&)>& fn) const {
  std::vector<Coord> out;
  for (const auto& kv : cells_)
    if (fn(kv.second)) out.push_back(kv.first);
  return out;
}
void Grid::set(const Coord& where, const Cell& c) {
  auto it = cells_.find(where);
  if (it == cells_.end()) throw Failure("Tried to set non-existent cell.");   // Grid.fs:35
  it->second = c;
  ++version_;
}
void Grid::add_cells(const std::vector<std::pair<Coord, Cell>>& cells) {
  for (const auto& kv : cells) {
    if (!cells_.emplace(kv.first, kv.second).second)
      throw Failure("An item with the same key has already been added.");
  }
  ++version_;
}
void Grid::delete_cells(const std::vector<Coord>& cells) {
  for (const auto& c : cells) cells_.erase(c);
  ++version_;
}
void Grid::update_velocities(const std::vector<std::pair<Coord, Vector3d>>& v) {
  for (const auto& kv : v) {                       // Grid.fs:49-53
    Cell c = raw_get(kv.first);
    c.velocity = kv.second;
    set(kv.first, c);
  }
}
void Grid::update_pressures(const std::vector<std::pair<Coord, double>>& p) {
  for (const auto& kv : p) {                       // Grid.fs:55-59
    Cell c = raw_get(kv.first);
    c.pressure = kv.second;
    set(kv.first, c);
  }
}
void Grid::update_media(const std::vector<std::pair<Coord, Media>>& m) {
  for (const auto& kv : m) {
    Cell c = raw_get(kv.first);
    c.media = kv.second;
    set(kv.first, c);
  }
}

// ---- packing ------------------------------------------------------------------------
long Box::index(const Coord& c) const {
  const int h = Constants::h_int;
  const int i = c.x / h - ox, j = c.y / h - oy, k = c.z / h - oz;
  if (i < 0 || i >= nx || j < 0 || j >= ny || k < 0 || k >= nz) return -1;
  return ((long)k * ny + j) * nx + i;
}

Box pack(const Grid& g) {
  const int h = Constants::h_int;
  Box b;
  if (g.size() == 0) throw Failure("cannot pack an empty grid");
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (const auto& kv : g.cells()) {
    const int c[3] = {kv.first.x / h, kv.first.y / h, kv.first.z / h};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], c[a]);
      hi[a] = std::max(hi[a], c[a]);
    }
  }
  b.ox = lo[0]; b.oy = lo[1]; b.oz = lo[2];
  b.nx = hi[0] - lo[0] + 1; b.ny = hi[1] - lo[1] + 1; b.nz = hi[2] - lo[2] + 1;
  const size_t n = (size_t)b.nx * b.ny * b.nz;
  b.media.assign(n, SPL_ABSENT);
  b.u.assign(n, 0.0); b.v.assign(n, 0.0); b.w.assign(n, 0.0); b.p.assign(n, 0.0);
  for (const auto& kv : g.cells()) {
    const long idx = b.index(kv.first);
    const Cell& c = kv.second;
    b.media[idx] = (uint8_t)c.media;
    b.u[idx] = c.velocity.x;          // +face convention, Pressure.fs:18-31
    b.v[idx] = c.velocity.y;
    b.w[idx] = c.velocity.z;
    b.p[idx] = c.pressure.value_or(0.0);
  }
  return b;
}

Grid unpack(const Box& b) {
  Grid g;
  std::vector<std::pair<Coord, Cell>> cells;
  const int h = Constants::h_int;
  for (int k = 0; k < b.nz; ++k)
    for (int j = 0; j < b.ny; ++j)
      for (int i = 0; i < b.nx; ++i) {
        const size_t idx = ((size_t)k * b.ny + j) * b.nx + i;
        if (b.media[idx] == SPL_ABSENT) continue;
        Cell c;
        c.media = (Media)b.media[idx];
        c.velocity = Vector3d(b.u[idx], b.v[idx], b.w[idx]);
        if (c.media != Media::Solid) c.pressure = b.p[idx];
        cells.push_back({Coord{(i + b.ox) * h, (j + b.oy) * h, (k + b.oz) * h}, c});
      }
  g.add_cells(cells);
  return g;
}

// ---- state files ------------------------------------------------------------------------
namespace {
constexpr char kMagic[8] = {'S', 'P', 'L', 'S', 'T', 'A', 'T', 'E'};
struct File {
  FILE* f;
  std::string path;
  File(const std::string& p, const char* mode) : f(std::fopen(p.c_str(), mode)), path(p) {
    if (!f) throw Failure("cannot open state file " + p);
  }
  ~File() { if (f) std::fclose(f); }
  void put(const void* d, size_t n) {
    if (n && std::fwrite(d, 1, n, f) != n) throw Failure("short write to " + path);
  }
  void get(void* d, size_t n) {
    if (n && std::fread(d, 1, n, f) != n) throw Failure("state file " + path + " is truncated");
  }
  template <typename V> void put(const V& v) { put(&v, sizeof v); }
  template <typename V> V get() { V v; get(&v, sizeof v); return v; }
};
void put_header(File& f, uint32_t kind) {
  f.put(kMagic, 8);
  f.put<uint32_t>(1);
  f.put<uint32_t>(kind);
}
void get_header(File& f, uint32_t kind) {
  char m[8];
  f.get(m, 8);
  if (std::memcmp(m, kMagic, 8) != 0) throw Failure(f.path + " is not a SPLSTATE file");
  const uint32_t version = f.get<uint32_t>(), k = f.get<uint32_t>();
  if (version != 1) throw Failure(f.path + ": unsupported SPLSTATE version");
  if (k != kind) throw Failure(f.path + (kind ? ": not a dense box" : ": not a sparse grid"));
}
}  // namespace

void save_grid(const std::string& path, const Grid& g) {
  std::vector<std::pair<Coord, Cell>> cells(g.cells().begin(), g.cells().end());
  std::sort(cells.begin(), cells.end(), [](const auto& a, const auto& b) {
    return std::tie(a.first.z, a.first.y, a.first.x) < std::tie(b.first.z, b.first.y, b.first.x);
  });
  File f(path, "wb");
  put_header(f, 0);
  f.put<double>(Constants::h);
  f.put<uint64_t>(cells.size());
  for (const auto& kv : cells) {
    const int32_t xyz[3] = {kv.first.x, kv.first.y, kv.first.z};
    const uint8_t tag[4] = {(uint8_t)kv.second.media, (uint8_t)(kv.second.pressure ? 1 : 0), 0, 0};
    const double val[4] = {kv.second.velocity.x, kv.second.velocity.y, kv.second.velocity.z,
                           kv.second.pressure.value_or(0.0)};
    f.put(xyz, sizeof xyz);
    f.put(tag, sizeof tag);
    f.put(val, sizeof val);
  }
}

Grid load_grid(const std::string& path) {
  File f(path, "rb");
  get_header(f, 0);
  const double h = f.get<double>();
  if (h != Constants::h) throw Failure(path + ": cell size differs from Constants.h");
  const uint64_t n = f.get<uint64_t>();
  std::vector<std::pair<Coord, Cell>> cells;
  cells.reserve(n);
  for (uint64_t i = 0; i < n; ++i) {
    int32_t xyz[3];
    uint8_t tag[4];
    double val[4];
    f.get(xyz, sizeof xyz);
    f.get(tag, sizeof tag);
    f.get(val, sizeof val);
    if (tag[0] > 2) throw Failure(path + ": unknown media value");
    Cell c;
    c.media = (Media)tag[0];
    c.velocity = Vector3d(val[0], val[1], val[2]);     // refuses NaN / Inf like Vector.fs:15-18
    if (tag[1]) c.pressure = val[3];
    cells.push_back({Coord{xyz[0], xyz[1], xyz[2]}, c});
  }
  Grid g;
  g.add_cells(cells);
  return g;
}

void save_box(const std::string& path, const Box& b) {
  File f(path, "wb");
  put_header(f, 1);
  const int32_t dims[6] = {b.nx, b.ny, b.nz, b.ox, b.oy, b.oz};
  f.put(dims, sizeof dims);
  f.put<double>(Constants::h);
  const size_t n = (size_t)b.nx * b.ny * b.nz;
  f.put(b.media.data(), n);
  for (const std::vector<double>* a : {&b.u, &b.v, &b.w, &b.p}) f.put(a->data(), n * sizeof(double));
}

Box load_box(const std::string& path) {
  File f(path, "rb");
  get_header(f, 1);
  int32_t dims[6];
  f.get(dims, sizeof dims);
  const double h = f.get<double>();
  if (h != Constants::h) throw Failure(path + ": cell size differs from Constants.h");
  if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) throw Failure(path + ": empty box");
  Box b;
  b.nx = dims[0]; b.ny = dims[1]; b.nz = dims[2];
  b.ox = dims[3]; b.oy = dims[4]; b.oz = dims[5];
  const size_t n = (size_t)b.nx * b.ny * b.nz;
  b.media.resize(n);
  f.get(b.media.data(), n);
  for (std::vector<double>* a : {&b.u, &b.v, &b.w, &b.p}) {
    a->resize(n);
    f.get(a->data(), n * sizeof(double));
  }
  return b;
}

namespace {
// NumPy format 1.0: magic, version, u16 header length, ASCII dict padded to a multiple of 64
void write_npy(const std::string& path, const char* descr, const Box& b, const void* data,
               size_t bytes) {
  std::string hd = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (" +
                   std::to_string(b.nz) + ", " + std::to_string(b.ny) + ", " +
                   std::to_string(b.nx) + "), }";
  while ((10 + hd.size() + 1) % 64 != 0) hd += ' ';
  hd += '\n';
  File f(path, "wb");
  const unsigned char pre[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.put(pre, 8);
  f.put<uint16_t>((uint16_t)hd.size());
  f.put(hd.data(), hd.size());
  f.put(data, bytes);
}
}  // namespace

void save_box_npy(const std::string& dir, const Box& b) {
  const size_t n = (size_t)b.nx * b.ny * b.nz;
  write_npy(dir + "/media.npy", "|u1", b, b.media.data(), n);
  write_npy(dir + "/u.npy", "<f8", b, b.u.data(), n * sizeof(double));
  write_npy(dir + "/v.npy", "<f8", b, b.v.data(), n * sizeof(double));
  write_npy(dir + "/w.npy", "<f8", b, b.w.data(), n * sizeof(double));
  write_npy(dir + "/p.npy", "<f8", b, b.p.data(), n * sizeof(double));
  File f(dir + "/meta.json", "wb");
  const std::string meta = "{\"origin\": [" + std::to_string(b.ox) + ", " + std::to_string(b.oy) + ", " +
                           std::to_string(b.oz) + "], \"h\": " + std::to_string(Constants::h_int) +
                           ", \"shape\": [" + std::to_string(b.nz) + ", " + std::to_string(b.ny) +
                           ", " + std::to_string(b.nx) + "]}\n";
  f.put(meta.data(), meta.size());
}

// ---- Solver ---------------------------------------------------------------------------
Solver::Solver(unsigned quirks, double tol, int device)
    : quirks_(quirks), tol_(tol), device_(device) {}

Solver::~Solver() {
  if (ctx_) spl_destroy(ctx_);
}

void Solver::fail_if(int rc) {
  if (rc != SPL_OK) throw Failure(spl_last_error(ctx_));   // text mirrors the failwith strings
}

void Solver::sync(const Grid& g) {
  Box b = pack(g);
  if (!ctx_ || b.nx != box_.nx || b.ny != box_.ny || b.nz != box_.nz || b.ox != box_.ox ||
      b.oy != box_.oy || b.oz != box_.oz) {
    if (ctx_) spl_destroy(ctx_);
    ctx_ = nullptr;
    spl_desc d;
    spl_desc_default(&d);
    d.nx = b.nx; d.ny = b.ny; d.nz = b.nz;
    d.dtype = SPL_F64;                     // the reference is fp64 (Vector.fs:11-13)
    d.tol = tol_;                          // the reference solves exactly (Pressure.fs:69)
    d.quirks = quirks_;
    d.origin[0] = b.ox; d.origin[1] = b.oy; d.origin[2] = b.oz;
    d.device = device_;
    const int rc = spl_create(&ctx_, &d);
    if (rc != SPL_OK) throw Failure(spl_last_error(nullptr));
  }
  box_ = std::move(b);
  fail_if(spl_upload(ctx_, box_.media.data(), box_.u.data(), box_.v.data(), box_.w.data(),
                     box_.p.data()));
  projected_ = false;
}

void Solver::download() {
  fail_if(spl_download(ctx_, box_.u.data(), box_.v.data(), box_.w.data(), box_.p.data(),
                       nullptr));
}

std::vector<std::pair<Coord, Vector3d>> Solver::convection_apply(
    const Grid& g, const std::vector<Coord>& markers, double dt) {
  sync(g);
  spl_info info;
  fail_if(spl_advect(ctx_, dt, &info));    // SPL_E_TRACE -> "Backwards particle trace went too far."
  download();
  std::vector<std::pair<Coord, Vector3d>> out;
  for (const Coord& m : markers) {
    const long idx = box_.index(m);
    if (idx < 0) throw Failure("Option.get: no cell at this coordinate.");
    out.emplace_back(m, Vector3d(box_.u[idx], box_.v[idx], box_.w[idx]));
  }
  return out;
}

std::vector<std::pair<Coord, Vector3d>> Solver::forces_apply(
    const Grid& g, const std::vector<Coord>& markers, double dt) {
  sync(g);
  fail_if(spl_add_forces(ctx_, dt));
  download();
  std::vector<std::pair<Coord, Vector3d>> out;
  for (const Coord& m : markers) {
    const long idx = box_.index(m);
    if (idx < 0) throw Failure("Option.get: no cell at this coordinate.");
    out.emplace_back(m, Vector3d(box_.u[idx], box_.v[idx], box_.w[idx]));
  }
  return out;
}

double Solver::pressure_divergence(const Grid& g, const Coord& where) {
  g.raw_get(where);
  for (const auto& n : where.neighbors()) g.raw_get(n.second);   // Pressure.fs:14 raw_get throws
  // the divergence is reported on Fluid cells; the reference evaluates it for any
  // cell, so the queried cell is packed as Fluid for this call
  sync(g);
  const long idx = box_.index(where);
  std::vector<uint8_t> media = box_.media;
  media[idx] = SPL_FLUID;
  fail_if(spl_upload(ctx_, media.data(), box_.u.data(), box_.v.data(), box_.w.data(), nullptr));
  std::vector<double> div(box_.u.size());
  fail_if(spl_download(ctx_, nullptr, nullptr, nullptr, nullptr, div.data()));
  return div[idx];
}

std::vector<std::pair<Coord, double>> Solver::pressure_calculate(
    const Grid& g, const std::vector<Coord>& markers, double dt) {
  sync(g);
  spl_info info;
  fail_if(spl_project(ctx_, dt, &info));   // solve + gradient on the device
  last_iters_ = info.iters;
  download();
  projected_ = true;
  projected_version_ = g.version();
  std::vector<std::pair<Coord, double>> out;
  for (const Coord& m : markers) {
    const long idx = box_.index(m);
    if (idx < 0) throw Failure("Option.get: no cell at this coordinate.");
    out.emplace_back(m, box_.p[idx]);
  }
  return out;
}

std::vector<std::pair<Coord, Vector3d>> Solver::pressure_apply(
    const Grid& g, const std::vector<Coord>& markers, double /*dt*/) {
  if (!projected_) throw Failure("Pressure.apply before Pressure.calculate");
  (void)g;
  // the projected velocities were produced by spl_project; every face the
  // projection may touch (markers and their -neighbours) is returned once
  std::vector<Coord> touched;
  auto push = [&](const Coord& c) {
    if (std::find(touched.begin(), touched.end(), c) == touched.end()) touched.push_back(c);
  };
  for (const Coord& m : markers) {
    push(m);
    for (const auto& n : m.backward_neighbors()) push(n.second);
  }
  std::vector<std::pair<Coord, Vector3d>> out;
  for (const Coord& c : touched) {
    const long idx = box_.index(c);
    if (idx < 0 || box_.media[idx] == SPL_ABSENT) continue;
    out.emplace_back(c, Vector3d(box_.u[idx], box_.v[idx], box_.w[idx]));
  }
  return out;
}

void Solver::check(const Grid& g, const std::vector<Coord>& markers) {
  (void)markers;
  sync(g);
  spl_info info;
  fail_if(spl_check(ctx_, 0.0001, &info));   // Pressure.fs:148 threshold
}

// ---- Simulator ----------------------------------------------------------------------
Simulator::Simulator(unsigned quirks, double tol, int margin, int device)
    : quirks_(quirks), tol_(tol), margin_(margin), device_(device) {}

Simulator::~Simulator() {
  if (ctx_) spl_destroy(ctx_);
}

void Simulator::fail_if(int rc) {
  if (rc != SPL_OK) throw Failure(spl_last_error(ctx_));
}

void Simulator::generate(const std::vector<Coord>& marker_cells) {
  std::vector<Coord> ms;
  for (const Coord& c : marker_cells) ms.push_back(Coord::construct(c.x, c.y, c.z));
  // Set.ofList: sorted by Coord.compare (x, then y, then z), duplicates dropped
  std::sort(ms.begin(), ms.end(), [](const Coord& a, const Coord& b) {
    return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z);
  });
  ms.erase(std::unique(ms.begin(), ms.end()), ms.end());
  locations_.clear();
  for (const Coord& m : ms) locations_.emplace_back((double)m.x, (double)m.y, (double)m.z);
  // Build.setup (fun () -> Build.add_new_markers markers |> Grid.add_cells)
  const float wh = 100.0f + (float)Constants::h_int / 2.0f;            // Constants.fs:32
  std::vector<std::pair<Coord, Cell>> fresh;
  for (const Coord& m : ms) {
    const bool in_world = std::fabs((float)m.x) <= wh && std::fabs((float)m.y) <= wh &&
                          std::fabs((float)m.z) <= wh;                  // World.contains
    if (!grid_.get(m) && in_world) {
      Cell c;
      c.media = Media::Fluid;
      fresh.insert(fresh.begin(), {m, c});
    }
  }
  grid_.add_cells(fresh);
  uploaded_ = false;
}

void Simulator::set_velocity(const Coord& where, const Vector3d& v) {
  if (uploaded_) throw Failure("set_velocity after the first advance");
  grid_.update_velocities({{where, v}});
}

void Simulator::ensure_uploaded() {
  if (uploaded_) return;
  Box b = pack(grid_);
  // room for Build.max_distance rings plus the cells a marker can reach
  Box big;
  big.ox = b.ox - margin_; big.oy = b.oy - margin_; big.oz = b.oz - margin_;
  big.nx = b.nx + 2 * margin_; big.ny = b.ny + 2 * margin_; big.nz = b.nz + 2 * margin_;
  const size_t n = (size_t)big.nx * big.ny * big.nz;
  big.media.assign(n, SPL_ABSENT);
  big.u.assign(n, 0.0); big.v.assign(n, 0.0); big.w.assign(n, 0.0); big.p.assign(n, 0.0);
  for (int k = 0; k < b.nz; ++k)
    for (int j = 0; j < b.ny; ++j)
      for (int i = 0; i < b.nx; ++i) {
        const size_t s = ((size_t)k * b.ny + j) * b.nx + i;
        const size_t d = ((size_t)(k + margin_) * big.ny + (j + margin_)) * big.nx + (i + margin_);
        big.media[d] = b.media[s];
        big.u[d] = b.u[s]; big.v[d] = b.v[s]; big.w[d] = b.w[s]; big.p[d] = b.p[s];
      }
  box_ = std::move(big);
  if (ctx_) spl_destroy(ctx_);
  ctx_ = nullptr;
  spl_desc d;
  spl_desc_default(&d);
  d.nx = box_.nx; d.ny = box_.ny; d.nz = box_.nz;
  d.dtype = SPL_F64;
  d.tol = tol_;
  d.quirks = quirks_;
  d.origin[0] = box_.ox; d.origin[1] = box_.oy; d.origin[2] = box_.oz;
  d.device = device_;
  if (spl_create(&ctx_, &d) != SPL_OK) throw Failure(spl_last_error(nullptr));
  fail_if(spl_upload(ctx_, box_.media.data(), box_.u.data(), box_.v.data(), box_.w.data(),
                     box_.p.data()));
  const int l = (int)((100.0f + (float)Constants::h_int / 2.0f) / (float)Constants::h_int);
  const int lo[3] = {-l - box_.ox, -l - box_.oy, -l - box_.oz};
  const int hi[3] = {l - box_.ox, l - box_.oy, l - box_.oz};
  fail_if(spl_set_world(ctx_, lo, hi));
  std::vector<double> xyz;
  for (const Vector3d& v : locations_) { xyz.push_back(v.x); xyz.push_back(v.y); xyz.push_back(v.z); }
  fail_if(spl_set_markers(ctx_, (int64_t)locations_.size(), xyz.data()));
  uploaded_ = true;
}

void Simulator::advance(double dt, bool sanity) {
  ensure_uploaded();
  spl_info info;
  fail_if(spl_advance(ctx_, dt, sanity ? 1 : 0, &info));
  last_iters_ = info.iters;
  std::vector<double> xyz(3 * locations_.size());
  fail_if(spl_get_markers(ctx_, xyz.data()));
  for (size_t m = 0; m < locations_.size(); ++m)
    locations_[m] = Vector3d(xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2]);
}

std::vector<Coord> Simulator::markers() const {
  std::vector<Coord> out;
  for (const Vector3d& l : locations_) out.push_back(Coord::construct(l.x, l.y, l.z));
  return out;
}

Grid Simulator::snapshot() {
  if (!uploaded_) return grid_;
  fail_if(spl_download(ctx_, box_.u.data(), box_.v.data(), box_.w.data(), box_.p.data(), nullptr));
  fail_if(spl_download_media(ctx_, box_.media.data()));
  return unpack(box_);
}

}  // namespace splashy

// ---------------------------------------------------------------------------------------
// C interface for the test-suite (ctypes)
// ---------------------------------------------------------------------------------------
using namespace splashy;

namespace {
thread_local std::string g_err;
template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
struct Host {
  Grid grid;
  std::unique_ptr<Solver> solver;
};
std::vector<Coord> coords_of(const int* xyz, int n) {
  std::vector<Coord> out;
  for (int i = 0; i < n; ++i) out.push_back({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
  return out;
}
}  // namespace

extern "C" {

const char* splh_last_error(void) { return g_err.c_str(); }

void* splh_new(unsigned quirks, double tol) {
  Host* h = new Host();
  h->solver.reset(new Solver(quirks, tol));
  return h;
}
void splh_free(void* h) { delete static_cast<Host*>(h); }

int splh_nearest(int n) { return Coord::nearest(n); }
void splh_construct_float(double x, double y, double z, int* out) {
  const Coord c = Coord::construct(x, y, z);
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

int splh_add_cell(void* hp, int x, int y, int z, int media, double vx, double vy, double vz,
                  int has_p, double p) {
  return guarded([&] {
    Cell c;
    c.media = (Media)media;
    c.velocity = Vector3d(vx, vy, vz);
    if (has_p) c.pressure = p;
    static_cast<Host*>(hp)->grid.add_cells({{Coord{x, y, z}, c}});
  });
}
int splh_get_cell(void* hp, int x, int y, int z, int* media, double* vel, int* has_p, double* p) {
  return guarded([&] {
    const Cell& c = static_cast<Host*>(hp)->grid.raw_get({x, y, z});
    *media = (int)c.media;
    vel[0] = c.velocity.x; vel[1] = c.velocity.y; vel[2] = c.velocity.z;
    *has_p = c.pressure.has_value();
    *p = c.pressure.value_or(0.0);
  });
}
int splh_size(void* hp) { return (int)static_cast<Host*>(hp)->grid.size(); }
// state files (splashy_host.hpp): kind 0 = sparse grid, 1 = dense box, 2 = NumPy directory
int splh_save(void* hp, int kind, const char* path) {
  return guarded([&] {
    const Grid& g = static_cast<Host*>(hp)->grid;
    if (kind == 0) save_grid(path, g);
    else if (kind == 1) save_box(path, pack(g));
    else save_box_npy(path, pack(g));
  });
}
int splh_load(void* hp, int kind, const char* path) {
  return guarded([&] {
    static_cast<Host*>(hp)->grid = (kind == 0) ? load_grid(path) : unpack(load_box(path));
  });
}
int splh_update_velocity(void* hp, int x, int y, int z, double vx, double vy, double vz) {
  return guarded([&] {
    static_cast<Host*>(hp)->grid.update_velocities({{Coord{x, y, z}, Vector3d(vx, vy, vz)}});
  });
}
int splh_update_media(void* hp, int x, int y, int z, int media) {
  return guarded([&] { static_cast<Host*>(hp)->grid.update_media({{Coord{x, y, z}, (Media)media}}); });
}

// Pressure.divergence coord
int splh_pressure_divergence(void* hp, int x, int y, int z, double* out) {
  Host* h = static_cast<Host*>(hp);
  return guarded([&] { *out = h->solver->pressure_divergence(h->grid, {x, y, z}); });
}
// Convection.apply markers dt |> Grid.update_velocities      (Simulator.fs:82)
int splh_convection(void* hp, const int* markers, int n, double dt) {
  Host* h = static_cast<Host*>(hp);
  return guarded([&] {
    h->grid.update_velocities(h->solver->convection_apply(h->grid, coords_of(markers, n), dt));
  });
}
// Forces.apply markers dt |> Grid.update_velocities          (Simulator.fs:84)
int splh_forces(void* hp, const int* markers, int n, double dt) {
  Host* h = static_cast<Host*>(hp);
  return guarded([&] {
    h->grid.update_velocities(h->solver->forces_apply(h->grid, coords_of(markers, n), dt));
  });
}
// Pressure.calculate |> update_pressures ; Pressure.apply |> update_velocities (:88-89)
int splh_pressure(void* hp, const int* markers, int n, double dt, int* iters) {
  Host* h = static_cast<Host*>(hp);
  return guarded([&] {
    const auto ms = coords_of(markers, n);
    const auto ps = h->solver->pressure_calculate(h->grid, ms, dt);
    const auto vs = h->solver->pressure_apply(h->grid, ms, dt);
    h->grid.update_pressures(ps);
    h->grid.update_velocities(vs);
    if (iters) *iters = h->solver->last_iterations();
  });
}
// Pressure.check_pressures ; Pressure.check_divergence        (Simulator.fs:91-97)
int splh_checks(void* hp, const int* markers, int n) {
  Host* h = static_cast<Host*>(hp);
  return guarded([&] { h->solver->check(h->grid, coords_of(markers, n)); });
}

// ---- Simulator (state on the device) ---------------------------------------------------
void* splh_sim_new(unsigned quirks, double tol, int margin) {
  return new Simulator(quirks, tol, margin);
}
void splh_sim_free(void* s) { delete static_cast<Simulator*>(s); }
int splh_sim_generate(void* s, const int* cells, int n) {
  return guarded([&] { static_cast<Simulator*>(s)->generate(coords_of(cells, n)); });
}
int splh_sim_set_velocity(void* s, int x, int y, int z, double vx, double vy, double vz) {
  return guarded([&] {
    static_cast<Simulator*>(s)->set_velocity({x, y, z}, Vector3d(vx, vy, vz));
  });
}
int splh_sim_advance(void* s, double dt, int sanity) {
  return guarded([&] { static_cast<Simulator*>(s)->advance(dt, sanity != 0); });
}
int splh_sim_marker_count(void* s) { return (int)static_cast<Simulator*>(s)->locations().size(); }
int splh_sim_locations(void* s, double* xyz) {
  return guarded([&] {
    const auto& l = static_cast<Simulator*>(s)->locations();
    for (size_t m = 0; m < l.size(); ++m) { xyz[3 * m] = l[m].x; xyz[3 * m + 1] = l[m].y; xyz[3 * m + 2] = l[m].z; }
  });
}
// the grid as flat arrays: returns the cell count; call with null pointers to size them
int splh_sim_snapshot(void* s, int cap, int* xyz, int* media, double* vel, double* p) {
  int count = -1;
  const int rc = guarded([&] {
    Grid g = static_cast<Simulator*>(s)->snapshot();
    count = (int)g.size();
    if (!xyz) return;
    int m = 0;
    for (const auto& kv : g.cells()) {
      if (m >= cap) break;
      xyz[3 * m] = kv.first.x; xyz[3 * m + 1] = kv.first.y; xyz[3 * m + 2] = kv.first.z;
      media[m] = (int)kv.second.media;
      vel[3 * m] = kv.second.velocity.x; vel[3 * m + 1] = kv.second.velocity.y;
      vel[3 * m + 2] = kv.second.velocity.z;
      p[m] = kv.second.pressure.value_or(0.0);
      ++m;
    }
  });
  return rc ? -1 : count;
}

}  // extern "C"
