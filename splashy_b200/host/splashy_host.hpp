// splashy_host.hpp -- C++ mirror of the reference's solver-facing API
// (src/splashy/{Coord,Cell,Grid,Convection,Forces,Pressure}.fs) on top of the C ABI
// of libsplashy_cuda.so.  The reference host is F# (compiled code) and no .NET
// toolchain exists in this image, so this is the compiled host side: same module
// and function names, argument meaning and error behaviour (std::runtime_error
// carrying the reference's failwith text), with the three hot functions executed
// on the GPU.  State stays in a sparse Dictionary-like Grid exactly as in
// Grid.fs:12; every GPU call packs it into the dense box.
#pragma once
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

struct spl_ctx;

namespace splashy {

// Constants.fs:10-36
namespace Constants {
constexpr int h_int = 10;
constexpr double h = 10.0;
constexpr double fluid_density = 999.97;
constexpr double atmospheric_pressure = 0.0;
constexpr double gravity[3] = {0.0, -9.81, 0.0};
}  // namespace Constants

struct Failure : std::runtime_error {      // the reference's failwith -> System.Exception
  using std::runtime_error::runtime_error;
};

// Vector.fs:9-34
struct Vector3d {
  double x = 0, y = 0, z = 0;
  Vector3d() = default;
  Vector3d(double x_, double y_, double z_);   // throws on NaN / +-Inf (Vector.fs:15-18)
};

enum class CoordDirection { NegX, NegY, NegZ, PosX, PosY, PosZ };   // Coord.fs:12

// Coord.fs:18-92
struct Coord {
  int x = 0, y = 0, z = 0;
  bool operator==(const Coord& o) const { return x == o.x && y == o.y && z == o.z; }
  static int nearest(int n);                               // Coord.fs:49-57
  static Coord construct(int x, int y, int z);             // Coord.fs:59-61
  static Coord construct(double x, double y, double z);    // Coord.fs:65-67
  std::vector<std::pair<CoordDirection, Coord>> backward_neighbors() const;
  std::vector<std::pair<CoordDirection, Coord>> forward_neighbors() const;
  std::vector<std::pair<CoordDirection, Coord>> neighbors() const;   // forward, then backward
  Coord get_neighbor(CoordDirection d) const;
};
struct CoordHash {                                          // Coord.fs:23-24
  size_t operator()(const Coord& c) const { return (size_t)(541 * c.x + 79 * c.y + 31 * c.z); }
};

enum class Media : uint8_t { Air = 0, Fluid = 1, Solid = 2 };   // Cell.fs:9

struct Cell {                                                // Cell.fs:11-16
  Media media = Media::Air;
  Vector3d velocity;                 // on the +faces
  std::optional<double> pressure;
  bool is_solid() const { return media == Media::Solid; }
  bool is_not_solid() const { return media != Media::Solid; }
};

// Grid.fs:12-59 (an instance instead of the reference's module-level mutable)
class Grid {
 public:
  std::optional<Cell> get(const Coord& where) const;
  const Cell& raw_get(const Coord& where) const;           // throws when absent
  std::vector<Coord> filter(const std::function<bool(const Cell&)>& fn) const;
  void add_cells(const std::vector<std::pair<Coord, Cell>>& cells);
  void delete_cells(const std::vector<Coord>& cells);
  void update_velocities(const std::vector<std::pair<Coord, Vector3d>>& v);
  void update_pressures(const std::vector<std::pair<Coord, double>>& p);
  void update_media(const std::vector<std::pair<Coord, Media>>& m);   // Tests.fs:69
  size_t size() const { return cells_.size(); }
  uint64_t version() const { return version_; }
  const std::unordered_map<Coord, Cell, CoordHash>& cells() const { return cells_; }

 private:
  void set(const Coord& where, const Cell& c);             // Grid.fs:33-36
  std::unordered_map<Coord, Cell, CoordHash> cells_;
  uint64_t version_ = 0;
};

// dense box packed from the grid
struct Box {
  int nx = 0, ny = 0, nz = 0, ox = 0, oy = 0, oz = 0;
  std::vector<uint8_t> media;
  std::vector<double> u, v, w, p;
  long index(const Coord& c) const;      // -1 when outside
};
Box pack(const Grid& g);
Grid unpack(const Box& b);             // absent cells dropped; pressure: Some p on non-Solid cells

// ---- state files: how a Grid / a dense Box leaves and enters the host ------
// One little-endian container, "SPLSTATE" + u32 version (1) + u32 kind:
//   kind 0, sparse grid   f64 h, u64 count, then per cell (sorted by z, y, x)
//                         i32 x, y, z (world coordinates, multiples of h), u8 media,
//                         u8 has_pressure, 2 pad bytes, f64 vx, vy, vz, pressure      (48 B)
//   kind 1, dense box     i32 nx, ny, nz, ox, oy, oz (origin in cells), f64 h, then
//                         u8 media[n] (255 = absent), f64 u[n], v[n], w[n], p[n], x fastest
// -- what Grid.fs:12's Dictionary<Coord,Cell> is written as by the F# shim of INTEGRATION.md
// (BinaryWriter) and read by splashy_b200/state_io.py, and the other way round.  save_box_npy
// writes the dense arrays as standard NumPy files (media.npy u8, u/v/w/p.npy f64, shape
// (nz, ny, nx)) plus meta.json (origin, h) into a directory.
void save_grid(const std::string& path, const Grid& g);
Grid load_grid(const std::string& path);
void save_box(const std::string& path, const Box& b);
Box load_box(const std::string& path);
void save_box_npy(const std::string& dir, const Box& b);

// The solver terms that run on the GPU.  One Solver = one spl_ctx (recreated when
// the packed box changes shape); fp64 like the reference.
class Solver {
 public:
  explicit Solver(unsigned quirks = 0, double tol = 1e-12, int device = 0);
  ~Solver();
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;

  // Convection.apply (Convection.fs:60-66): returned, not committed
  std::vector<std::pair<Coord, Vector3d>> convection_apply(const Grid& g,
                                                           const std::vector<Coord>& markers,
                                                           double dt);
  // Forces.apply (Forces.fs:6-12)
  std::vector<std::pair<Coord, Vector3d>> forces_apply(const Grid& g,
                                                       const std::vector<Coord>& markers,
                                                       double dt);
  // Pressure.divergence (Pressure.fs:16-32)
  double pressure_divergence(const Grid& g, const Coord& where);
  // Pressure.calculate (Pressure.fs:41-72)
  std::vector<std::pair<Coord, double>> pressure_calculate(const Grid& g,
                                                           const std::vector<Coord>& markers,
                                                           double dt);
  // Pressure.apply (Pressure.fs:115-129); must follow pressure_calculate on the same
  // grid state.  Each touched face is returned once.
  std::vector<std::pair<Coord, Vector3d>> pressure_apply(const Grid& g,
                                                         const std::vector<Coord>& markers,
                                                         double dt);
  // Pressure.check_pressures + check_divergence (Pressure.fs:132-149)
  void check(const Grid& g, const std::vector<Coord>& markers);
  int last_iterations() const { return last_iters_; }

 private:
  void sync(const Grid& g);          // pack + (re)create + upload
  void download();
  void fail_if(int rc);
  spl_ctx* ctx_ = nullptr;
  Box box_;
  unsigned quirks_;
  double tol_;
  int device_;
  int last_iters_ = 0;
  bool projected_ = false;
  uint64_t projected_version_ = 0;
};

// Simulator.fs:19-129 with the state resident on the device: `generate` packs the
// grid once (with a margin for the cells Build.setup creates), hands the marker
// locations over, and every `advance` is one spl_advance call -- move_markers,
// Build.setup block, convection, forces, viscosity, pressure (+ checks), cleanup.
// The grid is unpacked only on request (`snapshot`).
class Simulator {
 public:
  explicit Simulator(unsigned quirks = 0, double tol = 1e-12, int margin = 4, int device = 0);
  ~Simulator();
  Simulator(const Simulator&) = delete;
  Simulator& operator=(const Simulator&) = delete;

  // Simulator.generate (:113-129) for a given list of marker cells (the reference draws
  // them from System.Random(12349)): markers = sorted distinct cells, locations = their
  // centres, grid = the markers as new Fluid cells (Build.add_new_markers)
  void generate(const std::vector<Coord>& marker_cells);
  // overwrite a cell's velocity before the run (test scenes)
  void set_velocity(const Coord& where, const Vector3d& v);
  // Simulator.advance dt sanity (:54-110)
  void advance(double dt, bool sanity);
  const std::vector<Vector3d>& locations() const { return locations_; }   // metres
  std::vector<Coord> markers() const;                                     // closest cells
  Grid snapshot();                     // device state -> sparse grid (absent cells dropped)
  int last_iterations() const { return last_iters_; }

 private:
  void fail_if(int rc);
  void ensure_uploaded();
  spl_ctx* ctx_ = nullptr;
  Grid grid_;                          // host copy, valid until the first advance
  Box box_;
  std::vector<Vector3d> locations_;
  unsigned quirks_;
  double tol_;
  int margin_, device_;
  int last_iters_ = 0;
  bool uploaded_ = false;
};

}  // namespace splashy
