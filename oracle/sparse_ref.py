"""Faithful CPU restatement of the reference's per-timestep path (fp64, sparse).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED: the
F# reference cannot run here and ships no golden vectors; this file is pinned
only by the reference's three valid divergence KATs and SURVEY.md App. B.

Every function cites the reference file:line it follows (paths relative to
``/root/reference/src/splashy``).  Nothing is "fixed": the quirk register
Q1-Q11 of SURVEY.md is reproduced on purpose, including

* the lazy ``Seq.map`` + in-place ``Seq.iter`` commit order (Q2) -- Python
  generators consumed by ``Grid.update_*`` give exactly that order,
* ``Coord.nearest`` with .NET remainder semantics (Q6),
* the int-overload ``Coord.construct`` used on *index* values inside
  ``Convection.interpolate`` (Q3),
* the forward trace and nearest-cell copy (Q4, Q5),
* the double gradient hit on fluid-fluid faces (Q1) and the solid ghost
  pressure sign (Q7).

The linear solve is MathNet.Numerics 3.11.0 ``Matrix.Solve`` (paket.lock:19-23),
an exact LU solve that is not vendored; it is restated with
``numpy.linalg.solve`` (same algorithm class, fp64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Constants.fs:10-36
# --------------------------------------------------------------------------
H = 10.0                      # Constants.fs:11   h = 10 m
H_INT = 10
FLUID_VISCOSITY = 0.000001    # Constants.fs:13
FLUID_DENSITY = 999.97        # Constants.fs:15
ATMOSPHERIC_PRESSURE = 0.0    # Constants.fs:20
MAX_VELOCITY = 2.5            # Constants.fs:22
TIME_STEP_CONSTANT = 2.0      # Constants.fs:24
GRAVITY = (0.0, -9.81, 0.0)   # Constants.fs:26
WORLD_H = float(np.float32(100.0) + np.float32(H) / np.float32(2.0))  # :32
N_MARKERS = 1                 # Constants.fs:36

AIR, FLUID, SOLID = 0, 1, 2   # Cell.fs:9   Media = Air | Fluid | Solid

NEG_X, NEG_Y, NEG_Z, POS_X, POS_Y, POS_Z = range(6)  # Coord.fs:12

Coord = Tuple[int, int, int]
Vec = Tuple[float, float, float]


class ReferenceFailure(Exception):
    """Stands for the reference's ``failwith`` (System.Exception)."""


# --------------------------------------------------------------------------
# Vector.fs
# --------------------------------------------------------------------------
def vec(x: float, y: float, z: float) -> Vec:
    """Vector.fs:15-18 -- the constructor throws on NaN / +-Inf."""
    if not (math.isfinite(x) and math.isfinite(y) and math.isfinite(z)):
        raise ReferenceFailure(
            "Vector3d [%r; %r; %r] has an invalid quantity: either NaN or "
            "±Infinity." % (x, y, z))
    return (x, y, z)


ZERO: Vec = (0.0, 0.0, 0.0)


def vadd(a: Vec, b: Vec) -> Vec:            # Vector.fs:26
    return vec(a[0] + b[0], a[1] + b[1], a[2] + b[2])


def vsub(a: Vec, b: Vec) -> Vec:            # Vector.fs:28
    return vec(a[0] - b[0], a[1] - b[1], a[2] - b[2])


def vmul(a: Vec, f: float) -> Vec:          # Vector.fs:30
    return vec(a[0] * f, a[1] * f, a[2] * f)


def vsum(vs: Iterable[Vec]) -> Vec:         # Vector.fs:46
    acc = ZERO
    for v in vs:
        acc = vadd(acc, v)
    return acc


# --------------------------------------------------------------------------
# Coord.fs
# --------------------------------------------------------------------------
def _rem(n: int, d: int) -> int:
    """.NET ``%`` keeps the sign of the dividend (truncated division)."""
    return int(math.fmod(n, d))


def nearest(n: int) -> int:
    """Coord.fs:49-57."""
    h = H_INT
    r = _rem(n, h)
    if r == 0:
        return n
    if r >= h // 2:
        return n + (h - r)
    return n - r


def construct_int(x: int, y: int, z: int) -> Coord:
    """Coord.fs:59-61."""
    return (nearest(x), nearest(y), nearest(z))


def _round_half_even(x: float) -> float:
    """F# ``round`` = Math.Round = banker's rounding; Python's round() too."""
    return float(round(x))


def construct_float(x: float, y: float, z: float) -> Coord:
    """Coord.fs:65-67: float -> round -> int (truncate) -> nearest."""
    def to_nearest_int(v: float) -> int:
        return nearest(int(_round_half_even(v)))
    return (to_nearest_int(x), to_nearest_int(y), to_nearest_int(z))


def backward_neighbors(c: Coord) -> List[Tuple[int, Coord]]:
    """Coord.fs:69-73."""
    h = H_INT
    return [(NEG_X, (c[0] - h, c[1], c[2])),
            (NEG_Y, (c[0], c[1] - h, c[2])),
            (NEG_Z, (c[0], c[1], c[2] - h))]


def forward_neighbors(c: Coord) -> List[Tuple[int, Coord]]:
    """Coord.fs:75-79."""
    h = H_INT
    return [(POS_X, (c[0] + h, c[1], c[2])),
            (POS_Y, (c[0], c[1] + h, c[2])),
            (POS_Z, (c[0], c[1], c[2] + h))]


def neighbors(c: Coord) -> List[Tuple[int, Coord]]:
    """Coord.fs:81-82: forward first, then backward."""
    return forward_neighbors(c) + backward_neighbors(c)


def get_neighbor(c: Coord, d: int) -> Coord:
    """Coord.fs:84-92."""
    h = H_INT
    return {NEG_X: (c[0] - h, c[1], c[2]), POS_X: (c[0] + h, c[1], c[2]),
            NEG_Y: (c[0], c[1] - h, c[2]), POS_Y: (c[0], c[1] + h, c[2]),
            NEG_Z: (c[0], c[1], c[2] - h), POS_Z: (c[0], c[1], c[2] + h)}[d]


def reverse(d: int) -> int:
    """Coord.fs:94-101."""
    return {NEG_X: POS_X, POS_X: NEG_X, NEG_Y: POS_Y, POS_Y: NEG_Y,
            NEG_Z: POS_Z, POS_Z: NEG_Z}[d]


def _axis(d: int) -> int:
    return d % 3


def is_bordering(d: int, v: Vec) -> bool:
    """Coord.fs:103-111."""
    c = v[_axis(d)]
    return c < 0.0 if d < 3 else c > 0.0


def border(d: int, v: Vec) -> Vec:
    """Coord.fs:113-117."""
    a = _axis(d)
    out = [0.0, 0.0, 0.0]
    out[a] = v[a]
    return vec(*out)


def border_value(d: int, v: Vec) -> float:
    """Coord.fs:119-123."""
    return v[_axis(d)]


def merge(d: int, old_v: Vec, new_v: Vec) -> Vec:
    """Coord.fs:125-129."""
    a = _axis(d)
    out = list(old_v)
    out[a] = new_v[a]
    return vec(*out)


# --------------------------------------------------------------------------
# Cell.fs:11-16
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Cell:
    media: int
    velocity: Vec                 # on the +faces
    pressure: Optional[float]

    def is_solid(self) -> bool:
        return self.media == SOLID

    def is_not_solid(self) -> bool:
        return self.media != SOLID


# --------------------------------------------------------------------------
# World.fs:11-14 / Aabb.fs:13-17
# --------------------------------------------------------------------------
def world_contains(c: Coord) -> bool:
    w = np.float32(WORLD_H)
    p = [np.float32(c[0]), np.float32(c[1]), np.float32(c[2])]
    return all((-w <= q) and (q <= w) for q in p)


# --------------------------------------------------------------------------
# Grid.fs
# --------------------------------------------------------------------------
class Grid:
    """Grid.fs:12-59 -- the global mutable ``Dictionary<Coord, Cell>``.

    Python dicts iterate in insertion order; .NET's Dictionary iterates in
    bucket/entry order, which equals insertion order as long as nothing was
    removed.  No result on the hot path depends on that order except through
    the marker list, which the caller owns.
    """

    def __init__(self) -> None:
        self.cells: Dict[Coord, Cell] = {}

    def get(self, where: Coord) -> Optional[Cell]:          # Grid.fs:14-17
        return self.cells.get(where)

    def raw_get(self, where: Coord) -> Cell:                # Grid.fs:20
        c = self.cells.get(where)
        if c is None:
            raise ReferenceFailure("Option.get on None at %r" % (where,))
        return c

    def filter(self, fn) -> List[Coord]:                    # Grid.fs:22-27
        return [k for k, v in self.cells.items() if fn(v)]

    def _set(self, where: Coord, c: Cell) -> None:          # Grid.fs:33-36
        if where not in self.cells:
            raise ReferenceFailure("Tried to set non-existent cell.")
        self.cells[where] = c

    def add_cells(self, cells: Iterable[Tuple[Coord, Cell]]) -> None:  # :38
        for where, c in cells:
            if where in self.cells:
                raise ReferenceFailure("duplicate key %r" % (where,))
            self.cells[where] = c

    def delete_cells(self, cells: Iterable[Coord]) -> None:  # Grid.fs:40
        for where in cells:
            self.cells.pop(where, None)

    def move_cells(self, changes: Iterable[Tuple[Coord, Coord]]) -> None:
        """Grid.fs:42-47."""
        for old, new in changes:
            c = self.raw_get(old)
            self.cells.pop(new, None)
            self.cells[new] = c
            self.cells.pop(old, None)

    def update_velocities(self, velocities: Iterable[Tuple[Coord, Vec]]) -> None:
        """Grid.fs:49-53 -- ``Seq.iter`` over a lazy sequence: each element is
        produced, then committed, before the next one is produced (Q2)."""
        for where, new_v in velocities:
            c = self.raw_get(where)
            self._set(where, replace(c, velocity=new_v))

    def update_pressures(self, pressures: Iterable[Tuple[Coord, float]]) -> None:
        """Grid.fs:55-59."""
        for where, new_p in pressures:
            c = self.raw_get(where)
            self._set(where, replace(c, pressure=new_p))

    def update_media(self, media: Iterable[Tuple[Coord, int]]) -> None:
        """Not in the current Grid.fs; the (stale) test project calls it
        (Tests.fs:69,93,106).  Provided so KAT-3 can be expressed."""
        for where, m in media:
            c = self.raw_get(where)
            self._set(where, replace(c, media=m))


# --------------------------------------------------------------------------
# Convection.fs
# --------------------------------------------------------------------------
def get_velocity_index(g: Grid, where: Coord, index: int) -> Optional[float]:
    """Convection.fs:11-19."""
    c = g.get(where)
    if c is None:
        return None
    if index not in (0, 1, 2):
        raise ReferenceFailure("No such velocity index.")
    return c.velocity[index]


def interpolate(g: Grid, x: float, y: float, z: float, index: int) -> float:
    """Convection.fs:21-39.  Inputs are in index units (position / h); the
    corner key goes through the *int* overload of Coord.construct (Q3)."""
    i, j, k = math.floor(x), math.floor(y), math.floor(z)
    ii, jj, kk = int(i), int(j), int(k)
    idw = [i - x + 1.0, x - i]
    jdw = [j - y + 1.0, y - j]
    kdw = [k - z + 1.0, z - k]
    total = 0.0
    for xp in (0, 1):
        for yp in (0, 1):
            for zp in (0, 1):
                c = construct_int(ii + xp, jj + yp, kk + zp)
                v = get_velocity_index(g, c, index)
                if v is not None:
                    total += v * idw[xp] * jdw[yp] * kdw[zp]
                else:
                    total += 0.0
    return total


def get_interpolated_velocity(g: Grid, x: float, y: float, z: float) -> Vec:
    """Convection.fs:41-48."""
    xh, yh, zh = x / H, y / H, z / H
    vx = interpolate(g, xh, yh - 0.5, zh - 0.5, 0)
    vy = interpolate(g, xh - 0.5, yh, zh - 0.5, 1)
    vz = interpolate(g, xh - 0.5, yh - 0.5, zh, 2)
    return vec(vx, vy, vz)


def trace(g: Grid, c: Coord, t: float) -> Coord:
    """Convection.fs:50-57 -- RK2, forward sign (Q4), snapped result (Q5)."""
    cv = (float(c[0]), float(c[1]), float(c[2]))
    v = get_interpolated_velocity(g, *cv)
    nv = vadd(cv, vmul(v, t / 2.0))
    dv = get_interpolated_velocity(g, *nv)
    p = vadd(cv, vmul(dv, t))
    return construct_float(*p)


def convection_apply(g: Grid, markers: Iterable[Coord], dt: float
                     ) -> Iterator[Tuple[Coord, Vec]]:
    """Convection.fs:60-66 (lazy)."""
    for m in markers:
        new_position = trace(g, m, dt)
        c = g.get(new_position)
        if c is None:
            raise ReferenceFailure("Backwards particle trace went too far.")
        yield (m, c.velocity)


# --------------------------------------------------------------------------
# Forces.fs:6-12, Viscosity.fs:12-36
# --------------------------------------------------------------------------
def forces_apply(g: Grid, markers: Iterable[Coord], dt: float
                 ) -> Iterator[Tuple[Coord, Vec]]:
    f = vmul(GRAVITY, dt)
    for m in markers:
        c = g.raw_get(m)
        yield (m, vadd(c.velocity, f))


def get_shared_velocity(g: Grid, d: int, n: Coord) -> Vec:
    """Viscosity.fs:12-17."""
    c = g.get(n)
    if c is not None and c.media == FLUID and is_bordering(d, c.velocity):
        return border(d, c.velocity)
    return ZERO


def viscosity_laplacian(g: Grid, where: Coord) -> Vec:
    """Viscosity.fs:19-26 (not divided by h^2, Q10)."""
    c = g.get(where)
    where_v = vmul(c.velocity, 6.0) if c is not None else ZERO
    v = vsum(get_shared_velocity(g, d, n) for d, n in neighbors(where))
    return vmul(vsub(v, where_v), 1.0)


def viscosity_apply(g: Grid, markers: Iterable[Coord], dt: float
                    ) -> Iterator[Tuple[Coord, Vec]]:
    """Viscosity.fs:29-36."""
    const_v = FLUID_VISCOSITY * dt
    for m in markers:
        result = viscosity_laplacian(g, m)
        c = g.raw_get(m)
        yield (m, vadd(c.velocity, vmul(result, const_v)))


# --------------------------------------------------------------------------
# Pressure.fs
# --------------------------------------------------------------------------
def _is_nonsolid(g: Grid, n: Coord) -> bool:
    """Pressure.fs:14 (raw_get: throws when the neighbour is absent)."""
    return g.raw_get(n).is_not_solid()


def divergence(g: Grid, where: Coord) -> float:
    """Pressure.fs:16-32: (incoming - outgoing), x + y + z, not / h (Q9)."""
    cell = g.raw_get(where)
    outgoing = ZERO
    for d, n in forward_neighbors(where):
        if _is_nonsolid(g, n):
            outgoing = vadd(outgoing, border(d, cell.velocity))
    incoming = ZERO
    for d, n in backward_neighbors(where):
        if _is_nonsolid(g, n):
            incoming = vadd(incoming, border(d, g.raw_get(n).velocity))
    r = vsub(incoming, outgoing)
    return r[0] + r[1] + r[2]


def number_neighbors(g: Grid, fn, where: Coord) -> float:
    """Pressure.fs:34-38: absent neighbours do not count."""
    cnt = 0
    for _, n in neighbors(where):
        c = g.get(n)
        if c is not None and fn(c):
            cnt += 1
    return float(cnt)


def assemble(g: Grid, markers: List[Coord], dt: float
             ) -> Tuple[np.ndarray, np.ndarray]:
    """Pressure.fs:42-68: dense copy of the SparseMatrix and the RHS."""
    n = len(markers)
    lookups = {m: i for i, m in enumerate(markers)}          # :42-43
    a = np.zeros((n, n), dtype=np.float64)                   # :57
    for marker, c in lookups.items():                        # :58-60
        coeffs = [(lookups[marker],
                   number_neighbors(g, Cell.is_not_solid, marker))]
        singulars: List[Tuple[int, float]] = []
        for _, nb in neighbors(marker):
            if nb in lookups:
                singulars.insert(0, (lookups[nb], -1.0))
        for r, value in coeffs + singulars:
            a[r, c] = value
    if not np.array_equal(a, a.T):                           # :61-62
        raise ReferenceFailure("Coefficient matrix is not symmetric.")
    cst = (H * FLUID_DENSITY) / dt                           # :64
    b = np.array([cst * divergence(g, m) for m in markers],  # :65-68
                 dtype=np.float64)
    return a, b


def pressure_calculate(g: Grid, markers: List[Coord], dt: float
                       ) -> List[Tuple[Coord, float]]:
    """Pressure.fs:41-72.  ``m.Solve(b)`` (MathNet LU) -> numpy LU solve."""
    a, b = assemble(g, markers, dt)
    if len(markers) == 0:
        return []
    pressures = np.linalg.solve(a, b)                        # :69
    return list(zip(markers, (float(p) for p in pressures)))


def get_solid_pressure(g: Grid, solid_cell: Cell, origin: Coord, p: float,
                       inv_c: float, d: int) -> float:
    """Pressure.fs:75-85 (Q7 sign kept)."""
    if solid_cell.media != SOLID:
        raise ReferenceFailure("not a solid!")
    vsolid = solid_cell.velocity
    vborder = g.raw_get(origin).velocity
    if d in (NEG_X, NEG_Y, NEG_Z):
        diff = border_value(d, vborder) - border_value(d, vsolid)
    else:
        diff = border_value(d, vsolid) - border_value(d, vborder)
    return p + (diff / inv_c)


def get_pressure(g: Grid, origin: Coord, p: float, inv_c: float, d: int,
                 where: Coord) -> float:
    """Pressure.fs:87-94."""
    c = g.get(where)
    if c is not None and c.is_solid():
        return get_solid_pressure(g, c, origin, p, inv_c, d)
    if c is not None:
        if c.pressure is None:
            raise ReferenceFailure("Option.get on None pressure")
        return c.pressure
    raise ReferenceFailure("no pressure found.")


def forwards_gradient(g: Grid, where: Coord, p: float, inv_c: float) -> Vec:
    """Pressure.fs:96-100."""
    p1 = get_pressure(g, where, p, inv_c, POS_X, get_neighbor(where, POS_X))
    p2 = get_pressure(g, where, p, inv_c, POS_Y, get_neighbor(where, POS_Y))
    p3 = get_pressure(g, where, p, inv_c, POS_Z, get_neighbor(where, POS_Z))
    return vec(p1 - p, p2 - p, p3 - p)


def get_adjusted_velocity(g: Grid, where: Coord, n: Coord, p: float,
                          inv_c: float, d: int) -> Tuple[Coord, Vec]:
    """Pressure.fs:102-112."""
    c = g.raw_get(n)
    pressure = get_pressure(g, where, p, inv_c, d, n)
    offset = (p - pressure) * inv_c
    if d == NEG_X:
        vel = vec(offset, 0.0, 0.0)
    elif d == NEG_Y:
        vel = vec(0.0, offset, 0.0)
    elif d == NEG_Z:
        vel = vec(0.0, 0.0, offset)
    else:
        raise ReferenceFailure("Invalid direction.")
    return (n, vsub(c.velocity, vel))


def pressure_apply(g: Grid, markers: Iterable[Coord], dt: float
                   ) -> Iterator[Tuple[Coord, Vec]]:
    """Pressure.fs:115-129 (lazy; Q1 double hit arises at commit time)."""
    inv_c = dt / (H * FLUID_DENSITY)
    for m in markers:
        c = g.raw_get(m)
        if c.pressure is None:
            raise ReferenceFailure("Option.get on None pressure")
        p = c.pressure
        gradient = forwards_gradient(g, m, p, inv_c)
        v = vsub(c.velocity, vmul(gradient, inv_c))
        nonsolids = [(d, n) for d, n in backward_neighbors(m)
                     if _is_nonsolid(g, n)]
        nvs = [get_adjusted_velocity(g, m, n, p, inv_c, d)
               for d, n in nonsolids]
        for item in [(m, v)] + nvs:
            yield item


def check_pressures(g: Grid, markers: Iterable[Coord]) -> None:
    """Pressure.fs:132-142."""
    for m in markers:
        p = g.raw_get(m).pressure
        if p is None:
            raise ReferenceFailure("Option.get on None pressure")
        if math.isnan(p):
            raise ReferenceFailure("Invalid pressure.")
        if p == -math.inf:
            raise ReferenceFailure("Pressure is negative infinity.")
        if p == math.inf:
            raise ReferenceFailure("Pressure is positive infinity.")


def check_divergence(g: Grid, markers: Iterable[Coord]) -> None:
    """Pressure.fs:145-149."""
    for m in markers:
        f = divergence(g, m)
        if abs(f) > 0.0001:
            raise ReferenceFailure(
                "Velocity field is not correct: %r has divergence %r." % (m, f))


# --------------------------------------------------------------------------
# Build.fs (scene upkeep -- only what is needed to construct / finish a step)
# --------------------------------------------------------------------------
MAX_DISTANCE = max(2, int(math.ceil(TIME_STEP_CONSTANT)))     # Build.fs:34


class Build:
    """Build.fs:11-174 with its private ``Layers`` dictionary."""

    def __init__(self, g: Grid) -> None:
        self.g = g
        self.layers: Dict[Coord, Optional[int]] = {}

    def setup(self) -> None:                                  # Build.fs:37-41
        self.layers.clear()
        for m in self.g.filter(lambda _: True):
            self.layers[m] = None

    def add_new_markers(self, markers: Iterable[Coord]
                        ) -> List[Tuple[Coord, Cell]]:        # Build.fs:44-54
        acc: List[Tuple[Coord, Cell]] = []
        for marker in markers:
            if self.g.get(marker) is None and world_contains(marker):
                acc.insert(0, (marker, Cell(FLUID, ZERO, None)))
        return acc

    def set_fluid_layers(self, markers: Iterable[Coord]) -> None:  # :56-62
        for marker in markers:
            if self.g.get(marker) is None:
                raise ReferenceFailure("Fluid does not have corresponding cell.")
            self.layers[marker] = 0

    def cleanup(self) -> None:                                # Build.fs:65-71
        for m in self.g.filter(lambda _: True):
            c = self.g.raw_get(m)
            self.layers[m] = 0 if c.media == FLUID else None

    def delete_unused(self) -> List[Coord]:                   # Build.fs:73-76
        leftover = [k for k, v in self.layers.items() if v is None]
        for k in leftover:
            del self.layers[k]
        return leftover

    def create_air_buffer(self) -> List[Tuple[Coord, Cell]]:  # Build.fs:79-102
        additions: Dict[Coord, Cell] = {}
        for i in range(1, MAX_DISTANCE + 1):
            keys = self.g.filter(Cell.is_not_solid)           # eager List copy
            for c in keys:                                    # lazy filter
                if self.layers[c] != (i - 1) or self.layers[c] is None:
                    continue
                for _, where in neighbors(c):
                    cell = self.g.get(where)
                    if cell is not None:
                        if cell.media != FLUID and self.layers[where] is None:
                            self.layers[where] = i
                    else:
                        if world_contains(where):
                            new_cell = Cell(AIR, ZERO, ATMOSPHERIC_PRESSURE)
                        else:
                            new_cell = Cell(SOLID, ZERO, None)
                        additions[where] = new_cell
                        self.layers[where] = i
        return list(additions.items())

    def propagate_velocities(self) -> List[Tuple[Coord, Vec]]:  # :105-141
        g = self.g
        result: List[Tuple[Coord, Vec]] = []
        for i in range(1, MAX_DISTANCE + 1):
            nonfluid = [k for k, v in self.layers.items() if v is None]

            def prev_layer(nb: Coord) -> bool:
                return g.get(nb) is not None and self.layers.get(nb, "x") == (i - 1)

            for m in nonfluid:
                fwd = [(d, n) for d, n in forward_neighbors(m) if prev_layer(n)]
                bwd = [(d, n) for d, n in backward_neighbors(m) if prev_layer(n)]
                if len(fwd) + len(bwd) == 0:
                    continue
                c = g.raw_get(m)
                f = ZERO
                for d, _ in fwd:
                    f = vadd(f, border(d, c.velocity))
                nsum = ZERO
                for d, where in bwd:
                    nsum = vadd(nsum, border(d, g.raw_get(where).velocity))
                avg = vadd(f, nsum)
                dirs = [NEG_X if c.velocity[0] < 0.0 else POS_X,
                        NEG_Y if c.velocity[1] < 0.0 else POS_Y,
                        NEG_Z if c.velocity[2] < 0.0 else POS_Z]
                new_vs = []
                for d in dirs:
                    nb = g.get(get_neighbor(m, d))
                    if nb is not None and nb.media != FLUID:
                        new_vs.append(border_value(d, merge(d, c.velocity, avg)))
                    else:
                        new_vs.append(border_value(d, c.velocity))
                new_v = vec(*new_vs)
                if new_v != c.velocity:
                    result.insert(0, (m, new_v))
                self.layers[m] = i - 1
        return result

    def zero_solid_velocities(self) -> Iterator[Tuple[Coord, Vec]]:  # :144-160
        g = self.g
        for solid in g.filter(Cell.is_solid):
            for d, nb in backward_neighbors(solid):
                inward = reverse(d)
                n = g.get(nb)
                if n is not None and n.is_not_solid() and \
                        is_bordering(inward, n.velocity):
                    n2 = g.raw_get(nb)
                    yield (nb, merge(inward, n2.velocity, ZERO))

    def check_containment(self, markers: Iterable[Coord]) -> None:  # :162-166
        for m in markers:
            if not world_contains(m):
                raise ReferenceFailure(
                    "Error: Fluid marker went outside bounds (example: %r)." % (m,))

    def check_surroundings(self, markers: Iterable[Coord]) -> None:  # :168-174
        for m in markers:
            for _, nb in neighbors(m):
                if self.g.get(nb) is None:
                    raise ReferenceFailure(
                        "Fluid marker %r does not have a neighbor." % (m,))


# --------------------------------------------------------------------------
# Simulator.fs
# --------------------------------------------------------------------------
class Simulator:
    """Simulator.fs:19-129 -- markers, locations and the ``advance`` sequencer."""

    def __init__(self) -> None:
        self.grid = Grid()
        self.build = Build(self.grid)
        self.markers: List[Coord] = []
        self.locations: List[Vec] = []

    def get_velocity(self, m: Coord) -> Vec:                  # Simulator.fs:24-35
        g = self.grid
        c = g.raw_get(m)
        v = ZERO
        for d, n in backward_neighbors(m):
            nc = g.get(n)
            if nc is not None:
                v = vadd(v, border(d, nc.velocity))
        return vadd(v, c.velocity)

    def move_markers(self, dt: float) -> List[Tuple[Coord, Coord]]:  # :38-52
        new_locations = []
        for l in self.locations:
            m = construct_float(*l)
            v = self.get_velocity(m)
            new_locations.append(vadd(l, vmul(v, dt)))
        self.locations = new_locations
        moved: List[Tuple[Coord, Coord]] = []
        for m, l in zip(self.markers, self.locations):
            c = construct_float(*l)
            if m != c:
                moved.insert(0, (m, c))
        self.markers = [construct_float(*l) for l in self.locations]
        return moved

    def set_markers(self, markers: Iterable[Coord]) -> None:
        """Simulator.fs:124-128 -- the tail of ``generate`` for a given marker
        list (``Set.ofList`` sorts and de-duplicates with Coord.compare)."""
        self.markers = sorted(set(markers))
        self.locations = [(float(m[0]), float(m[1]), float(m[2]))
                          for m in self.markers]
        self.build.setup()
        self.grid.add_cells(self.build.add_new_markers(self.markers))

    def generate(self, n: int, marker_cells: Optional[List[Coord]] = None) -> None:
        """Simulator.fs:113-129.  .NET ``System.Random(12349)`` is not restated
        (Knuth subtractive generator); any in-world cell is an equivalent scene
        (SURVEY 8c), so callers pass the cells, default = the origin."""
        if marker_cells is None:
            marker_cells = [(0, 0, 0)][:n]
        self.set_markers(construct_int(*m) for m in marker_cells)

    # -- stages of ``advance`` exposed one by one for stage-level parity -----
    def stage_setup(self, sanity: bool = True) -> None:       # :60-79
        b = self.build
        if sanity:
            b.check_containment(self.markers)
        b.setup()
        self.grid.add_cells(b.add_new_markers(self.markers))
        b.set_fluid_layers(self.markers)
        self.grid.add_cells(b.create_air_buffer())
        self.grid.delete_cells(b.delete_unused())
        if sanity:
            b.check_surroundings(self.markers)

    def stage_convection(self, dt: float) -> None:            # :82
        self.grid.update_velocities(convection_apply(self.grid, self.markers, dt))

    def stage_forces(self, dt: float) -> None:                # :84
        self.grid.update_velocities(forces_apply(self.grid, self.markers, dt))

    def stage_viscosity(self, dt: float) -> None:             # :86
        self.grid.update_velocities(viscosity_apply(self.grid, self.markers, dt))

    def stage_pressure(self, dt: float) -> None:              # :88-89
        self.grid.update_pressures(
            pressure_calculate(self.grid, self.markers, dt))
        self.grid.update_velocities(pressure_apply(self.grid, self.markers, dt))

    def stage_checks(self) -> None:                           # :91-97
        check_pressures(self.grid, self.markers)
        check_divergence(self.grid, self.markers)

    def stage_cleanup(self) -> None:                          # :100-110
        b = self.build
        b.cleanup()
        self.grid.update_velocities(b.propagate_velocities())
        self.grid.update_velocities(b.zero_solid_velocities())
        self.grid.update_velocities(
            (s, ZERO) for s in self.grid.filter(Cell.is_solid))

    def advance(self, dt: float, sanity: bool = True) -> None:
        """Simulator.fs:54-110."""
        self.grid.move_cells(self.move_markers(dt))           # :60
        self.stage_setup(sanity)
        self.stage_convection(dt)
        self.stage_forces(dt)
        self.stage_viscosity(dt)
        self.stage_pressure(dt)
        if sanity:
            self.stage_checks()
        self.stage_cleanup()
