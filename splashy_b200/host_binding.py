"""ctypes access to ``lib/libsplashy_host.so`` -- the C++ mirror of the reference's
Grid / Convection / Forces / Pressure modules (``splashy_b200/host``).  Used by the
tests so they read like the reference's NUnit fixture (Tests.fs:39-52)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Iterable, Optional, Sequence, Tuple

LIB = Path(__file__).resolve().parent / "lib" / "libsplashy_host.so"
AIR, FLUID, SOLID = 0, 1, 2
_lib = None


class HostFailure(RuntimeError):
    """The C++ host's ``splashy::Failure`` (the reference's ``failwith``)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        raise RuntimeError(f"{LIB} is missing: run __graft_entry__.build()")
    lib = C.CDLL(str(LIB))
    lib.splh_last_error.restype = C.c_char_p
    lib.splh_new.restype = C.c_void_p
    lib.splh_new.argtypes = [C.c_uint, C.c_double]
    lib.splh_free.argtypes = [C.c_void_p]
    lib.splh_nearest.argtypes = [C.c_int]
    lib.splh_construct_float.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_int)]
    lib.splh_add_cell.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int, C.c_double]
    lib.splh_get_cell.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_double),
                                                              C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.splh_size.argtypes = [C.c_void_p]
    lib.splh_save.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.splh_load.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.splh_update_velocity.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_double] * 3
    lib.splh_update_media.argtypes = [C.c_void_p] + [C.c_int] * 4
    lib.splh_pressure_divergence.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.POINTER(C.c_double)]
    for name in ("splh_convection", "splh_forces"):
        getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_double]
    lib.splh_pressure.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_double, C.POINTER(C.c_int)]
    lib.splh_checks.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    lib.splh_sim_new.restype = C.c_void_p
    lib.splh_sim_new.argtypes = [C.c_uint, C.c_double, C.c_int]
    lib.splh_sim_free.argtypes = [C.c_void_p]
    lib.splh_sim_generate.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    lib.splh_sim_set_velocity.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_double] * 3
    lib.splh_sim_advance.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.splh_sim_marker_count.argtypes = [C.c_void_p]
    lib.splh_sim_locations.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.splh_sim_snapshot.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = lib
    return lib


def nearest(n: int) -> int:
    return load().splh_nearest(n)


def construct_float(x: float, y: float, z: float) -> Tuple[int, int, int]:
    out = (C.c_int * 3)()
    load().splh_construct_float(x, y, z, out)
    return tuple(out)


class Host:
    """A Grid plus the GPU-backed solver terms (module-level state in the reference)."""

    def __init__(self, quirks: int = 0, tol: float = 1e-12):
        self._lib = load()
        self._h = C.c_void_p(self._lib.splh_new(quirks, tol))

    def close(self):
        if self._h:
            self._lib.splh_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ok(self, rc):
        if rc != 0:
            raise HostFailure(self._lib.splh_last_error().decode())

    @staticmethod
    def _markers(markers: Iterable[Sequence[int]]):
        flat = [int(v) for m in markers for v in m]
        return (C.c_int * len(flat))(*flat), len(flat) // 3

    # Grid.fs
    def add_cell(self, coord, media, velocity=(0.0, 0.0, 0.0), pressure: Optional[float] = None):
        self._ok(self._lib.splh_add_cell(self._h, *coord, media, *velocity,
                                         pressure is not None, pressure or 0.0))

    def get_cell(self, coord):
        media, has_p, p = C.c_int(), C.c_int(), C.c_double()
        vel = (C.c_double * 3)()
        self._ok(self._lib.splh_get_cell(self._h, *coord, C.byref(media), vel, C.byref(has_p),
                                         C.byref(p)))
        return media.value, tuple(vel), (p.value if has_p.value else None)

    def size(self) -> int:
        return self._lib.splh_size(self._h)

    # state files (splashy_host.hpp save_grid / save_box / save_box_npy, load_grid / load_box)
    def save(self, path, kind: str = "grid"):
        self._ok(self._lib.splh_save(self._h, {"grid": 0, "box": 1, "npy": 2}[kind],
                                     str(path).encode()))

    def load(self, path, kind: str = "grid"):
        self._ok(self._lib.splh_load(self._h, {"grid": 0, "box": 1}[kind], str(path).encode()))

    def update_velocity(self, coord, v):
        self._ok(self._lib.splh_update_velocity(self._h, *coord, *v))

    def update_media(self, coord, media):
        self._ok(self._lib.splh_update_media(self._h, *coord, media))

    # Pressure.fs / Convection.fs / Forces.fs  (Simulator.fs:82-97)
    def pressure_divergence(self, coord) -> float:
        out = C.c_double()
        self._ok(self._lib.splh_pressure_divergence(self._h, *coord, C.byref(out)))
        return out.value

    def convection(self, markers, dt):
        a, n = self._markers(markers)
        self._ok(self._lib.splh_convection(self._h, a, n, dt))

    def forces(self, markers, dt):
        a, n = self._markers(markers)
        self._ok(self._lib.splh_forces(self._h, a, n, dt))

    def pressure(self, markers, dt) -> int:
        a, n = self._markers(markers)
        it = C.c_int()
        self._ok(self._lib.splh_pressure(self._h, a, n, dt, C.byref(it)))
        return it.value

    def checks(self, markers):
        a, n = self._markers(markers)
        self._ok(self._lib.splh_checks(self._h, a, n))


class Simulator:
    """The C++ mirror of Simulator.fs (generate / advance) with the state kept on the GPU."""

    def __init__(self, quirks: int = 0, tol: float = 1e-12, margin: int = 4):
        self._lib = load()
        self._s = C.c_void_p(self._lib.splh_sim_new(quirks, tol, margin))

    def close(self):
        if self._s:
            self._lib.splh_sim_free(self._s)
            self._s = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ok(self, rc):
        if rc != 0:
            raise HostFailure(self._lib.splh_last_error().decode())

    def generate(self, marker_cells):
        a, n = Host._markers(marker_cells)
        self._ok(self._lib.splh_sim_generate(self._s, a, n))

    def set_velocity(self, coord, v):
        self._ok(self._lib.splh_sim_set_velocity(self._s, *coord, *v))

    def advance(self, dt: float, sanity: bool = True):
        self._ok(self._lib.splh_sim_advance(self._s, dt, 1 if sanity else 0))

    def locations(self):
        n = self._lib.splh_sim_marker_count(self._s)
        out = (C.c_double * (3 * n))()
        self._ok(self._lib.splh_sim_locations(self._s, out))
        return [tuple(out[3 * m:3 * m + 3]) for m in range(n)]

    def snapshot(self):
        """{(x, y, z): (media, (vx, vy, vz), pressure)} of every existing cell."""
        n = self._lib.splh_sim_snapshot(self._s, 0, None, None, None, None)
        if n < 0:
            raise HostFailure(self._lib.splh_last_error().decode())
        xyz, media = (C.c_int * (3 * n))(), (C.c_int * n)()
        vel, p = (C.c_double * (3 * n))(), (C.c_double * n)()
        got = self._lib.splh_sim_snapshot(self._s, n, xyz, media, vel, p)
        if got < 0:
            raise HostFailure(self._lib.splh_last_error().decode())
        return {tuple(xyz[3 * m:3 * m + 3]): (media[m], tuple(vel[3 * m:3 * m + 3]), p[m])
                for m in range(n)}
