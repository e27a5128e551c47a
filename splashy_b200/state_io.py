"""State files of the splashy host: how a sparse grid (``Grid.fs:12``'s
``Dictionary<Coord, Cell>``) or a dense box leaves and enters a host process.

One little-endian container, written and read by the C++ host (``splashy_host.hpp``
``save_grid / load_grid / save_box / load_box``), by the F# shim of INTEGRATION.md
(``BinaryWriter``) and by this module:

    "SPLSTATE", u32 version = 1, u32 kind
    kind 0  sparse grid   f64 h, u64 count, then per cell (sorted by z, y, x): i32 x, y, z (world
                          coordinates, multiples of h), u8 media, u8 has_pressure, 2 pad bytes,
                          f64 vx, vy, vz, pressure                                     (48 bytes)
    kind 1  dense box     i32 nx, ny, nz, ox, oy, oz (origin in cells), f64 h, then u8 media[n]
                          (255 = absent), f64 u[n], v[n], w[n], p[n], x fastest

A dense box is also written as a directory of standard NumPy files (``save_box_npy``:
media.npy u8, u/v/w/p.npy f64, shape (nz, ny, nx), meta.json with origin and h).
"""
from __future__ import annotations

import json
import struct
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

MAGIC = b"SPLSTATE"
ABSENT = 255
Cells = Dict[Tuple[int, int, int], Tuple[int, Tuple[float, float, float], Optional[float]]]
_REC = np.dtype([("xyz", "<i4", 3), ("media", "u1"), ("has_p", "u1"), ("pad", "u1", 2),
                 ("val", "<f8", 4)])
assert _REC.itemsize == 48


def _header(f, kind: int) -> None:
    if f.read(8) != MAGIC:
        raise ValueError("not a SPLSTATE file")
    version, k = struct.unpack("<II", f.read(8))
    if version != 1:
        raise ValueError(f"unsupported SPLSTATE version {version}")
    if k != kind:
        raise ValueError("not a dense box" if kind else "not a sparse grid")


def save_grid(path, cells: Cells, h: float = 10.0) -> None:
    """cells: {(x, y, z) world coordinates: (media, (vx, vy, vz), pressure or None)}."""
    keys = sorted(cells, key=lambda c: (c[2], c[1], c[0]))
    rec = np.zeros(len(keys), dtype=_REC)
    for n, key in enumerate(keys):
        media, vel, p = cells[key]
        rec[n] = (key, media, 0 if p is None else 1, (0, 0), (*vel, 0.0 if p is None else p))
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", 1, 0) + struct.pack("<dQ", float(h), len(keys)))
        f.write(rec.tobytes())


def load_grid(path) -> Tuple[Cells, float]:
    with open(path, "rb") as f:
        _header(f, 0)
        h, n = struct.unpack("<dQ", f.read(16))
        rec = np.frombuffer(f.read(n * _REC.itemsize), dtype=_REC)
    if len(rec) != n:
        raise ValueError("state file is truncated")
    cells: Cells = {}
    for r in rec:
        v = r["val"]
        cells[tuple(int(c) for c in r["xyz"])] = (
            int(r["media"]), (float(v[0]), float(v[1]), float(v[2])),
            float(v[3]) if r["has_p"] else None)
    return cells, h


def save_box(path, labels, U, V, W, P, origin=(0, 0, 0), h: float = 10.0) -> None:
    """Dense arrays of shape (nz, ny, nx); fields are stored as f64."""
    nz, ny, nx = labels.shape
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", 1, 1) + struct.pack("<6id", nx, ny, nz, *origin, float(h)))
        f.write(np.ascontiguousarray(labels, dtype=np.uint8).tobytes())
        for a in (U, V, W, P):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())


def load_box(path) -> Dict:
    with open(path, "rb") as f:
        _header(f, 1)
        nx, ny, nz, ox, oy, oz, h = struct.unpack("<6id", f.read(32))
        n = nx * ny * nz
        labels = np.frombuffer(f.read(n), dtype=np.uint8)
        fields = [np.frombuffer(f.read(8 * n), dtype="<f8") for _ in range(4)]
    if len(labels) != n or any(len(a) != n for a in fields):
        raise ValueError("state file is truncated")
    out = {"labels": labels.reshape(nz, ny, nx).copy(), "origin": (ox, oy, oz), "h": h}
    for k, a in zip("UVWP", fields):
        out[k] = a.reshape(nz, ny, nx).copy()
    return out


def load_box_npy(directory) -> Dict:
    d = Path(directory)
    meta = json.loads((d / "meta.json").read_text())
    out = {"labels": np.load(d / "media.npy"), "origin": tuple(meta["origin"]), "h": float(meta["h"])}
    for k, name in zip("UVWP", ("u", "v", "w", "p")):
        out[k] = np.load(d / f"{name}.npy")
    return out


def grid_to_box(cells: Cells, h: float = 10.0) -> Dict:
    """index = coordinate / h - origin, tight bounding box."""
    if not cells:
        raise ValueError("cannot pack an empty grid")
    ijk = np.array([[c // int(h) for c in key] for key in cells])
    lo, hi = ijk.min(0), ijk.max(0)
    nx, ny, nz = (hi - lo + 1)
    box = {"labels": np.full((nz, ny, nx), ABSENT, dtype=np.uint8), "origin": tuple(int(v) for v in lo),
           "h": float(h)}
    for k in "UVWP":
        box[k] = np.zeros((nz, ny, nx))
    for (i, j, k), (media, vel, p) in zip(ijk - lo, cells.values()):
        box["labels"][k, j, i] = media
        box["U"][k, j, i], box["V"][k, j, i], box["W"][k, j, i] = vel
        box["P"][k, j, i] = 0.0 if p is None else p
    return box


def box_to_grid(box: Dict) -> Cells:
    """Absent cells dropped; pressure is Some p on non-Solid cells (Media.Solid = 2)."""
    h = int(box["h"])
    ox, oy, oz = box["origin"]
    cells: Cells = {}
    for k, j, i in np.argwhere(box["labels"] != ABSENT):
        media = int(box["labels"][k, j, i])
        cells[((i + ox) * h, (j + oy) * h, (k + oz) * h)] = (
            media, (float(box["U"][k, j, i]), float(box["V"][k, j, i]), float(box["W"][k, j, i])),
            None if media == 2 else float(box["P"][k, j, i]))
    return cells
