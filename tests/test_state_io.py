"""SURVEY 8f N4, the data-format half: the SPLSTATE container and the NumPy export that carry
a sparse grid / a dense box between the C++ host (the stand-in for the F# host), Python and
the F# shim of INTEGRATION.md.  No GPU call is made: the host library only has to load."""
import struct

import numpy as np
import pytest

from splashy_b200 import host_binding as hb
from splashy_b200 import state_io as sio


def random_cells(seed=3, n=40):
    rng = np.random.default_rng(seed)
    cells = {}
    for _ in range(n):
        key = tuple(int(v) * 10 for v in rng.integers(-6, 7, 3))
        media = int(rng.integers(0, 3))
        vel = tuple(float(v) for v in rng.uniform(-5, 5, 3))
        cells[key] = (media, vel, None if media == 2 else float(rng.uniform(0, 1e4)))
    return cells


def host_with(cells):
    h = hb.Host()
    for key, (media, vel, p) in cells.items():
        h.add_cell(key, media, vel, p)
    return h


def host_cells(h, keys):
    out = {}
    for key in keys:
        media, vel, p = h.get_cell(key)
        out[key] = (media, tuple(vel), p)
    return out


def test_sparse_grid_round_trips_between_python_and_the_cpp_host(tmp_path, built_lib):
    cells = random_cells()
    sio.save_grid(tmp_path / "a.spl", cells)
    back, h = sio.load_grid(tmp_path / "a.spl")
    assert back == cells and h == 10.0
    with hb.Host() as host:                                  # Python writes, C++ reads
        host.load(tmp_path / "a.spl", "grid")
        assert host.size() == len(cells) and host_cells(host, cells) == cells
        host.save(tmp_path / "b.spl", "grid")                # C++ writes, Python reads
    assert sio.load_grid(tmp_path / "b.spl")[0] == cells
    assert (tmp_path / "a.spl").read_bytes() == (tmp_path / "b.spl").read_bytes()   # one format


def test_dense_box_and_numpy_export(tmp_path, built_lib):
    cells = random_cells(seed=9)
    box = sio.grid_to_box(cells)
    with host_with(cells) as host:
        host.save(tmp_path / "box.spl", "box")               # pack + write
        (tmp_path / "npy").mkdir()
        host.save(tmp_path / "npy", "npy")
    got = sio.load_box(tmp_path / "box.spl")
    npy = sio.load_box_npy(tmp_path / "npy")
    for g in (got, npy):
        assert g["origin"] == box["origin"] and g["h"] == 10.0
        assert g["labels"].dtype == np.uint8 and np.array_equal(g["labels"], box["labels"])
        for k in "UVWP":
            assert g[k].dtype == np.float64 and np.array_equal(g[k], box[k]), k
    assert sio.box_to_grid(got) == cells
    sio.save_box(tmp_path / "py.spl", box["labels"], box["U"], box["V"], box["W"], box["P"],
                 box["origin"])
    assert (tmp_path / "py.spl").read_bytes() == (tmp_path / "box.spl").read_bytes()
    with hb.Host() as host:                                  # dense file -> sparse grid in C++
        host.load(tmp_path / "py.spl", "box")
        assert host.size() == len(cells) and host_cells(host, cells) == cells


def test_bad_files_are_refused(tmp_path, built_lib):
    cells = random_cells(n=5)
    sio.save_grid(tmp_path / "g.spl", cells)
    raw = (tmp_path / "g.spl").read_bytes()
    (tmp_path / "magic.spl").write_bytes(b"NOTSTATE" + raw[8:])
    (tmp_path / "short.spl").write_bytes(raw[:-7])
    (tmp_path / "ver.spl").write_bytes(raw[:8] + struct.pack("<I", 2) + raw[12:])
    nan = bytearray(raw)
    nan[32 + 16:32 + 24] = struct.pack("<d", float("nan"))     # vx of the first cell
    (tmp_path / "nan.spl").write_bytes(bytes(nan))
    with hb.Host() as host:
        for name, text in (("magic", "not a SPLSTATE"), ("short", "truncated"), ("ver", "version"),
                           ("nan", "invalid quantity")):
            with pytest.raises(hb.HostFailure, match=text):   # NaN: the Vector3d guard, Vector.fs:15-18
                host.load(tmp_path / f"{name}.spl", "grid")
        with pytest.raises(hb.HostFailure, match="not a dense box"):
            host.load(tmp_path / "g.spl", "box")
    for name in ("magic", "short", "ver"):
        with pytest.raises(ValueError):
            sio.load_grid(tmp_path / f"{name}.spl")
