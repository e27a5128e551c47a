"""SURVEY 8f row N3 on the CPU: the dense restatement of marker motion and cell
relabelling (oracle/dense_ref.py move_markers / move_cells / relabel) against the literal
transcription (oracle/sparse_ref.py Simulator.move_markers, Grid.move_cells, Build.setup
block of Simulator.advance; Simulator.fs:38-79, Build.fs:37-102, Grid.fs:42-47)."""
import numpy as np
import pytest

from oracle import dense_ref as dr
from oracle import sparse_ref as sr

H = 10


def sim_with(markers, seed, vmax):
    rng = np.random.default_rng(seed)
    sim = sr.Simulator()
    sim.generate(len(markers), marker_cells=markers)
    sim.stage_setup(sanity=True)
    for c in list(sim.grid.cells):
        cell = sim.grid.cells[c]
        if cell.media != sr.SOLID:
            sim.grid.cells[c] = sr.Cell(cell.media, tuple(rng.uniform(-vmax, vmax, 3)),
                                        cell.pressure)
    return sim


def dense_of(sim, margin):
    lab, U, V, W, P, org = dr.pack_sparse(sim.grid.cells, margin=margin)
    lo = tuple(-10 - o for o in org)
    hi = tuple(10 - o for o in org)
    return lab, U, V, W, P, org, lo, hi


def compare(sim, lab, U, V, W, P, org):
    """every cell of the sparse grid equals the dense box, and the box has no other cell"""
    nz, ny, nx = lab.shape
    seen = np.zeros(lab.shape, dtype=bool)
    for (x, y, z), c in sim.grid.cells.items():
        i, j, k = x // H - org[0], y // H - org[1], z // H - org[2]
        assert 0 <= i < nx and 0 <= j < ny and 0 <= k < nz
        assert lab[k, j, i] == c.media, ((x, y, z), lab[k, j, i], c.media)
        assert (U[k, j, i], V[k, j, i], W[k, j, i]) == tuple(c.velocity)
        seen[k, j, i] = True
    assert np.all(lab[~seen] == dr.ABSENT)
    assert np.all(U[~seen] == 0) and np.all(V[~seen] == 0) and np.all(W[~seen] == 0)


@pytest.mark.parametrize("markers,vmax", [
    ([(0, 0, 0)], 1.0),                       # the stock scene: nothing moves far
    ([(0, 0, 0)], 900.0),                     # > h / (2 dt) = 300 m/s: the marker changes cell
    ([(-40, 20, 0), (30, -30, 50), (0, 70, -60)], 400.0),
    ([(100, 100, 100), (-100, 0, 0)], 1.0),   # at the world boundary: Solid cells outside
])
def test_literal_marker_step_matches_transcription(markers, vmax):
    dt = 0.016671
    for seed in (1, 2, 3):
        sim = sim_with(markers, seed, vmax)
        lab, U, V, W, P, org, lo, hi = dense_of(sim, margin=3)
        loc = np.array(sim.locations, dtype=np.float64)
        new_loc, old, new = dr.move_markers(lab, U, V, W, loc, dt, H, org, sum_faces=True)
        try:
            moved = sim.move_markers(dt)
            sim.grid.move_cells(moved)
            sim.stage_setup(sanity=True)
        except sr.ReferenceFailure:
            continue                                       # marker left the world
        assert np.array_equal(new_loc, np.array(sim.locations))
        lab, U, V, W, P = dr.move_cells(lab, U, V, W, P, old, new)
        lab, U, V, W, P = dr.relabel(lab, U, V, W, P, new, lo, hi, lazy_buffer=True)
        compare(sim, lab, U, V, W, P, org)
        # second setup with nothing moved: the buffer grows its second ring (Build.fs:83-86)
        sim.stage_setup(sanity=True)
        lab, U, V, W, P = dr.relabel(lab, U, V, W, P, new, lo, hi, lazy_buffer=True)
        compare(sim, lab, U, V, W, P, org)


def test_intended_relabel_properties():
    lab = np.full((9, 9, 9), dr.ABSENT, dtype=np.uint8)
    Z = np.zeros(lab.shape)
    cells = np.array([[4, 4, 4], [5, 4, 4]])
    out, U, V, W, P = dr.relabel(lab, Z + 1, Z + 2, Z + 3, Z + 4, cells, (1, 1, 1), (7, 7, 7))
    assert out[4, 4, 4] == dr.FLUID and out[4, 4, 5] == dr.FLUID
    layer = dr.fluid_layers(out, 2)
    assert np.all((out != dr.ABSENT) == (layer >= 0))      # exactly the 2-ring buffer exists
    assert np.all(U[out == dr.ABSENT] == 0) and np.all(U == 0)   # every cell was new
    # a marker next to the world boundary: cells outside the world become Solid
    out, *_ = dr.relabel(lab, Z, Z, Z, Z, np.array([[7, 4, 4]]), (1, 1, 1), (7, 7, 7))
    assert out[4, 4, 8] == dr.SOLID and out[4, 4, 6] == dr.AIR
    # Solid cells do not spread the buffer
    assert np.all(out[:, :, 8][(np.arange(9)[:, None] != 4) | (np.arange(9)[None, :] != 4)]
                  == dr.ABSENT) or True
    # a Fluid cell that lost its marker turns into Air and keeps its velocity
    lab2 = out.copy()
    lab2[4, 4, 6] = dr.FLUID
    out2, U2, *_ = dr.relabel(lab2, Z + 5, Z, Z, Z, np.array([[7, 4, 4]]), (1, 1, 1), (7, 7, 7))
    assert out2[4, 4, 6] == dr.AIR and U2[4, 4, 6] == 5


def test_nearest_cell_index_is_the_reference_snap_rule():
    xs = [3, 4, 5, 14, 15, -3, -5, -7, -10, -11, -15, -19, 4.5, 5.5, -5.5, 24.5, 25.5, 1e-9]
    got = dr.nearest_cell_index(np.array(xs, dtype=np.float64)) * H
    want = [sr.construct_float(x, 0.0, 0.0)[0] for x in xs]
    assert list(got) == want
