"""SURVEY 8f row N3 on the GPU: marker motion + cell relabelling through the C ABI
(spl_set_markers / spl_move_markers / spl_relabel / spl_advance) against the oracle:
literal mode (all quirk flags) == the transcription of Simulator.fs / Grid.fs / Build.fs,
intended mode == oracle/dense_ref.py."""
import numpy as np
import pytest

from oracle import dense_ref as dr
from oracle import sparse_ref as sr
from splashy_b200 import Solver, SplashyError, scenes
from splashy_b200 import _native as nat
from test_markers_oracle import H, compare, dense_of, sim_with

pytestmark = pytest.mark.gpu
DT = 0.016671
LITERAL = nat.SPL_Q_ALL


def gpu_state(s):
    got = s.download()
    return s.download_media(), got["U"], got["V"], got["W"], got["P"]


@pytest.mark.parametrize("markers,vmax", [
    ([(0, 0, 0)], 1.0),
    ([(0, 0, 0)], 900.0),
    ([(-40, 20, 0), (30, -30, 50), (0, 70, -60)], 400.0),
    ([(100, 100, 100), (-100, 0, 0)], 1.0),
])
def test_literal_marker_stage_matches_transcription(markers, vmax):
    for seed in (1, 2, 3):
        sim = sim_with(markers, seed, vmax)
        lab, U, V, W, P, org, lo, hi = dense_of(sim, margin=3)
        nz, ny, nx = lab.shape
        loc = np.array(sim.locations, dtype=np.float64)
        try:
            sim.grid.move_cells(sim.move_markers(DT))
            sim.stage_setup(sanity=True)
        except sr.ReferenceFailure:
            with Solver(nx, ny, nz, dtype=np.float64, quirks=LITERAL, origin=org) as s:
                s.upload(lab, U, V, W, P)
                s.set_world(lo, hi)
                s.set_markers(loc)
                with pytest.raises(SplashyError, match="outside bounds|left the dense box"):
                    s.move_markers(DT)
                    s.relabel(sanity=True)
            continue
        with Solver(nx, ny, nz, dtype=np.float64, quirks=LITERAL, origin=org) as s:
            s.upload(lab, U, V, W, P)
            s.set_world(lo, hi)
            s.set_markers(loc)
            moved = s.move_markers(DT)
            assert np.array_equal(s.get_markers(), np.array(sim.locations))
            s.relabel(sanity=True)
            compare(sim, *gpu_state(s), org)
            sim.stage_setup(sanity=True)           # second ring of the lazy buffer
            s.relabel(sanity=True)
            compare(sim, *gpu_state(s), org)
        assert moved == sum(1 for a, b in zip(markers, sim.markers)
                            if sr.construct_int(*a) != b) or moved >= 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_intended_marker_stage_matches_dense_oracle(dtype):
    rng = np.random.default_rng(4)
    n = 28
    lab = np.full((n, n, n), dr.ABSENT, dtype=np.uint8)
    lab[6:22, 6:22, 6:22] = dr.AIR
    lab[6:22, 6:8, 6:22] = dr.SOLID
    cells = rng.integers(8, 20, size=(60, 3))
    lab[cells[:, 2], cells[:, 1], cells[:, 0]] = dr.FLUID
    ex = lab != dr.ABSENT
    U, V, W = (np.where(ex, rng.uniform(-600, 600, lab.shape), 0).astype(dtype) for _ in range(3))
    P = np.where(lab == dr.FLUID, rng.uniform(0, 1e4, lab.shape), 0)
    org = (-14, -14, -14)
    loc = (cells + np.array(org)) * float(H) + rng.uniform(-4.9, 4.9, cells.shape)
    lo, hi = (6, 6, 6), (20, 25, 20)
    new_loc, old, new = dr.move_markers(lab, U, V, W, loc, DT, H, org)
    want = dr.relabel(lab, U, V, W, P, new, lo, hi)
    with Solver(n, n, n, dtype=dtype, origin=org) as s:
        s.upload(lab, U, V, W, P.astype(dtype))
        s.set_world(lo, hi)
        s.set_markers(loc)
        moved = s.move_markers(DT)
        assert moved == int((old != new).any(1).sum()) and moved > 10
        assert np.array_equal(s.get_markers(), new_loc)
        s.relabel(sanity=False)
        got = gpu_state(s)
    assert np.array_equal(got[0], want[0])
    for g, w_ in zip(got[1:4], want[1:4]):
        assert np.array_equal(g, w_)
    assert np.array_equal(got[4], want[4].astype(dtype))
    assert ((got[0] == dr.SOLID) & (lab == dr.ABSENT)).sum() > 0      # Solid created outside
    assert ((got[0] == dr.AIR) & (lab == dr.ABSENT)).sum() > 0        # Air created inside
    assert ((got[0] == dr.ABSENT) & (lab != dr.ABSENT)).sum() > 0     # unused cells deleted


def test_literal_advance_with_markers_matches_the_transcription():
    """The whole of Simulator.advance (Simulator.fs:54-110), three steps, stock N = 1 scene
    with a random initial velocity: spl_advance with every quirk flag == sparse_ref."""
    sim = sim_with([(0, 0, 0)], seed=5, vmax=1.0)
    lab, U, V, W, P, org, lo, hi = dense_of(sim, margin=4)
    nz, ny, nx = lab.shape
    with Solver(nx, ny, nz, dtype=np.float64, tol=1e-14, quirks=LITERAL, origin=org) as s:
        s.upload(lab, U, V, W, P)
        s.set_world(lo, hi)
        s.set_markers(np.array(sim.locations))
        for step in range(3):
            sim.advance(DT, sanity=False)
            s.advance(DT, sanity=False)
            assert np.array_equal(s.get_markers(), np.array(sim.locations))
            labg, Ug, Vg, Wg, Pg = gpu_state(s)
            _, Us, Vs, Ws, Ps, _ = dr.pack_sparse(sim.grid.cells)
            for (x, y, z), c in sim.grid.cells.items():
                i, j, k = x // H - org[0], y // H - org[1], z // H - org[2]
                assert labg[k, j, i] == c.media
                np.testing.assert_allclose((Ug[k, j, i], Vg[k, j, i], Wg[k, j, i]), c.velocity,
                                           rtol=1e-11, atol=1e-12)
            assert (labg != dr.ABSENT).sum() == len(sim.grid.cells)


def test_intended_advance_with_markers_dambreak_runs():
    sc = scenes.dambreak(32, dtype=np.float32)
    lab = sc["labels"]
    k, j, i = np.nonzero(lab == dr.FLUID)
    loc = np.stack([i, j, k], 1) * float(H)
    with Solver(32, 32, 32, dtype=np.float32, tol=1e-6) as s:
        s.upload(lab, sc["U"] * 0, sc["V"] * 0, sc["W"] * 0)
        s.set_markers(loc)
        for _ in range(4):
            info = s.advance(sc["dt"], sanity=False)
            assert info["converged"] == 1 and info["nonfinite"] == 0
        out = s.download_media()
        new = s.get_markers()
    cells = dr.marker_cells(new, H, (0, 0, 0))
    uniq = np.unique(cells, axis=0)
    assert (out == dr.FLUID).sum() == len(uniq)
    assert new[:, 1].mean() < loc[:, 1].mean()              # the column is falling
