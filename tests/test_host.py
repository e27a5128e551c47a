"""The compiled host side (C++ mirror of Grid / Convection / Forces / Pressure):
host logic on CPU, and -- marked gpu -- the reference's own NUnit fixture and the
stock scene driven through it."""
import numpy as np
import pytest

from oracle import sparse_ref as sr
from splashy_b200 import host_binding as hb


@pytest.fixture(scope="module", autouse=True)
def _built(built_lib):
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-C", str(root / "splashy_b200" / "host")], check=True,
                   capture_output=True)


def test_coord_nearest_matches_the_transcription():        # Coord.fs:49-67
    for n in list(range(-45, 46)) + [105, -105, 1234567, -7654321]:
        assert hb.nearest(n) == sr.nearest(n)
    rng = np.random.default_rng(0)
    for x, y, z in rng.uniform(-120, 120, size=(200, 3)):
        assert hb.construct_float(x, y, z) == sr.construct_float(x, y, z)
    for v in (0.5, 1.5, 2.5, -0.5, 4.5, 5.5, -5.5, 14.5):  # banker's rounding
        assert hb.construct_float(v, v, v) == sr.construct_float(v, v, v)


def test_grid_semantics_match_grid_fs():
    with hb.Host() as h:
        h.add_cell((0, 0, 0), hb.AIR, pressure=0.0)
        with pytest.raises(hb.HostFailure):                # Dictionary.Add on a duplicate key
            h.add_cell((0, 0, 0), hb.AIR)
        with pytest.raises(hb.HostFailure):                # Grid.fs:20 raw_get on an absent cell
            h.get_cell((10, 0, 0))
        with pytest.raises(hb.HostFailure):                # Grid.fs:49-53 via raw_get
            h.update_velocity((10, 0, 0), (1.0, 0.0, 0.0))
        with pytest.raises(hb.HostFailure, match="NaN"):   # Vector.fs:15-18
            h.update_velocity((0, 0, 0), (float("nan"), 0.0, 0.0))
        h.update_velocity((0, 0, 0), (1.0, 2.0, 3.0))
        assert h.get_cell((0, 0, 0)) == (hb.AIR, (1.0, 2.0, 3.0), 0.0)
        assert h.size() == 1


# ---- the reference's fixture, on the GPU -----------------------------------------------
NBRS = [n for _, n in sr.neighbors((0, 0, 0))]


def fixture():
    """Tests.fs:39-47: Air cell at the origin plus its six Air neighbours."""
    h = hb.Host()
    for c in [(0, 0, 0)] + NBRS:
        h.add_cell(c, hb.AIR, pressure=0.0)
    return h


@pytest.mark.gpu
def test_divergence_no_velocities():                       # Tests.fs:54-57
    with fixture() as h:
        assert h.pressure_divergence((0, 0, 0)) == 0.0


@pytest.mark.gpu
def test_divergence_standalone_cell():                     # Tests.fs:59-64
    rng = np.random.default_rng(0)
    with fixture() as h:
        for _ in range(25):
            v = tuple(rng.uniform(-1e3, 1e3, 3))
            h.update_velocity((0, 0, 0), v)
            assert h.pressure_divergence((0, 0, 0)) == (-v[0] - v[1] - v[2])


@pytest.mark.gpu
def test_divergence_surrounded_by_rock():                  # Tests.fs:66-73
    rng = np.random.default_rng(1)
    with fixture() as h:
        for n in NBRS:
            h.update_media(n, hb.SOLID)
        for _ in range(10):
            h.update_velocity((0, 0, 0), tuple(rng.uniform(-9, 9, 3)))
            assert h.pressure_divergence((0, 0, 0)) == 0.0


@pytest.mark.gpu
def test_divergence_needs_all_six_neighbours():            # Pressure.fs:14 raw_get
    with hb.Host() as h:
        h.add_cell((0, 0, 0), hb.FLUID)
        with pytest.raises(hb.HostFailure):
            h.pressure_divergence((0, 0, 0))


@pytest.mark.gpu
def test_stock_scene_step_matches_appendix_b_and_the_transcription():
    """Simulator.fs:82-97 on the N = 1 scene through the C++ host: same numbers as
    oracle/sparse_ref.py at round-off."""
    dt = 0.016671
    sim = sr.Simulator()
    sim.generate(1)
    sim.stage_setup()
    markers = list(sim.markers)
    with hb.Host(quirks=7) as h:
        for c, cell in sim.grid.cells.items():
            h.add_cell(c, cell.media, cell.velocity, cell.pressure)
        h.convection(markers, dt)
        h.forces(markers, dt)
        iters = h.pressure(markers, dt)
        h.checks(markers)
        sim.stage_convection(dt)
        sim.stage_forces(dt)
        sim.stage_pressure(dt)
        sim.stage_checks()
        assert iters == 1
        for c, cell in sim.grid.cells.items():
            media, vel, p = h.get_cell(c)
            assert media == cell.media
            assert vel == pytest.approx(cell.velocity, rel=1e-12, abs=1e-15)
            if cell.media == sr.FLUID:
                assert p == pytest.approx(cell.pressure, rel=1e-12)
        _, v, p = h.get_cell(markers[0])
        assert p == pytest.approx(16349.5098, rel=1e-6)    # App. B (without viscosity)
        assert v[0] == pytest.approx(0.0272571, rel=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [4, 32])
def test_isolated_markers_step_matches_the_transcription(k):
    from test_oracle_dense import iso_scene
    dt = 0.016671
    sim = iso_scene(k)
    markers = list(sim.markers)
    with hb.Host(quirks=7) as h:
        for c, cell in sim.grid.cells.items():
            h.add_cell(c, cell.media, cell.velocity, cell.pressure)
        h.convection(markers, dt)
        h.forces(markers, dt)
        h.pressure(markers, dt)
        h.checks(markers)
        sim.stage_convection(dt)
        sim.stage_forces(dt)
        sim.stage_pressure(dt)
        for c, cell in sim.grid.cells.items():
            _, vel, p = h.get_cell(c)
            assert vel == pytest.approx(cell.velocity, rel=1e-11, abs=1e-13)


@pytest.mark.gpu
def test_adjacent_markers_pass_the_check_the_reference_fails():
    """two adjacent fluid cells -- the reference's own Pressure.apply fails
    its check_divergence; the intended semantics pass it."""
    rng = np.random.default_rng(1)
    sim = sr.Simulator()
    sim.set_markers([(0, 0, 0), (10, 0, 0)])
    sim.stage_setup()
    with hb.Host() as h:
        for c, cell in sim.grid.cells.items():
            h.add_cell(c, cell.media, tuple(rng.uniform(-1, 1, 3)), cell.pressure)
        h.pressure(sim.markers, 0.016671)
        h.checks(sim.markers)


@pytest.mark.gpu
def test_simulator_mirror_runs_the_whole_advance_on_the_device():
    """Simulator.generate + three Simulator.advance steps through the C++ mirror (state on
    the GPU, every quirk flag) == the transcription (oracle/sparse_ref.py), cell by cell."""
    dt = 0.016671
    v0 = (0.3, -0.2, 0.5)
    sim = sr.Simulator()
    sim.generate(1, marker_cells=[(0, 0, 0)])
    sim.grid.update_velocities([((0, 0, 0), v0)])
    with hb.Simulator(quirks=63, tol=1e-14) as h:
        h.generate([(0, 0, 0)])
        h.set_velocity((0, 0, 0), v0)
        for step in range(3):
            sim.advance(dt, sanity=False)
            h.advance(dt, sanity=False)
            assert h.locations() == [tuple(l) for l in sim.locations]
            got = h.snapshot()
            assert set(got) == set(sim.grid.cells)
            for c, cell in sim.grid.cells.items():
                media, vel, _ = got[c]
                assert media == cell.media
                np.testing.assert_allclose(vel, cell.velocity, rtol=1e-11, atol=1e-12)
        assert len(got) == 25                              # two rings of air around the marker


def test_simulator_mirror_generate_sorts_and_deduplicates_markers():
    with hb.Simulator() as h:                              # Set.ofList (Simulator.fs:122)
        h.generate([(30, 0, 0), (0, 0, 0), (30, 0, 0), (0, 10, 0)])
        assert h.locations() == [(0.0, 0.0, 0.0), (0.0, 10.0, 0.0), (30.0, 0.0, 0.0)]
        snap = h.snapshot()
        assert set(snap) == {(0, 0, 0), (0, 10, 0), (30, 0, 0)}
        assert all(m == hb.FLUID for m, _, _ in snap.values())
