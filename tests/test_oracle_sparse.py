"""Pins of the faithful oracle: the reference's own still-valid NUnit KATs on
Pressure.divergence (tests/splashy.Tests/Tests.fs:54-73) and the hand-derived
stock-scene numbers of SURVEY.md App. B."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import sparse_ref as sr

finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False)


def fixture_grid():
    """Tests.fs:39-47: Air cell at the origin plus its six Air neighbours."""
    g = sr.Grid()
    air = sr.Cell(sr.AIR, sr.ZERO, sr.ATMOSPHERIC_PRESSURE)
    origin = (0, 0, 0)
    g.add_cells([(origin, air)] + [(n, air) for _, n in sr.neighbors(origin)])
    return g, origin


def test_kat1_divergence_no_velocities():           # Tests.fs:54-57
    g, c = fixture_grid()
    assert sr.divergence(g, c) == 0.0


@settings(max_examples=200, deadline=None)
@given(finite, finite, finite)
def test_kat2_divergence_standalone_cell(x, y, z):  # Tests.fs:59-64
    g, c = fixture_grid()
    g.update_velocities([(c, (x, y, z))])
    assert sr.divergence(g, c) == (-x - y - z)


@settings(max_examples=200, deadline=None)
@given(finite, finite, finite)
def test_kat3_divergence_surrounded_by_rock(x, y, z):  # Tests.fs:66-73
    g, c = fixture_grid()
    g.update_media([(n, sr.SOLID) for _, n in sr.neighbors(c)])
    g.update_velocities([(c, (x, y, z))])
    assert sr.divergence(g, c) == 0.0


def test_stale_test4_is_not_the_current_convention():
    """Tests.fs:75-96 pins the mirrored (min-face) convention; the current
    Pressure.fs:16-32 uses -neighbour velocities as incoming."""
    g, c = fixture_grid()
    g.update_velocities([((10, 0, 0), (5.0, 0.0, 0.0))])     # +x neighbour's own face
    assert sr.divergence(g, c) == 0.0                         # stale test expects +5
    g.update_velocities([((-10, 0, 0), (5.0, 0.0, 0.0))])    # -x neighbour's +face
    assert sr.divergence(g, c) == 5.0


def test_nearest_spot_values():                     # Coord.fs:49-57
    vals = {3: 0, 4: 0, 5: 10, 14: 10, 15: 20, -3: 0, -5: 0, -7: 0, -10: -10,
            -11: -10, -15: -10, -19: -10}
    for n, want in vals.items():
        assert sr.nearest(n) == want
    assert sr.construct_float(4.5, 5.5, -5.5) == (0, 10, 0)   # banker's rounding


def test_vector_ctor_rejects_nonfinite():           # Vector.fs:15-18
    with pytest.raises(sr.ReferenceFailure):
        sr.vec(float("nan"), 0.0, 0.0)
    with pytest.raises(sr.ReferenceFailure):
        sr.vec(0.0, float("inf"), 0.0)


def test_appendix_b_stock_scene():
    """first `advance 0.016671` of the N = 1 scene."""
    dt = 0.016671
    sim = sr.Simulator()
    sim.generate(1)
    m = sim.markers[0]
    sim.grid.move_cells(sim.move_markers(dt))
    sim.stage_setup()
    assert len(sim.grid.cells) == 7          # 1 Fluid + 6 Air, no second layer yet
    sim.stage_convection(dt)
    assert sim.grid.raw_get(m).velocity == (0.0, 0.0, 0.0)
    sim.stage_forces(dt)
    assert sim.grid.raw_get(m).velocity == pytest.approx((0.0, -0.16354251, 0.0), abs=1e-15)
    sim.stage_viscosity(dt)
    assert sim.grid.raw_get(m).velocity[1] == pytest.approx(-0.1635424936414969, rel=1e-15)
    assert sr.H * sr.FLUID_DENSITY / dt == pytest.approx(599826.0452282408, rel=1e-15)
    assert sr.divergence(sim.grid, m) == pytest.approx(0.1635424936414969, rel=1e-15)
    a, b = sr.assemble(sim.grid, sim.markers, dt)
    assert a.tolist() == [[6.0]]
    assert b[0] == pytest.approx(98097.0471877438, rel=1e-14)
    sim.stage_pressure(dt)
    cell = sim.grid.raw_get(m)
    assert cell.pressure == pytest.approx(16349.507864623965, rel=1e-14)
    assert cell.velocity == pytest.approx(
        (0.027257082273582815, -0.13628541136791408, 0.027257082273582815), rel=1e-13)
    for d, n in sr.backward_neighbors(m):
        assert sr.border_value(d, sim.grid.raw_get(n).velocity) == pytest.approx(
            -0.027257082273582815, rel=1e-13)
    assert abs(sr.divergence(sim.grid, m)) < 1e-15
    sim.stage_checks()
    sim.stage_cleanup()
    sim.advance(dt)                           # second step grows the 2-layer buffer
    assert len(sim.grid.cells) == 25


def test_q1_double_gradient_breaks_adjacent_markers():
    """two adjacent fluid markers fail the reference's own check."""
    rng = np.random.default_rng(1)
    sim = sr.Simulator()
    sim.set_markers([(0, 0, 0), (10, 0, 0)])
    sim.stage_setup()
    for k, c in list(sim.grid.cells.items()):
        sim.grid.cells[k] = sr.Cell(c.media, tuple(rng.uniform(-1, 1, 3)), c.pressure)
    a, _ = sr.assemble(sim.grid, sim.markers, 0.016671)
    assert a.tolist() == [[6.0, -1.0], [-1.0, 6.0]]
    sim.stage_pressure(0.016671)
    with pytest.raises(sr.ReferenceFailure, match="Velocity field is not correct"):
        sim.stage_checks()


def test_q3_interpolate_only_touches_cells_next_to_origin():
    """the int-overload snap makes every lookup land in {-10,0,10}^3."""
    touched = set()

    class Spy(sr.Grid):
        def get(self, where):
            touched.add(where)
            return super().get(where)

    g = Spy()
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, y, z = rng.uniform(-100, 100, 3)
        sr.get_interpolated_velocity(g, x, y, z)
    assert touched <= {(a, b, c) for a in (-10, 0, 10) for b in (-10, 0, 10)
                       for c in (-10, 0, 10)}


def test_q5_convection_is_identity_below_300_m_per_s():
    sim = sr.Simulator()
    sim.set_markers([(0, 0, 0), (30, 0, 0)])
    sim.stage_setup()
    rng = np.random.default_rng(3)
    for k, c in list(sim.grid.cells.items()):
        sim.grid.cells[k] = sr.Cell(c.media, tuple(rng.uniform(-2.5, 2.5, 3)), c.pressure)
    before = {m: sim.grid.raw_get(m).velocity for m in sim.markers}
    sim.stage_convection(0.016671)
    assert {m: sim.grid.raw_get(m).velocity for m in sim.markers} == before
