"""SURVEY 8f "next" rows: N1 Forces + Viscosity between the two halves of the path,
N2 post-projection cleanup, and the whole solver part of Simulator.advance."""
import numpy as np
import pytest

from oracle import dense_ref as dr
from oracle import sparse_ref as sr
from splashy_b200 import scenes
from test_oracle_dense import iso_scene


def stock(steps=0, seed=None):
    sim = sr.Simulator()
    sim.generate(1)
    dt = 0.016671
    for _ in range(steps):
        sim.advance(dt)
    sim.grid.move_cells(sim.move_markers(dt))
    sim.stage_setup()
    if seed is not None:
        rng = np.random.default_rng(seed)
        for k, c in list(sim.grid.cells.items()):
            sim.grid.cells[k] = sr.Cell(c.media, tuple(rng.uniform(-1, 1, 3)), c.pressure)
    return sim


# ---- oracle: dense (order-free) vs the literal transcription ----------------------------
def test_dense_viscosity_equals_transcription_on_isolated_markers():
    sim = iso_scene(16)
    lab, U, V, W, _, _ = dr.pack_sparse(sim.grid.cells)
    sim.stage_viscosity(0.016671)
    _, Us, Vs, Ws, _, _ = dr.pack_sparse(sim.grid.cells)
    got = dr.viscosity(lab, U, V, W, 0.016671)
    for g, r in zip(got, (Us, Vs, Ws)):
        assert np.array_equal(g, r)


def test_dense_viscosity_neighbour_rule():
    """Viscosity.fs:12-17: only Fluid neighbours whose d-component has the sign of d."""
    lab = np.full((1, 1, 3), dr.FLUID, dtype=np.uint8)
    U = np.array([[[-2.0, 1.0, 3.0]]])
    Z = np.zeros_like(U)
    got, _, _ = dr.viscosity(lab, U, Z, Z, dt=1.0, nu=1.0)
    # centre: +x neighbour 3 > 0 counts, -x neighbour -2 < 0 counts -> 1 + (3 - 2 - 6*1)
    assert got[0, 0, 1] == 1.0 + (3.0 - 2.0 - 6.0)
    assert got[0, 0, 0] == -2.0 + (1.0 - 6.0 * -2.0)        # +x neighbour 1 > 0
    assert got[0, 0, 2] == 3.0 + (0.0 - 18.0)               # -x neighbour 1 is not < 0


def test_dense_cleanup_equals_transcription_on_the_stock_scene():
    sim = stock(0, seed=0)
    lab, U, V, W, _, _ = dr.pack_sparse(sim.grid.cells)
    sim.stage_cleanup()
    _, Us, Vs, Ws, _, _ = dr.pack_sparse(sim.grid.cells)
    got = dr.cleanup(lab, U, V, W)
    for g, r in zip(got, (Us, Vs, Ws)):
        assert np.array_equal(g, r)


def test_zero_solid_velocities_rule():
    """Build.fs:144-160 + Simulator.fs:106-109."""
    lab = np.array([[[dr.AIR, dr.SOLID, dr.FLUID]]], dtype=np.uint8)
    U = np.array([[[2.0, 5.0, 1.0]]])
    Z = np.zeros_like(U)
    got, _, _ = dr.zero_solid_velocities(lab, U, Z, Z)
    assert got.tolist() == [[[0.0, 0.0, 1.0]]]               # into-solid and solid zeroed
    got, _, _ = dr.zero_solid_velocities(lab, -U, Z, Z)
    assert got.tolist() == [[[-2.0, 0.0, -1.0]]]             # away from the solid: kept


def test_layers_are_bfs_distances():
    lab = np.full((1, 1, 6), dr.AIR, dtype=np.uint8)
    lab[0, 0, 0] = dr.FLUID
    lab[0, 0, 5] = dr.ABSENT
    assert dr.fluid_layers(lab)[0, 0].tolist() == [0, 1, 2, -1, -1, -1]


# ---- GPU ---------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(23, 17, 9), (36, 20, 12), (64, 64, 1), (5, 1, 7)])
def test_gpu_viscosity_matches_oracle(shape):
    from splashy_b200 import Solver
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=11, vel=3.0, absent_border=nx > 8)
    lab, U, V, W = sc["labels"], sc["U"], sc["V"], sc["W"]
    nu, dt = 0.37, 0.11                                      # large nu so the term is visible
    with Solver(nx, ny, nz, dtype=np.float64, viscosity=nu) as s:
        s.upload(lab, U, V, W)
        s.add_viscosity(dt)
        got = s.download()
    ref = dr.viscosity(lab, U, V, W, dt, nu)
    for n, r in zip("UVW", ref):
        np.testing.assert_allclose(got[n], r, rtol=1e-14, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(23, 17, 9), (36, 20, 12), (64, 64, 1), (5, 1, 7)])
def test_gpu_cleanup_matches_oracle(shape):
    from splashy_b200 import Solver
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=5, p_solid=0.2, p_air=0.55, absent_border=nx > 8)
    lab, U, V, W = sc["labels"], sc["U"], sc["V"], sc["W"]
    with Solver(nx, ny, nz, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        s.cleanup()
        got = s.download()
    ref = dr.cleanup(lab, U, V, W)
    for n, r in zip("UVW", ref):
        assert np.array_equal(got[n], r), n


@pytest.mark.gpu
def test_gpu_advance_matches_the_transcription_on_the_stock_scene():
    """Simulator.fs:82-110 in one call (spl_advance, quirk flags on) == the literal
    transcription's stages on the N = 1 scene, random initial velocities."""
    from splashy_b200 import Solver
    dt = 0.016671
    sim = stock(0, seed=3)
    for k, c in list(sim.grid.cells.items()):               # keep the marker below 300 m/s
        pass
    lab, U, V, W, P, org = dr.pack_sparse(sim.grid.cells)
    nz, ny, nx = lab.shape
    with Solver(nx, ny, nz, dtype=np.float64, tol=1e-14, quirks=7, origin=org) as s:
        s.upload(lab, U, V, W)
        info = s.advance(dt, sanity=True)
        got = s.download()
    sim.stage_convection(dt)
    sim.stage_forces(dt)
    sim.stage_viscosity(dt)
    sim.stage_pressure(dt)
    sim.stage_checks()
    sim.stage_cleanup()
    _, Us, Vs, Ws, Ps, _ = dr.pack_sparse(sim.grid.cells)
    for n, r in (("U", Us), ("V", Vs), ("W", Ws)):
        np.testing.assert_allclose(got[n], r, rtol=1e-12, atol=1e-13)
    fl = lab == dr.FLUID
    np.testing.assert_allclose(got["P"][fl], Ps[fl], rtol=1e-12)
    assert info["iters"] == 1 and info["max_abs_div"] < 1e-4


@pytest.mark.gpu
def test_gpu_advance_dambreak_runs_several_steps():
    from splashy_b200 import Solver
    sc = scenes.dambreak(32, dtype=np.float32)
    sc["V"][...] = 0
    with Solver(32, 32, 32, dtype=np.float32, tol=1e-6) as s:
        s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
        for _ in range(3):
            info = s.advance(sc["dt"], sanity=True)          # 1e-4 m/s check, like the reference
            assert info["converged"] == 1 and info["nonfinite"] == 0
        got = s.download()
    assert np.isfinite(got["U"]).all() and np.abs(got["V"]).max() > 0
