"""Dense (intended-semantics) oracle: agrees with the faithful transcription
on the scenes where the reference is self-consistent, and satisfies the
manufactured properties the CUDA path is later held to."""
import numpy as np
import pytest

from oracle import dense_ref as dr
from oracle import sparse_ref as sr
from splashy_b200 import scenes


def iso_scene(k, seed=12349, vel=1.0):
    """ref_iso_k: k isolated markers, random velocities on every
    existing cell (markers and their air buffer)."""
    rng = np.random.default_rng(seed)
    cells = set()
    while len(cells) < k:
        c = tuple(int(x) * 10 for x in rng.integers(-8, 9, size=3))
        if all(sum(abs(c[a] - o[a]) for a in range(3)) > 10 for o in cells):
            cells.add(c)
    sim = sr.Simulator()
    sim.set_markers(cells)
    sim.stage_setup()
    for key, c in list(sim.grid.cells.items()):
        sim.grid.cells[key] = sr.Cell(c.media, tuple(rng.uniform(-vel, vel, 3)), c.pressure)
    return sim


@pytest.mark.parametrize("k", [1, 4, 32])
def test_intended_equals_faithful_on_isolated_markers(k):
    dt = 0.016671
    sim = iso_scene(k)
    lab, U, V, W, P, org = dr.pack_sparse(sim.grid.cells)
    d = dr.divergence(lab, U, V, W)
    for m in sim.markers:
        i, j, kk = (m[a] // 10 - org[a] for a in range(3))
        assert d[kk, j, i] == sr.divergence(sim.grid, m)
    # convection: identity in the reference (Q5); intended advection is compared on the GPU
    sim.stage_pressure(dt)
    sim.stage_checks()
    U2, V2, W2, Pd, info = dr.project(lab, U, V, W, dt, sr.H, sr.FLUID_DENSITY, direct=True)
    _, Us, Vs, Ws, Ps, _ = dr.pack_sparse(sim.grid.cells)
    for a, b in ((U2, Us), (V2, Vs), (W2, Ws), (Pd, Ps)):
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)
    assert info["max_abs_div"] < 1e-12


def test_codes_and_operator_match_assembled_matrix():
    sc = scenes.random_box(9, 7, 5, seed=3)
    lab = sc["labels"]
    codes = dr.stencil_codes(lab)
    A, where = dr.assemble_csr(lab)
    rng = np.random.default_rng(0)
    s = np.zeros(lab.size)
    s[where] = rng.normal(size=len(where))
    q = dr.apply_A(codes, s.reshape(lab.shape)).ravel()
    np.testing.assert_allclose(q[where], A @ s[where], rtol=1e-13, atol=1e-13)
    assert np.all(q[np.setdiff1d(np.arange(lab.size), where)] == 0)
    assert (A != A.T).nnz == 0                       # Pressure.fs:61-62


@pytest.mark.parametrize("precond", [dr.PC_NONE, dr.PC_JACOBI, dr.PC_RBIC0])
def test_pcg_agrees_with_direct_solve(precond):
    sc = scenes.dambreak(16, dtype=np.float64)
    lab = sc["labels"]
    b = dr.rhs(lab, sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
    p_direct = dr.solve_direct(lab, b)
    p, info = dr.pcg(lab, b, precond, tol=1e-12, tau=0.0)
    assert info["r_inf"] <= 1e-12 * info["b_inf"]
    np.testing.assert_allclose(p, p_direct, rtol=1e-8, atol=1e-8 * np.abs(p_direct).max())


def test_rbic0_cuts_iterations():
    sc = scenes.vortex(24, 24, 24, dtype=np.float64)
    b = dr.rhs(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
    _, ij = dr.pcg(sc["labels"], b, dr.PC_JACOBI, tol=1e-6)
    _, ir = dr.pcg(sc["labels"], b, dr.PC_RBIC0, tol=1e-6)
    assert ir["iters"] < ij["iters"]


@pytest.mark.parametrize("name", ["dambreak", "vortex", "drop2d", "random"])
def test_projection_is_divergence_free(name):
    sc = {"dambreak": lambda: scenes.dambreak(12, dtype=np.float64),
          "vortex": lambda: scenes.vortex(14, 12, 10, dtype=np.float64),
          "drop2d": lambda: scenes.drop2d(32),
          "random": lambda: scenes.random_box(11, 9, 7, seed=5)}[name]()
    U2, V2, W2, P, info = dr.project(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"],
                                     sc["h"], sc["rho"], tol=1e-13)
    scale = max(1e-30, max(np.abs(sc[k]).max() for k in "UVW"))
    assert info["max_abs_div"] <= 1e-9 * scale
    assert info["nonfinite"] == 0


def test_advect_constant_field_is_constant_in_the_interior():
    n = 16
    lab = np.full((n, n, n), dr.FLUID, dtype=np.uint8)
    U = np.full((n, n, n), 3.0)
    V = np.full((n, n, n), -2.0)
    W = np.full((n, n, n), 1.0)
    U2, V2, W2 = dr.advect(lab, U, V, W, dt=2.0, h=10.0)
    core = (slice(3, -3),) * 3
    assert np.allclose(U2[core], 3.0) and np.allclose(V2[core], -2.0) and np.allclose(W2[core], 1.0)


def test_advect_translates_a_linear_profile_exactly():
    """Trilinear interpolation reproduces linear fields: with uniform velocity
    (a, 0, 0) the profile V = y-index is unchanged, and U = const + slope * x
    advected by itself obeys U'(x) = U(x - U dt / h ...) only to 2nd order, so
    use a passive check: V = x-index under uniform U = a shifts by a dt / h."""
    n = 24
    lab = np.full((n, n, n), dr.FLUID, dtype=np.uint8)
    a, dt, h = 4.0, 1.5, 10.0
    kk, jj, ii = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    U = np.full((n, n, n), a)
    V = 1e-9 * ii.astype(np.float64)          # passive: too small to move anything
    W = np.zeros((n, n, n))
    _, V2, _ = dr.advect(lab, U, V, W, dt, h)
    core = (slice(4, -4),) * 3
    np.testing.assert_allclose(V2[core], 1e-9 * (ii[core] - a * dt / h), rtol=1e-9)


def test_advect_only_touches_fluid_cells():
    sc = scenes.random_box(10, 9, 8, seed=7, vel=8.0)
    U2, V2, W2 = dr.advect(sc["labels"], sc["U"], sc["V"], sc["W"], 1.0, sc["h"])
    keep = sc["labels"] != dr.FLUID
    assert np.array_equal(U2[keep], sc["U"][keep])
    assert np.array_equal(W2[keep], sc["W"][keep])
    assert not np.array_equal(U2[~keep], sc["U"][~keep])


def test_scene_slabs_are_partition_invariant():
    whole = scenes.vortex(12, 10, 16)
    for z0, nz in ((0, 8), (8, 8), (4, 5)):
        part = scenes.vortex(12, 10, nz, z0=z0, nz_global=16)
        for k in ("labels", "U", "V", "W"):
            assert np.array_equal(part[k], whole[k][z0:z0 + nz])


# --------------------------------------------------------------------------
# multigrid preconditioner (oracle of SPL_PC_MG)
# --------------------------------------------------------------------------
def test_multigrid_vcycle_is_a_symmetric_positive_preconditioner():
    rng = np.random.default_rng(11)
    sc = scenes.dambreak(24, dtype=np.float64)
    lab = sc["labels"]
    fl = lab == dr.FLUID
    for nu in (1, 2, 3):
        mg = dr.Multigrid(lab, nu=nu)
        x = np.where(fl, rng.uniform(-1, 1, lab.shape), 0.0)
        y = np.where(fl, rng.uniform(-1, 1, lab.shape), 0.0)
        mx, my = mg.vcycle(x), mg.vcycle(y)
        assert abs((x * my).sum() - (mx * y).sum()) <= 1e-12 * abs((x * my).sum())
        assert (x * mx).sum() > 0 and (y * my).sum() > 0
        assert np.all(mx[~fl] == 0)


def test_multigrid_pcg_converges_to_the_direct_solve_in_fewer_iterations():
    sc = scenes.vortex(24, 20, 28, dtype=np.float64)
    lab, U, V, W = sc["labels"], sc["U"], sc["V"], sc["W"]
    b = dr.rhs(lab, U, V, W, sc["dt"], sc["h"], sc["rho"])
    p_mg, i_mg = dr.pcg(lab, b, dr.PC_MG, tol=1e-10)
    _, i_j = dr.pcg(lab, b, dr.PC_JACOBI, tol=1e-10)
    p = dr.solve_direct(lab, b)
    fl = lab == dr.FLUID
    assert np.abs(p_mg[fl] - p[fl]).max() <= 1e-8 * np.abs(p).max()
    assert i_mg["iters"] * 3 < i_j["iters"]


def test_multigrid_chebyshev_weights():
    w = dr.mg_weights(2)
    assert abs(w[0] - 0.5695) < 1e-3 and abs(w[1] - 1.7321) < 1e-3
    assert dr.mg_weights(1) == [6.0 / 7.0] or abs(dr.mg_weights(1)[0] - 6.0 / 7.0) < 1e-15
    assert dr.mg_weights(3, omega=2.0 / 3.0) == [2.0 / 3.0] * 3


# --------------------------------------------------------------------------
# scalar advection (north_star "velocity and scalar fields"; no reference counterpart)
# --------------------------------------------------------------------------
def test_scalar_advection_shifts_exactly_under_a_uniform_integer_flow():
    """dt * velocity = (2, -1, 1) cells everywhere: the departure point of every cell is a
    cell centre, so the new field is the old one shifted (zeros enter from outside)."""
    nz, ny, nx = 9, 10, 12
    rng = np.random.default_rng(5)
    lab = np.full((nz, ny, nx), dr.FLUID, dtype=np.uint8)
    h, dt = 10.0, 1.0
    U = np.full(lab.shape, 2.0 * h / dt)
    V = np.full(lab.shape, -1.0 * h / dt)
    W = np.full(lab.shape, 1.0 * h / dt)
    S = rng.uniform(0.0, 1.0, lab.shape)
    got = dr.advect_scalar(lab, U, V, W, S, dt, h)
    # interior cells whose whole RK2 footprint sees the uniform flow
    want = np.zeros_like(S)
    want[1:, :-1, 2:] = S[:-1, 1:, :-2]
    inner = (slice(3, -3), slice(3, -3), slice(4, -4))
    assert np.abs(got[inner] - want[inner]).max() <= 1e-13


def test_scalar_advection_keeps_constants_and_non_fluid_cells():
    sc = scenes.vortex(24, 20, 16, cfl=1.7, dtype=np.float64)
    lab = sc["labels"]
    S = np.full(lab.shape, 3.25)
    got = dr.advect_scalar(lab, sc["U"], sc["V"], sc["W"], S, sc["dt"], sc["h"])
    inner = (slice(4, -4),) * 3                 # footprint inside the box: weights sum to 1
    assert np.abs(got[inner] - 3.25).max() <= 1e-12
    S2 = np.random.default_rng(1).uniform(size=lab.shape)
    got2 = dr.advect_scalar(lab, sc["U"], sc["V"], sc["W"], S2, sc["dt"], sc["h"])
    assert np.array_equal(got2[lab != dr.FLUID], S2[lab != dr.FLUID])
    # bounded by the data (trilinear weights are a convex combination; outside = 0)
    assert got2.min() >= 0.0 and got2.max() <= S2.max() + 1e-15


def test_scalar_advection_c_port_matches_numpy():
    from oracle import c_port
    sc = scenes.vortex(40, 24, 16, cfl=2.0)
    lab = sc["labels"]
    U, V, W = (np.ascontiguousarray(sc[k], dtype=np.float32) for k in "UVW")
    S = scenes.blob(40, 24, 16)
    got = c_port.advect_scalar(lab, U, V, W, S, sc["dt"], sc["h"])
    ref = dr.advect_scalar(lab, *(a.astype(np.float64) for a in (U, V, W, S)), sc["dt"], sc["h"])
    assert np.abs(got - ref).max() <= 2e-6
