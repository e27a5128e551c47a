"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): advected fields within 1e-5 relative in fp32;
projected velocity / pressure within the solver tolerance; divergence-free
residual parity; iteration count reported.  Bit-exact where the arithmetic is
order-identical (divergence in fp64).
"""
import numpy as np
import pytest

from oracle import dense_ref as dr
from oracle import sparse_ref as sr
from splashy_b200 import Solver, SplashyError, scenes
from splashy_b200 import _native as nat

pytestmark = pytest.mark.gpu

RAGGED = [(1, 1, 1), (3, 3, 3), (23, 17, 9), (32, 8, 5), (5, 1, 7), (64, 64, 1),
          (130, 9, 4), (128, 16, 6)]


def solver_for(sc, **kw):
    nz, ny, nx = sc["labels"].shape
    kw.setdefault("h", sc["h"])
    kw.setdefault("rho", sc["rho"])
    return Solver(nx, ny, nz, **kw)


def fields(sc, dtype):
    return (sc["labels"],) + tuple(np.ascontiguousarray(sc[k], dtype=dtype) for k in "UVW")


def vscale(*arrs):
    return max(1e-30, max(float(np.abs(a).max()) for a in arrs))


def open_pockets(lab):
    """Every connected Fluid region gets an Air contact (one of its cells becomes Air if it has
    none): the Laplacian of Pressure.fs:41-60 is singular on a fully enclosed region."""
    from scipy import ndimage
    lab = lab.copy()
    comp, n = ndimage.label(lab == dr.FLUID)
    air = lab == dr.AIR
    touch = np.zeros(lab.shape, dtype=bool)
    for ax in range(3):
        for d in (1, -1):
            sh = np.roll(air, d, axis=ax)
            idx = [slice(None)] * 3
            idx[ax] = 0 if d == 1 else -1
            sh[tuple(idx)] = False          # no wrap-around
            touch |= sh
    for c in range(1, n + 1):
        cells = comp == c
        if not (cells & touch).any():
            k, j, i = np.argwhere(cells)[0]
            lab[k, j, i] = dr.AIR
    return lab


# --------------------------------------------------------------------------
# divergence (Pressure.fs:16-32)
# --------------------------------------------------------------------------
def seven_cell_box(media_nbrs=sr.AIR, v=(0.0, 0.0, 0.0)):
    """Tests.fs:39-47 fixture as a 3^3 box with the corners absent."""
    lab = np.full((3, 3, 3), dr.ABSENT, dtype=np.uint8)
    lab[1, 1, 1] = sr.AIR
    for k, j, i in ((1, 1, 0), (1, 1, 2), (1, 0, 1), (1, 2, 1), (0, 1, 1), (2, 1, 1)):
        lab[k, j, i] = media_nbrs
    U = np.zeros((3, 3, 3)); V = np.zeros((3, 3, 3)); W = np.zeros((3, 3, 3))
    U[1, 1, 1], V[1, 1, 1], W[1, 1, 1] = v
    return lab, U, V, W


def gpu_divergence(lab, U, V, W, centre_fluid=True):
    lab = lab.copy()
    if centre_fluid:
        lab[tuple(s // 2 for s in lab.shape)] = sr.FLUID   # div is reported on Fluid cells
    nz, ny, nx = lab.shape
    with Solver(nx, ny, nz, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        return s.download(div=True)["div"]


def test_kat1_no_velocities():                       # Tests.fs:54-57
    d = gpu_divergence(*seven_cell_box())
    assert d[1, 1, 1] == 0.0


def test_kat2_standalone_cell():                     # Tests.fs:59-64
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = tuple(rng.uniform(-1e3, 1e3, 3))
        d = gpu_divergence(*seven_cell_box(v=v))
        assert d[1, 1, 1] == (-v[0] - v[1] - v[2])


def test_kat3_surrounded_by_rock():                  # Tests.fs:66-73
    rng = np.random.default_rng(1)
    for _ in range(5):
        d = gpu_divergence(*seven_cell_box(media_nbrs=sr.SOLID, v=tuple(rng.uniform(-9, 9, 3))))
        assert d[1, 1, 1] == 0.0


@pytest.mark.parametrize("shape", RAGGED)
def test_divergence_bit_exact_fp64(shape):
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=nx * 100 + ny, absent_border=nx > 8)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        got = s.download(div=True)["div"]
    want = np.where(lab == dr.FLUID, dr.divergence(lab, U, V, W), 0.0)
    assert np.array_equal(got, want)


# --------------------------------------------------------------------------
# projection (Pressure.fs:41-129)
# --------------------------------------------------------------------------
SMALL_SCENES = {
    "dambreak20": lambda dt: scenes.dambreak(20, dtype=dt),
    "vortex_ragged": lambda dt: scenes.vortex(23, 17, 9, dtype=dt),
    "vortex_vec": lambda dt: scenes.vortex(36, 20, 12, dtype=dt),
    "drop2d_64": lambda dt: scenes.drop2d(64, dtype=dt),
    "smoke2d_96": lambda dt: scenes.smoke2d(96, dtype=dt),
}


@pytest.mark.parametrize("name", sorted(SMALL_SCENES))
@pytest.mark.parametrize("precond", ["none", "jacobi", "rbic0"])
def test_project_fp64_matches_direct_solve(name, precond):
    sc = SMALL_SCENES[name](np.float64)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64, tol=1e-12, precond=precond) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        got = s.download()
    U2, V2, W2, P, _ = dr.project(lab, U, V, W, sc["dt"], sc["h"], sc["rho"], direct=True)
    vs, ps = vscale(U2, V2, W2), vscale(P)
    assert info["converged"] == 1 and info["r_inf"] <= 1e-12 * info["b_inf"]
    for n, ref in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - ref).max() <= 1e-9 * vs, n
    fl = lab == dr.FLUID
    assert np.abs(got["P"][fl] - P[fl]).max() <= 1e-8 * ps
    assert info["max_abs_div"] <= 1e-9 * vscale(U, V, W)
    assert info["nonfinite"] == 0


@pytest.mark.parametrize("name", sorted(SMALL_SCENES))
def test_pcg_iteration_count_matches_oracle(name):
    sc = SMALL_SCENES[name](np.float64)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64, tol=1e-8, precond="jacobi") as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
    b = dr.rhs(lab, U, V, W, sc["dt"], sc["h"], sc["rho"])
    _, oinfo = dr.pcg(lab, b, dr.PC_JACOBI, tol=1e-8)
    assert abs(info["iters"] - oinfo["iters"]) <= 2, (info["iters"], oinfo["iters"])
    assert info["b_inf"] == pytest.approx(oinfo["b_inf"], rel=1e-13)


@pytest.mark.parametrize("name", sorted(SMALL_SCENES))
@pytest.mark.parametrize("tau", [0.0, 0.5])
def test_rbic0_iteration_count_matches_oracle_and_beats_jacobi(name, tau):
    """SPL_PC_RBIC0 (the parallel stand-in for MIC(0)): same
    iterates as oracle/dense_ref.py apply_rbic0, fewer iterations than Jacobi."""
    sc = SMALL_SCENES[name](np.float64)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64, tol=1e-8, precond="rbic0", tau=tau) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
    with solver_for(sc, dtype=np.float64, tol=1e-8, precond="jacobi") as s:
        s.upload(lab, U, V, W)
        jinfo = s.project(sc["dt"])
    b = dr.rhs(lab, U, V, W, sc["dt"], sc["h"], sc["rho"])
    _, oinfo = dr.pcg(lab, b, dr.PC_RBIC0, tol=1e-8, tau=tau)
    assert abs(info["iters"] - oinfo["iters"]) <= 2, (info["iters"], oinfo["iters"])
    assert info["iters"] <= jinfo["iters"]
    assert info["max_abs_div"] <= 1e-6 * vscale(U, V, W)


@pytest.mark.parametrize("name", sorted(SMALL_SCENES) + ["vortex64"])
def test_multigrid_pcg_matches_oracle_and_direct_solve(name):
    """SPL_PC_MG (opt-in): same V-cycle as oracle/dense_ref.Multigrid --
    same iteration count -- and the converged pressure equals the direct solve."""
    sc = (scenes.vortex(64, 64, 64, dtype=np.float64) if name == "vortex64"
          else SMALL_SCENES[name](np.float64))
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64, tol=1e-9, precond="mg") as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        got = s.download()
    with solver_for(sc, dtype=np.float64, tol=1e-9, precond="jacobi") as s:
        s.upload(lab, U, V, W)
        jinfo = s.project(sc["dt"])
    b = dr.rhs(lab, U, V, W, sc["dt"], sc["h"], sc["rho"])
    _, oinfo = dr.pcg(lab, b, dr.PC_MG, tol=1e-9)
    assert abs(info["iters"] - oinfo["iters"]) <= 1, (info["iters"], oinfo["iters"])
    assert info["iters"] < jinfo["iters"]
    # exact answer: sparse LU on the small scenes; a 64^3 LU takes minutes, so the
    # oracle's own MG-PCG at a tighter tolerance stands in there
    big = name == "vortex64"
    U2, V2, W2, P, _ = dr.project(lab, U, V, W, sc["dt"], sc["h"], sc["rho"], direct=not big,
                                  precond=dr.PC_MG, tol=1e-12)
    fl = lab == dr.FLUID
    assert np.abs(got["P"][fl] - P[fl]).max() <= 1e-6 * vscale(P)
    assert np.abs(got["U"] - U2).max() <= 1e-7 * vscale(U2, V2, W2)
    assert info["max_abs_div"] <= 1e-7 * vscale(U, V, W)
    print(name, "MG-PCG iterations", info["iters"], "Jacobi-PCG", jinfo["iters"])


def _vcycle_case(shape, seed, solid_blob=True):
    """Random residual on a box with a Solid shell, an Air layer on top and a Solid blob."""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    lab = np.full(shape, dr.FLUID, dtype=np.uint8)
    lab[0], lab[-1], lab[:, 0], lab[:, :, 0], lab[:, :, -1] = (dr.SOLID,) * 5
    lab[:, -3:] = dr.AIR
    if solid_blob:
        lab[nz // 4:nz // 2, ny // 4:ny // 2, nx // 3:nx // 2] = dr.SOLID
    r = np.where(lab == dr.FLUID, rng.uniform(-1, 1, shape), 0.0)
    return lab, r


@pytest.mark.parametrize("shape,dtype,env", [
    ((64, 64, 128), np.float32, {}),                       # fused marching kernels on 2 levels
    ((40, 72, 192), np.float32, {}),                       # ragged y / z, odd coarse extents
    ((37, 50, 68), np.float32, {}),                        # nx % 4 = 0 but odd nz
    ((64, 64, 128), np.float32, {"SPL_MG_NO_MARCH": "1"}),  # one-thread-per-cell kernels
    ((64, 64, 128), np.float32, {"SPL_MG_NU": "1"}),
    ((64, 64, 128), np.float32, {"SPL_MG_NU": "3"}),
    ((64, 64, 128), np.float32, {"SPL_MG_SUMFORM": "1"}),  # N s - sum form of the stencil
    ((64, 64, 128), np.float32, {"SPL_MG_NO_BOTTOM": "1"}),  # per-level launches to the coarsest
    ((33, 47, 61), np.float64, {"SPL_MG_NO_BOTTOM": "1"}),
    ((64, 64, 128), np.float32, {"SPL_MG_OMEGA": "0.6666666666666666"}),
    ((33, 47, 61), np.float64, {}),
])
def test_multigrid_vcycle_matches_oracle(shape, dtype, env, monkeypatch):
    """One application z = M^-1 r of SPL_PC_MG equals oracle/dense_ref.Multigrid.vcycle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    lab, r = _vcycle_case(shape, seed=7)
    nz, ny, nx = shape
    with Solver(nx, ny, nz, dtype=dtype, precond="mg") as s:
        s.upload(lab, np.zeros(shape), np.zeros(shape), np.zeros(shape))
        z, ms = s.apply_preconditioner(r)
    nu = int(env.get("SPL_MG_NU", "2"))
    om = float(env["SPL_MG_OMEGA"]) if "SPL_MG_OMEGA" in env else None
    ref = dr.Multigrid(lab, nu=nu, dtype=np.float64, omega=om).vcycle(
        r.astype(dtype).astype(np.float64))
    tol = 2e-5 if dtype == np.float32 else 1e-12
    assert np.abs(z - ref).max() <= tol * vscale(ref), (np.abs(z - ref).max(), vscale(ref))
    assert np.all(z[lab != dr.FLUID] == 0)


def test_multigrid_fused_residual_update_equals_the_separate_phase_b(monkeypatch):
    """On level 0 the PCG update r -= alpha q rides on the V-cycle's first pass
    (k_mg_march<FIRST2, ..., RUPD>); SPL_MG_NO_FUSE=1 keeps k_phaseB: same iterates."""
    sc = scenes.vortex(128, 72, 40, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    res = {}
    for nofuse in ("0", "1"):
        monkeypatch.setenv("SPL_MG_NO_FUSE", nofuse)
        with solver_for(sc, dtype=np.float32, tol=1e-6, precond="mg") as s:
            s.upload(lab, U, V, W)
            info = s.project(sc["dt"])
            res[nofuse] = (info, s.download())
    (ia, ga), (ib, gb) = res["0"], res["1"]
    assert ia["converged"] == 1 and ia["iters"] == ib["iters"]
    assert ia["r_inf"] == ib["r_inf"] and ia["b_inf"] == ib["b_inf"]
    for n in "UVWP":
        assert np.array_equal(ga[n], gb[n]), n


def test_iterative_refinement_restores_the_divergence_free_residual_in_fp32(monkeypatch):
    """A multigrid solve takes a few large fp32 steps, so its recursive residual drifts from
    b - A x by ~eps32 |A| |x|; one restart from the fp64 true residual (default for fp32 +
    SPL_PC_MG) brings max|div| after the projection down to the level of the many-small-steps
    Jacobi solve or better, for a handful of extra iterations."""
    sc = scenes.vortex(128, 128, 128, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    res = {}
    for name, pc, rounds in (("jacobi", "jacobi", 0), ("mg0", "mg", 0), ("mg1", "mg", 1)):
        with solver_for(sc, dtype=np.float32, tol=1e-6, precond=pc) as s:
            s.set_refinement(rounds)
            s.upload(lab, U, V, W)
            res[name] = s.project(sc["dt"])
    monkeypatch.setenv("SPL_NO_QUAD", "1")             # one-cell k_true_residual / div / grad / check
    with solver_for(sc, dtype=np.float32, tol=1e-6, precond="mg") as s:
        s.set_refinement(1)
        s.upload(lab, U, V, W)
        scalar = s.project(sc["dt"])
    monkeypatch.delenv("SPL_NO_QUAD")
    assert scalar["iters"] == res["mg1"]["iters"] and scalar["r_inf"] == res["mg1"]["r_inf"]
    assert scalar["max_abs_div"] == res["mg1"]["max_abs_div"]
    print({k: (v["iters"], v["max_abs_div"]) for k, v in res.items()})
    assert all(v["converged"] == 1 for v in res.values())
    assert res["mg1"]["max_abs_div"] < 0.5 * res["mg0"]["max_abs_div"]
    assert res["mg1"]["max_abs_div"] <= 2.0 * res["jacobi"]["max_abs_div"]
    assert res["mg0"]["iters"] <= res["mg1"]["iters"] <= res["mg0"]["iters"] + 8


def test_multigrid_pcg_fp32_dambreak_128():
    sc = scenes.dambreak(128, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    res = {}
    for pc in ("jacobi", "mg"):
        with solver_for(sc, dtype=np.float32, tol=1e-6, precond=pc) as s:
            s.upload(lab, U, V, W)
            res[pc] = s.project(sc["dt"])
    assert res["mg"]["converged"] == 1 and res["mg"]["iters"] * 8 < res["jacobi"]["iters"]
    assert res["mg"]["max_abs_div"] <= 1e-4 * vscale(V)
    print("dambreak128 iterations:", {k: (v["iters"], v["ms_solve"]) for k, v in res.items()})


@pytest.mark.parametrize("name", sorted(SMALL_SCENES))
def test_project_fp32_within_solver_tolerance(name):
    sc = SMALL_SCENES[name](np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        got = s.download()
    U2, V2, W2, P, _ = dr.project(lab, *(a.astype(np.float64) for a in (U, V, W)),
                                  sc["dt"], sc["h"], sc["rho"], direct=True)
    vs = vscale(U, V, W)
    assert info["converged"] == 1
    for n, ref in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - ref).max() <= 1e-4 * vs, n
    fl = lab == dr.FLUID
    assert np.abs(got["P"][fl] - P[fl]).max() <= 1e-4 * vscale(P)
    assert info["max_abs_div"] <= 1e-4 * vs          # divergence-free residual (fp32 floor)


def test_random_labels_enclosed_pockets():
    """Random media: some fluid pockets have no Air contact (singular but
    consistent blocks); velocities and the residual must still match."""
    sc = scenes.random_box(19, 13, 11, seed=42)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64, tol=1e-11, max_iters=5000) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        got = s.download()
    U2, V2, W2, P, oinfo = dr.project(lab, U, V, W, sc["dt"], sc["h"], sc["rho"], tol=1e-11)
    for n, ref in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - ref).max() <= 1e-7
    assert info["max_abs_div"] <= 1e-8
    assert abs(info["iters"] - oinfo["iters"]) <= 3


def test_appendix_b_on_the_gpu():
    """SURVEY App. B through forces -> project (viscosity is 1e-6-scale and off
    the path): c, b, p and the velocities of the N = 1 stock scene."""
    dt = 0.016671
    lab, U, V, W = seven_cell_box()
    lab[1, 1, 1] = sr.FLUID
    with Solver(3, 3, 3, dtype=np.float64, tol=1e-14) as s:
        s.upload(lab, U, V, W)
        s.add_forces(dt)
        info = s.project(dt)
        got = s.download()
        s.check(1e-4)
    g = -9.81 * dt
    p = (sr.H * sr.FLUID_DENSITY / dt) * (-g) / 6.0
    kappa = dt / (sr.H * sr.FLUID_DENSITY)
    assert got["P"][1, 1, 1] == pytest.approx(p, rel=1e-13)
    assert info["iters"] == 1
    assert got["U"][1, 1, 1] == pytest.approx(kappa * p, rel=1e-13)
    assert got["V"][1, 1, 1] == pytest.approx(g + kappa * p, rel=1e-13)
    assert got["U"][1, 1, 0] == pytest.approx(-kappa * p, rel=1e-13)   # -x neighbour's +face
    assert got["V"][1, 0, 1] == pytest.approx(-kappa * p, rel=1e-13)
    assert got["W"][0, 1, 1] == pytest.approx(-kappa * p, rel=1e-13)
    assert info["max_abs_div"] < 1e-15


@pytest.mark.parametrize("k", [4, 32])
def test_isolated_markers_match_the_faithful_reference(k):
    """ref_iso_k: GPU == literal transcription of the reference (fp64)."""
    from test_oracle_dense import iso_scene
    dt = 0.016671
    sim = iso_scene(k)
    lab, U, V, W, P, org = dr.pack_sparse(sim.grid.cells)
    nz, ny, nx = lab.shape
    with Solver(nx, ny, nz, dtype=np.float64, tol=1e-14, quirks=7, origin=org) as s:
        s.upload(lab, U, V, W)
        s.advect(dt)                                     # identity in the reference (Q5)
        adv = s.download()
        info = s.project(dt)
        got = s.download()
    assert np.array_equal(adv["U"], U) and np.array_equal(adv["W"], W)
    sim.stage_convection(dt)
    sim.stage_pressure(dt)
    sim.stage_checks()
    _, Us, Vs, Ws, Ps, _ = dr.pack_sparse(sim.grid.cells)
    for n, ref in (("U", Us), ("V", Vs), ("W", Ws), ("P", Ps)):
        np.testing.assert_allclose(got[n], ref, rtol=1e-12, atol=1e-12)
    assert info["max_abs_div"] <= 1e-4                   # Pressure.fs:148


# --------------------------------------------------------------------------
# advection (Convection.fs:21-66)
# --------------------------------------------------------------------------
ADVECT_SCENES = {
    "vortex_ragged": lambda dt: scenes.vortex(23, 17, 9, dtype=dt),
    "vortex_vec": lambda dt: scenes.vortex(40, 24, 16, dtype=dt),
    "smoke2d": lambda dt: scenes.smoke2d(80, dtype=dt),
    "random": lambda dt: scenes.random_box(21, 14, 10, seed=9, vel=15.0, dtype=dt, dt=1.0),
}


@pytest.mark.parametrize("name", sorted(ADVECT_SCENES))
def test_advect_fp64_matches_oracle(name):
    sc = ADVECT_SCENES[name](np.float64)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        s.advect(sc["dt"])
        got = s.download()
    ref = dr.advect(lab, U, V, W, sc["dt"], sc["h"])
    vs = vscale(U, V, W)
    for n, r in zip("UVW", ref):
        assert np.abs(got[n] - r).max() <= 1e-12 * vs, n


@pytest.mark.parametrize("name", sorted(ADVECT_SCENES))
def test_advect_fp32_within_1e5_relative(name):
    sc = ADVECT_SCENES[name](np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32) as s:
        s.upload(lab, U, V, W)
        s.advect(sc["dt"])
        got = s.download()
    ref = dr.advect(lab, *(a.astype(np.float64) for a in (U, V, W)), sc["dt"], sc["h"])
    vs = vscale(U, V, W)
    for n, r in zip("UVW", ref):
        assert np.abs(got[n] - r).max() <= 1e-5 * vs, n   # north_star tolerance


# ---- scalar fields (north_star "velocity and scalar fields") ---------------------------------
SCALAR_SCENES = {
    "vortex_ragged": lambda dt: scenes.vortex(23, 17, 9, dtype=dt),
    "vortex_vec": lambda dt: scenes.vortex(40, 24, 16, dtype=dt),
    "vortex_tiles": lambda dt: scenes.vortex(72, 40, 21, dtype=dt),      # several tiles, ragged tile edges
    "smoke2d": lambda dt: scenes.smoke2d(80, dtype=dt),
    "random": lambda dt: scenes.random_box(24, 14, 10, seed=9, vel=15.0, dtype=dt, dt=1.0),
}


def scalar_fields(sc, dtype, n=1):
    nz, ny, nx = sc["labels"].shape
    rng = np.random.default_rng(11)
    out = [scenes.blob(nx, ny, nz, dtype=dtype)]
    for _ in range(n - 1):
        out.append(rng.uniform(-1.0, 1.0, sc["labels"].shape).astype(dtype))
    return out


@pytest.mark.parametrize("name", sorted(SCALAR_SCENES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_scalar_advection_matches_oracle(name, dtype):
    """fp64: 1e-12; fp32: the north_star bar for advected fields, 1e-5 relative.  Three scalars:
    the first rides on the velocity pass, the others take scalar-only passes; the velocity
    result must not depend on scalars being present."""
    sc = SCALAR_SCENES[name](dtype)
    lab, U, V, W = fields(sc, dtype)
    S = scalar_fields(sc, dtype, 3)
    with solver_for(sc, dtype=dtype) as s:
        s.upload(lab, U, V, W)
        s.set_scalars(S)
        s.advect(sc["dt"])
        got = s.download()
        gs = s.get_scalars()
    with solver_for(sc, dtype=dtype) as s:
        s.upload(lab, U, V, W)
        s.advect(sc["dt"])
        plain = s.download()
    f64 = [a.astype(np.float64) for a in (U, V, W)]
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for k, Sk in enumerate(S):
        ref = dr.advect_scalar(lab, *f64, Sk.astype(np.float64), sc["dt"], sc["h"])
        assert np.abs(gs[k] - ref).max() <= tol * vscale(Sk), (name, k)
    vs = vscale(U, V, W)
    refv = dr.advect(lab, *f64, sc["dt"], sc["h"])
    for n, r in zip("UVW", refv):
        assert np.abs(got[n] - r).max() <= tol * vs, n
        assert np.abs(got[n] - plain[n]).max() <= (0 if dtype == np.float64 else 2e-6 * vs), n


def test_scalar_advection_exact_shift_on_the_gpu():
    """Uniform flow of (2, -1, 1) cells per step: every departure point is a cell centre, the
    interior of the new field is the old field shifted -- exact in fp32 as well."""
    nz, ny, nx = 20, 36, 40
    rng = np.random.default_rng(5)
    lab = np.full((nz, ny, nx), dr.FLUID, dtype=np.uint8)
    h, dt = 10.0, 1.0
    U = np.full(lab.shape, 2.0 * h / dt, dtype=np.float32)
    V = np.full(lab.shape, -1.0 * h / dt, dtype=np.float32)
    W = np.full(lab.shape, 1.0 * h / dt, dtype=np.float32)
    S = rng.uniform(0.0, 1.0, lab.shape).astype(np.float32)
    with Solver(nx, ny, nz, dtype=np.float32, h=h) as s:
        s.upload(lab, U, V, W)
        s.set_scalars([S])
        s.advect(dt)
        got = s.get_scalars()[0]
    want = np.zeros_like(S)
    want[1:, :-1, 2:] = S[:-1, 1:, :-2]
    inner = (slice(3, -3), slice(3, -3), slice(4, -4))
    assert np.array_equal(got[inner], want[inner])


def test_tile_staged_advection_equals_the_gather_kernel(monkeypatch):
    """k_advect_tma (shared-memory ring, packed pairs) against k_advect (L1 gathers): same
    samples, the lerps associate differently -> a few ulp."""
    sc = scenes.vortex(96, 40, 37, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    S = scalar_fields(sc, np.float32, 2)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SPL_ADV_TMA", mode)
        with solver_for(sc, dtype=np.float32) as s:
            s.upload(lab, U, V, W)
            s.set_scalars(S)
            s.advect(sc["dt"])
            res[mode] = (s.download(), s.get_scalars())
    vs = vscale(U, V, W)
    for n in "UVW":
        assert np.abs(res["1"][0][n] - res["0"][0][n]).max() <= 2e-6 * vs, n
    for a, b in zip(res["1"][1], res["0"][1]):
        assert np.abs(a - b).max() <= 2e-6
    # the two paths really are different kernels: the lerps round differently somewhere
    assert any(not np.array_equal(res["1"][0][n], res["0"][0][n]) for n in "UVW")


def test_scalars_ride_through_step_and_nan_velocities_do_not_fault():
    """spl_step advects the scalars too; a NaN velocity must surface as an error code, not as
    a shared-memory fault of the tile-staged kernel."""
    sc = scenes.vortex(40, 24, 16, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    S = scalar_fields(sc, np.float32, 1)
    with solver_for(sc, dtype=np.float32, tol=1e-6) as s:
        s.upload(lab, U, V, W)
        s.set_scalars(S)
        s.step(sc["dt"])
        gs = s.get_scalars()[0]
    ref = dr.advect_scalar(lab, *(a.astype(np.float64) for a in (U, V, W)),
                           S[0].astype(np.float64), sc["dt"], sc["h"])
    assert np.abs(gs - ref).max() <= 1e-5
    U2 = U.copy()
    U2[8, 12, 20] = np.nan
    U2[9, 12, 21] = np.inf
    with solver_for(sc, dtype=np.float32) as s:
        s.upload(lab, U2, V, W)
        s.set_scalars(S)
        with pytest.raises(SplashyError):
            s.step(sc["dt"])
        s.upload(lab, U, V, W)          # the context is still usable
        s.advect(sc["dt"])


@pytest.mark.parametrize("cfl", [(2.9, -2.9, 2.9), (-2.95, 2.95, -2.95), (2.4, 2.4, -2.4),
                                 (-0.3, 2.97, 0.1)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_advect_large_offsets_next_to_the_box_faces(cfl, dtype):
    """Backtraces of almost 3 cells in every direction on a box without a solid shell: the
    fast path (cells >= 3 from the x / y faces, footprint [-3, +3]) and the bounds-tested
    worklist path meet exactly where the footprint would leave the box; corners outside the
    box read as 0 (Convection.fs:36-38)."""
    nx, ny, nz = 24, 20, 16
    rng = np.random.default_rng(2)
    h, dt = 10.0, 1.0
    lab = np.full((nz, ny, nx), dr.FLUID, dtype=np.uint8)
    U = (cfl[0] * h / dt + rng.uniform(-0.2, 0.2, lab.shape)).astype(dtype)
    V = (cfl[1] * h / dt + rng.uniform(-0.2, 0.2, lab.shape)).astype(dtype)
    W = (cfl[2] * h / dt + rng.uniform(-0.2, 0.2, lab.shape)).astype(dtype)
    with Solver(nx, ny, nz, dtype=dtype, h=h) as s:
        s.upload(lab, U, V, W)
        try:
            s.advect(dt)
        except SplashyError:
            pass                       # z departures beyond the single-GPU halo are reported
        got = s.download()
    want = dr.advect(lab, U.astype(np.float64), V.astype(np.float64), W.astype(np.float64), dt, h)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    inner = (slice(4, -4),)            # planes whose z footprint stays inside the box + halo
    for n, r in zip("UVW", want):
        assert np.abs(got[n][inner] - r[inner]).max() <= tol * vscale(*want), n


def test_quirk_mode_reproduces_the_reference_trace():
    """SPL_Q_* = 7: forward RK2 trace + nearest-cell whole-vector copy with the
    reference's index-snap interpolation (Convection.fs:21-66), on a state where
    the trace actually leaves the cell (|v| > h / (2 dt))."""
    dt = 0.016671
    sim = sr.Simulator()
    sim.set_markers([(0, 0, 0)])
    sim.stage_setup()
    sim.advance(dt)                                      # grows the 25-cell buffer
    rng = np.random.default_rng(5)
    for key, c in list(sim.grid.cells.items()):
        sim.grid.cells[key] = sr.Cell(c.media, tuple(rng.uniform(-1, 1, 3)), c.pressure)
    m = sim.markers[0]
    sim.grid.cells[m] = sr.Cell(sr.FLUID, (700.0, -20.0, 30.0), None)
    lab, U, V, W, P, org = dr.pack_sparse(sim.grid.cells)
    want_pos = sr.trace(sim.grid, m, dt)
    assert want_pos != m
    nz, ny, nx = lab.shape
    with Solver(nx, ny, nz, dtype=np.float64, quirks=7, origin=org) as s:
        s.upload(lab, U, V, W)
        s.advect(dt)
        got = s.download()
    sim.stage_convection(dt)
    _, Us, Vs, Ws, _, _ = dr.pack_sparse(sim.grid.cells)
    assert np.array_equal(got["U"], Us) and np.array_equal(got["V"], Vs) \
        and np.array_equal(got["W"], Ws)


def test_quirk_mode_trace_too_far_raises_like_the_reference():
    dt = 0.016671
    lab, U, V, W = seven_cell_box()
    lab[1, 1, 1] = sr.FLUID
    U[:] = 5000.0                                        # lands on an absent cell
    with Solver(3, 3, 3, dtype=np.float64, quirks=7, origin=(-1, -1, -1)) as s:
        s.upload(lab, U, V, W)
        with pytest.raises(SplashyError, match="Backwards particle trace went too far"):
            s.advect(dt)


# --------------------------------------------------------------------------
# whole step + validators + error behaviour
# --------------------------------------------------------------------------
def test_step_fp32_vortex():
    sc = scenes.vortex(64, 40, 24, cfl=1.5, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6) as s:
        s.upload(lab, U, V, W)
        info = s.step(sc["dt"])
        got = s.download()
    U1, V1, W1 = dr.advect(lab, *(a.astype(np.float64) for a in (U, V, W)), sc["dt"], sc["h"])
    U2, V2, W2, P, oinfo = dr.project(lab, U1, V1, W1, sc["dt"], sc["h"], sc["rho"], tol=1e-10)
    vs = vscale(U2, V2, W2)
    for n, ref in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - ref).max() <= 1e-4 * vs
    assert info["max_abs_div"] <= 1e-4 * vs
    assert info["kernel_launches"] >= 2 * info["iters"]


def test_no_fluid_cells_is_a_noop():
    lab = np.full((4, 5, 6), sr.AIR, dtype=np.uint8)
    rng = np.random.default_rng(0)
    U, V, W = (rng.normal(size=lab.shape) for _ in range(3))
    with Solver(6, 5, 4, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        info = s.step(0.1)
        got = s.download()
    assert info["iters"] == 0 and info["converged"] == 1
    assert np.array_equal(got["U"], U) and np.array_equal(got["V"], V)


def test_check_reports_divergence_and_nonfinite():
    sc = scenes.vortex(16, 12, 8, dtype=np.float64)
    lab, U, V, W = fields(sc, np.float64)
    with solver_for(sc, dtype=np.float64) as s:
        s.upload(lab, U, V, W)
        with pytest.raises(SplashyError, match="Velocity field is not correct") as e:
            s.check(1e-4)                                # Pressure.fs:149
        assert e.value.code == nat.SPL_E_DIVERGENCE
        want = dr.check(lab, U, V, W)["max_abs_div"]
        assert s.last_info["max_abs_div"] == want
        U[3, 4, 5] = np.nan
        s.upload(lab, U, V, W)
        with pytest.raises(SplashyError, match="NaN") as e:
            s.check(1e30)                                # Vector.fs:17
        assert e.value.code == nat.SPL_E_NONFINITE


def test_max_iters_is_reported_as_not_converged():
    sc = scenes.vortex(32, 24, 16, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-7, max_iters=3) as s:
        s.upload(lab, U, V, W)
        with pytest.raises(SplashyError) as e:
            s.project(sc["dt"])
        assert e.value.code == nat.SPL_E_NOT_CONVERGED
        assert s.last_info["iters"] == 3


def test_call_order_errors():
    with Solver(4, 4, 4) as s:
        with pytest.raises(SplashyError) as e:
            s.project(0.1)
        assert e.value.code == nat.SPL_E_STATE
    with pytest.raises(SplashyError) as e:
        Solver(0, 4, 4)
    assert e.value.code == nat.SPL_E_BADARG


def test_fixed_iters_and_snapshot_restore():
    sc = scenes.vortex(32, 24, 16, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32) as s:
        s.upload(lab, U, V, W)
        s.snapshot()
        s.set_fixed_iters(7)
        a = s.step(sc["dt"])
        ra = s.download()
        s.restore()
        b = s.step(sc["dt"])
        rb = s.download()
    assert a["iters"] == b["iters"] == 7
    for n in "UVWP":
        assert np.array_equal(ra[n], rb[n])               # deterministic reductions


@pytest.mark.parametrize("shape", [(8, 8, 2), (16, 9, 3), (132, 5, 4), (64, 64, 5), (4, 4, 4),
                                   (260, 3, 2), (12, 1, 7)])
@pytest.mark.parametrize("precond", ["jacobi", "mg"])
def test_thin_boxes_fp32_project_matches_oracle(shape, precond):
    """Thin and tiny boxes with nx % 4 = 0 (fewer planes than the cp.async ring is deep, one
    or two z-chunks, single-row tiles) through the fp32 marching kernels: projected field
    equals the oracle's direct solve within the solver tolerance.  Fluid pockets without an
    Air contact (singular blocks) are opened first, so every case must converge."""
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=nx + ny + nz, dtype=np.float32, p_solid=0.1, p_air=0.3)
    lab = open_pockets(sc["labels"])
    _, U, V, W = fields(sc, np.float32)
    f64 = [a.astype(np.float64) for a in (U, V, W)]
    U2, V2, W2, P, _ = dr.project(lab, *f64, sc["dt"], sc["h"], sc["rho"], direct=True)
    with solver_for(sc, dtype=np.float32, tol=1e-7, max_iters=3000, precond=precond) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])        # raises if PCG does not converge
        got = s.download()
    assert info["nonfinite"] == 0 and info["converged"] == 1, (shape, precond, info)
    for n, r in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - r).max() <= 5e-4 * vscale(U2, V2, W2), (n, shape, precond)


@pytest.mark.parametrize("shape", [(192, 37, 21), (128, 8, 9), (516, 70, 12), (64, 64, 64)])
def test_phase_a_variants_are_bit_identical(shape, monkeypatch):
    """k_phaseA_tma (TMA producer warp, the default when nx % 16 = 0), k_phaseA_ws (halo warp)
    and k_phaseA_ring compute the same iterates bit for bit; the
    generic k_phaseA (different summation order of the dot product) to round-off: fixed 25
    Jacobi-PCG iterations on a ragged box."""
    nx, ny, nz = shape
    sc = scenes.vortex(nx, ny, nz, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    res = []
    for env in ({"SPL_TMA": "0"}, {"SPL_NO_WS": "1"}, {"SPL_NO_RING": "1"}, {"SPL_TMA": "1"}):
        for k in ("SPL_NO_WS", "SPL_NO_RING", "SPL_TMA"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with solver_for(sc, dtype=np.float32, tol=1e-6) as s:
            s.upload(lab, U, V, W)
            s.set_fixed_iters(25)
            info = s.project(sc["dt"])
            res.append((info, s.download()))
    info, got = res[1]
    assert info["r_inf"] == res[0][0]["r_inf"]
    for n in "UVWP":
        assert np.array_equal(got[n], res[0][1][n]), n
    info, got = res[3]                                   # TMA producer warp (nx % 16 = 0)
    assert info["r_inf"] == res[0][0]["r_inf"]
    for n in "UVWP":
        assert np.array_equal(got[n], res[0][1][n]), n
    info, got = res[2]
    assert abs(info["r_inf"] - res[0][0]["r_inf"]) <= 1e-5 * res[0][0]["r_inf"]
    for n in "UVWP":
        assert np.abs(got[n] - res[0][1][n]).max() <= 1e-5 * vscale(res[0][1][n]), n


@pytest.mark.parametrize("shape", [(64, 24, 12), (8, 5, 3), (132, 9, 1), (260, 37, 21),
                                   (128, 8, 40), (8, 2, 2)])
def test_quad_kernels_equal_the_one_cell_kernels_bit_for_bit(shape, monkeypatch):
    """k_div_rhs4 / k_grad_sub4 / k_check4 and the one-pass k_grad_check4 (fp32, nx % 4 = 0)
    against the generic kernels on random labels (every branch of the label logic; several
    x tiles, ragged y tiles, several z chunks), and against the oracle."""
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=9, dtype=np.float32, p_solid=0.2, p_air=0.3,
                           absent_border=True)
    lab, U, V, W = fields(sc, np.float32)
    res = {}
    for mode, env in (("0", {"SPL_NO_QUAD": "1"}), ("1", {"SPL_NO_QUAD": "0", "SPL_NO_GC_FUSE": "1"}),
                      ("f", {"SPL_NO_QUAD": "0", "SPL_NO_GC_FUSE": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=400) as s:
            s.upload(lab, U, V, W)
            div = s.download(div=True)["div"]
            s.set_fixed_iters(25)
            info = s.project(sc["dt"])
            res[mode] = (div, s.download(), info)
    d0, g0, i0 = res["0"]
    d1, g1, i1 = res["1"]
    for mode in ("1", "f"):
        dm, gm, im = res[mode]
        assert np.array_equal(d0, dm), mode
        for n in "UVWP":
            assert np.array_equal(g0[n], gm[n]), (mode, n)
        assert i0["max_abs_div"] == im["max_abs_div"] and i0["nonfinite"] == im["nonfinite"], mode
    assert res["f"][2]["kernel_launches"] == i1["kernel_launches"] - 1   # one pass instead of two
    want = dr.divergence(lab, U.astype(np.float64), V.astype(np.float64), W.astype(np.float64))
    fl = lab == dr.FLUID                                 # div is reported on Fluid cells
    assert np.abs(d1 - want)[fl].max() <= 1e-6 * vscale(U, V, W) and np.all(d1[~fl] == 0)


@pytest.mark.parametrize("shape", [(64, 48, 24), (8, 5, 3), (132, 9, 2), (36, 20, 12), (260, 37, 21)])
def test_rbic0_quad_sweeps_equal_the_one_cell_sweeps(shape, monkeypatch):
    """k_rb_march<black|red> (z-marching tiles, the default) and k_rb_black4 / k_rb_red4 (fp32,
    nx % 4 = 0) against k_rb_sweep on random labels: z = M^-1 r is bit for bit the same, so after
    K fixed iterations the solves differ only through the summation order of r.z (<= 1e-5
    relative on p, u, v, w); and the default path converges in the oracle's iteration count
    (fp64 numpy apply_rbic0) to the direct solve."""
    nx, ny, nz = shape
    sc = scenes.random_box(nx, ny, nz, seed=21, dtype=np.float32, p_solid=0.15, p_air=0.3)
    lab = open_pockets(sc["labels"])
    _, U, V, W = fields(sc, np.float32)
    res, zs = {}, {}
    rvec = np.where(lab == dr.FLUID, np.random.default_rng(5).uniform(-1, 1, lab.shape), 0
                    ).astype(np.float32)
    for mode, env in (("0", {"SPL_NO_QUAD": "1", "SPL_NO_RB_MARCH": "0"}),
                      ("q", {"SPL_NO_QUAD": "0", "SPL_NO_RB_MARCH": "1"}),
                      ("m", {"SPL_NO_QUAD": "0", "SPL_NO_RB_MARCH": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=2000, precond="rbic0") as s:
            s.upload(lab, U, V, W)
            zs[mode] = s.apply_preconditioner(rvec)[0]          # one application z = M^-1 r
            s.set_fixed_iters(6)
            info = s.project(sc["dt"])
            res[mode] = (s.download(), info)
    g0, i0 = res["0"]
    for mode in ("q", "m"):
        assert np.array_equal(zs[mode], zs["0"]), mode           # bit for bit
        g1, i1 = res[mode]
        assert i0["iters"] == i1["iters"] == 6, mode
        assert abs(i0["r_inf"] - i1["r_inf"]) <= 1e-4 * i0["b_inf"], mode
        for n in "UVWP":
            assert np.abs(g0[n] - g1[n]).max() <= 1e-5 * vscale(g0[n]), (mode, n)
    monkeypatch.setenv("SPL_NO_QUAD", "0")
    f64 = [a.astype(np.float64) for a in (U, V, W)]
    with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=2000, precond="rbic0") as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        got = s.download()
    b = dr.rhs(lab, *f64, sc["dt"], sc["h"], sc["rho"])
    _, oinfo = dr.pcg(lab, b, dr.PC_RBIC0, tol=1e-6)
    assert info["converged"] == 1 and abs(info["iters"] - oinfo["iters"]) <= 3, (info, oinfo)
    U2, V2, W2, P, _ = dr.project(lab, *f64, sc["dt"], sc["h"], sc["rho"], direct=True)
    for n, r in (("U", U2), ("V", V2), ("W", W2)):
        assert np.abs(got[n] - r).max() <= 5e-4 * vscale(U2, V2, W2), n


def test_async_transfers_and_two_boxes_in_flight_match_the_blocking_calls():
    """spl_upload_async / spl_download_async / spl_sync (pinned buffers): the same bytes as
    spl_upload / spl_step / spl_download, also with two contexts interleaved the way bench.py's
    end-to-end leg does it, and with a state error when nothing was uploaded."""
    import ctypes as C
    from splashy_b200 import PinnedArray
    sc = scenes.vortex(64, 40, 24, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    shape = lab.shape
    ptr = lambda a: C.c_void_p(a.array.ctypes.data)
    hin = {k: PinnedArray(shape, np.float32) for k in "UVW"}
    hlab = PinnedArray(shape, np.uint8)
    hlab.array[...] = lab
    for k, a in zip("UVW", (U, V, W)):
        hin[k].array[...] = a
    outs = [{k: PinnedArray(shape, np.float32) for k in "UVWP"} for _ in range(2)]
    with solver_for(sc, dtype=np.float32) as s:
        s.upload(lab, U, V, W)
        s.set_fixed_iters(30)
        s.step(sc["dt"])
        want = s.download()
    with solver_for(sc, dtype=np.float32) as a, solver_for(sc, dtype=np.float32) as b:
        with pytest.raises(SplashyError):
            a.step(sc["dt"])                                  # nothing uploaded yet
        boxes = [a, b]
        for x in boxes:
            x.set_fixed_iters(30)
        up = lambda x: x.upload_async_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        up(a)
        for i in range(5):
            x, out = boxes[i % 2], outs[i % 2]
            if i + 1 < 5:
                up(boxes[(i + 1) % 2])
            info = x.step(sc["dt"])
            assert info["iters"] == 30
            x.download_async_ptrs(ptr(out["U"]), ptr(out["V"]), ptr(out["W"]), ptr(out["P"]))
        for x in boxes:
            x.sync()
        for out in outs:
            for n in "UVW":
                assert np.array_equal(out[n].array, want[n]), n
            assert np.array_equal(out["P"].array, want["P"].astype(np.float32))
        # a blocking upload after asynchronous traffic, and a blocking download of a pending one
        up(a)
        got = a.download()
        for n, r in (("U", U), ("V", V), ("W", W)):
            assert np.array_equal(got[n], r), n


# --------------------------------------------------------------------------
# BASELINE sizes: size-independent properties
# --------------------------------------------------------------------------
def test_cfg3_dambreak_256_converges_and_is_divergence_free():
    sc = scenes.dambreak(256, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=20000) as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        assert info["converged"] == 1
        assert info["r_inf"] <= 1e-6 * info["b_inf"]
        vs = vscale(V)
        assert info["max_abs_div"] <= 1e-4 * vs
        s.set_fixed_iters(1)
        again = s.project(sc["dt"])                      # projection is idempotent:
        assert again["b_inf"] <= 1e-3 * info["b_inf"]    # nothing left to remove
    print("cfg3 iterations:", info["iters"], "ms_solve:", info["ms_solve"])


def test_cfg2_smoke2d_4096_project_is_divergence_free():
    """BASELINE config 2 at full size: 4096 x 4096 x 1 Taylor-Green + noise, PCG to 1e-6."""
    sc = scenes.smoke2d(4096, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=60000) as s:
        s.upload(lab, U, V, W)
        s.advect(sc["dt"])
        info = s.project(sc["dt"])
        assert info["converged"] == 1 and info["nonfinite"] == 0
        assert info["r_inf"] <= 1e-6 * info["b_inf"]
        assert info["max_abs_div"] <= 1e-4 * vscale(U, V)
    print("cfg2 iterations:", info["iters"], "ms_solve:", info["ms_solve"])


def test_cfg2_smoke2d_4096_multigrid_needs_a_fraction_of_the_iterations():
    """BASELINE config 2 at full size with SPL_PC_MG (2-D: 4x coarsening per level in the
    plane, nz stays 1): same divergence-free result, two orders of magnitude fewer
    iterations than the 14606 of Jacobi-PCG."""
    sc = scenes.smoke2d(4096, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=2000, precond="mg") as s:
        s.upload(lab, U, V, W)
        s.advect(sc["dt"])
        info = s.project(sc["dt"])
        assert info["converged"] == 1 and info["nonfinite"] == 0
        assert info["r_inf"] <= 1e-6 * info["b_inf"]
        assert info["max_abs_div"] <= 1e-4 * vscale(U, V)
        assert info["iters"] < 200
    print("cfg2 MG iterations:", info["iters"], "ms_solve:", info["ms_solve"])


def test_cfg3_dambreak_256_multigrid():
    sc = scenes.dambreak(256, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6, max_iters=500, precond="mg") as s:
        s.upload(lab, U, V, W)
        info = s.project(sc["dt"])
        assert info["converged"] == 1 and info["iters"] < 40
        assert info["max_abs_div"] <= 1e-4 * vscale(V)
    print("cfg3 MG iterations:", info["iters"], "ms_solve:", info["ms_solve"])


def test_cfg4_vortex_512_slab_step_properties():
    """BASELINE config 4 size (512^2 x 64 planes of it to keep the test short): the
    residual falls monotonically enough, nothing escapes, fixed-iteration runs are
    reproducible bit for bit."""
    sc = scenes.vortex(512, 512, 64, dtype=np.float32)
    lab, U, V, W = fields(sc, np.float32)
    with solver_for(sc, dtype=np.float32, tol=1e-6) as s:
        s.upload(lab, U, V, W)
        s.snapshot()
        s.set_fixed_iters(100)
        a = s.step(sc["dt"])
        pa = s.download()["P"]
        s.restore()
        b = s.step(sc["dt"])
        pb = s.download()["P"]
        assert a["trace_escapes"] == 0 and a["nonfinite"] == 0
        assert a["r_inf"] < 0.5 * a["b_inf"]
        assert np.array_equal(pa, pb)
        s.set_fixed_iters(0)
        s.restore()
        c = s.step(sc["dt"])
        assert c["converged"] == 1 and c["max_abs_div"] <= 1e-4 * vscale(U, V, W)


def test_cfg2_smoke2d_4096_advect_properties():
    n = 4096
    lab = np.full((1, n, n), sr.FLUID, dtype=np.uint8)
    U = np.full((1, n, n), 7.0, dtype=np.float32)
    V = np.full((1, n, n), -3.0, dtype=np.float32)
    W = np.zeros((1, n, n), dtype=np.float32)
    with Solver(n, n, 1, dtype=np.float32) as s:
        s.upload(lab, U, V, W)
        s.advect(2.0)
        got = s.download()
    core = (0, slice(4, -4), slice(4, -4))
    assert np.abs(got["U"][core] - 7.0).max() <= 1e-5 * 7.0   # constants are preserved
    assert np.abs(got["V"][core] + 3.0).max() <= 1e-5 * 7.0


# --------------------------------------------------------------------------
# field-level parity at BASELINE sizes, against the C + OpenMP port of the oracle
# (oracle/dense_ref.c, pinned to dense_ref.py by tests/test_oracle_c.py)
# --------------------------------------------------------------------------
BASELINE_SCENES = {
    "cfg3_dambreak_256": lambda: scenes.dambreak(256),
    "cfg4_vortex_256": lambda: scenes.vortex(256, 256, 256),
    "cfg4_vortex_384x320x128": lambda: scenes.vortex(384, 320, 128),
    "cfg2_smoke2d_1024": lambda: scenes.smoke2d(1024),
}


@pytest.mark.parametrize("name", sorted(BASELINE_SCENES))
def test_baseline_size_fields_match_the_c_oracle(name):
    """Value-by-value at cfg2 / cfg3 / cfg4-class sizes: advected u, v, w and a scalar within
    1e-5 relative (north_star); after 25 fixed Jacobi-PCG iterations the pressure and the
    projected u, v, w within fp32 round-off of the C port (same algorithm, fp32 fields / fp64
    dot products; the two differ in summation order only)."""
    from oracle import c_port
    c_port.use_all_cores()
    sc = BASELINE_SCENES[name]()
    lab, U, V, W = fields(sc, np.float32)
    nz, ny, nx = lab.shape
    S = scenes.blob(nx, ny, nz)
    dt, h, rho = sc["dt"], sc["h"], sc["rho"]
    aU, aV, aW = c_port.advect(lab, U, V, W, dt, h)
    aS = c_port.advect_scalar(lab, U, V, W, S, dt, h)
    K = 25
    pU, pV, pW, pP, oinfo = c_port.step(lab, U, V, W, dt, h, rho, fixed_iters=K)
    with solver_for(sc, dtype=np.float32) as s:
        s.upload(lab, U, V, W)
        s.set_scalars([S])
        s.snapshot()
        s.advect(dt)
        adv = s.download()
        gS = s.get_scalars()[0]
        s.restore()
        s.set_fixed_iters(K)
        info = s.step(dt)
        got = s.download()
    vs = vscale(U, V, W)
    for n, r in (("U", aU), ("V", aV), ("W", aW)):
        assert np.abs(adv[n] - r).max() <= 1e-5 * vs, (name, "advect", n)
    assert np.abs(gS - aS).max() <= 1e-5, (name, "scalar")
    assert info["iters"] == K == oinfo["iters"]
    ps = vscale(pP)
    assert np.abs(got["P"] - pP).max() <= 2e-5 * ps, (name, "p", float(np.abs(got["P"] - pP).max() / ps))
    vs2 = vscale(pU, pV, pW)
    for n, r in (("U", pU), ("V", pV), ("W", pW)):
        assert np.abs(got[n] - r).max() <= 2e-5 * vs2, (name, "projected", n)
    assert abs(info["max_abs_div"] - oinfo["max_abs_div"]) <= 1e-4 * max(1.0, oinfo["max_abs_div"])
