"""oracle/dense_ref.c (the timed CPU arm) agrees with oracle/dense_ref.py."""
import numpy as np
import pytest

from oracle import c_port, dense_ref as dr
from splashy_b200 import scenes


@pytest.mark.parametrize("make", [lambda: scenes.vortex(40, 24, 16, cfl=1.5),
                                  lambda: scenes.smoke2d(48),
                                  lambda: scenes.random_box(13, 9, 7, seed=2, vel=12.0,
                                                            dtype=np.float32, dt=1.0)])
def test_c_advect_matches_numpy(make):
    sc = make()
    lab = sc["labels"]
    U, V, W = (np.ascontiguousarray(sc[k], dtype=np.float32) for k in "UVW")
    got = c_port.advect(lab, U, V, W, sc["dt"], sc["h"])
    ref = dr.advect(lab, U.astype(np.float64), V.astype(np.float64), W.astype(np.float64),
                    sc["dt"], sc["h"])
    scale = max(np.abs(a).max() for a in (U, V, W))
    for g, r in zip(got, ref):
        assert np.abs(g - r).max() <= 2e-6 * scale


def test_c_step_matches_numpy_iterations_and_fields():
    sc = scenes.vortex(40, 24, 16, cfl=1.5)
    lab = sc["labels"]
    U, V, W = sc["U"], sc["V"], sc["W"]
    U2, V2, W2, P, info = c_port.step(lab, U, V, W, sc["dt"], sc["h"], sc["rho"], tol=1e-6)
    rU, rV, rW, rP, rinfo = dr.step(lab, *(a.astype(np.float64) for a in (U, V, W)),
                                   sc["dt"], sc["h"], sc["rho"], tol=1e-6)
    assert abs(info["iters"] - rinfo["iters"]) <= 2
    scale = np.abs(rU).max()
    assert np.abs(U2 - rU).max() <= 1e-4 * scale
    assert np.abs(P - rP).max() <= 1e-4 * np.abs(rP).max()


def test_c_fixed_iterations():
    sc = scenes.dambreak(24)
    *_, info = c_port.step(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"],
                           sc["rho"], fixed_iters=17)
    assert info["iters"] == 17
    assert c_port.threads() >= 1
