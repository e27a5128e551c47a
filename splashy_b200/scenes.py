"""Synthetic scenes for the parity tests and the bench.

Pure numpy, deterministic, and partition-invariant: every plane's noise is
drawn from ``default_rng([seed, global_k])`` so a z-slab generated on its own
rank equals the same planes cut out of the whole box.

Every scene returns ``dict(labels, U, V, W, dt, h, rho, name)`` with arrays
``[k, j, i]``; velocities are on the +faces.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

AIR, FLUID, SOLID, ABSENT = 0, 1, 2, 255
H = 10.0            # Constants.fs:11
RHO = 999.97        # Constants.fs:15
REF_DT = 0.016671   # Library.fs:101


def _noise(seed: int, z0: int, nz: int, ny: int, nx: int, amp: float, dtype):
    out = np.empty((3, nz, ny, nx), dtype=dtype)
    for k in range(nz):
        rng = np.random.default_rng([seed, z0 + k])
        out[:, k] = rng.uniform(-amp, amp, size=(3, ny, nx)).astype(dtype)
    return out


def _shell_labels(nx, ny, nz, z0, nzg, air_top=True, two_d=False):
    """All Fluid, 1-cell Solid shell on the global box, Air layer under the
    +y wall so every connected fluid region touches a Dirichlet cell."""
    lab = np.full((nz, ny, nx), FLUID, dtype=np.uint8)
    if air_top:
        lab[:, ny - 2, :] = AIR
    lab[:, 0, :] = SOLID
    lab[:, ny - 1, :] = SOLID
    lab[:, :, 0] = SOLID
    lab[:, :, nx - 1] = SOLID
    if not two_d:
        kg = np.arange(z0, z0 + nz)
        lab[kg == 0] = SOLID
        lab[kg == nzg - 1] = SOLID
    return lab


def vortex(nx: int, ny: int, nz: int, *, z0: int = 0, nz_global: Optional[int] = None,
           cfl: float = 2.0, dt: float = 1.0, seed: int = 2,
           dtype=np.float32, noise: float = 0.05) -> Dict:
    """cfg4 ``vortex_512`` / cfg5 ``weak_1024`` family: ABC flow + noise, CFL 2,
    all Fluid inside a Solid shell with an Air plane at j = ny-2."""
    nzg = nz_global or nz
    lab = _shell_labels(nx, ny, nz, z0, nzg)
    x = (np.arange(nx, dtype=np.float64) + 0.5) * (2 * np.pi / nx)
    y = (np.arange(ny, dtype=np.float64) + 0.5) * (2 * np.pi / ny)
    z = (np.arange(z0, z0 + nz, dtype=np.float64) + 0.5) * (2 * np.pi / nzg)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij", sparse=True)
    amp = cfl * H / dt / (2.0 + noise)        # |component| <= 2 + noise
    n = _noise(seed, z0, nz, ny, nx, noise, np.float64)
    U = amp * (np.sin(Z) + np.cos(Y) + n[0])
    V = amp * (np.sin(X) + np.cos(Z) + n[1])
    W = amp * (np.sin(Y) + np.cos(X) + n[2])
    solid = lab == SOLID
    out = {}
    for name, F in (("U", U), ("V", V), ("W", W)):
        F = np.broadcast_to(F, lab.shape).astype(dtype)
        F[solid] = 0           # Simulator.fs:106-109: solid velocities are zero
        out[name] = F
    out.update(labels=lab, dt=dt, h=H, rho=RHO, name=f"vortex_{nx}x{ny}x{nzg}")
    return out


def smoke2d(n: int, *, cfl: float = 2.0, dt: float = 1.0, seed: int = 1,
            dtype=np.float32, noise: float = 0.05) -> Dict:
    """cfg2 ``smoke2d_4096``: n x n x 1 Taylor-Green + noise, Solid border, Air
    strip under the top wall."""
    lab = _shell_labels(n, n, 1, 0, 1, two_d=True)
    x = (np.arange(n, dtype=np.float64) + 0.5) * (2 * np.pi / n)
    Y, X = np.meshgrid(x, x, indexing="ij", sparse=True)
    amp = cfl * H / dt / (1.0 + noise)
    nz = _noise(seed, 0, 1, n, n, noise, np.float64)
    U = amp * (np.sin(X) * np.cos(Y) + nz[0, 0])
    V = amp * (-np.cos(X) * np.sin(Y) + nz[1, 0])
    solid = lab[0] == SOLID
    U[solid] = 0
    V[solid] = 0
    return dict(labels=lab, U=U[None].astype(dtype), V=V[None].astype(dtype),
                W=np.zeros((1, n, n), dtype=dtype), dt=dt, h=H, rho=RHO,
                name=f"smoke2d_{n}")


def dambreak(n: int, *, dt: float = REF_DT, dtype=np.float32) -> Dict:
    """cfg3 ``dambreak_256``: Solid shell, Fluid column x < 3n/8, y < 5n/8,
    Air elsewhere, at rest plus one gravity kick (Forces.fs:6-12)."""
    lab = np.full((n, n, n), AIR, dtype=np.uint8)
    lab[:, : (5 * n) // 8, : (3 * n) // 8] = FLUID
    lab[0] = lab[-1] = SOLID
    lab[:, 0] = lab[:, -1] = SOLID
    lab[:, :, 0] = lab[:, :, -1] = SOLID
    U = np.zeros((n, n, n), dtype=dtype)
    W = np.zeros((n, n, n), dtype=dtype)
    V = np.where(lab == FLUID, dtype(-9.81 * dt), dtype(0)).astype(dtype)
    return dict(labels=lab, U=U, V=V, W=W, dt=dt, h=H, rho=RHO,
                name=f"dambreak_{n}")


def drop2d(n: int = 64, *, dt: float = REF_DT, dtype=np.float64) -> Dict:
    """Synthetic stand-in for BASELINE config 1's "64 x 64 drop" (the reference
    has no 2-D scene): Fluid disc r = 12 at (32, 40), Air,
    Solid border, gravity kick."""
    lab = np.full((1, n, n), AIR, dtype=np.uint8)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    r = 12.0 * n / 64.0
    disc = (ii - n / 2) ** 2 + (jj - 5 * n / 8) ** 2 <= r * r
    lab[0][disc] = FLUID
    lab[0, 0, :] = lab[0, -1, :] = SOLID
    lab[0, :, 0] = lab[0, :, -1] = SOLID
    V = np.where(lab == FLUID, -9.81 * dt, 0.0).astype(dtype)
    Z = np.zeros((1, n, n), dtype=dtype)
    return dict(labels=lab, U=Z.copy(), V=V, W=Z.copy(), dt=dt, h=H, rho=RHO,
                name=f"drop2d_{n}")


def random_box(nx: int, ny: int, nz: int, *, seed: int = 0, dtype=np.float64,
               p_solid: float = 0.15, p_air: float = 0.25, dt: float = REF_DT,
               vel: float = 1.0, absent_border: bool = False) -> Dict:
    """Random Air / Fluid / Solid labels and velocities U(-vel, vel): exercises
    every branch of the label logic (ragged sizes welcome)."""
    rng = np.random.default_rng(seed)
    r = rng.random((nz, ny, nx))
    lab = np.where(r < p_solid, SOLID, np.where(r < p_solid + p_air, AIR, FLUID)
                   ).astype(np.uint8)
    if absent_border:
        lab[:, 0, :] = ABSENT
        lab[:, :, -1] = ABSENT
    U, V, W = (rng.uniform(-vel, vel, size=(nz, ny, nx)).astype(dtype)
               for _ in range(3))
    return dict(labels=lab, U=U, V=V, W=W, dt=dt, h=H, rho=RHO,
                name=f"random_{nx}x{ny}x{nz}_s{seed}")


def blob(nx: int, ny: int, nz: int, *, z0: int = 0, nz_global: Optional[int] = None,
         dtype=np.float32) -> np.ndarray:
    """cfg2's "one scalar phi = Gaussian blob" (SURVEY 8d): centred at (0.5, 0.6, 0.5) of
    the global box, sigma = 1/8 of the smallest extent > 1; values in (0, 1]."""
    nzg = nz_global or nz
    ext = [e for e in (nx, ny, nzg) if e > 1]
    sig = max(1.0, min(ext) / 8.0)
    x = np.arange(nx, dtype=np.float64) - 0.5 * (nx - 1)
    y = np.arange(ny, dtype=np.float64) - 0.6 * (ny - 1)
    z = np.arange(z0, z0 + nz, dtype=np.float64) - 0.5 * (nzg - 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij", sparse=True)
    return np.exp(-(X * X + Y * Y + Z * Z) / (2.0 * sig * sig)).astype(dtype)


def slab_of(scene: Dict, z0: int, nz: int) -> Dict:
    """Planes [z0, z0 + nz) of a whole-box scene."""
    out = dict(scene)
    for k in ("labels", "U", "V", "W"):
        out[k] = np.ascontiguousarray(scene[k][z0:z0 + nz])
    return out
