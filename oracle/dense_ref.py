"""Dense-array CPU restatement of the advect -> project path, *intended* semantics.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED (the
reference cannot run here); pinned through ``sparse_ref`` on the scenes where
the reference is self-consistent (isolated fluid markers)
and through the reference's divergence KATs.

Conventions (kept from the reference so numbers compare 1:1):

* arrays are ``[k, j, i]`` (x fastest); ``U[k,j,i]`` is the x-velocity on the
  **+x face** of cell (i,j,k) (Pressure.fs:18-31); V, W likewise;
* labels: 0 Air, 1 Fluid, 2 Solid (Cell.fs:9), 255 absent; absent and
  out-of-box cells behave like Solid in every mask and are never counted
  (Pressure.fs:34-38 counts only *existing* neighbours);
* ``divergence`` = incoming - outgoing, not divided by h (Pressure.fs:31-32);
  ``b = (h rho / dt) * divergence`` (Pressure.fs:64-66);
* A = integer Laplacian: diag = #non-solid neighbours, -1 per Fluid neighbour,
  Air is Dirichlet p = 0 (Pressure.fs:44-55, Constants.fs:20).

What differs from the reference on purpose: each face gets
the pressure gradient exactly once, faces on Solid get the solid velocity (0),
updates are Jacobi (double-buffered), advection is a backward RK2 trace per
face with trilinear sampling.  The reference's direct LU solve
(Pressure.fs:69) is restated by ``solve_direct``; ``pcg`` is the iterative
substitute the CUDA path implements (same operation order).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

AIR, FLUID, SOLID, ABSENT = 0, 1, 2, 255

# stencil-code bits (shared with splashy_b200/csrc/spl_common.cuh)
B_XM, B_XP, B_YM, B_YP, B_ZM, B_ZP, B_FLUID = 1, 2, 4, 8, 16, 32, 64

PC_NONE, PC_JACOBI, PC_RBIC0, PC_MG = 0, 1, 2, 3


def _nb(a: np.ndarray, axis: int, d: int, fill=0) -> np.ndarray:
    """out[idx] = a[idx + d along axis]; ``fill`` where that leaves the box.
    axis: 0 = x (last numpy axis), 1 = y, 2 = z (first numpy axis)."""
    ax = 2 - axis
    out = np.full_like(a, fill)
    n = a.shape[ax]
    if abs(d) >= n:
        return out
    src = [slice(None)] * 3
    dst = [slice(None)] * 3
    if d > 0:
        src[ax] = slice(d, n)
        dst[ax] = slice(0, n - d)
    else:
        src[ax] = slice(0, n + d)
        dst[ax] = slice(-d, n)
    out[tuple(dst)] = a[tuple(src)]
    return out


def nonsolid(labels: np.ndarray) -> np.ndarray:
    """Pressure.fs:14 ``is_nonsolid`` as a mask (absent counts as solid)."""
    return (labels == AIR) | (labels == FLUID)


def stencil_codes(labels: np.ndarray) -> np.ndarray:
    """Per-cell u8: bits 0..5 = neighbour (-x,+x,-y,+y,-z,+z) exists and is not
    Solid (Pressure.fs:34-38,53), bit 6 = the cell is Fluid (a pressure unknown,
    Pressure.fs:42-43)."""
    ns = nonsolid(labels).astype(np.uint8)
    code = np.zeros(labels.shape, dtype=np.uint8)
    bit = 1
    for axis in range(3):
        for d in (-1, +1):
            code |= (_nb(ns, axis, d, 0) * bit).astype(np.uint8)
            bit <<= 1
    code |= ((labels == FLUID).astype(np.uint8) * B_FLUID).astype(np.uint8)
    return code


def diag_count(codes: np.ndarray) -> np.ndarray:
    """Pressure.fs:34-38,53: number of existing non-solid neighbours."""
    c = codes & 63
    cnt = np.zeros(codes.shape, dtype=np.int32)
    for b in range(6):
        cnt += (c >> b) & 1
    return cnt


def divergence(labels: np.ndarray, U: np.ndarray, V: np.ndarray,
               W: np.ndarray) -> np.ndarray:
    """Pressure.fs:16-32 for every cell (meaningful on Fluid cells):
    out_c = own +face if the +neighbour is non-solid, in_c = -neighbour's +face
    if the -neighbour is non-solid; (in - out) summed x + y + z."""
    ns = nonsolid(labels)
    dt = U.dtype
    res = []
    for axis, F in enumerate((U, V, W)):
        out = np.where(_nb(ns, axis, +1, False), F, dt.type(0))
        inc = np.where(_nb(ns, axis, -1, False), _nb(F, axis, -1, 0), dt.type(0))
        res.append(inc - out)
    return (res[0] + res[1]) + res[2]


def rhs(labels, U, V, W, dt: float, h: float, rho: float) -> np.ndarray:
    """Pressure.fs:64-68: b = (h rho / dt) * divergence on Fluid, 0 elsewhere."""
    c = U.dtype.type((h * rho) / dt)
    d = divergence(labels, U, V, W)
    return np.where(labels == FLUID, c * d, U.dtype.type(0))


def apply_A(codes: np.ndarray, s: np.ndarray) -> np.ndarray:
    """Matrix-free Pressure.fs:44-55: (A s)_i = N_i s_i - sum_{Fluid nbr} s_n.
    ``s`` must be 0 on non-Fluid cells (Air is p = 0), so the sum may run over
    every non-solid neighbour."""
    q = np.zeros_like(s)
    bit = 1
    for axis in range(3):
        for d in (-1, +1):
            m = (codes & bit) != 0
            q += np.where(m, s - _nb(s, axis, d, 0), s.dtype.type(0))
            bit <<= 1
    return np.where((codes & B_FLUID) != 0, q, s.dtype.type(0))


def assemble_csr(labels: np.ndarray):
    """Pressure.fs:42-60 as scipy CSR over the Fluid cells (row order = C order
    of the box).  Returns (A, flat indices of the unknowns)."""
    import scipy.sparse as sp
    codes = stencil_codes(labels)
    fluid = labels == FLUID
    idx = -np.ones(labels.shape, dtype=np.int64)
    n = int(fluid.sum())
    idx[fluid] = np.arange(n)
    rows = [idx[fluid]]
    cols = [idx[fluid]]
    vals = [diag_count(codes)[fluid].astype(np.float64)]
    for axis in range(3):
        for d in (-1, +1):
            nidx = _nb(idx, axis, d, -1)
            m = fluid & (nidx >= 0)
            rows.append(idx[m])
            cols.append(nidx[m])
            vals.append(-np.ones(int(m.sum())))
    A = sp.csr_matrix((np.concatenate(vals),
                       (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n, n))
    return A, np.flatnonzero(fluid.ravel())


def solve_direct(labels: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Pressure.fs:69 ``m.Solve(b)``: exact sparse LU in fp64."""
    import scipy.sparse.linalg as spla
    A, where = assemble_csr(labels)
    p = np.zeros(labels.size, dtype=np.float64)
    if len(where):
        p[where] = spla.spsolve(A.tocsc(), b.ravel()[where].astype(np.float64))
    return p.reshape(labels.shape)


# --------------------------------------------------------------------------
# preconditioners
# --------------------------------------------------------------------------
def rbic0_diag(codes: np.ndarray, tau: float = 0.0, safety: float = 0.25
               ) -> np.ndarray:
    """Diagonal E of the red-black incomplete Cholesky M = (E+L) E^-1 (E+L^T)
    (red = (i+j+k) even, ordered first).  tau = 0 is IC(0); 0 < tau < 1 adds
    the relaxed MIC(0) compensation of dropped black-black fill.  Substitution
    for the reference's exact solve (Pressure.fs:69) -- see DESIGN.md."""
    nz, ny, nx = codes.shape
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx),
                             indexing="ij")
    black = ((ii + jj + kk) & 1) == 1
    fluid = (codes & B_FLUID) != 0
    N = diag_count(codes).astype(np.float64)
    fl = fluid.astype(np.float64)
    nfl = np.zeros_like(N)
    for axis in range(3):
        for d in (-1, +1):
            nfl += _nb(fl, axis, d, 0.0)
    # contribution of each red Fluid cell j to its black neighbours
    with np.errstate(divide="ignore", invalid="ignore"):
        contrib = np.where(fluid & ~black & (N > 0),
                           (1.0 + tau * (nfl - 1.0)) / N, 0.0)
    acc = np.zeros_like(N)
    for axis in range(3):
        for d in (-1, +1):
            acc += _nb(contrib, axis, d, 0.0)
    E = N.copy()
    eb = N - acc
    eb = np.where(eb < safety * N, N, eb)
    E = np.where(black, eb, E)
    return np.where(fluid, E, 1.0)


def apply_rbic0(codes: np.ndarray, E: np.ndarray, r: np.ndarray) -> np.ndarray:
    """z = M^-1 r for the red-black IC: y_R = r_R/E_R; z_B = (r_B + sum_red-nbr
    y)/E_B; z_R = (r_R + sum_black-nbr z_B)/E_R."""
    nz, ny, nx = codes.shape
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx),
                             indexing="ij")
    black = ((ii + jj + kk) & 1) == 1
    fluid = (codes & B_FLUID) != 0
    t = r.dtype.type
    Ei = (1.0 / E).astype(r.dtype)
    y = np.where(fluid & ~black, r * Ei, t(0))
    acc = np.zeros_like(r)
    for axis in range(3):
        for d in (-1, +1):
            acc += _nb(y, axis, d, 0)
    z = np.where(fluid & black, (r + acc) * Ei, t(0))
    acc = np.zeros_like(r)
    for axis in range(3):
        for d in (-1, +1):
            acc += _nb(z, axis, d, 0)
    z = np.where(fluid & ~black, (r + acc) * Ei, z)
    return z


# ---- geometric multigrid V-cycle (splashy_b200/csrc/spl_mg.cuh) -----------------
def mg_coarsen_labels(lab: np.ndarray) -> np.ndarray:
    """Air if any child is Air, else Fluid if any child is Fluid, else Solid; children
    outside the box (and absent cells) count as Solid."""
    nz, ny, nx = lab.shape
    cz, cy, cx = (nz + 1) // 2, (ny + 1) // 2, (nx + 1) // 2
    pad = np.full((cz * 2, cy * 2, cx * 2), SOLID, dtype=np.uint8)
    pad[:nz, :ny, :nx] = np.where(lab == ABSENT, SOLID, lab)
    ch = pad.reshape(cz, 2, cy, 2, cx, 2).transpose(0, 2, 4, 1, 3, 5).reshape(cz, cy, cx, 8)
    any_air = (ch == AIR).any(-1)
    any_fl = (ch == FLUID).any(-1)
    return np.where(any_air, AIR, np.where(any_fl, FLUID, SOLID)).astype(np.uint8)


def mg_weights(nu: int, omega: Optional[float] = None):
    """Per-sweep damped-Jacobi weights: Chebyshev for eigenvalues of D^-1 A in [1/3, 2]
    (the high-frequency band of the 7-point operator under 2x coarsening), ascending;
    `omega` forces a constant weight instead."""
    if omega is not None:
        return [float(omega)] * nu
    return [1.0 / (7.0 / 6.0 + 5.0 / 6.0 * np.cos(np.pi * (2 * k + 1) / (2.0 * nu)))
            for k in range(nu)]


class Multigrid:
    """V(nu, nu) cycle: damped Jacobi with the weights of `mg_weights` (pre-smoother in
    ascending, post-smoother in descending order -- the cycle stays symmetric), restriction
    = mean of the children (8 in 3-D, 4 in 2-D), piecewise-constant prolongation, rediscretised operators
    4^-l * L_l and `coarse_sweeps` Jacobi sweeps on the coarsest level (smallest extent
    <= 8).  Restates splashy_b200/csrc/spl_mg.cuh + spl_api.cu Impl::vcycle."""

    def __init__(self, labels, nu: int = 2, coarse_sweeps: int = 30, dtype=np.float64,
                 max_levels: int = 12, omega: Optional[float] = None):
        self.nu, self.cs, self.dtype = nu, coarse_sweeps, np.dtype(dtype)
        self.omegas = [self.dtype.type(w) for w in mg_weights(nu, omega)]
        self.levels = []
        lab, scale = labels, 1.0
        while True:
            codes = stencil_codes(lab)
            fluid = (codes & B_FLUID) != 0
            N = diag_count(codes)
            invn = np.where(fluid & (N > 0), 1.0 / np.maximum(N, 1).astype(dtype), 0
                            ).astype(dtype)
            self.levels.append(dict(shape=lab.shape, codes=codes, fluid=fluid, invn=invn,
                                    scale=self.dtype.type(scale),
                                    inv_scale=self.dtype.type(1.0 / scale)))
            dims = [d for d in lab.shape if d > 1]
            if not dims or min(dims) <= 8 or len(self.levels) >= max_levels:
                break
            lab = mg_coarsen_labels(lab)
            scale *= 0.25

    def _lap(self, L, u):
        """N u - sum of in-box non-solid neighbours (u = 0 on non-Fluid cells)."""
        return apply_A(L["codes"], u)

    def _smooth(self, L, u, b, sweeps, first, reverse=False):
        for s in range(sweeps):
            k = s % self.nu
            w = L["invn"] * self.omegas[self.nu - 1 - k if reverse else k]
            if first and s == 0:
                u = w * (b * L["inv_scale"])
            else:
                u = u + w * (b * L["inv_scale"] - self._lap(L, u))
            u = np.where(L["fluid"], u, 0).astype(self.dtype)
        return u

    def vcycle(self, b, lev: int = 0):
        L = self.levels[lev]
        b = np.where(L["fluid"], b, 0).astype(self.dtype)
        last = lev == len(self.levels) - 1
        if last:
            sweeps = 2 * self.nu if len(self.levels) == 1 else self.cs
            return self._smooth(L, None, b, sweeps, True)
        u = self._smooth(L, None, b, self.nu, True)
        r = np.where(L["fluid"], b - L["scale"] * self._lap(L, u), 0)
        cz, cy, cx = self.levels[lev + 1]["shape"]
        pad = np.zeros((cz * 2, cy * 2, cx * 2), dtype=self.dtype)
        pad[:r.shape[0], :r.shape[1], :r.shape[2]] = r
        axes = sum(1 for d in r.shape if d > 1)       # an axis of extent 1 is not coarsened
        rc = pad.reshape(cz, 2, cy, 2, cx, 2).sum(axis=(1, 3, 5)) * self.dtype.type(0.5 ** axes)
        ec = self.vcycle(rc.astype(self.dtype), lev + 1)
        pe = np.repeat(np.repeat(np.repeat(ec, 2, 0), 2, 1), 2, 2)[:r.shape[0], :r.shape[1],
                                                                  :r.shape[2]]
        u = u + np.where(L["fluid"], pe, 0).astype(self.dtype)
        return self._smooth(L, u, b, self.nu, False, reverse=True)


def pcg(labels: np.ndarray, b: np.ndarray, precond: int = PC_JACOBI,
        tol: float = 1e-6, max_iters: int = 10000, tau: float = 0.0,
        dtype=np.float64, fixed_iters: Optional[int] = None
        ) -> Tuple[np.ndarray, Dict]:
    """Preconditioned CG on A p = b in the operation order of the CUDA path:
    stop when max|r| <= tol * max|b|; dot products accumulate in fp64."""
    codes = stencil_codes(labels)
    fluid = (codes & B_FLUID) != 0
    N = diag_count(codes)
    dt_ = np.dtype(dtype).type
    with np.errstate(divide="ignore"):
        invd = np.where(fluid & (N > 0), 1.0 / np.maximum(N, 1), 0.0).astype(dtype)
    E = rbic0_diag(codes, tau) if precond == PC_RBIC0 else None
    mg = Multigrid(labels, dtype=dtype) if precond == PC_MG else None

    def M(r):
        if precond == PC_NONE:
            return r.copy()
        if precond == PC_JACOBI:
            return r * invd
        if precond == PC_MG:
            return mg.vcycle(r)
        return apply_rbic0(codes, E, r)

    def dot(a, c):
        return float(np.dot(a.ravel().astype(np.float64),
                            c.ravel().astype(np.float64)))

    x = np.zeros(labels.shape, dtype=dtype)
    r = np.where(fluid, b, 0).astype(dtype)
    bnorm = float(np.max(np.abs(r))) if r.size else 0.0
    info = {"iters": 0, "b_inf": bnorm, "r_inf": bnorm, "history": [bnorm]}
    if bnorm == 0.0 and fixed_iters is None:
        return x, info
    z = M(r)
    s = z.copy()
    sigma = dot(r, z)
    limit = fixed_iters if fixed_iters is not None else max_iters
    for it in range(1, limit + 1):
        q = apply_A(codes, s)
        sq = dot(s, q)
        if sq == 0.0:
            break
        alpha = sigma / sq
        x = x + dt_(alpha) * s
        r = r - dt_(alpha) * q
        rnorm = float(np.max(np.abs(r)))
        info["iters"] = it
        info["r_inf"] = rnorm
        info["history"].append(rnorm)
        if fixed_iters is None and rnorm <= tol * bnorm:
            break
        z = M(r)
        sigma_new = dot(r, z)
        beta = sigma_new / sigma
        s = z + dt_(beta) * s
        sigma = sigma_new
    return x, info


# --------------------------------------------------------------------------
# gradient subtraction
# --------------------------------------------------------------------------
def grad_sub(labels, U, V, W, p, dt: float, h: float, rho: float):
    """Pressure.fs:96-129 with each face updated once (SURVEY A.7 intended):
    face between c and c+e: both non-solid and at least one Fluid ->
    vel -= (dt/(h rho)) (p[c+e] - p[c]) with p = 0 on Air; one Solid (or absent)
    and the other Fluid -> solid velocity 0; otherwise unchanged."""
    t = U.dtype.type
    kappa = t(dt / (h * rho))
    ns = nonsolid(labels)
    fl = labels == FLUID
    pz = np.where(fl, p, t(0)).astype(U.dtype)
    out = []
    for axis, F in enumerate((U, V, W)):
        ns_p = _nb(ns, axis, +1, False)
        fl_p = _nb(fl, axis, +1, False)
        p_p = _nb(pz, axis, +1, 0)
        both = ns & ns_p & (fl | fl_p)
        wall = (fl & ~ns_p) | (~ns & fl_p)
        G = np.where(both, F - kappa * (p_p - pz), F)
        G = np.where(wall, t(0), G)
        out.append(G.astype(U.dtype))
    return tuple(out)


def check(labels, U, V, W, p=None) -> Dict:
    """Pressure.fs:132-149: max |divergence| over Fluid, non-finite count."""
    d = divergence(labels, U, V, W)
    fl = labels == FLUID
    mx = float(np.max(np.abs(d[fl]))) if fl.any() else 0.0
    nonfinite = int((~np.isfinite(U)).sum() + (~np.isfinite(V)).sum()
                    + (~np.isfinite(W)).sum())
    if p is not None:
        nonfinite += int((~np.isfinite(p[fl])).sum())
    return {"max_abs_div": mx, "nonfinite": nonfinite}


# --------------------------------------------------------------------------
# advection
# --------------------------------------------------------------------------
def sample(F: np.ndarray, x, y, z) -> np.ndarray:
    """Convection.fs:21-39 with the intended corner key ((ii+x')h, ...): trilinear
    weights [i - x + 1, x - i]; a corner outside the box contributes 0; the sum
    runs x', y', z' nested exactly as the reference's list comprehension."""
    nz, ny, nx = F.shape
    t = F.dtype.type
    Fp = np.zeros((nz + 2, ny + 2, nx + 2), dtype=F.dtype)
    Fp[1:-1, 1:-1, 1:-1] = F
    i = np.floor(x)
    j = np.floor(y)
    k = np.floor(z)
    wx = (i - x + t(1), x - i)
    wy = (j - y + t(1), y - j)
    wz = (k - z + t(1), z - k)
    ii = i.astype(np.int64)
    jj = j.astype(np.int64)
    kk = k.astype(np.int64)
    total = np.zeros(np.shape(x), dtype=F.dtype)
    for a in (0, 1):
        ia = np.clip(ii + a, -1, nx) + 1
        for b in (0, 1):
            jb = np.clip(jj + b, -1, ny) + 1
            for c in (0, 1):
                kc = np.clip(kk + c, -1, nz) + 1
                total = total + Fp[kc, jb, ia] * wx[a] * wy[b] * wz[c]
    return total


def velocity_at(U, V, W, x, y, z):
    """Staggered sample of (u, v, w) at index-space point (x, y, z) with cell
    centres on the integers and each component on its +face (SURVEY A.2
    intended): U[i,j,k] lives at (i + 1/2, j, k)."""
    t = U.dtype.type
    hlf = t(0.5)
    return (sample(U, x - hlf, y, z), sample(V, x, y - hlf, z),
            sample(W, x, y, z - hlf))


def advect(labels, U, V, W, dt: float, h: float, forward: bool = False):
    """Convection.fs:50-66 with intended semantics: for every
    Fluid cell and each of its three +faces, RK2-midpoint backtrace from the
    face centre and trilinear sample of that component at the departure point.
    Non-Fluid cells keep their velocity (Convection.fs:60-61 maps over markers).
    Jacobi: reads the old field only (double-buffered)."""
    nz, ny, nx = U.shape
    t = U.dtype.type
    kk, jj, ii = np.meshgrid(np.arange(nz, dtype=U.dtype),
                             np.arange(ny, dtype=U.dtype),
                             np.arange(nx, dtype=U.dtype), indexing="ij")
    sgn = t(1.0) if forward else t(-1.0)
    c = t(dt / h)
    hlf = t(0.5)
    fl = labels == FLUID
    out = []
    for comp, F in enumerate((U, V, W)):
        fx = ii + (hlf if comp == 0 else t(0))
        fy = jj + (hlf if comp == 1 else t(0))
        fz = kk + (hlf if comp == 2 else t(0))
        u0, v0, w0 = velocity_at(U, V, W, fx, fy, fz)
        mx = fx + sgn * hlf * c * u0
        my = fy + sgn * hlf * c * v0
        mz = fz + sgn * hlf * c * w0
        u1, v1, w1 = velocity_at(U, V, W, mx, my, mz)
        dx = fx + sgn * c * u1
        dy = fy + sgn * c * v1
        dz = fz + sgn * c * w1
        off = [t(0), t(0), t(0)]
        off[comp] = hlf
        new = sample(F, dx - off[0], dy - off[1], dz - off[2])
        out.append(np.where(fl, new, F).astype(U.dtype))
    return tuple(out)


def advect_scalar(labels, U, V, W, S, dt: float, h: float, forward: bool = False):
    """A cell-centred scalar carried by the flow: the RK2-midpoint backtrace of
    Convection.fs:50-57 started at the cell centre (i, j, k) instead of a face centre,
    then a trilinear sample of the scalar itself at the departure point with the weights
    of Convection.fs:21-39 (a corner outside the box contributes 0).  The reference's Cell
    has no scalar (Cell.fs:11-16); this defines north_star's "velocity and scalar fields"
    in the same intended semantics as ``advect``.  Only Fluid cells are updated; reads the
    OLD velocity field (Jacobi, like ``advect``)."""
    nz, ny, nx = S.shape
    t = S.dtype.type
    kk, jj, ii = np.meshgrid(np.arange(nz, dtype=S.dtype),
                             np.arange(ny, dtype=S.dtype),
                             np.arange(nx, dtype=S.dtype), indexing="ij")
    sgn = t(1.0) if forward else t(-1.0)
    c = t(dt / h)
    hlf = t(0.5)
    u0, v0, w0 = velocity_at(U, V, W, ii, jj, kk)
    mx = ii + sgn * hlf * c * u0
    my = jj + sgn * hlf * c * v0
    mz = kk + sgn * hlf * c * w0
    u1, v1, w1 = velocity_at(U, V, W, mx, my, mz)
    new = sample(S, ii + sgn * c * u1, jj + sgn * c * v1, kk + sgn * c * w1)
    return np.where(labels == FLUID, new, S).astype(S.dtype)


def add_forces(labels, U, V, W, dt: float, gravity=(0.0, -9.81, 0.0)):
    """Forces.fs:6-12: v += g dt on Fluid cells."""
    fl = labels == FLUID
    t = U.dtype.type
    return tuple(np.where(fl, F + t(g * dt), F).astype(U.dtype)
                 for F, g in zip((U, V, W), gravity))


# --------------------------------------------------------------------------
# whole step
# --------------------------------------------------------------------------
def project(labels, U, V, W, dt, h, rho, precond=PC_JACOBI, tol=1e-6,
            max_iters=10000, tau=0.0, direct=False, fixed_iters=None):
    """Pressure.calculate + Pressure.apply (Simulator.fs:88-89)."""
    b = rhs(labels, U, V, W, dt, h, rho)
    if direct:
        p = solve_direct(labels, b).astype(U.dtype)
        info = {"iters": 0, "b_inf": float(np.max(np.abs(b))) if b.size else 0.0}
    else:
        p, info = pcg(labels, b, precond, tol, max_iters, tau, U.dtype,
                      fixed_iters)
    U2, V2, W2 = grad_sub(labels, U, V, W, p, dt, h, rho)
    info.update(check(labels, U2, V2, W2, p))
    return U2, V2, W2, p, info


def step(labels, U, V, W, dt, h, rho, **kw):
    """advect -> project (Simulator.fs:82,88-89 without Forces/Viscosity)."""
    U1, V1, W1 = advect(labels, U, V, W, dt, h)
    return project(labels, U1, V1, W1, dt, h, rho, **kw)


# --------------------------------------------------------------------------
# sparse <-> dense packing (used by parity tests against sparse_ref)
# --------------------------------------------------------------------------
def pack_sparse(cells, h: int = 10, margin: int = 0):
    """Pack a ``sparse_ref.Grid.cells`` dict into dense arrays.  Returns
    (labels, U, V, W, P, origin) with origin = world cell index of box cell 0."""
    keys = list(cells.keys())
    lo = [min(k[a] for k in keys) // h - margin for a in range(3)]
    hi = [max(k[a] for k in keys) // h + margin for a in range(3)]
    nx, ny, nz = (hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1)
    labels = np.full((nz, ny, nx), ABSENT, dtype=np.uint8)
    U = np.zeros((nz, ny, nx))
    V = np.zeros((nz, ny, nx))
    W = np.zeros((nz, ny, nx))
    P = np.zeros((nz, ny, nx))
    for (x, y, z), c in cells.items():
        i, j, k = x // h - lo[0], y // h - lo[1], z // h - lo[2]
        labels[k, j, i] = c.media
        U[k, j, i], V[k, j, i], W[k, j, i] = c.velocity
        P[k, j, i] = c.pressure if c.pressure is not None else 0.0
    return labels, U, V, W, P, tuple(lo)


# --------------------------------------------------------------------------
# "next" rows of SURVEY 8f: viscosity (N1) and post-projection cleanup (N2)
# --------------------------------------------------------------------------
def viscosity(labels, U, V, W, dt: float, nu: float = 0.000001):
    """Viscosity.fs:12-36 as a Jacobi update (SURVEY Q10 kept: no 1/h^2, always
    -6 v, a neighbour contributes only its d-axis component, only if it is Fluid
    and that component has the sign of d).  Only Fluid cells change."""
    t = U.dtype.type
    fl = labels == FLUID
    k = t(nu * dt)
    out = []
    for axis, F in enumerate((U, V, W)):
        Fp = _nb(F, axis, +1, 0)            # neighbour on the + side
        Fm = _nb(F, axis, -1, 0)
        flp = _nb(fl, axis, +1, False)
        flm = _nb(fl, axis, -1, False)
        shared = np.where(flp & (Fp > 0), Fp, t(0)) + np.where(flm & (Fm < 0), Fm, t(0))
        lap = shared - t(6) * F
        out.append(np.where(fl, F + lap * k, F).astype(U.dtype))
    return tuple(out)


def zero_solid_velocities(labels, U, V, W):
    """Build.fs:144-160: for every Solid cell, the -neighbour's +face component is
    zeroed when that neighbour is non-solid and the component points into the solid
    (> 0); then Simulator.fs:106-109: every Solid cell's velocity is zeroed."""
    t = U.dtype.type
    ns = nonsolid(labels)
    solid = labels == SOLID
    out = []
    for axis, F in enumerate((U, V, W)):
        solid_p = _nb(solid, axis, +1, False)       # + neighbour is Solid
        G = np.where(ns & solid_p & (F > 0), t(0), F)
        G = np.where(solid, t(0), G)
        out.append(G.astype(U.dtype))
    return tuple(out)


def fluid_layers(labels, max_distance: int = 2):
    """BFS distance from the Fluid cells through existing cells (Build.fs:79-102
    layer numbers): 0 on Fluid, 1..max_distance on the buffer, -1 elsewhere."""
    layer = np.where(labels == FLUID, 0, -1).astype(np.int32)
    exists = labels != ABSENT
    for i in range(1, max_distance + 1):
        near = np.zeros(labels.shape, dtype=bool)
        for axis in range(3):
            for d in (-1, +1):
                near |= _nb(layer == i - 1, axis, d, False)
        layer = np.where((layer < 0) & exists & near, i, layer)
    return layer


def propagate_velocities(labels, U, V, W, max_distance: int = 2):
    """Build.fs:105-141, order-free reading.  The reference relabels cells while
    it enumerates its Dictionary (Layers.set m (Some (i - 1)) inside the loop), so
    which buffer cells are reached -- and from which neighbours -- depends on the
    .NET enumeration order; it does read only pre-update velocities (the result
    list is committed afterwards).  Here the layers are the BFS distances from the
    Fluid cells, every cell of layer 1..max_distance is updated once from the
    cells of the previous layer, all from the pre-update field:
      avg_axis = [own component if the + neighbour is in the previous layer]
               + [- neighbour's component if it is in the previous layer]
      component := avg_axis if the neighbour on the side the own component points
                   to (Build.fs:123-125) exists and is not Fluid, else unchanged.
    Coincides with the transcription on the stock 7-cell scene."""
    t = U.dtype.type
    layer = fluid_layers(labels, max_distance)
    exists = labels != ABSENT
    fl = labels == FLUID
    out = []
    for axis, F in enumerate((U, V, W)):
        G = F.copy()
        for i in range(1, max_distance + 1):
            cur = layer == i
            prev = layer == i - 1
            touched = np.zeros(labels.shape, dtype=bool)
            for a2 in range(3):
                touched |= _nb(prev, a2, +1, False) | _nb(prev, a2, -1, False)
            prev_p = _nb(prev, axis, +1, False)
            prev_m = _nb(prev, axis, -1, False)
            avg = np.where(prev_p, F, t(0)) + np.where(prev_m, _nb(F, axis, -1, 0), t(0))
            neg = F < 0
            ex_dir = np.where(neg, _nb(exists, axis, -1, False), _nb(exists, axis, +1, False))
            fl_dir = np.where(neg, _nb(fl, axis, -1, False), _nb(fl, axis, +1, False))
            G = np.where(cur & touched & ex_dir & ~fl_dir, avg, G)
        out.append(G.astype(U.dtype))
    return tuple(out)


def cleanup(labels, U, V, W):
    """Simulator.fs:100-110 (Build.cleanup): propagate, zero into-solid, zero solids."""
    U, V, W = propagate_velocities(labels, U, V, W)
    return zero_solid_velocities(labels, U, V, W)


# --------------------------------------------------------------------------
# "next" row N3 of SURVEY 8f: marker motion + cell relabelling (the step before the path)
# --------------------------------------------------------------------------
def nearest_cell_index(x: np.ndarray, h: float = 10.0) -> np.ndarray:
    """Coord.construct(float) (Coord.fs:49-67) as a world cell index: round half even to an
    integer (metres), then snap to the multiple of h with the .NET remainder rule (ties of
    +-h/2 go up for positive, towards zero for negative coordinates)."""
    hi = int(h)
    n = np.rint(np.asarray(x, dtype=np.float64)).astype(np.int64)
    r = np.fmod(n, hi)                                   # sign of the dividend
    snapped = np.where(r == 0, n, np.where(r >= hi // 2, n + (hi - r), n - r))
    return (snapped // hi).astype(np.int64)


def marker_cells(loc: np.ndarray, h: float, origin) -> np.ndarray:
    """(n, 3) box indices (i, j, k) of the cells the marker locations (metres) snap to."""
    idx = nearest_cell_index(loc, h)
    return idx - np.asarray(origin, dtype=np.int64)[None, :]


def marker_velocity(labels, U, V, W, cells: np.ndarray, sum_faces: bool = False) -> np.ndarray:
    """Simulator.get_velocity (Simulator.fs:24-35): the cell's +face velocity plus the
    -neighbours' components on the shared faces (absent neighbour: 0).  The reference adds
    the two faces (`sum_faces`, quirk); intended = their mean, the cell-centre velocity."""
    nz, ny, nx = labels.shape
    out = np.zeros((len(cells), 3))
    for a, F in enumerate((U, V, W)):
        i, j, k = cells[:, 0], cells[:, 1], cells[:, 2]
        own = F[k, j, i].astype(np.float64)
        bi, bj, bk = i - (a == 0), j - (a == 1), k - (a == 2)
        ok = (bi >= 0) & (bj >= 0) & (bk >= 0)
        back = np.where(ok, F[np.maximum(bk, 0), np.maximum(bj, 0),
                              np.maximum(bi, 0)].astype(np.float64), 0.0)
        back = np.where(ok & (labels[np.maximum(bk, 0), np.maximum(bj, 0),
                                     np.maximum(bi, 0)] != ABSENT), back, 0.0)
        out[:, a] = own + back if sum_faces else 0.5 * (own + back)
    return out


def move_markers(labels, U, V, W, loc, dt, h, origin, sum_faces: bool = False):
    """Simulator.move_markers (Simulator.fs:38-52): loc += dt * velocity(cell(loc)).
    Returns (new locations, old cells, new cells) -- cells as box indices (i, j, k)."""
    loc = np.asarray(loc, dtype=np.float64)
    old = marker_cells(loc, h, origin)
    v = marker_velocity(labels, U, V, W, old, sum_faces)
    new_loc = loc + v * dt
    return new_loc, old, marker_cells(new_loc, h, origin)


def relabel(labels, U, V, W, P, cells, world_lo, world_hi, max_distance: int = 2,
            lazy_buffer: bool = False):
    """Build.setup block of Simulator.advance (Simulator.fs:66-75, Build.fs:37-102) on
    the dense box: Fluid = in-world, non-Solid cells holding a marker (a Fluid cell without
    one turns into Air, velocity kept); BFS layers 1..max_distance grow from the non-solid
    cells of the previous layer through every neighbour -- an absent neighbour is created
    as Air (inside the world) or Solid (outside), zero velocity / pressure; cells that get
    no layer are deleted (ABSENT, zeroed).  `lazy_buffer` = the reference's literal rule:
    only cells that existed before this call act as BFS sources (Build.fs:83-86 reads the
    grid before the additions are committed), so the buffer grows one ring per step.
    `world_lo/hi`: inclusive box-index bounds of World.contains (Aabb.fs:13-17)."""
    nz, ny, nx = labels.shape
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    in_world = ((ii >= world_lo[0]) & (ii <= world_hi[0]) & (jj >= world_lo[1]) &
                (jj <= world_hi[1]) & (kk >= world_lo[2]) & (kk <= world_hi[2]))
    has_marker = np.zeros(labels.shape, dtype=bool)
    c = np.asarray(cells, dtype=np.int64).reshape(-1, 3)
    inside = ((c >= 0).all(1) & (c[:, 0] < nx) & (c[:, 1] < ny) & (c[:, 2] < nz))
    has_marker[c[inside, 2], c[inside, 1], c[inside, 0]] = True
    fluid = has_marker & in_world & (labels != SOLID)
    lab = np.where(fluid, FLUID, np.where(labels == FLUID, AIR, labels)).astype(np.uint8)
    existed = lab != ABSENT
    layer = np.where(fluid, 0, -1).astype(np.int32)
    for i in range(1, max_distance + 1):
        src = (layer == i - 1) & (lab != SOLID) & (lab != ABSENT)
        if lazy_buffer:
            src &= existed
        near = np.zeros(labels.shape, dtype=bool)
        for axis in range(3):
            for d in (-1, +1):
                near |= _nb(src, axis, d, False)
        take = near & (layer < 0) & (lab != FLUID)
        created = take & (lab == ABSENT)
        lab = np.where(created, np.where(in_world, AIR, SOLID), lab).astype(np.uint8)
        layer = np.where(take, i, layer)
    gone = layer < 0
    lab = np.where(gone, ABSENT, lab).astype(np.uint8)
    fresh = gone | ((labels == ABSENT) & ~gone)
    out = [np.where(fresh, F.dtype.type(0), F) for F in (U, V, W)]
    Pn = np.where(fresh | (lab == AIR), P.dtype.type(0), P)
    return lab, out[0], out[1], out[2], Pn


def move_cells(labels, U, V, W, P, old, new):
    """Grid.move_cells (Grid.fs:42-47), the reference's literal rule (quirk): the whole
    cell record -- media, velocity, pressure -- travels with its marker to the new
    coordinate and the old coordinate is deleted.  `moved` is built by prepending
    (Simulator.fs:45-50), i.e. processed in reverse marker order.  A target outside the
    box cannot be represented: raises IndexError."""
    lab = labels.copy()
    F = [U.copy(), V.copy(), W.copy(), P.copy()]
    pairs = [(tuple(o), tuple(n)) for o, n in zip(np.asarray(old), np.asarray(new))
             if tuple(o) != tuple(n)]
    nz, ny, nx = labels.shape
    for (oi, oj, ok), (ni, nj, nk) in reversed(pairs):
        if not (0 <= ni < nx and 0 <= nj < ny and 0 <= nk < nz):
            raise IndexError("marker left the dense box")
        if lab[ok, oj, oi] == ABSENT:                     # raw_get on a missing cell
            raise KeyError("Grid.raw_get: no cell at the marker's old coordinate")
        lab[nk, nj, ni] = lab[ok, oj, oi]
        for A in F:
            A[nk, nj, ni] = A[ok, oj, oi]
            A[ok, oj, oi] = 0
        lab[ok, oj, oi] = ABSENT
    return lab, F[0], F[1], F[2], F[3]
