"""N > 1 host logic of the multigrid preconditioner on CPU: world_size-2 / -4 gloo runs of the
z-slab V-cycle protocol the library implements (spl_api.cu Impl::vcycle with world > 1):

* every level coarsens in lock step (a rank holds whole coarse planes), the labels / stencil
  codes of a level see one halo plane of the neighbouring slab;
* after every pass that writes a field a later pass reads through the z stencil, one plane
  travels to each neighbour (exchange_level);
* the coarsest level is gathered (all_gather) into one global grid that every rank solves
  redundantly, and each rank takes its own planes back.

With the same sweep counts the result equals the single-process oracle
(oracle/dense_ref.Multigrid on the whole box) to round-off, and slab MG-PCG takes the same
iterations.  Local compute is the oracle's numpy code.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_ref as dr
from splashy_b200 import scenes
from test_slab_gloo import exchange_plane, free_port, gather_slots

NU, CS = 2, 30


def with_halo(a, rank, world, fill):
    lo, hi = exchange_plane(a[0].astype(np.float64), a[-1].astype(np.float64), rank, world)
    ext = np.concatenate([lo[None], a.astype(np.float64), hi[None]], axis=0)
    if rank == 0:
        ext[0] = fill
    if rank == world - 1:
        ext[-1] = fill
    return ext


class SlabMultigrid:
    def __init__(self, labels_local, rank, world):
        self.rank, self.world = rank, world
        self.om = dr.mg_weights(NU)
        self.levels = []
        lab, scale = labels_local, 1.0
        while True:
            ext = with_halo(lab, rank, world, dr.ABSENT).astype(np.uint8)
            codes_ext = dr.stencil_codes(ext)
            codes = codes_ext[1:-1]
            fluid = (codes & dr.B_FLUID) != 0
            N = dr.diag_count(codes)
            invn = np.where(fluid & (N > 0), 1.0 / np.maximum(N, 1), 0.0)
            self.levels.append(dict(lab=lab, codes_ext=codes_ext, fluid=fluid, invn=invn,
                                    scale=scale))
            nz, ny, nx = lab.shape
            dims = [d for d in (nx, ny, nz) if d > 1]
            if min(dims) <= 8 or nz % 2:          # lock step: whole coarse planes per rank
                break
            lab = dr.mg_coarsen_labels(lab)
            scale *= 0.25
        # global coarsest grid: labels gathered once (ncclAllGather of the codes in the library)
        k = self.levels[-1]["lab"]
        parts = [torch.zeros(k.shape, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(k)))
        gl = np.concatenate([p.numpy() for p in parts], 0)
        gc = dr.stencil_codes(gl)
        gf = (gc & dr.B_FLUID) != 0
        gN = dr.diag_count(gc)
        self.g = dict(codes=gc, fluid=gf, invn=np.where(gf & (gN > 0), 1.0 / np.maximum(gN, 1), 0.0))

    def lap(self, L, u):
        ext = with_halo(u, self.rank, self.world, 0.0)      # exchange_level after the pass
        return dr.apply_A(L["codes_ext"], ext)[1:-1]

    def smooth(self, L, u, b, first, reverse):
        for s in range(NU):
            w = L["invn"] * self.om[NU - 1 - s if reverse else s]
            if first and s == 0:
                u = w * (b / L["scale"])
            else:
                u = u + w * (b / L["scale"] - self.lap(L, u))
            u = np.where(L["fluid"], u, 0.0)
        return u

    def coarsest(self, b):
        L = self.levels[-1]
        parts = [torch.zeros(b.shape, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(b)))
        gb = np.concatenate([p.numpy() for p in parts], 0)
        g = self.g
        gb = np.where(g["fluid"], gb, 0.0)
        u = None
        for s in range(CS):
            w = g["invn"] * self.om[s % NU]
            if s == 0:
                u = w * (gb / L["scale"])
            else:
                u = u + w * (gb / L["scale"] - dr.apply_A(g["codes"], u))
            u = np.where(g["fluid"], u, 0.0)
        nz = b.shape[0]
        return u[self.rank * nz:(self.rank + 1) * nz]

    def vcycle(self, b, lev=0):
        L = self.levels[lev]
        b = np.where(L["fluid"], b, 0.0)
        if lev == len(self.levels) - 1:
            return self.coarsest(b)
        u = self.smooth(L, None, b, True, False)
        r = np.where(L["fluid"], b - L["scale"] * self.lap(L, u), 0.0)
        nz, ny, nx = r.shape
        cz, cy, cx = self.levels[lev + 1]["lab"].shape
        pad = np.zeros((cz * 2, cy * 2, cx * 2))
        pad[:nz, :ny, :nx] = r
        rc = pad.reshape(cz, 2, cy, 2, cx, 2).sum(axis=(1, 3, 5)) * 0.125
        ec = self.vcycle(rc, lev + 1)
        pe = np.repeat(np.repeat(np.repeat(ec, 2, 0), 2, 1), 2, 2)[:nz, :ny, :nx]
        u = u + np.where(L["fluid"], pe, 0.0)
        return self.smooth(L, u, b, False, True)


def slab_mg_pcg(mg, codes_ext, b, tol, world):
    fluid = mg.levels[0]["fluid"]
    r = np.where(fluid, b, 0.0)
    x = np.zeros_like(r)
    z = mg.vcycle(r)
    s = z.copy()
    rho = sum(gather_slots(float((r * z).sum()), world))
    bmax = max(gather_slots(float(np.abs(r).max()), world))
    for it in range(1, 200):
        q = mg.lap(mg.levels[0], s)
        alpha = rho / sum(gather_slots(float((s * q).sum()), world))
        x += alpha * s
        r -= alpha * q
        if max(gather_slots(float(np.abs(r).max()), world)) <= tol * bmax:
            return x, it
        z = mg.vcycle(r)
        rho2 = sum(gather_slots(float((r * z).sum()), world))
        s = z + (rho2 / rho) * s
        rho = rho2
    return x, 200


def scene(world):
    return scenes.vortex(16, 12, 16 * world, dtype=np.float64)


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = scene(world)
        nz = 16
        lab = sc["labels"][rank * nz:(rank + 1) * nz]
        mg = SlabMultigrid(lab, rank, world)
        rng = np.random.default_rng(5)
        r_g = np.where(sc["labels"] == dr.FLUID, rng.uniform(-1, 1, sc["labels"].shape), 0.0)
        z = mg.vcycle(r_g[rank * nz:(rank + 1) * nz])
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z)
        b = dr.rhs(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
        x, it = slab_mg_pcg(mg, None, b[rank * nz:(rank + 1) * nz], 1e-9, world)
        np.save(os.path.join(out_dir, f"x{rank}.npy"), x)
        np.save(os.path.join(out_dir, f"it{rank}.npy"), np.array([it, len(mg.levels)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_vcycle_protocol_equals_the_whole_box_multigrid(tmp_path, world):
    port = free_port()
    mp.spawn(worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sc = scene(world)
    rng = np.random.default_rng(5)
    r_g = np.where(sc["labels"] == dr.FLUID, rng.uniform(-1, 1, sc["labels"].shape), 0.0)
    ref = dr.Multigrid(sc["labels"], nu=NU, coarse_sweeps=CS)
    its = [np.load(tmp_path / f"it{r}.npy") for r in range(world)]
    assert all(int(i[1]) == len(ref.levels) for i in its)          # same hierarchy depth
    z = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)], 0)
    z_ref = ref.vcycle(r_g)
    np.testing.assert_allclose(z, z_ref, rtol=0, atol=1e-12 * np.abs(z_ref).max())
    b = dr.rhs(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
    x_ref, info = dr.pcg(sc["labels"], b, dr.PC_MG, tol=1e-9)
    x = np.concatenate([np.load(tmp_path / f"x{r}.npy") for r in range(world)], 0)
    assert all(int(i[0]) == int(its[0][0]) for i in its)
    assert abs(int(its[0][0]) - info["iters"]) <= 1
    np.testing.assert_allclose(x, x_ref, rtol=0, atol=1e-7 * np.abs(x_ref).max())
