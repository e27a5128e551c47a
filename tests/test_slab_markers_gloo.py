"""Marker stage on z-slabs, host logic on CPU (world_size-2 and 3 gloo runs).

The protocol the library implements (spl_markers.cuh, spl_move_markers / spl_relabel with
world > 1): every rank holds ALL marker locations; a marker is moved by the rank whose slab
holds its cell, the others contribute zeros, and one all-reduce (sum) hands every rank the new
locations -- exact, because a single term is non-zero; the relabelling rings read the
neighbouring slab's boundary plane of the labels / layers through a one-plane halo exchanged
before each ring.  The local compute is the oracle's numpy code; the result must equal the
whole-box oracle (dense_ref.move_markers / relabel) bit for bit."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_ref as dr
from splashy_b200.slab import SlabPlan
from test_slab_gloo import free_port

H, DT = 10, 0.016671


def scene(world):
    rng = np.random.default_rng(11)
    n, nzg = 20, 9 * world + 2
    lab = np.full((nzg, n, n), dr.ABSENT, dtype=np.uint8)
    lab[3:nzg - 3, 3:17, 3:17] = dr.AIR
    lab[3:nzg - 3, 3:5, 3:17] = dr.SOLID
    cells = np.stack([rng.integers(5, 15, 40 * world), rng.integers(5, 15, 40 * world),
                      rng.integers(5, nzg - 5, 40 * world)], 1)
    lab[cells[:, 2], cells[:, 1], cells[:, 0]] = dr.FLUID
    ex = lab != dr.ABSENT
    U, V, W = (np.where(ex, rng.uniform(-600, 600, lab.shape), 0.0) for _ in range(3))
    P = np.where(lab == dr.FLUID, rng.uniform(0, 1e4, lab.shape), 0.0)
    org = (-10, -10, -10)
    loc = (cells + np.array(org)) * float(H) + rng.uniform(-4.9, 4.9, cells.shape)
    return lab, U, V, W, P, org, loc, (3, 3, 3), (15, 18, nzg - 5)


def swap_planes(lo_send, hi_send, rank, world, fill):
    """One plane to / from each slab neighbour (exchange_halo); `fill` at the global ends."""
    lo_recv = torch.full(lo_send.shape, fill, dtype=torch.int32)
    hi_recv = torch.full(hi_send.shape, fill, dtype=torch.int32)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(torch.from_numpy(lo_send.astype(np.int32)), rank - 1),
                 dist.irecv(lo_recv, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(torch.from_numpy(hi_send.astype(np.int32)), rank + 1),
                 dist.irecv(hi_recv, rank + 1)]
    for r in reqs:
        r.wait()
    return lo_recv.numpy(), hi_recv.numpy()


def slab_move(plan, rank, world, lab, U, V, W, loc, org):
    """k_move_markers + the all-reduce: the owner of a marker's cell moves it."""
    z0, nz = plan.z0(rank), plan.nz(rank)
    cells = dr.marker_cells(loc, H, org)
    nzg, ny, nx = lab.shape
    inside = ((cells >= 0).all(1) & (cells[:, 0] < nx) & (cells[:, 1] < ny) & (cells[:, 2] < nzg))
    mine = inside & (cells[:, 2] >= z0) & (cells[:, 2] < z0 + nz)
    keep = ~inside & (rank == 0)                       # outside the box: rank 0 keeps them
    # local fields with the one-plane halo below (the -z face of plane 0)
    lo = max(z0 - 1, 0)
    sl = slice(lo, z0 + nz)
    local = cells[mine] - np.array([0, 0, lo])
    vel = dr.marker_velocity(lab[sl], U[sl], V[sl], W[sl], local)
    out = np.zeros_like(loc)
    out[mine] = loc[mine] + vel * DT
    out[keep] = loc[keep]
    t = torch.from_numpy(out)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)            # one non-zero term per entry: exact
    moved = torch.tensor([int((dr.marker_cells(out[mine], H, org) != cells[mine]).any(1).sum())])
    dist.all_reduce(moved, op=dist.ReduceOp.SUM)        # the counters are summed over the ranks
    return t.numpy(), int(moved.item())


def slab_relabel(plan, rank, world, lab_g, new_cells, lo_w, hi_w):
    """k_relabel_init / step x 2 / finish on one slab with plane halos between the rings."""
    z0, nz = plan.z0(rank), plan.nz(rank)
    nzg, ny, nx = lab_g.shape
    labels = lab_g[z0:z0 + nz]
    kk, jj, ii = np.meshgrid(np.arange(z0, z0 + nz), np.arange(ny), np.arange(nx), indexing="ij")
    in_world = ((ii >= lo_w[0]) & (ii <= hi_w[0]) & (jj >= lo_w[1]) & (jj <= hi_w[1]) &
                (kk >= lo_w[2]) & (kk <= hi_w[2]))
    has = np.zeros(labels.shape, dtype=bool)
    c = new_cells[(new_cells[:, 2] >= z0) & (new_cells[:, 2] < z0 + nz)]
    has[c[:, 2] - z0, c[:, 1], c[:, 0]] = True
    fluid = has & in_world & (labels != dr.SOLID)
    lab = np.where(fluid, dr.FLUID, np.where(labels == dr.FLUID, dr.AIR, labels)).astype(np.uint8)
    layer = np.where(fluid, 0, -1).astype(np.int32)
    for ring in (1, 2):
        lay_lo, lay_hi = swap_planes(layer[0], layer[-1], rank, world, -1)
        lab_lo, lab_hi = swap_planes(lab[0], lab[-1], rank, world, int(dr.ABSENT))
        elay = np.concatenate([lay_lo[None], layer, lay_hi[None]], 0)
        elab = np.concatenate([lab_lo[None], lab, lab_hi[None]], 0).astype(np.uint8)
        src = (elay == ring - 1) & (elab != dr.SOLID) & (elab != dr.ABSENT)
        near = np.zeros(elab.shape, dtype=bool)
        for axis in range(3):
            for d in (-1, +1):
                near |= dr._nb(src, axis, d, False)
        near = near[1:-1]
        take = near & (layer < 0) & (lab != dr.FLUID)
        created = take & (lab == dr.ABSENT)
        lab = np.where(created, np.where(in_world, dr.AIR, dr.SOLID), lab).astype(np.uint8)
        layer = np.where(take, ring, layer)
    return np.where(layer < 0, dr.ABSENT, lab).astype(np.uint8)


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lab, U, V, W, P, org, loc, lo_w, hi_w = scene(world)
        plan = SlabPlan(lab.shape[0], world)
        new_loc, moved = slab_move(plan, rank, world, lab, U, V, W, loc, org)
        new_cells = dr.marker_cells(new_loc, H, org)
        out = slab_relabel(plan, rank, world, lab, new_cells, lo_w, hi_w)
        np.save(os.path.join(out_dir, f"loc{rank}.npy"), new_loc)
        np.save(os.path.join(out_dir, f"lab{rank}.npy"), out)
        np.save(os.path.join(out_dir, f"moved{rank}.npy"), np.array([moved]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_marker_stage_on_slabs_equals_the_whole_box(tmp_path, world):
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    lab, U, V, W, P, org, loc, lo_w, hi_w = scene(world)
    want_loc, old, new = dr.move_markers(lab, U, V, W, loc, DT, H, org)
    want = dr.relabel(lab, U, V, W, P, new, lo_w, hi_w)
    locs = [np.load(tmp_path / f"loc{r}.npy") for r in range(world)]
    for l in locs:                                     # every rank ends with every location
        assert np.array_equal(l, want_loc)
    got = np.concatenate([np.load(tmp_path / f"lab{r}.npy") for r in range(world)], 0)
    assert np.array_equal(got, want[0])
    moved = [int(np.load(tmp_path / f"moved{r}.npy")[0]) for r in range(world)]
    assert len(set(moved)) == 1 and moved[0] == int((old != new).any(1).sum())
    plan = SlabPlan(lab.shape[0], world)
    crossed = sum(plan.owner(int(a)) != plan.owner(int(b)) for a, b in zip(old[:, 2], new[:, 2]))
    assert crossed > 0                                  # some markers change slab
