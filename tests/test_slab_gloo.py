"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the z-slab protocol.

The data path of the product runs inside libsplashy_cuda.so over NCCL; what can
be checked without a GPU is the decomposition itself: the slab plan, the
ncclUniqueId hand-out, and that the protocol the library implements -- one
plane of the new search direction exchanged per iteration, per-rank partial dot
products all-gathered and summed in rank order -- reproduces the whole-box PCG
of the oracle exactly (same iterates, same iteration count).  The local compute
is the oracle's numpy code (tests may use the oracle).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_ref as dr
from splashy_b200 import scenes
from splashy_b200.slab import SlabPlan, share_unique_id


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def exchange_plane(lo_send, hi_send, rank, world):
    """Send my bottom / top plane to the slab below / above; return the planes
    received from them (zeros at the global ends), like exchange_halo()."""
    shape = lo_send.shape
    lo_recv = torch.zeros(shape, dtype=torch.float64)
    hi_recv = torch.zeros(shape, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(lo_send.copy()), rank - 1))
        reqs.append(dist.irecv(lo_recv, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(hi_send.copy()), rank + 1))
        reqs.append(dist.irecv(hi_recv, rank + 1))
    for r in reqs:
        r.wait()
    return lo_recv.numpy(), hi_recv.numpy()


def gather_slots(value, world):
    """ncclAllGather of one double per rank, summed in rank order."""
    slots = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(slots, torch.tensor([value], dtype=torch.float64))
    return [float(s.item()) for s in slots]


def slab_pcg(plan, rank, world, labels_g, b_g, tol, max_iters=5000):
    """The library's slab protocol with numpy local compute (Jacobi-PCG)."""
    z0, nz = plan.z0(rank), plan.nz(rank)
    # media halo: one plane from each neighbour (ABSENT at the global ends)
    lab = np.full((nz + 2,) + labels_g.shape[1:], dr.ABSENT, dtype=np.uint8)
    lo, hi = max(z0 - 1, 0), min(z0 + nz + 1, labels_g.shape[0])
    lab[(lo - (z0 - 1)):(hi - (z0 - 1))] = labels_g[lo:hi]
    codes = dr.stencil_codes(lab)[1:-1]          # codes of the local planes
    fluid = (codes & dr.B_FLUID) != 0
    N = dr.diag_count(codes)
    invd = np.where(fluid & (N > 0), 1.0 / np.maximum(N, 1), 0.0)
    r = np.where(fluid, b_g[z0:z0 + nz], 0.0)
    x = np.zeros_like(r)
    s = np.zeros_like(r)
    rho_prev = float("inf")
    rho = sum(gather_slots(float(np.dot(r.ravel(), (r * invd).ravel())), world))
    bmax = max(gather_slots(float(np.abs(r).max()), world))
    rmax = bmax
    it = 0
    while it < max_iters:
        if rmax <= tol * bmax or rho == 0.0:
            break
        it += 1
        beta = rho / rho_prev
        s = r * invd + beta * s
        below, above = exchange_plane(s[0], s[-1], rank, world)
        ext = np.concatenate([below[None], s, above[None]], axis=0)
        ecodes = np.concatenate([np.zeros_like(codes[:1]), codes, np.zeros_like(codes[:1])], 0)
        q = dr.apply_A(ecodes, ext)[1:-1]
        sq = sum(gather_slots(float(np.dot(s.ravel(), q.ravel())), world))
        alpha = rho / sq
        x += alpha * s
        r -= alpha * q
        rho_prev = rho
        rho = sum(gather_slots(float(np.dot(r.ravel(), (r * invd).ravel())), world))
        rmax = max(gather_slots(float(np.abs(r).max()), world))
    return x, it, rmax


def worker(rank, world, port, uneven, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nid = share_unique_id(lambda: bytes(range(128)), rank, world)
        assert nid == bytes(range(128))
        sc = scenes.vortex(12, 10, 5 * world + 1, dtype=np.float64)
        plan = SlabPlan(sc["labels"].shape[0], world, uneven=uneven)
        b = dr.rhs(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
        x, it, rmax = slab_pcg(plan, rank, world, sc["labels"], b, 1e-10)
        np.save(os.path.join(out_dir, f"x{rank}.npy"), x)
        np.save(os.path.join(out_dir, f"meta{rank}.npy"), np.array([it, rmax]))
        # per-slab scene generation equals cutting the whole box
        z0, nz = plan.z0(rank), plan.nz(rank)
        part = scenes.vortex(12, 10, nz, z0=z0, nz_global=sc["labels"].shape[0], dtype=np.float64)
        for k in ("labels", "U", "V", "W"):
            assert np.array_equal(part[k], sc[k][z0:z0 + nz])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,uneven", [(2, False), (2, True), (3, False)])
def test_slab_protocol_reproduces_whole_box_pcg(tmp_path, world, uneven):
    port = free_port()
    mp.spawn(worker, args=(world, port, uneven, str(tmp_path)), nprocs=world, join=True)
    sc = scenes.vortex(12, 10, 5 * world + 1, dtype=np.float64)
    b = dr.rhs(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"])
    x_ref, info = dr.pcg(sc["labels"], b, dr.PC_JACOBI, tol=1e-10)
    parts = [np.load(tmp_path / f"x{r}.npy") for r in range(world)]
    metas = [np.load(tmp_path / f"meta{r}.npy") for r in range(world)]
    x = np.concatenate(parts, axis=0)
    assert all(int(m[0]) == int(metas[0][0]) for m in metas)      # same stop decision
    assert abs(int(metas[0][0]) - info["iters"]) <= 1
    np.testing.assert_allclose(x, x_ref, rtol=1e-9, atol=1e-9 * np.abs(x_ref).max())


def test_slab_plan_partition():
    for nzg, world in ((16, 2), (17, 4), (9, 8), (512, 8), (5, 5)):
        for uneven in (False, True):
            p = SlabPlan(nzg, world, uneven)
            sizes = p.sizes()
            assert sum(sizes) == nzg and min(sizes) >= 1
            assert [p.z0(r) for r in range(world)] == list(np.cumsum([0] + sizes[:-1]))
            for k in range(nzg):
                r = p.owner(k)
                assert p.z0(r) <= k < p.z0(r) + p.nz(r)
            assert p.neighbours(0)[0] is None and p.neighbours(world - 1)[1] is None
    arrays = {"a": np.arange(7 * 2 * 3).reshape(7, 2, 3)}
    p = SlabPlan(7, 3)
    assert np.array_equal(p.assemble([p.cut(arrays, r) for r in range(3)])["a"], arrays["a"])
    with pytest.raises(ValueError):
        SlabPlan(3, 4)
