"""Diagnostic (torchrun): where do slabs and one GPU differ?  Compares, on rank 0, the advected
field and the field after the whole step, and prints the first mismatching cells."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from splashy_b200 import Solver, nccl_unique_id, scenes  # noqa: E402

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx, ny, nz = 128, 64, int(os.environ.get("DIAG_NZ", "48"))
nzg = nz * world
whole = scenes.vortex(nx, ny, nzg, dtype=np.float32)
part = scenes.slab_of(whole, rank * nz, nz)
box = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)


def gather(got, keys):
    mine = torch.from_numpy(np.stack([got[k] for k in keys])).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts, dim=1).cpu().numpy()


with Solver(nx, ny, nz, dtype=np.float32, h=whole["h"], rho=whole["rho"], device=local, rank=rank,
            world=world, nz_global=nzg, z0=rank * nz, nccl_id=box[0]) as s:
    s.upload(part["labels"], part["U"], part["V"], part["W"])
    s.snapshot()
    s.advect(whole["dt"])
    adv = gather(s.download(), "UVW")
    s.restore()
    s.set_fixed_iters(40)
    info = s.step(whole["dt"])
    stp = gather(s.download(), "UVWP")
if rank == 0:
    with Solver(nx, ny, nzg, dtype=np.float32, h=whole["h"], rho=whole["rho"], device=local) as s1:
        s1.upload(whole["labels"], whole["U"], whole["V"], whole["W"])
        s1.snapshot()
        s1.advect(whole["dt"])
        a1 = s1.download()
        s1.restore()
        s1.set_fixed_iters(40)
        s1.step(whole["dt"])
        o1 = s1.download()
    for name, got, ref, keys in (("advect", adv, a1, "UVW"), ("step", stp, o1, "UVWP")):
        for i, k in enumerate(keys):
            d = np.argwhere(got[i] != ref[k])
            print(f"{name} {k}: {len(d)} cells differ", flush=True)
            for (z, y, x) in d[:6]:
                print(f"    (k={z} j={y} i={x}) slab {got[i][z, y, x]!r} one {ref[k][z, y, x]!r} "
                      f"label {whole['labels'][z, y, x]} slab-plane {z % nz}", flush=True)
dist.barrier()
dist.destroy_process_group()
