"""Per-kernel CUDA-event times of the PCG iteration on z-slabs (run under torchrun):
   python -m torch.distributed.run --nproc-per-node N ... profiles/ubench/pcg_phase_time.py [n]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch, torch.distributed as dist
from splashy_b200 import Solver, nccl_unique_id, scenes

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
nid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    nid = box[0]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = scenes.vortex(n, n, n, z0=rank * n, nz_global=n * world, dtype=np.float32)
s = Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"], device=local, rank=rank, world=world,
           nz_global=n * world, z0=rank * n, nccl_id=nid)
s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
for _ in range(2):
    a, b, x = s.profile_pcg(sc["dt"], 64)
print(f"rank {rank}/{world} FUSED={os.environ.get('SPL_FUSED','')} MODE={os.environ.get('SPL_FUSED_MODE','')}: "
      f"phase A {a*1e3:.1f} us, gather A + phase B {b*1e3:.1f} us, x flush {x*1e3:.1f} us", flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
