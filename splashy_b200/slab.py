"""z-slab decomposition of the dense box: one process per GPU.

Host-side logic only: which planes a rank owns, how the ncclUniqueId reaches
every rank, and how whole-box arrays are cut into / assembled from slabs.  The
data-path exchanges (halo planes, CG scalars) happen inside
``libsplashy_cuda.so`` over NCCL; ``torch.distributed`` is rendezvous plumbing.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass(frozen=True)
class SlabPlan:
    """Contiguous z-slabs; the remainder planes go to the first ranks (or, with
    ``uneven``, a deliberately lopsided split used by the tests)."""
    nz_global: int
    world: int
    uneven: bool = False

    def __post_init__(self):
        if self.world < 1 or self.nz_global < self.world:
            raise ValueError(f"cannot cut {self.nz_global} planes into {self.world} slabs")

    def sizes(self) -> List[int]:
        base, rem = divmod(self.nz_global, self.world)
        sizes = [base + (1 if r < rem else 0) for r in range(self.world)]
        if self.uneven and self.world > 1 and sizes[-1] > 1:
            sizes[0] += 1          # move one plane from the last slab to the first
            sizes[-1] -= 1
        return sizes

    def nz(self, rank: int) -> int:
        return self.sizes()[rank]

    def z0(self, rank: int) -> int:
        return sum(self.sizes()[:rank])

    def owner(self, k: int) -> int:
        acc = 0
        for r, n in enumerate(self.sizes()):
            acc += n
            if k < acc:
                return r
        raise IndexError(k)

    def neighbours(self, rank: int) -> Tuple[Optional[int], Optional[int]]:
        """(rank owning the planes below, rank owning the planes above)."""
        return (rank - 1 if rank > 0 else None,
                rank + 1 if rank < self.world - 1 else None)

    def cut(self, arrays: Dict[str, np.ndarray], rank: int) -> Dict[str, np.ndarray]:
        z0, nz = self.z0(rank), self.nz(rank)
        return {k: np.ascontiguousarray(a[z0:z0 + nz]) for k, a in arrays.items()}

    def assemble(self, parts: List[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def share_unique_id(make_id, rank: int, world: int) -> Optional[bytes]:
    """Rank 0 calls ``make_id()`` (``spl_nccl_unique_id``); every rank returns
    the same 128 bytes.  Uses the default torch.distributed group (nccl or
    gloo)."""
    if world == 1:
        return None
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) != 128:
        raise RuntimeError("ncclUniqueId broadcast failed")
    return bytes(box[0])
