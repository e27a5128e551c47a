"""Timing of the advection stage alone at n^3 (one GPU):
   python profiles/ubench/advect_time.py [n] [n_scalars]
Environment: SPL_ADV_TMA=0 (L1-gather kernel), SPL_ADV_ZCHUNK=planes per CTA, SPL_ADV_FUSE=0."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
from splashy_b200 import Solver, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = scenes.vortex(n, n, n, dtype=np.float32)
with Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"]) as s:
    s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
    if nsc:
        s.set_scalars([scenes.blob(n, n, n)] * nsc)
    s.snapshot()
    best = 1e9
    for _ in range(5):
        s.restore()
        info = s.advect(sc["dt"])
        best = min(best, info["ms_advect"])
    chk = float(np.abs(s.download()["U"][n // 2, n // 2, ::7]).sum())
    chs = float(np.abs(s.get_scalars()[0][n // 2, n // 2, ::7]).sum()) if nsc else 0.0
b = 25 + 8 * nsc
print(f"n={n} scalars={nsc} advect {best:.3f} ms ({n ** 3 * b / best / 1e6:.0f} GB/s algorithmic, "
      f"{b} B/cell)  launches {info['kernel_launches']}  checksum {chk:.6f} {chs:.6f}")
