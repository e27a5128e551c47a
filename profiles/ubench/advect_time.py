"""Timing of the advection stage alone at n^3 (one GPU): python profiles/ubench/advect_time.py [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
from splashy_b200 import Solver, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = scenes.vortex(n, n, n, dtype=np.float32)
with Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"]) as s:
    s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
    s.snapshot()
    best = 1e9
    for _ in range(5):
        s.restore()
        info = s.advect(sc["dt"])
        best = min(best, info["ms_advect"])
    chk = float(np.abs(s.download()["U"][n // 2, n // 2, ::7]).sum())
print(f"n={n} advect {best:.3f} ms ({n ** 3 * 25 / best / 1e6:.0f} GB/s algorithmic)  checksum {chk:.6f}")
