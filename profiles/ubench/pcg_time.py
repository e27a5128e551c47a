"""Per-kernel timing of the Jacobi-PCG iteration at n^3 (one GPU, CUDA events inside the library):
python profiles/ubench/pcg_time.py [n] ; SPL_LIB=<variant .so> selects another build."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import os
import numpy as np
from splashy_b200 import Solver, scenes

M = int(os.environ.get("SPL_NSBUF", "16"))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = scenes.vortex(n, n, n, dtype=np.float32)
with Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"], precond="jacobi", tol=1e-6) as s:
    s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
    s.snapshot()
    best = None
    for _ in range(3):
        s.restore()
        a, b, x = s.profile_pcg(sc["dt"], 64)
        best = (a, b, x) if best is None or a < best[0] else best
a, b, x = best
cells = n ** 3
print(f"n={n} phaseA {a:.4f} ms ({cells * 17 / a / 1e6:.0f} GB/s)  phaseB {b:.4f} ms ({cells * 13 / b / 1e6:.0f} GB/s)  "
      f"xflush {x:.4f} ms (every {M})  iteration {a + b + x / M:.4f} ms")
