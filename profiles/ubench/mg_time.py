"""Quick timing of the multigrid step at n^3 (one GPU): python profiles/ubench/mg_time.py [n] [iters]
SPL_LIB=<path to a library variant> selects a differently compiled libsplashy_cuda."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
from splashy_b200 import Solver, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 14
sc = scenes.vortex(n, n, n, dtype=np.float32)
with Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"], precond="mg", tol=1e-6) as s:
    s.upload(sc["labels"], sc["U"], sc["V"], sc["W"])
    s.snapshot()
    s.set_fixed_iters(iters)
    best = None
    for _ in range(4):
        s.restore()
        info = s.step(sc["dt"])
        best = info if best is None or info["ms_solve"] < best["ms_solve"] else best
    s.set_fixed_iters(0)
    s.restore()
    t = s.step(sc["dt"])
print(f"n={n} fixed {iters} it: ms_solve {best['ms_solve']:.2f} ({best['ms_solve'] / iters:.3f}/it) "
      f"ms_total {best['ms_total']:.2f} | to 1e-6: {t['iters']} it, {t['ms_total']:.2f} ms, "
      f"max|div| {t['max_abs_div']:.2e}, r_inf/b_inf {t['r_inf'] / t['b_inf']:.2e}")
