"""Whole Simulator.advance steps on the device (markers -> labels -> advect -> forces -> viscosity ->
project -> cleanup) on the cfg3 dam-break: python profiles/ubench/advance_time.py [n] [steps] [per_cell]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
from splashy_b200 import Solver, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
per = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = scenes.dambreak(n, dtype=np.float32)
lab = sc["labels"]
k, j, i = np.nonzero(lab == 1)
rng = np.random.default_rng(0)
loc = (np.repeat(np.stack([i, j, k], 1), per, axis=0) + rng.uniform(-0.45, 0.45, (len(i) * per, 3))) * sc["h"]
print(f"n={n}: {len(i)} fluid cells, {len(loc)} markers")
with Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"], precond="mg", tol=1e-6) as s:
    s.upload(lab, sc["U"] * 0, sc["V"] * 0, sc["W"] * 0)
    s.set_markers(loc)
    for st in range(steps):
        t0 = time.perf_counter()
        s.timer_start()
        info = s.advance(sc["dt"], sanity=False)
        ms = s.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        print(f"step {st}: device {ms:.2f} ms (wall {wall:.2f}), iters {info['iters']}, solve {info['ms_solve']:.2f} ms, "
              f"advect {info['ms_advect']:.2f} ms, max|div| {info['max_abs_div']:.2e}")
    fl = int((s.download_media() == 1).sum())
    print("fluid cells now:", fl, "marker y mean", s.get_markers()[:, 1].mean() / sc["h"])
