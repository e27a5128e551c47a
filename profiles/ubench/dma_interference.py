"""How much do PCIe copies slow an HBM-bound step down?  One context runs the device-resident
512^3 step; a second context keeps one direction of the link busy from a host thread
(spl_upload_async / spl_download_async + spl_sync).  Prints the step's device time alone, with
H2D traffic, with D2H traffic, and with both.
usage: python profiles/ubench/dma_interference.py [n] [steps]"""
import ctypes as C
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from splashy_b200 import PinnedArray, Solver, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = scenes.vortex(n, n, n, dtype=np.float32)
shape = sc["labels"].shape
ptr = lambda a: C.c_void_p(a.array.ctypes.data)
hin = {k: PinnedArray(shape, np.float32) for k in "UVW"}
hlab = PinnedArray(shape, np.uint8)
hout = {k: PinnedArray(shape, np.float32) for k in "UVWP"}
for k in "UVW":
    hin[k].array[...] = sc[k]
hlab.array[...] = sc["labels"]
mk = lambda: Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"])
a, b, c = mk(), mk(), mk()
for s in (a, b, c):
    s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
a.snapshot()
a.set_fixed_iters(200)


def run(label, up, down):
    stop = threading.Event()
    moved = [0, 0]

    def h2d():
        while not stop.is_set():
            b.upload_async_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
            b.sync()
            moved[0] += 1

    def d2h():
        while not stop.is_set():
            c.download_async_ptrs(ptr(hout["U"]), ptr(hout["V"]), ptr(hout["W"]), ptr(hout["P"]))
            c.sync()
            moved[1] += 1

    ths = [threading.Thread(target=f) for f, on in ((h2d, up), (d2h, down)) if on]
    for t in ths:
        t.start()
    time.sleep(0.05)
    ms = []
    t0 = time.perf_counter()
    for _ in range(steps):
        a.restore()
        ms.append(a.step(sc["dt"])["ms_total"])
    wall = time.perf_counter() - t0
    stop.set()
    for t in ths:
        t.join()
    gb = (moved[0] * (n ** 3 * 13) + moved[1] * (n ** 3 * 16)) / 1e9
    print(f"{label:12s} step {np.mean(ms):7.2f} ms   copies: {moved[0]} up, {moved[1]} down "
          f"= {gb / wall:5.1f} GB/s over the link")


run("alone", False, False)
run("with H2D", True, False)
run("with D2H", False, True)
run("with both", True, True)
run("alone", False, False)
