#!/usr/bin/env python
"""bench.py -- advect + project throughput of splashy_b200 on N B200s.

Contract (see DESIGN.md "Measurement"):

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)

Workload (config.workload): cfg4/cfg5 ``vortex_512`` -- a 512 x 512 x 512
z-slab per GPU (weak scaling: the global box is 512 x 512 x 512*N), ABC vortex +
noise at CFL 2, Solid shell, Air plane; one step = semi-Lagrangian advection of
(u, v, w) + divergence/RHS + `--pcg-iters` (default 200) Jacobi-PCG
iterations + gradient subtraction + divergence check.  `value` =
cell-updates/s over all GPUs with inputs resident in HBM; `e2e` = the same
through spl_upload / spl_step / spl_download with pinned HOST buffers inside the
timed region.  A separate run-to-tolerance step reports the iteration count.
Inputs (3 x 512 MiB fields per GPU) are far larger than the 126 MB L2, so no
explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "advect+project cell-updates/s"
UNIT = "cell-updates/s"
NSBUF = int(os.environ.get("SPL_NSBUF", "16"))   # x is flushed every NSBUF iterations
BYTES_PHASE_A = 17     # r 4 + s 4 + code 1 -> s' 4 + q 4                 
BYTES_PHASE_B = 13     # r 4 + q 4 + code 1 -> r 4  (Jacobi)
BYTES_XFLUSH = 4 * NSBUF + 16   # per launch: NSBUF directions + fp64 x read / write
BYTES_ITER = BYTES_PHASE_A + BYTES_PHASE_B + BYTES_XFLUSH / NSBUF


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="splashy_b200", choices=["splashy_b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="x = y = per-GPU z extent")
    ap.add_argument("--pcg-iters", type=int, default=200)
    ap.add_argument("--precond", default="jacobi", choices=["none", "jacobi", "rbic0", "mg"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tolerance-run", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=128, help="edge of the CPU sample box")
    return ap.parse_args()


def ncu_traffic(kernel: str, n: int):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/r1_traffic.json, taken at 512^3); None for any other size."""
    p = ROOT / "profiles" / "r1_traffic.json"
    if n != 512 or not p.exists():
        return None
    try:
        t = json.loads(p.read_text())
        key = "k_phaseA_ws" if kernel == "k_phaseA" else kernel
        return float(t[key]["dram_bytes"])
    except Exception:
        return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle's C + OpenMP port on a bounded sample of the same workload
# ----------------------------------------------------------------------------------
def cpu_sample(n: int, pcg_iters: int, repeats: int = 1):
    from oracle import c_port
    from splashy_b200 import scenes
    c_port.use_all_cores()
    sc = scenes.vortex(n, n, n, dtype=np.float32)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = c_port.step(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"],
                          sc["rho"], fixed_iters=pcg_iters)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cells = n ** 3
    return {"value": cells / best, "unit": UNIT, "cores": c_port.threads(), "kind": "port",
            "sample": f"vortex {n}^3 cut of the workload, 1 step with {pcg_iters} Jacobi-PCG "
                      f"iterations, oracle/dense_ref.c (C + OpenMP, fp32 fields / fp64 dots), "
                      f"best of {repeats}; the F# reference itself cannot run (no .NET)",
            "seconds": best, "iters": out[4]["iters"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    from oracle import c_port
    from splashy_b200 import scenes
    c_port.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1 to the ranks
    sc = scenes.vortex(n, n, n, dtype=np.float32)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_port.step(sc["labels"], sc["U"], sc["V"], sc["W"], sc["dt"], sc["h"], sc["rho"],
                    fixed_iters=args.pcg_iters)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = n ** 3 / sec
    base = {"value": value, "unit": UNIT, "cores": c_port.threads(), "kind": "port",
            "sample": f"each step = vortex {n}^3 cut of the workload (bounded sample), "
                      f"{args.pcg_iters} Jacobi-PCG iterations, oracle/dense_ref.c"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"vortex_512 (512^3 per GPU), CPU arm on a {n}^3 sample",
                   "pcg_iters": args.pcg_iters, "precond": "jacobi"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from splashy_b200 import PinnedArray, Solver, nccl_unique_id, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    nid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nid = box[0]

    n = args.n
    nzg = n * world
    sc = scenes.vortex(n, n, n, z0=rank * n, nz_global=nzg, dtype=np.float32)
    dt = sc["dt"]
    h_, rho_ = sc["h"], sc["rho"]
    cells_local = n ** 3
    cells = cells_local * world

    s = Solver(n, n, n, dtype=np.float32, h=sc["h"], rho=sc["rho"], precond=args.precond,
               tol=1e-6, max_iters=20000, device=local, rank=rank, world=world,
               nz_global=nzg, z0=rank * n, nccl_id=nid)
    # pinned host buffers (the e2e leg copies from / to these every step)
    hin = {k: PinnedArray((n, n, n), np.float32) for k in "UVW"}
    hlab = PinnedArray((n, n, n), np.uint8)
    hout = {k: PinnedArray((n, n, n), np.float32) for k in "UVWP"}
    for k in "UVW":
        hin[k].array[...] = sc[k]
    hlab.array[...] = sc["labels"]
    del sc
    ptr = lambda a: C.c_void_p(a.array.ctypes.data)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
    s.snapshot()
    s.set_fixed_iters(args.pcg_iters)

    # ---- device-resident leg ("value") ---------------------------------------------
    launches = 0
    for _ in range(args.warmup):
        s.restore()
        s.step(dt)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    solve_ms = 0.0
    stage = {"ms_advect": 0.0, "ms_rhs": 0.0, "ms_solve": 0.0, "ms_grad": 0.0}
    for _ in range(args.steps):
        s.restore()                      # device-to-device reset, outside the timers
        s.timer_start()
        info = s.step(dt)
        dev_ms += s.timer_stop()
        launches += info["kernel_launches"]
        solve_ms += info["ms_solve"]
        for k in stage:
            stage[k] += info[k] / args.steps
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    solve_ms = max_over_ranks(solve_ms)
    ms_per_step = dev_ms / args.steps
    value = cells / (ms_per_step * 1e-3)
    iters_per_s = args.pcg_iters / (solve_ms / args.steps * 1e-3)
    last_info = info

    # ---- end-to-end leg: host buffers in, host buffers out -------------------------------
    h2d = n ** 3 * (3 * 4 + 1)
    d2h = n ** 3 * 4 * 4
    for _ in range(1):
        s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        s.step(dt)
        s.download_ptrs(ptr(hout["U"]), ptr(hout["V"]), ptr(hout["W"]), ptr(hout["P"]))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        s.step(dt)
        s.download_ptrs(ptr(hout["U"]), ptr(hout["V"]), ptr(hout["W"]), ptr(hout["P"]))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    e2e_value = cells / e2e_s
    checksum = float(hout["P"].array[n // 2, n // 2, ::64].astype(np.float64).sum())

    # ---- per-kernel roofline: CUDA events around every PCG launch -------------------------
    s.restore()
    ms_a, ms_b, ms_x = s.profile_pcg(dt, 32)
    ms_a, ms_b, ms_x = max_over_ranks(ms_a), max_over_ranks(ms_b), max_over_ranks(ms_x)
    peak, peak_src = peaks()
    gbs_a = cells_local * BYTES_PHASE_A / (ms_a * 1e-3) / 1e9
    gbs_b = cells_local * BYTES_PHASE_B / (ms_b * 1e-3) / 1e9
    gbs_x = cells_local * BYTES_XFLUSH / (ms_x * 1e-3) / 1e9 if ms_x > 0 else 0.0
    dom = "k_phaseB" if ms_b >= ms_a else "k_phaseA"
    dom_gbs = gbs_b if dom == "k_phaseB" else gbs_a
    step_bytes = cells_local * (75 + args.pcg_iters * BYTES_ITER)
    step_gbs = step_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- one run-to-tolerance step (iteration count) ------------------------
    tol_run = None
    if not args.no_tolerance_run:
        s.set_fixed_iters(0)
        s.restore()
        try:
            ti = s.step(dt)
        except Exception as e:          # not converged within max_iters: still report
            ti = dict(s.last_info)
            ti["error"] = str(e)
        tol_run = {"tol": 1e-6, "precond": args.precond, "iters": ti["iters"], "converged": ti["converged"],
                   "ms_step": max_over_ranks(ti["ms_total"]), "r_inf": ti["r_inf"],
                   "b_inf": ti["b_inf"], "max_abs_div": ti["max_abs_div"]}
        tol_run["cell_updates_per_s"] = cells / (tol_run["ms_step"] * 1e-3)
        if "error" in ti:
            tol_run["error"] = ti["error"]

    # ---- the same step with the multigrid preconditioner, to tolerance (1 GPU) ---------------
    tol_mg = None
    if not args.no_tolerance_run and args.precond != "mg":
        s.close()
        s = None
        nid2 = None
        if world > 1:
            box = [nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            nid2 = box[0]
        smg = Solver(n, n, n, dtype=np.float32, h=h_, rho=rho_, precond="mg", tol=1e-6,
                     max_iters=2000, device=local, rank=rank, world=world, nz_global=nzg,
                     z0=rank * n, nccl_id=nid2)
        smg.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        smg.snapshot()
        best = None
        for _ in range(3):
            smg.restore()
            barrier()
            smg.timer_start()
            ti = smg.step(dt)
            ms = max_over_ranks(smg.timer_stop())
            if best is None or ms < best[0]:
                best = (ms, ti)
        ms, ti = best
        # the same without the refinement restart (recursive residual only)
        smg.set_refinement(0)
        smg.restore()
        barrier()
        smg.timer_start()
        t0 = smg.step(dt)
        ms0 = max_over_ranks(smg.timer_stop())
        tol_mg = {"tol": 1e-6, "precond": "mg", "iters": ti["iters"], "converged": ti["converged"],
                  "refinement_rounds": 1,
                  "without_refinement": {"iters": t0["iters"], "ms_step": ms0,
                                         "max_abs_div": t0["max_abs_div"]},
                  "ms_step": ms, "ms_solve": ti["ms_solve"], "r_inf": ti["r_inf"],
                  "b_inf": ti["b_inf"], "max_abs_div": ti["max_abs_div"],
                  "cell_updates_per_s": cells / (ms * 1e-3), "best_of": 3}
        smg.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(args.cpu_n, args.pcg_iters)

    launches_total = int(sum_over_ranks(float(launches)))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"vortex_512: {n}x{n}x{n} cells per GPU (global {n}x{n}x{nzg}), "
                                   f"z-slabs, CFL 2, fp32 fields / fp64 dots",
                       "step": "advect + rhs + PCG + gradient + check",
                       "pcg_iters": args.pcg_iters, "precond": args.precond,
                       "l2": "inputs (3 x 512 MiB fields per GPU) exceed the 126 MB L2; no flush",
                       "parallelism": f"z-slab x{world}, NCCL halo + allgather"},
            "pcg_iters_per_s": iters_per_s,
            "stage_ms": stage,
            "wall_ms_per_step": wall / args.steps * 1e3,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_s * 1e3,
                    "result_checksum": checksum},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak,
                         "unit": "GB/s", "frac": dom_gbs / peak,
                         "traffic": ncu_traffic(dom, n),
                         "algorithmic_bytes": cells_local * (BYTES_PHASE_B if dom == "k_phaseB"
                                                             else BYTES_PHASE_A),
                         "peak_source": peak_src,
                         "kernels": {"k_phaseA": {"ms": ms_a, "bytes_per_cell": BYTES_PHASE_A,
                                                  "GBps": gbs_a, "frac": gbs_a / peak},
                                     "k_phaseB": {"ms": ms_b, "bytes_per_cell": BYTES_PHASE_B,
                                                  "GBps": gbs_b, "frac": gbs_b / peak},
                                     "k_xflush": {"ms": ms_x, "bytes_per_cell": BYTES_XFLUSH,
                                                  "every_iters": NSBUF, "GBps": gbs_x,
                                                  "frac": gbs_x / peak}},
                         "whole_step": {"bytes_per_cell": 75 + args.pcg_iters * BYTES_ITER,
                                        "GBps": step_gbs, "frac": step_gbs / peak}},
            "clocks": clocks,
            "to_tolerance": tol_run,
            "to_tolerance_mg": tol_mg,
            "check": {"max_abs_div": last_info["max_abs_div"], "nonfinite": last_info["nonfinite"]},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if s is not None:
        s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
