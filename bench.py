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
BYTES_ADVECT = 25      # u,v,w 12 + label 1 -> u',v',w' 12                 
BYTES_RHS = 17         # u,v,w 12 + code 1 -> b 4
BYTES_GRAD = 33        # p 8 (fp64) + label 1 + u,v,w 12 -> u,v,w 12
BYTES_CHECK = 21       # u,v,w 12 + p 8 + code 1
BYTES_GRADCHECK = 34   # one pass (k_grad_check4): label 1 + code 1 + p 8 + u,v,w 12 -> u,v,w 12


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
    ap.add_argument("--cpu-n", type=int, default=256,
                    help="edge of the CPU sample box (256^3: 0.7 GB working set, out of cache)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the same slab per GPU; strong: ONE n^3 box over the GPUs (cfg4)")
    ap.add_argument("--geometry", default="cube", choices=["cube", "cfg5"],
                    help="per-GPU slab: cube = n^3; cfg5 = 2n x 2n x n/4 (1024 x 1024 x 128)")
    ap.add_argument("--no-parity", action="store_true", help="skip the N > 1 slab-vs-one-GPU check")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="N > 1: skip the short cfg4-strong / cfg5 runs")
    return ap.parse_args()


def ncu_traffic(kernel: str, n: int):
    """DRAM bytes per launch of `kernel` from the committed ncu captures (taken at 512^3);
    None for any other size."""
    if n != 512:
        return None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = ROOT / "profiles" / name
        try:
            return float(json.loads(p.read_text())[kernel]["dram_bytes"])
        except Exception:
            continue
    return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_numa_cpus(gpu_index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader",
                              "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=20).stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                       # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return (node, cpus & os.sched_getaffinity(0)) if cpus & os.sched_getaffinity(0) else None
    except Exception:
        return None


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
class Dist:
    """torch.distributed plumbing: rendezvous, barriers, max / sum over ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def nccl_id(self):
        from splashy_b200 import nccl_unique_id
        if self.world == 1:
            return None
        box = [nccl_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(D, s, dt, warmup, steps, sampler=None):
    """`steps` device-timed spl_step calls after `warmup` untimed ones; the velocity field is
    reset (device to device, outside the timers) before each.  Returns per-step averages, the
    device time being the max over ranks."""
    for _ in range(warmup):
        s.restore()
        s.step(dt)
    D.barrier()
    if sampler is not None:
        sampler.start()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    launches = 0
    stage = {"ms_advect": 0.0, "ms_rhs": 0.0, "ms_solve": 0.0, "ms_grad": 0.0}
    info = None
    for _ in range(steps):
        s.restore()
        s.timer_start()
        info = s.step(dt)
        dev_ms += s.timer_stop()
        launches += info["kernel_launches"]
        for k in stage:
            stage[k] += info[k] / steps
    D.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler is not None else None
    ms = D.max(dev_ms) / steps
    stage = {k: D.max(v) for k, v in stage.items()}
    stage["ms_check"] = max(0.0, ms - sum(stage.values()))
    return {"ms_per_step": ms, "stage_ms": stage, "launches": launches, "info": info,
            "wall_ms_per_step": wall / steps * 1e3, "clocks": clocks}


def mgpu_parity(D, iters=40):
    """Slab decomposition against the same box on ONE GPU (both through the C ABI): a
    128 x 64 x 48N vortex, `iters` fixed Jacobi-PCG iterations -- enough z-chunks per slab for
    the peer-memory iteration, and two x flushes.  The slab result must equal the one-GPU result
    to fp32 round-off (`bit_identical` says whether every bit agreed)."""
    import torch
    from splashy_b200 import Solver, scenes
    world, rank = D.world, D.rank
    nx, ny, nz = 128, 64, 48
    nzg = nz * world
    whole = scenes.vortex(nx, ny, nzg, dtype=np.float32)
    part = scenes.slab_of(whole, rank * nz, nz)
    with Solver(nx, ny, nz, dtype=np.float32, h=whole["h"], rho=whole["rho"], device=D.local,
                rank=rank, world=world, nz_global=nzg, z0=rank * nz, nccl_id=D.nccl_id()) as s:
        s.upload(part["labels"], part["U"], part["V"], part["W"])
        s.set_fixed_iters(iters)
        s.step(whole["dt"])
        got = s.download()
    mine = torch.from_numpy(np.stack([got[k] for k in "UVWP"])).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    D.dist.all_gather(parts, mine)
    res = None
    if rank == 0:
        slab = torch.cat(parts, dim=1).cpu().numpy()
        with Solver(nx, ny, nzg, dtype=np.float32, h=whole["h"], rho=whole["rho"],
                    device=D.local) as s1:
            s1.upload(whole["labels"], whole["U"], whole["V"], whole["W"])
            s1.set_fixed_iters(iters)
            s1.step(whole["dt"])
            one = s1.download()
        vs = max(float(np.abs(one[k]).max()) for k in "UVW")
        verr = max(float(np.abs(slab[i] - one[k]).max()) for i, k in enumerate("UVW")) / vs
        perr = float(np.abs(slab[3] - one["P"]).max() / max(1e-30, np.abs(one["P"]).max()))
        # the rank partials of the fp64 dot products are summed in another association than one
        # GPU's CTA partials, so alpha / beta agree to ~1e-16 and a face whose new value nearly
        # cancels can round the other way: equal to fp32 round-off, bit-identical in most runs
        res = {"ok": bool(verr <= 1e-6 and perr <= 1e-6), "vel_err": verr, "p_err": perr,
               "case": f"vortex {nx}x{ny}x{nzg} over {world} slabs vs one GPU, {iters} fixed "
                       f"Jacobi-PCG iterations, fp32", "bit_identical": bool(verr == 0.0 and perr == 0.0)}
    D.barrier()
    return res


def run_cuda(args):
    from splashy_b200 import PinnedArray, Solver, scenes

    D = Dist(args)
    world, rank, local = D.world, D.rank, D.local
    parity = mgpu_parity(D) if world > 1 and not args.no_parity else None

    # ---- workload geometry ----------------------------------------------------------------
    n = args.n
    if args.scaling == "strong":          # cfg4: ONE n^3 box cut into `world` z-slabs
        if n % world:
            raise SystemExit(f"--scaling strong needs n % gpus == 0 ({n} % {world})")
        nx, ny, nz = n, n, n // world
        label = f"vortex_{n} (cfg4): one {n}^3 box over {world} z-slab(s) of {nz} planes"
    elif args.geometry == "cfg5":         # cfg5: 1024 x 1024 x 128 per GPU -> 1024^3 on 8 GPUs
        nx, ny, nz = 2 * n, 2 * n, n // 4
        label = f"weak_1024 (cfg5): {nx}x{ny}x{nz} cells per GPU (global {nx}x{ny}x{nz * world})"
    else:                                 # cfg4 at N = 1; the same slab per GPU as N grows
        nx, ny, nz = n, n, n
        label = f"vortex_512: {nx}x{ny}x{nz} cells per GPU (global {nx}x{ny}x{nz * world})"
    nzg = nz * world
    sc = scenes.vortex(nx, ny, nz, z0=rank * nz, nz_global=nzg, dtype=np.float32)
    dt, h_, rho_ = sc["dt"], sc["h"], sc["rho"]
    cells_local = nx * ny * nz
    cells = cells_local * world
    shape = (nz, ny, nx)

    def make(precond, tol=1e-6, max_iters=20000):
        return Solver(nx, ny, nz, dtype=np.float32, h=h_, rho=rho_, precond=precond, tol=tol,
                      max_iters=max_iters, device=local, rank=rank, world=world, nz_global=nzg,
                      z0=rank * nz, nccl_id=D.nccl_id())

    s = make(args.precond)
    # pinned host buffers (the e2e leg copies from / to these every step), allocated and first
    # touched on the GPU's own NUMA node: with one rank per GPU the copies of the 8 ranks then
    # do not all cross the socket interconnect
    numa = gpu_numa_cpus(local)
    old_aff = os.sched_getaffinity(0)
    if numa:
        try:
            os.sched_setaffinity(0, numa[1])
        except Exception:
            numa = None
    hin = {k: PinnedArray(shape, np.float32) for k in "UVW"}
    hlab = PinnedArray(shape, np.uint8)
    hout = {k: PinnedArray(shape, np.float32) for k in "UVWP"}
    hout2 = {k: PinnedArray(shape, np.float32) for k in "UVWP"}    # second box in flight
    for k in "UVW":
        hin[k].array[...] = sc[k]
    hlab.array[...] = sc["labels"]
    for hb in (hout, hout2):
        for k in "UVWP":
            hb[k].array[...] = 0
    if numa:
        os.sched_setaffinity(0, old_aff)
    del sc
    ptr = lambda a: C.c_void_p(a.array.ctypes.data)

    s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
    s.snapshot()
    s.set_fixed_iters(args.pcg_iters)

    # ---- device-resident leg ("value") ---------------------------------------------
    main = timed_steps(D, s, dt, args.warmup, args.steps, ClockSampler(local) if rank == 0 else None)
    ms_per_step = main["ms_per_step"]
    stage = main["stage_ms"]
    value = cells / (ms_per_step * 1e-3)
    iters_per_s = args.pcg_iters / (stage["ms_solve"] * 1e-3)
    last_info = main["info"]

    # ---- end-to-end leg: host buffers in, host buffers out -------------------------------
    h2d = cells_local * (3 * 4 + 1)
    d2h = cells_local * 4 * 4
    s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
    s.step(dt)
    s.download_ptrs(ptr(hout["U"]), ptr(hout["V"]), ptr(hout["W"]), ptr(hout["P"]))
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        s.step(dt)
        s.download_ptrs(ptr(hout["U"]), ptr(hout["V"]), ptr(hout["W"]), ptr(hout["P"]))
    D.barrier()
    e2e_one_s = D.max(time.perf_counter() - t0) / args.steps
    checksum = float(hout["P"].array[nz // 2, ny // 2, ::64].astype(np.float64).sum())
    # The same K steps with TWO boxes in flight per GPU (a second context; every step still
    # uploads its inputs from and downloads its result to pinned host memory): the copies of one
    # box ride on its copy streams while the other box's step runs.  One host thread, so the
    # collectives of the slab contexts are issued in the same order on every rank.
    e2e_mode = "two boxes in flight per GPU (spl_upload_async / spl_step / spl_download_async)"
    e2e_s = e2e_one_s
    e2e_match = True
    e2e_step_dev_ms = None
    if os.environ.get("SPL_BENCH_E2E", "pipelined") != "pipelined":
        e2e_mode = "one box: upload, step, download in sequence"
    else:
        s2 = make(args.precond)
        s2.set_fixed_iters(args.pcg_iters)
        boxes = [(s, hout), (s2, hout2)]
        up = lambda b: b.upload_async_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))

        per_box = [[], []]

        def run(n):
            dev = 0.0
            per_box[0].clear(); per_box[1].clear()
            up(boxes[0][0])
            boxes[0][0].sync()          # pipeline fill: the first upload has the link to itself
            for i in range(n):
                b, out = boxes[i % 2]
                if i + 1 < n:
                    up(boxes[(i + 1) % 2][0])
                info_ = b.step(dt)
                dev += info_["ms_total"]
                per_box[i % 2].append({k: round(info_[k], 2) for k in
                                       ("ms_total", "ms_advect", "ms_rhs", "ms_solve", "ms_grad")})
                b.download_async_ptrs(ptr(out["U"]), ptr(out["V"]), ptr(out["W"]), ptr(out["P"]))
            for b, _ in boxes:
                b.sync()
            return dev / n

        run(2)
        D.barrier()
        t0 = time.perf_counter()
        e2e_step_dev_ms = run(args.steps)
        D.barrier()
        e2e_s = D.max(time.perf_counter() - t0) / args.steps
        last = boxes[(args.steps - 1) % 2][1]
        checksum2 = float(last["P"].array[nz // 2, ny // 2, ::64].astype(np.float64).sum())
        e2e_match = checksum2 == checksum      # the same box through either path
        s2.close()
        s2 = None
    e2e_value = cells / e2e_s

    # ---- per-kernel roofline: CUDA events around every PCG launch -------------------------
    s.restore()
    ms_a, ms_b, ms_x = s.profile_pcg(dt, 32)
    ms_a, ms_b, ms_x = D.max(ms_a), D.max(ms_b), D.max(ms_x)
    peak, peak_src = peaks()
    gbs = lambda bytes_per_cell, ms: cells_local * bytes_per_cell / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    gbs_a, gbs_b, gbs_x = gbs(BYTES_PHASE_A, ms_a), gbs(BYTES_PHASE_B, ms_b), gbs(BYTES_XFLUSH, ms_x)
    dom = "k_phaseB" if ms_b >= ms_a else "k_phaseA_tma"
    dom_gbs = gbs_b if dom == "k_phaseB" else gbs_a
    # gradient + validators: one marching pass unless SPL_NO_GC_FUSE / SPL_NO_QUAD ask for two
    gc_fused = (args.n % 4 == 0 and os.environ.get("SPL_NO_GC_FUSE", "0") in ("", "0")
                and os.environ.get("SPL_NO_QUAD", "0") in ("", "0"))
    bytes_gc = BYTES_GRADCHECK if gc_fused else BYTES_GRAD + BYTES_CHECK
    step_bpc = BYTES_ADVECT + BYTES_RHS + bytes_gc + args.pcg_iters * BYTES_ITER
    step_gbs = gbs(step_bpc, ms_per_step)
    ms_gc = stage["ms_grad"] + stage["ms_check"]
    nonpcg_ms = stage["ms_advect"] + stage["ms_rhs"] + ms_gc
    nonpcg_bpc = BYTES_ADVECT + BYTES_RHS + bytes_gc
    kernels = {
        "k_phaseA_tma": {"ms": ms_a, "bytes_per_cell": BYTES_PHASE_A, "GBps": gbs_a, "frac": gbs_a / peak},
        "k_phaseB": {"ms": ms_b, "bytes_per_cell": BYTES_PHASE_B, "GBps": gbs_b, "frac": gbs_b / peak},
        "k_xflush": {"ms": ms_x, "bytes_per_cell": BYTES_XFLUSH, "every_iters": NSBUF,
                     "GBps": gbs_x, "frac": gbs_x / peak},
        # stages of the step outside the solve, from spl_info's CUDA events (sustained clocks:
        # they run inside the long step)
        "advect (k_advect_tma + k_advect_rest)": {
            "ms": stage["ms_advect"], "bytes_per_cell": BYTES_ADVECT,
            "GBps": gbs(BYTES_ADVECT, stage["ms_advect"]),
            "frac": gbs(BYTES_ADVECT, stage["ms_advect"]) / peak,
            "bound": "shared-memory gather pipe (114 LDS.32 / cell), not HBM"},
        "rhs (k_div_rhs4)": {"ms": stage["ms_rhs"], "bytes_per_cell": BYTES_RHS,
                             "GBps": gbs(BYTES_RHS, stage["ms_rhs"]),
                             "frac": gbs(BYTES_RHS, stage["ms_rhs"]) / peak},
        **({"grad + check (k_grad_check4)": {
                "ms": ms_gc, "bytes_per_cell": BYTES_GRADCHECK, "GBps": gbs(BYTES_GRADCHECK, ms_gc),
                "frac": gbs(BYTES_GRADCHECK, ms_gc) / peak}} if gc_fused else {
            "grad (k_grad_sub4)": {"ms": stage["ms_grad"], "bytes_per_cell": BYTES_GRAD,
                                   "GBps": gbs(BYTES_GRAD, stage["ms_grad"]),
                                   "frac": gbs(BYTES_GRAD, stage["ms_grad"]) / peak},
            "check (k_check4)": {"ms": stage["ms_check"], "bytes_per_cell": BYTES_CHECK,
                                 "GBps": gbs(BYTES_CHECK, stage["ms_check"]),
                                 "frac": gbs(BYTES_CHECK, stage["ms_check"]) / peak}}),
        "non_pcg (advect + rhs + grad + check)": {
            "ms": nonpcg_ms, "bytes_per_cell": nonpcg_bpc,
            "GBps": gbs(nonpcg_bpc, nonpcg_ms), "frac": gbs(nonpcg_bpc, nonpcg_ms) / peak},
    }

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
                   "ms_step": D.max(ti["ms_total"]), "r_inf": ti["r_inf"],
                   "b_inf": ti["b_inf"], "max_abs_div": ti["max_abs_div"]}
        tol_run["cell_updates_per_s"] = cells / (tol_run["ms_step"] * 1e-3)
        if "error" in ti:
            tol_run["error"] = ti["error"]
    s.close()
    s = None

    # ---- the same step with the multigrid preconditioner, to tolerance ------------------------
    tol_mg = None
    if not args.no_tolerance_run and args.precond != "mg":
        smg = make("mg", max_iters=2000)
        smg.upload_ptrs(ptr(hlab), ptr(hin["U"]), ptr(hin["V"]), ptr(hin["W"]))
        smg.snapshot()
        best = None
        for _ in range(3):
            smg.restore()
            D.barrier()
            smg.timer_start()
            ti = smg.step(dt)
            ms = D.max(smg.timer_stop())
            if best is None or ms < best[0]:
                best = (ms, ti)
        ms, ti = best
        # the same without the refinement restart (recursive residual only)
        smg.set_refinement(0)
        smg.restore()
        D.barrier()
        smg.timer_start()
        t0 = smg.step(dt)
        ms0 = D.max(smg.timer_stop())
        tol_mg = {"tol": 1e-6, "precond": "mg", "iters": ti["iters"], "converged": ti["converged"],
                  "refinement_rounds": 1,
                  "without_refinement": {"iters": t0["iters"], "ms_step": ms0,
                                         "max_abs_div": t0["max_abs_div"]},
                  "ms_step": ms, "ms_solve": ti["ms_solve"], "r_inf": ti["r_inf"],
                  "b_inf": ti["b_inf"], "max_abs_div": ti["max_abs_div"],
                  "cell_updates_per_s": cells / (ms * 1e-3), "best_of": 3}
        smg.close()
    for a in list(hin.values()) + list(hout.values()) + [hlab]:
        a.free()

    # ---- N > 1: the other two named configurations, short runs, in the same record ---------------
    extra = {}
    if world > 1 and not args.no_extra_configs and args.scaling == "weak" and args.geometry == "cube":
        def short_run(gx, gy, gz_local, name):
            scx = scenes.vortex(gx, gy, gz_local, z0=rank * gz_local, nz_global=gz_local * world,
                                dtype=np.float32)
            sx = Solver(gx, gy, gz_local, dtype=np.float32, h=scx["h"], rho=scx["rho"],
                        precond=args.precond, device=local, rank=rank, world=world,
                        nz_global=gz_local * world, z0=rank * gz_local, nccl_id=D.nccl_id())
            sx.upload(scx["labels"], scx["U"], scx["V"], scx["W"])
            sx.snapshot()
            sx.set_fixed_iters(args.pcg_iters)
            r = timed_steps(D, sx, scx["dt"], 2, 2)
            sx.close()
            tot = gx * gy * gz_local * world
            return {"workload": name, "ms_per_step": r["ms_per_step"],
                    "value": tot / (r["ms_per_step"] * 1e-3), "unit": UNIT,
                    "pcg_iters_per_s": args.pcg_iters / (r["stage_ms"]["ms_solve"] * 1e-3),
                    "stage_ms": r["stage_ms"], "steps": 2, "warmup": 2}
        if n % world == 0 and n // world >= 16:
            extra["strong_cfg4"] = short_run(
                n, n, n // world, f"cfg4 strong scaling: one {n}^3 box over {world} z-slabs of "
                                  f"{n // world} planes; compare with the N = 1 line of this bench")
            extra["strong_cfg4"]["scaling"] = "strong"
        extra["weak_cfg5"] = short_run(
            2 * n, 2 * n, n // 4, f"cfg5 geometry: {2 * n}x{2 * n}x{n // 4} cells per GPU "
                                  f"(global {2 * n}x{2 * n}x{n // 4 * world}), "
                                  f"{4 * (2 * n) ** 2 // 2 ** 20} MiB halo planes")
        extra["weak_cfg5"]["scaling"] = "weak"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(args.cpu_n, args.pcg_iters)

    launches_total = int(D.sum(float(main["launches"])))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{label}, z-slabs, CFL 2, fp32 fields / fp64 dots",
                       "step": "advect + rhs + PCG + gradient + check",
                       "pcg_iters": args.pcg_iters, "precond": args.precond,
                       "l2": "inputs (3 x 512 MiB fields per GPU) exceed the 126 MB L2; no flush",
                       "parallelism": ("one GPU, one box" if world == 1 else
                                       f"z-slab x{world}; PCG iteration over peer memory (NVLink "
                                       f"stores of halo planes and CG scalars), NCCL for the rest")},
            "pcg_iters_per_s": iters_per_s,
            "stage_ms": stage,
            "wall_ms_per_step": main["wall_ms_per_step"],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_s * 1e3,
                    "result_checksum": checksum, "mode": e2e_mode,
                    "same_result_as_one_box": e2e_match,
                    "step_device_ms": e2e_step_dev_ms,
                    "last_steps_per_box": [pb[-1] if pb else None for pb in per_box]
                    if e2e_step_dev_ms else None,
                    "one_box": {"value": cells / e2e_one_s, "ms_per_step": e2e_one_s * 1e3,
                                "mode": "upload, step, download of ONE box in sequence"},
                    "pinned_numa_node": numa[0] if numa else None},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak,
                         "unit": "GB/s", "frac": dom_gbs / peak,
                         "traffic": ncu_traffic(dom, n if (nx, ny, nz) == (512, 512, 512) else 0),
                         "algorithmic_bytes": cells_local * (BYTES_PHASE_B if dom == "k_phaseB"
                                                             else BYTES_PHASE_A),
                         "peak_source": peak_src,
                         "byte_model": "DESIGN.md section 4 (this build's arrays: 35 B/cell/iteration; "
                                       "SURVEY 8d's 45 B model would read 1.29x higher)",
                         "kernels": kernels,
                         "whole_step": {"bytes_per_cell": step_bpc,
                                        "GBps": step_gbs, "frac": step_gbs / peak}},
            "clocks": main["clocks"],
            "to_tolerance": tol_run,
            "to_tolerance_mg": tol_mg,
            "check": {"max_abs_div": last_info["max_abs_div"], "nonfinite": last_info["nonfinite"]},
        }
        if parity is not None:
            line["mgpu_parity"] = parity
        if extra:
            line["other_configs"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    D.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
