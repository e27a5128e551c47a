"""ctypes access to ``oracle/_build/liboracle.so`` (dense_ref.c).

TEST / BENCH INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        subprocess.run(["make", "-C", str(HERE)], check=True)
    lib = C.CDLL(str(LIB))
    lib.orc_threads.restype = C.c_int
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_codes.argtypes = [C.c_int] * 3 + [_u8p, _u8p]
    lib.orc_div_rhs.argtypes = [C.c_int] * 3 + [_u8p, _f32p, _f32p, _f32p, C.c_float, _f32p]
    lib.orc_pcg_jacobi.restype = C.c_int
    lib.orc_pcg_jacobi.argtypes = [C.c_int] * 3 + [_u8p, _f32p, _f32p, C.c_double, C.c_int,
                                                   C.c_int, C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double)]
    lib.orc_grad_sub.argtypes = [C.c_int] * 3 + [_u8p, _f32p, _f32p, _f32p, _f32p, C.c_float]
    lib.orc_advect.argtypes = [C.c_int] * 3 + [_u8p] + [_f32p] * 6 + [C.c_float]
    lib.orc_advect_scalar.argtypes = [C.c_int] * 3 + [_u8p] + [_f32p] * 5 + [C.c_float]
    lib.orc_max_div.restype = C.c_double
    lib.orc_max_div.argtypes = [C.c_int] * 3 + [_u8p, _f32p, _f32p, _f32p]
    lib.orc_step.restype = C.c_int
    lib.orc_step.argtypes = [C.c_int] * 3 + [_u8p] + [_f32p] * 4 + [C.c_double] * 4 + \
        [C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3
    _lib = lib
    return lib


def threads() -> int:
    return int(load().orc_threads())


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().orc_set_threads(n)
    return threads()


def advect(labels, U, V, W, dt, h):
    lib = load()
    nz, ny, nx = labels.shape
    out = [np.empty_like(U) for _ in range(3)]
    lib.orc_advect(nx, ny, nz, labels, U, V, W, *out, float(dt / h))
    return tuple(out)


def advect_scalar(labels, U, V, W, S, dt, h):
    lib = load()
    nz, ny, nx = labels.shape
    out = np.empty_like(S)
    lib.orc_advect_scalar(nx, ny, nz, labels, U, V, W, S, out, float(dt / h))
    return out


def step(labels, U, V, W, dt, h, rho, tol=1e-6, max_iters=20000, fixed_iters=0):
    """In-place advect -> project on copies; returns (U, V, W, P, info)."""
    lib = load()
    nz, ny, nx = labels.shape
    U, V, W = (np.array(a, dtype=np.float32, order="C", copy=True) for a in (U, V, W))
    P = np.zeros_like(U)
    r, b, d = C.c_double(), C.c_double(), C.c_double()
    it = lib.orc_step(nx, ny, nz, np.ascontiguousarray(labels), U, V, W, P, dt, h, rho, tol,
                      max_iters, fixed_iters, C.byref(r), C.byref(b), C.byref(d))
    return U, V, W, P, {"iters": it, "r_inf": r.value, "b_inf": b.value,
                        "max_abs_div": d.value}
