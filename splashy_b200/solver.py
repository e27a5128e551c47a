"""Python harness over the C ABI (``include/splashy_cuda.h``).

Thin, allocation-free wrapper used by the tests and ``bench.py``; numpy arrays
are passed as plain host pointers exactly as the F# shim of INTEGRATION.md
passes pinned ``float32[]`` / ``byte[]``.  Arrays are ``[k, j, i]`` (x fastest).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _native as nat

PRECOND = {"none": nat.SPL_PC_NONE, "jacobi": nat.SPL_PC_JACOBI,
           "rbic0": nat.SPL_PC_RBIC0, "mg": nat.SPL_PC_MG}


class SplashyError(RuntimeError):
    """Raised for every non-zero return code; ``str(e)`` carries the text of
    ``spl_last_error`` which mirrors the reference's ``failwith`` messages
    (Convection.fs:65, Pressure.fs:138-149, Vector.fs:17)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{nat.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


def device_count() -> int:
    return int(nat.load().spl_device_count())


def nccl_unique_id() -> bytes:
    """128 bytes to hand to every rank's ``Solver(nccl_id=...)``."""
    lib = nat.load()
    buf = C.create_string_buffer(128)
    rc = lib.spl_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise SplashyError(rc, lib.spl_last_error(None).decode())
    return buf.raw


class PinnedArray:
    """numpy view over page-locked host memory from ``spl_host_alloc``."""

    def __init__(self, shape, dtype):
        self._lib = nat.load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = self._lib.spl_host_alloc(nbytes)
        if not self._ptr:
            raise SplashyError(nat.SPL_E_CUDA, "spl_host_alloc failed")
        buf = (C.c_char * nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.spl_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Solver:
    """One context = one dense box (or one z-slab of it) on one GPU."""

    def __init__(self, nx: int, ny: int, nz: int, *, dtype=np.float32,
                 h: float = 10.0, rho: float = 999.97, precond: str = "jacobi",
                 tol: float = 1e-6, max_iters: int = 20000, tau: float = 0.0,
                 quirks: int = 0, origin: Sequence[int] = (0, 0, 0),
                 device: int = 0, rank: int = 0, world: int = 1,
                 nz_global: int = 0, z0: int = 0,
                 nccl_id: Optional[bytes] = None,
                 gravity: Sequence[float] = (0.0, -9.81, 0.0),
                 viscosity: float = 0.000001):
        self._lib = nat.load()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        self.shape = (nz, ny, nx)
        d = nat.SplDesc()
        self._lib.spl_desc_default(C.byref(d))
        d.nx, d.ny, d.nz = nx, ny, nz
        d.h, d.rho = h, rho
        d.dtype = nat.SPL_F64 if self.dtype == np.float64 else nat.SPL_F32
        d.precond = PRECOND[precond]
        d.tol, d.max_iters, d.tau = tol, max_iters, tau
        d.quirks = quirks
        d.origin[:] = list(origin)
        d.device = device
        d.rank, d.world, d.nz_global, d.z0 = rank, world, nz_global, z0
        d.gravity[:] = list(gravity)
        d.viscosity = viscosity
        self._id_buf = None
        if nccl_id is not None:
            self._id_buf = C.create_string_buffer(nccl_id, 128)
            d.nccl_id = C.cast(self._id_buf, C.c_void_p)
        self._ctx = C.c_void_p()
        rc = self._lib.spl_create(C.byref(self._ctx), C.byref(d))
        if rc != 0:
            raise SplashyError(rc, self._lib.spl_last_error(None).decode())
        self.desc = d

    # -- plumbing --------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise SplashyError(rc, self._lib.spl_last_error(self._ctx).decode())

    def _in(self, a, dtype) -> Optional[C.c_void_p]:
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.shape != self.shape:
            raise ValueError(f"expected shape {self.shape}, got {a.shape}")
        self._keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def close(self):
        if self._ctx:
            self._lib.spl_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state -------------------------------------------------------------------
    def upload(self, labels, U=None, V=None, W=None, P=None):
        self._keep = []
        self._check(self._lib.spl_upload(
            self._ctx, self._in(labels, np.uint8), self._in(U, self.dtype),
            self._in(V, self.dtype), self._in(W, self.dtype),
            self._in(P, self.dtype)))
        self._keep = []

    def download(self, div: bool = False, out: Optional[Dict] = None) -> Dict:
        """Returns {"U","V","W","P"[,"div"]}; pass ``out`` to reuse buffers."""
        res = out if out is not None else {}
        names = ["U", "V", "W", "P"] + (["div"] if div else [])
        for n in names:
            if n not in res:
                res[n] = np.empty(self.shape, dtype=self.dtype)
        ptr = lambda n: res[n].ctypes.data_as(C.c_void_p) if n in res else None
        self._check(self._lib.spl_download(
            self._ctx, ptr("U"), ptr("V"), ptr("W"), ptr("P"),
            ptr("div") if div else None))
        return res

    def set_scalars(self, fields: Sequence[np.ndarray]) -> None:
        """Cell-centred scalar fields advected with the velocity by every later advect / step."""
        arrs = [np.ascontiguousarray(f, dtype=self.dtype) for f in fields]
        for a in arrs:
            if a.shape != self.shape:
                raise ValueError(f"expected shape {self.shape}, got {a.shape}")
        ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        self._nscal = len(arrs)
        self._check(self._lib.spl_set_scalars(self._ctx, len(arrs), ptrs))

    def get_scalars(self):
        outs = [np.empty(self.shape, dtype=self.dtype) for _ in range(self._nscal)]
        ptrs = (C.c_void_p * max(len(outs), 1))(*[a.ctypes.data for a in outs])
        self._check(self._lib.spl_get_scalars(self._ctx, len(outs), ptrs))
        return outs

    # -- the hot path --------------------------------------------------------------
    def _call(self, fn, *args) -> Dict:
        info = nat.SplInfo()
        rc = fn(self._ctx, *args, C.byref(info))
        self.last_info = info.as_dict()
        self._check(rc)
        return self.last_info

    def advect(self, dt: float) -> Dict:
        return self._call(self._lib.spl_advect, dt)

    def add_forces(self, dt: float):
        self._check(self._lib.spl_add_forces(self._ctx, dt))

    def add_viscosity(self, dt: float):
        self._check(self._lib.spl_add_viscosity(self._ctx, dt))

    def cleanup(self):
        self._check(self._lib.spl_cleanup(self._ctx))

    def advance(self, dt: float, sanity: bool = True) -> Dict:
        return self._call(self._lib.spl_advance, dt, int(sanity))

    def divergence(self, dt: float) -> Dict:
        return self._call(self._lib.spl_divergence, dt)

    def project(self, dt: float) -> Dict:
        return self._call(self._lib.spl_project, dt)

    def check(self, div_tol: float = 1e-4) -> Dict:
        return self._call(self._lib.spl_check, div_tol)

    def step(self, dt: float) -> Dict:
        return self._call(self._lib.spl_step, dt)

    # -- bench helpers ----------------------------------------------------------------
    def snapshot(self):
        self._check(self._lib.spl_snapshot(self._ctx))

    def restore(self):
        self._check(self._lib.spl_restore(self._ctx))

    def set_fixed_iters(self, iters: int):
        self._check(self._lib.spl_set_fixed_iters(self._ctx, iters))

    def timer_start(self):
        self._check(self._lib.spl_timer_start(self._ctx))

    def set_refinement(self, rounds: int) -> None:
        """Rounds of iterative refinement (true fp64 residual + PCG restart) per solve."""
        self._check(self._lib.spl_set_refinement(self._ctx, rounds))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.spl_timer_stop(self._ctx, C.byref(ms)))
        return float(ms.value)

    def profile_pcg(self, dt: float, iters: int = 32):
        """(ms per phase-A launch, per phase-B launch, per x-flush launch),
        measured with CUDA events on the library's stream."""
        a, b, x = C.c_float(), C.c_float(), C.c_float()
        self._check(self._lib.spl_profile_pcg(self._ctx, dt, iters, C.byref(a),
                                              C.byref(b), C.byref(x)))
        return float(a.value), float(b.value), float(x.value)

    # -- marker stage (Simulator.fs:19-52,60-79) ---------------------------
    def set_world(self, lo: Sequence[int], hi: Sequence[int]) -> None:
        """World.contains as inclusive box-index bounds (i, j, k)."""
        a = (C.c_int * 3)(*lo)
        b = (C.c_int * 3)(*hi)
        self._check(self._lib.spl_set_world(self._ctx, a, b))

    def set_markers(self, xyz: np.ndarray) -> None:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self._nmark = len(xyz)
        self._check(self._lib.spl_set_markers(self._ctx, len(xyz), C.c_void_p(xyz.ctypes.data)))

    def get_markers(self) -> np.ndarray:
        out = np.empty((self._nmark, 3), dtype=np.float64)
        self._check(self._lib.spl_get_markers(self._ctx, C.c_void_p(out.ctypes.data)))
        return out

    def move_markers(self, dt: float) -> int:
        moved = C.c_int64()
        self._check(self._lib.spl_move_markers(self._ctx, dt, C.byref(moved)))
        return int(moved.value)

    def relabel(self, sanity: bool = True) -> None:
        self._check(self._lib.spl_relabel(self._ctx, 1 if sanity else 0))

    def download_media(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.uint8)
        self._check(self._lib.spl_download_media(self._ctx, C.c_void_p(out.ctypes.data)))
        return out

    def apply_preconditioner(self, r: np.ndarray):
        """z = M^-1 r (SPL_PC_RBIC0 / SPL_PC_MG) for a host array of the box; returns
        (z, device milliseconds of the application)."""
        r = np.ascontiguousarray(r, dtype=self.dtype).reshape(self.shape)
        z = np.empty_like(r)
        ms = C.c_float()
        self._check(self._lib.spl_apply_preconditioner(
            self._ctx, C.c_void_p(r.ctypes.data), C.c_void_p(z.ctypes.data), C.byref(ms)))
        return z, float(ms.value)

    def upload_ptrs(self, labels_ptr, u_ptr, v_ptr, w_ptr, p_ptr=None):
        """Raw host pointers (pinned buffers) -- no numpy conversion."""
        self._check(self._lib.spl_upload(self._ctx, labels_ptr, u_ptr, v_ptr,
                                         w_ptr, p_ptr))

    def download_ptrs(self, u_ptr, v_ptr, w_ptr, p_ptr, div_ptr=None):
        self._check(self._lib.spl_download(self._ctx, u_ptr, v_ptr, w_ptr,
                                           p_ptr, div_ptr))

    # -- asynchronous transfers (pinned host buffers): spl_upload_async / spl_download_async --
    def upload_async_ptrs(self, labels_ptr, u_ptr, v_ptr, w_ptr, p_ptr=None):
        """Queue the host -> device copies and return; the next call that needs the state
        waits for them on the device.  The buffers must stay untouched until then."""
        self._check(self._lib.spl_upload_async(self._ctx, labels_ptr, u_ptr, v_ptr,
                                               w_ptr, p_ptr))

    def download_async_ptrs(self, u_ptr, v_ptr, w_ptr, p_ptr):
        """Queue the device -> host copies; `sync()` waits for them."""
        self._check(self._lib.spl_download_async(self._ctx, u_ptr, v_ptr, w_ptr, p_ptr))

    def sync(self):
        self._check(self._lib.spl_sync(self._ctx))
