"""ctypes binding of ``libsplashy_cuda.so`` (include/splashy_cuda.h).

This is the same stub a maintainer of the reference would write as F#
``[<DllImport>]`` declarations; Python is used here because no
.NET toolchain exists in the image.  The library is looked up in-tree
(``splashy_b200/lib``) and loading fails loudly: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_DIR = Path(__file__).resolve().parent / "lib"
LIB_PATH = Path(os.environ["SPL_LIB"]) if os.environ.get("SPL_LIB") else LIB_DIR / "libsplashy_cuda.so"

SPL_F32, SPL_F64 = 0, 1
SPL_PC_NONE, SPL_PC_JACOBI, SPL_PC_RBIC0, SPL_PC_MG = 0, 1, 2, 3
SPL_AIR, SPL_FLUID, SPL_SOLID, SPL_ABSENT = 0, 1, 2, 255
SPL_Q_REF_INTERP, SPL_Q_FORWARD_TRACE, SPL_Q_NEAREST_COPY = 1, 2, 4
SPL_Q_MARKER_SUM, SPL_Q_MOVE_CELLS, SPL_Q_LAZY_BUFFER = 8, 16, 32
SPL_Q_ALL = 63

SPL_OK = 0
SPL_E_BADARG, SPL_E_CUDA, SPL_E_TRACE, SPL_E_NONFINITE = -1, -2, -3, -4
SPL_E_DIVERGENCE, SPL_E_NOT_CONVERGED, SPL_E_STATE = -5, -6, -7

ERROR_NAMES = {
    SPL_E_BADARG: "SPL_E_BADARG", SPL_E_CUDA: "SPL_E_CUDA",
    SPL_E_TRACE: "SPL_E_TRACE", SPL_E_NONFINITE: "SPL_E_NONFINITE",
    SPL_E_DIVERGENCE: "SPL_E_DIVERGENCE",
    SPL_E_NOT_CONVERGED: "SPL_E_NOT_CONVERGED", SPL_E_STATE: "SPL_E_STATE",
}


class SplDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("h", C.c_double), ("rho", C.c_double),
        ("dtype", C.c_int), ("precond", C.c_int),
        ("tol", C.c_double), ("max_iters", C.c_int),
        ("tau", C.c_double),
        ("quirks", C.c_uint),
        ("origin", C.c_int * 3),
        ("device", C.c_int),
        ("rank", C.c_int), ("world", C.c_int),
        ("nz_global", C.c_int), ("z0", C.c_int),
        ("nccl_id", C.c_void_p),
        ("gravity", C.c_double * 3),
        ("viscosity", C.c_double),
    ]


class SplInfo(C.Structure):
    _fields_ = [
        ("iters", C.c_int), ("converged", C.c_int),
        ("r_inf", C.c_double), ("b_inf", C.c_double),
        ("max_abs_div", C.c_double),
        ("nonfinite", C.c_int64), ("trace_escapes", C.c_int64),
        ("ms_advect", C.c_float), ("ms_rhs", C.c_float),
        ("ms_solve", C.c_float), ("ms_grad", C.c_float),
        ("ms_total", C.c_float),
        ("kernel_launches", C.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/splashy_cuda.h declares: (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "spl_desc_default": (None, [C.POINTER(SplDesc)]),
    "spl_create": (C.c_int, [C.POINTER(_VP), C.POINTER(SplDesc)]),
    "spl_destroy": (None, [_VP]),
    "spl_last_error": (C.c_char_p, [_VP]),
    "spl_abi_version": (C.c_int, []),
    "spl_device_count": (C.c_int, []),
    "spl_nccl_unique_id": (C.c_int, [_VP]),
    "spl_upload": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "spl_download": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "spl_upload_async": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "spl_download_async": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "spl_sync": (C.c_int, [_VP]),
    "spl_set_scalars": (C.c_int, [_VP, C.c_int, C.POINTER(_VP)]),
    "spl_get_scalars": (C.c_int, [_VP, C.c_int, C.POINTER(_VP)]),
    "spl_advect": (C.c_int, [_VP, C.c_double, C.POINTER(SplInfo)]),
    "spl_add_forces": (C.c_int, [_VP, C.c_double]),
    "spl_add_viscosity": (C.c_int, [_VP, C.c_double]),
    "spl_cleanup": (C.c_int, [_VP]),
    "spl_advance": (C.c_int, [_VP, C.c_double, C.c_int, C.POINTER(SplInfo)]),
    "spl_divergence": (C.c_int, [_VP, C.c_double, C.POINTER(SplInfo)]),
    "spl_project": (C.c_int, [_VP, C.c_double, C.POINTER(SplInfo)]),
    "spl_check": (C.c_int, [_VP, C.c_double, C.POINTER(SplInfo)]),
    "spl_step": (C.c_int, [_VP, C.c_double, C.POINTER(SplInfo)]),
    "spl_host_alloc": (_VP, [C.c_size_t]),
    "spl_host_free": (None, [_VP]),
    "spl_snapshot": (C.c_int, [_VP]),
    "spl_restore": (C.c_int, [_VP]),
    "spl_set_fixed_iters": (C.c_int, [_VP, C.c_int]),
    "spl_set_refinement": (C.c_int, [_VP, C.c_int]),
    "spl_timer_start": (C.c_int, [_VP]),
    "spl_timer_stop": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "spl_profile_pcg": (C.c_int, [_VP, C.c_double, C.c_int, C.POINTER(C.c_float),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "spl_device_ptr": (_VP, [_VP, C.c_int, C.POINTER(C.c_int)]),
    "spl_apply_preconditioner": (C.c_int, [_VP, _VP, _VP, C.POINTER(C.c_float)]),
    "spl_set_world": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spl_set_markers": (C.c_int, [_VP, C.c_int64, _VP]),
    "spl_get_markers": (C.c_int, [_VP, _VP]),
    "spl_move_markers": (C.c_int, [_VP, C.c_double, C.POINTER(C.c_int64)]),
    "spl_relabel": (C.c_int, [_VP, C.c_int]),
    "spl_download_media": (C.c_int, [_VP, _VP]),
}

_lib = None


def load() -> C.CDLL:
    """Load the in-tree CUDA library; raise if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SPLASHY_CUDA_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import "
            f"__graft_entry__ as g; g.build()'` (splashy_b200 has no CPU "
            f"fallback)")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.spl_abi_version() != 1:
        raise RuntimeError("libsplashy_cuda ABI version mismatch")
    _lib = lib
    return lib
