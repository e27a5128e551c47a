"""splashy_b200 -- B200-native advect -> project hot path of cantsin/splashy.

The product is ``lib/libsplashy_cuda.so`` (hand-written sm_100a kernels behind
the C ABI of ``include/splashy_cuda.h``); this package is the ctypes harness
plus the synthetic scenes.  Importing it does not load CUDA; the first use of
:class:`splashy_b200.solver.Solver` does, and fails loudly without the library
or without a GPU (there is no CPU fallback).
"""
from .solver import Solver, SplashyError, PinnedArray, device_count, nccl_unique_id  # noqa: F401
from . import scenes  # noqa: F401

__all__ = ["Solver", "SplashyError", "PinnedArray", "device_count",
           "nccl_unique_id", "scenes"]
