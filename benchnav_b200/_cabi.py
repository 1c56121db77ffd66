"""ctypes binding of libbnvmppi.so -- the C ABI declared in ``include/bnv_mppi.h``.

This is the stub a maintainer of the reference would add to call the engine from Python (see
INTEGRATION.md).  Nothing here touches PyTorch: pointers are plain integers (``tensor.data_ptr()``)
and the stream is a ``cudaStream_t`` integer.  Loading fails loudly if the library is missing --
there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import re
import weakref
from typing import List

from . import build as _build

BNV_OK = 0
BNV_FLAG_RECORD_STATES = 0x1
BNV_FLAG_STOCHASTIC_SLIP = 0x2
BNV_RISK_CLOSED_FORM, BNV_RISK_MONTE_CARLO = 0, 1
ABI_VERSION = 3


class BnvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbnvmppi error {code}: {msg}")
        self.code = code


class MppiCfg(C.Structure):
    """Mirror of ``bnv_mppi_cfg`` (include/bnv_mppi.h)."""

    _fields_ = [
        ("num_samples", C.c_int32),
        ("horizon", C.c_int32),
        ("sigma", C.c_float * 2),
        ("lambda_", C.c_float),
        ("u_min", C.c_float * 2),
        ("u_max", C.c_float * 2),
        ("dt", C.c_float),
        ("seed", C.c_uint64),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("num_envs", C.c_int32),
    ]


class Grid(C.Structure):
    """Mirror of ``bnv_grid`` (include/bnv_mppi.h)."""

    _fields_ = [
        ("grid_size", C.c_int32),
        ("pitch", C.c_int32),
        ("resolution", C.c_float),
        ("x_min", C.c_float),
        ("x_max", C.c_float),
        ("y_min", C.c_float),
        ("y_max", C.c_float),
    ]


_VP = C.c_void_p
_FP = C.POINTER(C.c_float)
_SIGNATURES = {
    "bnv_abi_version": (C.c_int, []),
    "bnv_last_error": (C.c_char_p, []),
    "bnv_mppi_create": (C.c_int, [C.POINTER(_VP), C.POINTER(MppiCfg)]),
    "bnv_mppi_destroy": (None, [_VP]),
    "bnv_mppi_set_problem": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.POINTER(C.c_float), C.c_float, _VP]),
    "bnv_mppi_set_problem_ex": (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_float, _FP, C.c_float, _VP]),
    "bnv_mppi_forward": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_forward_ex": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_forward_state": (C.c_int, [_VP, C.POINTER(C.c_float), _VP, _VP, _VP, _VP]),
    "bnv_mppi_forward_host": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_forward_host_batch": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_forward_host_action": (C.c_int, [_VP, _VP, _VP, _VP]),
    "bnv_mppi_wait_states": (C.c_int, [_VP, _VP]),
    "bnv_mppi_forward_follow": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_mailbox_handle": (C.c_int, [_VP, C.POINTER(C.c_ubyte)]),
    "bnv_mppi_attach_peers": (C.c_int, [_VP, C.POINTER(C.c_ubyte)]),
    "bnv_mppi_partial": (_VP, [_VP]),
    "bnv_mppi_partial_len": (C.c_int32, [_VP]),
    "bnv_mppi_finalize": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_top_samples": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP]),
    "bnv_mppi_merge_top": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP]),
    "bnv_mppi_weights": (_VP, [_VP]),
    "bnv_mppi_costs": (_VP, [_VP]),
    "bnv_mppi_states": (_VP, [_VP]),
    "bnv_mppi_noise": (_VP, [_VP]),
    "bnv_mppi_u_prev": (_VP, [_VP]),
    "bnv_mppi_local_samples": (C.c_int32, [_VP]),
    "bnv_mppi_sample_offset": (C.c_int32, [_VP]),
    "bnv_mppi_reset": (C.c_int, [_VP, _VP]),
    "bnv_mppi_draw_noise": (C.c_int, [_VP, C.c_uint64, _VP]),
    "bnv_mppi_draw_xi": (C.c_int, [_VP, C.c_uint64, _VP, _VP, _VP]),
    "bnv_mppi_device_counter": (C.c_int, [_VP, C.c_int32, _VP]),
    "bnv_mppi_iteration_counter": (_VP, [_VP]),
    "bnv_closed_loop_step": (C.c_int, [C.POINTER(Grid), _VP, _VP, C.c_int64, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP,
                                       C.c_uint64, _VP, _FP, _FP, C.c_float, C.c_float, C.c_float, _VP, _VP, _VP, _VP,
                                       _VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_prelaunch": (C.c_int, [_VP, C.c_int32, C.c_uint32]),
    "bnv_mppi_set_keep_mean": (C.c_int, [_VP, C.c_int32]),
    "bnv_mppi_set_terminal_goal": (C.c_int, [_VP, _FP]),
    "bnv_mppi_set_goal_dev": (C.c_int, [_VP, _VP]),
    "bnv_mppi_argmin": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bnv_mppi_dwa_subgoal": (C.c_int, [_VP, _VP, C.c_int32, _VP, _VP, C.c_float, _VP, _VP]),
    "bnv_mppi_launch_count": (C.c_uint64, [_VP]),
    "bnv_mppi_check": (C.c_int, [_VP, _VP]),
    "bnv_mppi_launch_geometry": (C.c_int, [_VP, C.POINTER(C.c_int32)]),
    "bnv_mppi_kernel_timing": (C.c_int, [_VP, C.c_int32]),
    "bnv_mppi_kernel_time": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "bnv_debug_timestamps": (C.c_int, [_VP, C.POINTER(C.c_longlong)]),
    "bnv_debug_flush": (C.c_int, [_VP, C.c_uint64, C.c_uint32, C.c_uint32, _VP]),
    "bnv_debug_philox": (C.c_int, [_VP, _VP, C.c_int32, _VP]),
    "bnv_debug_sincos": (C.c_int, [_VP, _VP, _VP, C.c_int32, _VP]),
    "bnv_trav_lookup": (C.c_int, [C.POINTER(Grid), _VP, _VP, C.c_int64, C.c_int64, _VP, C.c_int64, C.c_int32, _VP,
                                  C.c_uint64, C.c_uint64, _VP, C.c_float, _VP, _VP, _VP]),
    "bnv_env_step": (C.c_int, [C.POINTER(Grid), _VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP, _VP, C.c_uint64,
                               C.c_uint64, _VP, _FP, _FP, C.c_float, C.c_float, _VP, _VP, _VP]),
    "bnv_risk_map": (C.c_int, [C.c_int32, C.c_float, C.c_int32, _VP, _VP, C.c_int64, _VP, C.c_int32, C.c_uint64, _VP,
                               _VP, _VP]),
    "bnv_dwa_actions": (C.c_int, [_VP, _FP, _FP, _FP, C.c_float, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP]),
}

_lib = None


def header_path() -> str:
    return os.path.join(_build.INCLUDE, "bnv_mppi.h")


def declared_symbols() -> List[str]:
    """Every function name declared in include/bnv_mppi.h (used by the export test)."""
    text = open(header_path()).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnv_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load (building first if stale and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.ensure_built()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m benchnav_b200.build` (needs nvcc)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.bnv_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.bnv_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def _destroy_handle(lib, raw: C.c_void_p) -> None:
    if raw.value:
        lib.bnv_mppi_destroy(raw)
        raw.value = None


class SolverHandle:
    """Owner of one ``bnv_mppi*`` (bnv_mppi_create / bnv_mppi_destroy).

    The solver object and every zero-copy view of an engine buffer hold a reference to THIS object, never to the
    solver, so there is no reference cycle through the C++ side of ``torch.as_tensor``: the handle is destroyed when
    the last of them goes away, or at once through ``close()`` (afterwards every ABI call fails with "null argument").
    ctypes passes the object wherever a ``bnv_mppi*`` is expected (``_as_parameter_``)."""

    def __init__(self, lib: C.CDLL, cfg: "MppiCfg") -> None:
        raw = C.c_void_p()
        check(lib.bnv_mppi_create(C.byref(raw), C.byref(cfg)))
        self._as_parameter_ = raw
        self._finalizer = weakref.finalize(self, _destroy_handle, lib, raw)
        self._finalizer.atexit = False  # at interpreter exit the CUDA context may already be gone

    @property
    def value(self):
        return self._as_parameter_.value

    def close(self) -> None:
        self._finalizer()


def make_grid(grid_size: int, pitch: int, resolution: float, x_limits, y_limits) -> Grid:
    return Grid(int(grid_size), int(pitch), float(resolution), float(x_limits[0]), float(x_limits[1]),
                float(y_limits[0]), float(y_limits[1]))


def check(code: int) -> None:
    if code != BNV_OK:
        raise BnvError(code, load().bnv_last_error().decode("utf-8", "replace"))
