"""ctypes binding of libbiahub_b200.so (C ABI: include/biahub_b200.h).

The library is the product: if it is missing or no sm_100 device is present, every compute
entry point raises — there is no CPU or PyTorch fallback anywhere in this package.
"""

from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

ABI_VERSION = 2

DTYPE_U16 = 0
DTYPE_F32 = 1
DTYPE_F64 = 2
BOUNDARY_CONSTANT = 0
BOUNDARY_ITK = 1
PATH_AUTO = 0
PATH_GATHER = 1
PATH_TMA = 2

ERR_INVALID = 1
ERR_UNSUPPORTED = 2
ERR_NO_DEVICE = 3

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BIAHUB_B200_LIB") or os.path.join(_PKG_DIR, "_lib", "libbiahub_b200.so")

EXPORTS = (
    "b2_abi_version", "b2_last_error", "b2_device_count", "b2_check_device", "b2_deskew",
    "b2_affine3d", "b2_deskew_pitched", "b2_affine3d_pitched", "b2_overhang_fill_workspace", "b2_overhang_fill", "b2h_deskew",
    "b2h_affine3d", "b2h_release", "b2_launch_count",
    "b2_flatfield_workspace", "b2_flatfield_u16", "b2h_flatfield_u16", "b2h_deskew_affine3d",
    "b2_debug_oob_count", "b2_debug_bounds_check_build", "b2_overhang_fill_ex", "b2_average_slices", "b2h_deskew_fill", "b2_spline3_workspace", "b2_affine3d_spline3", "b2h_affine3d_spline3",
)


class B2Error(RuntimeError):
    """A libbiahub_b200 call failed (message from b2_last_error())."""


class B2Unsupported(B2Error):
    """The explicitly requested kernel path is not eligible for this input."""


_lib = None
_lock = threading.Lock()

_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_vp = ctypes.c_void_p


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B2Error(
                f"{LIB_PATH} is not built. Run `python -m biahub_b200._build` (needs nvcc); "
                "biahub_b200 has no CPU fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        handle.b2_abi_version.restype = _int
        handle.b2_last_error.restype = ctypes.c_char_p
        handle.b2_device_count.restype = _int
        handle.b2_check_device.argtypes = [_int]
        handle.b2_deskew.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _int,
                                     _f32, _f32, _f32, _int, _vp]
        handle.b2_affine3d.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                       ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64),
                                       _int, _int, _int, _int, _vp]
        handle.b2_deskew_pitched.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64,
                                             _i64, _int, _f32, _f32, _f32, _int, _vp]
        handle.b2_affine3d_pitched.argtypes = [_vp, _int, _i64, _i64, _i64, _i64, _vp, _i64, _i64,
                                               _i64, _i64, ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(_i64), _int, _int, _int, _int, _vp]
        handle.b2_overhang_fill_workspace.argtypes = [_i64, _i64, _i64]
        handle.b2_overhang_fill_workspace.restype = ctypes.c_size_t
        handle.b2_overhang_fill.argtypes = [_vp, _i64, _i64, _i64, _int, _f32, _int, _vp,
                                            ctypes.c_size_t, _vp]
        handle.b2h_deskew.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _int,
                                      _f32, _f32, _f32, _int]
        handle.b2h_affine3d.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64),
                                        _int, _int, _int, _int]
        handle.b2_flatfield_workspace.argtypes = [_i64, _i64]
        handle.b2_flatfield_workspace.restype = ctypes.c_size_t
        handle.b2_flatfield_u16.argtypes = [_vp, _i64, _i64, _i64, _vp, _int, _vp, ctypes.c_size_t, _vp]
        handle.b2h_flatfield_u16.argtypes = [_vp, _i64, _i64, _i64, _vp, _int, _int]
        handle.b2h_deskew_affine3d.argtypes = [_vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int,
                                               _f32, _f32, _f32, _vp, _i64, _i64, _i64,
                                               ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64),
                                               _int, _int, _int, _int]
        handle.b2_overhang_fill_ex.argtypes = [_vp, _i64, _i64, _i64, _int, _f32, _int, _int, _vp,
                                               ctypes.c_size_t, _vp]
        handle.b2_average_slices.argtypes = [_vp, _i64, _i64, _int, _vp, _vp]
        handle.b2h_deskew_fill.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64,
                                           _int, _f32, _f32, _f32, _int, _f32, _int]
        handle.b2_spline3_workspace.argtypes = [_i64, _i64, _i64]
        handle.b2_spline3_workspace.restype = ctypes.c_size_t
        handle.b2_affine3d_spline3.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _i64,
                                               _i64, ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(_i64), _int, _vp, ctypes.c_size_t, _vp]
        handle.b2h_affine3d_spline3.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _i64,
                                                _i64, ctypes.POINTER(ctypes.c_double),
                                                ctypes.POINTER(_i64), _int, _int]
        handle.b2h_release.restype = _int
        handle.b2_launch_count.restype = ctypes.c_uint64
        handle.b2_debug_oob_count.restype = ctypes.c_uint64
        handle.b2_debug_bounds_check_build.restype = _int
        if handle.b2_abi_version() != ABI_VERSION:
            raise B2Error(f"ABI mismatch: library {handle.b2_abi_version()} != binding {ABI_VERSION}")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = lib().b2_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise B2Unsupported(msg)
    raise B2Error(f"[rc={rc}] {msg}")


def device_count() -> int:
    return int(lib().b2_device_count())


def require_device(dev: int = 0) -> None:
    check(lib().b2_check_device(int(dev)))


def launch_count() -> int:
    return int(lib().b2_launch_count())


def matrix12(matrix) -> "ctypes.Array":
    m = np.asarray(matrix, dtype=np.float64)
    if m.shape == (4, 4):
        m = m[:3, :]
    if m.shape != (3, 4):
        raise ValueError(f"expected a 4x4 (or 3x4) affine matrix, got shape {m.shape}")
    return (ctypes.c_double * 12)(*m.ravel().tolist())


def int64x3(values) -> "ctypes.Array":
    return (_i64 * 3)(*[int(v) for v in values])
