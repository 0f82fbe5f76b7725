"""Flat-field correction on the B200 — the reference's ``biahub/flat_field.py`` compute functions
with the same signatures (the pipeline stage before deskew in the mantis workflow).

* ``flat_field_zyx(zyx_data, axis=0)``        reference biahub/flat_field.py:105-122  → float64
* ``flat_field_correction``                    deprecated alias, :125-149
* ``_flat_field_czyx(czyx_data, target_indices)``  the callable the CLI hands to
  ``process_single_position`` (:152-166, :299-310) → float32 CZYX

``pattern = median(zyx, axis)``, result ``zyx / pattern * pattern.mean()``.  Computed by
``b2_flatfield_u16`` / ``b2h_flatfield_u16`` (csrc/b2_flatfield.cu): exact radix-select medians
and correctly rounded float64 arithmetic, so the results are BIT-IDENTICAL to numpy's.  The GPU
path takes uint16 acquisitions (what the camera writes and what the reference's pipeline feeds
this step); there is no CPU fallback, other dtypes raise.
"""

from __future__ import annotations

import ctypes
import warnings

import numpy as np

from . import _cabi
from ._device import check_out, is_torch_tensor, resolve_device

__all__ = ["flat_field_zyx", "flat_field_correction", "_flat_field_czyx"]


def _require_u16(dtype):
    if np.dtype(dtype) != np.uint16:
        raise NotImplementedError(
            f"biahub_b200 flat-field runs on uint16 acquisitions (got {np.dtype(dtype)}); "
            "there is no CPU fallback"
        )


def _flatfield_tensor(t, out_dtype):
    """CUDA uint16 tensor (Z, Y, X) → CUDA tensor of ``out_dtype`` on the current stream."""
    import torch

    if not t.is_cuda:
        raise RuntimeError("biahub_b200 computes on the GPU only: pass a CUDA tensor or a numpy array")
    if t.dtype != torch.uint16:
        raise NotImplementedError(f"flat-field runs on uint16 tensors (got {t.dtype})")
    if t.ndim != 3:
        raise ValueError("expected a (Z, Y, X) tensor")
    t = t.contiguous()
    Z, Y, X = (int(v) for v in t.shape)
    lib = _cabi.lib()
    with torch.cuda.device(t.device):
        out = torch.empty((Z, Y, X), dtype=out_dtype, device=t.device)
        if out.numel() == 0:
            return out
        ws_bytes = int(lib.b2_flatfield_workspace(Y, X))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=t.device)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        code = _cabi.DTYPE_F32 if out_dtype == torch.float32 else _cabi.DTYPE_F64
        _cabi.check(lib.b2_flatfield_u16(t.data_ptr(), Z, Y, X, out.data_ptr(), code, ws_ptr, ws_bytes,
                                         torch.cuda.current_stream().cuda_stream))
        ws.record_stream(torch.cuda.current_stream())
    return out


def _flatfield_host(zyx, out, code, device):
    Z, Y, X = zyx.shape
    if zyx.size:
        _cabi.check(_cabi.lib().b2h_flatfield_u16(
            zyx.ctypes.data_as(ctypes.c_void_p), Z, Y, X, out.ctypes.data_as(ctypes.c_void_p), code,
            resolve_device(device)))
    return out


def flat_field_zyx(zyx_data, axis: int = 0, device=None):
    """Divide out the median pattern along ``axis`` (reference biahub/flat_field.py:105-122).

    numpy uint16 (Z, Y, X) → numpy float64, or CUDA uint16 tensor → CUDA float64 tensor."""
    if is_torch_tensor(zyx_data):
        import torch

        t = zyx_data if axis == 0 else torch.movedim(zyx_data, axis, 0)
        res = _flatfield_tensor(t, torch.float64)
        return res if axis == 0 else torch.movedim(res, 0, axis)
    arr = np.asarray(zyx_data)
    _require_u16(arr.dtype)
    if arr.ndim != 3:
        raise ValueError("expected a (Z, Y, X) array")
    moved = arr if axis == 0 else np.moveaxis(arr, axis, 0)
    src = np.ascontiguousarray(moved)
    out = check_out(None, src.shape, np.float64)
    _flatfield_host(src, out, _cabi.DTYPE_F64, device)
    return out if axis == 0 else np.moveaxis(out, 0, axis)


def flat_field_correction(zyx_data, axis: int = 0):
    """Deprecated alias of :func:`flat_field_zyx` (reference biahub/flat_field.py:125-149)."""
    warnings.warn("flat_field_correction is deprecated; use flat_field_zyx instead.",
                  DeprecationWarning, stacklevel=2)
    return flat_field_zyx(zyx_data, axis=axis)


def _flat_field_czyx(czyx_data: np.ndarray, target_indices, device=None) -> np.ndarray:
    """Flat-field the channels in ``target_indices`` of a CZYX volume, pass the others through;
    float32 result (reference biahub/flat_field.py:152-166)."""
    czyx = np.asarray(czyx_data)
    if czyx.ndim != 4:
        raise ValueError("expected a (C, Z, Y, X) array")
    target = set(int(i) for i in target_indices)
    if target:
        _require_u16(czyx.dtype)
    out = check_out(None, czyx.shape, np.float32)
    for c in range(czyx.shape[0]):
        if c in target:
            _flatfield_host(np.ascontiguousarray(czyx[c]), out[c], _cabi.DTYPE_F32, device)
        else:
            out[c] = czyx[c].astype(np.float32)
    return out
