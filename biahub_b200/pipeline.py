"""Chained deskew → register with the intermediate volume resident in HBM (SURVEY.md §8f next-2;
BASELINE.json configs[4]: "8-position plate, deskew then register").

The reference runs the two steps as separate CLI commands with a zarr round trip in between
(reference nextflow/mantis-v2.nf; biahub/deskew.py:739-748 then biahub/register.py:561-572).
Here the deskewed float32 volume never leaves the device: the deskew kernel writes it with a
16-byte-aligned row pitch so that the affine kernel can stage it with TMA.
"""

from __future__ import annotations

import numpy as np

from ._device import check_out, is_torch_tensor, resolve_device
from .deskew import fast_deskew_zyx
from .register import _interpolation_order, affine_warp

__all__ = ["deskew_then_register", "flatfield_then_deskew"]


def _to_device(raw, device):
    """numpy (Z, Y, X) → CUDA tensor (uint16 stays uint16, everything else float32)."""
    import torch

    arr = np.ascontiguousarray(raw)
    dev = torch.device("cuda", resolve_device(device))
    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).to(dev, non_blocking=True).view(torch.uint16)
    return torch.from_numpy(arr.astype(np.float32, copy=False)).to(dev, non_blocking=True)


def _to_host(res, out):
    """CUDA float32 tensor → numpy: into ``out``, or into a pooled pinned array (one DMA)."""
    import torch

    out = check_out(out, tuple(res.shape))
    torch.from_numpy(out).copy_(res)
    return out


def deskew_then_register(raw, matrix, output_shape_zyx, *, ls_angle_deg, px_to_scan_ratio,
                         keep_overhang, average_n_slices=1, interpolation="linear",
                         crop_output_slicing=None, device=None, out=None):
    """``apply_affine_transform(_fast_deskew_czyx(raw), matrix, output_shape_zyx)`` without the
    host round trip.  ``raw``: (Z, Y, X) numpy array (result: numpy float32, optionally into the
    pinned/pageable ``out``) or CUDA tensor (result: CUDA tensor)."""
    import torch

    order = _interpolation_order(interpolation)
    if is_torch_tensor(raw):
        mid = fast_deskew_zyx(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices,
                              row_align=4)
        return affine_warp(mid, matrix, output_shape_zyx, order=order, boundary="itk",
                           crop_output_slicing=crop_output_slicing)
    # host arrays: ONE pipelined C-ABI call (b2h_deskew_affine3d) — upload of the next tilt-row
    # band, deskew, warp of the output planes whose source planes are complete and their
    # download all overlap
    import ctypes

    from . import _cabi
    from ._device import host_source
    from .deskew import deskew_scalars
    from .register import _crop_box

    src, code = host_source(raw)
    if src.ndim != 3:
        raise ValueError("raw data must have ndim == 3 (Z, Y, X)")
    s = deskew_scalars(src.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    starts, sizes = _crop_box(output_shape_zyx, crop_output_slicing)
    out = check_out(out, sizes)
    if out.size:
        _cabi.check(_cabi.lib().b2h_deskew_affine3d(
            src.ctypes.data_as(ctypes.c_void_p), code, s["Zi"], s["Yi"], s["Xi"], s["Zavg"], s["Yo"],
            s["Xo"], s["Zo"], s["N"], s["px32"], s["pxct32"], s["off32"],
            out.ctypes.data_as(ctypes.c_void_p), *sizes, _cabi.matrix12(matrix),
            _cabi.int64x3(starts), int(order), _cabi.BOUNDARY_ITK, 1, resolve_device(device)))
    return out


def flatfield_then_deskew(raw, *, ls_angle_deg, px_to_scan_ratio, keep_overhang,
                          average_n_slices=1, overhang_fill=0, device=None, out=None):
    """``_fast_deskew_czyx(_flat_field_czyx(raw))`` — the first two stages of the mantis workflow
    (reference nextflow/mantis-v2.nf:106-122: deskew reads flat-field's output zarr) — with the flat-fielded float32
    volume resident in HBM instead of a zarr round trip: one uint16 upload, one float32 download.
    Bit-identical to running the two steps separately.  ``raw``: (Z, Y, X) uint16 numpy array
    (result numpy float32) or CUDA tensor (result CUDA tensor)."""
    import torch

    from .flat_field import _flatfield_tensor

    on_device = is_torch_tensor(raw)
    src = raw if on_device else _to_device(raw, device)
    flat = _flatfield_tensor(src, torch.float32)
    res = fast_deskew_zyx(flat, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices,
                          overhang_fill)
    return res if on_device else _to_host(res, out)
