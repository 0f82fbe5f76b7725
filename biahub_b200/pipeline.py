"""Chained deskew → register with the intermediate volume resident in HBM (SURVEY.md §8f next-2;
BASELINE.json configs[4]: "8-position plate, deskew then register").

The reference runs the two steps as separate CLI commands with a zarr round trip in between
(reference nextflow/mantis-v2.nf; biahub/deskew.py:739-748 then biahub/register.py:561-572).
Here the deskewed float32 volume never leaves the device: the deskew kernel writes it with a
16-byte-aligned row pitch so that the affine kernel can stage it with TMA.
"""

from __future__ import annotations

import numpy as np

from ._device import is_torch_tensor, resolve_device
from .deskew import fast_deskew_zyx
from .register import _interpolation_order, affine_warp

__all__ = ["deskew_then_register"]


def deskew_then_register(raw, matrix, output_shape_zyx, *, ls_angle_deg, px_to_scan_ratio,
                         keep_overhang, average_n_slices=1, interpolation="linear",
                         crop_output_slicing=None, device=None, out=None):
    """``apply_affine_transform(_fast_deskew_czyx(raw), matrix, output_shape_zyx)`` without the
    host round trip.  ``raw``: (Z, Y, X) numpy array (result: numpy float32, optionally into the
    pinned/pageable ``out``) or CUDA tensor (result: CUDA tensor)."""
    import torch

    order = _interpolation_order(interpolation)
    on_device = is_torch_tensor(raw)
    if on_device:
        src = raw
    else:
        arr = np.ascontiguousarray(raw)
        dev = torch.device("cuda", resolve_device(device))
        if arr.dtype == np.uint16:
            src = torch.from_numpy(arr.view(np.int16)).to(dev, non_blocking=True).view(torch.uint16)
        else:
            src = torch.from_numpy(arr.astype(np.float32, copy=False)).to(dev, non_blocking=True)
    mid = fast_deskew_zyx(src, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices,
                          row_align=4)
    res = affine_warp(mid, matrix, output_shape_zyx, order=order, boundary="itk",
                      crop_output_slicing=crop_output_slicing)
    if on_device:
        return res
    if out is None:
        return res.cpu().numpy()
    torch.from_numpy(out).copy_(res)
    return out
