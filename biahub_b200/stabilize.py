"""Apply path of ``biahub stabilize`` — mirror of reference ``biahub/stabilize.py:32-90``.

``apply_stabilization_transform(zyx_data, list_of_shifts, input_time_index, output_shape=None)``
picks the 4x4 matrix of the timepoint and pull-warps every channel with linear interpolation
and the ITK boundary rule (the reference calls ANTs ``apply_to_image`` with its default
``interpolation="linear"``).  ``iohub``'s ``process_single_position`` injects
``input_time_index`` by name (reference stabilize.py:288-300 vs :32-37), so the parameter name
is part of the contract.
"""

from __future__ import annotations

import numpy as np

from ._device import is_torch_tensor
from .register import affine_warp

__all__ = ["apply_stabilization_transform"]


def apply_stabilization_transform(
    zyx_data: np.ndarray,
    list_of_shifts: list[np.ndarray],
    input_time_index: int,
    output_shape: tuple[int, int, int] = None,
):
    """Stabilise a (Z, Y, X) or (C, Z, Y, X) volume with ``list_of_shifts[input_time_index]``;
    returns float32 of ``output_shape`` (default: the input's ZYX shape)."""
    if output_shape is None:
        output_shape = tuple(zyx_data.shape[-3:])
    output_shape = tuple(int(v) for v in output_shape)
    matrix = np.asarray(list_of_shifts[input_time_index], dtype=np.float64)
    if matrix.shape != (4, 4):
        raise ValueError("each stabilization transform must be a 4x4 matrix")

    if zyx_data.ndim == 4:
        if is_torch_tensor(zyx_data):
            import torch

            return torch.stack([affine_warp(zyx_data[c], matrix, output_shape, order=1, boundary="itk")
                                for c in range(zyx_data.shape[0])])
        stabilized = np.zeros((zyx_data.shape[0],) + output_shape, dtype=np.float32)
        for c in range(zyx_data.shape[0]):
            stabilized[c] = affine_warp(zyx_data[c], matrix, output_shape, order=1, boundary="itk")
        return stabilized
    if zyx_data.ndim != 3:
        raise ValueError("zyx_data must be (Z, Y, X) or (C, Z, Y, X)")
    return affine_warp(zyx_data, matrix, output_shape, order=1, boundary="itk")
