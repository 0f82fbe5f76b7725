"""Oblique-plane deskew — host-side mirror of the reference's array-compute functions.

Same names, arguments, output-shape logic and error behaviour as reference
``biahub/deskew.py:43-579``; the arithmetic runs in libbiahub_b200.so (one fused sm_100a
kernel: axis flip/transpose + fp32 scan-axis lerp + N-slice average) instead of
``permute().flip().contiguous()`` + ``F.grid_sample`` + ``mean`` / MONAI ``Affine``.

Drop-in points (what ``iohub.ngff.utils.process_single_position`` is handed in
reference biahub/deskew.py:739-748): ``_fast_deskew_czyx`` (production) and ``_deskew_czyx``.
All functions are module-level so they pickle by reference into spawn-ed workers.
"""

from __future__ import annotations

import ctypes
import math
from pathlib import Path
from typing import Literal

import numpy as np

from . import _cabi
from ._device import (check_out, device_source, host_source, is_torch_tensor, padded_empty,
                      resolve_device, row_pitch)

__all__ = [
    "_average_n_slices", "_get_averaged_shape", "_get_transform_matrix", "get_deskewed_data_shape",
    "fast_deskew_zyx", "deskew_zyx", "_fast_deskew_czyx", "_deskew_czyx", "deskew_scalars",
]


# ------------------------------------------------------------------------------------------
# host-side shape / parameter logic (float64 python scalars, exactly as the reference)
# ------------------------------------------------------------------------------------------
def _get_averaged_shape(deskewed_data_shape: tuple, average_window_width: int) -> tuple:
    """Shape after averaging every `average_window_width` slices (reference deskew.py:157-177)."""
    first = int(np.ceil(deskewed_data_shape[0] / average_window_width))
    return (first,) + tuple(deskewed_data_shape[1:])


def _get_transform_matrix(ls_angle_deg: float, px_to_scan_ratio: float):
    """Centred-coordinate deskew affine (reference deskew.py:180-210)."""
    ct = np.cos(ls_angle_deg * np.pi / 180)
    m = np.zeros((4, 4))
    m[0, 0] = -px_to_scan_ratio * ct
    m[0, 2] = px_to_scan_ratio
    m[1, 0] = -1
    m[2, 1] = -1
    m[3, 3] = 1
    return m


def get_deskewed_data_shape(
    raw_data_shape: tuple,
    ls_angle_deg: float,
    px_to_scan_ratio: float,
    keep_overhang: bool,
    average_n_slices: int = 1,
    pixel_size_um: float = 1,
):
    """Output (Z, Y, X) shape and voxel size of the deskewed volume (reference deskew.py:213-274).

    Raises ``ValueError("Dataset contains only overhang ...")`` when ``keep_overhang=False``
    leaves nothing (reference deskew.py:262-267).
    """
    theta = ls_angle_deg * np.pi / 180
    st = np.sin(theta)
    ct = np.cos(theta)
    Z, Y, X = raw_data_shape
    if keep_overhang:
        Xp = int(np.ceil((Z / px_to_scan_ratio) + (Y * ct)))
    else:
        Xp = int(np.ceil((Z / px_to_scan_ratio) - (Y * ct)))
        if Xp <= 0:
            raise ValueError(
                f"Dataset contains only overhang when keep_overhang=False. "
                f"Computed Xp={Xp} <= 0. Either set keep_overhang=True or use a dataset "
                f"with non-overhang content."
            )
    output_shape = (Y, X, Xp)
    voxel_size = (average_n_slices * st * pixel_size_um, pixel_size_um, pixel_size_um)
    return _get_averaged_shape(output_shape, average_n_slices), voxel_size


def deskew_scalars(raw_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices=1):
    """Everything the C ABI needs: output dims + the three fp32 scalars that define rounding
    (computed in float64 as reference deskew.py:136-138, then rounded once to fp32)."""
    Zi, Yi, Xi = (int(v) for v in raw_shape)
    (Zo, Yo, Xo), _ = get_deskewed_data_shape((Zi, Yi, Xi), ls_angle_deg, px_to_scan_ratio, keep_overhang)
    N = int(average_n_slices)
    if N < 1:
        raise ValueError("average_n_slices must be >= 1")
    if Zi < 2:
        raise ValueError("deskew needs at least 2 scan planes (the reference divides by Z_in - 1)")
    ct = np.cos(ls_angle_deg * np.pi / 180)
    px = px_to_scan_ratio
    off = px * ct * (Zo - 1) / 2 - px * (Xo - 1) / 2 + (Zi - 1) / 2
    return dict(
        Zi=Zi, Yi=Yi, Xi=Xi, Zo=int(Zo), Yo=int(Yo), Xo=int(Xo), N=N,
        Zavg=int(math.ceil(Zo / N)),
        px32=float(np.float32(px)), pxct32=float(np.float32(px * ct)), off32=float(np.float32(off)),
    )


def _average_n_slices(data, average_window_width=1):
    """Average a HOST array over groups of slices along axis 0, edge-padding the last group
    (reference deskew.py:43-68).  Host utility kept for API parity (pinned by the reference's
    golden vector, tests/test_cli/test_deskew_cli.py:11-30); the deskew kernels fuse the
    averaging and never call this."""
    data = np.asarray(data)
    w = int(average_window_width)
    rem = data.shape[0] % w
    if rem:
        data = np.concatenate([data, np.repeat(data[-1:], w - rem, axis=0)], axis=0)
    return data.reshape((data.shape[0] // w, w) + data.shape[1:]).mean(axis=1)


# ------------------------------------------------------------------------------------------
# device path
# ------------------------------------------------------------------------------------------
def _fill_args(keep_overhang, overhang_fill):
    """(do_fill, use_mean, value) following reference deskew.py:538-540."""
    if not keep_overhang:
        return False, 0, 0.0
    if isinstance(overhang_fill, str):
        if overhang_fill == "mean":
            return True, 1, 0.0
        if overhang_fill == "zero":
            return False, 0, 0.0
        raise ValueError(f"overhang_fill must be 'mean' or a number, got {overhang_fill!r}")
    if overhang_fill != 0:
        return True, 0, float(overhang_fill)
    return False, 0, 0.0


def fast_deskew_zyx(
    raw_data,
    ls_angle_deg: float,
    px_to_scan_ratio: float,
    keep_overhang: bool,
    average_n_slices: int = 1,
    overhang_fill: Literal["mean"] | float = 0,
    *,
    row_align: int = 1,
    _path: int = _cabi.PATH_AUTO,
):
    """Deskew a (Z_scan, Y_tilt, X_coverslip) CUDA tensor → float32 CUDA tensor
    (ceil(Y/N), X, X_out).  Signature and semantics of reference deskew.py:456-542; float32 and
    uint16 tensors are consumed as they are, other dtypes are cast to float32 first.
    ``row_align`` (extension): round the output row pitch up to that many elements and return a
    view — ``row_align=4`` makes the result a TMA-eligible source for a chained ``affine_warp``."""
    import torch

    if not is_torch_tensor(raw_data):
        raise TypeError("fast_deskew_zyx expects a torch.Tensor already on the CUDA device")
    if raw_data.ndim != 3:
        raise ValueError("raw_data must have ndim == 3 (Z, Y, X)")
    src, code = device_source(raw_data)
    s = deskew_scalars(tuple(src.shape), ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    do_fill, use_mean, value = _fill_args(keep_overhang, overhang_fill)
    lib = _cabi.lib()
    with torch.cuda.device(src.device):
        out = padded_empty((s["Zavg"], s["Yo"], s["Xo"]), src.device, max(1, int(row_align)))
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.b2_deskew_pitched(
            src.data_ptr(), code, s["Zi"], s["Yi"], s["Xi"], out.data_ptr(), row_pitch(out),
            s["Zavg"], s["Yo"], s["Xo"], s["Zo"], s["N"], s["px32"], s["pxct32"], s["off32"],
            int(_path), stream))
        if do_fill:
            if not out.is_contiguous():
                raise NotImplementedError("overhang_fill with a padded row pitch")
            nbytes = lib.b2_overhang_fill_workspace(s["Zavg"], s["Yo"], s["Xo"])
            ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
            _cabi.check(lib.b2_overhang_fill(
                out.data_ptr(), s["Zavg"], s["Yo"], s["Xo"], use_mean, value, 3, ws.data_ptr(),
                nbytes, stream))
    return out


def _deskew_host(zyx, device, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices=1,
                 overhang_fill=0, out=None):
    """numpy (Z, Y, X) → numpy float32 through the pinned-buffer host pipeline."""
    do_fill, _, _ = _fill_args(keep_overhang, overhang_fill)
    dev = resolve_device(device)
    if do_fill:
        # the fill needs the whole volume on the device between kernel and download
        import torch

        src, _ = host_source(zyx)
        if src.dtype == np.uint16:
            t = torch.from_numpy(src.view(np.int16)).to(f"cuda:{dev}").view(torch.uint16)
        else:
            t = torch.from_numpy(src).to(f"cuda:{dev}")
        res = fast_deskew_zyx(t, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices,
                              overhang_fill)
        out = check_out(out, tuple(res.shape))  # pooled pinned array when the caller gave none
        torch.from_numpy(out).copy_(res)
        return out
    src, code = host_source(zyx)
    if src.ndim != 3:
        raise ValueError("raw data must have ndim == 3 (Z, Y, X)")
    s = deskew_scalars(src.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    out = check_out(out, (s["Zavg"], s["Yo"], s["Xo"]))
    _cabi.check(_cabi.lib().b2h_deskew(
        src.ctypes.data_as(ctypes.c_void_p), code, s["Zi"], s["Yi"], s["Xi"],
        out.ctypes.data_as(ctypes.c_void_p), s["Zavg"], s["Yo"], s["Xo"], s["Zo"], s["N"],
        s["px32"], s["pxct32"], s["off32"], dev))
    return out


def deskew_zyx(
    raw_data: np.ndarray,
    ls_angle_deg: float,
    px_to_scan_ratio: float,
    keep_overhang: bool,
    device: str = "cpu",
    average_n_slices: int = 1,
    overhang_fill: Literal["zero", "mean"] = "zero",
    debug_plot_path: Path | None = None,
) -> np.ndarray:
    """Legacy entry point (reference deskew.py:371-453, MONAI ``Affine`` path).  Routed to the
    same fused kernel as the production path; ``device`` only selects the GPU (see _device.py).
    Known differences from the MONAI arithmetic are listed in DESIGN.md (last averaged slice when
    Y % N != 0; cube instead of cross dilation for ``overhang_fill="mean"``)."""
    raw = np.asarray(raw_data)
    if raw.ndim != 3:
        raise ValueError("raw_data must have ndim == 3 (Z, Y, X)")
    # shape check first so the reference's error surfaces before any GPU work
    get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    if overhang_fill not in ("zero", "mean"):
        raise ValueError("overhang_fill must be 'zero' or 'mean'")
    return _deskew_host(raw, device, ls_angle_deg, px_to_scan_ratio, keep_overhang,
                        average_n_slices, overhang_fill)


# Adapt ZYX functions to CZYX — module-level for multiprocessing pickling (reference deskew.py:545-579)
def _deskew_czyx(data, **kwargs):
    return deskew_zyx(data[0], **kwargs)[None]


def _fast_deskew_czyx(data, device="cuda", num_splits=1, out=None, **kwargs):
    """CZYX wrapper used by ``biahub deskew`` (reference deskew.py:551-579): takes ``data[0]``,
    returns ``(1, Z', Y', X')`` float32.  ``num_splits`` is accepted for compatibility: the
    kernel is tile-based and the host pipeline already streams the volume in slabs, so no
    host-side split/concatenate is needed (splitting along input X is exact).
    ``out`` (extension): optional float32 (Z', Y', X') destination, e.g. a pinned buffer."""
    zyx = np.asarray(data)[0]
    if int(num_splits) < 1:
        raise ValueError("num_splits must be >= 1")
    return _deskew_host(zyx, device, out=out, **kwargs)[None]
