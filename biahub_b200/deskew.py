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
    "_average_n_slices", "_average_n_slices_torch", "_get_averaged_shape", "_get_transform_matrix",
    "get_deskewed_data_shape", "fast_deskew_zyx", "deskew_zyx", "_fast_deskew_czyx", "_deskew_czyx",
    "deskew_scalars",
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


def _average_n_slices_torch(data, average_window_width: int):
    """Average a CUDA float32 tensor over groups of slices along axis 0, the last slice repeated
    to fill the last group (reference deskew.py:71-96).  Runs ``b2_average_slices``; the
    production path fuses the averaging into the deskew kernel and never calls this — the legacy
    ``deskew_zyx`` does (reference deskew.py:438-440)."""
    import torch

    w = int(average_window_width)
    if w == 1:
        return data
    if w < 1:
        raise ValueError("average_window_width must be >= 1")
    if not (is_torch_tensor(data) and data.is_cuda):
        raise RuntimeError("_average_n_slices_torch computes on the GPU only: pass a CUDA tensor "
                           "(no CPU fallback)")
    src = data.to(torch.float32).contiguous()
    Z = int(src.shape[0])
    plane = int(np.prod(src.shape[1:])) if src.ndim > 1 else 1
    out = torch.empty((-(-Z // w),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    if out.numel():
        with torch.cuda.device(src.device):
            _cabi.check(_cabi.lib().b2_average_slices(
                src.data_ptr(), Z, plane, w, out.data_ptr(),
                torch.cuda.current_stream().cuda_stream))
    return out


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
    """numpy (Z, Y, X) → numpy float32 through the pinned-buffer host pipeline: slab uploads,
    deskew kernels and downloads overlap (``b2h_deskew``); with ``keep_overhang`` and a non-zero
    ``overhang_fill`` the deskewed slabs stay on the device, the fill runs once on the resident
    volume and the slabs are downloaded afterwards (``b2h_deskew_fill``)."""
    do_fill, use_mean, value = _fill_args(keep_overhang, overhang_fill)
    dev = resolve_device(device)
    src, code = host_source(zyx)
    if src.ndim != 3:
        raise ValueError("raw data must have ndim == 3 (Z, Y, X)")
    s = deskew_scalars(src.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices)
    out = check_out(out, (s["Zavg"], s["Yo"], s["Xo"]))
    _cabi.check(_cabi.lib().b2h_deskew_fill(
        src.ctypes.data_as(ctypes.c_void_p), code, s["Zi"], s["Yi"], s["Xi"],
        out.ctypes.data_as(ctypes.c_void_p), s["Zavg"], s["Yo"], s["Xo"], s["Zo"], s["N"],
        s["px32"], s["pxct32"], s["off32"], (1 if use_mean else 2) if do_fill else 0, value, dev))
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
    """Legacy entry point (reference deskew.py:371-453), with the legacy order of operations:

    1. deskew every tilt row (``average_n_slices=1``) — the reference's MONAI ``Affine`` with
       this matrix is a flip/transpose plus a 1-D lerp along the scan axis with zero padding, the
       same sampling the production kernel does (MONAI's own fp32 rounding is unpinned: MONAI is
       not installable here);
    2. ``_average_n_slices_torch`` on the DESKEWED stack, the last deskewed slice repeated when
       ``Y % N != 0`` (reference :438-440) — NOT the production path's padding of the input row;
    3. ``overhang_fill="mean"`` with ``keep_overhang``: the numpy variant
       ``_fill_overhang_with_mean`` (reference :277-336, 448-451): zero mask dilated three times
       with scipy's default 3-D CROSS, filled with the mean of the un-masked voxels.

    ``device`` only selects the GPU (see _device.py); ``debug_plot_path`` is accepted and
    ignored (a matplotlib diagnostic of the reference)."""
    import torch

    raw = np.asarray(raw_data)
    if raw.ndim != 3:
        raise ValueError("raw_data must have ndim == 3 (Z, Y, X)")
    # shape check first so the reference's error surfaces before any GPU work
    get_deskewed_data_shape(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    if overhang_fill not in ("zero", "mean"):
        raise ValueError("overhang_fill must be 'zero' or 'mean'")
    N = int(average_n_slices)
    if N == 1 and not (keep_overhang and overhang_fill == "mean"):
        return _deskew_host(raw, device, ls_angle_deg, px_to_scan_ratio, keep_overhang, 1, 0)
    dev = resolve_device(device)
    src, _ = host_source(raw)
    with torch.cuda.device(dev):
        if src.dtype == np.uint16:
            t = torch.from_numpy(src.view(np.int16)).to(f"cuda:{dev}").view(torch.uint16)
        else:
            t = torch.from_numpy(src).to(f"cuda:{dev}")
        res = _average_n_slices_torch(
            fast_deskew_zyx(t, ls_angle_deg, px_to_scan_ratio, keep_overhang, 1, 0), N)
        if keep_overhang and overhang_fill == "mean":
            lib = _cabi.lib()
            shape = tuple(int(v) for v in res.shape)
            nbytes = lib.b2_overhang_fill_workspace(*shape)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=res.device)
            _cabi.check(lib.b2_overhang_fill_ex(
                res.data_ptr(), *shape, 1, 0.0, 3, 6, ws.data_ptr(), nbytes,
                torch.cuda.current_stream().cuda_stream))
        out = check_out(None, tuple(res.shape))
        torch.from_numpy(out).copy_(res)
    return out


# Adapt ZYX functions to CZYX — module-level for multiprocessing pickling (reference deskew.py:545-579)
def _deskew_czyx(data, **kwargs):
    return deskew_zyx(data[0], **kwargs)[None]


def _fast_deskew_czyx(data, device="cuda", num_splits=1, out=None, **kwargs):
    """CZYX wrapper used by ``biahub deskew`` (reference deskew.py:551-579): takes ``data[0]``,
    returns ``(1, Z', Y', X')`` float32.

    ``num_splits`` caps device memory as in the reference (deskew.py:563-573): the input is split
    along its X axis (``np.array_split``), every chunk is deskewed on its own — device buffers
    are sized by the chunk — and lands in its rows of the output's Y axis (output Y = reversed
    input X, rows are independent: bit-identical to the unsplit call unless an overhang fill is
    requested, which then sees one chunk at a time exactly as in the reference).  With the
    default ``num_splits=1`` the device holds one input volume plus a three-slab output ring.
    ``out`` (extension): optional float32 (Z', Y', X') destination, e.g. a pinned buffer."""
    zyx = np.asarray(data[0])   # index first: lazy (zarr/dask) inputs decode one channel only
    num_splits = int(num_splits)
    if num_splits < 1:
        raise ValueError("num_splits must be >= 1")
    if num_splits == 1 or zyx.ndim != 3:
        return _deskew_host(zyx, device, out=out, **kwargs)[None]
    # (an overhang fill is applied per chunk, as the reference does: mask dilation and mean see
    # only the chunk)
    s = deskew_scalars(zyx.shape, kwargs["ls_angle_deg"], kwargs["px_to_scan_ratio"],
                       kwargs["keep_overhang"], kwargs.get("average_n_slices", 1))
    out = check_out(out, (s["Zavg"], s["Yo"], s["Xo"]))
    bounds = np.cumsum([0] + [len(c) for c in np.array_split(np.arange(s["Xi"]), num_splits)])
    for x0, x1 in zip(bounds[:-1], bounds[1:]):
        if x1 == x0:
            continue
        part = _deskew_host(np.ascontiguousarray(zyx[:, :, x0:x1]), device, **kwargs)
        out[:, s["Xi"] - x1:s["Xi"] - x0, :] = part   # output row y <- input column Xi-1-y
    return out[None]
