"""Affine apply path of ``biahub register`` — host-side mirror of reference
``biahub/register.py:32-281, 397-398``.

``apply_affine_transform`` keeps the reference's signature and matrix convention (4x4
homogeneous, ZYX, pull: ``source_index = M[:3,:3] @ out_index + M[:3,3]``, no centring —
reference register.py:148-168, pinned by reference tests/test_affine.py:43-59) and returns a
float32 array; the resampling runs in libbiahub_b200.so instead of ANTs/ITK or scipy.

``method="ants"`` (default) uses the ITK boundary rule; ``interpolation`` accepts the two ANTs
modes this path is specified for: ``"linear"`` (order 1) and ``"nearestneighbor"`` (order 0);
ANTs' other interpolators (gaussian, bspline, windowed sinc, label modes) raise
``NotImplementedError`` — there is no CPU fallback.
``method="scipy"`` reproduces the reference's literal call
``scipy.ndimage.affine_transform(zyx_data, matrix, output_shape_zyx)`` (register.py:272): cubic
spline (order 3), ``mode="constant"``; ``output_shape_zyx`` lands in scipy's ``offset`` slot and
is ignored for a homogeneous matrix, so the result has the INPUT's shape and dtype
(``spline_warp``, csrc/b2_spline.cu).  ``affine_warp`` exposes order/boundary explicitly (scipy
``mode="constant"`` rule = the oracle's primary mode).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._device import (check_out, device_source, host_source, is_torch_tensor, padded_empty,
                      resolve_device, row_pitch)

__all__ = [
    "get_3D_rescaling_matrix", "get_3D_rotation_matrix", "get_3D_fliplr_matrix",
    "convert_transform_to_ants", "convert_transform_to_numpy", "apply_affine_transform",
    "affine_warp", "spline_warp", "rescale_voxel_size", "ItkAffineParameters", "largest_interior_rectangle",
    "find_lir", "find_overlapping_volume",
]

_INTERPOLATION_ORDER = {"linear": 1, "nearestneighbor": 0}


# ------------------------------------------------------------------------------------------
# matrix builders (reference register.py:32-145) — YX centre = shape/2
# ------------------------------------------------------------------------------------------
def _yx_centres(start_shape_zyx, end_shape_zyx):
    cy, cx = np.array(start_shape_zyx)[-2:] / 2
    if end_shape_zyx is None:
        return cy, cx, cy, cx
    ey, ex = np.array(end_shape_zyx)[-2:] / 2
    return cy, cx, ey, ex


def get_3D_rescaling_matrix(start_shape_zyx, scaling_factor_zyx=(1, 1, 1), end_shape_zyx=None):
    cy, cx, ey, ex = _yx_centres(start_shape_zyx, end_shape_zyx)
    sz, sy, sx = scaling_factor_zyx[-3], scaling_factor_zyx[-2], scaling_factor_zyx[-1]
    m = np.eye(4)
    m[0, 0] = sz
    m[1, 1] = sy
    m[2, 2] = sx
    m[1, 3] = -cy * sy + ey
    m[2, 3] = -cx * sx + ex
    return m


def get_3D_rotation_matrix(start_shape_zyx, angle: float = 0.0, end_shape_zyx=None) -> np.ndarray:
    cy, cx, ey, ex = _yx_centres(start_shape_zyx, end_shape_zyx)
    theta = np.radians(angle)
    c, s = np.cos(theta), np.sin(theta)
    m = np.eye(4)
    m[1, 1], m[1, 2] = c, -s
    m[2, 1], m[2, 2] = s, c
    m[1, 3] = -cy * c + s * cx + ey
    m[2, 3] = -cy * s - cx * c + ex
    return m


def get_3D_fliplr_matrix(start_shape_zyx, end_shape_zyx=None) -> np.ndarray:
    cx_end = (start_shape_zyx[-1] if end_shape_zyx is None else end_shape_zyx[-1]) / 2
    m = np.eye(4)
    m[2, 2] = -1
    m[2, 3] = 2 * cx_end
    return m


def rescale_voxel_size(affine_matrix, input_scale):
    """Row norms of the linear part times the input scale (reference register.py:397-398)."""
    return np.linalg.norm(affine_matrix, axis=1) * input_scale


# ------------------------------------------------------------------------------------------
# ITK/ANTs parameter layout (reference register.py:148-199)
# ------------------------------------------------------------------------------------------
class ItkAffineParameters:
    """Stand-in for an ANTs ``AffineTransform``: 12 parameters = row-major 3x3 then translation,
    plus the fixed parameters (centre).  ``apply_to_image`` resamples with the B200 kernel."""

    transform_type = "AffineTransform"
    dimension = 3

    def __init__(self, parameters=None, fixed_parameters=None):
        self.parameters = (np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
                           if parameters is None else np.asarray(parameters, dtype=np.float64).copy())
        self.fixed_parameters = (np.zeros(3) if fixed_parameters is None
                                 else np.asarray(fixed_parameters, dtype=np.float64).copy())

    def set_parameters(self, parameters):
        parameters = np.asarray(parameters, dtype=np.float64).ravel()
        if parameters.size != 12:
            raise ValueError("an affine transform has 12 parameters")
        self.parameters = parameters.copy()

    def set_fixed_parameters(self, fixed_parameters):
        self.fixed_parameters = np.asarray(fixed_parameters, dtype=np.float64).ravel().copy()

    def as_matrix(self) -> np.ndarray:
        return convert_transform_to_numpy(self)

    def invert(self):
        """``ants.ANTsTransform.invert()`` (reference biahub/estimate_registration.py:190, 330):
        the inverse map as a new transform with a zero centre."""
        inv = np.linalg.inv(self.as_matrix())
        return ItkAffineParameters(np.concatenate([inv[:3, :3].ravel(), inv[:3, 3]]))

    def apply_to_image(self, image, reference=None, interpolation="linear"):
        """``ants.ANTsTransform.apply_to_image`` for the estimation loops that warp inside their
        optimisation (reference biahub/optimize_registration.py:111, registration/ants.py:204,
        registration/beads.py:117,194,920,966, estimate_registration.py:192,332).  Accepts numpy
        arrays or image objects with ``.numpy()`` (ANTs images); an image object in gives an
        image object out, so ``t.apply_to_image(ants.from_numpy(a), reference=r).numpy()`` works
        unchanged."""
        wrapped = hasattr(image, "numpy") and not isinstance(image, np.ndarray)
        arr = np.asarray(image.numpy() if wrapped else image)
        if reference is None:
            shape = arr.shape
        else:
            shape = tuple(reference.shape) if hasattr(reference, "shape") else np.shape(reference)
        out = affine_warp(arr, self.as_matrix(), tuple(int(v) for v in shape),
                          order=_interpolation_order(interpolation), boundary="itk")
        return WarpedImage(out) if wrapped else out


class WarpedImage:
    """Minimal image object returned by ``ItkAffineParameters.apply_to_image`` when it was given
    one: ``.numpy()``, ``.shape`` and ``np.asarray`` support, which is all the reference's call
    sites use of the ANTs image that comes back."""

    def __init__(self, array):
        self._array = array
        self.shape = array.shape
        self.dimension = array.ndim

    def numpy(self):
        return self._array

    def __array__(self, dtype=None, copy=None):
        return self._array if dtype is None else self._array.astype(dtype)


def convert_transform_to_ants(T_numpy: np.ndarray):
    """4x4 numpy → ITK-style parameters (reference register.py:148-168)."""
    T_numpy = np.asarray(T_numpy, dtype=np.float64)
    assert T_numpy.shape == (4, 4)
    params = np.concatenate([T_numpy[:3, :3].ravel(), T_numpy[:3, 3]])
    return ItkAffineParameters(params)


def convert_transform_to_numpy(T_ants) -> np.ndarray:
    """ITK-style parameters → 4x4 numpy, folding the centre into the translation
    (reference register.py:171-199: t += (I - A) @ centre)."""
    p = np.asarray(T_ants.parameters, dtype=np.float64)
    A = p[:9].reshape(3, 3)
    t = p[9:12] + (np.eye(3) - A) @ np.asarray(T_ants.fixed_parameters, dtype=np.float64)
    M = np.eye(4)
    M[:3, :3] = A
    M[:3, 3] = t
    return M


# ------------------------------------------------------------------------------------------
# the warp
# ------------------------------------------------------------------------------------------
def _interpolation_order(interpolation) -> int:
    key = str(interpolation).lower()
    if key not in _INTERPOLATION_ORDER:
        raise NotImplementedError(
            f"interpolation={interpolation!r}: the B200 path implements ANTs' 'linear' and "
            f"'nearestneighbor' (the modes BASELINE.json's north_star specifies) and, through "
            f"method='scipy', the cubic B-spline; ANTs' gaussian / bspline / windowed-sinc / "
            f"label interpolators are not built and there is no CPU fallback")
    return _INTERPOLATION_ORDER[key]


def _crop_box(output_shape_zyx, crop_output_slicing):
    shape = tuple(int(v) for v in output_shape_zyx)
    if len(shape) != 3:
        raise ValueError("output_shape_zyx must have 3 entries")
    if crop_output_slicing is None:
        return (0, 0, 0), shape
    starts, sizes = [], []
    for sl, n in zip(crop_output_slicing, shape):
        start, stop, step = sl.indices(n)
        if step != 1:
            raise ValueError("crop_output_slicing must use unit steps")
        starts.append(start)
        sizes.append(max(stop - start, 0))
    return tuple(starts), tuple(sizes)


def affine_warp(data, matrix, output_shape_zyx, order: int = 1, boundary: str = "itk",
                crop_output_slicing=None, scrub_nonfinite: bool = True, device=None,
                out=None, row_align: int = 1, _path: int = _cabi.PATH_AUTO):
    """Pull-warp a (Z, Y, X) volume: numpy in → numpy float32 out (host pipeline), or CUDA
    tensor in → CUDA float32 tensor out (kernel only, on the current stream).

    boundary: ``"itk"`` (ANTs/ITK rule) or ``"constant"`` (scipy ``mode="constant", cval=0``).
    """
    bcode = {"itk": _cabi.BOUNDARY_ITK, "constant": _cabi.BOUNDARY_CONSTANT}.get(boundary)
    if bcode is None:
        raise ValueError(f"boundary must be 'itk' or 'constant', got {boundary!r}")
    if order not in (0, 1):
        raise NotImplementedError("only order 0 (nearest) and 1 (linear) are implemented")
    starts, sizes = _crop_box(output_shape_zyx, crop_output_slicing)
    m12 = _cabi.matrix12(matrix)
    crop = _cabi.int64x3(starts)
    lib = _cabi.lib()

    if is_torch_tensor(data):
        import torch

        if data.ndim != 3:
            raise ValueError("expected a (Z, Y, X) tensor")
        src, code = device_source(data, allow_pitched=True)
        with torch.cuda.device(src.device):
            out = padded_empty(sizes, src.device, max(1, int(row_align)))
            if out.numel():
                _cabi.check(lib.b2_affine3d_pitched(
                    src.data_ptr(), code, row_pitch(src), *src.shape, out.data_ptr(),
                    row_pitch(out), *sizes, m12, crop, int(order), bcode,
                    int(bool(scrub_nonfinite)), int(_path),
                    torch.cuda.current_stream().cuda_stream))
        return out

    src, code = host_source(data)
    if src.ndim != 3:
        raise ValueError("expected a (Z, Y, X) array")
    out = check_out(out, sizes)
    if out.size:
        _cabi.check(lib.b2h_affine3d(
            src.ctypes.data_as(ctypes.c_void_p), code, *src.shape,
            out.ctypes.data_as(ctypes.c_void_p), *sizes, m12, crop, int(order), bcode,
            int(bool(scrub_nonfinite)), resolve_device(device)))
    return out


def _scipy_round(values: np.ndarray, dtype) -> np.ndarray:
    """scipy.ndimage's conversion of an interpolated value to an integer output dtype
    (NI_GeometricTransform): unsigned ``t > 0 ? t + 0.5 : 0``, signed ``t ± 0.5``, clamped to the
    dtype's range, truncated."""
    info = np.iinfo(dtype)
    t = values.astype(np.float64)
    t = np.where(t > 0, t + 0.5, 0.0 if info.min == 0 else t - 0.5)
    return np.trunc(np.clip(t, info.min, info.max)).astype(dtype)


def spline_warp(data, matrix, crop_output_slicing=None, device=None, out=None):
    """Cubic B-spline pull-warp of a (Z, Y, X) volume onto ITS OWN grid — the arithmetic of
    reference ``method="scipy"`` (register.py:271-272: scipy order 3, ``mode="constant"``,
    ``cval=0``, prefilter on; output dtype = input dtype).  numpy in → numpy out (host pipeline
    ``b2h_affine3d_spline3``) or CUDA tensor in → CUDA tensor out (``b2_affine3d_spline3``).

    uint16 → uint16 (scipy's round-half-up conversion in the kernel) and float32 → float32 run
    natively; float64 is computed from its float32 cast and returned as float64; other integer
    dtypes are computed in float32 and converted with scipy's rule on the host."""
    lib = _cabi.lib()
    m12 = _cabi.matrix12(matrix)
    if is_torch_tensor(data):
        import torch

        if data.ndim != 3:
            raise ValueError("expected a (Z, Y, X) tensor")
        src, code = device_source(data)
        starts, sizes = _crop_box(tuple(src.shape), crop_output_slicing)
        with torch.cuda.device(src.device):
            res = torch.empty(sizes, dtype=torch.uint16 if code == _cabi.DTYPE_U16 else torch.float32,
                              device=src.device)
            if res.numel():
                nbytes = int(lib.b2_spline3_workspace(*src.shape))
                ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=src.device)
                ws_ptr = (ws.data_ptr() + 255) // 256 * 256
                stream = torch.cuda.current_stream()
                _cabi.check(lib.b2_affine3d_spline3(
                    src.data_ptr(), code, *src.shape, res.data_ptr(), code, *sizes, m12,
                    _cabi.int64x3(starts), 1, ws_ptr, nbytes, stream.cuda_stream))
                ws.record_stream(stream)
        return res
    arr = np.asarray(data)
    if arr.ndim != 3:
        raise ValueError("expected a (Z, Y, X) array")
    in_dtype = arr.dtype
    if in_dtype == np.float64:
        arr = np.nan_to_num(arr, nan=0)   # reference register.py:254, in the input dtype
    src, code = host_source(arr)
    starts, sizes = _crop_box(src.shape, crop_output_slicing)
    native = in_dtype in (np.uint16, np.float32)
    res = check_out(out if native else None, sizes, np.uint16 if code == _cabi.DTYPE_U16 else np.float32)
    if res.size:
        _cabi.check(lib.b2h_affine3d_spline3(
            src.ctypes.data_as(ctypes.c_void_p), code, *src.shape,
            res.ctypes.data_as(ctypes.c_void_p), code, *sizes, m12, _cabi.int64x3(starts), 1,
            resolve_device(device)))
    if native:
        return res
    if in_dtype.kind in "ui":
        return _scipy_round(res, in_dtype)
    return res.astype(in_dtype)


def apply_affine_transform(
    zyx_data: np.ndarray,
    matrix: np.ndarray,
    output_shape_zyx: tuple,
    method="ants",
    interpolation: str = "linear",
    crop_output_slicing: bool = None,
) -> np.ndarray:
    """Drop-in for reference ``apply_affine_transform`` (register.py:202-281).

    3-D (Z, Y, X) or 4-D (C, Z, Y, X) input; NaNs are scrubbed to 0 (and ±inf to ±float32 max,
    ``np.nan_to_num`` semantics) on load.  ``method="ants"``: float32 result on
    ``output_shape_zyx`` cropped to ``crop_output_slicing``.  ``method="scipy"``: cubic spline on
    the INPUT's grid in the input's dtype, then cropped (see the module docstring); a 4-D input
    is assigned channel by channel into a float32 ``(C,) + cropped output_shape`` array exactly
    as the reference does, so a mismatch raises numpy's broadcast ``ValueError`` there too.
    """
    if method == "ants":
        order = _interpolation_order(interpolation)
        boundary = "itk"
    elif method != "scipy":
        raise ValueError(f"Unknown method {method}")

    ndim = zyx_data.ndim
    if ndim == 4:
        if method == "ants":
            _, sizes = _crop_box(output_shape_zyx, crop_output_slicing)
        elif crop_output_slicing is None:
            sizes = tuple(int(v) for v in output_shape_zyx)
        else:  # reference register.py:233-238: stop - start, unclipped
            sizes = tuple(int(sl.stop - sl.start) for sl in crop_output_slicing)
        if is_torch_tensor(zyx_data):
            import torch

            return torch.stack([
                apply_affine_transform(zyx_data[c], matrix, output_shape_zyx, method, interpolation,
                                       crop_output_slicing).to(torch.float32)
                for c in range(zyx_data.shape[0])])
        registered = np.zeros((zyx_data.shape[0],) + sizes, dtype=np.float32)
        for c in range(zyx_data.shape[0]):
            registered[c] = apply_affine_transform(zyx_data[c], matrix, output_shape_zyx, method,
                                                   interpolation, crop_output_slicing)
        return registered
    if ndim != 3:
        raise ValueError("zyx_data must be (Z, Y, X) or (C, Z, Y, X)")
    if method == "scipy":
        return spline_warp(zyx_data, matrix, crop_output_slicing)
    return affine_warp(zyx_data, matrix, output_shape_zyx, order, boundary, crop_output_slicing)


# ------------------------------------------------------------------------------------------
# output-shape logic of `biahub register` when keep_overhang=False (reference register.py:284-394)
# ------------------------------------------------------------------------------------------
def largest_interior_rectangle(mask2d) -> tuple:
    """Largest axis-aligned all-True rectangle of a 2-D boolean mask → ``(x, y, width, height)``
    (the return convention of the third-party ``largestinteriorrectangle.lir`` the reference
    calls at register.py:289-309).  Histogram/stack sweep, O(H*W); ties keep the first
    rectangle found scanning rows top to bottom."""
    m = np.asarray(mask2d, dtype=bool)
    if m.ndim != 2:
        raise ValueError("mask must be 2-D")
    H, W = m.shape
    heights = np.zeros(W + 1, dtype=np.int64)  # sentinel column of height 0
    best = (0, 0, 0, 0)
    best_area = 0
    for row in range(H):
        heights[:W] = np.where(m[row], heights[:W] + 1, 0)
        stack = []  # (start column, height)
        for col in range(W + 1):
            h = int(heights[col])
            start = col
            while stack and stack[-1][1] >= h:
                s0, h0 = stack.pop()
                area = h0 * (col - s0)
                if area > best_area:
                    best_area = area
                    best = (s0, row - h0 + 1, col - s0, h0)
                start = s0
            if h > 0:
                stack.append((start, h))
    return best


def find_lir(registered_zyx: np.ndarray, plot: bool = False) -> tuple:
    """ZYX slices of a large interior cuboid of a boolean volume, with the reference's recipe
    (register.py:284-342): largest interior rectangle in YX at Z//2, then the Z extent common to
    the ZY / ZX rectangles probed at three x and three y positions."""
    vol = np.asarray(registered_zyx, dtype=bool)
    x, y, width, height = largest_interior_rectangle(vol[vol.shape[0] // 2])
    x_start, x_stop = x, x + width
    y_start, y_stop = y, y + height
    x_slice, y_slice = slice(x_start, x_stop), slice(y_start, y_stop)
    z_ranges = []
    for _x in (x_start, x_start + (x_stop - x_start) // 2, x_stop - 1):
        _, z, _, depth = largest_interior_rectangle(vol[:, y_slice, _x])
        z_ranges.append((z, z + depth))
    for _y in (y_start, y_start + (y_stop - y_start) // 2, y_stop - 1):
        _, z, _, depth = largest_interior_rectangle(vol[:, _y, x_slice])
        z_ranges.append((z, z + depth))
    z_ranges = np.asarray(z_ranges)
    z_slice = slice(int(z_ranges[:, 0].max()), int(z_ranges[:, 1].min()))
    return (z_slice, y_slice, x_slice)


def find_overlapping_volume(input_zyx_shape, target_zyx_shape, transformation_matrix,
                            method: str = "LIR", plot: bool = False, device=None) -> tuple:
    """ZYX slices of the overlap of a warped source with the target grid (reference
    register.py:345-394): warp a volume of ones onto the target grid, ``mask = warped > 0``, then
    ``find_lir``.  The ones volume is created and warped on the GPU (uint16 ones, linear
    interpolation, ITK boundary rule); only the boolean mask comes back to the host."""
    import torch

    if method != "LIR":
        raise ValueError(f"Unknown method {method}")
    dev = torch.device("cuda", resolve_device(device))
    ones = torch.ones(tuple(int(v) for v in input_zyx_shape), dtype=torch.float32, device=dev)
    warped = affine_warp(ones, transformation_matrix, tuple(int(v) for v in target_zyx_shape),
                         order=1, boundary="itk", scrub_nonfinite=False)
    mask = (warped > 0).cpu().numpy()
    return find_lir(mask, plot=plot)
