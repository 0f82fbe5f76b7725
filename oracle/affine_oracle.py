"""TEST INFRASTRUCTURE ONLY — CPU oracles for the affine warp of ``biahub register`` /
``biahub stabilize`` (reference biahub/register.py:202-281, biahub/stabilize.py:32-90).

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.

Where the arithmetic lives: the reference delegates to third-party libraries that are not in
``/root/reference`` —

* ``method="scipy"``: ``scipy.ndimage.affine_transform`` (scipy 1.15.3 pinned in the reference's
  ``uv.lock``; scipy 1.18.1 is installed here and on the GPU box).  ``affine_oracle_scipy`` calls
  that library directly with ``order`` ∈ {0, 1}, ``mode="constant"``, ``cval=0`` — this is the
  PRIMARY oracle (float64 coordinates and weights).
* ``method="ants"`` (default): antspyx 0.6.1 → ITK ``ResampleImageFilter`` with
  ``LinearInterpolateImageFunction`` / ``NearestNeighborInterpolateImageFunction``.  antspyx is
  not installable here, so the ITK boundary rule is restated from its published algorithm
  (``boundary="itk"`` below).  **Parity unpinned at the ANTs boundary**: the only numeric pin the
  reference holds is ``tests/test_affine.py:43-59`` (translation known answer), which
  ``tests/test_oracle_golden.py`` checks for both modes.

``affine_oracle_numpy`` restates both boundary rules in float64 and is itself pinned against
``scipy.ndimage`` (bit-exact for order 0, ≤ 1 ulp(float32) for order 1) in the CPU tests.

Matrix convention (reference biahub/register.py:148-168, pinned by tests/test_affine.py:43-59):
4x4 homogeneous, ZYX, pull — source index = M[:3,:3] @ out_index + M[:3,3]; no centring,
no axis reordering.
"""

from __future__ import annotations

import numpy as np

_f32 = np.float32

FLT_MAX = float(np.finfo(np.float32).max)


def scrub_nonfinite(vol):
    """``np.nan_to_num(x, nan=0)`` in the INPUT dtype, then ``astype(float32)`` — the reference's
    order (biahub/register.py:254, 266; biahub/stabilize.py:84-85).  NaN→0, +inf→max, −inf→−max
    of the input dtype.  For float32 inputs that is ±FLT_MAX.  For float64 inputs the cast comes
    AFTER the scrub, so ±inf (scrubbed to ±DBL_MAX) and finite magnitudes beyond FLT_MAX reach
    the resampler as float32 ±inf: this function returns them as such.  What ITK then does with
    an infinite tap is unpinned; ``affine_oracle_numpy`` (and the CUDA path) scrub once more at
    the tap, i.e. treat them as ±FLT_MAX (golden case ``f64_overflow``)."""
    vol = np.nan_to_num(np.asarray(vol), nan=0)
    with np.errstate(over="ignore"):
        vol = vol.astype(_f32)
    return vol


def _crop_box(output_shape_zyx, crop_output_slicing):
    if crop_output_slicing is None:
        return (0, 0, 0), tuple(int(v) for v in output_shape_zyx)
    starts, sizes = [], []
    for sl, n in zip(crop_output_slicing, output_shape_zyx):
        start, stop, step = sl.indices(int(n))
        if step != 1:
            raise ValueError("crop slices must have unit step")
        starts.append(start)
        sizes.append(max(stop - start, 0))
    return tuple(starts), tuple(sizes)


def affine_oracle_scipy(vol, matrix, output_shape_zyx, order=1, crop_output_slicing=None):
    """PRIMARY oracle: the reference's ``method="scipy"`` library call
    (biahub/register.py:271-272) with order 0/1 instead of the default 3."""
    import scipy.ndimage

    src = scrub_nonfinite(vol)
    out = scipy.ndimage.affine_transform(
        src, np.asarray(matrix, dtype=np.float64), output_shape=tuple(output_shape_zyx),
        order=int(order), mode="constant", cval=0.0, prefilter=False, output=np.float32,
    )
    if crop_output_slicing is not None:
        out = out[tuple(crop_output_slicing)]
    return np.ascontiguousarray(out)


def affine_oracle_numpy(vol, matrix, output_shape_zyx, order=1, boundary="constant",
                        crop_output_slicing=None, z_chunk=8):
    """float64 restatement of the pull warp with either boundary rule.

    boundary="constant" (scipy ``mode="constant", cval=0``):
        any coordinate outside [0, n-1] (strictly) → 0.
        order 1: f=floor(c), w=c-f, trilinear; a +1 tap at c==n-1 has zero weight.
        order 0: index floor(c+0.5).
    boundary="itk" (ITK ResampleImageFilter + Linear/NearestNeighbor interpolators):
        inside iff -0.5 <= c < n-0.5 on every axis, else 0 (default pixel value);
        order 1: clamp-to-edge inside the half-voxel band (base clamped to >= 0, a neighbour
        beyond n-1 is dropped, non-positive distance → no blend);
        order 0: index floor(c+0.5) (round-half-up).

    Coordinates follow scipy's op order (probed on scipy 1.18.1: the accumulator starts at the
    shift): c = ((shift + z*m0) + y*m1) + x*m2 in float64 with separate multiply/add — this
    order reproduces ``scipy.ndimage`` order-0 output bit for bit on a generic matrix, the
    shift-last order does not (tests/test_oracle_golden.py).
    """
    if boundary not in ("constant", "itk"):
        raise ValueError(boundary)
    src = np.nan_to_num(scrub_nonfinite(vol), nan=0).astype(np.float64)  # tap scrub, see above
    M = np.asarray(matrix, dtype=np.float64)
    A, t = M[:3, :3], M[:3, 3]
    n = src.shape
    start, size = _crop_box(output_shape_zyx, crop_output_slicing)
    out = np.zeros(size, dtype=_f32)
    if 0 in size:
        return out
    ys = np.arange(start[1], start[1] + size[1], dtype=np.float64)[None, :, None]
    xs = np.arange(start[2], start[2] + size[2], dtype=np.float64)[None, None, :]
    for z0 in range(0, size[0], z_chunk):
        z1 = min(z0 + z_chunk, size[0])
        zs = np.arange(start[0] + z0, start[0] + z1, dtype=np.float64)[:, None, None]
        c = []
        for d in range(3):
            c.append(((t[d] + zs * A[d, 0]) + ys * A[d, 1]) + xs * A[d, 2])
        if boundary == "constant":
            inside = np.ones(c[0].shape, dtype=bool)
            for d in range(3):
                inside &= (c[d] >= 0.0) & (c[d] <= n[d] - 1)
        else:
            inside = np.ones(c[0].shape, dtype=bool)
            for d in range(3):
                inside &= (c[d] >= -0.5) & (c[d] < n[d] - 0.5)
        if order == 0:
            idx = [np.clip(np.floor(c[d] + 0.5).astype(np.int64), 0, n[d] - 1) for d in range(3)]
            val = src[idx[0], idx[1], idx[2]]
        elif order == 1:
            base, frac, nxt = [], [], []
            for d in range(3):
                b = np.floor(c[d])
                b = np.clip(b, 0, n[d] - 1)  # itk: base clamped to start; constant: no-op when inside
                w = np.clip(c[d] - b, 0.0, 1.0)  # itk: distance <= 0 → 0
                bi = b.astype(np.int64)
                ni = np.minimum(bi + 1, n[d] - 1)  # +1 neighbour beyond the edge is dropped
                w = np.where(bi + 1 > n[d] - 1, 0.0, w)
                base.append(bi)
                nxt.append(ni)
                frac.append(w)
            val = np.zeros(c[0].shape, dtype=np.float64)
            for dz in (0, 1):
                iz = nxt[0] if dz else base[0]
                wz = frac[0] if dz else 1.0 - frac[0]
                for dy in (0, 1):
                    iy = nxt[1] if dy else base[1]
                    wy = frac[1] if dy else 1.0 - frac[1]
                    for dx in (0, 1):
                        ix = nxt[2] if dx else base[2]
                        wx = frac[2] if dx else 1.0 - frac[2]
                        val += src[iz, iy, ix] * (wz * wy * wx)
        else:
            raise ValueError("order must be 0 or 1")
        with np.errstate(over="ignore"):
            out[z0:z1] = np.where(inside, val, 0.0).astype(_f32)
    return out


# ---- reference matrix helpers restated for building test matrices ------------------------
def rotation_matrix_yx(shape_zyx, angle_deg, end_shape_zyx=None):
    """In-plane (YX) rotation about shape/2 — convention of reference biahub/register.py:60-111."""
    cy, cx = np.array(shape_zyx)[-2:] / 2
    ey, ex = (cy, cx) if end_shape_zyx is None else np.array(end_shape_zyx)[-2:] / 2
    th = np.radians(angle_deg)
    c, s = np.cos(th), np.sin(th)
    return np.array([
        [1, 0, 0, 0],
        [0, c, -s, -cy * c + s * cx + ey],
        [0, s, c, -cy * s - cx * c + ex],
        [0, 0, 0, 1],
    ], dtype=np.float64)


def scaling_matrix_zyx(shape_zyx, scale_zyx=(1, 1, 1), end_shape_zyx=None):
    """Scaling about the YX centre — convention of reference biahub/register.py:32-57."""
    cy, cx = np.array(shape_zyx)[-2:] / 2
    ey, ex = (cy, cx) if end_shape_zyx is None else np.array(end_shape_zyx)[-2:] / 2
    sz, sy, sx = scale_zyx
    return np.array([
        [sz, 0, 0, 0],
        [0, sy, 0, -cy * sy + ey],
        [0, 0, sx, -cx * sx + ex],
        [0, 0, 0, 1],
    ], dtype=np.float64)


def translation_matrix_zyx(shift_zyx):
    M = np.eye(4)
    M[:3, 3] = shift_zyx
    return M


def register_matrix_c3(shape_zyx, angle_deg=7.3, scale_yx=1.07, shift_zyx=(0.4, 3.25, -11.5)):
    """The C3 benchmark matrix of SURVEY.md §8(d): rotate·scale then translate (pull form)."""
    return (translation_matrix_zyx(shift_zyx)
            @ rotation_matrix_yx(shape_zyx, angle_deg)
            @ scaling_matrix_zyx(shape_zyx, (1.0, scale_yx, scale_yx)))


def affine_oracle_points(vol, matrix, points, order=1, boundary="constant"):
    """``affine_oracle_numpy`` evaluated only at ``points`` — (M, 3) int array of UNCROPPED output
    indices (z, y, x).  For spot-checking full-size volumes."""
    src = np.asarray(vol)
    M = np.asarray(matrix, dtype=np.float64)
    A, t = M[:3, :3], M[:3, 3]
    n = src.shape
    pts = np.asarray(points, dtype=np.float64)
    c = []
    for d in range(3):
        c.append(((t[d] + pts[:, 0] * A[d, 0]) + pts[:, 1] * A[d, 1]) + pts[:, 2] * A[d, 2])
    inside = np.ones(len(pts), dtype=bool)
    for d in range(3):
        if boundary == "constant":
            inside &= (c[d] >= 0.0) & (c[d] <= n[d] - 1)
        else:
            inside &= (c[d] >= -0.5) & (c[d] < n[d] - 0.5)

    def tap(iz, iy, ix):
        return np.nan_to_num(scrub_nonfinite(src[iz, iy, ix]), nan=0).astype(np.float64)

    if order == 0:
        idx = [np.clip(np.floor(c[d] + 0.5).astype(np.int64), 0, n[d] - 1) for d in range(3)]
        val = tap(*idx)
    else:
        base, nxt, frac = [], [], []
        for d in range(3):
            b = np.clip(np.floor(np.where(inside, c[d], 0.0)), 0, n[d] - 1)
            w = np.clip(np.where(inside, c[d], 0.0) - b, 0.0, 1.0)
            bi = b.astype(np.int64)
            w = np.where(bi + 1 > n[d] - 1, 0.0, w)
            base.append(bi)
            nxt.append(np.minimum(bi + 1, n[d] - 1))
            frac.append(w)
        val = np.zeros(len(pts))
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    wz = frac[0] if dz else 1.0 - frac[0]
                    wy = frac[1] if dy else 1.0 - frac[1]
                    wx = frac[2] if dx else 1.0 - frac[2]
                    val += tap(nxt[0] if dz else base[0], nxt[1] if dy else base[1],
                               nxt[2] if dx else base[2]) * (wz * wy * wx)
    with np.errstate(over="ignore"):
        return np.where(inside, val, 0.0).astype(_f32)


# ---- method="scipy": cubic B-spline, scipy ``mode="constant"`` (reference register.py:271-272) ----
SPLINE3_POLE = float(np.sqrt(3.0) - 2.0)


def spline3_prefilter_numpy(vol):
    """``scipy.ndimage.spline_filter(vol, 3, output=float64, mode="constant")`` restated: per
    axis (length > 1) scale by (1-z)(1-1/z) = 6, causal pass ``c[i] += z c[i-1]`` started from
    the mirror-extended sum ``c[0] = sum_k z^k s[mirror(k)]``, anti-causal pass
    ``c[i] = z (c[i+1] - c[i])`` started from ``c[n-1] = z/(z^2-1) (z c[n-2] + c[n-1])``; scipy
    uses the MIRROR initialisation for ``mode="constant"`` (probed on scipy 1.18.1: identical
    bits to ``mode="mirror"``).  Pinned to scipy ≤ 1e-15 relative in tests/test_affine_reference.py."""
    c = np.asarray(vol, dtype=np.float64).copy()
    z = SPLINE3_POLE
    for axis in range(c.ndim):
        n = c.shape[axis]
        if n < 2:
            continue
        c = np.moveaxis(c, axis, 0)
        c *= (1.0 - z) * (1.0 - 1.0 / z)
        k = np.arange(64)  # z^64 ~ 3e-37: the infinite mirror sum to double precision
        idx = k % (2 * n - 2)
        idx = np.where(idx >= n, 2 * n - 2 - idx, idx)
        c[0] = np.tensordot(z ** k, c[idx], axes=(0, 0))
        for i in range(1, n):
            c[i] += z * c[i - 1]
        c[n - 1] = (z / (z * z - 1.0)) * (z * c[n - 2] + c[n - 1])
        for i in range(n - 2, -1, -1):
            c[i] = z * (c[i + 1] - c[i])
        c = np.moveaxis(c, 0, axis)
    return c


def _bspline3_weights(t):
    return ((1 - t) ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6,
            (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6, t ** 3 / 6)


def _round_like_scipy(val, dtype):
    """Output conversion of scipy's NI_GeometricTransform: floats cast; unsigned
    ``t > 0 ? t + 0.5 : 0`` clamped then truncated; signed ``t ± 0.5`` clamped then truncated."""
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        with np.errstate(over="ignore"):
            return val.astype(dtype)
    info = np.iinfo(dtype)
    if dtype.kind == "u":
        t = np.where(val > 0, val + 0.5, 0.0)
    else:
        t = np.where(val > 0, val + 0.5, val - 0.5)
    return np.trunc(np.clip(t, info.min, info.max)).astype(dtype)


def affine_oracle_spline3(vol, matrix, crop_output_slicing=None, z_chunk=4):
    """The reference's ``method="scipy"`` call ``scipy.ndimage.affine_transform(zyx, M, shape)``
    restated: the third positional argument is scipy's ``offset`` (ignored for a homogeneous 4x4
    matrix), so the output grid is ALWAYS the input's shape, order 3, ``mode="constant"``,
    ``cval=0``, prefilter on, output dtype = input dtype.  Coordinates outside [0, n-1] → 0; the
    4x4x4 taps ``floor(c)-1 … floor(c)+2`` are mirror-extended."""
    vol = np.asarray(vol)
    coef = spline3_prefilter_numpy(vol)
    M = np.asarray(matrix, dtype=np.float64)
    A, t = M[:3, :3], M[:3, 3]
    n = vol.shape
    start, size = _crop_box(n, crop_output_slicing)
    out = np.zeros(size, dtype=vol.dtype)
    if 0 in size:
        return out

    def mirror(i, nn):
        if nn == 1:
            return np.zeros_like(i)
        p = 2 * nn - 2
        i = np.mod(i, p)
        return np.where(i >= nn, p - i, i)

    ys = np.arange(start[1], start[1] + size[1], dtype=np.float64)[None, :, None]
    xs = np.arange(start[2], start[2] + size[2], dtype=np.float64)[None, None, :]
    for z0 in range(0, size[0], z_chunk):
        z1 = min(z0 + z_chunk, size[0])
        zs = np.arange(start[0] + z0, start[0] + z1, dtype=np.float64)[:, None, None]
        c = [((t[d] + zs * A[d, 0]) + ys * A[d, 1]) + xs * A[d, 2] for d in range(3)]
        inside = np.ones(c[0].shape, dtype=bool)
        for d in range(3):
            inside &= (c[d] >= 0.0) & (c[d] <= n[d] - 1)
        fl = [np.floor(np.where(inside, c[d], 0.0)) for d in range(3)]
        W = [_bspline3_weights(np.where(inside, c[d], 0.0) - fl[d]) for d in range(3)]
        val = np.zeros(c[0].shape, dtype=np.float64)
        for a0 in range(4):
            iz = mirror(fl[0].astype(np.int64) - 1 + a0, n[0])
            for a1 in range(4):
                iy = mirror(fl[1].astype(np.int64) - 1 + a1, n[1])
                for a2 in range(4):
                    ix = mirror(fl[2].astype(np.int64) - 1 + a2, n[2])
                    val += coef[iz, iy, ix] * (W[0][a0] * W[1][a1] * W[2][a2])
        out[z0:z1] = _round_like_scipy(np.where(inside, val, 0.0), vol.dtype)
    return out


# ---- the reference's wrappers restated (register.py:202-281, stabilize.py:32-90) ----------
_ANTS_ORDER = {"linear": 1, "nearestneighbor": 0}


def apply_affine_transform_oracle(zyx_data, matrix, output_shape_zyx, method="ants",
                                  interpolation="linear", crop_output_slicing=None):
    """Restatement of the reference's ``apply_affine_transform`` with the ANTs call replaced by
    the ITK-rule oracle.  Pinned against the reference function itself (run with the fake ants of
    oracle/ref_loader.py) through tests/golden/golden_affine_v1.npz."""
    zyx_data = np.asarray(zyx_data)
    if zyx_data.ndim == 4:
        _, size = _crop_box(output_shape_zyx, crop_output_slicing)
        out = np.zeros((zyx_data.shape[0],) + size, dtype=_f32)
        for c in range(zyx_data.shape[0]):
            out[c] = apply_affine_transform_oracle(zyx_data[c], matrix, output_shape_zyx, method,
                                                   interpolation, crop_output_slicing)
        return out
    if method == "ants":
        return affine_oracle_numpy(zyx_data, matrix, output_shape_zyx, _ANTS_ORDER[interpolation],
                                   "itk", crop_output_slicing)
    if method == "scipy":
        return affine_oracle_spline3(np.nan_to_num(zyx_data, nan=0), matrix, crop_output_slicing)
    raise ValueError(f"Unknown method {method}")


def apply_stabilization_oracle(zyx_data, list_of_shifts, input_time_index, output_shape=None):
    """Restatement of the reference's ``apply_stabilization_transform`` (stabilize.py:32-90)."""
    zyx_data = np.asarray(zyx_data)
    if output_shape is None:
        output_shape = zyx_data.shape[-3:]
    M = np.asarray(list_of_shifts[input_time_index], dtype=np.float64)
    if zyx_data.ndim == 4:
        return np.stack([affine_oracle_numpy(zyx_data[c], M, output_shape, 1, "itk")
                         for c in range(zyx_data.shape[0])]).astype(_f32)
    return affine_oracle_numpy(zyx_data, M, output_shape, 1, "itk")
