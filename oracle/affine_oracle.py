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
    """``np.nan_to_num(x, nan=0)`` then float32 — reference biahub/register.py:254, 266;
    biahub/stabilize.py:84-85.  NaN→0, +inf→FLT_MAX, −inf→−FLT_MAX (after the float32 cast
    float64 maxima would overflow, so the scrub is applied in the input dtype like numpy does,
    then re-applied after the cast for float64 inputs whose maxima exceed float32)."""
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
    src = scrub_nonfinite(vol).astype(np.float64)
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
        return scrub_nonfinite(src[iz, iy, ix]).astype(np.float64)

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
