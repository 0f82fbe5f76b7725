"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's deskew path.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  ``biahub_b200/`` must never import it (the product path fails loudly
when the CUDA library is missing; there is no CPU fallback).

Parity pin: ``tests/golden/*.npz`` were produced by running the *unmodified*
reference (``/root/reference/biahub/deskew.py`` → ``_fast_deskew_czyx(device="cpu")``)
through ``oracle/ref_loader.py`` (script: ``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` checks both restatements below against them and
against the reference's own known-answer tests
(reference ``tests/test_cli/test_deskew_cli.py:11-59, 189-204``).
On the x86 hosts used so far ``deskew_oracle_numpy`` is BIT-IDENTICAL to the
reference for ``average_n_slices`` ≤ 4 (≤ 2e-7 of range above that: torch's
reduction order over 5+ elements is not sequential).

Two restatements:

* ``deskew_oracle_numpy``  — explicit per-voxel formula with the reference's fp32
  rounding sequence.  This is the checker.
* ``deskew_oracle_torch``  — the same algorithm phrased with the library ops the
  reference uses on CPU (transpose copy, 2-D ``grid_sample``, ``mean``), so
  that its run time on the host cores is a fair stand-in for the reference's
  CPU path when the reference itself cannot travel to the GPU box.  Used as
  ``bench.py``'s ``cpu_baseline`` / reference arm (kind ``"port"``).
"""

from __future__ import annotations

import math

import numpy as np

_f32 = np.float32
_f64 = np.float64


# --------------------------------------------------------------------------------------
# shape logic — reference biahub/deskew.py:157-177 and 213-274
# --------------------------------------------------------------------------------------
def averaged_shape_oracle(shape, window):
    """reference biahub/deskew.py:157-177 (`_get_averaged_shape`)."""
    return (int(np.ceil(shape[0] / window)),) + tuple(shape[1:])


def deskewed_shape_oracle(raw_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang,
                          average_n_slices=1, pixel_size_um=1):
    """reference biahub/deskew.py:213-274 (`get_deskewed_data_shape`)."""
    theta = ls_angle_deg * np.pi / 180
    st, ct = np.sin(theta), np.cos(theta)
    Z, Y, X = raw_shape
    if keep_overhang:
        Xp = int(np.ceil((Z / px_to_scan_ratio) + (Y * ct)))
    else:
        Xp = int(np.ceil((Z / px_to_scan_ratio) - (Y * ct)))
        if Xp <= 0:
            raise ValueError(
                "Dataset contains only overhang when keep_overhang=False. "
                f"Computed Xp={Xp} <= 0."
            )
    out = averaged_shape_oracle((Y, X, Xp), average_n_slices)
    voxel = (average_n_slices * st * pixel_size_um, pixel_size_um, pixel_size_um)
    return out, voxel


def average_n_slices_oracle(data, window=1):
    """reference biahub/deskew.py:43-68 (`_average_n_slices`): edge-pad then group mean."""
    data = np.asarray(data)
    rem = data.shape[0] % window
    if rem:
        reps = np.repeat(data[-1:], window - rem, axis=0)
        data = np.concatenate([data, reps], axis=0)
    grouped = data.reshape((data.shape[0] // window, window) + data.shape[1:])
    return grouped.mean(axis=1)


# --------------------------------------------------------------------------------------
# coordinate pipeline — reference biahub/deskew.py:133-148 + ATen GridSampler unnormalize
# --------------------------------------------------------------------------------------
def deskew_scalars(raw_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang):
    """float64 host scalars of reference biahub/deskew.py:136-138, then rounded to fp32."""
    Zi, Yi, Xi = raw_shape
    (Zo, Yo, Xo), _ = deskewed_shape_oracle(raw_shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    ct = np.cos(ls_angle_deg * np.pi / 180)
    px = px_to_scan_ratio
    off = px * ct * (Zo - 1) / 2 - px * (Xo - 1) / 2 + (Zi - 1) / 2
    return dict(Zo=Zo, Yo=Yo, Xo=Xo, px32=_f32(px), pxct32=_f32(px * ct), off32=_f32(off))


def scan_coordinate_fp32(x_idx, zo_idx, Zi, px32, pxct32, off32):
    """p' for every (zo, x): fp32 ops in the reference's order.

    reference biahub/deskew.py:147  in_z_f  = px*x - px*ct*z_out + offset
    reference biahub/deskew.py:148  norm    = 2*in_z_f/(Z_in-1) - 1
    ATen grid_sampler (align_corners=True) un-normalise: ((g+1)/2)*(Z_in-1)
    """
    x = np.asarray(x_idx, dtype=_f32)
    zo = np.asarray(zo_idx, dtype=_f32)
    p = (px32 * x)[None, :] - (pxct32 * zo)[:, None]
    p = (p + off32).astype(_f32)
    g = ((_f32(2.0) * p) / _f32(Zi - 1)) - _f32(1.0)
    pp = ((g + _f32(1.0)) / _f32(2.0)) * _f32(Zi - 1)
    assert pp.dtype == _f32
    return pp


def _fma32(a, b, c):
    # fp32 fused multiply-add emulated through float64 (product exact; one extra rounding
    # that matters only in double-rounding corner cases).
    return (a.astype(_f64) * b.astype(_f64) + c.astype(_f64)).astype(_f32)


def deskew_oracle_numpy(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices=1):
    """Per-voxel restatement of reference ``fast_deskew_zyx`` (biahub/deskew.py:456-542)
    without the overhang fill.  ``raw``: (Z_scan, Y_tilt, X_coverslip) any numeric dtype.

    out[a,y,x] = ( Σ_k S(aN+k, y, x) ) / N            torch.mean over dim 2 (:536)
    S(zo,y,x)  = fma(T(f+1), w, T(f)*e)               ATen vectorised bilinear, zeros padding (:531-533)
    T(j)       = raw[j, Yi-1-min(zo,Zo-1), Xi-1-y]    _rearrange_axes (:99-110) + edge pad (:517-519)
    """
    raw = np.asarray(raw)
    if raw.ndim != 3:
        raise ValueError("raw must be (Z, Y, X)")
    Zi, Yi, Xi = raw.shape
    if Zi < 2:
        raise ValueError("Z_in must be >= 2 (reference divides by Z_in-1, biahub/deskew.py:148)")
    N = int(average_n_slices)
    s = deskew_scalars(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    Zo, Yo, Xo = s["Zo"], s["Yo"], s["Xo"]
    Zavg = int(np.ceil(Zo / N))

    pp = scan_coordinate_fp32(np.arange(Xo), np.arange(Zavg * N), Zi, s["px32"], s["pxct32"], s["off32"])
    f = np.floor(pp)
    w = (pp - f).astype(_f32)
    e = ((f + _f32(1.0)) - pp).astype(_f32)
    fi = f.astype(np.int64)

    rawf = raw.astype(_f32)
    out = np.empty((Zavg, Yo, Xo), dtype=_f32)
    for a in range(Zavg):
        acc = None
        for k in range(N):
            zo = a * N + k
            iy = Yi - 1 - min(zo, Zo - 1)
            plane = rawf[:, iy, ::-1]  # [j, y]
            j0 = fi[zo]
            j1 = j0 + 1
            in0 = (j0 >= 0) & (j0 <= Zi - 1)
            in1 = (j1 >= 0) & (j1 <= Zi - 1)
            t0 = np.where(in0[None, :], plane[np.clip(j0, 0, Zi - 1)].T, _f32(0))
            t1 = np.where(in1[None, :], plane[np.clip(j1, 0, Zi - 1)].T, _f32(0))
            E = np.broadcast_to(e[zo][None, :], t0.shape)
            W = np.broadcast_to(w[zo][None, :], t0.shape)
            sk = _fma32(t1, W, (t0 * E).astype(_f32))
            acc = sk if acc is None else (acc + sk).astype(_f32)
        out[a] = acc / _f32(N)
    return out


def deskew_oracle_torch(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices=1,
                        num_threads=None):
    """Library-op phrasing (CPU torch) of the same algorithm; the timed CPU baseline.

    Follows the stages of reference biahub/deskew.py:505-536: cast → axis permutation/flip
    copy → edge pad → (Z_avg, Y_out, N, Z_in) view → 2-D bilinear ``grid_sample``
    (zeros, align_corners=True) → mean over the N sub-slices.
    """
    import torch
    import torch.nn.functional as F

    if num_threads is not None:
        torch.set_num_threads(int(num_threads))
    raw = np.asarray(raw)
    Zi, Yi, Xi = raw.shape
    N = int(average_n_slices)
    s = deskew_scalars(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    Zo, Yo, Xo = s["Zo"], s["Yo"], s["Xo"]
    Zavg = int(math.ceil(Zo / N))

    vol = torch.from_numpy(raw).to(dtype=torch.float32)
    # (Z_scan, Y_tilt, X_cov) -> (zo, y, j) with zo = Yi-1-iy, y = Xi-1-ix
    ra = vol.permute(1, 2, 0).flip([0, 1]).contiguous()
    pad = Zavg * N - Zo
    if pad:
        ra = torch.cat([ra, ra[-1:].expand(pad, -1, -1)], dim=0)
    ra = ra.reshape(Zavg, N, Yo, Zi).permute(0, 2, 1, 3)

    x = torch.arange(Xo, dtype=torch.float32)
    zo = torch.arange(Zavg * N, dtype=torch.float32).reshape(Zavg, N)
    px = float(px_to_scan_ratio)
    ct = float(np.cos(ls_angle_deg * np.pi / 180))
    off = px * ct * (Zo - 1) / 2 - px * (Xo - 1) / 2 + (Zi - 1) / 2
    p = px * x - px * ct * zo.unsqueeze(2) + off
    gw = 2.0 * p / (Zi - 1) - 1.0
    k = torch.arange(N, dtype=torch.float32)
    gh = (2.0 * k / max(N - 1, 1) - 1.0) if N > 1 else torch.zeros(1)
    gh = gh.view(1, N, 1).expand(Zavg, N, Xo)
    grid = torch.stack([gw, gh], dim=-1)
    sampled = F.grid_sample(ra, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return sampled.mean(dim=2).numpy()


# --------------------------------------------------------------------------------------
# overhang fill — reference biahub/deskew.py:339-368 (torch variant, 3x3x3 cube dilation)
# --------------------------------------------------------------------------------------
def fill_overhang_oracle(vol, fill_value=None, dilation_iterations=3):
    """mask=(vol==0); dilate with a 3x3x3 max filter `iterations` times (borders never seed:
    max_pool3d pads with -inf); fill = mean(vol[~mask]) in fp32 if fill_value is None."""
    vol = np.asarray(vol, dtype=_f32)
    mask = vol == 0
    for _ in range(dilation_iterations):
        m = np.pad(mask, 1, mode="constant", constant_values=False)
        acc = np.zeros_like(mask)
        for dz in range(3):
            for dy in range(3):
                for dx in range(3):
                    acc |= m[dz:dz + mask.shape[0], dy:dy + mask.shape[1], dx:dx + mask.shape[2]]
        mask = acc
    if fill_value is None:
        import torch

        valid = torch.from_numpy(vol)[torch.from_numpy(~mask)]
        fill = _f32(valid.mean().item())  # fp32 reduction like the reference (torch)
    else:
        fill = _f32(fill_value)
    return np.where(mask, fill, vol).astype(_f32), mask


def deskew_oracle_points(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices, points):
    """Same per-voxel formula as ``deskew_oracle_numpy`` evaluated only at ``points`` — an
    (M, 3) int array of output indices (a, y, x).  For spot-checking full-size volumes."""
    raw = np.asarray(raw)
    Zi, Yi, Xi = raw.shape
    N = int(average_n_slices)
    s = deskew_scalars(raw.shape, ls_angle_deg, px_to_scan_ratio, keep_overhang)
    Zo = s["Zo"]
    pts = np.asarray(points, dtype=np.int64)
    a, y, x = pts[:, 0], pts[:, 1], pts[:, 2]
    acc = None
    for k in range(N):
        zo = a * N + k
        x32 = x.astype(_f32)
        zo32 = zo.astype(_f32)
        p = ((s["px32"] * x32) - (s["pxct32"] * zo32)).astype(_f32)
        p = (p + s["off32"]).astype(_f32)
        g = ((_f32(2.0) * p) / _f32(Zi - 1)) - _f32(1.0)
        pp = ((g + _f32(1.0)) / _f32(2.0)) * _f32(Zi - 1)
        f = np.floor(pp)
        w = (pp - f).astype(_f32)
        e = ((f + _f32(1.0)) - pp).astype(_f32)
        j0 = f.astype(np.int64)
        j1 = j0 + 1
        iy = Yi - 1 - np.minimum(zo, Zo - 1)
        ix = Xi - 1 - y
        t0 = np.where((j0 >= 0) & (j0 <= Zi - 1), raw[np.clip(j0, 0, Zi - 1), iy, ix], 0).astype(_f32)
        t1 = np.where((j1 >= 0) & (j1 <= Zi - 1), raw[np.clip(j1, 0, Zi - 1), iy, ix], 0).astype(_f32)
        sk = _fma32(t1, w, (t0 * e).astype(_f32))
        acc = sk if acc is None else (acc + sk).astype(_f32)
    return (acc / _f32(N)).astype(_f32)


# --------------------------------------------------------------------------------------
# legacy `deskew_zyx` order of operations — reference biahub/deskew.py:371-453
# --------------------------------------------------------------------------------------
def average_n_slices_torch_oracle(vol, window):
    """``_average_n_slices_torch`` (reference deskew.py:71-96): groups of ``window`` slices of the
    DESKEWED stack, the last slice repeated to fill the last group; fp32 mean."""
    vol = np.asarray(vol, dtype=_f32)
    w = int(window)
    if w == 1:
        return vol
    rem = vol.shape[0] % w
    if rem:
        vol = np.concatenate([vol, np.repeat(vol[-1:], w - rem, axis=0)], axis=0)
    grouped = vol.reshape((vol.shape[0] // w, w) + vol.shape[1:])
    acc = np.zeros(grouped.shape[:1] + grouped.shape[2:], dtype=_f32)
    for k in range(w):   # sequential fp32 sum, then true division (torch.mean over a short dim)
        acc = (acc + grouped[:, k]).astype(_f32)
    return (acc / _f32(w)).astype(_f32)


def fill_overhang_with_mean_oracle(vol, dilation_iterations=3):
    """``_fill_overhang_with_mean`` (reference deskew.py:277-336, the numpy variant the legacy
    path uses): zero mask, scipy ``binary_dilation`` with its default 3-D cross, fill with
    ``vol[~mask].mean()``."""
    from scipy.ndimage import binary_dilation

    vol = np.asarray(vol, dtype=_f32)
    mask = binary_dilation(vol == 0, iterations=dilation_iterations)
    out = vol.copy()
    out[mask] = vol[~mask].mean()
    return out, mask


def deskew_legacy_oracle(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, average_n_slices=1,
                         overhang_fill="zero"):
    """Legacy ``deskew_zyx``: deskew every tilt row, average the deskewed stack, numpy-variant
    fill.  The deskew itself is the production sampling (MONAI's rounding is unpinned)."""
    vol = deskew_oracle_numpy(raw, ls_angle_deg, px_to_scan_ratio, keep_overhang, 1)
    vol = average_n_slices_torch_oracle(vol, average_n_slices)
    if keep_overhang and overhang_fill == "mean":
        vol, _ = fill_overhang_with_mean_oracle(vol)
    return vol
