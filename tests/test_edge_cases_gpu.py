"""Edge cases of the reference-facing calls on the GPU: ragged/minimal shapes, dtypes, empty
crops, pageable vs pinned host buffers, error behaviour."""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
def test_deskew_input_dtypes(dtype):
    """Every numeric dtype is legal (reference casts to float32, deskew.py:578)."""
    import biahub_b200 as b2

    rng = np.random.default_rng(1)
    raw = (rng.random((48, 10, 72)) * 200).astype(dtype)
    want = do.deskew_oracle_numpy(raw, 30.0, 0.386, True, 2)
    got = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                               keep_overhang=True, average_n_slices=2)[0]
    assert np.abs(got - want).max() <= 2e-7 * 200


@pytest.mark.parametrize("shape,n", [((2, 1, 8), 1), ((2, 3, 4), 1), ((3, 2, 64), 5), ((5, 7, 1), 3),
                                     ((64, 5, 65), 7), ((33, 4, 136), 3)])
def test_deskew_minimal_and_ragged_shapes(shape, n):
    """Z=2 (the minimum the reference can divide by), single rows/columns, N > Y, N > 4,
    widths just off the TMA tile sizes."""
    import biahub_b200 as b2

    rng = np.random.default_rng(2)
    raw = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    want = do.deskew_oracle_numpy(raw, 36.0, 0.386, True, n)
    got = b2._fast_deskew_czyx(raw[None], ls_angle_deg=36.0, px_to_scan_ratio=0.386,
                               keep_overhang=True, average_n_slices=n)[0]
    assert got.shape == want.shape
    # N >= 5: torch's reduction order differs from the sequential sum by 1 ulp at most
    assert np.abs(got - want).max() <= 2e-7 * 65535.0


def test_deskew_rejects_single_plane_and_overhang_only():
    import biahub_b200 as b2

    with pytest.raises(ValueError):
        b2._fast_deskew_czyx(np.zeros((1, 1, 4, 8), np.uint16), ls_angle_deg=30.0,
                             px_to_scan_ratio=0.386, keep_overhang=True)
    with pytest.raises(ValueError, match="Dataset contains only overhang"):
        b2._fast_deskew_czyx(np.zeros((1, 10, 500, 100), np.uint16), ls_angle_deg=30.0,
                             px_to_scan_ratio=0.1, keep_overhang=False)


def test_pinned_and_pageable_host_buffers_agree():
    import biahub_b200 as b2
    from biahub_b200._device import pinned_empty

    rng = np.random.default_rng(3)
    raw = rng.integers(0, 65536, size=(200, 40, 256), dtype=np.uint16)   # several slabs
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    pageable = b2._fast_deskew_czyx(raw[None], **kw)[0]
    pin_in = pinned_empty(raw.shape, np.uint16)
    pin_in[...] = raw
    pin_out = pinned_empty(pageable.shape, np.float32)
    res = b2._fast_deskew_czyx(pin_in[None], out=pin_out, **kw)
    assert res.base is not None and np.shares_memory(res, pin_out)
    assert np.array_equal(pin_out, pageable)
    vol = rng.random((40, 300, 256), dtype=np.float32)
    M = ao.register_matrix_c3(vol.shape)
    a = b2.affine_warp(vol, M, vol.shape)
    pin_v = pinned_empty(vol.shape, np.float32)
    pin_v[...] = vol
    pin_o = pinned_empty(vol.shape, np.float32)
    b2.affine_warp(pin_v, M, vol.shape, out=pin_o)
    assert np.array_equal(a, pin_o)
    with pytest.raises(ValueError):
        b2.affine_warp(vol, M, vol.shape, out=np.empty((1, 2, 3), np.float32))


def test_affine_empty_and_degenerate_outputs():
    from biahub_b200 import affine_warp, apply_affine_transform

    vol = np.random.default_rng(4).random((6, 20, 24), dtype=np.float32)
    out = apply_affine_transform(vol, np.eye(4), (6, 20, 24),
                                 crop_output_slicing=(slice(2, 2), slice(0, 20), slice(0, 24)))
    assert out.shape == (0, 20, 24) and out.dtype == np.float32
    far = np.eye(4)
    far[:3, 3] = (1000, 0, 0)                       # everything maps outside → zeros
    assert not affine_warp(vol, far, vol.shape).any()
    one = affine_warp(vol[:1, :1, :8], np.eye(4), (1, 1, 8))   # single row volume
    assert np.array_equal(one, vol[:1, :1, :8])
    up = np.diag([0.5, 0.5, 0.5, 1.0])              # 2x upsampling onto a bigger grid
    big = affine_warp(vol, up, (12, 40, 48), order=1, boundary="constant")
    assert np.abs(big - ao.affine_oracle_numpy(vol, up, (12, 40, 48), 1, "constant")).max() <= 1e-4


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.float64])
def test_affine_input_dtypes(dtype):
    from biahub_b200 import apply_affine_transform

    rng = np.random.default_rng(5)
    vol = (rng.random((8, 40, 72)) * 250).astype(dtype)
    M = ao.register_matrix_c3(vol.shape)
    want = ao.affine_oracle_numpy(vol, M, vol.shape, 1, "itk")
    got = apply_affine_transform(vol, M, vol.shape)
    assert np.abs(got - want).max() <= 1e-4 * 250


def test_nan_scrub_semantics():
    """np.nan_to_num(nan=0): NaN -> 0, +-inf -> +-float32 max, before interpolation."""
    from biahub_b200 import affine_warp

    vol = np.ones((4, 16, 64), np.float32)
    vol[1, 5, 9] = np.nan
    vol[2, 6, 20] = np.inf
    vol[2, 7, 30] = -np.inf
    out = affine_warp(vol, np.eye(4), vol.shape)            # integer shift (0): shifted-copy path
    assert out[1, 5, 9] == 0 and out[2, 6, 20] == np.finfo(np.float32).max
    assert out[2, 7, 30] == -np.finfo(np.float32).max
    M = np.eye(4)
    M[2, 3] = 0.5
    out = affine_warp(vol, M, vol.shape)                     # fractional: NaN neighbours blend as 0
    assert np.isfinite(out).all()
    assert out[1, 5, 8] == 0.5 and out[1, 5, 9] == 0.5
    want = ao.affine_oracle_numpy(vol, M, vol.shape, 1, "itk")
    fin = np.abs(want) < 1e30
    assert np.abs(out[fin] - want[fin]).max() <= 1e-6
