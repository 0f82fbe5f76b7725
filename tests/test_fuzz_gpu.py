"""Randomised parity sweep: random shapes / parameters / matrices against the CPU oracle.
Seeds are fixed so failures reproduce; sizes keep the dense oracle within seconds."""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_deskew_random_cases(seed):
    import biahub_b200 as b2

    rng = np.random.default_rng(1000 + seed)
    Zi = int(rng.integers(2, 140))
    Yi = int(rng.integers(1, 40))
    Xi = int(rng.choice([int(rng.integers(1, 300)), 64, 128, 72, 136, 256, 8 * int(rng.integers(8, 40))]))
    theta = float(np.round(rng.uniform(5.0, 45.0), 2))
    px = float(np.round(rng.uniform(0.15, 1.6), 3))
    n = int(rng.integers(1, 7))
    keep = bool(rng.integers(0, 2))
    dtype = [np.uint16, np.float32][int(rng.integers(0, 2))]
    if dtype == np.uint16:
        raw = rng.integers(0, 65536, size=(Zi, Yi, Xi), dtype=np.uint16)
        span = 65535.0
    else:
        raw = (rng.random((Zi, Yi, Xi), dtype=np.float32) * 4095).astype(np.float32)
        span = 4095.0
    try:
        want = do.deskew_oracle_numpy(raw, theta, px, keep, n)
    except ValueError:
        with pytest.raises(ValueError, match="only overhang"):
            b2._fast_deskew_czyx(raw[None], ls_angle_deg=theta, px_to_scan_ratio=px,
                                 keep_overhang=keep, average_n_slices=n)
        return
    got = b2._fast_deskew_czyx(raw[None], ls_angle_deg=theta, px_to_scan_ratio=px,
                               keep_overhang=keep, average_n_slices=n)[0]
    assert got.shape == want.shape, (Zi, Yi, Xi, theta, px, n, keep)
    err = np.abs(got.astype(np.float64) - want).max() / span
    assert err <= 2e-7, (seed, (Zi, Yi, Xi), theta, px, n, keep, dtype.__name__, err)


def _random_matrix(rng, shape, kind):
    M = np.eye(4)
    if kind == "translation":
        M[:3, 3] = rng.uniform(-6, 6, 3)
    elif kind == "int_translation":
        M[:3, 3] = rng.integers(-5, 6, 3)
    elif kind == "zsep":
        th = np.radians(rng.uniform(-25, 25))
        s = rng.uniform(0.7, 1.4)
        M[1:3, 1:3] = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        M[0, 0] = rng.uniform(0.5, 2.0)
        c = np.array(shape) / 2
        M[:3, 3] = c - M[:3, :3] @ c + rng.uniform(-4, 4, 3)
    else:  # generic
        A = np.eye(3) + rng.uniform(-0.15, 0.15, (3, 3))
        if rng.random() < 0.3:
            A[0] *= -1
        c = np.array(shape) / 2
        M[:3, :3] = A
        M[:3, 3] = c - A @ c + rng.uniform(-3, 3, 3)
    return M


@pytest.mark.parametrize("seed", range(32))
def test_affine_random_cases(seed):
    from biahub_b200 import affine_warp

    rng = np.random.default_rng(2000 + seed)
    kind = ["translation", "int_translation", "zsep", "generic"][seed % 4]
    shape = (int(rng.integers(1, 24)), int(rng.integers(1, 90)),
             int(rng.choice([int(rng.integers(1, 150)), 64, 96, 128, 4 * int(rng.integers(2, 40))])))
    out_shape = tuple(int(max(1, v + rng.integers(-3, 4))) for v in shape)
    order = int(rng.integers(0, 2))
    boundary = ["constant", "itk"][int(rng.integers(0, 2))]
    if rng.random() < 0.5:
        vol = rng.integers(0, 65536, size=shape, dtype=np.uint16)
        span = 65535.0
    else:
        vol = (rng.random(shape, dtype=np.float32) * 4095).astype(np.float32)
        if vol.size > 50:
            vol.ravel()[rng.integers(0, vol.size, 3)] = np.nan
        span = 4095.0
    M = _random_matrix(rng, shape, kind)
    want = ao.affine_oracle_numpy(vol, M, out_shape, order, boundary)
    got = affine_warp(vol, M, out_shape, order=order, boundary=boundary)
    assert got.shape == want.shape and got.dtype == np.float32
    if order == 0 or kind == "int_translation":
        assert np.array_equal(got, want), (seed, kind, shape, out_shape, order, boundary,
                                           int((got != want).sum()))
    else:
        err = np.abs(got.astype(np.float64) - want).max() / span
        assert err <= 1e-4, (seed, kind, shape, out_shape, order, boundary, err)
