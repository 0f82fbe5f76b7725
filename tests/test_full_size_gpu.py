"""Full BASELINE.json sizes: size-independent properties + sparse oracle checks (the dense CPU
oracle would take minutes at these sizes)."""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu


def _to_cuda(arr):
    import torch

    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(arr).cuda()


def _sample_points(shape, n, seed):
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.integers(0, s, size=n) for s in shape], axis=1)
    corners = np.array([[(shape[0] - 1) * a, (shape[1] - 1) * b, (shape[2] - 1) * c]
                        for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    return np.concatenate([pts, corners])


def test_c2_mantis_deskew_full_size():
    """configs[1]: uint16 (800,300,2048), theta 30, px 0.386, N=3, crop → (100,2048,1813)."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    rng = np.random.default_rng(1000)
    raw = rng.integers(0, 65536, size=(800, 300, 2048), dtype=np.uint16)
    t = _to_cuda(raw)
    a = b2.fast_deskew_zyx(t, 30.0, 0.386, False, 3, _path=_cabi.PATH_TMA)
    assert tuple(a.shape) == (100, 2048, 1813)
    g = b2.fast_deskew_zyx(t, 30.0, 0.386, False, 3, _path=_cabi.PATH_GATHER)
    assert torch.equal(a, g)                       # two independent kernels, bit for bit
    del g
    host = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                                keep_overhang=False, average_n_slices=3)[0]
    a_h = a.cpu().numpy()
    assert np.array_equal(host, a_h)               # slab pipeline == single launch
    pts = _sample_points(a_h.shape, 40000, 7)
    want = do.deskew_oracle_points(raw, 30.0, 0.386, False, 3, pts)
    got = a_h[pts[:, 0], pts[:, 1], pts[:, 2]]
    assert np.abs(got - want).max() <= 2e-7 * 65535.0
    # split invariance on a column band: rows of the output depend only on their own
    # input columns
    band = b2.fast_deskew_zyx(_to_cuda(np.ascontiguousarray(raw[:, :, 1024:1280])), 30.0, 0.386,
                              False, 3)
    assert torch.equal(band, a[:, 2048 - 1280:2048 - 1024, :])


def test_c3_register_full_size():
    """configs[2]: float32 (120,2048,2048), rotate 7.3 deg + scale 1.07 + shift, order 1 and 0."""
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (120, 2048, 2048)
    rng = np.random.default_rng(2000)
    vol = rng.random(shape, dtype=np.float32)
    vol *= np.float32(4095.0)
    nan_idx = rng.integers(0, vol.size, size=vol.size // 1000)
    vol.ravel()[nan_idx] = np.nan                 # 0.1 % NaNs exercise the scrub
    M = ao.register_matrix_c3(shape)
    t = _to_cuda(vol)
    pts = _sample_points(shape, 60000, 3)
    for boundary in ("itk", "constant"):
        got = affine_warp(t, M, shape, order=1, boundary=boundary, _path=_cabi.PATH_TMA)
        torch.cuda.synchronize()
        assert torch.isfinite(got).all()
        sel = got[torch.from_numpy(pts[:, 0]).cuda(), torch.from_numpy(pts[:, 1]).cuda(),
                  torch.from_numpy(pts[:, 2]).cuda()].cpu().numpy()
        want = ao.affine_oracle_points(vol, M, pts, 1, boundary)
        assert np.abs(sel.astype(np.float64) - want).max() <= 1e-4 * 4095.0
        del got
    got0 = affine_warp(t, M, shape, order=0, boundary="constant", _path=_cabi.PATH_TMA)
    sel0 = got0[torch.from_numpy(pts[:, 0]).cuda(), torch.from_numpy(pts[:, 1]).cuda(),
                torch.from_numpy(pts[:, 2]).cuda()].cpu().numpy()
    assert np.array_equal(sel0, ao.affine_oracle_points(vol, M, pts, 0, "constant"))


def test_c4_stabilize_full_size_integer_and_fractional():
    """configs[3]: float32 (64,2048,2048); integer shifts are bit-exact shifted copies, fractional
    shifts match the float64 oracle on a sample."""
    import torch

    from biahub_b200 import apply_stabilization_transform

    shape = (64, 2048, 2048)
    rng = np.random.default_rng(3000)
    vol = rng.random(shape, dtype=np.float32)
    vol *= np.float32(4095.0)
    mats = [np.eye(4) for _ in range(3)]
    mats[1][:3, 3] = (2, -5, 7)
    mats[2][:3, 3] = (-1.25, 3.5, -0.75)
    out = apply_stabilization_transform(vol, mats, 1)
    want = np.zeros_like(vol)
    want[0:62, 5:2048, 0:2041] = vol[2:64, 0:2043, 7:2048]
    assert np.array_equal(out, want)
    del want
    out = apply_stabilization_transform(vol, mats, 2)
    pts = _sample_points(shape, 60000, 5)
    ref = ao.affine_oracle_points(vol, mats[2], pts, 1, "itk")
    assert np.abs(out[pts[:, 0], pts[:, 1], pts[:, 2]].astype(np.float64) - ref).max() <= 1e-4 * 4095.0
    # device API returns the same bits as the host pipeline
    dev = apply_stabilization_transform(torch.from_numpy(vol).cuda(), mats, 2).cpu().numpy()
    assert np.array_equal(dev, out)
