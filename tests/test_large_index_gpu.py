"""Volumes beyond 2^31 voxels (the 'maximum sizes' edge): 64-bit indexing of every kernel on the
path, checked at sampled output voxels against the point oracles (the dense CPU oracle would take
hours at these sizes)."""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu


def _points(shape, n, seed):
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.integers(0, s, size=n) for s in shape], axis=1)
    hi = np.stack([rng.integers(max(0, s - 3), s, size=64) for s in shape], axis=1)  # last planes / rows
    corners = np.array([[(shape[0] - 1) * a, (shape[1] - 1) * b, (shape[2] - 1) * c]
                        for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    return np.concatenate([pts, hi, corners])


def test_deskew_beyond_int32_voxels():
    """uint16 (1400, 400, 4096) = 2.29e9 source voxels -> float32 (400, 4096, 3281) = 5.4e9 output
    voxels (> 2^32): TMA kernel == gather kernel, both == the oracle at sampled voxels."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    shape = (1400, 400, 4096)
    g = torch.Generator(device="cuda").manual_seed(77)
    t = torch.randint(0, 65536, shape, generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    out = b2.fast_deskew_zyx(t, 30.0, 0.386, False, 1, _path=_cabi.PATH_TMA)
    assert out.numel() > 2 ** 32 and tuple(out.shape) == (400, 4096, 3281)
    pts = _points(out.shape, 30000, 3)
    raw = t.view(torch.int16).cpu().numpy().view(np.uint16)
    want = do.deskew_oracle_points(raw, 30.0, 0.386, False, 1, pts)
    idx = torch.from_numpy(pts).cuda()
    got = out[idx[:, 0], idx[:, 1], idx[:, 2]].cpu().numpy()
    assert np.abs(got - want).max() <= 2e-7 * 65535.0
    ref = b2.fast_deskew_zyx(t, 30.0, 0.386, False, 1, _path=_cabi.PATH_GATHER)
    # compare in z slabs to keep the temporary small
    for a0 in range(0, out.shape[0], 50):
        assert torch.equal(out[a0:a0 + 50], ref[a0:a0 + 50])


def test_affine_beyond_int32_voxels():
    """float32 (640, 2048, 2048) = 2.68e9 voxels, C3-style matrix, order 1 and 0: the z-marching
    kernel at sampled voxels == the oracle; the last planes / rows are in the sample."""
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (640, 2048, 2048)
    g = torch.Generator(device="cuda").manual_seed(78)
    t = torch.rand(shape, generator=g, device="cuda") * 4095.0
    M = ao.register_matrix_c3(shape)
    host = t.cpu().numpy()
    pts = _points(shape, 20000, 4)
    idx = torch.from_numpy(pts).cuda()
    for order in (1, 0):
        out = affine_warp(t, M, shape, order=order, boundary="itk", _path=_cabi.PATH_TMA)
        assert out.numel() > 2 ** 31
        want = ao.affine_oracle_points(host, M, pts, order, "itk")
        got = out[idx[:, 0], idx[:, 1], idx[:, 2]].cpu().numpy()
        if order == 0:
            assert np.array_equal(got, want)
        else:
            assert np.abs(got - want).max() <= 1e-4 * 4095.0
        del out


def test_flatfield_beyond_int32_voxels():
    """uint16 (1200, 1024, 2048) = 2.5e9 voxels: sampled columns == numpy, bit for bit."""
    import torch

    import biahub_b200 as b2

    Z, Y, X = 1200, 1024, 2048
    g = torch.Generator(device="cuda").manual_seed(79)
    t = torch.randint(90, 900, (Z, Y, X), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    out = b2.flat_field._flatfield_tensor(t, torch.float32)
    assert out.numel() > 2 ** 31
    ti = t.view(torch.int16)
    pat_sum2 = 0
    for y0 in range(0, Y, 32):
        s = torch.sort(ti[:, y0:y0 + 32].to(torch.int32) & 0xFFFF, dim=0).values
        pat_sum2 += int((s[Z // 2 - 1] + s[Z // 2]).to(torch.int64).sum())
    mean = (pat_sum2 / 2.0) / (Y * X)
    rng = np.random.default_rng(5)
    ys = np.concatenate([rng.integers(0, Y, 60), [0, Y - 1, Y - 1, 0]])
    xs = np.concatenate([rng.integers(0, X, 60), [0, X - 1, 0, X - 1]])
    cols = ti[:, ys, xs].cpu().numpy().view(np.uint16)
    srt = np.sort(cols.astype(np.int64), axis=0)
    pat = (srt[Z // 2 - 1] + srt[Z // 2]) / 2.0
    want = (cols.astype(np.float64) / pat[None, :] * mean).astype(np.float32)
    assert np.array_equal(out[:, ys, xs].cpu().numpy(), want)
