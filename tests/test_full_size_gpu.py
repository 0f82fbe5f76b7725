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


def test_c2_variants_full_size_keep_overhang_and_production_ratio():
    """configs[1] variants: keep_overhang=True (X_out 2333) and the production pixel ratio 0.755
    (nextflow/configs/*/deskew.yml: 0.1133 / 0.150 → (100,2048,800)); TMA kernel against the gather
    kernel bit for bit and against the sampled-voxel oracle."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    rng = np.random.default_rng(1001)
    raw = rng.integers(0, 65536, size=(800, 300, 2048), dtype=np.uint16)
    t = _to_cuda(raw)
    for px, keep, shape in ((0.386, True, (100, 2048, 2333)), (0.755, False, (100, 2048, 800))):
        a = b2.fast_deskew_zyx(t, 30.0, px, keep, 3, _path=_cabi.PATH_TMA)
        assert tuple(a.shape) == shape == b2.get_deskewed_data_shape(raw.shape, 30.0, px, keep, 3)[0]
        g = b2.fast_deskew_zyx(t, 30.0, px, keep, 3, _path=_cabi.PATH_GATHER)
        assert torch.equal(a, g)
        del g
        a_h = a.cpu().numpy()
        pts = _sample_points(shape, 40000, 11)
        want = do.deskew_oracle_points(raw, 30.0, px, keep, 3, pts)
        assert np.abs(a_h[pts[:, 0], pts[:, 1], pts[:, 2]] - want).max() <= 2e-7 * 65535.0
        del a, a_h
    # keep_overhang + overhang_fill through the pipelined host path at full size: no zero is left
    # and the un-masked voxels are untouched
    filled = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                                  keep_overhang=True, average_n_slices=3, overhang_fill="mean")[0]
    plain = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                                 keep_overhang=True, average_n_slices=3)[0]
    assert filled.shape == (100, 2048, 2333) and (filled == 0).sum() == 0
    changed = filled != plain
    assert changed.any() and np.unique(filled[changed]).size == 1      # one fill value
    assert abs(float(filled[changed][0]) - float(plain[~changed].mean(dtype=np.float64))) <= 1e-2


def test_c5_chained_unit_full_size():
    """configs[4] unit at full size: deskew uint16 (800,300,2048) N=3 → register the float32
    (100,2048,1813) result with the C3-style matrix; the one pipelined host call
    (b2h_deskew_affine3d) equals the device chain bit for bit, and sampled voxels match the oracle
    chain (oracle deskew at the taps' source voxels → oracle warp)."""
    import torch

    import biahub_b200 as b2

    rng = np.random.default_rng(4000)
    raw = rng.integers(0, 65536, size=(800, 300, 2048), dtype=np.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    mid_shape = (100, 2048, 1813)
    M = ao.register_matrix_c3(mid_shape)
    host = b2.deskew_then_register(raw, M, mid_shape, **kw)
    dev = b2.deskew_then_register(_to_cuda(raw), M, mid_shape, **kw)
    assert tuple(dev.shape) == mid_shape
    assert np.array_equal(dev.cpu().numpy(), host)
    # two-step product path (what the two CLI commands do): the dense 1813-float rows are not
    # 16-byte aligned, so the warp runs on the gather kernel instead of the TMA kernel the
    # pitched chain uses — same arithmetic, float64 vs tile-local fp32 coordinates
    mid = b2.fast_deskew_zyx(_to_cuda(raw), 30.0, 0.386, False, 3)
    two = b2.affine_warp(mid, M, mid_shape, order=1, boundary="itk")
    assert float((two - dev).abs().max()) <= 1e-4 * 65535.0
    del two, mid
    # oracle chain on sampled output voxels: the 8 taps of each sample come from the oracle deskew
    pts = _sample_points(mid_shape, 20000, 13)
    A, tr = M[:3, :3], M[:3, 3]
    c = pts @ A.T + tr
    base = np.floor(c).astype(np.int64)
    need = set()
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                q = base + np.array([dz, dy, dx])
                ok = np.all((q >= 0) & (q < np.array(mid_shape)), axis=1)
                need.update(map(tuple, q[ok]))
    need = np.array(sorted(need))
    vals = do.deskew_oracle_points(raw, 30.0, 0.386, False, 3, need)
    sparse = np.zeros(mid_shape, dtype=np.float32)          # only the needed taps are filled in
    sparse[need[:, 0], need[:, 1], need[:, 2]] = vals
    want = ao.affine_oracle_points(sparse, M, pts, 1, "itk")
    got = host[pts[:, 0], pts[:, 1], pts[:, 2]]
    assert np.abs(got.astype(np.float64) - want).max() <= 1e-4 * 65535.0
