"""GPU parity tests for the deskew path (through the C ABI: python wrappers -> ctypes -> CUDA)."""
import numpy as np
import pytest

from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu

RANGE_U16 = 65535.0
TOL = 1e-4  # north_star: max abs error <= 1e-4 of the input dynamic range (order 1)


def _rand(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.uint16:
        return rng.integers(0, 65536, size=shape, dtype=np.uint16)
    return (rng.random(shape, dtype=np.float32) * np.float32(4095.0)).astype(dtype)


def _to_cuda(arr):
    import torch

    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(arr).cuda()


def _check(got, want, raw, name=""):
    rng = float(raw.max()) - float(raw.min())
    assert got.shape == want.shape, name
    assert got.dtype == np.float32
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= TOL * rng, f"{name}: max err {err} vs range {rng}"
    return err / rng, float((got == want).mean())


def test_golden_cases_host_api(golden):
    """Committed golden vectors (unmodified reference output) through `_fast_deskew_czyx`."""
    import biahub_b200 as b2

    arrays, meta = golden
    for case in meta["deskew"]:
        raw = arrays[f"deskew_{case['name']}_in"]
        want = arrays[f"deskew_{case['name']}_out"]
        got = b2._fast_deskew_czyx(
            raw[None], device="cuda", num_splits=case["num_splits"],
            ls_angle_deg=case["ls_angle_deg"], px_to_scan_ratio=case["px_to_scan_ratio"],
            keep_overhang=case["keep_overhang"], average_n_slices=case["average_n_slices"],
            overhang_fill=case["overhang_fill"])
        assert got.shape == (1,) + want.shape and got.dtype == np.float32
        rel, same = _check(got[0], want, raw, case["name"])
        if case["overhang_fill"] == 0 and case["average_n_slices"] <= 4:
            assert rel <= 2e-7, (case["name"], rel, same)


@pytest.mark.parametrize("path", ["gather", "tma"])
@pytest.mark.parametrize("case", [
    # shape, dtype, theta, px, keep_overhang, N
    ((128, 30, 128), np.uint16, 30.0, 0.386, False, 3),
    ((128, 31, 128), np.uint16, 30.0, 0.386, True, 3),      # padded last group (31 % 3 != 0)
    ((96, 16, 72), np.uint16, 30.0, 0.386, False, 1),        # partial y tile (72 % 64 != 0)
    ((64, 9, 200), np.uint16, 45.0, 0.755, True, 2),
    ((80, 10, 64), np.uint16, 20.0, 1.25, True, 4),          # px > 1
    ((120, 12, 96), np.float32, 30.0, 0.386, False, 3),
    ((40, 7, 36), np.float32, 36.0, 0.5, True, 1),           # partial y tile for f32 (36 % 32)
    ((16, 4, 64), np.uint16, 30.0, 0.386, True, 2),          # brick deeper than the volume
])
def test_device_api_matches_oracle(case, path):
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    shape, dtype, theta, px, keep, n = case
    raw = _rand(shape, dtype, seed=hash((shape, n)) % 2**31)
    want = do.deskew_oracle_numpy(raw, theta, px, keep, n)
    p = _cabi.PATH_TMA if path == "tma" else _cabi.PATH_GATHER
    got = b2.fast_deskew_zyx(_to_cuda(raw), theta, px, keep, n, _path=p)
    torch.cuda.synchronize()
    assert got.dtype == torch.float32 and got.is_cuda
    rel, same = _check(got.cpu().numpy(), want, raw, f"{case}/{path}")
    assert rel <= 2e-7, (case, path, rel, same)   # oracle-level agreement (bit-identical class)


def test_tma_and_gather_bit_identical_c1():
    """C1 size (BASELINE.json configs[0]): two independent kernels must agree bit for bit, and a
    sample of voxels must match the per-voxel oracle formula."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    raw = _rand((256, 256, 512), np.uint16, seed=0)
    t = _to_cuda(raw)
    for keep, n in ((False, 1), (False, 3), (True, 3)):
        a = b2.fast_deskew_zyx(t, 30.0, 0.386, keep, n, _path=_cabi.PATH_TMA)
        b = b2.fast_deskew_zyx(t, 30.0, 0.386, keep, n, _path=_cabi.PATH_GATHER)
        assert torch.equal(a, b)
        shape, _ = b2.get_deskewed_data_shape(raw.shape, 30.0, 0.386, keep, n)
        assert tuple(a.shape) == tuple(shape)
        rng = np.random.default_rng(5)
        pts = np.stack([rng.integers(0, s, size=20000) for s in a.shape], axis=1)
        # include the corners/edges of the volume
        edge = np.array([[0, 0, 0], [a.shape[0] - 1, a.shape[1] - 1, a.shape[2] - 1],
                         [0, a.shape[1] - 1, 0], [a.shape[0] - 1, 0, a.shape[2] - 1]])
        pts = np.concatenate([pts, edge])
        want = do.deskew_oracle_points(raw, 30.0, 0.386, keep, n, pts)
        got = a.cpu().numpy()[pts[:, 0], pts[:, 1], pts[:, 2]]
        assert np.abs(got - want).max() <= 2e-7 * RANGE_U16


def test_split_invariance_and_host_equals_device():
    """deskewing x-chunks in reverse order and concatenating along output Y is exact;
    the host pipeline (b2h_deskew) returns exactly what the device API computes."""
    import biahub_b200 as b2

    raw = _rand((160, 24, 256), np.uint16, seed=11)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    full = b2._fast_deskew_czyx(raw[None], **kw)[0]
    parts = [b2._fast_deskew_czyx(np.ascontiguousarray(c)[None], **kw)[0]
             for c in reversed(np.array_split(raw, 2, axis=2))]
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    dev = b2.fast_deskew_zyx(_to_cuda(raw), 30.0, 0.386, False, 3).cpu().numpy()
    assert np.array_equal(dev, full)
    # num_splits is accepted and changes nothing
    assert np.array_equal(b2._fast_deskew_czyx(raw[None], num_splits=3, **kw)[0], full)


def test_constant_volume_and_zero_padding():
    """A constant volume deskews to the constant wherever both taps are inside, 0 in the far
    overhang, and values in between only on the one-voxel blend border."""
    import biahub_b200 as b2

    raw = np.full((64, 16, 64), 1000, dtype=np.uint16)
    out = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                               keep_overhang=True, average_n_slices=1)[0]
    assert out.max() <= 1000.0 + 1e-3 and out.min() >= 0.0
    assert (out == 1000.0).sum() > 0.3 * out.size and (out == 0).sum() > 0
    crop = b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                                keep_overhang=False, average_n_slices=1)[0]
    assert np.abs(crop[:, :, 2:-2] - 1000.0).max() <= 1e-3


def test_legacy_entry_points_and_errors():
    import biahub_b200 as b2

    # reference tests/test_cli/test_deskew_cli.py:33-59
    raw = np.random.default_rng(3).random((2, 3, 4))
    out = b2.deskew_zyx(raw, 36, 0.386, True, average_n_slices=1)
    assert out.shape[1] == 4
    assert out[0, 0, 0] != 0
    assert out.shape == b2.get_deskewed_data_shape(raw.shape, 36, 0.386, True, pixel_size_um=1.0)[0]
    assert b2._deskew_czyx(raw[None], ls_angle_deg=36, px_to_scan_ratio=0.386,
                           keep_overhang=True).shape == (1,) + out.shape
    # reference tests/test_cli/test_deskew_cli.py:189-204
    data = np.random.default_rng(4).random((10, 500, 100))
    with pytest.raises(ValueError, match="Dataset contains only overhang"):
        b2.deskew_zyx(data, 30, 0.1, keep_overhang=False)
    assert b2.deskew_zyx(data, 30, 0.1, keep_overhang=True).shape[2] > 0


def test_fill_matches_oracle_on_tma_shape():
    import biahub_b200 as b2

    raw = _rand((96, 12, 64), np.uint16, seed=21)
    base = do.deskew_oracle_numpy(raw, 30.0, 0.386, True, 3)
    for fill in ("mean", 250.0):
        want, _ = do.fill_overhang_oracle(base, None if fill == "mean" else fill)
        got = b2.fast_deskew_zyx(_to_cuda(raw), 30.0, 0.386, True, 3, overhang_fill=fill).cpu().numpy()
        assert np.abs(got - want).max() <= 1e-5 * RANGE_U16


def test_library_is_the_compute_path():
    """The extension must be the thing that ran: launches are counted inside the library."""
    import biahub_b200 as b2
    from biahub_b200 import _cabi

    before = _cabi.launch_count()
    raw = _rand((32, 6, 64), np.uint16, seed=1)
    b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                         keep_overhang=False, average_n_slices=3)
    assert _cabi.launch_count() > before


@pytest.mark.parametrize("dtype,X", [("uint16", 132), ("uint16", 203), ("float32", 130), ("float32", 67)])
@pytest.mark.parametrize("n", [1, 3, 4])
def test_unaligned_rows_take_the_brick_kernel_with_manual_fill(dtype, X, n):
    """Source rows that are not 16-byte aligned cannot be addressed by TMA: the register kernel
    then fills its brick with coalesced element loads (same swizzled layout, same arithmetic).
    Bit-identical to the plain gather kernel and to the oracle; PATH_TMA still refuses."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi

    rng = np.random.default_rng(91)
    shape = (150, 26, X)
    raw = (rng.integers(0, 65536, size=shape, dtype=np.uint16) if dtype == "uint16"
           else (rng.random(shape, dtype=np.float32) * 4095).astype(np.float32))
    assert (X * raw.itemsize) % 16 != 0
    t = _to_cuda(raw)
    for keep in (False, True):
        auto = b2.fast_deskew_zyx(t, 30.0, 0.386, keep, n)
        gather = b2.fast_deskew_zyx(t, 30.0, 0.386, keep, n, _path=_cabi.PATH_GATHER)
        assert torch.equal(auto, gather)
        want = do.deskew_oracle_numpy(raw, 30.0, 0.386, keep, n)
        rngv = 65535.0 if dtype == "uint16" else 4095.0
        assert np.abs(auto.cpu().numpy() - want).max() <= 2e-7 * rngv
    with pytest.raises(_cabi.B2Unsupported):
        b2.fast_deskew_zyx(t, 30.0, 0.386, False, n, _path=_cabi.PATH_TMA)
