"""GPU parity tests against outputs of the reference's OWN functions (golden_affine_v1.npz,
golden_legacy_v1.npz; generators next to them) — everything through the C ABI.

Tolerances: order 0 / integer shifts bit-exact; order 1 and the cubic spline ≤ 1e-4 of the input
dynamic range (north_star); integer outputs of the scipy branch ≤ 1 count (a rounding flip at k+0.5).
"""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

pytestmark = pytest.mark.gpu


def _slicing(case):
    return None if case["crop"] is None else tuple(slice(a, b) for a, b in case["crop"])


def _kw(case):
    return {k: case[k] for k in ("method", "interpolation") if k in case}


def _range(a):
    a = np.asarray(a, dtype=np.float64)
    a = a[np.isfinite(a)]
    return max(float(a.max() - a.min()), 1.0) if a.size else 1.0


def test_apply_affine_transform_matches_reference_goldens(golden_affine):
    import biahub_b200 as b2

    arrays, meta = golden_affine
    for case in meta["apply"]:
        n = case["name"]
        vol, want = arrays[f"apply_{n}_in"], arrays[f"apply_{n}_out"]
        with np.errstate(over="ignore"):
            got = b2.apply_affine_transform(vol, arrays[f"apply_{n}_M"], tuple(case["output_shape"]),
                                            crop_output_slicing=_slicing(case), **_kw(case))
        assert got.shape == want.shape, n
        assert got.dtype == want.dtype, n
        if case.get("method", "ants") == "scipy":
            if want.dtype.kind in "ui":
                d = np.abs(got.astype(np.int64) - want.astype(np.int64))
                assert d.max() <= 1 and (d != 0).mean() < 5e-3, n
            else:
                assert np.abs(got.astype(np.float64) - want).max() <= 1e-4 * _range(vol), n
        elif case.get("interpolation", "linear") == "nearestneighbor":
            assert np.array_equal(got, want), n
        else:
            finite = np.isfinite(want)
            assert np.array_equal(finite, np.isfinite(got)), n
            # taps scrubbed to +-FLT_MAX dominate their neighbourhood: relative term for those
            tol = 1e-4 * _range(vol) + 1e-6 * np.abs(want[finite].astype(np.float64))
            assert (np.abs(got[finite].astype(np.float64) - want[finite]) <= tol).all(), n
            if n in ("u8_identity", "ref_kat_translation"):
                assert np.array_equal(got, want), n


def test_apply_stabilization_matches_reference_goldens(golden_affine):
    import biahub_b200 as b2

    arrays, meta = golden_affine
    for case in meta["stabilize"]:
        n = case["name"]
        vol, want = arrays[f"stab_{n}_in"], arrays[f"stab_{n}_out"]
        shape = None if case["output_shape"] is None else tuple(case["output_shape"])
        got = b2.apply_stabilization_transform(vol, list(arrays[f"stab_{n}_mats"]), case["t"], shape)
        assert got.shape == want.shape and got.dtype == np.float32, n
        if n == "int_shift_3d":
            assert np.array_equal(got, want)
        else:
            tol = (1e-4 * _range(np.nan_to_num(vol.astype(np.float64), posinf=0, neginf=0))
                   + 1e-6 * np.abs(want.astype(np.float64)))
            assert (np.abs(got.astype(np.float64) - want) <= tol).all(), n


def test_find_overlapping_volume_matches_reference_goldens(golden_affine):
    import biahub_b200 as b2

    arrays, meta = golden_affine
    for case in meta["overlap"]:
        sl = b2.find_overlapping_volume(tuple(case["input_shape"]), tuple(case["target_shape"]),
                                        arrays[f"overlap_{case['name']}_M"])
        assert [[s.start, s.stop] for s in sl] == case["slices"], case["name"]


@pytest.mark.parametrize("dtype", ["float32", "uint16"])
def test_spline_warp_matches_scipy_at_size(dtype):
    """The reference's literal call on a volume that spans several host slabs, host and tensor API."""
    import scipy.ndimage
    import torch

    import biahub_b200 as b2

    rng = np.random.default_rng(9)
    shape = (24, 300, 420)
    vol = (rng.integers(0, 65536, size=shape, dtype=np.uint16) if dtype == "uint16"
           else (rng.random(shape, dtype=np.float32) * 4095).astype(np.float32))
    M = ao.register_matrix_c3(shape)
    M[0, 1], M[1, 0], M[0, 2] = 0.02, -0.015, 0.01
    want = scipy.ndimage.affine_transform(vol, M, shape)      # reference register.py:272
    got = b2.apply_affine_transform(vol, M, shape, method="scipy")
    assert got.dtype == want.dtype and got.shape == want.shape
    if dtype == "uint16":
        d = np.abs(got.astype(np.int64) - want.astype(np.int64))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
        t = torch.from_numpy(vol.view(np.int16)).cuda().view(torch.uint16)
    else:
        assert np.abs(got - want).max() <= 2e-6 * 4095
        t = torch.from_numpy(vol).cuda()
    crop = (slice(3, 20), slice(10, 280), slice(7, 401))
    dev = b2.spline_warp(t, M, crop).cpu()
    dev = dev.view(torch.int16).numpy().view(np.uint16) if dtype == "uint16" else dev.numpy()
    assert np.array_equal(dev, got[crop])


def test_spline_nonfinite_scrub_and_degenerate_axes():
    import scipy.ndimage

    import biahub_b200 as b2

    rng = np.random.default_rng(10)
    vol = (rng.random((1, 40, 37), dtype=np.float32) * 100).astype(np.float32)
    vol[0, 5, 5] = np.nan
    M = np.eye(4)
    M[1, 3], M[2, 3], M[1, 2] = 0.37, -1.6, 0.05
    want = scipy.ndimage.affine_transform(np.nan_to_num(vol, nan=0), M, vol.shape)
    got = b2.apply_affine_transform(vol, M, vol.shape, method="scipy")
    assert np.abs(got - want).max() <= 1e-5 * 100
    one = np.full((1, 1, 1), 7.0, dtype=np.float32)
    assert b2.apply_affine_transform(one, np.eye(4), (1, 1, 1), method="scipy")[0, 0, 0] == 7.0


def test_legacy_deskew_zyx_matches_reference_stage_goldens():
    """Legacy order of operations (average AFTER deskew with the last slice repeated; numpy-variant
    fill with scipy's cross): golden_legacy_v1.npz holds the reference's own
    `_average_n_slices_torch` / `_fill_overhang_with_mean` outputs."""
    import os

    import biahub_b200 as b2

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_legacy_v1.npz"))
    for name in ("keep_n3_mean", "keep_n2_zero", "crop_n4", "keep_n1_mean"):
        th, px, keep, n, mean = g[f"{name}_params"]
        raw, want = g[f"{name}_in"], g[f"{name}_out"]
        got = b2.deskew_zyx(raw, float(th), float(px), bool(keep), average_n_slices=int(n),
                            overhang_fill="mean" if mean else "zero")
        assert got.shape == want.shape and got.dtype == np.float32, name
        assert np.abs(got - want).max() <= 1e-5 * 65535, name
        if not mean:
            assert np.abs(got - want).max() <= 2e-7 * 65535, name
        # the legacy semantics differ from the production path exactly where documented
        if int(n) > 1 and raw.shape[1] % int(n):
            fast = b2._fast_deskew_czyx(raw[None], ls_angle_deg=float(th), px_to_scan_ratio=float(px),
                                        keep_overhang=bool(keep), average_n_slices=int(n))[0]
            assert np.array_equal(fast[:-1], got[:-1]) or mean
            assert not np.array_equal(fast[-1], got[-1]) or mean


def test_average_n_slices_torch_matches_reference_golden():
    import os

    import torch

    import biahub_b200 as b2

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_legacy_v1.npz"))
    for name in ("keep_n3_mean", "keep_n2_zero", "crop_n4"):
        n = int(g[f"{name}_params"][3])
        got = b2._average_n_slices_torch(torch.from_numpy(g[f"{name}_deskewed"]).cuda(), n).cpu().numpy()
        assert np.array_equal(got, g[f"{name}_avg"]), name
    t = torch.arange(24, dtype=torch.float32).reshape(4, 3, 2).cuda()
    assert b2._average_n_slices_torch(t, 1) is t
    with pytest.raises(RuntimeError, match="GPU only"):
        b2._average_n_slices_torch(torch.zeros(4, 2), 2)


def test_fill_pipeline_many_slabs_equals_tensor_path_and_oracle():
    """keep_overhang + overhang_fill through b2h_deskew_fill (uploads / deskew slabs overlap, fill
    on the resident volume, slab downloads) on a volume of several slabs."""
    import torch

    import biahub_b200 as b2

    raw = np.random.default_rng(31).integers(0, 65536, size=(400, 45, 512), dtype=np.uint16)
    t = torch.from_numpy(raw.view(np.int16)).cuda().view(torch.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=True, average_n_slices=3)
    for fill in ("mean", 123.0):
        host = b2._fast_deskew_czyx(raw[None], overhang_fill=fill, **kw)[0]
        dev = b2.fast_deskew_zyx(t, 30.0, 0.386, True, 3, overhang_fill=fill).cpu().numpy()
        assert host.shape == dev.shape
        # the mean is an atomically accumulated float64 sum: equal up to its last float32 bit
        assert np.abs(host - dev).max() <= 1e-6 * 65535
        assert (host == 0).sum() == 0
    # pageable destination of the same call
    out = np.empty(host.shape, dtype=np.float32)
    b2._fast_deskew_czyx(raw[None], overhang_fill=123.0, out=out, **kw)
    assert np.array_equal(out, host)


def test_num_splits_is_honoured_and_exact():
    import biahub_b200 as b2

    raw = np.random.default_rng(12).integers(0, 65536, size=(160, 24, 250), dtype=np.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    full = b2._fast_deskew_czyx(raw[None], **kw)[0]
    for k in (2, 3, 7):
        assert np.array_equal(b2._fast_deskew_czyx(raw[None], num_splits=k, **kw)[0], full), k
    # with a fill every chunk is filled on its own, as the reference does (deskew.py:563-573)
    kwf = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=True, average_n_slices=3,
               overhang_fill="mean")
    split = b2._fast_deskew_czyx(raw[None], num_splits=2, **kwf)[0]
    parts = [b2._fast_deskew_czyx(np.ascontiguousarray(c)[None], **kwf)[0]
             for c in reversed(np.array_split(raw, 2, axis=2))]
    assert np.array_equal(split, np.concatenate(parts, axis=1))


def test_host_calls_leave_the_current_device_alone():
    """b2h_* select their GPU internally and must restore the caller's device."""
    import torch

    import biahub_b200 as b2

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    vol = np.random.default_rng(1).random((8, 64, 64), dtype=np.float32)
    M = ao.register_matrix_c3(vol.shape)
    got = b2.affine_warp(vol, M, vol.shape, device=1)
    assert torch.cuda.current_device() == 0
    raw = np.random.default_rng(2).integers(0, 65536, size=(64, 12, 64), dtype=np.uint16)
    b2._fast_deskew_czyx(raw[None], device="cuda:1", ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                         keep_overhang=False, average_n_slices=3)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(got, b2.affine_warp(vol, M, vol.shape, device=0))


def test_float64_overflow_corner_is_pinned():
    """float64 inputs holding ±inf or |v| > FLT_MAX: scrubbed in float64, cast to
    float32 (→ ±inf), and scrubbed once more at the tap → ±FLT_MAX.  The golden is the reference
    wrapper run with the scrubbing ITK-rule resampler."""
    import biahub_b200 as b2

    vol = np.zeros((4, 5, 6))
    vol[1, 2, 3] = np.inf
    vol[2, 2, 3] = -1e39
    vol[3, 3, 3] = np.nan
    with np.errstate(over="ignore"):
        out = b2.apply_affine_transform(vol, np.eye(4), vol.shape, interpolation="nearestneighbor")
        want = ao.apply_affine_transform_oracle(vol, np.eye(4), vol.shape, interpolation="nearestneighbor")
    fmax = np.finfo(np.float32).max
    assert out[1, 2, 3] == fmax and out[2, 2, 3] == -fmax and out[3, 3, 3] == 0
    assert np.array_equal(out, want)
