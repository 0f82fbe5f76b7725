"""CPU tests: pin oracle/ against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle import deskew_oracle as do
from oracle.ref_loader import reference_available


def _deskew_cases(meta):
    return [c for c in meta["deskew"]]


def test_deskew_numpy_oracle_matches_reference_golden(golden):
    arrays, meta = golden
    for case in _deskew_cases(meta):
        if case["overhang_fill"] != 0:
            continue
        raw = arrays[f"deskew_{case['name']}_in"]
        want = arrays[f"deskew_{case['name']}_out"]
        got = do.deskew_oracle_numpy(raw, case["ls_angle_deg"], case["px_to_scan_ratio"],
                                     case["keep_overhang"], case["average_n_slices"])
        assert got.shape == want.shape, case["name"]
        rng = float(raw.max()) - float(raw.min())
        # bit-identical on the hosts used so far; the gate is 2e-7 of range (1 ulp class)
        assert np.abs(got - want).max() <= 2e-7 * rng, case["name"]
        assert (got == want).mean() > 0.999, case["name"]


def test_deskew_torch_oracle_matches_reference_golden(golden):
    arrays, meta = golden
    for case in _deskew_cases(meta):
        if case["overhang_fill"] != 0:
            continue
        raw = arrays[f"deskew_{case['name']}_in"]
        want = arrays[f"deskew_{case['name']}_out"]
        got = do.deskew_oracle_torch(raw, case["ls_angle_deg"], case["px_to_scan_ratio"],
                                     case["keep_overhang"], case["average_n_slices"])
        rng = float(raw.max()) - float(raw.min())
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-7 * rng, case["name"]


def test_fill_oracle_matches_reference_golden(golden):
    arrays, meta = golden
    for case in _deskew_cases(meta):
        if case["overhang_fill"] == 0:
            continue
        raw = arrays[f"deskew_{case['name']}_in"]
        want = arrays[f"deskew_{case['name']}_out"]
        base = do.deskew_oracle_numpy(raw, case["ls_angle_deg"], case["px_to_scan_ratio"],
                                      case["keep_overhang"], case["average_n_slices"])
        fill = None if case["overhang_fill"] == "mean" else case["overhang_fill"]
        got, mask = do.fill_overhang_oracle(base, fill)
        assert mask.any() and not mask.all()
        assert np.abs(got - want).max() <= 1e-6 * 65535, case["name"]


def test_points_oracle_equals_dense_oracle(golden):
    arrays, meta = golden
    case = meta["deskew"][0]
    raw = arrays[f"deskew_{case['name']}_in"]
    dense = do.deskew_oracle_numpy(raw, case["ls_angle_deg"], case["px_to_scan_ratio"],
                                   case["keep_overhang"], case["average_n_slices"])
    pts = np.argwhere(np.ones(dense.shape, bool))[::7]
    sparse = do.deskew_oracle_points(raw, case["ls_angle_deg"], case["px_to_scan_ratio"],
                                     case["keep_overhang"], case["average_n_slices"], pts)
    assert np.array_equal(sparse, dense[pts[:, 0], pts[:, 1], pts[:, 2]])


def test_shape_logic_matches_reference_golden(golden):
    _, meta = golden
    for c in meta["shapes"]:
        shape, voxel = do.deskewed_shape_oracle(tuple(c["raw_shape"]), c["ls_angle_deg"],
                                                c["px_to_scan_ratio"], c["keep_overhang"],
                                                c["average_n_slices"], c["pixel_size_um"])
        assert list(shape) == c["out_shape"]
        assert np.allclose(voxel, c["voxel_size"], rtol=0, atol=0)


def test_average_n_slices_known_answer(golden):
    # reference tests/test_cli/test_deskew_cli.py:11-30
    arrays, _ = golden
    data = np.arange(1, 17).reshape(4, 2, 2)
    assert np.array_equal(do.average_n_slices_oracle(data, 3),
                          np.array([[[5, 6], [7, 8]], [[13, 14], [15, 16]]]))
    assert np.array_equal(do.average_n_slices_oracle(data, 2),
                          np.array([[[3, 4], [5, 6]], [[11, 12], [13, 14]]]))
    assert np.array_equal(do.average_n_slices_oracle(data, 1), data)
    for w in (1, 2, 3):
        assert np.array_equal(do.average_n_slices_oracle(data, w), arrays[f"avg_w{w}"])
        assert do.average_n_slices_oracle(data, w).shape == do.averaged_shape_oracle(data.shape, w)


def test_overhang_only_error():
    # reference tests/test_cli/test_deskew_cli.py:189-204
    with pytest.raises(ValueError, match="Dataset contains only overhang"):
        do.deskewed_shape_oracle((10, 500, 100), 30, 0.1, keep_overhang=False)
    shape, _ = do.deskewed_shape_oracle((10, 500, 100), 30, 0.1, keep_overhang=True)
    assert shape[2] > 0


def test_affine_numpy_oracle_matches_scipy_golden(golden):
    arrays, meta = golden
    vol = arrays["affine_in"]
    rng = 4095.0
    for c in meta["affine"]:
        want = arrays[f"affine_{c['name']}_o{c['order']}"]
        got = ao.affine_oracle_numpy(vol, np.array(c["matrix"]), tuple(c["out_shape"]), c["order"],
                                     "constant")
        fin = np.isfinite(want) & np.isfinite(got)
        assert fin.mean() > 0.99
        if c["order"] == 0:
            assert np.array_equal(got[fin], want[fin]), c["name"]
        else:
            big = np.abs(want) > 1e30   # scrubbed +-inf taps: compare relatively
            assert np.abs(got[fin & ~big] - want[fin & ~big]).max() <= 1e-6 * rng, c["name"]
        # and the live scipy on this host agrees with the stored fixture
        live = ao.affine_oracle_scipy(vol, np.array(c["matrix"]), tuple(c["out_shape"]), c["order"])
        assert np.array_equal(live[fin], want[fin]), c["name"]


def test_affine_translation_known_answer():
    # reference tests/test_affine.py:43-59 — pins the pull convention and the sign
    ones = np.ones((10, 10, 10))
    M = np.eye(4)
    M[:3, 3] = (-3, 1, 4)
    for boundary in ("constant", "itk"):
        for order in (0, 1):
            out = ao.affine_oracle_numpy(ones, M, (10, 10, 10), order, boundary)
            assert out.shape == (10, 10, 10) and out.dtype == np.float32
            assert np.all(out[3:10, 0:9, 0:6] == 1)
            assert out.sum() == 7 * 9 * 6
    out = ao.affine_oracle_scipy(ones, M, (10, 10, 10), 1)
    assert np.all(out[3:10, 0:9, 0:6] == 1)


def test_affine_itk_band_semantics():
    # half-voxel band: clamp-to-edge inside [-0.5, n-0.5), zero outside; constant mode is strict
    vol = np.arange(1, 9, dtype=np.float32).reshape(2, 2, 2)
    M = np.eye(4)
    M[2, 3] = -0.25   # x coordinate = x - 0.25 → -0.25 for x = 0
    itk = ao.affine_oracle_numpy(vol, M, (2, 2, 2), 1, "itk")
    con = ao.affine_oracle_numpy(vol, M, (2, 2, 2), 1, "constant")
    assert np.array_equal(itk[:, :, 0], vol[:, :, 0])     # clamped to the edge voxel
    assert np.all(con[:, :, 0] == 0)                      # strictly outside → 0
    assert np.allclose(itk[:, :, 1], 0.25 * vol[:, :, 0] + 0.75 * vol[:, :, 1])
    pts = np.argwhere(np.ones((2, 2, 2), bool))
    assert np.array_equal(ao.affine_oracle_points(vol, M, pts, 1, "itk"), itk.ravel())
    assert np.array_equal(ao.affine_oracle_points(vol, M, pts, 1, "constant"), con.ravel())


@pytest.mark.skipif(not reference_available(), reason="reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    from oracle.ref_loader import load_reference_deskew

    ref = load_reference_deskew()
    rng = np.random.default_rng(42)
    raw = rng.integers(0, 65536, size=(72, 19, 24), dtype=np.uint16)
    for keep, n in ((False, 3), (True, 2), (False, 1)):
        want = ref._fast_deskew_czyx(raw[None], device="cpu", ls_angle_deg=30.0,
                                     px_to_scan_ratio=0.386, keep_overhang=keep,
                                     average_n_slices=n)[0]
        got = do.deskew_oracle_numpy(raw, 30.0, 0.386, keep, n)
        assert np.abs(got - want).max() <= 2e-7 * 65535
