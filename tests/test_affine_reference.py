"""CPU tests: the affine half of the path pinned against the reference's OWN functions.

``tests/golden/golden_affine_v1.npz`` holds outputs of the unmodified reference
``apply_affine_transform`` / ``apply_stabilization_transform`` / matrix builders /
``convert_transform_*`` / ``rescale_voxel_size`` / ``find_lir`` / ``find_overlapping_volume``
(imported from /root/reference by oracle/ref_loader.py; ``method="scipy"`` with the real scipy,
``method="ants"`` with the fake ants whose resampler is the ITK-rule oracle).  Here:

* the oracle's wrapper restatements must reproduce them (so the GPU tests can use the oracle at
  other sizes),
* the live reference, when mounted, must still reproduce them,
* the product's host-side functions (no GPU needed) must reproduce them.
"""
import contextlib
import io

import numpy as np
import pytest

from oracle import affine_oracle as ao
from oracle.ref_loader import reference_available


def _slicing(case):
    return None if case["crop"] is None else tuple(slice(a, b) for a, b in case["crop"])


def _kw(case):
    kw = {}
    for k in ("method", "interpolation"):
        if k in case:
            kw[k] = case[k]
    return kw


def _close(got, want, name):
    assert got.shape == want.shape, name
    assert got.dtype == want.dtype, name
    if want.dtype.kind in "ui":
        # identical up to a rounding flip at k+0.5 (float64 spline sums differ in the last bits)
        d = np.abs(got.astype(np.int64) - want.astype(np.int64))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3, name
    else:
        scale = max(float(np.abs(want[np.isfinite(want)]).max()), 1.0)
        assert np.allclose(got, want, rtol=0, atol=1e-6 * scale, equal_nan=True), name


def test_oracle_wrapper_matches_reference_apply(golden_affine):
    arrays, meta = golden_affine
    for case in meta["apply"]:
        n = case["name"]
        got = ao.apply_affine_transform_oracle(arrays[f"apply_{n}_in"], arrays[f"apply_{n}_M"],
                                               tuple(case["output_shape"]),
                                               crop_output_slicing=_slicing(case), **_kw(case))
        want = arrays[f"apply_{n}_out"]
        assert str(want.dtype) == case["out_dtype"]
        if case.get("method", "ants") == "ants":
            assert np.array_equal(got, want), n   # same resampler behind the wrapper → identical
        else:
            _close(got, want, n)


def test_oracle_wrapper_matches_reference_stabilize(golden_affine):
    arrays, meta = golden_affine
    for case in meta["stabilize"]:
        n = case["name"]
        shape = None if case["output_shape"] is None else tuple(case["output_shape"])
        got = ao.apply_stabilization_oracle(arrays[f"stab_{n}_in"], list(arrays[f"stab_{n}_mats"]),
                                            case["t"], shape)
        assert np.array_equal(got, arrays[f"stab_{n}_out"]), n
        assert got.dtype == np.float32


def test_reference_kat_inside_golden(golden_affine):
    arrays, _ = golden_affine
    out = arrays["apply_ref_kat_translation_out"]    # reference tests/test_affine.py:43-59
    assert out.shape == (10, 10, 10) and np.all(out[3:10, 0:9, 0:6] == 1)


def test_scipy_branch_quirks_are_what_the_reference_does(golden_affine):
    """register.py:272 hands output_shape_zyx to scipy's `offset` slot: the scipy branch returns
    the INPUT's shape and the INPUT's dtype."""
    arrays, meta = golden_affine
    by = {c["name"]: c for c in meta["apply"]}
    assert by["scipy_f32_generic"]["output_shape"] == [12, 20, 44]
    assert arrays["apply_scipy_f32_generic_out"].shape == arrays["apply_scipy_f32_generic_in"].shape
    assert arrays["apply_scipy_u16_out"].dtype == np.uint16
    assert arrays["apply_scipy_f64_nan_out"].dtype == np.float64
    assert arrays["apply_scipy_u16_4d_crop_out"].dtype == np.float32


def test_spline_prefilter_restatement_matches_scipy():
    import scipy.ndimage

    rng = np.random.default_rng(5)
    for shape in [(7, 9, 11), (1, 9, 13), (2, 3, 50), (40, 2, 3)]:
        vol = rng.random(shape) * 1000
        want = scipy.ndimage.spline_filter(vol, 3, output=np.float64, mode="constant")
        got = ao.spline3_prefilter_numpy(vol)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), shape


def test_spline_oracle_matches_scipy_call():
    import scipy.ndimage

    rng = np.random.default_rng(6)
    vol = (rng.random((6, 17, 21)) * 4095).astype(np.float32)
    M = ao.register_matrix_c3(vol.shape)
    M[0, 1], M[1, 0], M[0, 3] = 0.03, -0.02, 0.7
    want = scipy.ndimage.affine_transform(vol, M, vol.shape)   # the reference's literal call
    got = ao.affine_oracle_spline3(vol, M)
    assert got.dtype == np.float32 and np.abs(got - want).max() <= 1e-6 * 4095
    u = rng.integers(0, 65536, size=(5, 12, 14), dtype=np.uint16)
    want = scipy.ndimage.affine_transform(u, M, u.shape)
    got = ao.affine_oracle_spline3(u, M)
    assert got.dtype == np.uint16 and np.abs(got.astype(int) - want.astype(int)).max() <= 1


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_live_reference_still_produces_the_goldens(golden_affine):
    from oracle.ref_loader import load_reference_register_stabilize

    reg, stab = load_reference_register_stabilize()
    arrays, meta = golden_affine
    for case in meta["apply"]:
        n = case["name"]
        with np.errstate(over="ignore"):
            got = reg.apply_affine_transform(arrays[f"apply_{n}_in"], arrays[f"apply_{n}_M"],
                                             tuple(case["output_shape"]),
                                             crop_output_slicing=_slicing(case), **_kw(case))
        assert np.array_equal(got, arrays[f"apply_{n}_out"], equal_nan=True), n
    for case in meta["stabilize"]:
        n = case["name"]
        shape = None if case["output_shape"] is None else tuple(case["output_shape"])
        with contextlib.redirect_stdout(io.StringIO()):
            got = stab.apply_stabilization_transform(arrays[f"stab_{n}_in"], list(arrays[f"stab_{n}_mats"]),
                                                     case["t"], shape)
        assert np.array_equal(got, arrays[f"stab_{n}_out"]), n
    for case in meta["lir"]:
        sl = reg.find_lir(arrays[f"lir_{case['name']}_mask"])
        assert [[s.start, s.stop] for s in sl] == case["slices"]


# ---- product host logic against the reference-produced answers (no GPU involved) -----------
def test_product_matrix_builders_match_reference(golden_affine):
    import biahub_b200.register as reg

    arrays, meta = golden_affine
    for case in meta["matrices"]:
        args = [tuple(a) if isinstance(a, list) else a for a in case["args"]]
        got = getattr(reg, case["fn"])(*args)
        assert np.array_equal(np.asarray(got, dtype=np.float64), arrays[f"matrix_{case['name']}"]), case["name"]


def test_product_transform_packing_matches_reference(golden_affine):
    import biahub_b200.register as reg

    arrays, _ = golden_affine
    T = reg.convert_transform_to_ants(arrays["convert_generic_M"])
    assert np.array_equal(np.asarray(T.parameters), arrays["convert_generic_params"])
    assert np.array_equal(reg.convert_transform_to_numpy(T), arrays["convert_generic_back"])
    T.set_fixed_parameters(arrays["convert_generic_fixed"])
    assert np.array_equal(reg.convert_transform_to_numpy(T), arrays["convert_generic_back_fixed"])


def test_product_rescale_voxel_size_matches_reference(golden_affine):
    import biahub_b200.register as reg

    arrays, meta = golden_affine
    for k in meta["voxel"]:
        got = reg.rescale_voxel_size(arrays[f"voxel_{k}_M"], arrays[f"voxel_{k}_scale"])
        assert np.array_equal(got, arrays[f"voxel_{k}_out"])


def test_product_find_lir_matches_reference(golden_affine):
    import biahub_b200.register as reg

    arrays, meta = golden_affine
    for case in meta["lir"]:
        sl = reg.find_lir(arrays[f"lir_{case['name']}_mask"])
        assert [[s.start, s.stop] for s in sl] == case["slices"], case["name"]
