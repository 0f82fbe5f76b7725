"""CPU tests: the flat-field oracle against the committed golden vectors (outputs of the
unmodified reference, tests/golden/make_golden_flatfield.py), the reference's own known-answer
test, and — when /root/reference is mounted — the live reference."""

import os
import warnings

import numpy as np
import pytest

from oracle import flatfield_oracle as fo
from oracle.ref_loader import reference_available

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_flatfield_v1.npz"))
NAMES = sorted({k.split("__")[0] for k in GOLD.files})


def _same(a, b):
    """bit-identical including NaN / inf positions"""
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(name):
    data = GOLD[f"{name}__in"]
    with np.errstate(all="ignore"):
        assert _same(fo.flat_field_zyx_oracle(data), GOLD[f"{name}__zyx_f64"])
        czyx = np.stack([data, data[::-1].copy()])
        assert _same(fo.flat_field_czyx_oracle(czyx, [0]), GOLD[f"{name}__czyx_f32"])


def test_reference_known_answer():
    # reference tests/test_flat_field.py:76-81
    rng = np.random.default_rng(0)
    data = rng.integers(1, 1000, size=(12, 9, 11), dtype=np.uint16)
    expected = data / np.median(data, axis=0) * np.median(data, axis=0).mean()
    np.testing.assert_array_equal(fo.flat_field_zyx_oracle(data), expected)


def test_mean_of_pattern_is_order_independent():
    """Pattern values are multiples of 0.5 below 65536: every partial sum is exact in float64, so
    numpy's pairwise mean equals the sequential / integer one (what the CUDA kernel computes)."""
    rng = np.random.default_rng(1)
    pattern = rng.integers(0, 131071, size=(300, 2048)).astype(np.float64) / 2.0
    exact = int((pattern * 2).astype(np.int64).sum())
    assert pattern.mean() == (exact / 2.0) / pattern.size
    assert pattern.sum() == exact / 2.0


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference():
    from oracle.ref_loader import load_reference_flat_field

    ff = load_reference_flat_field()
    rng = np.random.default_rng(7)
    for shape in [(33, 7, 5), (40, 3, 9)]:
        data = rng.integers(0, 65536, size=shape, dtype=np.uint16)
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            assert _same(fo.flat_field_zyx_oracle(data), ff.flat_field_zyx(data))
