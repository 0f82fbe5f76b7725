"""GPU parity tests of the flat-field path (csrc/b2_flatfield.cu) through the C ABI: bit-identical
to the oracle (numpy restatement pinned to the unmodified reference) and to the committed golden
vectors; edge cases as the reference's own tests have them (tests/test_flat_field.py)."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import biahub_b200 as b2  # noqa: E402
from oracle import flatfield_oracle as fo  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_flatfield_v1.npz"))
NAMES = sorted({k.split("__")[0] for k in GOLD.files})


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def _cuda_u16(a):
    return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)


@pytest.mark.parametrize("name", NAMES)
def test_golden_vectors_host_and_device_api(name):
    data = GOLD[f"{name}__in"]
    assert _same(b2.flat_field_zyx(data), GOLD[f"{name}__zyx_f64"])
    czyx = np.stack([data, data[::-1].copy()])
    assert _same(b2._flat_field_czyx(czyx, target_indices=[0]), GOLD[f"{name}__czyx_f32"])
    dev = b2.flat_field_zyx(_cuda_u16(data))
    assert dev.is_cuda and _same(dev.cpu().numpy(), GOLD[f"{name}__zyx_f64"])


@pytest.mark.parametrize("shape", [(9, 5, 7), (10, 6, 8), (33, 17, 29), (64, 30, 16), (5, 1, 1),
                                   (2, 3, 1030), (101, 64, 64), (256, 9, 1024)])
@pytest.mark.parametrize("kind", ["noise", "camera", "sparse"])
def test_matches_oracle_bit_for_bit(shape, kind):
    rng = np.random.default_rng(hash((shape, kind)) & 0xffff)
    if kind == "noise":
        data = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    elif kind == "camera":  # dark counts + Poisson-ish signal: many ties
        data = (100 + rng.poisson(20, size=shape)).astype(np.uint16)
    else:                   # mostly zeros: medians of 0 -> inf / nan as numpy
        data = (rng.random(shape) < 0.4).astype(np.uint16) * rng.integers(0, 5, size=shape, dtype=np.uint16)
    with np.errstate(all="ignore"):
        want64 = fo.flat_field_zyx_oracle(data)
        want32 = fo.flat_field_czyx_oracle(data[None], [0])
    assert _same(b2.flat_field_zyx(data), want64)
    assert _same(b2._flat_field_czyx(data[None], [0]), want32)
    got32 = b2.flat_field._flatfield_tensor(_cuda_u16(data), torch.float32).cpu().numpy()
    assert _same(got32, want32[0])


@pytest.mark.parametrize("shape", [(2, 3, 5), (4, 7, 9), (6, 5, 64), (7, 3, 33), (40, 6, 130), (801, 2, 64),
                                   (800, 3, 66)])
def test_median_across_radix_bin_boundaries(shape):
    """The median kernel selects the high byte first and the low byte second; the upper middle
    sample of an even Z may sit in another low-byte bin or another high-byte bin.  Columns built
    to hit each of those cases, at odd and even plane sizes (scalar and 4-byte loads)."""
    rng = np.random.default_rng(sum(shape))
    Z, Y, X = shape
    data = np.empty(shape, np.uint16)
    flat = data.reshape(Z, -1)
    for p in range(flat.shape[1]):
        kind = p % 8
        if kind == 0:    # two values either side of a high-byte boundary, half and half
            base = int(rng.integers(1, 255)) << 8
            col = np.where(np.arange(Z) < (Z + 1) // 2, base - 1, base)
        elif kind == 1:  # lower half in one high-byte bin, upper half far above
            col = np.where(np.arange(Z) < Z // 2, rng.integers(0x0100, 0x01ff), rng.integers(0x7000, 0xffff))
        elif kind == 2:  # constant
            col = np.full(Z, rng.integers(0, 65536))
        elif kind == 3:  # all samples in one high-byte bin, many ties
            col = 0x1200 + rng.integers(0, 4, size=Z)
        elif kind == 4:  # full range
            col = rng.integers(0, 65536, size=Z)
        elif kind == 5:  # extremes only
            col = np.where(rng.random(Z) < 0.5, 0, 65535)
        elif kind == 6:  # distinct consecutive values straddling a boundary
            col = 0x3400 - Z // 2 + np.arange(Z)
        else:            # low byte 0xff / 0x00 neighbours
            col = np.where(rng.random(Z) < 0.5, 0x20ff, 0x2100 + rng.integers(0, 2, size=Z) * 0x100)
        flat[:, p] = rng.permutation(np.asarray(col, dtype=np.int64)).astype(np.uint16)
    with np.errstate(all="ignore"):
        want64 = fo.flat_field_zyx_oracle(data)
    assert _same(b2.flat_field_zyx(data), want64)
    # the medians themselves (the pattern is what the kernel under test produces)
    med = np.median(data, axis=0)
    with np.errstate(all="ignore"):
        want = data.astype(np.float64) / med * med.mean()
    assert _same(want, want64)


def test_reference_known_answer_and_passthrough():
    # reference tests/test_flat_field.py:76-90
    rng = np.random.default_rng(0)
    data = rng.integers(1, 1000, size=(12, 9, 11), dtype=np.uint16)
    expected = data / np.median(data, axis=0) * np.median(data, axis=0).mean()
    np.testing.assert_array_equal(b2.flat_field_zyx(data), expected)
    czyx = rng.integers(1, 1000, size=(3, 6, 5, 7), dtype=np.uint16)
    out = b2._flat_field_czyx(czyx, target_indices=[1])
    assert out.dtype == np.float32 and out.shape == czyx.shape
    np.testing.assert_array_equal(out[0], czyx[0].astype(np.float32))
    np.testing.assert_array_equal(out[2], czyx[2].astype(np.float32))
    np.testing.assert_array_equal(out[1], (czyx[1] / np.median(czyx[1], axis=0)
                                           * np.median(czyx[1], axis=0).mean()).astype(np.float32))


def test_other_axis_and_errors():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 4096, size=(6, 11, 8), dtype=np.uint16)
    expected = data / np.median(data, axis=1, keepdims=True) * np.median(data, axis=1).mean()
    np.testing.assert_array_equal(b2.flat_field_zyx(data, axis=1), expected)
    with pytest.raises(NotImplementedError):
        b2.flat_field_zyx(data.astype(np.float32))
    with pytest.warns(DeprecationWarning):
        b2.flat_field_correction(data)


def test_mantis_sized_properties():
    """(800, 300, 2048) uint16: checked through size-independent properties — the median pattern of
    the RESULT is flat (== mean of the input pattern, up to the float32 rounding of the output), a
    random sample of columns equals the oracle bit for bit."""
    Z, Y, X = 800, 300, 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    t = (torch.randint(90, 1200, (Z, Y, X), generator=g, device="cuda", dtype=torch.int32)
         + torch.arange(X, device="cuda", dtype=torch.int32) // 8).to(torch.uint16)
    out = b2.flat_field._flatfield_tensor(t, torch.float32)
    assert out.shape == (Z, Y, X)
    rng = np.random.default_rng(9)
    ys, xs = rng.integers(0, Y, 64), rng.integers(0, X, 64)
    ti = t.view(torch.int16)  # torch cannot index uint16 tensors; the bits are the same
    cols = ti[:, ys, xs].cpu().numpy().view(np.uint16)          # (Z, 64)
    srt = np.sort(cols.astype(np.int64), axis=0)
    pat_cols = (srt[Z // 2 - 1] + srt[Z // 2]) / 2.0
    # the full pattern sum (exact integer arithmetic, torch sort on the device, band by band)
    pat_sum2 = 0
    for y0 in range(0, Y, 50):
        band = ti[:, y0:y0 + 50].to(torch.int32) & 0xFFFF
        s = torch.sort(band, dim=0).values
        pat_sum2 += int((s[Z // 2 - 1] + s[Z // 2]).to(torch.int64).sum())
    mean = (pat_sum2 / 2.0) / (Y * X)
    want = (cols.astype(np.float64) / pat_cols[None, :] * mean).astype(np.float32)
    got = out[:, ys, xs].cpu().numpy()
    assert np.array_equal(got, want)


def test_flatfield_then_deskew_equals_the_two_steps():
    """Chained unit (flat-field output stays in HBM) == _fast_deskew_czyx(_flat_field_czyx(raw))."""
    rng = np.random.default_rng(21)
    raw = (100 + rng.poisson(40, size=(96, 30, 128))).astype(np.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    flat = b2._flat_field_czyx(raw[None], [0])
    want = b2._fast_deskew_czyx(flat, **kw)[0]
    got = b2.flatfield_then_deskew(raw, **kw)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    dev = b2.flatfield_then_deskew(_cuda_u16(raw), **kw)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)


def test_host_pipeline_many_bands_and_slabs():
    """b2h_flatfield_u16 with a volume that spans several upload bands and download slabs
    (48 MB each): pageable and pinned inputs, float32 and float64 results == the device API."""
    from biahub_b200._device import pinned_empty

    Z, Y, X = 96, 700, 1024   # 138 MB uint16 -> 3 bands; 275 MB float32 -> 6 slabs
    g = torch.Generator(device="cuda").manual_seed(12)
    t = torch.randint(80, 3000, (Z, Y, X), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    want32 = b2.flat_field._flatfield_tensor(t, torch.float32).cpu().numpy()
    host = t.view(torch.int16).cpu().numpy().view(np.uint16)
    got = b2._flat_field_czyx(host[None], [0])
    assert np.array_equal(got[0], want32)
    pinned = pinned_empty(host.shape, np.uint16)
    pinned[...] = host
    assert np.array_equal(b2._flat_field_czyx(pinned[None], [0])[0], want32)
    del got
    want64 = b2.flat_field._flatfield_tensor(t[:40], torch.float64).cpu().numpy()
    assert np.array_equal(b2.flat_field_zyx(host[:40]), want64)
