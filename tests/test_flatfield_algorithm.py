"""CPU model of the median kernel's ALGORITHM (csrc/b2_flatfield.cu, flatfield_median_kernel): the
two 8-bit radix sweeps, the branch-free two-level scan, the rank bookkeeping and the three ways the
upper middle sample of an even Z is found.  The model follows the kernel statement by statement
(same variable names) and is checked against np.median over many thousands of columns, including
every corner the selection logic has: ties, samples on high-byte / low-byte bin boundaries, the
upper middle sample in the same low-byte bin, in a later low-byte bin, and above the high-byte bin.
The GPU tests (tests/test_flatfield_gpu.py) check the kernel itself against the oracle."""

import numpy as np
import pytest


def _scan(hist, k):
    """fm_scan: bin where the cumulative count passes rank k -> (bin, rank within bin, count)."""
    assert hist.shape == (256,)
    gs = hist.reshape(16, 16).sum(axis=1)
    acc, rem, grp = 0, k, 0
    for g in range(16):
        acc += int(gs[g])
        d = (k - acc) & 0xFFFFFFFF          # unsigned wrap, as in the kernel
        rem = min(rem, d)
        grp += 1 if acc <= k else 0
    grp = min(grp, 15)
    acc2, rem2, b = 0, rem, 0
    for u in range(16):
        acc2 += int(hist[grp * 16 + u])
        d = (rem - acc2) & 0xFFFFFFFF
        rem2 = min(rem2, d)
        b += 1 if acc2 <= rem else 0
    b = min(b, 15)
    return grp * 16 + b, rem2, int(hist[grp * 16 + b])


def median2_model(col):
    """2 * median of a uint16 column, computed the way the kernel does."""
    col = np.asarray(col, dtype=np.int64)
    Z = col.size
    k1 = (Z - 1) >> 1
    even = (Z & 1) == 0
    # sweep 1: high bytes
    hist = np.bincount(col >> 8, minlength=256)
    hi, krem, here = _scan(hist, k1)
    pre = hi << 8
    outside = even and krem + 1 >= here
    # sweep 2: low bytes of the samples in the selected bin; the others go to the discard bin;
    # `above` = min(sample - (pre + 256)) in unsigned arithmetic (only tracked when some pixel of
    # the warp needs it: track = any(outside); the model tracks exactly when this pixel does)
    x = col ^ pre
    hist2 = np.bincount(x[x < 256], minlength=256)
    above = 0xFFFFFFFF
    if outside:
        above = int(((col - (pre + 256)) & 0xFFFFFFFF).min())
    lo, rank, here2 = _scan(hist2, krem)
    m1 = pre | lo
    m2 = m1
    if even and rank + 1 >= here2:
        b = lo + 1
        while b < 256 and hist2[b] == 0:
            b += 1
        m2 = (pre | b) if b < 256 else (above + pre + 256) & 0xFFFFFFFF
    return m1 + m2


def _check(col):
    want = 2.0 * float(np.median(np.asarray(col, dtype=np.uint16)))
    got = median2_model(col)
    assert got == want, (list(map(int, col)), got, want)


def test_scan_matches_a_plain_cumulative_search():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n_occupied = int(rng.integers(1, 40))
        hist = np.zeros(256, np.int64)
        hist[rng.choice(256, size=n_occupied, replace=False)] = rng.integers(1, 900, size=n_occupied)
        total = int(hist.sum())
        for k in {0, total - 1, total // 2, int(rng.integers(0, total))}:
            cum = np.cumsum(hist)
            want_bin = int(np.searchsorted(cum, k, side="right"))
            below = int(cum[want_bin - 1]) if want_bin else 0
            assert _scan(hist, k) == (want_bin, k - below, int(hist[want_bin]))


@pytest.mark.parametrize("Z", list(range(1, 13)) + [31, 32, 33, 64, 255, 256, 257, 800, 801])
def test_random_columns(Z):
    rng = np.random.default_rng(Z)
    for rep in range(120):
        kind = rep % 6
        if kind == 0:
            col = rng.integers(0, 65536, size=Z)
        elif kind == 1:      # camera-like: narrow range, many ties
            col = 100 + rng.poisson(20, size=Z)
        elif kind == 2:      # two clusters in different high-byte bins
            col = np.where(rng.random(Z) < 0.5, rng.integers(0x0100, 0x0200, size=Z),
                           rng.integers(0x4000, 0xFFFF, size=Z))
        elif kind == 3:      # around one high-byte boundary
            base = int(rng.integers(1, 255)) << 8
            col = base + rng.integers(-3, 4, size=Z)
        elif kind == 4:      # few distinct values
            col = rng.choice(rng.integers(0, 65536, size=3), size=Z)
        else:                # extremes
            col = rng.choice([0, 255, 256, 0xFF00, 0xFFFF], size=Z)
        _check(col)


def test_every_way_the_upper_middle_sample_is_found():
    # same value (rank stays inside the low-byte bin)
    _check([7, 7, 7, 7])
    # next occupied low-byte bin of the same high-byte bin
    _check([0x1203, 0x1203, 0x1280, 0x12FF])
    # the lower middle sample is the last of its high-byte bin: smallest sample above the bin
    _check([0x12FF, 0x1200, 0x4567, 0x1300])
    _check([0x00FF, 0x0100])
    _check([0, 65535])
    _check([65535, 65535])
    _check([0, 0])
    # halves either side of a boundary, odd and even
    for Z in (2, 3, 10, 11, 800):
        col = np.where(np.arange(Z) < (Z + 1) // 2, 0x33FF, 0x3400)
        _check(col)
        _check(col[::-1])


def test_exhaustive_small_alphabet():
    """All columns of length 1..6 over an alphabet that straddles both kinds of bin boundary."""
    import itertools

    alphabet = [0x00FE, 0x00FF, 0x0100, 0x0101, 0xFFFF]
    for Z in range(1, 7):
        for col in itertools.product(alphabet, repeat=Z):
            _check(col)
