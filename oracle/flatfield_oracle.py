"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's flat-field correction.

Reference: ``biahub/flat_field.py``
  * ``flat_field_zyx`` (:105-122): ``pattern = median(zyx, axis)`` (``_median_tiled`` is
    ``np.median`` over cache-sized tiles, identical output, :57-102), result
    ``zyx / pattern * pattern.mean()`` — float64 for integer input.
  * ``_flat_field_czyx`` (:152-166): the callable handed to ``process_single_position``; target
    channels are corrected and CAST TO FLOAT32 on assignment, the others are ``astype(float32)``.

Restated per element so that the CUDA kernels can be checked bit for bit:
  * median of Z samples: sort, take the mean of the two middle elements (``np.median`` for even Z
    computes ``mean([a, b])`` = ``(a + b) / 2`` in float64; for uint16 data this is exact);
  * mean of the pattern: the pattern values are multiples of 0.5 below 65536, so ANY summation
    order gives the same exactly representable float64 sum; ``mean = sum / count`` (one rounding);
  * element: ``float64(v) / pattern`` (correctly rounded), ``* mean`` (correctly rounded), then
    for the czyx adapter one more rounding to float32.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
Parity status: PINNED — ``tests/golden/golden_flatfield_v1.npz`` holds outputs of the unmodified
reference (``tests/golden/make_golden_flatfield.py``) and the reference's own known-answer test
(``tests/test_flat_field.py:76-81``) is reproduced in ``tests/test_flatfield_oracle.py``.
"""

from __future__ import annotations

import numpy as np


def median_pattern_oracle(zyx: np.ndarray) -> np.ndarray:
    """float64 median along axis 0 by explicit selection (reference flat_field.py:118)."""
    z = zyx.shape[0]
    s = np.sort(zyx, axis=0).astype(np.float64)
    lo, hi = (z - 1) // 2, z // 2
    return (s[lo] + s[hi]) / 2.0 if lo != hi else s[lo].copy()


def flat_field_zyx_oracle(zyx: np.ndarray) -> np.ndarray:
    """float64 ``zyx / median * median.mean()`` (reference flat_field.py:118-119)."""
    pattern = median_pattern_oracle(zyx)
    mean = pattern.sum(dtype=np.float64) / pattern.size
    with np.errstate(divide="ignore", invalid="ignore"):
        return zyx.astype(np.float64) / pattern * mean


def flat_field_czyx_oracle(czyx: np.ndarray, target_indices) -> np.ndarray:
    """float32 CZYX adapter (reference flat_field.py:152-166)."""
    out = np.empty(czyx.shape, dtype=np.float32)
    target = set(target_indices)
    for c in range(czyx.shape[0]):
        if c in target:
            with np.errstate(over="ignore", invalid="ignore"):
                out[c] = flat_field_zyx_oracle(czyx[c])
        else:
            out[c] = czyx[c].astype(np.float32)
    return out
