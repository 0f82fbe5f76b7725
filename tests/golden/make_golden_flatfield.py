"""Generate the flat-field golden vectors by running the UNMODIFIED reference
(``biahub.flat_field.flat_field_zyx`` / ``_flat_field_czyx``, reference biahub/flat_field.py:105-166).

Run in the build container only (``/root/reference`` is not on the GPU box):

    PYTHONPATH=/root/repo python tests/golden/make_golden_flatfield.py
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference_flat_field  # noqa: E402

CASES = [
    # name, (Z, Y, X), seed, low, high
    ("odd_z", (9, 5, 7), 0, 0, 65536),
    ("even_z", (10, 6, 8), 1, 0, 65536),
    ("camera_like", (16, 8, 12), 2, 90, 400),      # dark camera counts: many ties
    ("with_zero_columns", (8, 4, 6), 3, 0, 3),     # medians of 0 / 0.5: division by zero paths
    ("z2", (2, 3, 5), 4, 0, 65536),
    ("z1", (1, 3, 5), 5, 1, 65536),
]


def main():
    ff = load_reference_flat_field()
    out = {}
    for name, shape, seed, lo, hi in CASES:
        rng = np.random.default_rng(seed)
        data = rng.integers(lo, hi, size=shape, dtype=np.uint16)
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            res64 = ff.flat_field_zyx(data)
            czyx = np.stack([data, data[::-1].copy()])
            res32 = ff._flat_field_czyx(czyx, target_indices=[0])
        out[f"{name}__in"] = data
        out[f"{name}__zyx_f64"] = res64
        out[f"{name}__czyx_f32"] = res32
    path = os.path.join(HERE, "golden_flatfield_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
