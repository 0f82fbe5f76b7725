"""Golden vectors for the LEGACY `deskew_zyx` stages that live in the reference's own files
(MONAI's `Affine` is third-party and absent; these two are not):

* ``_average_n_slices_torch`` (reference biahub/deskew.py:71-96) and
* ``_fill_overhang_with_mean`` (reference biahub/deskew.py:277-336, scipy cross dilation),

run UNMODIFIED (oracle/ref_loader.py) on the reference's own ``fast_deskew_zyx(…, N=1)`` output.

    PYTHONPATH=/root/repo python tests/golden/make_golden_legacy.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference_deskew  # noqa: E402

CASES = [
    # name, shape, theta, px, keep_overhang, N, fill
    ("keep_n3_mean", (64, 31, 17), 30.0, 0.386, True, 3, "mean"),
    ("keep_n2_zero", (48, 21, 9), 36.0, 0.755, True, 2, "zero"),
    ("crop_n4", (96, 26, 8), 45.0, 0.5, False, 4, "zero"),
    ("keep_n1_mean", (40, 12, 12), 30.0, 0.386, True, 1, "mean"),
]


def main():
    ref = load_reference_deskew()
    arrays = {}
    for i, (name, shape, theta, px, keep, n, fill) in enumerate(CASES):
        raw = np.random.default_rng(700 + i).integers(0, 65536, size=shape, dtype=np.uint16)
        d1 = ref.fast_deskew_zyx(torch.from_numpy(raw.astype(np.float32)), theta, px, keep, 1)
        avg = ref._average_n_slices_torch(d1, n).numpy()
        out = ref._fill_overhang_with_mean(avg) if (keep and fill == "mean") else avg
        arrays[f"{name}_in"] = raw
        arrays[f"{name}_deskewed"] = d1.numpy()
        arrays[f"{name}_avg"] = avg
        arrays[f"{name}_out"] = out
        arrays[f"{name}_params"] = np.array([theta, px, float(keep), float(n), float(fill == "mean")])
    np.savez_compressed(os.path.join(HERE, "golden_legacy_v1.npz"), **arrays)
    print(f"wrote {len(arrays)} arrays")


if __name__ == "__main__":
    main()
