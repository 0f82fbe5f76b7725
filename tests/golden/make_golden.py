"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is not on the GPU box):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Deskew vectors come from the reference's own production path
``biahub.deskew._fast_deskew_czyx(device="cpu")`` (reference biahub/deskew.py:551-579) and
``get_deskewed_data_shape`` (:213-274), imported through ``oracle/ref_loader.py``.
Affine vectors come from ``scipy.ndimage.affine_transform`` — the library behind the
reference's ``method="scipy"`` branch (biahub/register.py:271-272); the ANTs branch cannot be
executed here (antspyx not installable), see oracle/affine_oracle.py.
Inputs are stored next to outputs so the fixtures are self-contained.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference_deskew  # noqa: E402
from oracle import affine_oracle as ao  # noqa: E402

DESKEW_CASES = [
    # name, shape, dtype, theta, px, keep_overhang, N, overhang_fill, num_splits
    ("u16_crop_n3", (64, 30, 16), "uint16", 30.0, 0.386, False, 3, 0, 1),
    ("u16_keep_n3", (64, 30, 16), "uint16", 30.0, 0.386, True, 3, 0, 1),
    ("u16_crop_n1", (256, 24, 8), "uint16", 30.0, 0.386, False, 1, 0, 1),
    ("f32_keep_n2_odd", (48, 21, 9), "float32", 36.0, 0.755, True, 2, 0, 1),
    ("u16_crop_n4_pad", (96, 26, 8), "uint16", 45.0, 0.5, False, 4, 0, 1),
    ("ref_test_shape", (2, 3, 4), "float64", 36.0, 0.386, True, 1, 0, 1),
    ("u16_keep_fill_mean", (64, 30, 17), "uint16", 30.0, 0.386, True, 3, "mean", 1),
    ("u16_keep_fill_100", (64, 30, 17), "uint16", 30.0, 0.386, True, 3, 100.0, 1),
    ("u16_crop_n3_splits2", (64, 30, 16), "uint16", 30.0, 0.386, False, 3, 0, 2),
    ("u16_px_gt1", (40, 12, 8), "uint16", 20.0, 1.25, True, 1, 0, 1),
]

SHAPE_CASES = [
    ((256, 256, 512), 30.0, 0.386, False, 1),
    ((256, 256, 512), 30.0, 0.386, False, 3),
    ((256, 256, 512), 30.0, 0.386, True, 3),
    ((800, 300, 2048), 30.0, 0.386, False, 3),
    ((800, 300, 2048), 30.0, 0.755, False, 3),
    ((2, 3, 4), 36.0, 0.386, True, 1),
    ((10, 500, 100), 30.0, 0.1, True, 1),
    ((64, 30, 16), 30.0, 0.386, False, 3),
    ((123, 77, 31), 12.34, 0.271, True, 5),
]


def make_input(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    if dtype == "uint16":
        return rng.integers(0, 65536, size=shape, dtype=np.uint16)
    vol = rng.random(shape, dtype=np.float32) * np.float32(4095.0)
    return vol.astype(dtype)


def main():
    ref = load_reference_deskew()
    arrays, meta = {}, {"deskew": [], "shapes": [], "affine": []}

    for i, (name, shape, dtype, theta, px, keep, n, fill, splits) in enumerate(DESKEW_CASES):
        raw = make_input(shape, dtype, 100 + i)
        out = ref._fast_deskew_czyx(
            raw[None], device="cpu", num_splits=splits, ls_angle_deg=theta, px_to_scan_ratio=px,
            keep_overhang=keep, average_n_slices=n, overhang_fill=fill,
        )
        assert out.dtype == np.float32 and out.ndim == 4
        arrays[f"deskew_{name}_in"] = raw
        arrays[f"deskew_{name}_out"] = out[0]
        meta["deskew"].append(dict(name=name, shape=shape, dtype=dtype, ls_angle_deg=theta,
                                   px_to_scan_ratio=px, keep_overhang=keep, average_n_slices=n,
                                   overhang_fill=fill, num_splits=splits))

    for shape, theta, px, keep, n in SHAPE_CASES:
        out_shape, voxel = ref.get_deskewed_data_shape(shape, theta, px, keep, n, 0.116)
        meta["shapes"].append(dict(raw_shape=shape, ls_angle_deg=theta, px_to_scan_ratio=px,
                                   keep_overhang=keep, average_n_slices=n, pixel_size_um=0.116,
                                   out_shape=[int(v) for v in out_shape],
                                   voxel_size=[float(v) for v in voxel]))

    # reference tests/test_cli/test_deskew_cli.py:11-30 — recomputed with the reference function
    data = np.arange(1, 17).reshape(4, 2, 2)
    for w in (1, 2, 3):
        arrays[f"avg_w{w}"] = ref._average_n_slices(data, average_window_width=w)
    arrays["avg_in"] = data

    # affine (scipy library = reference method="scipy" arithmetic, order 0/1)
    rng = np.random.default_rng(7)
    vol = (rng.random((12, 40, 48), dtype=np.float32) * np.float32(4095.0)).astype(np.float32)
    vol[3, 5, 7] = np.nan
    vol[4, 4, 4] = np.inf
    vol[5, 6, 8] = -np.inf
    arrays["affine_in"] = vol
    mats = {
        "c3_rot_scale_shift": ao.register_matrix_c3(vol.shape),
        "int_shift": ao.translation_matrix_zyx((-3, 1, 4)),
        "frac_shift": ao.translation_matrix_zyx((0.4, -2.25, 3.5)),
        "generic": np.array([[0.98, 0.05, -0.03, 1.2], [0.04, 1.02, 0.11, -3.3],
                             [-0.06, -0.09, 0.95, 4.7], [0, 0, 0, 1.0]]),
    }
    for mname, M in mats.items():
        for order in (0, 1):
            out_shape = (12, 40, 48) if mname != "generic" else (10, 44, 40)
            arrays[f"affine_{mname}_o{order}"] = ao.affine_oracle_scipy(vol, M, out_shape, order)
            meta["affine"].append(dict(name=mname, order=order, out_shape=out_shape,
                                       matrix=np.asarray(M).tolist()))

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **arrays)
    with open(os.path.join(HERE, "golden_v1.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    size = os.path.getsize(os.path.join(HERE, "golden_v1.npz"))
    print(f"wrote golden_v1.npz ({size/1e3:.0f} kB) with {len(arrays)} arrays")


if __name__ == "__main__":
    main()
