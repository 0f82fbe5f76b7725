"""Generate the affine-wrapper golden vectors by running the UNMODIFIED reference functions.

Run in the build container only (``/root/reference`` is not on the GPU box):

    PYTHONPATH=/root/repo python tests/golden/make_golden_affine.py

Every output below is produced by the reference's own code, imported from its files through
``oracle/ref_loader.load_reference_register_stabilize``:

* ``apply_affine_transform`` (reference biahub/register.py:202-281) — ``method="ants"`` with the
  fake ``ants`` module (its ``apply_to_image`` = the ITK-rule restatement; what the goldens pin
  is the reference's wrapper: NaN scrub order, float32 cast, 4-D channel loop, crop slicing,
  output shapes, parameter packing), and ``method="scipy"`` with the REAL scipy (order-3 spline,
  output dtype = input dtype; nothing faked on that branch);
* ``apply_stabilization_transform`` (biahub/stabilize.py:32-90);
* ``get_3D_rescaling_matrix`` / ``get_3D_rotation_matrix`` / ``get_3D_fliplr_matrix``
  (register.py:32-145), ``convert_transform_to_ants`` / ``convert_transform_to_numpy``
  (:148-199), ``rescale_voxel_size`` (:397-398), ``find_lir`` (:284-342, ``lir.lir`` =
  ``ref_loader.fake_lir_module``), ``find_overlapping_volume`` (:345-394).

Inputs are stored next to outputs so the fixtures are self-contained.
"""

from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import load_reference_register_stabilize  # noqa: E402


def _rot3(ax_deg):
    from scipy.spatial.transform import Rotation

    return Rotation.from_euler("zyx", ax_deg, degrees=True).as_matrix()


def _generic_matrix(shape, out_shape, angles=(4.0, -3.0, 6.5), scale=(1.02, 0.95, 1.06),
                    shift=(0.3, -1.25, 2.5)):
    """Out-of-plane rotation·scale about the volume centres (pull form)."""
    A = _rot3(angles) @ np.diag(scale)
    c_in = (np.array(shape) - 1) / 2
    c_out = (np.array(out_shape) - 1) / 2
    M = np.eye(4)
    M[:3, :3] = A
    M[:3, 3] = c_in - A @ c_out + np.array(shift)
    return M


def _c3_matrix(reg, shape):
    T = np.eye(4)
    T[:3, 3] = (0.4, 3.25, -11.5)
    return T @ reg.get_3D_rotation_matrix(shape, 7.3) @ reg.get_3D_rescaling_matrix(shape, (1.0, 1.07, 1.07))


def _volume(rng, shape, dtype):
    if np.dtype(dtype).kind == "u":
        return rng.integers(0, np.iinfo(dtype).max + 1, size=shape, dtype=dtype)
    if np.dtype(dtype).kind == "i":
        return rng.integers(-3000, 3000, size=shape).astype(dtype)
    return (rng.random(shape) * 4095.0).astype(dtype)


def _slices_to_list(sl):
    return [[int(s.start), int(s.stop)] for s in sl]


def main():
    reg, stab = load_reference_register_stabilize()
    rng = np.random.default_rng(20260101)
    arrays, meta = {}, {"apply": [], "stabilize": [], "matrices": [], "convert": [], "lir": [],
                        "overlap": [], "voxel": []}

    # ---- apply_affine_transform -------------------------------------------------------
    def apply_case(name, vol, matrix, out_shape, **kw):
        crop = kw.pop("crop", None)
        slicing = None if crop is None else tuple(slice(a, b) for a, b in crop)
        out = reg.apply_affine_transform(vol, matrix, out_shape, crop_output_slicing=slicing, **kw)
        arrays[f"apply_{name}_in"] = vol
        arrays[f"apply_{name}_M"] = np.asarray(matrix, dtype=np.float64)
        arrays[f"apply_{name}_out"] = out
        meta["apply"].append(dict(name=name, output_shape=list(out_shape), crop=crop,
                                  out_dtype=str(out.dtype), **kw))

    s3 = (10, 24, 40)
    v = _volume(rng, s3, np.float32)
    apply_case("f32_c3_linear", v, _c3_matrix(reg, s3), s3)
    apply_case("f32_c3_nearest", v, _c3_matrix(reg, s3), s3, interpolation="nearestneighbor")
    apply_case("f32_generic_linear", v, _generic_matrix(s3, (12, 20, 44)), (12, 20, 44))
    apply_case("f32_generic_nearest", v, _generic_matrix(s3, (12, 20, 44)), (12, 20, 44),
               interpolation="nearestneighbor")
    apply_case("u16_4d_crop", _volume(rng, (2, 8, 20, 24), np.uint16),
               _generic_matrix((8, 20, 24), (9, 22, 26)), (9, 22, 26), crop=[[1, 8], [2, 20], [3, 25]])
    bad = v.copy()
    bad[2, 5, 7] = np.nan
    bad[4, 10, 20] = np.inf
    bad[7, 18, 33] = -np.inf
    bad[0, 0, 0] = np.nan
    apply_case("f32_nonfinite", bad, _c3_matrix(reg, s3), s3)
    apply_case("f32_nonfinite_nearest", bad, _c3_matrix(reg, s3), s3, interpolation="nearestneighbor")
    v64 = _volume(rng, s3, np.float64)
    v64[3, 3, 3] = np.nan
    apply_case("f64_nan", v64, _c3_matrix(reg, s3), s3)
    # float64 holding +-inf and magnitudes beyond float32: np.nan_to_num scrubs in float64, the
    # float32 cast then overflows to +-inf (reference register.py:254, 266) — see DESIGN.md §8
    v64b = v64.copy()
    v64b[5, 5, 5] = np.inf
    v64b[6, 6, 6] = -1e39
    apply_case("f64_overflow", v64b, np.eye(4), s3, interpolation="nearestneighbor")
    apply_case("i16_linear", _volume(rng, s3, np.int16), _c3_matrix(reg, s3), s3)
    apply_case("u8_identity", _volume(rng, (4, 6, 8), np.uint8), np.eye(4), (4, 6, 8))
    M = np.eye(4)
    M[:3, 3] = (-3, 1, 4)
    apply_case("ref_kat_translation", np.ones((10, 10, 10)), M, (10, 10, 10))
    M = np.eye(4)
    M[:3, 3] = (0.5, -0.5, 9.5)   # the half-voxel band of the ITK rule on every axis
    apply_case("f32_half_voxel_band", _volume(rng, (6, 7, 12), np.float32), M, (7, 8, 12))
    apply_case("f32_half_voxel_band_nearest", arrays["apply_f32_half_voxel_band_in"], M, (7, 8, 12),
               interpolation="nearestneighbor")

    # method="scipy": the reference's literal call (order-3 spline, output dtype = input dtype)
    apply_case("scipy_f32", v, _c3_matrix(reg, s3), s3, method="scipy")
    apply_case("scipy_f32_generic", v, _generic_matrix(s3, (12, 20, 44)), (12, 20, 44), method="scipy")
    apply_case("scipy_u16", _volume(rng, s3, np.uint16), _c3_matrix(reg, s3), s3, method="scipy")
    apply_case("scipy_f64_nan", v64, _generic_matrix(s3, s3), s3, method="scipy")
    # NB reference register.py:272 passes output_shape_zyx in scipy's `offset` slot, which scipy
    # ignores for a homogeneous 4x4 matrix: the scipy branch ALWAYS resamples onto the INPUT's
    # shape ("scipy_f32_generic" above asks for (12,20,44) and gets (10,24,40)).  A 4-D call only
    # works when the cropped input-shaped result fits the (C,)+crop array it is assigned into.
    apply_case("scipy_u16_4d_crop", arrays["apply_u16_4d_crop_in"],
               _generic_matrix((8, 20, 24), (8, 20, 24)), (8, 20, 24),
               crop=[[1, 8], [2, 20], [3, 24]], method="scipy")
    apply_case("scipy_thin", _volume(rng, (1, 9, 13), np.float32), _c3_matrix(reg, (1, 9, 13)),
               (1, 9, 13), method="scipy")
    apply_case("scipy_i16", arrays["apply_i16_linear_in"], _c3_matrix(reg, s3), s3, method="scipy")

    # ---- apply_stabilization_transform ------------------------------------------------
    def stab_case(name, vol, mats, t, output_shape=None):
        with contextlib.redirect_stdout(io.StringIO()):
            out = stab.apply_stabilization_transform(vol, mats, t, output_shape)
        arrays[f"stab_{name}_in"] = vol
        arrays[f"stab_{name}_mats"] = np.asarray(mats, dtype=np.float64)
        arrays[f"stab_{name}_out"] = out
        meta["stabilize"].append(dict(name=name, t=t, output_shape=None if output_shape is None
                                      else list(output_shape), out_dtype=str(out.dtype)))

    mats = []
    walk = np.zeros(3)
    for _ in range(4):
        walk = walk + rng.integers(-3, 4, size=3)
        m = np.eye(4)
        m[:3, 3] = walk
        mats.append(m)
    stab_case("int_shift_3d", v, mats, 2)
    fr = [np.eye(4) for _ in range(3)]
    for k, m in enumerate(fr):
        m[:3, 3] = rng.normal(0, 1.5, size=3) * (k + 1)
    stab_case("frac_shift_4d_u16", _volume(rng, (2, 6, 18, 22), np.uint16), fr, 1)
    stab_case("frac_shift_shape", bad, fr, 2, (8, 20, 36))

    # ---- matrix builders ----------------------------------------------------------------
    for name, fn, args in [
        ("rescale_same", reg.get_3D_rescaling_matrix, ((120, 2048, 2048), (1.0, 1.07, 1.07))),
        ("rescale_end", reg.get_3D_rescaling_matrix, ((60, 500, 700), (2.0, 0.5, 1.5), (61, 400, 900))),
        ("rescale_default", reg.get_3D_rescaling_matrix, ((60, 500, 701),)),
        ("rotate_same", reg.get_3D_rotation_matrix, ((120, 2048, 2048), 7.3)),
        ("rotate_end", reg.get_3D_rotation_matrix, ((60, 501, 700), 90, (60, 700, 501))),
        ("rotate_neg", reg.get_3D_rotation_matrix, ((3, 11, 17), -33.3)),
        ("fliplr_same", reg.get_3D_fliplr_matrix, ((60, 501, 700),)),
        ("fliplr_end", reg.get_3D_fliplr_matrix, ((60, 501, 700), (60, 400, 901))),
    ]:
        out = np.asarray(fn(*args), dtype=np.float64)
        arrays[f"matrix_{name}"] = out
        meta["matrices"].append(dict(name=name, fn=fn.__name__, args=[list(a) if isinstance(a, tuple) else a
                                                                       for a in args]))

    # ---- ITK parameter packing ----------------------------------------------------------
    Mg = _generic_matrix(s3, (12, 20, 44))
    T = reg.convert_transform_to_ants(Mg)
    arrays["convert_generic_M"] = Mg
    arrays["convert_generic_params"] = np.asarray(T.parameters, dtype=np.float64)
    arrays["convert_generic_back"] = reg.convert_transform_to_numpy(T)
    T.set_fixed_parameters([3.0, -7.5, 11.25])
    arrays["convert_generic_fixed"] = np.array([3.0, -7.5, 11.25])
    arrays["convert_generic_back_fixed"] = reg.convert_transform_to_numpy(T)
    meta["convert"].append(dict(name="generic"))

    # ---- rescale_voxel_size ---------------------------------------------------------------
    for k, (m, scale) in enumerate([(Mg[:3, :3], [0.25, 0.1, 0.1]), (_c3_matrix(reg, s3)[:3, :3], [1, 2, 3]),
                                    (np.eye(3) * 2.0, [2, 1, 3])]):
        arrays[f"voxel_{k}_M"] = np.asarray(m, dtype=np.float64)
        arrays[f"voxel_{k}_scale"] = np.asarray(scale, dtype=np.float64)
        arrays[f"voxel_{k}_out"] = np.asarray(reg.rescale_voxel_size(m, scale), dtype=np.float64)
        meta["voxel"].append(k)

    # ---- find_lir / find_overlapping_volume ---------------------------------------------
    kat = np.zeros((10, 10, 10))
    kat[2:8, 0:9, 3:10] = 1
    blob = np.zeros((9, 14, 16), dtype=bool)
    zz, yy, xx = np.ogrid[:9, :14, :16]
    blob[((zz - 4) / 4.2) ** 2 + ((yy - 6.5) / 6.0) ** 2 + ((xx - 8) / 7.0) ** 2 <= 1.0] = True
    wedge = np.ones((8, 12, 12), dtype=bool)
    wedge[:2] = False
    wedge[:, :, :3] = False
    wedge[5:, 8:, :] = False
    for name, mask in [("ref_kat", kat), ("ellipsoid", blob), ("wedge", wedge)]:
        arrays[f"lir_{name}_mask"] = np.asarray(mask).astype(np.uint8)
        meta["lir"].append(dict(name=name, slices=_slices_to_list(reg.find_lir(mask))))

    for name, ishape, tshape, Mx in [
        ("translation", (10, 12, 14), (10, 12, 14), mats[1]),
        ("c3_style", (8, 40, 48), (8, 40, 48), _c3_matrix(reg, (8, 40, 48))),
        ("generic", (10, 24, 40), (12, 20, 44), _generic_matrix((10, 24, 40), (12, 20, 44))),
    ]:
        with contextlib.redirect_stdout(io.StringIO()):
            sl = reg.find_overlapping_volume(ishape, tshape, Mx)
        arrays[f"overlap_{name}_M"] = np.asarray(Mx, dtype=np.float64)
        meta["overlap"].append(dict(name=name, input_shape=list(ishape), target_shape=list(tshape),
                                    slices=_slices_to_list(sl)))

    np.savez_compressed(os.path.join(HERE, "golden_affine_v1.npz"), **arrays)
    with open(os.path.join(HERE, "golden_affine_v1.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"wrote {len(arrays)} arrays; cases: " + ", ".join(f"{k}={len(v)}" for k, v in meta.items()))


if __name__ == "__main__":
    main()
