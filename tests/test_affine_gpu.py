"""GPU parity tests for the affine apply path of register / stabilize (through the C ABI)."""
import numpy as np
import pytest

from oracle import affine_oracle as ao

pytestmark = pytest.mark.gpu

TOL = 1e-4  # of the input dynamic range, order 1 (north_star); order 0 / integer shifts: bit-exact


def _vol(shape, seed, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    v = (rng.random(shape, dtype=np.float32) * np.float32(4095.0)).astype(np.float32)
    if nan_frac:
        m = rng.random(shape) < nan_frac
        v[m] = np.nan
        v.flat[7] = np.inf
        v.flat[11] = -np.inf
    return v


def _to_cuda(arr):
    import torch

    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(arr).cuda()


def _compare(got, want, order, rng=4095.0, name=""):
    assert got.shape == want.shape and got.dtype == np.float32, name
    if order == 0:
        assert np.array_equal(got, want), f"{name}: {np.sum(got != want)} voxels differ"
    else:
        big = np.abs(want) > 1e30  # taps scrubbed from +-inf: compare relatively
        err = np.abs(got[~big].astype(np.float64) - want[~big]).max()
        assert err <= TOL * rng, f"{name}: max err {err}"
        if big.any():
            assert np.allclose(got[big], want[big], rtol=1e-5)


def test_golden_scipy_vectors(golden):
    from biahub_b200 import affine_warp

    arrays, meta = golden
    vol = arrays["affine_in"]
    for c in meta["affine"]:
        want = arrays[f"affine_{c['name']}_o{c['order']}"]
        got = affine_warp(vol, np.array(c["matrix"]), tuple(c["out_shape"]), order=c["order"],
                          boundary="constant")
        _compare(got, want, c["order"], name=f"{c['name']}/o{c['order']}")


MATRICES = {
    "c3": lambda s: ao.register_matrix_c3(s),
    "int_shift": lambda s: ao.translation_matrix_zyx((-3, 1, 4)),
    "frac_shift": lambda s: ao.translation_matrix_zyx((0.4, -2.25, 3.5)),
    "zscale": lambda s: ao.translation_matrix_zyx((0.3, 1.5, -2.0)) @ ao.scaling_matrix_zyx(s, (0.5, 1.0, 1.0)),
    "zscale_up": lambda s: ao.scaling_matrix_zyx(s, (2.5, 0.9, 1.1)),
    "fliplr": lambda s: np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, s[2] - 1.0], [0, 0, 0, 1.0]]),
    "generic": lambda s: np.array([[0.98, 0.05, -0.03, 1.2], [0.04, 1.02, 0.11, -3.3],
                                   [-0.06, -0.09, 0.95, 4.7], [0, 0, 0, 1.0]]),
    "zflip": lambda s: np.array([[-1, 0, 0, s[0] - 1.0], [0, 1, 0, 0.5], [0, 0, 1, 0], [0, 0, 0, 1.0]]),
}


@pytest.mark.parametrize("path", ["auto", "gather"])
@pytest.mark.parametrize("boundary", ["constant", "itk"])
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("mname", sorted(MATRICES))
def test_device_api_matches_oracle(mname, order, boundary, path):
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (20, 72, 136)
    vol = _vol(shape, seed=3, nan_frac=0.001)
    M = MATRICES[mname](shape)
    out_shape = (22, 70, 140)
    want = ao.affine_oracle_numpy(vol, M, out_shape, order, boundary)
    p = _cabi.PATH_GATHER if path == "gather" else _cabi.PATH_AUTO
    got = affine_warp(_to_cuda(vol), M, out_shape, order=order, boundary=boundary, _path=p)
    torch.cuda.synchronize()
    _compare(got.cpu().numpy(), want, order, name=f"{mname}/{order}/{boundary}/{path}")


@pytest.mark.parametrize("mname", ["c3", "frac_shift", "int_shift", "zscale"])
def test_tma_path_is_taken_and_matches_gather(mname):
    """The z-separable matrices must be eligible for the TMA brick kernel (PATH_TMA raises
    otherwise) and agree with the independent gather kernel."""
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (24, 200, 264)
    vol = _vol(shape, seed=9)
    M = MATRICES[mname](shape)
    t = _to_cuda(vol)
    for order in (0, 1):
        for boundary in ("constant", "itk"):
            a = affine_warp(t, M, shape, order=order, boundary=boundary, _path=_cabi.PATH_TMA)
            b = affine_warp(t, M, shape, order=order, boundary=boundary, _path=_cabi.PATH_GATHER)
            if order == 0:
                assert torch.equal(a, b)
            else:
                assert (a - b).abs().max().item() <= 2e-6 * 4095.0  # 4-weight vs nested-lerp rounding


def test_unaligned_rows_not_tma_eligible():
    from biahub_b200 import _cabi, affine_warp
    from biahub_b200._cabi import B2Unsupported

    vol = _vol((8, 64, 63), seed=1)          # 63 floats per row: rows are not 16-byte aligned
    with pytest.raises(B2Unsupported):
        affine_warp(_to_cuda(vol), MATRICES["generic"](vol.shape), vol.shape, _path=_cabi.PATH_TMA)
    got = affine_warp(_to_cuda(vol), MATRICES["generic"](vol.shape), vol.shape).cpu().numpy()
    _compare(got, ao.affine_oracle_numpy(vol, MATRICES["generic"](vol.shape), vol.shape, 1, "itk"), 1)


@pytest.mark.parametrize("mname", ["generic", "zflip", "fliplr_tilt", "big_tilt", "half_out"])
def test_generic_brick_kernel(mname):
    """Non z-separable matrices take the 3-D TMA brick kernel (PATH_TMA raises otherwise): order 0
    bit-exact vs the float64 oracle, order 1 within tolerance, both boundary rules, uint16 too."""
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (28, 120, 200)
    mats = dict(MATRICES)
    mats["fliplr_tilt"] = lambda s: MATRICES["fliplr"](s) @ MATRICES["generic"](s)
    th = np.radians(12.0)
    mats["big_tilt"] = lambda s: np.array([[np.cos(th), 0, -np.sin(th), 8.0], [0, 1, 0, -0.5],
                                           [np.sin(th), 0, np.cos(th), -3.0], [0, 0, 0, 1.0]])
    # most tiles of the right half and of the top planes map outside the source: the zero-fill path
    mats["half_out"] = lambda s: ao.translation_matrix_zyx((-9.0, 3.5, 110.25)) @ MATRICES["generic"](s)
    M = mats[mname](shape)
    out_shape = (30, 116, 204)
    vol = _vol(shape, seed=17, nan_frac=0.001)
    u16 = np.random.default_rng(3).integers(0, 65536, size=shape, dtype=np.uint16)
    for boundary in ("constant", "itk"):
        for order in (0, 1):
            want = ao.affine_oracle_numpy(vol, M, out_shape, order, boundary)
            got = affine_warp(_to_cuda(vol), M, out_shape, order=order, boundary=boundary,
                              _path=_cabi.PATH_TMA)
            torch.cuda.synchronize()
            _compare(got.cpu().numpy(), want, order, name=f"{mname}/{order}/{boundary}")
            want = ao.affine_oracle_numpy(u16, M, out_shape, order, boundary)
            got = affine_warp(_to_cuda(u16), M, out_shape, order=order, boundary=boundary,
                              _path=_cabi.PATH_TMA)
            _compare(got.cpu().numpy(), want, order, rng=65535.0, name=f"{mname}/u16/{order}/{boundary}")


def test_uint16_source_and_crop():
    from biahub_b200 import affine_warp

    rng = np.random.default_rng(5)
    vol = rng.integers(0, 65536, size=(18, 80, 128), dtype=np.uint16)
    M = ao.register_matrix_c3(vol.shape)
    crop = (slice(2, 15), slice(5, 70), slice(16, 120))
    for order in (0, 1):
        want = ao.affine_oracle_numpy(vol, M, vol.shape, order, "itk", crop_output_slicing=crop)
        got = affine_warp(vol, M, vol.shape, order=order, boundary="itk", crop_output_slicing=crop)
        _compare(got, want, order, rng=65535.0, name=f"u16/crop/o{order}")
        got_d = affine_warp(_to_cuda(vol), M, vol.shape, order=order, boundary="itk",
                            crop_output_slicing=crop).cpu().numpy()
        assert np.array_equal(got, got_d)


def test_reference_known_answers():
    """reference tests/test_affine.py:26-59."""
    from biahub_b200 import apply_affine_transform

    ones = np.ones((10, 10, 10))
    for interp in ("linear", "nearestneighbor"):
        out = apply_affine_transform(ones, np.eye(4), (10, 10, 10), interpolation=interp)
        assert isinstance(out, np.ndarray) and out.shape == (10, 10, 10)
        assert np.all(out == 1)
    M = np.eye(4)
    M[:3, -1] = (-3, 1, 4)
    out = apply_affine_transform(ones, M, (10, 10, 10))
    assert out.shape == (10, 10, 10) and out.dtype == np.float32
    assert np.all(out[3:10, 0:9, 0:6] == 1)
    assert out.sum() == 7 * 9 * 6
    with pytest.raises(ValueError, match="Unknown method"):
        apply_affine_transform(ones, M, (10, 10, 10), method="cupy")
    # 4-D input → per-channel loop (reference register.py:240-251)
    out4 = apply_affine_transform(np.ones((2, 10, 10, 10)), M, (10, 10, 10),
                                  crop_output_slicing=(slice(3, 10), slice(0, 9), slice(0, 6)))
    assert out4.shape == (2, 7, 9, 6) and np.all(out4 == 1)


def test_stabilize_integer_translations_bit_exact():
    """Z-focus stabilisation matrices carry integer shifts: output must be a bit-exact shifted copy
    (zero outside), for every timepoint of a random-walk list."""
    from biahub_b200 import apply_stabilization_transform

    rng = np.random.default_rng(8)
    czyx = _vol((2, 12, 96, 160), seed=4)
    shifts = np.cumsum(rng.integers(-3, 4, size=(5, 3)), axis=0)
    mats = []
    for s in shifts:
        m = np.eye(4)
        m[:3, 3] = s
        mats.append(m)
    for t in range(len(mats)):
        out = apply_stabilization_transform(czyx, mats, t)
        assert out.shape == czyx.shape and out.dtype == np.float32
        dz, dy, dx = (int(v) for v in shifts[t])
        want = np.zeros_like(czyx)
        Z, Y, X = czyx.shape[1:]
        zs = slice(max(0, -dz), min(Z, Z - dz)); ys = slice(max(0, -dy), min(Y, Y - dy)); xs = slice(max(0, -dx), min(X, X - dx))
        zi = slice(zs.start + dz, zs.stop + dz); yi = slice(ys.start + dy, ys.stop + dy); xi = slice(xs.start + dx, xs.stop + dx)
        want[:, zs, ys, xs] = czyx[:, zi, yi, xi]
        assert np.array_equal(out, want), t


def test_stabilize_fractional_and_output_shape():
    from biahub_b200 import apply_stabilization_transform

    zyx = _vol((10, 64, 96), seed=6, nan_frac=0.002)
    mats = [np.eye(4) for _ in range(3)]
    mats[2][:3, 3] = (0.75, -1.5, 2.25)
    out = apply_stabilization_transform(zyx, mats, 2, output_shape=(12, 60, 100))
    want = ao.affine_oracle_numpy(zyx, mats[2], (12, 60, 100), 1, "itk")
    _compare(out, want, 1, name="stabilize/frac")


def test_full_size_spot_check_c3_plane_count():
    """C3-like geometry at full YX size (few planes): TMA kernel vs the float64 oracle on a random
    sample of voxels plus the volume corners."""
    import torch

    from biahub_b200 import _cabi, affine_warp

    shape = (6, 2048, 2048)
    vol = _vol(shape, seed=2000)
    M = ao.register_matrix_c3(shape)
    got = affine_warp(_to_cuda(vol), M, shape, order=1, boundary="constant", _path=_cabi.PATH_TMA)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    rng = np.random.default_rng(1)
    pts = np.stack([rng.integers(0, s, size=50000) for s in shape], axis=1)
    want = ao.affine_oracle_points(vol, M, pts, 1, "constant")
    err = np.abs(got[pts[:, 0], pts[:, 1], pts[:, 2]].astype(np.float64) - want).max()
    assert err <= TOL * 4095.0
    got0 = affine_warp(_to_cuda(vol), M, shape, order=0, boundary="constant").cpu().numpy()
    want0 = ao.affine_oracle_points(vol, M, pts, 0, "constant")
    assert np.array_equal(got0[pts[:, 0], pts[:, 1], pts[:, 2]], want0)


def test_pitched_chain_deskew_then_register():
    """Chained deskew -> register keeps the intermediate on the device with a padded row pitch
    (TMA-eligible) and must equal the two separate reference-facing calls bit for bit."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi
    from biahub_b200._device import row_pitch

    rng = np.random.default_rng(12)
    raw = rng.integers(0, 65536, size=(160, 24, 128), dtype=np.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    mid = b2._fast_deskew_czyx(raw[None], **kw)[0]            # (8, 128, 389): 389 % 4 != 0
    assert mid.shape[2] % 4 != 0
    M = ao.register_matrix_c3(mid.shape)
    # dense odd-width rows go through the gather kernel (nested lerps), the padded chain through
    # the TMA kernel (4-weight form): order 1 agrees to rounding, order 0 bit for bit
    want = b2.apply_affine_transform(mid, M, mid.shape)
    got = b2.deskew_then_register(raw, M, mid.shape, **kw)
    assert np.abs(got - want).max() <= 2e-6 * 65535.0
    want0 = b2.apply_affine_transform(mid, M, mid.shape, interpolation="nearestneighbor")
    got0 = b2.deskew_then_register(raw, M, mid.shape, interpolation="nearestneighbor", **kw)
    assert np.array_equal(got0, want0)
    oracle = ao.affine_oracle_numpy(mid, M, mid.shape, 1, "itk")
    assert np.abs(got - oracle).max() <= 1e-4 * 65535.0
    # the padded intermediate is a strided view and takes the TMA path
    t = torch.from_numpy(raw.view(np.int16)).cuda().view(torch.uint16)
    padded = b2.fast_deskew_zyx(t, 30.0, 0.386, False, 3, row_align=4)
    assert not padded.is_contiguous() and row_pitch(padded) % 4 == 0
    assert np.array_equal(padded.cpu().numpy(), mid)
    a = b2.affine_warp(padded, M, mid.shape, _path=_cabi.PATH_TMA)
    assert np.array_equal(a.cpu().numpy(), got)
    with pytest.raises(_cabi.B2Unsupported):                  # dense odd-width rows are not TMA-able
        b2.affine_warp(torch.from_numpy(mid).cuda(), M, mid.shape, _path=_cabi.PATH_TMA)


def test_find_overlapping_volume():
    """register's keep_overhang=False crop (reference register.py:345-394): warp ones on the GPU,
    largest interior rectangle on the host."""
    import biahub_b200 as b2

    M = np.eye(4)
    M[:3, 3] = (-3, 1, 4)                      # reference tests/test_affine.py:43-59 geometry
    z, y, x = b2.find_overlapping_volume((10, 10, 10), (10, 10, 10), M)
    assert (z, y, x) == (slice(3, 10), slice(0, 9), slice(0, 6))
    shape = (12, 96, 128)
    M2 = ao.register_matrix_c3(shape)
    z, y, x = b2.find_overlapping_volume(shape, shape, M2)
    mask = ao.affine_oracle_numpy(np.ones(shape, np.float32), M2, shape, 1, "itk") > 0
    assert mask[z, y, x].all()                  # the crop lies inside the warped support
    assert (z.stop - z.start) * (y.stop - y.start) * (x.stop - x.start) > 0.5 * mask.sum()


def test_ants_transform_shim_apply_to_image():
    """The estimation loops call ``convert_transform_to_ants(M).apply_to_image(ants_img,
    reference=ants_img).numpy()`` (reference biahub/optimize_registration.py:111): the shim takes
    image objects or arrays and resamples with the ITK boundary rule on the GPU."""
    import biahub_b200 as b2

    class FakeAntsImage:  # what ants.from_numpy returns, as far as the call sites use it
        def __init__(self, a):
            self._a, self.shape = a, a.shape

        def numpy(self):
            return self._a

    rng = np.random.default_rng(11)
    mov = (rng.random((10, 40, 48), dtype=np.float32) * 100).astype(np.float32)
    ref = np.zeros((12, 36, 52), np.float32)
    M = np.eye(4)
    M[:3, 3] = (0.5, -1.25, 2.0)
    M[1, 1] = 1.05
    t = b2.convert_transform_to_ants(M)
    want = b2.affine_warp(mov, M, ref.shape, order=1, boundary="itk")
    got = t.apply_to_image(FakeAntsImage(mov), reference=FakeAntsImage(ref))
    assert np.array_equal(got.numpy(), want) and got.shape == ref.shape
    assert np.array_equal(t.apply_to_image(mov, reference=ref), want)
    assert np.array_equal(np.asarray(got), want)


@pytest.mark.parametrize("mname", ["rot90", "rot90_scale_fliplr", "rot270_shift", "rot60"])
def test_lanes_along_y_variant(mname):
    """Matrices that map output y onto source x (the manual-registration family: scaling @ rotate90
    @ fliplr, reference biahub/estimate_registration.py:174-189) take the zsep kernel with lanes
    along y and a shared-memory transposed store: must match the oracle and the gather kernel,
    including ragged tiles (shape not a multiple of 64 x 16) and a cropped output."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi, affine_warp

    shape = (20, 150, 204)   # source rows 16-byte aligned (TMA-eligible); output ragged on purpose
    out_shape = (18, 210, 141)
    vol = _vol(shape, seed=31)
    mats = {
        "rot90": b2.get_3D_rotation_matrix(shape, 90),
        "rot90_scale_fliplr": (b2.get_3D_rescaling_matrix(shape, (1.0, 1.07, 1.07))
                               @ b2.get_3D_rotation_matrix(shape, 90) @ b2.get_3D_fliplr_matrix(shape)),
        "rot270_shift": ao.translation_matrix_zyx((0.4, 3.25, -11.5)) @ b2.get_3D_rotation_matrix(shape, 270),
        "rot60": b2.get_3D_rotation_matrix(shape, 60),
    }
    M = mats[mname]
    assert abs(M[2, 1]) > abs(M[2, 2])  # d src_x / d out_y dominates: the LY variant is selected
    t = _to_cuda(vol)
    for order in (0, 1):
        for boundary in ("constant", "itk"):
            want = ao.affine_oracle_numpy(vol, M, out_shape, order, boundary)
            a = affine_warp(t, M, out_shape, order=order, boundary=boundary, _path=_cabi.PATH_TMA)
            b = affine_warp(t, M, out_shape, order=order, boundary=boundary, _path=_cabi.PATH_GATHER)
            _compare(a.cpu().numpy(), want, order, name=f"{mname}/o{order}/{boundary}")
            if order == 0:
                assert torch.equal(a, b)
    crop = (slice(2, 15), slice(7, 190), slice(5, 133))
    want = ao.affine_oracle_numpy(vol, M, out_shape, 1, "itk")[crop]
    got = affine_warp(vol, M, out_shape, order=1, boundary="itk", crop_output_slicing=crop)
    _compare(got, want, 1, name=f"{mname}/crop")


@pytest.mark.parametrize("out_shape", [(21, 210, 141), (21, 210, 140), (24, 224, 160)])
@pytest.mark.parametrize("order", [0, 1])
def test_generic_lanes_along_y_variant(order, out_shape):
    """Non z-separable matrix with a 90-degree in-plane part (an ESTIMATED registration on top of
    the manual rotate90 approximation): brick kernel with lanes along y and staged stores.
    Output shapes: ragged with unaligned rows (scalar copy-out), ragged with 16-byte aligned rows
    (16-byte copy-out and zero fill on the full tiles, scalar on the last x tile), whole tiles."""
    import torch

    import biahub_b200 as b2
    from biahub_b200 import _cabi, affine_warp

    shape = (20, 150, 204)
    vol = _vol(shape, seed=33)
    c = (np.array(shape) - 1) / 2.0
    a, b = np.radians(1.5), np.radians(-0.8)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    tilt = np.eye(4)
    tilt[:3, :3] = Ry @ Rx
    tilt[:3, 3] = c - (Ry @ Rx) @ c
    M = (b2.get_3D_rescaling_matrix(shape, (1.0, 1.05, 1.05)) @ b2.get_3D_rotation_matrix(shape, 90)
         @ b2.get_3D_fliplr_matrix(shape) @ tilt)
    assert abs(M[2, 1]) > abs(M[2, 2]) and abs(M[0, 1]) + abs(M[0, 2]) > 0
    t = _to_cuda(vol)
    for boundary in ("constant", "itk"):
        want = ao.affine_oracle_numpy(vol, M, out_shape, order, boundary)
        got = affine_warp(t, M, out_shape, order=order, boundary=boundary, _path=_cabi.PATH_TMA)
        _compare(got.cpu().numpy(), want, order, name=f"generic-ly/o{order}/{boundary}")
        if order == 0:
            assert torch.equal(got, affine_warp(t, M, out_shape, order=0, boundary=boundary,
                                                _path=_cabi.PATH_GATHER))


@pytest.mark.parametrize("mname", ["c3", "generic", "zflip"])
def test_host_chain_pipeline_many_slabs(mname):
    """b2h_deskew_affine3d over several deskew slabs: output plane ranges are warped and
    downloaded as soon as their deskewed source planes exist.  Must equal the device chain (one
    deskew launch + one warp launch): bit for bit for z-separable matrices (the z-marching kernel
    does not depend on how the output is cut along z), to fp32 rounding for the brick kernel."""
    import torch

    import biahub_b200 as b2

    rng = np.random.default_rng(44)
    raw = rng.integers(0, 65536, size=(400, 120, 1024), dtype=np.uint16)
    kw = dict(ls_angle_deg=30.0, px_to_scan_ratio=0.386, keep_overhang=False, average_n_slices=3)
    t = torch.from_numpy(raw.view(np.int16)).cuda().view(torch.uint16)
    mid_shape = tuple(b2.fast_deskew_zyx(t, 30.0, 0.386, False, 3).shape)   # (40, 1024, 933)
    assert mid_shape[0] * mid_shape[1] * mid_shape[2] * 4 > 3 * (48 << 20)   # >= 4 slabs
    M = ao.register_matrix_c3(mid_shape)
    if mname == "generic":
        c = (np.array(mid_shape) - 1) / 2.0
        a = np.radians(0.8)
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        tilt = np.eye(4)
        tilt[:3, :3] = R
        tilt[:3, 3] = c - R @ c
        M = M @ tilt
    elif mname == "zflip":
        F = np.eye(4)
        F[0, 0] = -1.0
        F[0, 3] = mid_shape[0] - 1
        M = F @ M
    out_shape = (mid_shape[0] + 3, mid_shape[1] - 10, mid_shape[2] + 7)
    dev = b2.deskew_then_register(t, M, out_shape, **kw).cpu().numpy()
    host = b2.deskew_then_register(raw, M, out_shape, **kw)
    assert host.shape == out_shape and host.dtype == np.float32
    if mname == "c3":
        assert np.array_equal(host, dev)
    else:
        assert np.abs(host - dev).max() <= 2e-5 * 65535.0
    crop = (slice(3, 30), slice(5, 900), slice(10, 800))
    hc = b2.deskew_then_register(raw, M, out_shape, crop_output_slicing=crop, **kw)
    assert np.abs(hc - dev[crop]).max() <= (0 if mname == "c3" else 2e-5 * 65535.0)
