"""CPU tests: host-side logic of the package, the C-ABI library's exports, and loud failure
without a GPU (no compute calls succeed here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import biahub_b200 as b2
from biahub_b200 import _cabi
from oracle import affine_oracle as ao
from oracle import deskew_oracle as do

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _no_gpu():
    import torch

    return not torch.cuda.is_available()


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "biahub_b200.h")).read()
    declared = set(re.findall(r"B2_API\s+[\w\s\*]+?\b(b2h?_\w+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(handle, name), name
    assert _cabi.lib().b2_abi_version() == _cabi.ABI_VERSION


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    assert _cabi.device_count() == 0
    raw = np.zeros((8, 4, 8), dtype=np.uint16)
    with pytest.raises(_cabi.B2Error, match="no CPU fallback"):
        b2._fast_deskew_czyx(raw[None], ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                             keep_overhang=True, average_n_slices=1)
    with pytest.raises(_cabi.B2Error, match="no CPU fallback"):
        b2.apply_affine_transform(np.ones((4, 4, 4)), np.eye(4), (4, 4, 4))
    with pytest.raises(_cabi.B2Error):
        b2.apply_stabilization_transform(np.ones((4, 4, 4)), [np.eye(4)], 0)
    u16 = np.ones((4, 4, 8), dtype=np.uint16)
    with pytest.raises(_cabi.B2Error, match="no CPU fallback"):
        b2.flat_field_zyx(u16)
    with pytest.raises(_cabi.B2Error, match="no CPU fallback"):
        b2._flat_field_czyx(u16[None], [0])
    with pytest.raises(_cabi.B2Error, match="no CPU fallback"):
        b2.deskew_then_register(raw, np.eye(4), (8, 8, 8), ls_angle_deg=30.0, px_to_scan_ratio=0.386,
                                keep_overhang=True)


def test_flat_field_host_logic_without_gpu():
    """Argument handling that needs no device: dtype / rank errors, pass-through channels."""
    with pytest.raises(NotImplementedError, match="uint16"):
        b2.flat_field_zyx(np.ones((3, 4, 5), dtype=np.float32))
    with pytest.raises(ValueError):
        b2.flat_field_zyx(np.ones((4, 5), dtype=np.uint16))
    with pytest.raises(ValueError):
        b2._flat_field_czyx(np.ones((3, 4, 5), dtype=np.uint16), [0])
    # no target channel: every channel is passed through as float32, nothing touches the device
    czyx = np.arange(2 * 3 * 4 * 5, dtype=np.uint16).reshape(2, 3, 4, 5)
    out = b2._flat_field_czyx(czyx, [])
    assert out.dtype == np.float32 and np.array_equal(out, czyx.astype(np.float32))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "biahub_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_shape_logic_matches_golden_and_oracle(golden):
    _, meta = golden
    for c in meta["shapes"]:
        shape, voxel = b2.get_deskewed_data_shape(tuple(c["raw_shape"]), c["ls_angle_deg"],
                                                  c["px_to_scan_ratio"], c["keep_overhang"],
                                                  c["average_n_slices"], c["pixel_size_um"])
        assert list(shape) == c["out_shape"]
        assert list(voxel) == c["voxel_size"]
    with pytest.raises(ValueError, match="Dataset contains only overhang"):
        b2.get_deskewed_data_shape((10, 500, 100), 30, 0.1, keep_overhang=False)
    assert b2.get_deskewed_data_shape((256, 256, 512), 30.0, 0.386, False, 1)[0] == (256, 512, 442)
    assert b2.get_deskewed_data_shape((800, 300, 2048), 30.0, 0.386, False, 3)[0] == (100, 2048, 1813)


def test_deskew_scalars_equal_oracle():
    from biahub_b200.deskew import deskew_scalars

    for shape, th, px, keep, n in [((256, 256, 512), 30.0, 0.386, False, 3), ((64, 31, 16), 12.5, 0.755, True, 2)]:
        s = deskew_scalars(shape, th, px, keep, n)
        o = do.deskew_scalars(shape, th, px, keep)
        assert (s["Zo"], s["Yo"], s["Xo"]) == (o["Zo"], o["Yo"], o["Xo"])
        assert np.float32(s["px32"]) == o["px32"] and np.float32(s["pxct32"]) == o["pxct32"]
        assert np.float32(s["off32"]) == o["off32"]
        assert s["Zavg"] == int(np.ceil(s["Zo"] / n))
    with pytest.raises(ValueError):
        deskew_scalars((1, 4, 4), 30.0, 0.386, True)


def test_average_n_slices_known_answer():
    # reference tests/test_cli/test_deskew_cli.py:11-30
    data = np.arange(1, 17).reshape(4, 2, 2)
    assert np.array_equal(b2._average_n_slices(data, 3), np.array([[[5, 6], [7, 8]], [[13, 14], [15, 16]]]))
    assert np.array_equal(b2._average_n_slices(data, 2), np.array([[[3, 4], [5, 6]], [[11, 12], [13, 14]]]))
    assert np.array_equal(b2._average_n_slices(data, 1), data)
    for w in (1, 2, 3):
        assert b2._average_n_slices(data, w).shape == b2._get_averaged_shape(data.shape, w)


def test_matrix_helpers():
    shape = (10, 200, 300)
    assert np.allclose(b2.get_3D_rotation_matrix(shape, 7.3), ao.rotation_matrix_yx(shape, 7.3))
    assert np.allclose(b2.get_3D_rescaling_matrix(shape, (1, 1.07, 1.07)), ao.scaling_matrix_zyx(shape, (1, 1.07, 1.07)))
    end = (10, 100, 150)
    assert np.allclose(b2.get_3D_rotation_matrix(shape, 30, end), ao.rotation_matrix_yx(shape, 30, end))
    f = b2.get_3D_fliplr_matrix(shape)
    assert f[2, 2] == -1 and f[2, 3] == 300
    # the rotation keeps the YX centre fixed
    c = np.array([0, 100, 150, 1.0])
    assert np.allclose(b2.get_3D_rotation_matrix(shape, 33.0) @ c, c)
    # reference tests/test_cli/test_register_cli.py:42-72
    one = np.array([1, 1, 1])
    assert np.allclose(b2.rescale_voxel_size(np.diag([2, 3, 4]), one), [2, 3, 4])
    assert np.allclose(b2.rescale_voxel_size(np.diag([2, -3, 4]), one), [2, 3, 4])
    assert np.allclose(b2.rescale_voxel_size(np.array([[0, 2, 0], [1, 0, 0], [0, 0, 3]]), one), [2, 1, 3])
    th = np.pi / 3
    m4 = np.array([[2, 0, 0], [0, 3 * np.cos(th), -3 * np.sin(th)], [0, 3 * np.sin(th), 3 * np.cos(th)]])
    assert np.allclose(b2.rescale_voxel_size(m4, one), [2, 3, 3])


def test_ants_parameter_round_trip():
    # reference tests/test_affine.py:12-23 (types/shapes) + exact round trip
    T = b2.convert_transform_to_ants(np.eye(4))
    assert T.parameters.shape == (12,)
    M = ao.register_matrix_c3((10, 200, 300))
    back = b2.convert_transform_to_numpy(b2.convert_transform_to_ants(M))
    assert back.shape == (4, 4) and np.array_equal(back, M)
    T2 = b2.register.ItkAffineParameters()
    T2.set_parameters(np.arange(12.0))
    T2.set_fixed_parameters([1.0, 2.0, 3.0])
    A = np.arange(9.0).reshape(3, 3)
    want_t = np.arange(9.0, 12.0) + (np.eye(3) - A) @ np.array([1.0, 2.0, 3.0])
    assert np.allclose(b2.convert_transform_to_numpy(T2)[:3, 3], want_t)


def test_argument_validation_happens_before_gpu():
    with pytest.raises(ValueError, match="Unknown method"):
        b2.apply_affine_transform(np.ones((4, 4, 4)), np.eye(4), (4, 4, 4), method="cupy")
    with pytest.raises(NotImplementedError):
        b2.apply_affine_transform(np.ones((4, 4, 4)), np.eye(4), (4, 4, 4), interpolation="bspline")
    with pytest.raises(ValueError):   # unit-step crop slices only
        b2.spline_warp(np.ones((4, 4, 4), np.float32), np.eye(4), (slice(0, 4, 2),) * 3)
    with pytest.raises(ValueError):
        b2.affine_warp(np.ones((4, 4, 4)), np.eye(3), (4, 4, 4))
    with pytest.raises(ValueError, match="Dataset contains only overhang"):
        b2.deskew_zyx(np.zeros((10, 500, 100), np.uint16), 30, 0.1, keep_overhang=False)
    with pytest.raises(ValueError):
        b2.apply_stabilization_transform(np.ones((4, 4, 4)), [np.eye(3)], 0)


def test_device_resolution(monkeypatch):
    from biahub_b200 import _device

    monkeypatch.setenv("BIAHUB_B200_DEVICE", "cuda:3")
    assert _device.resolve_device("cpu") == 3
    monkeypatch.setenv("BIAHUB_B200_DEVICE", "2")
    assert _device.resolve_device(None) == 2
    monkeypatch.delenv("BIAHUB_B200_DEVICE")
    assert _device.resolve_device("cuda:5") == 5
    assert _device.resolve_device(1) == 1
    with pytest.raises(ValueError):
        _device.resolve_device("tpu")


def test_find_lir_known_answer():
    # reference tests/test_cli/test_register_cli.py:75-86
    data = np.zeros((10, 10, 10))
    data[2:8, 0:9, 3:10] = 1
    z_slice, y_slice, x_slice = b2.find_lir(data)
    assert (z_slice, y_slice, x_slice) == (slice(2, 8), slice(0, 9), slice(3, 10))


def test_largest_interior_rectangle_bruteforce():
    from biahub_b200.register import largest_interior_rectangle

    rng = np.random.default_rng(0)
    for _ in range(30):
        m = rng.random((7, 9)) > 0.3
        x, y, w, h = largest_interior_rectangle(m)
        assert w * h == 0 or m[y:y + h, x:x + w].all()
        best = 0
        for y0 in range(7):
            for y1 in range(y0 + 1, 8):
                for x0 in range(9):
                    for x1 in range(x0 + 1, 10):
                        if m[y0:y1, x0:x1].all():
                            best = max(best, (y1 - y0) * (x1 - x0))
        assert w * h == best
    assert largest_interior_rectangle(np.zeros((3, 3), bool)) == (0, 0, 0, 0)


def test_pinned_result_pool_recycles_blocks():
    """Result arrays come from a recycling pool: a block is reused once the array AND its views
    are gone, the cap is honoured, and the arrays behave like fresh C-contiguous float32 arrays."""
    import gc

    from biahub_b200._device import PinnedResultPool

    allocs = []

    def fake_alloc(nbytes):
        buf = np.empty(nbytes, dtype=np.uint8)
        allocs.append(nbytes)
        return {"ptr": buf.ctypes.data, "nbytes": nbytes, "keep": buf}

    pool = PinnedResultPool(alloc=fake_alloc, cap_bytes=3 * 4 * 1000)
    a = pool.empty((10, 10, 10), np.float32)
    assert a.shape == (10, 10, 10) and a.dtype == np.float32
    assert a.flags.c_contiguous and a.flags.writeable
    a[...] = 7.0
    ptr_a = a.ctypes.data
    view = a[2:5]
    del a
    gc.collect()
    assert pool.handed_out == 1 and not pool.free          # the view keeps the block alive
    b = pool.empty((10, 10, 10), np.float32)               # second block
    assert b.ctypes.data != ptr_a and len(allocs) == 2
    del view
    gc.collect()
    assert len(pool.free) == 1
    c = pool.empty((5, 10, 10), np.float32)                # smaller request reuses the idle block
    assert c.ctypes.data == ptr_a and len(allocs) == 2
    d = pool.empty((10, 10, 10), np.float32)               # third block: at the cap now
    assert d is not None and len(allocs) == 3
    assert pool.empty((10, 10, 10), np.float32) is None    # exhausted -> caller falls back
    del b, c, d
    gc.collect()
    assert pool.handed_out == 0 and len(pool.free) == 3
    # a request larger than any idle block evicts idle blocks to stay under the cap
    e = pool.empty((20, 10, 10), np.float32)
    assert e is not None and pool.total_bytes <= pool.cap_bytes
