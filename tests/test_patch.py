"""patch.install() re-points a (stand-in) biahub package at the B200 functions."""
import sys
import types

import biahub_b200
from biahub_b200 import patch


def test_install_and_uninstall(monkeypatch):
    pkg = types.ModuleType("biahub")
    pkg.__path__ = []
    deskew = types.ModuleType("biahub.deskew")
    deskew._fast_deskew_czyx = lambda *a, **k: "reference"
    deskew.unrelated = 1
    stabilize = types.ModuleType("biahub.stabilize")
    stabilize.apply_stabilization_transform = lambda *a, **k: "reference"
    for name, mod in (("biahub", pkg), ("biahub.deskew", deskew), ("biahub.stabilize", stabilize)):
        monkeypatch.setitem(sys.modules, name, mod)
    done = patch.install()
    assert done["biahub.deskew"] == ["_fast_deskew_czyx"]
    assert deskew._fast_deskew_czyx is biahub_b200._fast_deskew_czyx
    assert stabilize.apply_stabilization_transform is biahub_b200.apply_stabilization_transform
    assert deskew._fast_deskew_czyx.__module__ == "biahub_b200.deskew"  # pickles by reference
    patch.uninstall()
    assert deskew._fast_deskew_czyx() == "reference"


def test_install_repoints_the_live_reference_estimation_loops():
    """With the reference's own modules imported (fake ants, stubbed third parties): every name
    the estimation loops bound with ``from biahub.register import …`` and ``Transform.to_ants``
    resolve to the B200 implementations after install(), and back after uninstall()."""
    import numpy as np
    import pytest

    from oracle import ref_loader as rl

    if not rl.reference_available():
        pytest.skip("/root/reference not mounted")
    import biahub_b200.register as br

    mods = rl.load_reference_estimation_modules()
    try:
        done = patch.install()
        opt = mods["biahub.optimize_registration"]
        assert opt.convert_transform_to_ants is br.convert_transform_to_ants   # optimize_registration.py:96,275
        assert opt.find_lir is br.find_lir                                     # :137
        assert mods["biahub.registration.ants"].find_lir is br.find_lir        # registration/ants.py:233
        assert mods["biahub.estimate_registration"].convert_transform_to_ants is br.convert_transform_to_ants
        assert mods["biahub.stabilize"].convert_transform_to_ants is br.convert_transform_to_ants
        utils = mods["biahub.registration.utils"]
        assert utils.apply_affine_transform is br.apply_affine_transform
        assert utils.find_overlapping_volume is br.find_overlapping_volume
        assert mods["biahub.register"].find_overlapping_volume is br.find_overlapping_volume
        assert "Transform.to_ants" in done["biahub.core.transform"]
        # registration/beads.py:117 `approx_transform.to_ants().apply_to_image(...)`
        T = mods["biahub.core.transform"].Transform
        M = np.eye(4)
        M[:3, 3] = (1.5, -2.0, 3.25)
        M[1, 2] = 0.1
        t = T(M).to_ants()
        assert isinstance(t, br.ItkAffineParameters)
        assert np.array_equal(t.as_matrix(), M)
        # estimate_registration.py:190 `convert_transform_to_ants(M).invert()`
        assert np.allclose(br.convert_transform_to_ants(M).invert().as_matrix(), np.linalg.inv(M), atol=1e-12)
        patch.uninstall()
        assert opt.convert_transform_to_ants.__module__ == "biahub.register"
        assert T.to_ants is not patch._transform_to_ants
    finally:
        patch.uninstall()
        rl.unload_reference_package()
