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
