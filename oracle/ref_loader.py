"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference compute module.

Imports ``/root/reference/biahub/deskew.py`` in THIS container (the GPU box has
no ``/root/reference``) with inert stand-ins for the third-party packages that
are not installed (iohub, monai, submitit, natsort), so that the reference's
own ``fast_deskew_zyx`` / ``_fast_deskew_czyx(device="cpu")`` /
``get_deskewed_data_shape`` / ``_average_n_slices`` / ``_fill_overhang_torch``
can be executed on CPU torch.  It is used by ``tests/golden/make_golden.py`` to
generate the committed golden vectors and by CPU tests that pin ``oracle/``
against the real reference when it is mounted.  Nothing under ``biahub_b200/``
may import this file.

Recipe: SURVEY.md Appendix B.  Reference import list: biahub/deskew.py:1-36.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BIAHUB_REFERENCE_ROOT", "/root/reference")


class _Sink:
    """Attribute/call sink: importable, but raises if the reference calls it."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("third-party stub called (not available in this container)")

    def __getattr__(self, name):
        return _Sink()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "biahub", "deskew.py"))


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # behaves as a package
    sys.modules[name] = mod
    return mod


_LOADED = None


def load_reference_deskew():
    """Return the reference's ``biahub.deskew`` module (CPU torch only)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")

    import torch.multiprocessing as tmp

    start_method_before = tmp.get_start_method(allow_none=True)

    for name in (
        "submitit",
        "iohub",
        "iohub.ngff",
        "iohub.ngff.utils",
        "monai",
        "monai.transforms",
        "monai.transforms.spatial",
        "monai.transforms.spatial.array",
        "natsort",
    ):
        if name not in sys.modules:
            _stub(name)
    sys.modules["iohub"].open_ome_zarr = _Sink()
    sys.modules["iohub.ngff"].open_ome_zarr = _Sink()
    sys.modules["iohub.ngff"].Plate = type("Plate", (), {})
    sys.modules["iohub.ngff.utils"].create_empty_plate = _Sink()
    sys.modules["iohub.ngff.utils"].process_single_position = _Sink()
    sys.modules["monai.transforms.spatial.array"].Affine = _Sink
    sys.modules["natsort"].natsorted = sorted
    sys.modules["submitit"].Job = object
    sys.modules["submitit"].AutoExecutor = _Sink
    sys.modules["submitit"].helpers = _Sink()

    saved_pkg = sys.modules.get("biahub")
    pkg = types.ModuleType("biahub")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "biahub")]
    sys.modules["biahub"] = pkg
    try:
        mod = importlib.import_module("biahub.deskew")
    finally:
        # biahub/deskew.py:40 forces the start method process-wide; undo that side effect.
        try:
            tmp.set_start_method(start_method_before, force=True)
        except Exception:
            pass
        if saved_pkg is not None:
            sys.modules["biahub"] = saved_pkg
    _LOADED = mod
    return mod


def load_reference_settings():
    """Return the reference's ``biahub.settings`` (pydantic models; needs no stubs)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    if "biahub" not in sys.modules:
        pkg = types.ModuleType("biahub")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "biahub")]
        sys.modules["biahub"] = pkg
    return importlib.import_module("biahub.settings")


def load_reference_flat_field():
    """Return the reference's ``biahub.flat_field`` module (numpy only; same stubs as deskew).

    ``flat_field_zyx`` / ``_flat_field_czyx`` (reference biahub/flat_field.py:105-122, 152-166)
    are plain numpy and run unmodified."""
    load_reference_deskew()  # installs the third-party stand-ins
    saved_pkg = sys.modules.get("biahub")
    pkg = types.ModuleType("biahub")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "biahub")]
    sys.modules["biahub"] = pkg
    try:
        return importlib.import_module("biahub.flat_field")
    finally:
        if saved_pkg is not None:
            sys.modules["biahub"] = saved_pkg
