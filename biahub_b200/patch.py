"""Re-point an importable ``biahub`` at the B200 implementations.

    import biahub_b200.patch as p; p.install()

After ``install()`` the reference's own drivers (``biahub deskew/register/stabilize`` CLI, YAML
settings, submitit executors, iohub I/O) run unchanged on top of the CUDA path: the names they
look up at call time (reference biahub/deskew.py:739-748, biahub/register.py:561-572,
biahub/stabilize.py:287-300) now resolve to this package's module-level functions, which pickle
by reference into spawn-ed workers as ``biahub_b200.<module>.<name>``.
"""

from __future__ import annotations

import importlib

_TARGETS = {
    "biahub.deskew": ("biahub_b200.deskew", [
        "_fast_deskew_czyx", "_deskew_czyx", "fast_deskew_zyx", "deskew_zyx",
        "get_deskewed_data_shape", "_average_n_slices", "_get_averaged_shape"]),
    "biahub.register": ("biahub_b200.register", [
        "apply_affine_transform", "convert_transform_to_ants", "convert_transform_to_numpy"]),
    "biahub.stabilize": ("biahub_b200.stabilize", ["apply_stabilization_transform"]),
    # reference biahub/flat_field.py:299-310 looks `_flat_field_czyx` up at call time
    "biahub.flat_field": ("biahub_b200.flat_field", [
        "flat_field_zyx", "flat_field_correction", "_flat_field_czyx"]),
    # duplicate helpers (reference biahub/registration/utils.py:774-853)
    "biahub.registration.utils": ("biahub_b200.register", ["apply_affine_transform"]),
}

_saved = {}


def install(strict: bool = False):
    """Patch every importable reference module; returns {module: [patched names]}."""
    patched = {}
    for ref_name, (our_name, names) in _TARGETS.items():
        try:
            ref_mod = importlib.import_module(ref_name)
        except Exception:
            if strict:
                raise
            continue
        ours = importlib.import_module(our_name)
        for name in names:
            if hasattr(ref_mod, name):
                _saved.setdefault((ref_name, name), getattr(ref_mod, name))
                setattr(ref_mod, name, getattr(ours, name))
                patched.setdefault(ref_name, []).append(name)
    return patched


def uninstall():
    for (ref_name, name), fn in list(_saved.items()):
        setattr(importlib.import_module(ref_name), name, fn)
        del _saved[(ref_name, name)]
