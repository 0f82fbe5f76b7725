"""Re-point an importable ``biahub`` at the B200 implementations.

    import biahub_b200.patch as p; p.install()

After ``install()`` the reference's own drivers (``biahub deskew/register/stabilize`` CLI, YAML
settings, submitit executors, iohub I/O) run unchanged on top of the CUDA path: the names they
look up at call time (reference biahub/deskew.py:739-748, biahub/register.py:561-572,
biahub/stabilize.py:287-300) now resolve to this package's module-level functions, which pickle
by reference into spawn-ed workers as ``biahub_b200.<module>.<name>``.

Three kinds of call sites are covered:

* functions looked up in their defining module (``_TARGETS``);
* names other reference modules bound with ``from biahub.register import …`` at import time —
  the estimation loops (reference biahub/optimize_registration.py:15-19,96,111,137,275-282;
  biahub/registration/ants.py:43-46,204,233; biahub/estimate_registration.py:21-24,190-192,
  330-332) and ``biahub/stabilize.py:24``: every already imported ``biahub.*`` module whose
  attribute IS a replaced original is re-bound (``_rebind``), and modules named in
  ``_IMPORTERS`` are imported first so that they are seen;
* ``Transform.to_ants()`` (reference biahub/core/transform.py:427-456), which the bead
  registration uses to warp (biahub/registration/beads.py:117,194,920,966): it returns the
  ``ItkAffineParameters`` stand-in, whose ``apply_to_image`` / ``invert`` run on the GPU.
"""

from __future__ import annotations

import importlib
import sys

_TARGETS = {
    "biahub.deskew": ("biahub_b200.deskew", [
        "_fast_deskew_czyx", "_deskew_czyx", "fast_deskew_zyx", "deskew_zyx",
        "get_deskewed_data_shape", "_average_n_slices", "_average_n_slices_torch",
        "_get_averaged_shape"]),
    "biahub.register": ("biahub_b200.register", [
        "apply_affine_transform", "convert_transform_to_ants", "convert_transform_to_numpy",
        "find_lir", "find_overlapping_volume"]),
    "biahub.stabilize": ("biahub_b200.stabilize", ["apply_stabilization_transform"]),
    # reference biahub/flat_field.py:299-310 looks `_flat_field_czyx` up at call time
    "biahub.flat_field": ("biahub_b200.flat_field", [
        "flat_field_zyx", "flat_field_correction", "_flat_field_czyx"]),
    # duplicate helpers (reference biahub/registration/utils.py:467-640, 774-853)
    "biahub.registration.utils": ("biahub_b200.register", [
        "apply_affine_transform", "convert_transform_to_ants", "convert_transform_to_numpy",
        "find_lir", "find_overlapping_volume"]),
}

# modules that bind the names above with `from … import …` (imported so that _rebind sees them)
_IMPORTERS = (
    "biahub.optimize_registration", "biahub.registration.ants", "biahub.registration.beads",
    "biahub.estimate_registration", "biahub.estimate_stabilization",
)

_saved = {}          # (module name, attribute) -> original object
_saved_methods = {}  # (module name, class name, attribute) -> original function


def _transform_to_ants(self):
    """Replacement of ``biahub.core.transform.Transform.to_ants`` for 3-D transforms."""
    from .register import ItkAffineParameters

    import numpy as np

    if getattr(self, "_ndim", 3) != 3:
        orig = _saved_methods.get(("biahub.core.transform", "Transform", "to_ants"))
        if orig is None:
            raise NotImplementedError("only 3-D transforms run on the B200 path")
        return orig(self)
    m = np.asarray(self._matrix, dtype=np.float64)
    return ItkAffineParameters(np.concatenate([m[:3, :3].ravel(), m[:3, 3]]))


def _rebind(replaced: dict) -> dict:
    """Re-bind ``from``-imported copies: any attribute of an imported ``biahub.*`` module that IS
    a replaced original."""
    done = {}
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "biahub" or mod_name.startswith("biahub.")):
            continue
        for attr, value in list(vars(mod).items()):
            new = replaced.get(id(value))
            if new is not None and value is not new:
                _saved.setdefault((mod_name, attr), value)
                setattr(mod, attr, new)
                done.setdefault(mod_name, []).append(attr)
    return done


def install(strict: bool = False):
    """Patch every importable reference module; returns {module: [patched names]}."""
    for name in _IMPORTERS:
        try:
            importlib.import_module(name)
        except Exception:
            if strict:
                raise
    patched = {}
    replaced = {}
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
                orig = getattr(ref_mod, name)
                new = getattr(ours, name)
                if orig is new:
                    continue
                _saved.setdefault((ref_name, name), orig)
                replaced[id(orig)] = new
                setattr(ref_mod, name, new)
                patched.setdefault(ref_name, []).append(name)
    for mod_name, names in _rebind(replaced).items():
        patched.setdefault(mod_name, []).extend(n for n in names if n not in patched.get(mod_name, []))
    try:
        tmod = importlib.import_module("biahub.core.transform")
        cls = getattr(tmod, "Transform", None)
        if cls is not None and hasattr(cls, "to_ants") and cls.to_ants is not _transform_to_ants:
            _saved_methods.setdefault(("biahub.core.transform", "Transform", "to_ants"), cls.to_ants)
            cls.to_ants = _transform_to_ants
            patched.setdefault("biahub.core.transform", []).append("Transform.to_ants")
    except Exception:
        if strict:
            raise
    return patched


def uninstall():
    for (ref_name, name), fn in list(_saved.items()):
        mod = sys.modules.get(ref_name) or importlib.import_module(ref_name)
        setattr(mod, name, fn)
        del _saved[(ref_name, name)]
    for (mod_name, cls_name, name), fn in list(_saved_methods.items()):
        setattr(getattr(importlib.import_module(mod_name), cls_name), name, fn)
        del _saved_methods[(mod_name, cls_name, name)]
