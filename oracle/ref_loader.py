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

``load_reference_register_stabilize`` does the same for ``biahub/register.py`` and
``biahub/stabilize.py`` with a fake ``ants`` / ``largestinteriorrectangle`` (see below), so that
the reference's own wrapper code around the ANTs call is executed and pinned.

Recipe: SURVEY.md Appendix B.  Reference import list: biahub/deskew.py:1-36.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BIAHUB_REFERENCE_ROOT", "/root/reference")
# Staged copy of the same files for the GPU box (scripts/make_baseline_ref.py; git-ignored)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def use_staged_reference() -> bool:
    """Point the loader at ``baseline/_ref`` (what ``bench.py`` does, so that the same files are
    timed here and on the GPU box).  Returns False when nothing has been staged."""
    global REFERENCE_ROOT, _LOADED, _REG
    if not os.path.isfile(os.path.join(STAGED_ROOT, "biahub", "deskew.py")):
        return False
    if REFERENCE_ROOT != STAGED_ROOT:
        REFERENCE_ROOT = STAGED_ROOT
        _LOADED = None
        _REG = None
        for name in [n for n in sys.modules if n == "biahub" or n.startswith("biahub.")]:
            del sys.modules[name]
    return True


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
_REG = None


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


# ------------------------------------------------------------------------------------------
# register / stabilize wrappers (reference biahub/register.py:32-398, biahub/stabilize.py:32-90)
# ------------------------------------------------------------------------------------------
# The arithmetic of ``method="ants"`` lives in antspyx 0.6.1 (uv.lock:196-197; ITK
# ResampleImageFilter), which is not installable here.  What CAN be executed is everything the
# reference itself wrote around that call: NaN scrub order, the float32 cast, the 4-D channel
# loop, crop slicing, output shapes, the ITK parameter packing of ``convert_transform_to_ants``
# / ``convert_transform_to_numpy``, the matrix builders, ``rescale_voxel_size``, ``find_lir``
# and ``find_overlapping_volume``.  ``FakeAnts`` below is the smallest ``ants`` module those
# functions need; its ``apply_to_image`` is the ITK-rule restatement of
# ``oracle/affine_oracle.py`` (float64 coordinates — ANTs' own float32 transform precision is
# deliberately not mimicked).  ``method="scipy"`` needs no fake: the reference's
# literal ``scipy.ndimage.affine_transform(zyx_data, matrix, output_shape_zyx)`` runs as is.


class FakeAntsImage:
    """``ants.from_numpy`` result: ``.numpy()``, ``.shape``, ``.dimension``."""

    def __init__(self, array):
        import numpy as np

        self._array = np.array(array)  # ants copies
        self.shape = self._array.shape
        self.dimension = self._array.ndim

    def numpy(self):
        return self._array.copy()


class FakeAntsTransform:
    """``ants.new_ants_transform(transform_type="AffineTransform")``: 12 parameters (row-major
    3x3 then translation) + 3 fixed parameters (centre), ITK's ``y = A (x - c) + c + t``."""

    def __init__(self, transform_type="AffineTransform", dimension=3, precision="float",
                 parameters=None, fixed_parameters=None, **_):
        import numpy as np

        if transform_type != "AffineTransform" or dimension != 3:
            raise NotImplementedError("the fake ants module only models 3-D AffineTransform")
        self.transform_type = transform_type
        self.dimension = dimension
        self.precision = precision
        self._parameters = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
        self._fixed = np.zeros(3, dtype=np.float64)
        if parameters is not None:
            self.set_parameters(parameters)
        if fixed_parameters is not None:
            self.set_fixed_parameters(fixed_parameters)

    def set_parameters(self, parameters):
        import numpy as np

        p = np.asarray(parameters, dtype=np.float64).ravel()
        assert p.size == 12, "AffineTransform has 12 parameters"
        self._parameters = p.copy()

    # ants returns a fresh array per access (reference register.py:187 reshapes and writes into it)
    @property
    def parameters(self):
        return self._parameters.copy()

    @property
    def fixed_parameters(self):
        return self._fixed.copy()

    def set_fixed_parameters(self, fixed_parameters):
        import numpy as np

        self._fixed = np.asarray(fixed_parameters, dtype=np.float64).ravel().copy()

    def pull_matrix(self):
        import numpy as np

        A = self._parameters[:9].reshape(3, 3)
        M = np.eye(4)
        M[:3, :3] = A
        M[:3, 3] = self._parameters[9:] + (np.eye(3) - A) @ self._fixed
        return M

    def apply_to_image(self, image, reference=None, interpolation="linear"):
        from oracle import affine_oracle as ao

        order = {"linear": 1, "nearestneighbor": 0}.get(str(interpolation).lower())
        if order is None:
            raise NotImplementedError(f"fake ants: interpolation {interpolation!r} is not modelled")
        shape = image.shape if reference is None else reference.shape
        out = ao.affine_oracle_numpy(image.numpy(), self.pull_matrix(), tuple(shape), order, "itk")
        return FakeAntsImage(out)


def fake_ants_module():
    mod = types.ModuleType("ants")
    mod.from_numpy = FakeAntsImage
    mod.new_ants_transform = FakeAntsTransform
    mod.ANTsImage = FakeAntsImage
    mod.ANTsTransform = FakeAntsTransform
    mod.__fake__ = True
    return mod


def fake_lir_module():
    """``largestinteriorrectangle`` 0.2.1 (uv.lock; numba, absent here): ``lir(mask) ->
    (x, y, width, height)`` of the largest axis-aligned all-True rectangle.  Restated
    independently of the product (all heights per column, all row spans: O(H^2 W))."""
    import numpy as np

    def lir(mask, contour=None):
        m = np.asarray(mask, dtype=bool)
        H, W = m.shape
        best, best_area = (0, 0, 0, 0), 0
        for top in range(H):
            col_ok = np.ones(W, dtype=bool)
            for bottom in range(top, H):
                col_ok &= m[bottom]
                if not col_ok.any():
                    break
                h = bottom - top + 1
                # longest run of True in col_ok
                padded = np.concatenate([[False], col_ok, [False]]).astype(np.int8)
                d = np.diff(padded)
                starts, stops = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
                k = int(np.argmax(stops - starts))
                w = int(stops[k] - starts[k])
                if w * h > best_area:
                    best_area, best = w * h, (int(starts[k]), top, w, h)
        return np.array(best)

    mod = types.ModuleType("largestinteriorrectangle")
    mod.lir = lir
    mod.__fake__ = True
    return mod


def load_reference_register_stabilize():
    """Return the reference's ``(biahub.register, biahub.stabilize)`` modules, imported from
    their files with the fake ``ants`` / ``largestinteriorrectangle`` above and inert
    matplotlib / humanize stand-ins.  Recipe: SURVEY.md Appendix B."""
    global _REG
    if _REG is not None:
        return _REG
    load_reference_deskew()  # iohub / submitit / monai / natsort stand-ins
    for name in ("matplotlib", "matplotlib.pyplot", "humanize"):
        if name not in sys.modules:
            _stub(name)
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib.pyplot"], "subplots"):
        sys.modules["matplotlib.pyplot"].subplots = _Sink()
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["humanize"], "naturalsize"):
        sys.modules["humanize"].naturalsize = lambda n, *a, **k: f"{n} B"
    saved = {k: sys.modules.get(k) for k in ("ants", "largestinteriorrectangle", "biahub")}
    sys.modules["ants"] = fake_ants_module()
    sys.modules["largestinteriorrectangle"] = fake_lir_module()
    pkg = types.ModuleType("biahub")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "biahub")]
    sys.modules["biahub"] = pkg
    try:
        reg = importlib.import_module("biahub.register")
        stab = importlib.import_module("biahub.stabilize")
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _REG = (reg, stab)
    return _REG


# ------------------------------------------------------------------------------------------
# estimation-loop modules: only their import-time name bindings matter
# here (tests/test_patch.py checks that patch.install() re-points them), so every third-party
# module they import and this container lacks (napari, dask, skimage, ...) becomes an inert stub.
# ------------------------------------------------------------------------------------------
class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Sink()


# third-party packages the estimation modules import that this container lacks
_ABSENT_THIRD_PARTY = ("napari", "dask", "skimage", "waveorder", "cmap", "imageio", "tifffile",
                       "ultrack", "pystackreg", "stitch", "xarray", "zarr", "numcodecs",
                       "sklearn_extra", "cellpose", "toolz", "numba", "llvmlite", "seaborn",
                       "networkx", "ome_zarr", "tensorstore", "psutil_extra", "colorspacious")


class _StubFinder:
    """meta_path finder of last resort: the known-absent third-party packages above (and their
    sub-modules) become inert stubs.  Never shadows ``biahub`` or anything installed."""

    def __init__(self):
        self.made = []

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery

        top = fullname.split(".")[0]
        if top not in _ABSENT_THIRD_PARTY:
            return None
        self.made.append(fullname)
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def load_reference_estimation_modules():
    """Import the reference modules whose warps the estimation loops run (biahub/
    optimize_registration.py, registration/ants.py, registration/beads.py,
    estimate_registration.py, core/transform.py) into ``sys.modules`` with the fake ants and
    auto-stubbed third parties; returns {name: module}.  The reference package stays installed
    as ``biahub`` in ``sys.modules`` until ``unload_reference_package()``."""
    reg, stab = load_reference_register_stabilize()
    sys.modules["ants"] = fake_ants_module()
    sys.modules["largestinteriorrectangle"] = fake_lir_module()
    pkg = types.ModuleType("biahub")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "biahub")]
    sys.modules["biahub"] = pkg
    sys.modules["biahub.register"] = reg
    sys.modules["biahub.stabilize"] = stab
    finder = _StubFinder()
    sys.meta_path.append(finder)
    out = {"biahub.register": reg, "biahub.stabilize": stab}
    try:
        for name in ("biahub.core.transform", "biahub.registration.utils",
                     "biahub.optimize_registration", "biahub.registration.ants",
                     "biahub.registration.beads", "biahub.estimate_registration"):
            out[name] = importlib.import_module(name)
    finally:
        sys.meta_path.remove(finder)
    return out


def unload_reference_package():
    """Drop the reference ``biahub`` package (and the fakes) from ``sys.modules``."""
    global _LOADED, _REG
    for name in [n for n in sys.modules if n == "biahub" or n.startswith("biahub.")]:
        del sys.modules[name]
    for name in ("ants", "largestinteriorrectangle"):
        if getattr(sys.modules.get(name), "__fake__", False):
            del sys.modules[name]
    for name in [n for n, m in sys.modules.items() if isinstance(m, _StubModule)]:
        del sys.modules[name]
    _LOADED = None
    _REG = None
