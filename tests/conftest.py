import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    here = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(here, "golden_v1.npz"))
    with open(os.path.join(here, "golden_v1.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def golden_affine():
    """Outputs of the reference's OWN register/stabilize functions (tests/golden/make_golden_affine.py)."""
    import json

    import numpy as np

    here = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(here, "golden_affine_v1.npz"))
    with open(os.path.join(here, "golden_affine_v1.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def built_lib():
    from biahub_b200 import _build

    return _build.build()
