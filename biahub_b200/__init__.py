"""biahub_b200 — B200-native (sm_100a) 3-D affine resampling path of czbiohub-sf/biahub.

Drop-in replacements (same names/signatures as the reference) for the array-compute layer of
``biahub deskew`` / ``biahub register`` / ``biahub stabilize``:

    biahub_b200.deskew     ↔ reference biahub/deskew.py:43-579
    biahub_b200.register   ↔ reference biahub/register.py:32-281, 397-398
    biahub_b200.stabilize  ↔ reference biahub/stabilize.py:32-90
    biahub_b200.flat_field ↔ reference biahub/flat_field.py:105-166 (the stage before deskew)

All arithmetic runs in ``_lib/libbiahub_b200.so`` (hand-written CUDA, C ABI in
``include/biahub_b200.h``); there is no CPU or PyTorch fallback.  ``patch.install()`` re-points
an importable ``biahub`` at these functions.
"""

from . import _cabi  # noqa: F401
from .deskew import (  # noqa: F401
    _average_n_slices,
    _average_n_slices_torch,
    _deskew_czyx,
    _fast_deskew_czyx,
    _get_averaged_shape,
    _get_transform_matrix,
    deskew_zyx,
    fast_deskew_zyx,
    get_deskewed_data_shape,
)
from .register import (  # noqa: F401
    affine_warp,
    apply_affine_transform,
    convert_transform_to_ants,
    convert_transform_to_numpy,
    find_lir,
    find_overlapping_volume,
    get_3D_fliplr_matrix,
    get_3D_rescaling_matrix,
    get_3D_rotation_matrix,
    rescale_voxel_size,
    spline_warp,
)
from .flat_field import _flat_field_czyx, flat_field_correction, flat_field_zyx  # noqa: F401
from .pipeline import deskew_then_register, flatfield_then_deskew  # noqa: F401
from .stabilize import apply_stabilization_transform  # noqa: F401

__version__ = "0.1.0"
