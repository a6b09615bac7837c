"""hrbffusion3d_b200 -- B200-native (sm_100a) implementation of HRBFFusion3D's per-frame hot path.

Host-side mirrors of the reference classes (RGBDOdometry, IndexMap, GlobalModel, HRBFFusion) over
the C ABI in include/hrbf_b200.h.  PyTorch is used for device memory and streams only.
"""
from ._lib import HrbfError, lib, LIB_PATH  # noqa: F401

__all__ = ["HrbfError", "lib", "LIB_PATH"]
