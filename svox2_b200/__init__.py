"""svox2_b200: B200-native (sm_100a) implementation of the Plenoxels hot path.

Public surface mirrors ``svox2`` of sxyu/svox2 (svox2/__init__.py:1-3):
``SparseGrid``, ``Camera``, ``Rays``, ``RenderOptions`` plus the operator module ``csrc``.
"""
from .grid import (BASIS_TYPE_3D_TEXTURE, BASIS_TYPE_MLP, BASIS_TYPE_SH, Camera, Rays,
                   RenderOptions, SparseGrid)
from . import csrc, utils

__version__ = "0.1.0"
__all__ = ["SparseGrid", "Camera", "Rays", "RenderOptions", "csrc", "utils",
           "BASIS_TYPE_SH", "BASIS_TYPE_3D_TEXTURE", "BASIS_TYPE_MLP"]
