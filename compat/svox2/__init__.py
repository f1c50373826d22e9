"""Drop-in alias: ``import svox2`` resolves to the B200-native implementation.

Put this directory on PYTHONPATH *instead of* the reference checkout:
    PYTHONPATH=/path/to/repo/compat:/path/to/repo python opt/opt.py ...
``svox2.SparseGrid / Camera / Rays / RenderOptions`` (svox2/__init__.py:1-3 of the reference) and
the operator module ``svox2.csrc`` keep their names.
"""
from svox2_b200 import *  # noqa: F401,F403
from svox2_b200 import csrc, utils  # noqa: F401
from svox2_b200.grid import (BASIS_TYPE_3D_TEXTURE, BASIS_TYPE_MLP, BASIS_TYPE_SH, Camera, Rays,  # noqa: F401
                             RenderOptions, SparseGrid)
from svox2_b200 import __version__  # noqa: F401
