"""``svox2.csrc``: same names as the reference's pybind module (svox2/csrc/svox2.cpp:74-172)."""
from svox2_b200.csrc import *  # noqa: F401,F403
from svox2_b200.csrc import __all__  # noqa: F401
