"""``svox2.utils``: importable as a submodule (``from svox2.utils import convert_to_ndc``,
``import svox2.utils``), as the reference's opt/util/llff_dataset.py and opt/render_imgs*.py do."""
from svox2_b200.utils import *  # noqa: F401,F403
from svox2_b200.utils import (MAX_SH_BASIS, SH_C0, SH_C1, SH_C2, Timing, _get_c_extension,  # noqa: F401
                              convert_to_ndc, eval_sh_bases, gen_morton, is_pow2, isqrt,
                              morton_code_3, xyz2equirect)
