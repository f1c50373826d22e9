"""Host-side helpers of the SparseGrid API: Z-order codes, real SH bases, NDC rays, timing.

Semantics follow the reference's ``svox2/utils.py`` (Morton :46-79, ``eval_sh_bases``
:115-162, ``convert_to_ndc`` :576-596, ``Timing`` :611-631); the implementations are our own.
"""
import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
MAX_SH_BASIS = 10


def is_pow2(x: int) -> bool:
    return x > 0 and (x & (x - 1)) == 0


def isqrt(x: int):
    """Integer square root if ``x`` is a perfect square, else None."""
    r = math.isqrt(int(x))
    return r if r * r == x else None


def _spread3(v: torch.Tensor) -> torch.Tensor:
    """Insert two zero bits between the low 10 bits of ``v`` (bit i -> bit 3i)."""
    out = torch.zeros_like(v)
    for bit in range(10):
        out |= ((v >> bit) & 1) << (3 * bit)
    return out


def morton_code_3(x, y, z):
    """Z-order code with x as the most significant axis of each bit triple."""
    return (_spread3(x) << 2) | (_spread3(y) << 1) | _spread3(z)


def gen_morton(D: int, device="cpu", dtype=torch.long) -> torch.Tensor:
    """[D,D,D] tensor whose (x,y,z) entry is the Morton index of that voxel."""
    assert is_pow2(D), "Morton code requires power of 2 reso"
    a = torch.arange(D, device=device, dtype=torch.long)
    X, Y, Z = torch.meshgrid(a, a, a, indexing="ij")
    return morton_code_3(X, Y, Z).to(dtype)


def eval_sh_bases(basis_dim: int, dirs: torch.Tensor) -> torch.Tensor:
    """Real SH basis values (degree <= 2 here: 1, 4 or 9 functions) at unit ``dirs`` (...,3)."""
    assert basis_dim in (1, 4, 9), "the B200 path supports SH basis_dim 1, 4 or 9"
    out = torch.empty((*dirs.shape[:-1], basis_dim), dtype=dirs.dtype, device=dirs.device)
    out[..., 0] = SH_C0
    if basis_dim > 1:
        x, y, z = dirs.unbind(-1)
        out[..., 1] = -SH_C1 * y
        out[..., 2] = SH_C1 * z
        out[..., 3] = -SH_C1 * x
        if basis_dim > 4:
            out[..., 4] = SH_C2[0] * (x * y)
            out[..., 5] = SH_C2[1] * (y * z)
            out[..., 6] = SH_C2[2] * (2.0 * z * z - x * x - y * y)
            out[..., 7] = SH_C2[3] * (x * z)
            out[..., 8] = SH_C2[4] * (x * x - y * y)
    return out


def convert_to_ndc(origins, directions, ndc_coeffs, near: float = 1.0):
    """Forward-facing NDC ray warp (origins shifted to the near plane first)."""
    t = (near - origins[..., 2]) / directions[..., 2]
    origins = origins + t[..., None] * directions
    dx, dy, dz = directions.unbind(-1)
    ox, oy, oz = origins.unbind(-1)
    o = torch.stack([ndc_coeffs[0] * (ox / oz), ndc_coeffs[1] * (oy / oz), 1 - 2 * near / oz], -1)
    d = torch.stack([ndc_coeffs[0] * (dx / dz - ox / oz), ndc_coeffs[1] * (dy / dz - oy / oz),
                     2 * near / oz], -1)
    return o, d


def xyz2equirect(bearings, reso):
    """Unit vectors -> equirectangular pixel coordinates ([0,2*reso) x [0,reso))."""
    lat = torch.asin(bearings[..., 1])
    lon = torch.atan2(bearings[..., 0], bearings[..., 2])
    return torch.stack([reso * 2 * (0.5 + lon / (2 * math.pi)), reso * (0.5 - lat / math.pi)], -1)


class Timing:
    """``with Timing("name"):`` prints the CUDA-event time of the block on the current stream."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.start = torch.cuda.Event(enable_timing=True)
        self.end = torch.cuda.Event(enable_timing=True)
        self.start.record()
        return self

    def __exit__(self, *exc):
        self.end.record()
        self.end.synchronize()
        self.ms = self.start.elapsed_time(self.end)
        print(self.name, "elapsed", self.ms, "ms")


def _get_c_extension():
    """The operator module.  Unlike the reference (utils.py:29-42) a missing library is an
    error, not a warning: there is no slow path to fall back to."""
    from . import _lib, csrc
    _lib.lib()
    return csrc
