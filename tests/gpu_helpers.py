"""Builders that put a scenes.make_scene() dict on the GPU for our library and for the reference."""
import numpy as np
import torch

import svox2_b200 as sv
from svox2_b200 import csrc as C


def t(a, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def our_spec(sc, device="cuda"):
    spec = C.SparseGridSpec()
    spec.density_data = t(sc["density"], device)
    spec.sh_data = t(sc["sh"], device)
    spec.links = t(sc["links"], device)
    spec._offset = torch.from_numpy(sc["offset"].copy())
    spec._scaling = torch.from_numpy(sc["scaling"].copy())
    spec.basis_dim = sc["basis_dim"]
    spec.basis_type = 1
    if "bg_links" in sc:
        spec.background_links = t(sc["bg_links"], device)
        spec.background_data = t(sc["bg_data"], device)
    return spec


def our_rays(sc, device="cuda"):
    r = C.RaysSpec()
    r.origins = t(sc["origins"], device)
    r.dirs = t(sc["dirs"], device)
    return r


def our_opt(**kw):
    o = C.RenderOptions()
    for k, v in kw.items():
        assert hasattr(o, k), k
        setattr(o, k, v)
    return o


def our_grads(spec, with_mask=True):
    g = C.GridOutputGrads()
    g.grad_density_out = torch.zeros_like(spec.density_data)
    g.grad_sh_out = torch.zeros_like(spec.sh_data)
    if with_mask:
        g.mask_out = torch.zeros(spec.density_data.size(0), dtype=torch.bool, device=spec.sh_data.device)
    if spec.background_data is not None:
        g.grad_background_out = torch.zeros_like(spec.background_data)
        if with_mask:
            g.mask_background_out = torch.zeros(spec.background_data.shape[:2], dtype=torch.bool,
                                                device=spec.sh_data.device)
    return g


def _fill_grid(grid, sc, device):
    from torch import nn
    grid.links = t(sc["links"], device)
    grid.density_data = nn.Parameter(t(sc["density"], device))
    grid.sh_data = nn.Parameter(t(sc["sh"], device))
    grid.capacity = sc["density"].shape[0]
    if "bg_links" in sc:
        grid.background_links = t(sc["bg_links"], device)
        grid.background_data = nn.Parameter(t(sc["bg_data"], device))
    return grid


def make_grid(module, sc, device="cuda"):
    """module = svox2_b200 or the reference svox2 package; same constructor signature."""
    kw = {}
    if "bg_links" in sc:
        kw = dict(background_nlayers=sc["bg_data"].shape[1], background_reso=sc["bg_links"].shape[1])
    grid = module.SparseGrid(reso=list(sc["links"].shape), radius=sc["radius"].tolist(),
                             center=sc["center"].tolist(), basis_dim=sc["basis_dim"],
                             device=device, **kw)
    return _fill_grid(grid, sc, device)


def set_opt(grid, **kw):
    for k, v in kw.items():
        assert hasattr(grid.opt, k), k
        setattr(grid.opt, k, v)
