"""Experiment (round 2): is a COMPACT per-step gradient buffer worth building?  On the real C3 touched
set (masks of one bench step): masked RMSProp on dense-layout gradients (today) vs RMSProp that
streams gradients from a packed [rows, 28] buffer and touches only data / rms sparsely."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
import svox2_b200 as sv
from svox2_b200 import csrc as C

dev = torch.device("cuda")
sc = B.build_scene("c3", dev)
grid = sv.SparseGrid(reso=[2, 2, 2], basis_dim=9, device=dev)
grid.links = sc["links"]
grid.density_data = torch.nn.Parameter(sc["density"])
grid.sh_data = torch.nn.Parameter(sc["sh"])
grid.capacity = sc["n_rows"]
grid.accelerate()
for k, v in B.RENDER.items():
    setattr(grid.opt, k, v)
o, d = B.build_rays(5000, dev)
with torch.no_grad():
    gt = grid.volume_render(sv.Rays(o, d))
np.random.seed(1)
grid.volume_render_fused(sv.Rays(o, d), gt + 0.02)
grid.inplace_tv_grad(grid.density_data.grad, scaling=1e-5)
grid.inplace_tv_color_grad(grid.sh_data.grad, scaling=1e-3)
md, msh = grid.sparse_grad_indexer.clone(), grid.sparse_sh_grad_indexer.clone()
grid._ensure_rms()
rows = torch.nonzero(md | msh).flatten().contiguous()
print("rows", rows.numel(), "density mask", int(md.sum()), "sh mask", int(msh.sum()))
gsh0, gd0 = grid.sh_data.grad.clone(), grid.density_data.grad.clone()
packed = torch.empty(rows.numel(), 28, device=dev)
C.pack_grad_rows(grid.sh_data.grad, grid.density_data.grad, rows, packed)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(f, reps=8, prep=None):
    tot = 0.0
    for i in range(reps + 2):
        if prep:
            prep()
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def restore():
    grid.sh_data.grad.copy_(gsh0)
    grid.density_data.grad.copy_(gd0)
    grid.sparse_grad_indexer, grid.sparse_sh_grad_indexer = md, msh


def today():
    grid.optim_density_step(30.0, beta=0.95)
    grid.optim_sh_step(1e-2, beta=0.95)


print("masked RMSProp x2 (today)        : %.1f us" % timeit(today, prep=restore))
print("  density only                   : %.1f us" % timeit(lambda: grid.optim_density_step(30.0, beta=0.95), prep=restore))
print("  sh only                        : %.1f us" % timeit(lambda: grid.optim_sh_step(1e-2, beta=0.95), prep=restore))
f = lambda: C.rmsprop_step_packed(grid.sh_data.data, grid.sh_rms, grid.density_data.data, grid.density_rms,
                                  rows, packed, msh, md, 0.95, 1e-2, 30.0, 1e-8)
print("packed RMSProp, no dense zeroing : %.1f us" % timeit(f))
print("memset of the packed buffer      : %.1f us" % timeit(lambda: packed.zero_()))
print("copy 1 GiB (peak check)          : %.1f us" % timeit(lambda: flush[:128 << 20].copy_(flush[128 << 20:])))
