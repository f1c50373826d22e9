import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
import bench
from test_gpu_fullsize import _grid, DEV
sv, g = _grid("c2")
with torch.no_grad():
    cam = sv.Camera(bench.bench_cameras(DEV)[0], fx=300.0, width=160, height=120)
    img = g.volume_render_image(cam).view(-1, 3)
    rays = cam.gen_rays()
    ref = g.volume_render(rays)
    diff = (img - ref).abs().max(1)[0]
    print("image vs rays: max", diff.max().item(), "frac>1e-3", (diff > 1e-3).float().mean().item(), "frac>1e-5", (diff > 1e-5).float().mean().item())
    bad = torch.nonzero(diff > 1e-3).flatten()[:5]
    print("bad pixels", bad.tolist(), img[bad], ref[bad])
    print("dirs", rays.dirs[bad], rays.origins[bad])
n = g.links.numel()
dense = torch.zeros_like(g.density_data.data)
g.inplace_tv_grad(dense, scaling=0.37, sparse_frac=1.0, contiguous=True)
g.sparse_grad_indexer = None
sparse = torch.zeros_like(g.density_data.data)
g.inplace_tv_grad(sparse, scaling=0.37, _cell_run=(12345, n), _scale_cells=n)
d = (dense - sparse).abs().flatten()
print("tv: max dense", dense.abs().max().item(), "max sparse", sparse.abs().max().item(), "max diff", d.max().item(), "frac", (d > 1e-5 * dense.abs().max()).float().mean().item())
i = torch.argmax(d); print("worst row", i.item(), dense.flatten()[i].item(), sparse.flatten()[i].item())
print("ratio", (sparse.flatten()[dense.flatten()!=0][:10] / dense.flatten()[dense.flatten()!=0][:10]).tolist())
