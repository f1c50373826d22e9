"""Render a few 800x800 images of the C3 bench scene (for ncu: --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import svox2_b200 as sv

dev = torch.device("cuda:0")
sc = bench.build_scene("c3", dev)
g = sv.SparseGrid(reso=sc["reso"], basis_dim=9, device=dev)
g.links = sc["links"]
g.density_data = torch.nn.Parameter(sc["density"])
g.sh_data = torch.nn.Parameter(sc["sh"])
g.capacity = sc["n_rows"]
g.accelerate()
for k, v in bench.RENDER.items():
    setattr(g.opt, k, v)
cams = bench.bench_cameras(dev)
with torch.no_grad():
    for c in cams[:2]:
        g.volume_render_image(sv.Camera(c, **bench.CAM_KW))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    g.volume_render_image(sv.Camera(cams[2], **bench.CAM_KW))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
