"""Generate golden vectors from the UNMODIFIED reference (sxyu/svox2) pure-PyTorch CPU path.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Writes tests/golden/*.npz (committed).  The reference has no stored goldens of its own
(SURVEY.md 8c); these vectors come from its own code: svox2/svox2.py:658-884
(_volume_render_gradcheck_lerp) + autograd, :603-653 (sample fallback), :414-447 (links
construction), utils.py:46-79 (Morton), :115-162 (eval_sh_bases), svox2.py:157-183 (gen_rays).
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("SVOX2_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")
import svox2  # noqa: E402  (the reference; _C is None here)

OUT = os.path.dirname(os.path.abspath(__file__))


def rays_into_grid(n, radius, center, gen, spread=0.5, dist=2.7):
    v = torch.randn(n, 3, generator=gen)
    origins = F.normalize(v, dim=-1) * dist + torch.tensor(center)
    targets = (torch.rand(n, 3, generator=gen) - 0.5) * 2 * spread * torch.tensor(radius) + torch.tensor(center)
    dirs = F.normalize(targets - origins, dim=-1)
    return origins.float(), dirs.float()


def render_case(name, reso, basis_dim, radius, center, sphere, n_rays, step, bgb, seed, near_clip=0.0,
                backend="cuvol", dens_scale=3.0, zero_frac=0.3):
    """backend="nvol": the reference's _volume_render_gradcheck_nvol_lerp (svox2.py:886-1034), the
    path its own test/test_render_gradcheck.py compares the nvol kernels with.  nvol fixtures keep
    every stored density positive (zero_frac=0): at a sample whose sigma is exactly 0 the two
    reference implementations disagree (PyTorch: relu'(0) = 0; the CUDA remat pass,
    render_lerp_kernel_nvol.cu:248-293, differentiates sub-threshold samples too), and the C
    oracle follows the CUDA kernels."""
    gen = torch.Generator().manual_seed(seed)
    grid = svox2.SparseGrid(reso=reso, radius=radius, center=center, basis_dim=basis_dim,
                            use_z_order=False, use_sphere_bound=sphere, device="cpu")
    N = grid.density_data.shape[0]
    with torch.no_grad():
        grid.sh_data[:] = torch.randn(N, 3 * basis_dim, generator=gen) * 0.3
        grid.sh_data[:, ::basis_dim] += 0.3
        dens = torch.rand(N, 1, generator=gen) * dens_scale
        dens[torch.rand(N, 1, generator=gen) < zero_frac] = 0.0
        grid.density_data[:] = dens
    grid.opt.step_size = step
    grid.opt.background_brightness = bgb
    grid.opt.sigma_thresh = 0.0
    grid.opt.stop_thresh = 0.0
    grid.opt.near_clip = near_clip
    grid.opt.backend = backend
    r = [radius] * 3 if isinstance(radius, float) else radius
    origins, dirs = rays_into_grid(n_rays, r, center, gen)
    rgb_gt = torch.rand(n_rays, 3, generator=gen)
    rays = svox2.Rays(origins, dirs)
    rgb = grid.volume_render(rays, use_kernel=False)
    loss = F.mse_loss(rgb, rgb_gt)
    loss.backward()
    gsz = torch.tensor(grid.links.shape, dtype=torch.float32)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        links=grid.links.numpy(), density=grid.density_data.detach().numpy(),
        sh=grid.sh_data.detach().numpy(),
        offset=(grid._offset * gsz - 0.5).numpy(), scaling=(grid._scaling * gsz).numpy(),
        origins=origins.numpy(), dirs=dirs.numpy(), rgb_gt=rgb_gt.numpy(),
        step_size=np.float32(step), background_brightness=np.float32(bgb),
        near_clip=np.float32(near_clip),
        rgb=rgb.detach().numpy(), grad_density=grid.density_data.grad.numpy(),
        grad_sh=grid.sh_data.grad.numpy(), loss=np.float32(loss.item()),
        radius=np.array(r, np.float32), center=np.array(center, np.float32),
    )
    print(name, "N", N, "loss", loss.item())


def sample_case():
    gen = torch.Generator().manual_seed(11)
    grid = svox2.SparseGrid(reso=[12, 10, 14], radius=[1.0, 0.9, 1.1], center=[0.1, 0.0, -0.1],
                            basis_dim=4, use_sphere_bound=True, device="cpu")
    N = grid.density_data.shape[0]
    with torch.no_grad():
        grid.sh_data[:] = torch.randn(N, 12, generator=gen)
        grid.density_data[:] = torch.rand(N, 1, generator=gen)
    pts = (torch.rand(500, 3, generator=gen) - 0.5) * 2.6
    sig, col = grid.sample(pts.clone(), use_kernel=False)
    gs = torch.randn(sig.shape, generator=gen)
    gc = torch.randn(col.shape, generator=gen)
    (sig * gs).sum().backward(retain_graph=True)
    gd_only = grid.density_data.grad.clone()
    grid.density_data.grad = None
    ((sig * gs).sum() + (col * gc).sum()).backward()
    gsz = torch.tensor(grid.links.shape, dtype=torch.float32)
    np.savez_compressed(
        os.path.join(OUT, "sample_sphere_12x10x14.npz"),
        links=grid.links.numpy(), density=grid.density_data.detach().numpy(),
        sh=grid.sh_data.detach().numpy(),
        offset=(grid._offset * gsz - 0.5).numpy(), scaling=(grid._scaling * gsz).numpy(),
        points=pts.numpy(), out_density=sig.detach().numpy(), out_sh=col.detach().numpy(),
        grad_out_density=gs.numpy(), grad_out_sh=gc.numpy(),
        grad_density=grid.density_data.grad.numpy(), grad_sh=grid.sh_data.grad.numpy(),
        grad_density_only=gd_only.numpy(),
    )
    print("sample case N", N)


def structure_case():
    out = {}
    for reso, zo, sph in [(16, True, True), (16, False, True), ([8, 12, 10], False, True),
                          (8, True, False)]:
        grid = svox2.SparseGrid(reso=reso, use_z_order=zo, use_sphere_bound=sph, device="cpu")
        key = f"links_{reso}_{int(zo)}_{int(sph)}".replace(" ", "")
        out[key] = grid.links.numpy()
        out[key + "_cap"] = np.int64(int(grid.capacity))
    out["morton8"] = svox2.utils.gen_morton(8).numpy()
    gen = torch.Generator().manual_seed(5)
    d = F.normalize(torch.randn(64, 3, generator=gen), dim=-1)
    out["sh_dirs"] = d.numpy()
    for bd in (1, 4, 9):
        out[f"sh_bases_{bd}"] = svox2.utils.eval_sh_bases(bd, d).numpy()
    c2w = torch.tensor([[0.8, -0.36, 0.48, 1.0], [0.6, 0.48, -0.64, -2.0], [0.0, 0.8, 0.6, 1.5],
                        [0, 0, 0, 1.0]])
    cam = svox2.Camera(c2w, fx=40.0, fy=42.0, cx=15.5, cy=11.0, width=32, height=24)
    r = cam.gen_rays()
    out["cam_c2w"] = c2w.numpy()
    out["cam_origins"] = r.origins.numpy()
    out["cam_dirs"] = r.dirs.numpy()
    cam2 = svox2.Camera(c2w, fx=40.0, fy=42.0, cx=15.5, cy=11.0, width=32, height=24,
                        ndc_coeffs=(2 * 40.0 / 32, 2 * 42.0 / 24))
    c2w2 = c2w.clone()
    c2w2[:3, :3] = torch.tensor([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]])
    c2w2[:3, 3] = torch.tensor([0.1, -0.05, 0.2])
    cam2.c2w = c2w2
    r2 = cam2.gen_rays()
    out["ndc_c2w"] = c2w2.numpy()
    out["ndc_origins"] = r2.origins.numpy()
    out["ndc_dirs"] = r2.dirs.numpy()
    np.savez_compressed(os.path.join(OUT, "structure.npz"), **out)
    print("structure case", list(out.keys())[:6], "...")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "nvol":  # only the nvol fixtures (added in round 2)
        render_case("nvol_dense14_sh4", 14, 4, 1.0, [0.0, 0.0, 0.0], False, 80, 0.5, 0.8, 4,
                    backend="nvol", dens_scale=0.9, zero_frac=0.0)
        render_case("nvol_sphere_16x12x14_sh9", [16, 12, 14], 9, [1.0, 0.8, 0.9], [0.0, 0.05, 0.0],
                    True, 64, 0.43, 1.0, 5, backend="nvol", dens_scale=1.2, zero_frac=0.0)
        sys.exit(0)
    render_case("render_dense16_sh9", 16, 9, 1.0, [0.0, 0.0, 0.0], False, 96, 0.5, 1.0, 1)
    render_case("render_sphere_20x16x12_sh4", [20, 16, 12], 4, [1.0, 0.8, 0.6], [0.1, -0.05, 0.0],
                True, 64, 0.5, 0.5, 2)
    render_case("render_dense12_sh1_step037", 12, 1, 1.3, [0.0, 0.0, 0.0], False, 48, 0.37, 0.0, 3)
    # nvol backend: densities chosen so that part of the rays saturate (alpha_total reaches 1)
    render_case("nvol_dense14_sh4", 14, 4, 1.0, [0.0, 0.0, 0.0], False, 80, 0.5, 0.8, 4,
                backend="nvol", dens_scale=0.9, zero_frac=0.0)
    render_case("nvol_sphere_16x12x14_sh9", [16, 12, 14], 9, [1.0, 0.8, 0.9], [0.0, 0.05, 0.0], True,
                64, 0.43, 1.0, 5, backend="nvol", dens_scale=1.2, zero_frac=0.0)
    sample_case()
    structure_case()
