"""GPU side of the dataset / evaluation rows: rays generated in HBM equal the
host rays, and the training driver runs the other two dataset families end to end -- an NSVF / T&T
format scene with an MSI background (tnt.json's path: fused step with background, MSI TV,
optim_background_step, sparsify_background at the upsample) and a forward-facing NDC scene
(llff.json's path)."""
import math
import os

import numpy as np
import pytest
import torch

import svox2_b200 as sv
from svox2_b200 import datasets as D
from svox2_b200 import train as T
from test_datasets_eval import _write_llff, _write_nsvf

pytestmark = pytest.mark.gpu


def test_device_rays_equal_host_rays(tmp_path):
    rng = np.random.default_rng(1)
    _write_nsvf(str(tmp_path / "nsvf"), rng)
    _write_llff(str(tmp_path / "llff"), rng)
    for loader, root, kw in ((D.load_nsvf, "nsvf", {}), (D.load_llff, "llff", dict(scale=0.5))):
        host = loader(str(tmp_path / root), "train", device="cpu", **kw)
        dev = loader(str(tmp_path / root), "train", device="cuda", **kw)
        (rh, gh), (rd, gd) = host.gen_rays(), dev.gen_rays()
        assert rd.origins.is_cuda and rd.origins.shape == rh.origins.shape
        assert (rd.origins.cpu() - rh.origins).abs().max().item() < 2e-6
        assert (rd.dirs.cpu() - rh.dirs).abs().max().item() < 2e-6
        assert torch.equal(gd.cpu(), gh)


def _save_nsvf(data, root, prefix_of):
    from PIL import Image
    os.makedirs(os.path.join(root, "rgb"))
    os.makedirs(os.path.join(root, "pose"))
    for i in range(data.c2w.size(0)):
        name = f"{prefix_of(i)}_{i:04d}"
        np.savetxt(os.path.join(root, "pose", name + ".txt"), data.c2w[i].double().cpu().numpy())
        im = (data.images[i].clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).cpu().numpy()
        Image.fromarray(im).save(os.path.join(root, "rgb", name + ".png"))
    K = np.eye(4)
    K[0, 0] = K[1, 1] = data.fx
    K[0, 2], K[1, 2] = data.width * 0.5, data.height * 0.5
    np.savetxt(os.path.join(root, "intrinsics.txt"), K)


def test_trainer_on_nsvf_format_scene_with_msi_background(tmp_path):
    dev = torch.device("cuda:0")
    data = T.synthetic_dataset(dev, n_images=22, width=72, height=72, reso=48)
    root = str(tmp_path / "tnt_like")
    _save_nsvf(data, root, lambda i: "0" if i < 18 else "1")
    train, test = D.auto_dataset(root, "train", dev), D.auto_dataset(root, "test", dev)
    assert train.n_images == 18 and test.n_images == 4 and train.should_use_background
    # cameras were normalised to a median distance of cam_scale_factor: they sit INSIDE the grid's
    # bounding sphere region as in T&T scenes, so foreground and background both matter
    assert abs(train.c2w[:, :3, 3].norm(dim=-1).median().item() - 0.95) < 0.05
    cfg = T.TrainConfig(reso="[[32,32,32],[64,64,64]]", upsamp_every=90, n_iters=220, batch_size=2048,
                        lr_sigma_delay_steps=40, print_every=1000, weight_thresh=0.0005 * 64,
                        background_nlayers=8, background_reso=64, lambda_tv_background_sigma=1e-3,
                        lambda_tv_background_color=1e-3, lambda_beta=1e-5, lambda_sparsity=1e-11,
                        tv_early_only=0, lambda_tv=5e-5, lambda_tv_sh=5e-3)
    tr = T.Trainer(cfg, train, test, log=lambda m: None)
    assert tr.grid.use_background and tr.grid.background_data.size(1) == 8
    before = tr.eval_psnr()
    n_bg0 = tr.grid.background_data.size(0)
    tr.run()
    res = tr.eval_metrics(log=None)
    assert list(tr.grid.links.shape) == [64, 64, 64] and tr.reso_id == 1
    assert tr.grid.background_data.size(0) <= n_bg0          # sparsify_background ran at the upsample
    assert res["psnr"] > before + 3.0 and res["psnr"] > 14.0, (before, res["psnr"])
    assert 0.0 < res["ssim"] <= 1.0 and res["n_images"] == 4
    assert torch.isfinite(tr.grid.background_data).all()


def test_trainer_on_forward_facing_ndc_scene():
    """llff.json's path: non-cubic NDC grid, sigma-threshold resampling, last_sample_opaque off,
    grey background; images come from a hidden NDC grid rendered with the NDC image kernel."""
    dev = torch.device("cuda:0")
    W, H, f = 64, 48, 52.0
    ndc = (2 * f / W, 2 * f / H)
    radius = [1.0 + 2 * 10 / W, 1.0 + 2 * 10 / H, 1.0]
    hidden = sv.SparseGrid(reso=[48, 40, 32], radius=radius, basis_dim=1, device=dev)
    gen = torch.Generator().manual_seed(5)
    ax = [(torch.arange(n, dtype=torch.float32) + 0.5) / n * 2 - 1 for n in (48, 40, 32)]
    X, Y, Z = torch.meshgrid(*ax, indexing="ij")
    sigma, col = torch.zeros(48, 40, 32), torch.zeros(48, 40, 32, 3)
    for _ in range(5):
        c = (torch.rand(3, generator=gen) - 0.5) * torch.tensor([1.2, 1.2, 1.0])
        r = torch.rand(3, generator=gen) * 0.2 + 0.15
        inside = ((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2 < 1
        sigma[inside] = 40.0
        col[inside] = torch.rand(3, generator=gen)
    with torch.no_grad():
        hidden.density_data.data[:] = sigma.reshape(-1, 1).to(dev)
        hidden.sh_data.data[:] = ((col.reshape(-1, 3) - 0.5) / 0.28209479177387814).to(dev)
    hidden.opt.background_brightness = 0.5
    n = 14
    c2w = torch.eye(4).repeat(n, 1, 1)
    c2w[:, 1, 1] = c2w[:, 2, 2] = -1.0                       # looking down -z (LLFF convention)
    c2w[:, 0, 3] = torch.linspace(-0.25, 0.25, n)
    c2w[:, 1, 3] = 0.1 * torch.sin(torch.linspace(0, 6.0, n))
    c2w = c2w.to(dev)
    with torch.no_grad():
        imgs = torch.stack([hidden.volume_render_image(
            sv.Camera(c2w[i], f, f, W * 0.5, H * 0.5, W, H, ndc_coeffs=ndc)) for i in range(n)])
    assert imgs.std().item() > 0.05
    mk = lambda sl: D.PosedImageSet(c2w[sl], imgs[sl].clamp(0, 1), D.Intrin(f, f, W * 0.5, H * 0.5),
                                    ndc_coeffs=ndc, scene_radius=radius, use_sphere_bound=False,
                                    should_use_background=False)
    train, test = mk(slice(0, 12)), mk(slice(12, 14))
    cfg = T.TrainConfig(reso="[[24,20,16],[48,40,32]]", upsamp_every=100, n_iters=260, batch_size=2048,
                        lr_sigma_delay_steps=0, print_every=1000, thresh_type="sigma", density_thresh=1.0,
                        lambda_tv=5e-4, lambda_tv_sh=5e-3, lambda_sparsity=1e-12,
                        background_brightness=0.5, tv_early_only=0, background_nlayers=64)
    tr = T.Trainer(cfg, train, test, log=lambda m: None)
    assert not tr.grid.use_background and list(tr.grid.links.shape) == [24, 20, 16]
    before = tr.eval_psnr()
    tr.run()
    after = tr.eval_metrics(log=None)
    assert list(tr.grid.links.shape) == [48, 40, 32]
    assert after["psnr"] > before + 2.0 and math.isfinite(after["ssim"]), (before, after["psnr"])
