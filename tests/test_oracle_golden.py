"""Pins the CPU oracle (oracle/svox2_oracle.c) against vectors produced by the reference's own
pure-PyTorch renderer / sampler (tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

import oracle as og
from scenes import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "render_*.npz"))))
def test_render_forward_and_gradients(path):
    z = np.load(path)
    g = og.Grid(z["links"], z["density"], z["sh"], z["offset"], z["scaling"])
    opt = og.make_opt(background_brightness=float(z["background_brightness"]),
                      step_size=float(z["step_size"]), sigma_thresh=0.0, stop_thresh=0.0,
                      near_clip=float(z["near_clip"]))
    rgb, _ = og.volume_render(g, opt, z["origins"], z["dirs"])
    assert np.abs(rgb - z["rgb"]).max() < 2e-6
    res = og.volume_render_fused(g, opt, z["origins"], z["dirs"], z["rgb_gt"])
    assert np.abs(res["rgb"] - z["rgb"]).max() < 2e-6
    # F.mse_loss gradient == the fused seed (rgb - gt) * 2 / (3Q)
    assert rel_err(res["grad_density"], z["grad_density"].reshape(-1)) < 1e-4
    assert rel_err(res["grad_sh"], z["grad_sh"]) < 1e-4
    # non-fused backward with the same upstream gradient
    Q = z["origins"].shape[0]
    gout = (z["rgb"] - z["rgb_gt"]) * (2.0 / (3 * Q))
    res2 = og.volume_render_backward(g, opt, z["origins"], z["dirs"], gout.astype(np.float32), z["rgb"])
    assert rel_err(res2["grad_density"], z["grad_density"].reshape(-1)) < 1e-4
    assert rel_err(res2["grad_sh"], z["grad_sh"]) < 1e-4


def test_sample_forward_backward():
    z = np.load(os.path.join(GOLD, "sample_sphere_12x10x14.npz"))
    g = og.Grid(z["links"], z["density"], z["sh"], z["offset"], z["scaling"])
    d, s = og.sample_grid(g, z["points"])
    assert np.abs(d - z["out_density"]).max() < 1e-6
    assert np.abs(s - z["out_sh"]).max() < 2e-6
    gd, gs = og.sample_grid_backward(g, z["points"], z["grad_out_density"], z["grad_out_sh"])
    assert rel_err(gd, z["grad_density"].reshape(-1)) < 1e-5
    assert rel_err(gs, z["grad_sh"]) < 1e-5


def test_accel_and_dilate_properties():
    rng = np.random.default_rng(0)
    links = np.full((16, 12, 20), -1, np.int32)
    occ = rng.random(links.shape) < 0.02
    links[occ] = np.arange(occ.sum())
    acc = og.accel_dist_prop(links)
    assert np.array_equal(acc[occ], links[occ])          # occupied voxels untouched
    assert (acc[~occ] <= -1).all()
    assert np.array_equal(og.accel_dist_prop(acc), acc)  # idempotent
    # level k => the aligned 2^k cube containing the voxel is empty
    for (x, y, z) in np.argwhere(acc <= -2)[:200]:
        k = -acc[x, y, z] - 1
        x0, y0, z0 = (x >> k) << k, (y >> k) << k, (z >> k) << k
        assert not occ[x0:x0 + (1 << k), y0:y0 + (1 << k), z0:z0 + (1 << k)].any()
    m = rng.random((9, 7, 11)) < 0.05
    d = og.dilate(m)
    from scenes import _dilate_np
    assert np.array_equal(d, _dilate_np(m, 1))


def test_optimizers_match_formula():
    rng = np.random.default_rng(1)
    data = rng.standard_normal((50, 4)).astype(np.float32)
    rms = np.abs(rng.standard_normal((50, 4))).astype(np.float32)
    rms[::3] = 0
    grad = rng.standard_normal((50, 4)).astype(np.float32)
    mask = rng.random(50) < 0.5
    d2, r2, g2 = og.rmsprop_step(data, rms, grad, mask, 0.95, 0.1, 1e-8, -1e9, 0.05)
    exp_rms = np.where(rms == 0, grad ** 2, 0.95 * rms + 0.05 * grad ** 2)
    lr = np.array([0.1, 0.1, 0.1, 0.05], np.float32)
    exp_d = data - lr * grad / (np.sqrt(exp_rms) + 1e-8)
    assert np.allclose(r2[mask], exp_rms[mask], rtol=1e-5)
    assert np.allclose(d2[mask], exp_d[mask], rtol=1e-5, atol=1e-6)
    assert (g2[mask] == 0).all() and np.array_equal(g2[~mask], grad[~mask])
    assert np.array_equal(d2[~mask], data[~mask])
    idx = np.nonzero(mask)[0]
    d3, r3, g3 = og.rmsprop_step(data, rms, grad, idx, 0.95, 0.1, 1e-8, -1e9, 0.05)
    assert np.array_equal(d3, d2) and np.array_equal(r3, r2) and np.array_equal(g3, g2)
    d4, g4 = og.sgd_step(data, grad, None, 0.1, -1.0)
    assert np.allclose(d4, data - 0.1 * grad, rtol=1e-6) and (g4 == 0).all()


def test_oracle_matches_the_reference_pytorch_renderer_live(ref_svox2):
    """Not a stored vector: the unmodified reference (oracle/_ref, pure-PyTorch path + autograd) and the
    C oracle on a fresh tiny case, forward colours and all gradients (guards against stale fixtures)."""
    import torch
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(77)
    grid = ref_svox2.SparseGrid(reso=[9, 8, 10], radius=[1.0, 0.9, 1.1], center=[0.05, 0.0, -0.05],
                                basis_dim=4, use_z_order=False, use_sphere_bound=True, device="cpu")
    N = grid.density_data.shape[0]
    with torch.no_grad():
        grid.sh_data[:] = torch.randn(N, 12, generator=gen) * 0.3
        grid.sh_data[:, ::4] += 0.3
        grid.density_data[:] = torch.rand(N, 1, generator=gen) * 4.0
    grid.opt.step_size, grid.opt.background_brightness = 0.41, 0.7
    grid.opt.sigma_thresh = grid.opt.stop_thresh = 0.0
    n = 24
    v = F.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    origins = v * 2.6 + torch.tensor([0.05, 0.0, -0.05])
    dirs = F.normalize((torch.rand(n, 3, generator=gen) - 0.5) * 0.9 - v * 2.6, dim=-1)
    gt = torch.rand(n, 3, generator=gen)
    rgb = grid.volume_render(ref_svox2.Rays(origins, dirs), use_kernel=False)
    F.mse_loss(rgb, gt).backward()
    gsz = torch.tensor(grid.links.shape, dtype=torch.float32)
    og_grid = og.Grid(grid.links.numpy(), grid.density_data.detach().numpy(), grid.sh_data.detach().numpy(),
                      (grid._offset * gsz - 0.5).numpy(), (grid._scaling * gsz).numpy())
    opt = og.make_opt(background_brightness=0.7, step_size=0.41, sigma_thresh=0.0, stop_thresh=0.0)
    out = og.volume_render_fused(og_grid, opt, origins.numpy(), dirs.numpy(), gt.numpy())
    assert np.abs(out["rgb"] - rgb.detach().numpy()).max() < 5e-6
    gd, gsh = grid.density_data.grad.numpy().reshape(-1), grid.sh_data.grad.numpy()
    assert np.abs(out["grad_density"] - gd).max() <= 1e-4 * np.abs(gd).max()
    assert np.abs(out["grad_sh"] - gsh).max() <= 1e-4 * np.abs(gsh).max()
