"""Config-scale parity against the REFERENCE CUDA extension (oracle/_ref, sm_100 build of the
unmodified sxyu/svox2) at the sizes BASELINE.json names:

  C2  SparseGrid(256, use_z_order=True, use_sphere_bound=True): opt.py initialisation and a
      trained-like field; fused step + TV + RMSProp
  C3  the optimisers on identical gradients at 512^3 / 13.4 M rows (rms bit-exact, data 1e-6)
  C4  LLFF NDC grid 1408x1156x128, one 1008x756 image (render_lerp_kernel_cuvol.cu:960-1008)
  C5  tnt.json: 640^3 sparsified foreground + MSI background_reso=1024, nlayers=64, cameras inside
      the grid, lambda_beta / lambda_sparsity / TV weights of opt/configs/tnt.json

plus `sparsify_background` / `resize` (svox2.py:1426-1485).  Bars (north_star): links and masks
bit-exact (torch.equal), RGB / gradients within 1e-4 relative.
"""
import math

import numpy as np
import pytest
import torch

import bench
import svox2_b200 as sv
from scenes import frac_bad, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _np(x):
    return x.detach().cpu().numpy()


def _free():
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def _blob_inside(reso, seed, scale, device=DEV, nb=40):
    gen = torch.Generator().manual_seed(seed)
    centers = (torch.rand(nb, 3, generator=gen) - 0.5) * 1.3
    radii = torch.rand(nb, 3, generator=gen) * 0.17 + 0.08
    return bench._inside_blobs(reso, centers, radii, scale, torch.device(device))


def _grads_close(ours, ref, what=("density_data", "sh_data")):
    for name in what:
        go, gr = _np(getattr(ours, name).grad), _np(getattr(ref, name).grad)
        assert rel_err(go, gr) < TOL, (name, rel_err(go, gr))


# ----------------------------------------------------------------------------------------------
# C2
# ----------------------------------------------------------------------------------------------
def test_config_c2_sphere_bound_zorder_grid_vs_reference_cuda(ref_svox2):
    """BASELINE.json configs[1]: 256^3, SH deg 2, z-order rows inside the sphere bound
    (N = 8 962 370), 5000 NeRF-synthetic-shaped rays.  (a) opt.py's initial state (sh = 0,
    sigma = 0.1: every in-sphere sample is active, ~313 per ray); (b) a trained-like field
    (sigma = 30 in seeded ellipsoids) through one whole step: fused render + sparse TV x2 +
    RMSProp x2."""
    ours = sv.SparseGrid(256, basis_dim=9, use_z_order=True, use_sphere_bound=True, device=DEV)
    ref = ref_svox2.SparseGrid(256, basis_dim=9, use_z_order=True, use_sphere_bound=True, device=DEV)
    assert ours.sh_data.size(0) == ref.sh_data.size(0) == 8962370
    assert torch.equal(ours.links, ref.links)                      # Morton init + accelerate
    for g in (ours, ref):
        for k, v in bench.RENDER.items():
            setattr(g.opt, k, v)
    o, d = bench.build_rays(2 * bench.BATCH, torch.device(DEV))
    gt = torch.rand(o.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))

    # (a) opt.py:300-301
    with torch.no_grad():
        for g in (ours, ref):
            g.sh_data.data[:] = 0.0
            g.density_data.data[:] = 0.1
    s = slice(0, bench.BATCH)
    ro = ours.volume_render_fused(sv.Rays(o[s], d[s]), gt[s])
    rr = ref.volume_render_fused(ref_svox2.Rays(o[s], d[s]), gt[s])
    assert (ro - rr).abs().max().item() < 1e-5
    assert torch.equal(ours.sparse_grad_indexer, ref.sparse_grad_indexer)
    assert int(ours.sparse_grad_indexer.sum()) > 1_000_000
    _grads_close(ours, ref)
    for g in (ours, ref):
        g.density_data.grad.zero_()
        g.sh_data.grad.zero_()

    # (b) trained-like: rows are in Morton order, so the field goes through the links
    inside = _blob_inside(256, 20200823, 0.8) & (ours.links >= 0)
    rows = ours.links[inside].long()
    gen = torch.Generator(device=DEV).manual_seed(5)
    sh = torch.randn(ours.sh_data.shape, generator=gen, device=DEV) * 0.1
    sh[:, ::9] += 0.5
    with torch.no_grad():
        for g in (ours, ref):
            g.density_data.data[:] = 0.0
            g.density_data.data[rows] = 30.0
            g.sh_data.data[:] = sh
    s = slice(bench.BATCH, 2 * bench.BATCH)
    syn = bench.SYN
    np.random.seed(77)
    ro = ours.volume_render_fused(sv.Rays(o[s], d[s]), gt[s])
    ours.inplace_tv_grad(ours.density_data.grad, scaling=syn["lambda_tv"], sparse_frac=syn["tv_sparsity"])
    ours.inplace_tv_color_grad(ours.sh_data.grad, scaling=syn["lambda_tv_sh"], sparse_frac=syn["tv_sparsity"])
    np.random.seed(77)
    rr = ref.volume_render_fused(ref_svox2.Rays(o[s], d[s]), gt[s])
    ref.inplace_tv_grad(ref.density_data.grad, scaling=syn["lambda_tv"], sparse_frac=syn["tv_sparsity"])
    ref.inplace_tv_color_grad(ref.sh_data.grad, scaling=syn["lambda_tv_sh"], sparse_frac=syn["tv_sparsity"])
    assert (ro - rr).abs().max().item() < 1e-5
    assert torch.equal(ours.sparse_grad_indexer, ref.sparse_grad_indexer)
    assert torch.equal(ours.sparse_sh_grad_indexer, ref.sparse_sh_grad_indexer)
    _grads_close(ours, ref)
    # identical gradients into both optimisers: rms bit for bit, parameters 1e-6 (A.9)
    ref.density_data.grad.copy_(ours.density_data.grad)
    ref.sh_data.grad.copy_(ours.sh_data.grad)
    for g in (ours, ref):
        g.optim_density_step(syn["lr_sigma"], beta=syn["rms_beta"])
        g.optim_sh_step(syn["lr_sh"], beta=syn["rms_beta"])
    assert torch.equal(ours.density_rms, ref.density_rms)
    assert torch.equal(ours.sh_rms, ref.sh_rms)
    assert rel_err(_np(ours.density_data), _np(ref.density_data)) < 1e-6
    assert rel_err(_np(ours.sh_data), _np(ref.sh_data)) < 1e-6
    assert not ours.sh_data.grad.any() and not ref.sh_data.grad.any()
    del ours, ref
    _free()


# ----------------------------------------------------------------------------------------------
# C3: optimisers on identical inputs at full size
# ----------------------------------------------------------------------------------------------
def test_config_c3_optimisers_on_identical_gradients_vs_reference_cuda(ref_svox2):
    """north_star: 'updated parameters within 1e-4'.  A whole train step cannot meet that literally
    (the first RMSProp step is g/|g|: atomics-order noise on cancelling gradients flips signs), so
    the bar is split: gradients <= 1e-4 (test_config_c3_training_steps_follow_the_reference_cuda)
    and, here, the optimiser kernels fed the SAME gradient at C3 size: rms bit-exact, data 1e-6,
    over two steps (first-touch and running-average branches of optim_kernel.cu:15-25)."""
    RC = ref_svox2.svox2._C
    from svox2_b200 import csrc as C
    N = 13_390_000
    gen = torch.Generator(device=DEV).manual_seed(11)
    mask = torch.rand(N, generator=gen, device=DEV) < 0.12
    data0 = torch.randn(N, 27, generator=gen, device=DEV) * 0.1
    pad = lambda x: sv.grid._padded_copy(x)          # our HBM layout: 28-float rows
    do, ro = pad(data0), pad(torch.zeros_like(data0))
    dr, rr = data0.clone(), torch.zeros_like(data0)
    for it in range(2):
        g = torch.randn(N, 27, generator=gen, device=DEV) * 1e-3 * mask.unsqueeze(-1)
        go, gr = pad(g), g.clone()
        C.rmsprop_step(do, ro, go, mask, 0.95, 1e-2, 1e-8, -1e9, 1e-2)
        RC.rmsprop_step(dr, rr, gr, mask, 0.95, 1e-2, 1e-8, -1e9, 1e-2)
        assert torch.equal(ro, rr), it
        assert not go.any() and not gr.any()
        assert (do - dr).abs().max().item() <= 1e-6 * dr.abs().max().item(), it
    # density: [N,1], index-list indexer (what the reference converts small masks to)
    idx = torch.nonzero(mask).flatten()
    d1, r1, g1 = (torch.rand(N, 1, generator=gen, device=DEV) for _ in range(3))
    d2, r2, g2 = d1.clone(), r1.clone(), g1.clone()
    C.rmsprop_step(d1, r1, g1, idx, 0.95, 30.0, 1e-8, -1e9, 30.0)
    RC.rmsprop_step(d2, r2, g2, idx, 0.95, 30.0, 1e-8, -1e9, 30.0)
    assert torch.equal(r1, r2) and torch.equal(g1, g2)
    assert (d1 - d2).abs().max().item() <= 1e-6 * d2.abs().max().item()
    _free()


# ----------------------------------------------------------------------------------------------
# C4
# ----------------------------------------------------------------------------------------------
def test_config_c4_llff_ndc_image_vs_reference_cuda(ref_svox2):
    """BASELINE.json configs[3]: forward-facing NDC grid 1408x1156x128 (radius 1.496/1.661/1.0,
    llff.json options), one full 1008x756 image against ref.volume_render_image <= 1e-5."""
    reso = [1408, 1156, 128]
    dev = torch.device(DEV)
    gen = torch.Generator(device=DEV).manual_seed(11)
    keep = torch.zeros(reso, dtype=torch.bool, device=dev)
    for z0 in (20, 52, 90):        # content slabs at a few NDC depths, ~4 % occupancy
        blob = torch.rand(reso[0] // 16 + 1, reso[1] // 16 + 1, generator=gen, device=dev) < 0.45
        keep[:, :, z0:z0 + 4] |= blob.repeat_interleave(16, 0).repeat_interleave(16, 1)[:reso[0], :reso[1], None]
    n_rows = int(keep.sum())
    links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
    links[~keep.view(-1)] = -1
    links = links.view(*reso).contiguous()
    del keep
    density = torch.rand(n_rows, 1, generator=gen, device=dev) * 40.0
    sh = torch.randn(n_rows, 27, generator=gen, device=dev) * 0.1
    sh[:, ::9] += 0.5
    grids = []
    for mod in (sv, ref_svox2):
        g = mod.SparseGrid(reso=[8, 8, 8], radius=[1.496031746031746, 1.6613756613756614, 1.0],
                           basis_dim=9, use_sphere_bound=False, device=dev)
        g.links = links.clone()
        g.density_data = torch.nn.Parameter(density.clone())
        g.sh_data = torch.nn.Parameter(sh.clone())
        g.capacity = n_rows
        g.accelerate()
        g.opt.step_size, g.opt.sigma_thresh, g.opt.stop_thresh = 0.5, 1e-8, 1e-7
        g.opt.background_brightness, g.opt.last_sample_opaque = 0.5, False
        grids.append(g)
    ours, ref = grids
    assert torch.equal(ours.links, ref.links)
    W, H, fx = 1008, 756, 820.0
    ndc = (2 * fx / W, 2 * fx / H)
    c2w = torch.eye(4, device=dev)
    c2w[:3, 3] = torch.tensor([0.05, -0.03, 0.0])
    with torch.no_grad():
        img_o = ours.volume_render_image(sv.Camera(c2w, fx=fx, width=W, height=H, ndc_coeffs=ndc))
        img_r = ref.volume_render_image(ref_svox2.Camera(c2w, fx=fx, width=W, height=H, ndc_coeffs=ndc))
    assert img_o.shape == img_r.shape == (H, W, 3)
    assert float((img_r - 0.5).abs().max()) > 0.05           # content was hit
    assert (img_o - img_r).abs().max().item() < 1e-5
    del ours, ref, grids
    _free()


# ----------------------------------------------------------------------------------------------
# C5
# ----------------------------------------------------------------------------------------------
def _c5_pair(ref_svox2, reso=640, bg_reso=1024, bg_layers=64):
    """tnt.json-sized scene: sparsified 640^3 foreground (~6 % occupancy), MSI 2048x1024x64."""
    dev = torch.device(DEV)
    inside = _blob_inside(reso, 99, 0.85)
    keep = bench._dilate(inside, 2)
    n_rows = int(keep.sum())
    links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
    links[~keep.view(-1)] = -1
    links = links.view(reso, reso, reso).contiguous()
    gen = torch.Generator(device=DEV).manual_seed(17)
    density = (inside[keep].float() * 25.0).view(n_rows, 1).contiguous()
    del inside, keep
    sh = torch.randn(n_rows, 27, generator=gen, device=dev) * 0.1
    sh[:, ::9] += 0.5
    nbg = 2 * bg_reso * bg_reso
    bg = torch.zeros(nbg, bg_layers, 4, device=dev)
    bg[..., :3] = torch.randn(nbg, bg_layers, 3, generator=gen, device=dev) * 0.4
    # sparse shells of density: ~20 % of the texels carry sigma in a few layers
    lay = torch.rand(nbg, bg_layers, generator=gen, device=dev)
    tex = torch.rand(nbg, 1, generator=gen, device=dev) < 0.2
    bg[..., 3] = torch.where((lay < 0.15) & tex, lay * 40.0, torch.zeros_like(lay))
    del lay, tex
    grids = []
    for mod in (sv, ref_svox2):
        g = mod.SparseGrid(reso=[8, 8, 8], basis_dim=9, use_sphere_bound=False, device=dev,
                           background_nlayers=bg_layers, background_reso=bg_reso)
        g.links = links.clone()
        g.density_data = torch.nn.Parameter(density.clone())
        g.sh_data = torch.nn.Parameter(sh.clone())
        g.capacity = n_rows
        g.background_data = torch.nn.Parameter(bg.clone())
        assert tuple(g.background_links.shape) == (2 * bg_reso, bg_reso)
        g.accelerate()
        g.opt.step_size, g.opt.sigma_thresh, g.opt.stop_thresh = 0.5, 1e-8, 1e-7
        g.opt.background_brightness = 1.0
        grids.append(g)
    del bg, sh
    return grids[0], grids[1], n_rows


def _inside_rays(n, seed):
    """Cameras INSIDE the grid at radius ~0.9*0.95 looking roughly at the centre region (the NSVF
    loader scales T&T cameras by 0.95, nsvf_dataset.py:46,167-170); many rays leave the foreground
    box and composite the MSI shells."""
    gen = torch.Generator().manual_seed(seed)
    v = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    origins = v * 0.85
    target = (torch.rand(n, 3, generator=gen) - 0.5) * 1.2
    dirs = torch.nn.functional.normalize(target - origins, dim=-1)
    return origins.to(DEV).contiguous(), dirs.to(DEV).contiguous()


def test_config_c5_tnt_msi_train_step_vs_reference_cuda(ref_svox2):
    """BASELINE.json configs[4] at tnt.json size: fused step with beta + sparsity losses, bg mask
    bit-exact, sparse TV (sigma, SH), msi_tv_grad_sparse, optim_background_step, then
    sparsify_background on the result -- every stage against the reference CUDA extension."""
    ours, ref, n_rows = _c5_pair(ref_svox2)
    assert torch.equal(ours.links, ref.links)
    o, d = _inside_rays(bench.BATCH, 5)
    gt = torch.rand(o.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9))
    tnt = dict(lambda_beta=1e-5, lambda_sparsity=1e-11, lambda_tv=5e-5, lambda_tv_sh=5e-3,
               lambda_tv_bg=1e-3)
    np.random.seed(31)
    ro = ours.volume_render_fused(sv.Rays(o, d), gt, beta_loss=tnt["lambda_beta"],
                                  sparsity_loss=tnt["lambda_sparsity"])
    np.random.seed(31)
    rr = ref.volume_render_fused(ref_svox2.Rays(o, d), gt, beta_loss=tnt["lambda_beta"],
                                 sparsity_loss=tnt["lambda_sparsity"])
    assert (ro - rr).abs().max().item() < 1e-5
    assert torch.equal(ours.sparse_grad_indexer, ref.sparse_grad_indexer)
    assert torch.equal(ours.sparse_background_indexer, ref.sparse_background_indexer)
    n_bg = int(ours.sparse_background_indexer.sum())
    assert n_bg > 10_000, n_bg                         # the MSI was really exercised
    _grads_close(ours, ref, ("density_data", "sh_data", "background_data"))
    # regularisers (opt.py:551-583); same numpy RNG stream on both sides
    for g, seed in ((ours, 32), (ref, 32)):
        np.random.seed(seed)
        g.inplace_tv_grad(g.density_data.grad, scaling=tnt["lambda_tv"], sparse_frac=0.01)
        g.inplace_tv_color_grad(g.sh_data.grad, scaling=tnt["lambda_tv_sh"], sparse_frac=0.01)
        torch.manual_seed(33)                          # contiguous=False draws torch.randint cells
        g.inplace_tv_background_grad(g.background_data.grad, scaling=tnt["lambda_tv_bg"],
                                     scaling_density=tnt["lambda_tv_bg"], sparse_frac=0.01,
                                     contiguous=True)
    assert torch.equal(ours.sparse_grad_indexer, ref.sparse_grad_indexer)
    assert torch.equal(ours.sparse_sh_grad_indexer, ref.sparse_sh_grad_indexer)
    assert torch.equal(ours.sparse_background_indexer, ref.sparse_background_indexer)
    _grads_close(ours, ref, ("density_data", "sh_data", "background_data"))
    # optimisers on identical gradients (see the C3 test for why): rms bit-exact
    ref.background_data.grad.copy_(ours.background_data.grad)
    ref.density_data.grad.copy_(ours.density_data.grad)
    ref.sh_data.grad.copy_(ours.sh_data.grad)
    for g in (ours, ref):
        g.optim_density_step(30.0, beta=0.95)
        g.optim_sh_step(1e-2, beta=0.95)
        g.optim_background_step(3.0, 0.1, beta=0.95)
    assert torch.equal(ours.background_rms, ref.background_rms)
    assert torch.equal(ours.sh_rms, ref.sh_rms)
    assert rel_err(_np(ours.background_data), _np(ref.background_data)) < 1e-6
    assert not ours.background_data.grad.any()
    # second render from the updated parameters
    with torch.no_grad():
        r2o = ours.volume_render(sv.Rays(o, d))
        r2r = ref.volume_render(ref_svox2.Rays(o, d))
    assert (r2o - r2r).abs().max().item() < 1e-4
    # opt.py:639 after the last upsample
    ours.sparsify_background(1.0)
    ref.sparsify_background(1.0)
    assert torch.equal(ours.background_links, ref.background_links)
    assert torch.equal(ours.background_data.data, ref.background_data.data)
    assert 0 < ours.background_data.size(0) < 2 * 1024 * 1024
    with torch.no_grad():
        r3o = ours.volume_render(sv.Rays(o, d))
        r3r = ref.volume_render(ref_svox2.Rays(o, d))
    assert (r3o - r3r).abs().max().item() < 1e-4
    del ours, ref
    _free()


# ----------------------------------------------------------------------------------------------
# a18: sparsify_background / resize
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("thresh,dil", [(1.0, 1), (0.4, 0), (1.15, 2)])
def test_sparsify_background_bit_exact_vs_reference(ref_svox2, thresh, dil):
    """svox2.py:1426-1448: links renumbering and the retained texel rows, incl. already-sparse
    background_links (a second call) and dilation across layers."""
    from gpu_helpers import make_grid
    from scenes import make_scene
    sc = make_scene(reso=16, basis_dim=1, seed=5, n_rays=4, bg=(48, 12))
    ours, ref = make_grid(sv, sc), make_grid(ref_svox2, sc)
    for rep in range(2):                                  # the second pass starts from sparse links
        ours.sparsify_background(thresh + 0.3 * rep, dilate=dil)
        ref.sparsify_background(thresh + 0.3 * rep, dilate=dil)
        assert torch.equal(ours.background_links, ref.background_links)
        assert torch.equal(ours.background_data.data, ref.background_data.data)
        assert ours.background_data.size(0) == int((ours.background_links >= 0).sum())


def test_resize_basis_vs_reference(ref_svox2):
    """svox2.py:1451-1485.  Growing the basis: the shared low-order coefficients of each colour keep
    their values, the new ones are 0 -- identical to the reference, which however returns a stale
    svox1-era extra column ([N, 3*basis_dim + 1]) that makes its result unusable by its own
    kernels; ours returns the [N, 3*basis_dim] layout.  Shrinking raises a shape error upstream
    (same extra column); ours keeps the low-order coefficients."""
    from gpu_helpers import make_grid
    from scenes import make_scene
    sc = make_scene(reso=16, basis_dim=4, seed=6, n_rays=4)
    ours, ref = make_grid(sv, sc), make_grid(ref_svox2, sc)
    old = ours.sh_data.data.clone()
    ours.resize(9)
    ref.resize(9)
    assert ours.basis_dim == ref.basis_dim == 9 and ours.sh_data.shape == (old.size(0), 27)
    assert ref.sh_data.shape == (old.size(0), 28) and not ref.sh_data.data[:, 27].any()
    assert torch.equal(ours.sh_data.data, ref.sh_data.data[:, :27])
    for c in range(3):
        assert torch.equal(ours.sh_data.data[:, 9 * c:9 * c + 4], old[:, 4 * c:4 * c + 4])
        assert not ours.sh_data.data[:, 9 * c + 4:9 * c + 9].any()
    rays = sv.Rays(torch.from_numpy(sc["origins"]).to(DEV), torch.from_numpy(sc["dirs"]).to(DEV))
    with torch.no_grad():
        assert torch.isfinite(ours.volume_render(rays)).all()      # the resized grid renders
    ours.resize(1)
    assert ours.sh_data.shape == (old.size(0), 3)
    assert torch.equal(ours.sh_data.data, old[:, ::4])
    with pytest.raises(RuntimeError):
        ref.resize(1)
