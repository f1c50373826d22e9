"""SURVEY 8(f) rank 4 regularisers behind the operator boundary: lumisphere TV
(loss_kernel.cu:335-434 / svox2.py:1822-1895) and the L2 colour regulariser (svox2.py:1897-1928),
against the C oracle and against the reference's CUDA extension on the same inputs."""
import numpy as np
import pytest
import torch

import oracle as og
import svox2_b200 as sv
from svox2_b200 import csrc as C
from gpu_helpers import make_grid, our_spec, set_opt, t
from scenes import make_scene, rel_err
from test_gpu_parity import OPTS, SCENES

pytestmark = pytest.mark.gpu


def _cells(shape, n, seed):
    X, Y, Z = shape
    # indices that decode inside the grid with the kernel's (Y-1, Z-1) strides; the reference reads
    # past the end of `links` for the rest (see loss.cu), so they cannot be compared
    return np.random.default_rng(seed).integers(0, (X - 1) * (Y - 1) * (Z - 1), n).astype(np.int32)


@pytest.mark.parametrize("scene,dir_factor", [("sparse32_sh9", 0.0), ("sparse32_sh9", 0.7),
                                              ("sparse_24x32x20_sh4", 1.0), ("sparse32_sh1", 0.3)])
def test_lumisphere_kernel_vs_oracle_and_reference_cuda(ref_svox2, scene, dir_factor):
    RC = ref_svox2.svox2._C
    sc = make_scene(**SCENES[scene])
    B = sc["basis_dim"]
    rng = np.random.default_rng(11)
    d = rng.standard_normal((2, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d[1] = d[0] + 0.05 * d[1]
    d[1] /= np.linalg.norm(d[1])
    fn = sv.utils.eval_sh_bases(B, torch.from_numpy(d).float())
    fn0, fn1 = fn[0].contiguous().cuda(), fn[1].contiguous().cuda()
    cells = _cells(sc["links"].shape, 6000, 5)
    want, want_mask = og.lumisphere_tv_grad_sparse(sc["links"], sc["sh"], cells, fn0.cpu().numpy(),
                                                   fn1.cpu().numpy(), 0.37, dir_factor)

    spec = our_spec(sc)
    gh = C.GridOutputGrads()
    gh.grad_sh_out = torch.zeros_like(spec.sh_data)
    gh.mask_out = torch.zeros(spec.sh_data.size(0), dtype=torch.bool, device="cuda")
    C.lumisphere_tv_grad_sparse(spec, t(cells), fn0, fn1, 0.37, -1.0, -1.0, dir_factor, gh)
    assert rel_err(gh.grad_sh_out.cpu().numpy(), want) < 1e-4
    assert np.array_equal(gh.mask_out.cpu().numpy(), want_mask)

    rspec = RC.SparseGridSpec()
    rspec.density_data, rspec.sh_data, rspec.links = spec.density_data, spec.sh_data, spec.links
    rspec._offset, rspec._scaling = spec._offset, spec._scaling
    rspec.basis_dim, rspec.basis_type = B, 1
    rspec.basis_data = torch.empty(0, device="cuda")
    rspec.background_links = torch.empty((0, 0), dtype=torch.int32, device="cuda")
    rspec.background_data = torch.empty((0, 0, 0), device="cuda")
    rh = RC.GridOutputGrads()
    rh.grad_sh_out = torch.zeros_like(spec.sh_data)
    rh.mask_out = torch.zeros_like(gh.mask_out)
    RC.lumisphere_tv_grad_sparse(rspec, t(cells), fn0, fn1, 0.37, -1.0, -1.0, dir_factor, rh)
    torch.cuda.synchronize()
    assert torch.equal(gh.mask_out, rh.mask_out), "lumisphere mask must be bit-exact"
    assert rel_err(gh.grad_sh_out.cpu().numpy(), rh.grad_sh_out.cpu().numpy()) < 1e-4


def test_inplace_tv_lumisphere_grad_api_vs_reference(ref_svox2):
    """Through SparseGrid.inplace_tv_lumisphere_grad with the RNGs seeded alike: same random cells
    (numpy), same view direction (torch CUDA RNG), same perturbation (torch CPU RNG + scipy)."""
    sc = make_scene(**SCENES["sparse32_sh9"])
    ours, ref = make_grid(sv, sc), make_grid(ref_svox2, sc)
    outs = []
    for grid, mod in ((ours, sv), (ref, ref_svox2)):
        set_opt(grid, **OPTS["default"])
        grid.volume_render_fused(mod.Rays(t(sc["origins"]), t(sc["dirs"])), t(sc["rgb_gt"]))
        grid.sh_data.grad.zero_()
        # the contiguous cell run starts where numpy says; pick a seed that keeps it inside the range
        # the reference can decode without reading past `links` (see the kernel test above)
        n = grid.links.numel()
        seed = next(s for s in range(100) if np.random.RandomState(s).randint(0, n) <= 0.8 * n)
        np.random.seed(seed)
        torch.manual_seed(4)
        grid.inplace_tv_lumisphere_grad(grid.sh_data.grad, scaling=2e-2, sparse_frac=0.05,
                                        dir_factor=0.5, dir_perturb_radians=0.05)
        torch.cuda.synchronize()
        outs.append((grid.sh_data.grad.detach().clone(), grid.sparse_sh_grad_indexer.clone()))
    (go, mo), (gr, mr) = outs
    assert gr.abs().max().item() > 0
    assert torch.equal(mo, mr)
    assert rel_err(go.cpu().numpy(), gr.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("sparse", [True, False])
def test_inplace_l2_color_grad_vs_reference(ref_svox2, sparse):
    sc = make_scene(**SCENES["sparse32_sh9"])
    ours, ref = make_grid(sv, sc), make_grid(ref_svox2, sc)
    res = []
    for grid, mod in ((ours, sv), (ref, ref_svox2)):
        set_opt(grid, **OPTS["default"])
        grid.volume_render_fused(mod.Rays(t(sc["origins"]), t(sc["dirs"])), t(sc["rgb_gt"]))
        if not sparse:
            grid.sparse_grad_indexer = grid.sparse_sh_grad_indexer = None
        grid.sh_data.grad.zero_()
        grid.inplace_l2_color_grad(grid.sh_data.grad, start_dim=1, end_dim=-2, scaling=0.3)
        res.append(grid.sh_data.grad.detach().clone())
    assert res[1].abs().max().item() > 0
    assert torch.allclose(res[0], res[1], rtol=1e-6, atol=0)
