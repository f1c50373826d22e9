"""The fused train step (csrc/step.cu: one C-ABI call, compact gradient buffer) against the
operator-by-operator sequence of opt.py on the same grid / batch / TV cells, and against the
reference's CUDA extension running that sequence.  Bars: colours 1e-5, the two sparse masks
bit-exact, RMS state and parameters within 1e-4 (float atomics reorder; see test_gpu_vs_reference
for why parameters after an RMSProp step are compared by fraction)."""
import numpy as np
import pytest
import torch

import svox2_b200 as sv
from svox2_b200.dist import ShardedTrainer
from gpu_helpers import make_grid, set_opt, t
from scenes import frac_bad, make_scene, rel_err
from test_gpu_parity import OPTS, SCENES

pytestmark = pytest.mark.gpu

HYPER = dict(lr_sigma=30.0, lr_sh=1e-2, lambda_tv=1e-4, lambda_tv_sh=1e-2, tv_sparsity=0.05, rms_beta=0.95,
             beta_loss=1e-3, sparsity_loss=1e-4)


def _np(x):
    return x.detach().cpu().numpy()


@pytest.mark.parametrize("scene,optname", [("sparse32_sh9", "default"), ("sparse_24x32x20_sh4", "default"),
                                           ("sparse32_sh1", "step037_near"), ("sparse64_sh9", "default"),
                                           ("sparse32_sh9", "opaque"), ("dense16_sh9", "nothresh")])
def test_fused_step_equals_operator_sequence(scene, optname):
    sc = make_scene(**SCENES[scene])
    a, b = make_grid(sv, sc), make_grid(sv, sc)
    for g in (a, b):
        set_opt(g, **OPTS[optname])
    fused = ShardedTrainer(a, fused=True, **HYPER)
    plain = ShardedTrainer(b, fused=False, **HYPER)
    rays, gt = sv.Rays(t(sc["origins"]), t(sc["dirs"])), t(sc["rgb_gt"])
    for it in range(3):
        np.random.seed(100 + it)
        ra = fused.step(rays, gt)
        np.random.seed(100 + it)
        rb = plain.step(rays, gt)
        torch.cuda.synchronize()
        assert a.density_data.grad is None or not a.density_data.grad.any()   # dense .grad untouched
        assert (ra - rb).abs().max().item() < (1e-5 if it == 0 else 2e-3)
        want = ((ra - gt) ** 2).mean().item()                       # the loss the step reports
        assert abs(fused.last_mse.item() - want) <= 1e-5 * max(want, 1e-12)
        st = fused.export_masks()
        assert st[1] == 0 and st[2] == 0                                    # no queue overflow
        if it == 0:     # same parameters going in: the masks must agree bit for bit
            assert torch.equal(a.sparse_grad_indexer, b.sparse_grad_indexer)
            assert torch.equal(a.sparse_sh_grad_indexer, b.sparse_sh_grad_indexer)
            # slots are a superset of the masks: the TV marking pass is conservative (data independent)
            assert st[0] >= int((b.sparse_grad_indexer | b.sparse_sh_grad_indexer).sum())
            assert rel_err(_np(a.density_rms), _np(b.density_rms)) < 1e-4
            assert rel_err(_np(a.sh_rms), _np(b.sh_rms)) < 1e-4
            assert frac_bad(_np(a.density_data), _np(b.density_data)) < 2e-3
            assert frac_bad(_np(a.sh_data), _np(b.sh_data)) < 2e-3
            # rows outside the masks were not written at all
            untouched = ~(b.sparse_grad_indexer | b.sparse_sh_grad_indexer)
            assert torch.equal(a.sh_data.data[untouched], t(sc["sh"])[untouched])
            assert torch.equal(a.density_data.data[untouched], t(sc["density"])[untouched])
    fused.check_overflow()


def test_fused_step_masks_and_rms_vs_reference_cuda(ref_svox2):
    """Against the reference's own step (volume_render_fused + inplace_tv_* + optim_*_step): masks
    bit-exact, colours 1e-5, RMS state 1e-4."""
    sc = make_scene(**SCENES["sparse64_sh9"])
    ours, ref = make_grid(sv, sc), make_grid(ref_svox2, sc)
    for g in (ours, ref):
        set_opt(g, **OPTS["default"])
    tr = ShardedTrainer(ours, fused=True, **HYPER)
    o, d, gt = t(sc["origins"]), t(sc["dirs"]), t(sc["rgb_gt"])
    np.random.seed(7)
    rgb_o = tr.step(sv.Rays(o, d), gt)
    np.random.seed(7)
    rgb_r = ref.volume_render_fused(ref_svox2.Rays(o, d), gt, beta_loss=HYPER["beta_loss"],
                                    sparsity_loss=HYPER["sparsity_loss"])
    ref.inplace_tv_grad(ref.density_data.grad, scaling=HYPER["lambda_tv"], sparse_frac=HYPER["tv_sparsity"])
    ref.inplace_tv_color_grad(ref.sh_data.grad, scaling=HYPER["lambda_tv_sh"], sparse_frac=HYPER["tv_sparsity"])
    md, msh = ref.sparse_grad_indexer.clone(), ref.sparse_sh_grad_indexer.clone()
    ref.optim_density_step(HYPER["lr_sigma"], beta=HYPER["rms_beta"])
    ref.optim_sh_step(HYPER["lr_sh"], beta=HYPER["rms_beta"])
    torch.cuda.synchronize()
    tr.export_masks()
    assert (rgb_o - rgb_r).abs().max().item() < 1e-5
    assert torch.equal(ours.sparse_grad_indexer, md) and torch.equal(ours.sparse_sh_grad_indexer, msh)
    assert rel_err(_np(ours.density_rms), _np(ref.density_rms)) < 1e-4
    assert rel_err(_np(ours.sh_rms), _np(ref.sh_rms)) < 1e-4
    assert frac_bad(_np(ours.sh_data), _np(ref.sh_data)) < 2e-3


def test_fused_step_queue_overflow_is_reported():
    sc = make_scene(**SCENES["sparse64_sh9"])
    g = make_grid(sv, sc)
    set_opt(g, **OPTS["default"])
    tr = ShardedTrainer(g, fused=True, **HYPER)
    tr.slots_per_ray = 64           # the C entry refuses less than 64 records per ray
    rays, gt = sv.Rays(t(sc["origins"]), t(sc["dirs"])), t(sc["rgb_gt"])
    np.random.seed(1)
    tr.step(rays, gt)
    tr.check_overflow()             # 64 per ray is plenty for this scene
    assert tr._step_ws.status()[0][0] > 0


def test_rays_longer_than_the_block_table():
    """A tube grid of 2304 cells in z crossed end to end: ~120 record blocks per ray, more than the
    64 entries of K1's per-ray block table, so K3 goes table -> chain inside one ray.  The reference
    result is the operator sequence on a second grid: its queue (512 records per ray on average)
    overflows for these rays and the lattice kernel -- a different implementation -- redoes the batch."""
    sc = make_scene(reso=(6, 6, 2304), basis_dim=4, seed=11, n_rays=48, sparse=False, accel=False)
    rng = np.random.default_rng(5)
    n = 48
    o = np.zeros((n, 3), np.float32)
    o[:, :2] = rng.uniform(-0.3, 0.3, (n, 2))
    o[:, 2] = -1.02     # (the march never goes past t = 2000 grid units, render_util.cuh:581: start at the face)
    d = np.zeros((n, 3), np.float32)
    d[:, :2] = rng.uniform(-0.02, 0.02, (n, 2))
    d[:, 2] = 1.0
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    sc["origins"], sc["dirs"] = o, d.astype(np.float32)
    sc["rgb_gt"] = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    sc["density"] = (sc["density"] * 0.05).astype(np.float32)     # keep the rays alive to the far end
    a, b = make_grid(sv, sc), make_grid(sv, sc)
    for g in (a, b):
        set_opt(g, **OPTS["default"])
    hyper = dict(HYPER, tv_sparsity=0.01)
    fused = ShardedTrainer(a, fused=True, **hyper)
    plain = ShardedTrainer(b, fused=False, **hyper)
    rays, gt = sv.Rays(t(sc["origins"]), t(sc["dirs"])), t(sc["rgb_gt"])
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    from svox2_b200 import csrc as C
    C.volume_render_cuvol(a._to_cpp(), rays._to_cpp(), a.opt._to_cpp(), _counters=counters)
    assert counters[2].item() / n > 64 * 32, "the scene must give more than 64 full blocks per ray"
    np.random.seed(3)
    ra = fused.step(rays, gt)
    np.random.seed(3)
    rb = plain.step(rays, gt)
    torch.cuda.synchronize()
    st = fused.export_masks()
    assert st[1] == 0 and st[2] == 0                                        # the fused step's queue held them
    assert (ra - rb).abs().max().item() < 1e-5
    assert torch.equal(a.sparse_grad_indexer, b.sparse_grad_indexer)
    assert torch.equal(a.sparse_sh_grad_indexer, b.sparse_sh_grad_indexer)
    assert rel_err(_np(a.density_rms), _np(b.density_rms)) < 1e-4           # rms = g^2 after the first step
    assert rel_err(_np(a.sh_rms), _np(b.sh_rms)) < 1e-4
