"""2-GPU NCCL parity: a ray-sharded step on global batch B must equal the 1-GPU step on the same B
(masks bit-exact, gradients / parameters 1e-4 relative).  Skipped with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import svox2_b200 as sv
    from svox2_b200.dist import ShardedTrainer
    from gpu_helpers import make_grid, set_opt, t
    from scenes import make_scene, rel_err

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        sc = make_scene(reso=48, basis_dim=9, seed=41, n_rays=4096)
        B = sc["origins"].shape[0]
        opts = dict(sigma_thresh=1e-8, stop_thresh=1e-7)
        # single-GPU reference on the full batch (every rank computes it on its own device)
        ref = make_grid(sv, sc, dev)
        set_opt(ref, **opts)
        np.random.seed(7)
        tr1 = ShardedTrainer(ref, world_size=1, rank=0)
        rgb1 = tr1.step(sv.Rays(t(sc["origins"], dev), t(sc["dirs"], dev)), t(sc["rgb_gt"], dev), global_batch=B)
        # sharded: this rank marches its contiguous slice
        grid = make_grid(sv, sc, dev)
        set_opt(grid, **opts)
        np.random.seed(7)
        b, e = rank * B // world, (rank + 1) * B // world
        tr = ShardedTrainer(grid, world_size=world, rank=rank)
        rgb = tr.step(sv.Rays(t(sc["origins"][b:e], dev), t(sc["dirs"][b:e], dev)), t(sc["rgb_gt"][b:e], dev),
                      global_batch=B)
        torch.cuda.synchronize()
        ok = (rgb - rgb1[b:e]).abs().max().item() < 1e-5
        # after the optimiser: parameters of the sharded replica == single-GPU parameters
        ok = ok and rel_err(grid.density_data.detach().cpu().numpy(), ref.density_data.detach().cpu().numpy()) < 1e-4
        dsh = (grid.sh_data.detach() - ref.sh_data.detach()).abs()
        ok = ok and torch.quantile(dsh.flatten()[:2000000], 0.999).item() < 1e-4
        # replicas bit-identical across ranks
        mine = grid.sh_data.detach().contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_gpu_step_equals_single_gpu_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
