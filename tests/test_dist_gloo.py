"""World-size-2 gloo (CPU) tests of the multi-GPU host logic: sharding and the masked, packed
gradient exchange must equal a dense all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svox2_b200.dist import exchange_sparse_grads, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D = 5000, 27
        gen = torch.Generator().manual_seed(100 + rank)
        mask_d = torch.rand(N, generator=gen) < 0.05
        mask_s = mask_d | (torch.rand(N, generator=gen) < 0.02)      # TV touches extra rows
        gd = torch.zeros(N, 1)
        gs = torch.zeros(N, D)
        gd[mask_d] = torch.randn(int(mask_d.sum()), 1, generator=gen)
        gs[mask_s] = torch.randn(int(mask_s.sum()), D, generator=gen)
        dense_d, dense_s = gd.clone(), gs.clone()
        dist.all_reduce(dense_d)
        dist.all_reduce(dense_s)
        md, ms = mask_d.to(torch.uint8), mask_s.to(torch.uint8)
        dist.all_reduce(md, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rmd, rms = exchange_sparse_grads(gd, gs, mask_d, mask_s)
        ok = (torch.equal(gd, dense_d) and torch.equal(gs, dense_s) and torch.equal(rmd, md.bool())
              and torch.equal(rms, ms.bool()))
        # replicas identical after the exchange
        gathered = [torch.zeros_like(gs) for _ in range(world)]
        dist.all_gather(gathered, gs)
        ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_exchange_equals_dense_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 5000, 167772):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
