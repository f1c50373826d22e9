"""Ray-sharded multi-GPU training step (one process per GPU, replicated grid).

The reference has no distributed code at all; this is the north-star extension.
Rank r of G marches the contiguous slice [r*B/G, (r+1)*B/G) of the global batch B.  To make the
sum over ranks EQUAL the single-GPU gradient on the same B:

  * the fused MSE / beta seed is normalised with the GLOBAL batch size (``_norm_rays``; the
    reference uses the local Q, render_lerp_kernel_cuvol.cu:736,1127);
  * the sparse-TV cell run is drawn once (same numpy seed on every rank), split across ranks and
    scaled by the global cell count (``_scale_cells``; loss_kernel.cu:584-595), and added BEFORE
    the exchange;
  * gradients are exchanged restricted to the union of the touched-row masks: an OR all-reduce of
    the two masks, then ONE SUM all-reduce of a packed [rows, 3*basis_dim+1] buffer.  A dense
    all-reduce would move N*28*4 bytes (1.5 GB at 512^3/10%) per step;
  * every rank then runs the identical masked RMSProp on identical reduced gradients, so replicas
    stay bit-identical (an all-reduce returns the same bits on every rank).

``exchange_sparse_grads`` is plain ``torch.distributed`` code and is exercised on CPU with the
gloo backend in tests/test_dist_gloo.py; on the GPU box the backend is NCCL over NVLink.
"""
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, world_size: int, rank: int):
    """Contiguous, balanced [begin, end) of n items for this rank."""
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_sparse_grads(grad_density: torch.Tensor, grad_sh: torch.Tensor,
                          mask_density: torch.Tensor, mask_sh: torch.Tensor, group=None):
    """Sum gradients over ranks, touching only rows in the union of all ranks' masks.

    grad_density [N,1], grad_sh [N,D] are updated in place; returns the reduced
    (mask_density, mask_sh) as bool tensors.  Costs one host sync (the packed row count)."""
    masks = torch.stack([mask_density, mask_sh]).to(torch.uint8)
    dist.all_reduce(masks, op=dist.ReduceOp.MAX, group=group)
    rows = torch.nonzero(masks[0] | masks[1], as_tuple=False).flatten()
    D = grad_sh.size(1)
    packed = torch.empty((rows.numel(), D + 1), dtype=grad_sh.dtype, device=grad_sh.device)
    packed[:, :D] = grad_sh[rows]
    packed[:, D] = grad_density[rows, 0]
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    grad_sh[rows] = packed[:, :D]
    grad_density[rows, 0] = packed[:, D]
    return masks[0].bool(), masks[1].bool()


class ShardedTrainer:
    """opt/opt.py's train step (opt.py:474-603) for a replicated grid and sharded rays."""

    def __init__(self, grid, world_size: int = 1, rank: int = 0, lr_sigma: float = 30.0,
                 lr_sh: float = 1e-2, lambda_tv: float = 1e-5, lambda_tv_sh: float = 1e-3,
                 tv_sparsity: float = 0.01, rms_beta: float = 0.95, beta_loss: float = 0.0,
                 sparsity_loss: float = 0.0, group=None, lr_sigma_bg: float = 3e0,
                 lr_color_bg: float = 1e-1, lambda_tv_background_sigma: float = 0.0,
                 lambda_tv_background_color: float = 0.0, tv_background_sparsity: float = 0.01,
                 fused: Optional[bool] = None):
        self.grid, self.world_size, self.rank, self.group = grid, world_size, rank, group
        self.lr_sigma_bg, self.lr_color_bg = lr_sigma_bg, lr_color_bg
        self.lambda_tv_bg_sigma, self.lambda_tv_bg_color = lambda_tv_background_sigma, lambda_tv_background_color
        self.tv_background_sparsity = tv_background_sparsity
        self.lr_sigma, self.lr_sh = lr_sigma, lr_sh
        self.lambda_tv, self.lambda_tv_sh = lambda_tv, lambda_tv_sh
        self.tv_sparsity, self.rms_beta = tv_sparsity, rms_beta
        self.beta_loss, self.sparsity_loss = beta_loss, sparsity_loss
        if tv_sparsity >= 1.0 and world_size > 1 and (lambda_tv > 0.0 or lambda_tv_sh > 0.0):
            # sparse_frac >= 1 selects the dense TV kernel and a dense optimiser step in the
            # reference (svox2.py:1760-1766): there is no sparse mask to exchange then
            raise ValueError("ShardedTrainer with world_size > 1 needs tv_sparsity < 1")
        # march fwd + march bwd + 2 TV + 2 optimiser kernels
        # clear + trace + shade + seed + scatter + (conditional) lattice fallback, 2 TV, then
        # 2 optimiser kernels (1 GPU) or pack + packed optimiser (multi-GPU)
        self.kernels_per_step = 6 + (lambda_tv > 0) + (lambda_tv_sh > 0) + 2
        self.profile, self._events, self.exchanged_rows = False, [], 0
        self._packed = None
        # single GPU: the whole step as ONE call on a compact gradient buffer (csrc/step.cu) unless
        # ``fused=False`` asks for the operator-by-operator sequence of opt.py (dense .grad arrays
        # and bool masks visible between the calls)
        self.fused = fused
        self._step_ws = None
        self.slots_per_ray = None       # record slots per ray of the fused step's queue (None: worst case)
        self.last_mse = None            # fused step: 0-dim device tensor, F.mse_loss(rgb, rgb_gt) of the last step
        self._mse_pair = None

    def _tv_run(self):
        """This rank's slice (start, count) of the step's contiguous wrapped cell run (drawn with
        numpy's global RNG, which must be seeded identically on every rank) and the global count."""
        start, n = self.grid._draw_cell_run(self.tv_sparsity)
        if self.world_size == 1:
            return (start, n), n
        b, e = shard_range(n, self.world_size, self.rank)
        return (start + b, e - b), n

    def _can_fuse(self) -> bool:
        g = self.grid
        return (self.fused is not False and self.world_size == 1 and g.sh_data.is_cuda and
                not g.use_background and g.opt.backend == "cuvol" and self.tv_sparsity < 1.0)

    def _fused_step(self, rays, rgb_gt, lr_sigma, lr_sh):
        """One svox2_train_step call: same sample set, TV cells (numpy RNG consumed as by the two
        inplace_tv_* calls), masks and RMSProp arithmetic as the sequence below; gradients live in a
        compact buffer for the duration of the call and the dense .grad arrays are not touched."""
        from . import csrc as _C
        g = self.grid
        g._ensure_layout()
        g._ensure_rms()
        Q = rgb_gt.size(0)
        slots = self.slots_per_ray
        if slots is None:   # worst case: every lattice step along the diagonal is an active sample
            diag = sum(float(s) ** 2 for s in g.links.shape) ** 0.5
            slots = int(diag / max(float(g.opt.step_size), 1e-3)) + 2
            slots = max(64, min(slots, (4 << 30) // (96 * max(Q, 1))))
        key = (g.sh_data.size(0), int(g.basis_dim), Q, int(slots))
        if self._step_ws is None or self._step_ws.key != key:
            self._step_ws = None
            self._step_ws = _C.TrainStepWorkspace(*key, g.sh_data.device)
        tv_d = tv_s = (0, 0, 0.0)
        if self.lambda_tv > 0.0:
            start, n = g._draw_cell_run(self.tv_sparsity)
            tv_d = (start, n, self.lambda_tv)
        if self.lambda_tv_sh > 0.0:
            start, n = g._draw_cell_run(self.tv_sparsity)
            tv_s = (start, n, self.lambda_tv_sh)
        rgb = torch.empty_like(rgb_gt)
        # two alternating device scalars: a reader on another stream (bench e2e loop) may still be
        # copying step i's loss out while step i+1's clear kernel zeroes its own
        if self._mse_pair is None:
            self._mse_pair = [torch.zeros((), dtype=torch.float32, device=rgb_gt.device) for _ in range(2)]
        self._mse_pair.reverse()
        self.last_mse = self._mse_pair[0]
        _C.train_step(g._to_cpp(), g.density_rms, g.sh_rms, rays._to_cpp(), g.opt._to_cpp(), rgb_gt.contiguous(),
                      rgb, self._step_ws, tv_density=tv_d, tv_sh=tv_s, beta_loss=self.beta_loss,
                      sparsity_loss=self.sparsity_loss, norm_rays=Q, lr_density=lr_sigma, lr_sh=lr_sh,
                      rms_beta=self.rms_beta, mse_out=self.last_mse)
        g.sparse_grad_indexer = g.sparse_sh_grad_indexer = None     # see export_masks()
        # clear, trace, shade, tv_mark, rank count + scan, seed, scatter, TV x2, RMSProp
        self.kernels_per_step = 9 + (self.lambda_tv > 0) + (self.lambda_tv_sh > 0)
        return rgb

    def export_masks(self):
        """After a fused step: materialise grid.sparse_grad_indexer / sparse_sh_grad_indexer (bool[N]) of
        that step and return (touched rows, queue overflowed, overflowed steps before)."""
        if self._step_ws is None:
            return None
        st, md, msh = self._step_ws.status(want_masks=True)
        self.grid.sparse_grad_indexer, self.grid.sparse_sh_grad_indexer = md, msh
        return st

    def check_overflow(self):
        if self._step_ws is not None:
            st, _, _ = self._step_ws.status()
            if st[1] or st[2]:
                raise RuntimeError("ShardedTrainer: the sample-record queue of the fused step overflowed "
                                   "(ray gradients of those steps were dropped); raise slots_per_ray")

    def step(self, rays, rgb_gt: torch.Tensor, global_batch: Optional[int] = None,
             lr_sigma: Optional[float] = None, lr_sh: Optional[float] = None):
        g = self.grid
        if self._can_fuse() and (global_batch is None or global_batch == rgb_gt.size(0)):
            return self._fused_step(rays, rgb_gt, self.lr_sigma if lr_sigma is None else lr_sigma,
                                    self.lr_sh if lr_sh is None else lr_sh)
        rgb = g.volume_render_fused(rays, rgb_gt, beta_loss=self.beta_loss,
                                    sparsity_loss=self.sparsity_loss, _norm_rays=global_batch)
        if self.tv_sparsity >= 1.0:     # single GPU: the reference's dense TV + dense step
            if self.lambda_tv > 0.0:
                g.inplace_tv_grad(g.density_data.grad, scaling=self.lambda_tv, sparse_frac=self.tv_sparsity)
            if self.lambda_tv_sh > 0.0:
                g.inplace_tv_color_grad(g.sh_data.grad, scaling=self.lambda_tv_sh,
                                        sparse_frac=self.tv_sparsity)
        else:
            if self.lambda_tv > 0.0:
                run, n = self._tv_run()
                g.inplace_tv_grad(g.density_data.grad, scaling=self.lambda_tv, _cell_run=run,
                                  _scale_cells=n)
            if self.lambda_tv_sh > 0.0:
                run, n = self._tv_run()
                g.inplace_tv_color_grad(g.sh_data.grad, scaling=self.lambda_tv_sh, _cell_run=run,
                                        _scale_cells=n)
        lr_sigma = self.lr_sigma if lr_sigma is None else lr_sigma
        lr_sh = self.lr_sh if lr_sh is None else lr_sh
        if g.use_background:
            self._background_step()
        if self.world_size > 1 and g.sh_data.is_cuda:
            self._exchange_and_step_packed(lr_sigma, lr_sh)
            return rgb
        if self.world_size > 1:
            g.sparse_grad_indexer, g.sparse_sh_grad_indexer = exchange_sparse_grads(
                g.density_data.grad, g.sh_data.grad, g.sparse_grad_indexer,
                g.sparse_sh_grad_indexer, group=self.group)
        g.optim_density_step(lr_sigma, beta=self.rms_beta)
        g.optim_sh_step(lr_sh, beta=self.rms_beta)
        return rgb

    def _background_step(self):
        """MSI background (opt.py:566-603): sparse TV on this rank's share of the drawn cells, sum of
        the masked gradient rows over ranks, identical masked RMSProp on every rank."""
        g = self.grid
        grad = g.background_data.grad
        if self.lambda_tv_bg_sigma > 0.0 or self.lambda_tv_bg_color > 0.0:
            cells = g._get_rand_cells_background(self.tv_background_sparsity, contiguous=True)
            n = cells.numel()
            if self.world_size > 1:
                b, e = shard_range(n, self.world_size, self.rank)
                cells = cells[b:e].contiguous()
            g.inplace_tv_background_grad(grad, scaling=self.lambda_tv_bg_color,
                                         scaling_density=self.lambda_tv_bg_sigma, _rand_cells=cells,
                                         _scale_cells=n)
        if self.world_size > 1:
            mask = g.sparse_background_indexer.view(torch.uint8)
            dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=self.group)
            idx = torch.nonzero(mask.view(-1)).flatten()
            flat = grad.view(-1, grad.size(-1))
            packed = flat[idx]
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
            flat[idx] = packed
        g.optim_background_step(self.lr_sigma_bg, self.lr_color_bg, beta=self.rms_beta)

    def _exchange_and_step_packed(self, lr_sigma: float, lr_sh: float):
        """CUDA path of the exchange: OR the masks, pack the union's rows ([rows, 28] = 27 SH
        gradients + the density gradient) with one kernel, ONE NCCL all-reduce, then a single
        masked RMSProp kernel that reads the reduced gradients from the packed buffer (no unpack
        pass).  Same result as exchange_sparse_grads + the two optimiser steps."""
        from . import csrc as _C
        g = self.grid
        mark = self._mark
        mark("local")
        g._ensure_layout()
        masks = torch.stack([g.sparse_grad_indexer, g.sparse_sh_grad_indexer]).view(torch.uint8)
        dist.all_reduce(masks, op=dist.ReduceOp.MAX, group=self.group)
        mark("mask_allreduce")
        rows = torch.nonzero(masks[0] | masks[1]).flatten()     # sorted; one host sync
        mark("row_list")
        D = g.sh_data.size(1)
        # persistent, geometrically grown buffer: a fresh several-hundred-MB tensor per step would
        # miss the caching allocator (the NCCL stream still holds the previous one) and cudaMalloc
        if self._packed is None or self._packed.size(0) < rows.numel():
            self._packed = None
            self._packed = torch.empty((int(rows.numel() * 1.25) + 1024, _C.packed_row_floats(D)),
                                       dtype=torch.float32, device=g.sh_data.device)
        packed = self._packed[:rows.numel()]
        _C.pack_grad_rows(g.sh_data.grad, g.density_data.grad, rows, packed)
        mark("pack")
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        mark("grad_allreduce")
        g._ensure_rms()
        _C.rmsprop_step_packed(g.sh_data.data, g.sh_rms, g.density_data.data, g.density_rms, rows,
                               packed, masks[1], masks[0], self.rms_beta, lr_sh, lr_sigma, 1e-8,
                               grad_sh=g.sh_data.grad, grad_density=g.density_data.grad)
        mark("optimiser")
        self.exchanged_rows = rows.numel()
        g.sparse_grad_indexer, g.sparse_sh_grad_indexer = masks[0].view(torch.bool), masks[1].view(torch.bool)

    # ---- optional phase timing (CUDA events; enable with trainer.profile = True) -------------
    def _mark(self, name):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events.append((name, ev))

    def phase_times_ms(self):
        """Mean milliseconds per phase since profiling was enabled ("local" = march + TV)."""
        torch.cuda.synchronize()
        out, n = {}, {}
        for (_, e0), (name, e1) in zip(self._events[:-1], self._events[1:]):
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
            n[name] = n.get(name, 0) + 1
        return {k: out[k] / n[k] for k in out}
