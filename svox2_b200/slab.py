"""Slab-sharded multi-GPU training step: the GRID is partitioned, not the rays.

Why: with a replicated grid every rank needs every touched row's reduced gradient each step --
at 512^3 / 5000 rays per GPU that is ~110 B x 1.5 M rows PER RANK over NVLink, more than the
march itself costs (dist.py).  Here rank k owns the cells with x in [b_k, b_{k+1})
and the voxel rows of those planes.  Every rank traverses ALL rays of the global batch over the
(replicated, 4 B/voxel) links grid -- so the sample lattice, including the empty-space skips, is
exactly the reference's (render_lerp_kernel_cuvol.cu:56-109) -- but gathers, composites and
back-propagates only the samples inside its slab.  Per step the ranks exchange per-RAY scalars:

  A. all-gather of logT_local[Q]       -> log-transmittance at slab entry (prefix over the slabs
                                          in the order the ray crosses them, i.e. by sign of d_x)
  B. all-gather of rgb_partial[Q,3(+1)] -> total colour, dL/d(rgb), and the value of the
                                          reference's `accum` (cuvol.cu:248-264) at slab entry
  C. halo: samples in the last cell plane of a slab interpolate from the first voxel plane of the
     next one.  That plane is CO-OWNED: the two neighbours swap the gradients / mask bits they
     accumulated for it (one P2P message each way per boundary) and both apply the identical
     update, so no parameters ever travel.

16 B per ray per rank instead of 110 B per touched row: the exchange no longer scales with the
grid.  TV and RMSProp run on the owned rows only.  Everything outside a rank's slab (+ halo
plane) in its copy of density_data / sh_data goes stale; ``gather_params`` re-replicates.

Parity: same samples, same arithmetic per sample; the per-ray log-transmittance is a sum of
per-slab partial sums instead of one sequential sum (~1e-7 relative), covered by the 1e-4 bar.

SHIP MODE (default when the ranks can map each other's memory).  The scheme above makes every rank
traverse the whole global batch -- traversal cost grows with the number of GPUs.  In ship mode the
RAYS are sharded for the traversal: rank r traces rays [r*Q, (r+1)*Q) of the global batch exactly
as the single-GPU kernel does (one sequential transmittance per ray, so the sample set, the stop
decision and the weights are bit-identical to one GPU) and ships every active sample's 96-byte
record to the rank that owns the sample's cell, with plain NVLink stores into region r of that
rank's record queue (svox2_ship_trace).  Exchange A disappears; after one device-side barrier each
rank shades the records it received, exchange B / halo / optimiser are as above.
"""
from typing import List, Optional

import os

import numpy as np
import torch
import torch.distributed as dist

from . import csrc as _C
from .grid import Rays


# "auto" traversal: ship records from this many ranks up.  Measured on 8 B200s (C3, weak scaling,
# profiles/r2_ab_variants.txt): ship 0.976 ms/step vs 0.844 with every rank traversing the whole
# batch -- the records leave the SM as isolated 16-byte NVLink stores (trace 0.115 -> 0.22 ms per
# rank) and the density publication as isolated 4-byte ones (optimiser 0.14 -> 0.28 ms), which costs
# more than the redundant traversal it removes (0.26 ms at 8 ranks).  Ship mode therefore stays
# opt-in (SVOX2_B200_SLAB_TRAVERSAL=ship) until its stores are staged into full 128-byte lines.
SHIP_FROM_WORLD = 1 << 30


def balanced_slabs(links: torch.Tensor, world_size: int) -> List[int]:
    """x boundaries b_0=0 < ... < b_G=X splitting the stored voxels evenly between ranks."""
    X = links.size(0)
    per_plane = (links >= 0).view(X, -1).sum(1).double().cpu().numpy()
    cum = np.concatenate([[0.0], np.cumsum(per_plane)])
    bounds = [0]
    for k in range(1, world_size):
        b = int(np.searchsorted(cum, cum[-1] * k / world_size))
        bounds.append(min(max(b, bounds[-1] + 1), X - (world_size - k)))
    bounds.append(X)
    return bounds


class PeerMemory:
    """One symmetric buffer per rank, all of them mapped into every process (torch symmetric
    memory over NVLink).  Exchanges become: a kernel that stores into the peers' buffers, a
    device-side barrier, a kernel that reads the local buffer -- no collective library call.

    Layout (floats; identical on every rank): A [G,Q] | B [G,Q,4] | halo packets received from the
    left neighbour [H,32] | from the right neighbour [H,32] | row-0 gradient packets [G,32] |
    row-0 parameters published by its owner [32]."""

    WIDTH = 32

    def __init__(self, group, world_size: int, rank: int, device, n_rays: int, halo_rows: int,
                 ship_region_cap: int = 0):
        import torch.distributed._symmetric_memory as symm
        G, Q = world_size, n_rays
        self.world_size, self.rank, self.n_rays = G, rank, Q
        self.off_a, self.off_b = 0, G * Q
        self.off_hl = self.off_b + G * Q * 4
        self.off_hr = self.off_hl + halo_rows * self.WIDTH
        self.off_z = self.off_hr + halo_rows * self.WIDTH      # [G, 32]: row-0 packets (see _row0_*)
        self.off_p = self.off_z + G * self.WIDTH               # [32]: row-0 parameters (see _publish_row0)
        total = self.off_p + self.WIDTH
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._views = {}        # tensor views of the inboxes (hdl.get_buffer is not free: once per shape)
        self.buf.zero_()
        # ship mode: one record workspace per rank, every rank's mapped into every process
        self.ship = None
        if ship_region_cap > 0:
            self.ship = _C.ShipWorkspace(Q, G, rank, ship_region_cap)
            self.ship_buf = symm.empty(self.ship.nbytes, dtype=torch.uint8, device=device)
            self.ship_hdl = symm.rendezvous(self.ship_buf, group if group is not None else dist.group.WORLD)
            self.ship.bind(self.ship_buf, self.ship_hdl.buffer_ptrs)
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)

    def all_gather(self, local: torch.Tensor, region: str) -> torch.Tensor:
        """local [Q] (region "a") or [Q,4] (region "b") -> [G,Q(,4)] of every rank."""
        G, Q = self.world_size, self.n_rays
        off, per = (self.off_a, Q) if region == "a" else (self.off_b, Q * 4)
        _C.peer_store(self.ptrs, (off + self.rank * per) * 4, local)
        self.hdl.barrier(channel=1 if region == "a" else 2)
        view = self.buf[off:off + G * per]
        return view.view(G, Q) if region == "a" else view.view(G, Q, 4)

    def neighbour_inbox(self, peer: int, from_left: bool, rows: int) -> torch.Tensor:
        """The [rows,32] region of ``peer``'s buffer that receives packets from its left / right."""
        key = ("n", peer, from_left, rows)
        if key not in self._views:
            off = self.off_hl if from_left else self.off_hr
            self._views[key] = self.hdl.get_buffer(peer, (rows, self.WIDTH), torch.float32, off)
        return self._views[key]

    def inbox(self, from_left: bool, rows: int) -> torch.Tensor:
        key = ("i", from_left, rows)
        if key not in self._views:
            off = self.off_hl if from_left else self.off_hr
            self._views[key] = self.buf[off:off + rows * self.WIDTH].view(rows, self.WIDTH)
        return self._views[key]


class SlabTrainer:
    """opt/opt.py's train step (opt.py:474-603) with the grid sharded by x-slabs.

    ``step(rays, rgb_gt)`` takes the GLOBAL batch on every rank and returns its colours."""

    def __init__(self, grid, world_size: int = 1, rank: int = 0, lr_sigma: float = 30.0,
                 lr_sh: float = 1e-2, lambda_tv: float = 1e-5, lambda_tv_sh: float = 1e-3,
                 tv_sparsity: float = 0.01, rms_beta: float = 0.95, beta_loss: float = 0.0,
                 sparsity_loss: float = 0.0, group=None, bounds: Optional[List[int]] = None):
        assert not grid.use_background, "slab mode does not support an MSI background"
        if tv_sparsity >= 1.0 and (lambda_tv > 0.0 or lambda_tv_sh > 0.0):
            # the reference switches to the dense TV kernel and a dense optimiser step there
            # (svox2.py:1760-1766); the slab step is built on the sparse contiguous cell run
            raise ValueError("SlabTrainer needs tv_sparsity < 1 (sparse contiguous TV run)")
        self.grid, self.world_size, self.rank, self.group = grid, world_size, rank, group
        self.lr_sigma, self.lr_sh = lr_sigma, lr_sh
        self.lambda_tv, self.lambda_tv_sh = lambda_tv, lambda_tv_sh
        self.tv_sparsity, self.rms_beta = tv_sparsity, rms_beta
        self.beta_loss, self.sparsity_loss = beta_loss, sparsity_loss
        self.bounds = bounds if bounds is not None else balanced_slabs(grid.links, world_size)
        assert len(self.bounds) == world_size + 1
        self.x_lo, self.x_hi = self.bounds[rank], self.bounds[rank + 1]
        links = grid.links
        dev = links.device
        self.own_mask = torch.zeros(grid.density_data.size(0), dtype=torch.bool, device=dev)
        own = links[self.x_lo:self.x_hi].reshape(-1)
        self.own_mask[own[own >= 0].long()] = True
        # boundary planes, co-owned: first voxel plane of the right neighbour / my own first plane
        self.halo_right = self._plane_rows(self.x_hi) if rank < world_size - 1 else None
        self.halo_left = self._plane_rows(self.x_lo) if rank > 0 else None
        self.excl_mask = self.own_mask.clone()          # exclusive ownership (gather_params)
        if self.halo_right is not None:
            self.own_mask[self.halo_right] = True
        # Reference quirk: TV cells on the grid faces read row 0 for their out-of-range
        # neighbour and add gradient / mask bits to it.  Such cells lie in every slab, row 0 in one.
        self._row0 = torch.zeros(1, dtype=torch.int64, device=dev)
        self._holds_row0 = bool(self.own_mask[0].item()) if self.own_mask.numel() else True
        # its exclusive owner (known to every rank from the replicated links) publishes the updated
        # parameters of row 0 after each step, so that the face cells of the other slabs keep
        # reading the value the single-GPU step would read
        at = (links == 0).nonzero()
        x0 = int(at[0, 0].item()) if at.numel() else 0
        self._row0_owner = max(k for k in range(world_size) if self.bounds[k] <= x0)
        self._row0_pending = False      # a published row 0 waits in the peer buffer
        self._ws = None
        self.slots_per_ray = None  # record slots per ray of the workspace (None: min(diagonal, 512))
        self._peer = None          # PeerMemory for the current batch size, False if unavailable
        self._density_ptrs = None  # ship mode: every rank's density_data, mapped into this process
        # the step's gradients live in a compact [touched rows, 28] buffer (step.cu) instead of the
        # dense .grad arrays; SVOX2_B200_SLAB_SLOTS=0 keeps the dense-gradient step (also the path
        # without peer memory)
        self.use_slots = os.environ.get("SVOX2_B200_SLAB_SLOTS", "1") != "0"
        self._mg = None
        self._scr = None
        rows = [h for h in (self.halo_left, self.halo_right) if h is not None] + [self._row0]
        self._boundary_rows = torch.unique(torch.cat(rows)).contiguous()
        self.use_peer_memory = os.environ.get("SVOX2_B200_PEER", "1") != "0"
        # "ship": rays sharded for the traversal, records shipped to the owning slab (needs peer
        # memory and a batch divisible by the world size); "replicated": every rank traverses all rays
        # "auto": see SHIP_FROM_WORLD
        self.traversal = os.environ.get("SVOX2_B200_SLAB_TRAVERSAL", "auto")
        if self.traversal == "auto":
            self.traversal = "ship" if world_size >= SHIP_FROM_WORLD else "replicated"
        self._steps = 0
        # clear + trace, prefix, shade, combine, seed + scatter, TV x2, halo pack/unpack, optimiser x2
        self.kernels_per_step = 2 + 1 + 1 + 1 + 2 + (lambda_tv > 0) + (lambda_tv_sh > 0) + 2 + 2
        self.profile, self._events = False, []

    def _plane_rows(self, x: int) -> torch.Tensor:
        p = self.grid.links[x].reshape(-1)
        return p[p >= 0].long().contiguous()

    # ---- helpers ------------------------------------------------------------------------
    def _workspace(self, n_rays: int):
        g = self.grid
        diag = sum(float(s) ** 2 for s in g.links.shape) ** 0.5
        slots = max(64, min(int(diag / max(float(g.opt.step_size), 1e-3)) + 2, 512))
        if self.slots_per_ray is not None:
            slots = max(64, int(self.slots_per_ray))
        if self._ws is None or self._ws[0] != (n_rays, slots):
            self._ws = ((n_rays, slots), _C.slab_workspace(n_rays, slots, g.links.device))
        return self._ws[1]

    def _peer_memory(self, n_rays: int):
        """Symmetric-memory exchange buffers for this batch size, or None (-> NCCL collectives)."""
        if self.world_size == 1 or not self.use_peer_memory or n_rays % 4 != 0:
            return None
        if self._peer is None or (self._peer is not False and self._peer.n_rays != n_rays):
            try:
                halo = torch.tensor([max([h.numel() for h in (self.halo_left, self.halo_right)
                                          if h is not None] + [1])], device=self.grid.links.device)
                dist.all_reduce(halo, op=dist.ReduceOp.MAX, group=self.group)
                cap = 0
                if self.traversal == "ship" and n_rays % self.world_size == 0:
                    g = self.grid
                    diag = sum(float(s) ** 2 for s in g.links.shape) ** 0.5
                    slots = max(64, min(int(diag / max(float(g.opt.step_size), 1e-3)) + 2, 512))
                    if self.slots_per_ray is not None:
                        slots = max(64, int(self.slots_per_ray))
                    # region r of a rank's queue holds what rank r traced inside this rank's slab:
                    # at most all of its samples; the whole queue is capped at 2 GiB per rank
                    cap = min(slots * (n_rays // self.world_size), (2 << 30) // 96 // self.world_size)
                self._peer = PeerMemory(self.group, self.world_size, self.rank,
                                        self.grid.links.device, n_rays, int(halo.item()), cap)
                if self._peer.ship is not None and self._density_ptrs is None:
                    self._bind_symmetric_density()
            except Exception as e:      # no symmetric-memory support on this system / build
                if self.rank == 0:
                    print(f"SlabTrainer: peer memory unavailable ({type(e).__name__}: {e}); using NCCL")
                self._peer = False
        return self._peer or None

    def _bind_symmetric_density(self):
        """Ship mode keeps density_data replicated: re-home it in symmetric memory so that the owners
        can store their updated rows straight into every rank's copy (svox2_rmsprop_step_publish /
        svox2_mg_step_finish)."""
        import torch.distributed._symmetric_memory as symm
        g = self.grid
        old = g.density_data.data
        buf = symm.empty(old.numel(), dtype=torch.float32, device=old.device)
        self._density_hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        buf.copy_(old.reshape(-1))
        requires_grad = g.density_data.requires_grad
        g.density_data = torch.nn.Parameter(buf.view(-1, 1), requires_grad=requires_grad)
        self._density_buf = buf
        self._density_ptrs = [int(p) for p in self._density_hdl.buffer_ptrs]
        # the first voxel plane of a slab is co-owned with the left neighbour, which applies the
        # identical update to its own copy: publishing those rows to it would race with that update
        self._send = self.excl_mask.to(torch.uint8)          # 1: publish to every peer
        if self.halo_left is not None:
            self._send[self.halo_left] = 2                   # 2: every peer but the left neighbour
        torch.cuda.synchronize()
        self._peer.hdl.barrier(channel=0)

    def _density_step_publish(self, lr):
        """Masked RMSProp on the (replicated) density array with the publication of the updated rows
        fused in (NVLink stores; ordered before anybody's next traversal by the barrier that opens
        the next step)."""
        g = self.grid
        g._ensure_rms()
        _C.rmsprop_step_publish(g.density_data.data, g.density_rms, g.density_data.grad, g.sparse_grad_indexer,
                                self._send, self._density_ptrs, self.rank, self.rank - 1, self.rms_beta, lr, 1e-8)

    def _gather(self, t: torch.Tensor, region: str) -> torch.Tensor:
        """[G, *t.shape]: t of every rank (this rank's own row is t itself)."""
        if self.world_size == 1:
            return t.unsqueeze(0)
        peer = self._peer_memory(t.size(0))
        if peer is not None:
            return peer.all_gather(t, region)
        out = torch.empty((self.world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _tv_runs(self):
        """This rank's part of the step's wrapped contiguous cell run (x-major linear index)."""
        g = self.grid
        start, n = g._draw_cell_run(self.tv_sparsity)
        X, Y, Z = g.links.shape
        total = X * Y * Z
        lo, hi = self.x_lo * Y * Z, self.x_hi * Y * Z
        pieces = [(start, min(start + n, total))]
        if start + n > total:
            pieces.append((0, start + n - total))
        runs = []
        for a, b in pieces:
            a, b = max(a, lo), min(b, hi)
            if b > a:
                runs.append((a, b - a))
        return runs, n

    def _mark(self, name):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events.append((name, ev))

    def phase_times_ms(self):
        torch.cuda.synchronize()
        out, n = {}, {}
        for (_, e0), (name, e1) in zip(self._events[:-1], self._events[1:]):
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
            n[name] = n.get(name, 0) + 1
        return {k: out[k] / n[k] for k in out}

    # ---- the step -----------------------------------------------------------------------
    def step(self, rays: Rays, rgb_gt: torch.Tensor, lr_sigma: Optional[float] = None,
             lr_sh: Optional[float] = None) -> torch.Tensor:
        g = self.grid
        g._require_cuda("SlabTrainer.step")
        g._ensure_layout()
        Q = rgb_gt.size(0)
        dev = rgb_gt.device
        peer = self._peer_memory(Q) if self.world_size > 1 else None
        if self.use_slots and (self.world_size == 1 or
                               (peer is not None and _C.packed_row_floats(g.sh_data.size(1) + 2) == peer.WIDTH)):
            return self._step_slots(rays, rgb_gt, lr_sigma, lr_sh, peer)
        if peer is not None and peer.ship is not None:
            return self._step_ship(rays, rgb_gt, lr_sigma, lr_sh, peer)
        ws = self._workspace(Q)
        spec, ropt, rspec = g._to_cpp(), g.opt._to_cpp(), rays._to_cpp()
        self._mark("start")

        # 1. traversal + this slab's transmittance
        logT_local = torch.empty(Q, dtype=torch.float32, device=dev)
        _C.slab_trace(spec, rspec, ropt, self.x_lo, self.x_hi, logT_local, ws)
        self._mark("trace")
        logT_all = self._gather(logT_local, "a")                                 # exchange A
        self._refresh_row0()            # the barrier of exchange A made the owner's store visible
        lt_in = torch.empty(Q, dtype=torch.float32, device=dev)
        _C.slab_prefix(logT_all, rays.dirs, self.rank, lt_in)
        self._mark("exchange_a")

        # 2. colours of this slab's samples
        part = torch.zeros((Q, 4), dtype=torch.float32, device=dev)          # rgb, stop flag
        _C.slab_shade(spec, Q, ropt, lt_in, part, ws)
        self._mark("shade")
        part_all = self._gather(part, "b")                                       # exchange B
        rgb = torch.empty((Q, 3), dtype=torch.float32, device=dev)
        gout = torch.empty((Q, 3), dtype=torch.float32, device=dev)
        accum_start = torch.empty(Q, dtype=torch.float32, device=dev)
        _C.slab_combine(part_all, logT_all, rays.dirs, rgb_gt, self.rank,
                        g.opt.background_brightness, self.beta_loss, rgb, gout, accum_start)
        self._mark("exchange_b")

        # 3. gradients of this slab's samples, TV on the owned cells
        grad_density, grad_sh, _ = g._get_data_grads()
        g.sparse_grad_indexer = torch.zeros(g.density_data.size(0), dtype=torch.bool, device=dev)
        holder = _C.GridOutputGrads()
        holder.grad_density_out, holder.grad_sh_out = grad_density, grad_sh
        holder.mask_out = g.sparse_grad_indexer
        _C.slab_backward(spec, Q, ropt, lt_in, gout, accum_start, self.sparsity_loss, holder, ws)
        g.sparse_sh_grad_indexer = g.sparse_grad_indexer.clone()
        self._mark("backward")
        return self._finish_step(rgb, lr_sigma, lr_sh)

    def _mg_step(self, n_rays: int, peer):
        g = self.grid
        ship = peer.ship if peer is not None else None
        key = (n_rays, g.sh_data.size(0), id(ship))
        if self._mg is None or self._mg[0] != key:
            diag = sum(float(s) ** 2 for s in g.links.shape) ** 0.5
            slots = max(64, min(int(diag / max(float(g.opt.step_size), 1e-3)) + 2, 512))
            if self.slots_per_ray is not None:
                slots = max(64, int(self.slots_per_ray))
            if ship is not None:
                slots = 64          # the records live in the ship workspace; keep the unused queue small
            ws = _C.TrainStepWorkspace(g.sh_data.size(0), g.basis_dim, n_rays, slots, g.links.device)
            self._mg = (key, _C.MgStep(ws, self.world_size, self.rank, self.x_lo, self.x_hi, ship, self.bounds))
        return self._mg[1]

    def export_masks(self):
        """After a slot-based step: materialise grid.sparse_grad_indexer / sparse_sh_grad_indexer
        (bool[N], this rank's owned rows) of that step; returns (touched rows, overflow, sticky)."""
        if self._mg is None:
            return None
        st, md, msh = self._mg[1].ws.status(want_masks=True)
        self.grid.sparse_grad_indexer, self.grid.sparse_sh_grad_indexer = md, msh
        return st

    def _step_slots(self, rays: Rays, rgb_gt: torch.Tensor, lr_sigma, lr_sh, peer) -> torch.Tensor:
        """The step on the compact gradient buffer (include/svox2_b200.h svox2_mg_step_*): the
        single-GPU fused step cut where the slabs exchange per-ray scalars and boundary planes."""
        g = self.grid
        Qg, G = rgb_gt.size(0), self.world_size
        dev = rgb_gt.device
        ship = peer is not None and peer.ship is not None
        mg = self._mg_step(Qg, peer)
        lr_d = self.lr_sigma if lr_sigma is None else lr_sigma
        lr_s = self.lr_sh if lr_sh is None else lr_sh
        # the step's two TV cell runs, drawn in the reference's order (density, then SH)
        runs_d, n_d = self._tv_runs() if self.lambda_tv > 0.0 else ([], 0)
        runs_s, n_s = self._tv_runs() if self.lambda_tv_sh > 0.0 else ([], 0)
        mg.set_step(runs_d, n_d, self.lambda_tv, runs_s, n_s, self.lambda_tv_sh, self.sparsity_loss, lr_d, lr_s,
                    self.rms_beta)
        # clear, TV marking, boundary rows, 2 rank kernels, seed, scatter, TV x2, RMSProp + march / exchange kernels
        self.kernels_per_step = 12 + (3 if ship else 4) + (6 if G > 1 else 0)
        spec, ropt = g._to_cpp(), g.opt._to_cpp()
        if ship:
            peer.hdl.barrier(channel=0)     # everybody's density rows of the last step have landed
            self._refresh_row0()
        self._mark("start")
        sc = self._scratch(Qg, dev)
        logT_local = None if ship else sc["logT_local"]
        mg.begin(spec, rays._to_cpp(), ropt, self._boundary_rows, logT_local)
        self._mark("trace")
        lt_in = logT_all = None
        if ship:
            peer.hdl.barrier(channel=1)     # the records shipped to me have landed
        else:
            logT_all = self._gather(logT_local, "a")                             # exchange A
            self._refresh_row0()
            lt_in = sc["lt_in"]
            _C.slab_prefix(logT_all, rays.dirs, self.rank, lt_in)
        self._mark("exchange_a")
        part = sc["part"].zero_()
        mg.shade(spec, ropt, lt_in, part)
        self._mark("shade")
        part_all = self._gather(part, "b")                                       # exchange B
        rgb = torch.empty((Qg, 3), dtype=torch.float32, device=dev)     # handed to the caller: fresh
        gout, accum_start = sc["gout"], sc["accum_start"]
        if ship:
            _C.ship_combine(part_all, peer.ship.ray_logT(), rays.dirs, rgb_gt, self.rank,
                            g.opt.background_brightness, self.beta_loss, rgb, gout, accum_start)
        else:
            _C.slab_combine(part_all, logT_all, rays.dirs, rgb_gt, self.rank,
                            g.opt.background_brightness, self.beta_loss, rgb, gout, accum_start)
        self._mark("exchange_b")
        mg.backward(spec, ropt, lt_in, gout, accum_start)                         # + TV into the slots
        self._mark("backward")
        if G > 1:
            self._halo_slots(mg, peer)
        self._mark("halo_grads")
        g._ensure_rms()
        mg.finish(spec, g.density_rms, g.sh_rms, self._send if ship else None,
                  self._density_ptrs if ship else None)
        self._publish_row0()
        self._mark("optimiser")
        g.sparse_grad_indexer = g.sparse_sh_grad_indexer = None      # slot flags; export_masks() on demand
        if self._steps % 256 == 0:
            self.check_overflow()
        self._steps += 1
        return rgb

    def _scratch(self, n_rays: int, dev):
        """Per-ray scratch of the slot-based step, allocated once per batch size."""
        if self._scr is None or self._scr[0] != n_rays:
            f = dict(dtype=torch.float32, device=dev)
            self._scr = (n_rays, {
                "logT_local": torch.empty(n_rays, **f), "lt_in": torch.empty(n_rays, **f),
                "part": torch.zeros((n_rays, 4), **f), "gout": torch.empty((n_rays, 3), **f),
                "accum_start": torch.empty(n_rays, **f), "z": torch.zeros((1, PeerMemory.WIDTH), **f)})
        return self._scr[1]

    def _halo_slots(self, mg, peer):
        """The halo swap of _halo_grads in slot space: boundary planes and the row-0 packet go from
        the slot rows straight into the neighbours' inboxes, and what arrives is added to the slots."""
        width = peer.WIDTH
        z = self._scr[1]["z"]           # all zeros at the row-0 holder; fully rewritten by the pack elsewhere
        if not self._holds_row0:
            mg.slot_pack(self._row0, z, clear=True)
        _C.peer_store(peer.ptrs, (peer.off_z + self.rank * width) * 4, z)
        nl = self.halo_left.numel() if self.halo_left is not None else 0
        nr = self.halo_right.numel() if self.halo_right is not None else 0
        if nr:
            mg.slot_pack(self.halo_right, peer.neighbour_inbox(self.rank + 1, True, nr))
        if nl:
            mg.slot_pack(self.halo_left, peer.neighbour_inbox(self.rank - 1, False, nl))
        peer.hdl.barrier(channel=3)
        if nl:
            mg.slot_unpack_add(peer.inbox(True, nl), self.halo_left)
        if nr:
            mg.slot_unpack_add(peer.inbox(False, nr), self.halo_right)
        if self._holds_row0:
            zs = peer.buf[peer.off_z:peer.off_z + self.world_size * width].view(self.world_size, width)
            mg.slot_unpack_add(zs.sum(0, keepdim=True), self._row0)

    def _step_ship(self, rays: Rays, rgb_gt: torch.Tensor, lr_sigma, lr_sh, peer) -> torch.Tensor:
        """Ship mode: trace my share of the rays, ship the records to their slabs, then shade /
        back-propagate what I received (module docstring)."""
        g = self.grid
        Qg, G = rgb_gt.size(0), self.world_size
        Q = Qg // G
        dev = rgb_gt.device
        ws = peer.ship
        # basis + trace + publish, shade, peer store, combine, seed + scatter + reset, TV x2,
        # halo pack/unpack, optimiser x2
        self.kernels_per_step = 3 + 1 + 1 + 1 + 3 + (self.lambda_tv > 0) + (self.lambda_tv_sh > 0) + 2 + 2
        spec, ropt = g._to_cpp(), g.opt._to_cpp()
        # 0. everybody's density rows of the last step have landed in my copy
        peer.hdl.barrier(channel=0)
        self._refresh_row0()
        self._mark("start")
        # 1. traversal of my rays; records land in the owners' queues (NVLink stores)
        _C.ship_trace(spec, rays._to_cpp(), ropt, self.rank * Q, Q, self.bounds, ws)
        self._mark("trace")
        peer.hdl.barrier(channel=1)
        self._mark("exchange_a")
        # 2. colours of the samples inside my slab, whoever traced them
        part = torch.zeros((Qg, 4), dtype=torch.float32, device=dev)
        _C.ship_shade(spec, part, ws)
        self._mark("shade")
        part_all = self._gather(part, "b")                                       # exchange B
        rgb = torch.empty((Qg, 3), dtype=torch.float32, device=dev)
        gout = torch.empty((Qg, 3), dtype=torch.float32, device=dev)
        accum_start = torch.empty(Qg, dtype=torch.float32, device=dev)
        _C.ship_combine(part_all, ws.ray_logT(), rays.dirs, rgb_gt, self.rank,
                        g.opt.background_brightness, self.beta_loss, rgb, gout, accum_start)
        self._mark("exchange_b")
        # 3. gradients of my slab's samples; the queue is reset behind the scatter (the step's last
        #    barrier, in the halo swap, orders that reset before anybody's next trace)
        grad_density, grad_sh, _ = g._get_data_grads()
        g.sparse_grad_indexer = torch.zeros(g.density_data.size(0), dtype=torch.bool, device=dev)
        holder = _C.GridOutputGrads()
        holder.grad_density_out, holder.grad_sh_out = grad_density, grad_sh
        holder.mask_out = g.sparse_grad_indexer
        _C.ship_backward(spec, ropt, gout, accum_start, self.sparsity_loss, holder, ws)
        g.sparse_sh_grad_indexer = g.sparse_grad_indexer.clone()
        self._mark("backward")
        return self._finish_step(rgb, lr_sigma, lr_sh)

    def _finish_step(self, rgb, lr_sigma, lr_sh):
        """TV on the owned cells, halo swap, RMSProp on the owned rows."""
        g = self.grid
        if self.lambda_tv > 0.0:
            runs, n = self._tv_runs()
            for run in runs:
                g.inplace_tv_grad(g.density_data.grad, scaling=self.lambda_tv, _cell_run=run,
                                  _scale_cells=n)
        if self.lambda_tv_sh > 0.0:
            runs, n = self._tv_runs()
            for run in runs:
                g.inplace_tv_color_grad(g.sh_data.grad, scaling=self.lambda_tv_sh, _cell_run=run,
                                        _scale_cells=n)
        self._mark("tv")

        # 4. boundary planes: both neighbours add each other's gradients (a + b == b + a, so the
        #    two copies of a co-owned plane stay bit-identical), then RMSProp on the owned rows
        lr_d = self.lr_sigma if lr_sigma is None else lr_sigma
        lr_s = self.lr_sh if lr_sh is None else lr_sh
        ship = self._density_ptrs is not None and self._peer not in (None, False) and self._peer.ship is not None
        self._halo_grads()
        if not self._holds_row0:    # the only row outside own_mask a step can touch (TV face cells)
            g.sparse_grad_indexer[0] = False
            g.sparse_sh_grad_indexer[0] = False
        self._mark("halo_grads")
        if ship:
            self._density_step_publish(lr_d)
        else:
            g.optim_density_step(lr_d, beta=self.rms_beta)
        g.optim_sh_step(lr_s, beta=self.rms_beta)
        self._publish_row0()
        self._mark("optimiser")

        if self._steps % 256 == 0:
            self.check_overflow()
        self._steps += 1
        return rgb

    def check_overflow(self):
        """Raise if any step since the last check ran out of sample-record space (its colours and
        gradients were dropped).  The trace kernel keeps a sticky count in the workspace, so the
        check can be made at any host sync point (every 256 steps, logging, gather_params)."""
        ship = self._peer.ship if self._peer not in (None, False) else None
        slots_ov = False
        if self._mg is not None:
            st, _, _ = self._mg[1].ws.status()
            slots_ov = bool(st[1] or st[2])
        if (self._ws is not None and _C.slab_overflowed(self._ws[1])) or (ship is not None and ship.overflowed()) \
                or slots_ov:
            raise RuntimeError(
                "SlabTrainer: the sample-record workspace overflowed in at least one step since "
                "the last check (those steps contributed no colour / gradient); raise "
                "SlabTrainer.slots_per_ray")

    def _publish_row0(self):
        """Row-0 parameters -> every rank (reference quirk B5: TV face cells read row 0 wherever
        they are).  Peer memory: the owner stores [sh, density] of row 0 into every rank's buffer;
        readers pick it up behind the next step's first barrier.  NCCL: one tiny broadcast."""
        if self.world_size == 1:
            return
        g = self.grid
        width = _C.packed_row_floats(g.sh_data.size(1) + 2)
        peer = self._peer if (self._peer not in (None, False)) else None
        z = torch.empty((1, width), dtype=torch.float32, device=g.sh_data.device)
        if self.rank == self._row0_owner:       # the pack kernel writes every float of the packet
            _C.pack_grad_rows(g.sh_data.data, g.density_data.data, self._row0, z)
        if peer is not None and width == peer.WIDTH:
            if self.rank == self._row0_owner:
                _C.peer_store(peer.ptrs, peer.off_p * 4, z)
            self._row0_pending = True
            return
        dist.broadcast(z, src=self._row0_owner, group=self.group)
        if not self._holds_row0:
            _C.unpack_rows(z, self._row0, g.sh_data.data, g.density_data.data, False)

    def _refresh_row0(self):
        if not self._row0_pending:
            return
        self._row0_pending = False
        peer = self._peer if (self._peer not in (None, False)) else None
        if peer is None or self._holds_row0:
            return
        g = self.grid
        slot = peer.buf[peer.off_p:peer.off_p + peer.WIDTH].view(1, peer.WIDTH)
        _C.unpack_rows(slot, self._row0, g.sh_data.data, g.density_data.data, False)

    def _halo_grads(self):
        """The first voxel plane of every slab is co-owned with the left neighbour (its last cell
        plane interpolates from it).  Both sides send what they accumulated for the plane --
        gradients and mask bits, packed 32 floats per row -- and add what they receive."""
        if self.world_size == 1:
            return
        g = self.grid
        gsh, gd = g.sh_data.grad, g.density_data.grad
        md, msh = g.sparse_grad_indexer, g.sparse_sh_grad_indexer
        width = _C.packed_row_floats(gsh.size(1) + 2)
        peer = self._peer if (self._peer not in (None, False)) else None
        # row 0: ranks that do not hold it publish (and clear) what their face cells added to it
        z = torch.zeros((1, width), dtype=torch.float32, device=gsh.device)
        if not self._holds_row0:
            _C.pack_grad_rows(gsh, gd, self._row0, z, mask_a=md, mask_b=msh)
            gsh[0] = 0.0
            gd[0] = 0.0
        if peer is not None and width == peer.WIDTH:
            _C.peer_store(peer.ptrs, (peer.off_z + self.rank * width) * 4, z)
            # pack straight into the neighbour's inbox (NVLink stores), barrier, add my inbox
            if self.halo_right is not None and self.halo_right.numel():
                _C.pack_grad_rows(gsh, gd, self.halo_right,
                                  peer.neighbour_inbox(self.rank + 1, True, self.halo_right.numel()),
                                  mask_a=md, mask_b=msh)
            if self.halo_left is not None and self.halo_left.numel():
                _C.pack_grad_rows(gsh, gd, self.halo_left,
                                  peer.neighbour_inbox(self.rank - 1, False, self.halo_left.numel()),
                                  mask_a=md, mask_b=msh)
            peer.hdl.barrier(channel=3)
            if self.halo_left is not None and self.halo_left.numel():
                _C.unpack_rows(peer.inbox(True, self.halo_left.numel()), self.halo_left, gsh, gd, True,
                               mask_a=md, mask_b=msh)
            if self.halo_right is not None and self.halo_right.numel():
                _C.unpack_rows(peer.inbox(False, self.halo_right.numel()), self.halo_right, gsh, gd, True,
                               mask_a=md, mask_b=msh)
            if self._holds_row0:
                zs = peer.buf[peer.off_z:peer.off_z + self.world_size * width].view(self.world_size, width)
                _C.unpack_rows(zs.sum(0, keepdim=True), self._row0, gsh, gd, True, mask_a=md, mask_b=msh)
            return
        dist.all_reduce(z, op=dist.ReduceOp.SUM, group=self.group)
        if self._holds_row0:
            _C.unpack_rows(z, self._row0, gsh, gd, True, mask_a=md, mask_b=msh)
        ops, todo = [], []
        for rows, peer in ((self.halo_right, self.rank + 1), (self.halo_left, self.rank - 1)):
            if rows is None or rows.numel() == 0:
                continue
            send = torch.empty((rows.numel(), width), dtype=torch.float32, device=gsh.device)
            recv = torch.empty_like(send)
            _C.pack_grad_rows(gsh, gd, rows, send, mask_a=md, mask_b=msh)
            ops.append(dist.P2POp(dist.isend, send, peer, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv, peer, group=self.group))
            todo.append((rows, recv))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for rows, recv in todo:
            _C.unpack_rows(recv, rows, gsh, gd, True, mask_a=md, mask_b=msh)

    def gather_params(self):
        """Re-replicate: every rank receives the owners' density / SH rows (one-off, e.g. before
        rendering or saving)."""
        self.check_overflow()
        if self.world_size == 1:
            return
        g = self.grid
        for p in (g.density_data.data, g.sh_data.data):
            p.mul_(self.excl_mask.unsqueeze(-1).to(p.dtype))
            # padded SH rows: reduce the whole [N, row_stride] storage behind the [:, :27] view
            flat = p if p.is_contiguous() else p.as_strided((p.size(0), p.stride(0)), (p.stride(0), 1))
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
