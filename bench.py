#!/usr/bin/env python
"""Benchmark of the Plenoxels hot path: train rays/s of the fused train step on a 512^3 SH-deg-2
sparse grid (BASELINE.json metric, config C3 of SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm, rank 0 only)

A step = SparseGrid.volume_render_fused (fused fwd+bwd ray march, MSE gradient) +
inplace_tv_grad + inplace_tv_color_grad + optim_density_step + optim_sh_step on one batch of
5000 rays per GPU (syn.json hyper-parameters; opt/opt.py:474-603), replicated grid, rays
sharded over ranks, masked gradient exchange over NCCL when N>1.  Prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 5000                      # rays per GPU per step (opt/opt.py batch_size)
POOL = 8                          # distinct ray batches cycled through (no L2-resident reuse)
SYN = dict(lr_sigma=30.0, lr_sh=1e-2, lambda_tv=1e-5, lambda_tv_sh=1e-3, tv_sparsity=0.01,
           rms_beta=0.95)         # opt/configs/syn.json + opt.py defaults
RENDER = dict(step_size=0.5, sigma_thresh=1e-8, stop_thresh=1e-7, background_brightness=1.0)
WORKLOADS = {
    "c3": dict(reso=512, occupancy=0.10, name="C3: 512^3 sparse (~10% occupancy) SH deg 2"),
    "c2": dict(reso=256, occupancy=0.25, name="C2-like: 256^3 sparse SH deg 2"),
    "small": dict(reso=128, occupancy=0.25, name="dev: 128^3 sparse SH deg 2"),
}


# ----------------------------------------------------------------------------------------------
# synthetic scene (identical for both arms; torch ops only, seeded)
# ----------------------------------------------------------------------------------------------
def _inside_blobs(reso, centers, radii, scale, device):
    ax = (torch.arange(reso, device=device, dtype=torch.float32) + 0.5) / reso * 2 - 1
    inside = torch.zeros((reso, reso, reso), dtype=torch.bool, device=device)
    X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    for c, r in zip(centers.tolist(), radii.tolist()):
        d2 = ((X - c[0]) / (r[0] * scale)) ** 2 + ((Y - c[1]) / (r[1] * scale)) ** 2 + \
             ((Z - c[2]) / (r[2] * scale)) ** 2
        inside |= d2 < 1.0
    return inside


def _dilate(mask, iters):
    m = mask[None, None].float()
    for _ in range(iters):
        m = torch.nn.functional.max_pool3d(m, 3, stride=1, padding=1)
    return m[0, 0] > 0


def build_scene(workload, device, seed=20200823):
    """sigma = 30 inside a union of seeded ellipsoids, SH ~ N(0, 0.1^2) with DC 0.5, voxels kept =
    dilate^2(occupied); blob scale bisected so that the stored fraction hits the target."""
    cfg = WORKLOADS[workload]
    reso, target = cfg["reso"], cfg["occupancy"]
    gen = torch.Generator().manual_seed(seed)
    nb = 40
    centers = (torch.rand(nb, 3, generator=gen) - 0.5) * 1.3
    radii = torch.rand(nb, 3, generator=gen) * 0.17 + 0.08
    lo, hi = 0.2, 3.0
    for _ in range(10):
        mid = 0.5 * (lo + hi)
        keep = _dilate(_inside_blobs(reso, centers, radii, mid, device), 2)
        frac = keep.float().mean().item()
        lo, hi = (mid, hi) if frac < target else (lo, mid)
    scale = 0.5 * (lo + hi)
    inside = _inside_blobs(reso, centers, radii, scale, device)
    keep = _dilate(inside, 2)
    links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
    links[~keep.view(-1)] = -1
    links = links.view(reso, reso, reso).contiguous()
    n_rows = int(keep.sum().item())
    density = (inside[keep].float() * 30.0).view(n_rows, 1).contiguous()
    ggen = torch.Generator(device=device).manual_seed(seed + 1)
    sh = torch.randn(n_rows, 27, generator=ggen, device=device) * 0.1
    sh[:, ::9] += 0.5
    return dict(links=links, density=density, sh=sh.contiguous(), n_rows=n_rows, reso=reso,
                occupancy=n_rows / reso ** 3, name=cfg["name"])


def build_rays(n_total, device, seed=20200823):
    """NeRF-synthetic-shaped rays: 100 cameras on the r=2.687 upper hemisphere looking at the
    origin, 800x800, fx=1111.11 (OpenCV convention); random (camera, pixel) per ray."""
    gen = torch.Generator().manual_seed(seed + 7)
    ncam = 100
    u = torch.rand(ncam, generator=gen)
    phi = torch.rand(ncam, generator=gen) * 2 * math.pi
    zc = 0.1 + 0.85 * u
    rxy = torch.sqrt(1 - zc * zc)
    pos = torch.stack([rxy * torch.cos(phi), rxy * torch.sin(phi), zc], -1).double() * 2.687
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    R = torch.stack([right, down, fwd], dim=-1)                 # columns: x, y, z of the camera
    cam = torch.randint(0, ncam, (n_total,), generator=gen)
    px = torch.randint(0, 800, (n_total, 2), generator=gen).double() + 0.5
    d = torch.stack([(px[:, 0] - 400.0) / 1111.11, (px[:, 1] - 400.0) / 1111.11,
                     torch.ones(n_total, dtype=torch.float64)], -1)
    d = d / d.norm(dim=-1, keepdim=True)
    dirs = torch.einsum("nij,nj->ni", R[cam], d)
    dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).float()
    origins = pos[cam].float()
    return origins.to(device).contiguous(), dirs.to(device).contiguous()


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
N_CAMS = 8
CAM_KW = dict(fx=1111.11, width=800, height=800)


def bench_cameras(device, seed=5):
    """N_CAMS OpenCV-convention c2w matrices on the r=2.687 upper hemisphere, looking at 0."""
    gen = torch.Generator().manual_seed(seed)
    zc = 0.15 + 0.7 * torch.rand(N_CAMS, generator=gen)
    phi = torch.rand(N_CAMS, generator=gen) * 2 * math.pi
    rxy = torch.sqrt(1 - zc * zc)
    pos = torch.stack([rxy * torch.cos(phi), rxy * torch.sin(phi), zc], -1).double() * 2.687
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    c2w = torch.eye(4, dtype=torch.float64).repeat(N_CAMS, 1, 1)
    c2w[:, :3, :3] = torch.stack([right, down, fwd], dim=-1)
    c2w[:, :3, 3] = pos
    return [m.float().to(device).contiguous() for m in c2w]


def time_images(render, cams, reps=3):
    """ms per image (CUDA events), after one warm-up pass over the cameras."""
    with torch.no_grad():
        for c in cams:
            render(c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for c in cams:
                render(c)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(cams))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms in a side process.  nvidia-smi
    needs ~0.1 s before its first line, longer than a short timed region, so it is started before
    the warm-up; `mark()` brackets the timed region on the host clock and only samples whose
    receive time falls inside are reported (all samples under load as a fallback, flagged)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.lines = []          # (host receive time, line)
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Call right after the synchronising barrier at the start and at the end of the region."""
        if self.t0 is None:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 6:
                    continue
                try:
                    sm.append(float(p[0]))
                    mx.append(float(p[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons

        inside = [x for x in self.lines if self.t0 is not None and self.t1 is not None and
                  self.t0 <= x[0] <= self.t1 + 0.02]
        sm, mx, reasons = digest(inside)
        window = "timed region"
        if not sm:      # region shorter than the sampling period: everything sampled under load
            sm, mx, reasons = digest(self.lines)
            window = "warm-up + timed region"
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_fused(q, c):
    """fused fwd+bwd bytes = 84 Q + 8 S_skip + 128 S_vis + 3528 S_act (SH deg 2)."""
    return 84 * q + 8 * c[0] + 128 * c[1] + 3528 * c[2]


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import svox2_b200 as sv
    from svox2_b200 import csrc as C
    from svox2_b200.dist import ShardedTrainer
    from svox2_b200.slab import SlabTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    sc = build_scene(args.workload, dev)
    grid = sv.SparseGrid(reso=sc["reso"], basis_dim=9, device=dev)
    grid.links = sc["links"]
    grid.density_data = torch.nn.Parameter(sc["density"])
    grid.sh_data = torch.nn.Parameter(sc["sh"])
    grid.capacity = sc["n_rows"]
    grid.accelerate()
    for k, v in RENDER.items():
        setattr(grid.opt, k, v)

    # global ray pool: POOL batches of BATCH*world rays; rank r owns a contiguous slice of each
    gq = BATCH * world
    origins, dirs = build_rays(POOL * gq, dev)
    with torch.no_grad():
        gt = torch.cat([grid.volume_render(sv.Rays(origins[i:i + 50000], dirs[i:i + 50000]))
                        for i in range(0, POOL * gq, 50000)])
        ggen = torch.Generator(device=dev).manual_seed(99)
        gt = (gt + 0.02 * torch.randn(gt.shape, generator=ggen, device=dev)).clamp_(0, 1)
    slab = world > 1 and args.dist == "slab"
    # this rank's BATCH rays of every pooled global batch (kernel-only timing, counters)
    sl = [slice(b * gq + rank * BATCH, b * gq + (rank + 1) * BATCH) for b in range(POOL)]
    own = [(origins[s].contiguous(), dirs[s].contiguous(), gt[s].contiguous()) for s in sl]
    if slab:    # grid sharded by x-slabs: every rank takes the whole global batch
        gsl = [slice(b * gq, (b + 1) * gq) for b in range(POOL)]
        batches = [(origins[s].contiguous(), dirs[s].contiguous(), gt[s].contiguous()) for s in gsl]
        trainer = SlabTrainer(grid, world_size=world, rank=rank, **SYN)
    else:       # rays sharded, grid replicated
        batches = own
        trainer = ShardedTrainer(grid, world_size=world, rank=rank, **SYN)
    # e2e inputs in pinned host memory (one stacked buffer / F.mse_loss were measured SLOWER)
    host = [tuple(x.cpu().pin_memory() for x in b) for b in batches]
    np.random.seed(1234)

    def step(o, d, g):
        if slab:
            return trainer.step(sv.Rays(o, d), g)
        return trainer.step(sv.Rays(o, d), g, global_batch=gq)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # counts of the roofline model, one counting pass per pooled batch (untimed)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    for o, d, g in own:
        C.volume_render_cuvol(grid._to_cpp(), sv.Rays(o, d)._to_cpp(), grid.opt._to_cpp(),
                              _counters=counters)
    torch.cuda.synchronize()
    cnt = [c / POOL for c in counters.cpu().tolist()]

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(*batches[i % POOL])
    barrier()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(*batches[i % POOL])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # end to end: pinned host inputs -> H2D, step, D2H of the loss, every step.  The loss of step i
    # lands in a double-buffered pinned scalar and is read while step i+1 is already enqueued (the
    # driver in svox2_b200/train.py never blocks on the step it just launched either).
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    t0.record()
    loss_val = 0.0
    for i in range(args.steps):
        o, d, g = (x.to(dev, non_blocking=True) for x in host[i % POOL])
        rgb = step(o, d, g)
        loss_host[i & 1].copy_(torch.mean((rgb - g) ** 2), non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            loss_val = float(loss_host[(i - 1) & 1])
    loss_ev[(args.steps - 1) & 1].synchronize()
    loss_val = float(loss_host[(args.steps - 1) & 1])
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    if sampler:
        sampler.mark()
    clocks = sampler.stop() if sampler else None

    phase_ms = None
    if world > 1:   # untimed extra steps with per-phase CUDA events, one entry per rank
        barrier()
        trainer.profile = True
        for i in range(min(args.steps, 10)):
            step(*batches[i % POOL])
        mine = {k: round(v, 4) for k, v in trainer.phase_times_ms().items()}
        trainer.profile = False
        every = [None] * world
        dist.all_gather_object(every, mine)
        phase_ms = {k: [e.get(k) for e in every] for k in mine}     # per phase, one entry per rank
    if slab:
        trainer.gather_params()     # re-replicate before the single-GPU style measurements

    # dominant kernels alone: the fused fwd+bwd march pair
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    spec, ropt = grid._to_cpp(), grid.opt._to_cpp()
    holder = C.GridOutputGrads()
    holder.grad_density_out, holder.grad_sh_out = grid.density_data.grad, grid.sh_data.grad
    holder.mask_out = torch.zeros(sc["n_rows"], dtype=torch.bool, device=dev)
    rgb_out = torch.zeros_like(own[0][2])
    rspecs = [sv.Rays(o, d)._to_cpp() for o, d, _ in own]
    for i in range(3):
        C.volume_render_cuvol_fused(spec, rspecs[i % POOL], ropt, own[i % POOL][2], 0.0, 0.0,
                                    rgb_out, holder, _norm_rays=gq)
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.start()     # ncu --profile-from-start off captures from here
    k0.record()
    for i in range(args.steps):
        C.volume_render_cuvol_fused(spec, rspecs[i % POOL], ropt, own[i % POOL][2], 0.0, 0.0,
                                    rgb_out, holder, _norm_rays=gq)
    k1.record()
    torch.cuda.synchronize()
    if args.profile:
        np.random.seed(4321)
        for i in range(2):
            step(*batches[i % POOL])    # two full steps (TV + optimiser kernels) for the launch list
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    ms_fused = k0.elapsed_time(k1) / args.steps
    grid.density_data.grad.zero_()
    grid.sh_data.grad.zero_()

    # second half of the headline metric: whole-image render, 800x800 test cameras
    cams = bench_cameras(dev)
    ms_img = time_images(lambda c: grid.volume_render_image(sv.Camera(c, **CAM_KW)), cams)

    times = torch.tensor([ms, ms_e2e, ms_fused, ms_img], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_fused, ms_img = times.tolist()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    abytes = algorithmic_bytes_fused(BATCH, cnt)
    achieved = abytes / (ms_fused * 1e-3) / 1e9
    line = {
        "metric": "train rays/sec (fused fwd+bwd step) at 512^3 SH2",
        "value": gq * args.steps / (ms * 1e-3),
        "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": sc["name"], "reso": sc["reso"], "stored_voxels": sc["n_rows"],
            "occupancy": round(sc["occupancy"], 4), "rays_per_gpu": BATCH, "global_batch": gq,
            "step": "volume_render_fused + tv_grad_sparse(sigma,sh) + rmsprop(sigma,sh)",
            "l2": "inputs larger than L2 (grid+grads+rms >> 126 MB); 8 distinct ray batches cycled",
            "parallelism": (f"grid sharded into {world} x-slabs, every rank traverses the global "
                            "batch; per-ray scalar all-gathers + halo planes (NCCL)") if slab else
                           (f"ray-sharded x{world}, replicated grid" +
                            (", masked packed grad all-reduce (NCCL)" if world > 1 else "")),
            "samples_per_ray": {"skip": cnt[0] / BATCH, "visited": cnt[1] / BATCH,
                                "active": cnt[2] / BATCH},
        },
        "e2e": {"value": gq * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": 3 * batches[0][0].size(0) * 3 * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": loss_val,
                "loss_readback": "every step, one step late (double-buffered pinned scalar)"},
        "gpu_launches": trainer.kernels_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fused call: trace + shade + seed + scatter kernels",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes,
                     "ms_per_launch": ms_fused,
                     # dram__bytes_read+write of K1..K4 of one fused call, ncu --set full capture
                     # of this command (profiles/r1_ncu_step_kernels_full.txt): 335+252+45+419 MB
                     "traffic": 1.051e9 if args.workload == "c3" else None},
    }
    line["render"] = {"value": world * 0.64 / (ms_img * 1e-3), "unit": "MPix/s",
                      "ms_per_image": ms_img, "what": "volume_render_image, 800x800, fx 1111.11, "
                      f"{N_CAMS} hemisphere cameras per GPU (images sharded over GPUs)"}
    if slab:
        halo = sum(h.numel() for h in (trainer.halo_left, trainer.halo_right) if h is not None)
        line["exchange"] = {"mode": "slab", "slab_bounds": trainer.bounds,
                            "bytes_all_gathered_per_step": gq * 20 * world,
                            "halo_rows_rank0": halo, "phase_ms": phase_ms}
    elif world > 1:
        line["exchange"] = {"mode": "rays", "rows_per_step": trainer.exchanged_rows,
                            "bytes_all_reduced_per_step": trainer.exchanged_rows * 28 * 4,
                            "phase_ms": phase_ms}
    if args.cpu_baseline and world == 1:        # reported baselines: rank 0 at N=1 only
        line["cpu_baseline"] = cpu_port_baseline(sc, grid, batches[0], budget_rays=args.cpu_rays)
    if args.ref_cuda and world == 1:
        line["ref_cuda"] = reference_cuda_baseline(sc, grid, batches, args, gq)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# baselines
# ----------------------------------------------------------------------------------------------
class CpuPortStep:
    """One train step of the reference ALGORITHM on the host cores: the oracle's C port (pthreads
    over rays / cells / rows) of the fused fwd+bwd march, the sparse TV on a contiguous cell run
    and the masked RMSProp, on `sample` rays.  The TV cell count is scaled by sample/BATCH so
    that the per-ray cost is that of a whole step (opt.py:474-603)."""

    def __init__(self, sc, links, sample):
        import ctypes as CT
        import oracle as og
        self.og, self.CT, self.sample = og, CT, sample
        gsz = float(sc["reso"])
        self.links = np.ascontiguousarray(links, np.int32)
        self.g = og.Grid(self.links, sc["density"].cpu().numpy(), sc["sh"].cpu().numpy(),
                         np.full(3, 0.5 * gsz - 0.5, np.float32), np.full(3, 0.5 * gsz, np.float32))
        self.opt = og.make_opt(**RENDER)
        self.cores = os.cpu_count() or 1
        og.set_num_threads(self.cores)
        og.set_accum_float(True)
        n = self.g.density.shape[0]
        self.gd, self.gsh = np.zeros(n, np.float32), np.zeros(self.g.sh.shape, np.float32)
        self.drms, self.shrms = np.zeros_like(self.g.density), np.zeros_like(self.g.sh)
        self.mask = np.zeros(n, np.uint8)
        self.rng = np.random.default_rng(0)
        self.total_cells = int(np.prod(self.links.shape))
        self.n_tv = max(int(SYN["tv_sparsity"] * self.total_cells * sample / BATCH), 1)
        self.gc, self.cnt = self.g.c(), og._Counts()

    def __call__(self, o, d, gt):
        og, CT, L, g = self.og, self.CT, self.og.lib(), self.g
        X, Y, Z = self.links.shape
        rgb = np.zeros_like(o)
        L.og_volume_render_fused(CT.byref(self.gc), CT.byref(self.opt), og._ptr(o), og._ptr(d),
                                 CT.c_int64(o.shape[0]), CT.c_int64(o.shape[0]), og._ptr(gt),
                                 CT.c_float(0.0), CT.c_float(0.0), og._ptr(rgb), og._ptr(self.gd),
                                 og._ptr(self.gsh), None, og._ptr(self.mask), None, CT.byref(self.cnt))
        for data, ch, grad, lam in ((g.density, 1, self.gd, SYN["lambda_tv"]),
                                    (g.sh, g.sh.shape[1], self.gsh, SYN["lambda_tv_sh"])):
            start = int(self.rng.integers(0, self.total_cells))
            cells = np.ascontiguousarray((np.arange(start, start + self.n_tv, dtype=np.int64)
                                          % self.total_cells).astype(np.int32))
            L.og_tv_grad_sparse(og._ptr(self.links), X, Y, Z, og._ptr(data), ch, og._ptr(cells),
                                CT.c_int64(self.n_tv), og._ptr(self.mask), 0, ch, CT.c_float(lam), 0, 0,
                                og._ptr(grad))
        for data, rms, ch, grad, lr in ((g.density, self.drms, 1, self.gd, SYN["lr_sigma"]),
                                        (g.sh, self.shrms, g.sh.shape[1], self.gsh, SYN["lr_sh"])):
            L.og_rmsprop_step(og._ptr(data), og._ptr(rms), og._ptr(grad), CT.c_int64(self.mask.shape[0]),
                              ch, 1, og._ptr(self.mask), None, CT.c_int64(0), CT.c_float(SYN["rms_beta"]),
                              CT.c_float(lr), CT.c_float(1e-8), CT.c_float(-1e9), CT.c_float(lr))
        self.mask[:] = 0

    def close(self):
        self.og.set_num_threads(1)
        self.og.set_accum_float(False)

    def describe(self):
        return (f"fused fwd+bwd march + sparse TV on {self.n_tv} cells x2 + masked RMSProp x2, oracle C "
                f"port of the reference algorithm, {self.cores} pthreads")


def cpu_port_baseline(sc, grid, batch, budget_rays=1024, reps=3):
    """cpu_baseline of the N=1 line: CpuPortStep on the first `budget_rays` rays of batch 0."""
    step = CpuPortStep(sc, grid.links.cpu().numpy(), budget_rays)
    o, d, gt = (np.ascontiguousarray(x[:budget_rays].cpu().numpy()) for x in batch)
    step(o, d, gt)                                   # warm-up (page faults of the 1.4 GB arrays)
    t = time.perf_counter()
    for _ in range(reps):
        step(o, d, gt)
    dt = (time.perf_counter() - t) / reps
    step.close()
    return {"value": o.shape[0] / dt, "unit": "rays/s", "cores": step.cores, "kind": "port",
            "sample": f"first {o.shape[0]} rays of batch 0, one whole step: {step.describe()}, "
                      f"{dt:.2f} s per step"}


def reference_cuda_baseline(sc, grid, batches, args, gq):
    """The reference's own CUDA extension (built for sm_100, oracle/_ref) on the same workload,
    same step, same GPU: the kernel to beat."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "svox2")):
        return {"unavailable": "oracle/_ref not built"}
    sys.path.insert(0, ref_dir)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import svox2 as ref
    dev = grid.links.device
    rg = ref.SparseGrid(reso=sc["reso"], basis_dim=9, device=dev)
    rg.links = grid.links.clone()
    rg.density_data = torch.nn.Parameter(sc["density"].clone())
    rg.sh_data = torch.nn.Parameter(sc["sh"].clone())
    rg.capacity = sc["n_rows"]
    for k, v in RENDER.items():
        setattr(rg.opt, k, v)
    np.random.seed(1234)

    def rstep(o, d, g):
        rgb = rg.volume_render_fused(ref.Rays(o, d), g)
        rg.inplace_tv_grad(rg.density_data.grad, scaling=SYN["lambda_tv"], sparse_frac=SYN["tv_sparsity"])
        rg.inplace_tv_color_grad(rg.sh_data.grad, scaling=SYN["lambda_tv_sh"], sparse_frac=SYN["tv_sparsity"])
        rg.optim_density_step(SYN["lr_sigma"], beta=SYN["rms_beta"])
        rg.optim_sh_step(SYN["lr_sh"], beta=SYN["rms_beta"])
        return rgb

    for i in range(args.warmup):
        rstep(*batches[i % POOL])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        rstep(*batches[i % POOL])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    e0.record()
    for i in range(args.steps):
        rg.volume_render_fused(ref.Rays(batches[i % POOL][0], batches[i % POOL][1]), batches[i % POOL][2])
    e1.record()
    torch.cuda.synchronize()
    ms_f = e0.elapsed_time(e1) / args.steps
    ms_img = time_images(lambda c: rg.volume_render_image(ref.Camera(c, **CAM_KW)),
                         bench_cameras(dev))
    return {"value": BATCH / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
            "ms_fused_call": ms_f, "render_mpix_s": 0.64 / (ms_img * 1e-3), "ms_per_image": ms_img, "what": "sxyu/svox2 CUDA extension (TORCH_CUDA_ARCH_LIST=10.0), "
            "same grid/rays/step on this GPU, 1 GPU"}


def run_reference(args):
    """Reference arm: the reference algorithm on the host cores.  The reference's own
    pure-PyTorch path cannot run the 512^3 workload (its autograd materialises a dense [N,27]
    gradient per gather: ~50 TB of traffic per batch), so the timed number is the C port of the
    same algorithm (oracle, all cores) on a bounded sample of each batch; the reference's own
    Python path is timed beside it on its own CPU-runnable config (C1) and reported as an
    extra key."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    sc = build_scene(args.workload, dev)
    import oracle as og
    world = args.gpus
    gq = BATCH * world
    origins, dirs = build_rays(POOL * gq, "cpu")
    sample = args.cpu_rays
    step = CpuPortStep(sc, og.accel_dist_prop(sc["links"].cpu().numpy()), sample)
    cores, n_tv = step.cores, step.n_tv
    rng = np.random.default_rng(1)

    def one(i):
        s = (i % POOL) * gq
        o = np.ascontiguousarray(origins[s:s + sample].numpy())
        d = np.ascontiguousarray(dirs[s:s + sample].numpy())
        step(o, d, rng.uniform(0, 1, o.shape).astype(np.float32))

    for i in range(min(args.warmup, 1)):
        one(i)
    t = time.perf_counter()
    for i in range(args.steps):
        one(i)
    dt = time.perf_counter() - t
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "train rays/sec (fused fwd+bwd step) at 512^3 SH2",
        "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": sc["name"], "reso": sc["reso"], "stored_voxels": sc["n_rows"],
                   "rays_per_step": sample},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} rays of each pooled batch per step: fused fwd+bwd march + "
                                   f"sparse TV on {n_tv} cells x2 + masked RMSProp x2 "
                                   f"(oracle C port of the reference algorithm, {cores} pthreads)"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.ref_pytorch:
        line["reference_pytorch_c1"] = reference_pytorch_c1(args.ref_pytorch_rays)
    print(json.dumps(line), flush=True)


def reference_pytorch_c1(n_rays):
    """The reference's own pure-PyTorch lerp renderer + autograd (oracle/_ref, unmodified) on
    its CPU-runnable config C1: 128^3 dense, SH9, rays from the r=2.687 sphere."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "svox2")):
        return {"unavailable": "oracle/_ref not built"}
    sys.path.insert(0, ref_dir)
    import warnings
    warnings.filterwarnings("ignore")
    import svox2 as ref
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(1234)
    grid = ref.SparseGrid(reso=128, basis_dim=9, use_z_order=False, use_sphere_bound=False, device="cpu")
    with torch.no_grad():
        grid.sh_data[:] = torch.randn(grid.sh_data.shape, generator=gen) * 0.1
        grid.sh_data[:, ::9] = 0.5
        grid.density_data[:] = torch.rand(grid.density_data.shape, generator=gen) * 2
    for k, v in RENDER.items():
        setattr(grid.opt, k, v)
    v = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=gen), dim=-1)
    origins = v * 2.687
    targets = torch.rand(n_rays, 3, generator=gen) - 0.5
    dirs = torch.nn.functional.normalize(targets - origins, dim=-1)
    gt = torch.rand(n_rays, 3, generator=gen)
    t = time.perf_counter()
    rgb = grid.volume_render(ref.Rays(origins, dirs), use_kernel=False)
    t_f = time.perf_counter() - t
    torch.nn.functional.mse_loss(rgb, gt).backward()
    t_all = time.perf_counter() - t
    return {"config": "C1: SparseGrid 128^3 dense, basis_dim=9, volume_render(use_kernel=False) + autograd",
            "rays": n_rays, "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
            "fwd_s": t_f, "fwd_bwd_s": t_all, "fwd_rays_per_s": n_rays / t_f,
            "fwd_bwd_rays_per_s": n_rays / t_all}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--dist", default="slab", choices=["slab", "rays"],
                    help="multi-GPU decomposition: grid sharded by x-slabs (default) or rays "
                         "sharded over a replicated grid with a packed gradient all-reduce")
    ap.add_argument("--cpu-rays", type=int, default=2048, help="bounded CPU sample (rays)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-ref-cuda", dest="ref_cuda", action="store_false")
    ap.add_argument("--no-ref-pytorch", dest="ref_pytorch", action="store_false")
    ap.add_argument("--ref-pytorch-rays", type=int, default=256)
    ap.add_argument("--profile", action="store_true",
                    help="bracket the kernel-only loop + 2 full steps with cudaProfilerStart/Stop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
