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


def build_scene(workload, device, seed=20200823, zorder=False):
    """sigma = 30 inside a union of seeded ellipsoids, SH ~ N(0, 0.1^2) with DC 0.5, voxels kept =
    dilate^2(occupied); blob scale bisected so that the stored fraction hits the target.
    zorder: rows numbered along the Morton curve -- what SparseGrid.resample(use_z_order=True)
    produces (svox2.py:1399-1406) -- instead of the row-major cumsum opt.py's resample call yields."""
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
    n_rows = int(keep.sum().item())
    if zorder:
        from svox2_b200.utils import gen_morton
        morton = gen_morton(reso, device=device, dtype=torch.long).view(-1)   # cell -> curve position
        order = torch.argsort(morton)                                          # curve position -> cell
        kept_cells = order[keep.view(-1)[order]]                               # kept cells along the curve
        links = torch.full((reso ** 3,), -1, dtype=torch.int32, device=device)
        links[kept_cells] = torch.arange(n_rows, dtype=torch.int32, device=device)
        density = torch.zeros(n_rows, 1, device=device)
        density[links[inside.view(-1)].long(), 0] = 30.0
        del morton, order, kept_cells
    else:
        links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
        links[~keep.view(-1)] = -1
        density = (inside[keep].float() * 30.0).view(n_rows, 1).contiguous()
    links = links.view(reso, reso, reso).contiguous()
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


def time_images(render, cams, reps=3, sampler=None):
    """ms per image (CUDA events), after one warm-up pass over the cameras."""
    with torch.no_grad():
        for c in cams:
            render(c)
        torch.cuda.synchronize()
        if sampler:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for c in cams:
                render(c)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(cams))


_NVML_POLLER = r"""
import sys, time
import pynvml as N
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
period = float(sys.argv[2])
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
R = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
     "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
     "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
     "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
out = sys.stdout
while True:
    sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
    try:
        bits = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    except Exception:
        bits = 0
    names = "|".join(k for k, v in R.items() if bits & v)
    out.write("%.6f,%d,%d,%s\n" % (time.monotonic(), sm, mx, names))
    out.flush()
    time.sleep(period)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 2 ms by a side PROCESS (no GIL
    contention with the launching thread; `nvidia-smi -lms` cannot go below ~20 ms, longer than a
    20-step timed region).  Samples carry the poller's CLOCK_MONOTONIC stamp (system-wide on Linux);
    `mark()` brackets the timed region and only samples inside it are reported (all samples under
    load as a flagged fallback)."""

    def __init__(self, index, period_s=0.002):
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None
        try:
            phys = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_POLLER, str(phys), str(period_s)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        # second source, used only if the NVML poller yields nothing: nvidia-smi at its 20 ms floor
        self.smi, self.smi_lines = None, []
        try:
            self.smi = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except Exception:
            self.smi = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.smi.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 6:
                why = "|".join(n for n, v in zip(names, p[2:6]) if v.lower().startswith("active"))
                self.smi_lines.append(f"{time.monotonic():.6f},{p[0]},{p[1]},{why}")

    def mark(self):
        """Call right after the synchronising barrier at the start and at the end of the region."""
        if self.t0 is None:
            self.t0 = time.monotonic()
        else:
            self.t1 = time.monotonic()

    def stop(self):
        time.sleep(0.01)
        for pr in (self.proc, self.smi):
            if pr is None:
                continue
            pr.terminate()
            try:
                pr.wait(timeout=2)
            except Exception:
                pr.kill()
        how = "NVML (pynvml) polled every 2 ms by a side process"
        lines = self.lines
        if not lines:
            lines, how = self.smi_lines, "nvidia-smi -lms 20 (NVML poller produced nothing)"
        if not lines:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sampler available"]}
        rows = []
        for ln in lines:
            p = ln.split(",")
            if len(p) < 4:
                continue
            try:
                rows.append((float(p[0]), float(p[1]), float(p[2]), [x for x in p[3].split("|") if x]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and
                  self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if len(inside) < 3:     # a region shorter than a few periods: everything sampled under load
            inside, window = rows, "warm-up + timed region"
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(r[2] for r in inside)) if inside else None,
                "reasons": reasons, "samples": len(sm), "window": window,
                "how": how}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload):
    """DRAM bytes (read + write) of the dominant kernels from the committed `ncu --set full` capture
    of THIS workload, written by tools/ncu_traffic.py into profiles/r2_traffic.json together with
    the git head it was taken at.  None (and why) when there is no such file / entry."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed for this build (profiles/r2_traffic.json)"
    rec = json.load(open(path)).get(workload)
    if not rec:
        return None, f"profiles/r2_traffic.json has no entry for workload {workload}"
    return rec, f"profiles/r2_traffic.json ({rec.get('source')}, head {rec.get('head')})"


def _frac_bad(a, b, tol=1e-4):
    """Fraction of elements whose error exceeds tol * max|b| (tests/scenes.py::frac_bad on device)."""
    if b.numel() == 0:
        return 0.0
    denom = max(float(b.abs().max()), 1e-30)
    return float(((a - b).abs() > tol * denom).float().mean())


def multi_gpu_parity(sv, grid, trainer, slab, step, global_batches, local_batches, rank, world, dev):
    """Before anything is timed: the N-GPU step on global batch B must equal the 1-GPU step on the
    same B.  Every rank clones the starting grid, runs the single-GPU trainer on the
    whole global batch with the same TV seed, then the multi-GPU step, and compares colours, the
    sparse masks (bit for bit on the rows this rank is responsible for), the updated parameters
    (fraction of elements off by more than 1e-4 of the max) and the colours of a second step.
    Returns the per-job summary (worst rank); the caller fails the run when `ok` is false."""
    import torch.distributed as dist
    from svox2_b200.dist import ShardedTrainer
    ref = sv.SparseGrid(reso=[2, 2, 2], basis_dim=grid.basis_dim, device="cpu")
    ref.links = grid.links
    ref.density_data = torch.nn.Parameter(grid.density_data.data.clone())
    ref.sh_data = torch.nn.Parameter(grid.sh_data.data.clone())
    ref.capacity = grid.capacity
    for k in RENDER:
        setattr(ref.opt, k, getattr(grid.opt, k))
    single = ShardedTrainer(ref, world_size=1, rank=0, fused=False, **SYN)     # operator-by-operator
    out = {}
    for it in range(2):
        go, gd, gg = global_batches[it]
        np.random.seed(4242 + it)
        rgb1 = single.step(sv.Rays(go, gd), gg)
        np.random.seed(4242 + it)
        rgb = step(*local_batches[it])
        torch.cuda.synchronize()
        mine = slice(None) if slab else slice(rank * rgb.size(0), (rank + 1) * rgb.size(0))
        out["rgb_max_abs" if it == 0 else "rgb_max_abs_step2"] = float((rgb - rgb1[mine]).abs().max())
        if it == 0:
            if grid.sparse_grad_indexer is None and hasattr(trainer, "export_masks"):
                trainer.export_masks()      # slot-based step: masks live as slot flags until asked for
            rows = trainer.own_mask if slab else torch.ones_like(ref.sparse_grad_indexer)
            out["mask_equal"] = bool(
                torch.equal(grid.sparse_grad_indexer[rows], ref.sparse_grad_indexer[rows]) and
                torch.equal(grid.sparse_sh_grad_indexer[rows], ref.sparse_sh_grad_indexer[rows]))
            out["rows_checked"] = int(rows.sum())
            out["rows_touched"] = int(ref.sparse_sh_grad_indexer[rows].sum())
            out["fb_density"] = _frac_bad(grid.density_data.data[rows], ref.density_data.data[rows])
            out["fb_sh"] = _frac_bad(grid.sh_data.data[rows], ref.sh_data.data[rows])
    ok = (out["rgb_max_abs"] < 1e-5 and out["mask_equal"] and out["fb_density"] < 1e-3 and
          out["fb_sh"] < 1e-3 and out["rgb_max_abs_step2"] < 2e-3)
    vec = torch.tensor([out["rgb_max_abs"], out["rgb_max_abs_step2"], out["fb_density"], out["fb_sh"],
                        0.0 if out["mask_equal"] else 1.0, 0.0 if ok else 1.0],
                       dtype=torch.float64, device=dev)
    cnt = torch.tensor([out["rows_checked"], out["rows_touched"]], dtype=torch.int64, device=dev)
    dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    v = vec.tolist()
    del ref, single
    torch.cuda.empty_cache()
    return {"ok": v[5] == 0.0,
            "against": "1-GPU step (ShardedTrainer, world 1) on the same global batch, grid and TV seed",
            "rgb_max_abs": v[0], "mask_equal_on_owned_rows": v[4] == 0.0,
            "frac_bad_density_1e-4": v[2], "frac_bad_sh_1e-4": v[3], "rgb_max_abs_step2": v[1],
            "rows_checked": int(cnt[0]), "rows_touched": int(cnt[1]), "ranks": world,
            "bars": {"rgb_max_abs": 1e-5, "mask": "bit-exact", "frac_bad": 1e-3, "rgb_max_abs_step2": 2e-3},
            "reduction": "worst rank (max), row counts summed"}


def single_gpu_parity(sv, grid, batches):
    """N = 1: the fused step (one C call, compact gradient buffer) must equal the operator-by-operator
    sequence of opt.py (volume_render_fused, inplace_tv_grad, inplace_tv_color_grad, optim_density_step,
    optim_sh_step through SparseGrid) on the same grid, batch and TV cells: colours, both sparse masks
    bit for bit, RMS state and parameters.  Runs on two clones before anything is timed."""
    from svox2_b200.dist import ShardedTrainer

    def clone():
        g = sv.SparseGrid(reso=[2, 2, 2], basis_dim=grid.basis_dim, device="cpu")
        g.links = grid.links
        g.density_data = torch.nn.Parameter(grid.density_data.data.clone())
        g.sh_data = torch.nn.Parameter(grid.sh_data.data.clone())
        g.capacity = grid.capacity
        for k in RENDER:
            setattr(g.opt, k, getattr(grid.opt, k))
        return g

    a, b = clone(), clone()
    fused, plain = ShardedTrainer(a, fused=True, **SYN), ShardedTrainer(b, fused=False, **SYN)
    out = {}
    for it in range(2):
        o, d, g = batches[it]
        np.random.seed(777 + it)
        ra = fused.step(sv.Rays(o, d), g)
        np.random.seed(777 + it)
        rb = plain.step(sv.Rays(o, d), g)
        torch.cuda.synchronize()
        out["rgb_max_abs" if it == 0 else "rgb_max_abs_step2"] = float((ra - rb).abs().max())
        if it == 0:
            st = fused.export_masks()
            out["mask_equal"] = bool(torch.equal(a.sparse_grad_indexer, b.sparse_grad_indexer) and
                                     torch.equal(a.sparse_sh_grad_indexer, b.sparse_sh_grad_indexer))
            out["rows_touched"] = int(st[0])
            out["overflow"] = bool(st[1] or st[2])
            out["rms_rel_err"] = max(float((a.sh_rms - b.sh_rms).abs().max() / b.sh_rms.abs().max()),
                                     float((a.density_rms - b.density_rms).abs().max() / b.density_rms.abs().max()))
            out["fb_density"] = _frac_bad(a.density_data.data, b.density_data.data)
            out["fb_sh"] = _frac_bad(a.sh_data.data, b.sh_data.data)
    ok = (out["rgb_max_abs"] < 1e-5 and out["mask_equal"] and not out["overflow"] and out["rms_rel_err"] < 1e-4
          and out["fb_density"] < 1e-3 and out["fb_sh"] < 1e-3 and out["rgb_max_abs_step2"] < 2e-3)
    del a, b, fused, plain
    torch.cuda.empty_cache()
    return {"ok": bool(ok), "against": "operator-by-operator step through SparseGrid (dense .grad, bool masks) "
            "on the same batch, grid and TV seed", "rgb_max_abs": out["rgb_max_abs"],
            "masks_bit_exact": out["mask_equal"], "rms_rel_err": out["rms_rel_err"],
            "frac_bad_density_1e-4": out["fb_density"], "frac_bad_sh_1e-4": out["fb_sh"],
            "rgb_max_abs_step2": out["rgb_max_abs_step2"], "rows_touched": out["rows_touched"],
            "bars": {"rgb_max_abs": 1e-5, "masks": "bit-exact", "rms_rel_err": 1e-4, "frac_bad": 1e-3,
                     "rgb_max_abs_step2": 2e-3}}


def algorithmic_bytes_fused(q, c):
    """fused fwd+bwd bytes = 84 Q + 8 S_skip + 128 S_vis + 3528 S_act (SH deg 2)."""
    return 84 * q + 8 * c[0] + 128 * c[1] + 3528 * c[2]


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def ship_mode(trainer):
    """True when the slab trainer runs with ray-sharded traversal + record shipping (slab.py)."""
    peer = getattr(trainer, "_peer", None)
    return peer not in (None, False) and getattr(peer, "ship", None) is not None


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

    sc = build_scene(args.workload, dev, zorder=args.zorder)
    grid = sv.SparseGrid(reso=[2, 2, 2], basis_dim=9, device=dev)   # geometry follows links.shape
    grid.links = sc["links"]
    grid.density_data = torch.nn.Parameter(sc["density"])
    grid.sh_data = torch.nn.Parameter(sc["sh"])
    grid.capacity = sc["n_rows"]
    grid.accelerate()
    for k, v in RENDER.items():
        setattr(grid.opt, k, v)

    # global ray pool: POOL global batches; rank r owns a contiguous slice of each.  Weak scaling:
    # BATCH rays per GPU (global batch BATCH*world); strong: the reference's global batch of BATCH
    # rays (opt.py:79-81) split over the ranks.
    strong = args.scaling == "strong"
    per_rank = BATCH // world if strong else BATCH
    gq = per_rank * world
    origins, dirs = build_rays(POOL * gq, dev)
    with torch.no_grad():
        gt = torch.cat([grid.volume_render(sv.Rays(origins[i:i + 50000], dirs[i:i + 50000]))
                        for i in range(0, POOL * gq, 50000)])
        ggen = torch.Generator(device=dev).manual_seed(99)
        gt = (gt + 0.02 * torch.randn(gt.shape, generator=ggen, device=dev)).clamp_(0, 1)
    slab = world > 1 and args.dist == "slab"
    # this rank's BATCH rays of every pooled global batch (kernel-only timing, counters)
    sl = [slice(b * gq + rank * per_rank, b * gq + (rank + 1) * per_rank) for b in range(POOL)]
    own = [(origins[s].contiguous(), dirs[s].contiguous(), gt[s].contiguous()) for s in sl]
    gsl = [slice(b * gq, (b + 1) * gq) for b in range(POOL)]
    if slab:    # grid sharded by x-slabs: every rank takes the whole global batch
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

    parity = None
    if world > 1 and not args.no_parity:
        # N-GPU step == 1-GPU step on the same global batch, checked on every rank before timing
        gb = [(origins[s].contiguous(), dirs[s].contiguous(), gt[s].contiguous()) for s in gsl[:2]]
        parity = multi_gpu_parity(sv, grid, trainer, slab, step, gb, batches, rank, world, dev)
        np.random.seed(1234)

    if world == 1 and not args.no_parity and trainer._can_fuse():
        parity = single_gpu_parity(sv, grid, batches)
        np.random.seed(1234)

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(*batches[i % POOL])
    barrier()
    # the end-to-end loop below repeats the SAME K training steps as the device-resident
    # loop (state after the warm-up restored, same TV seeds), so that the two numbers differ by the
    # end-to-end overhead only.  (A grid that has trained K steps longer is a different workload: the
    # optimiser alone went from 0.23 to 0.28-0.33 ms per step over 400 steps in tools/timeline.py.)
    snapshot = {k: getattr(grid, k).detach().clone() for k in ("density_data", "sh_data", "density_rms", "sh_rms")
                if getattr(grid, k, None) is not None}    # (N > 1: every rank its own arrays, slab and halo rows alike)
    rng_state = np.random.get_state()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(*batches[i % POOL])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # rows the last step's optimisers updated (the masks survive the step): whole-step roofline
    if getattr(trainer, "_step_ws", None) is not None or getattr(trainer, "_mg", None) is not None:
        trainer.export_masks()          # slot-based steps keep the masks as slot flags until asked for
        trainer.check_overflow()
    touched = (int(grid.sparse_grad_indexer.sum()), int(grid.sparse_sh_grad_indexer.sum()))

    barrier()
    with torch.no_grad():
        for k, v in snapshot.items():
            getattr(grid, k).detach().copy_(v)
    np.random.set_state(rng_state)
    snapshot = None
    barrier()
    # end to end: pinned host inputs -> H2D, step, D2H of the loss, every step.  The loss of step i
    # lands in a double-buffered pinned scalar and is read while step i+1 is already enqueued (the
    # driver in svox2_b200/train.py never blocks on the step it just launched either).
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    # H2D on a copy stream into double-buffered device staging, one step ahead of the compute stream
    # (every step's inputs still cross PCIe inside the timed region); the loss comes from the step
    # itself when it is the fused call (svox2_train_step's mse_out), else from three torch kernels.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    stage = [tuple(torch.empty_like(x, device=dev) for x in host[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i & 1])
            for dst, src in zip(stage[i & 1], host[i % POOL]):
                dst.copy_(src, non_blocking=True)
            ready[i & 1].record(copy_stream)

    for ev in free:
        ev.record(main_stream)
    t0.record()
    loss_val = 0.0
    prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch(i + 1)
        main_stream.wait_event(ready[i & 1])
        o, d, g = stage[i & 1]
        rgb = step(o, d, g)
        fused_mse = getattr(trainer, "last_mse", None) if getattr(trainer, "_step_ws", None) is not None else None
        free[i & 1].record(main_stream)
        if fused_mse is not None:
            # the step's own MSE (alternating device scalars): read out on a side stream behind the
            # step, so the D2H copy does not sit between two steps on the compute stream
            with torch.cuda.stream(loss_stream):
                loss_stream.wait_event(free[i & 1])
                loss_host[i & 1].copy_(fused_mse, non_blocking=True)
                loss_ev[i & 1].record(loss_stream)
        else:
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
    holder.grad_density_out, holder.grad_sh_out, _ = grid._get_data_grads()     # dense .grad arrays
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
        for i in range(args.profile_steps):
            step(*batches[i % POOL])    # full steps (TV + optimiser kernels) for the launch list
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
        if parity is not None and not parity["ok"]:
            sys.exit(1)
        return
    peak, peak_src = peaks()
    abytes = algorithmic_bytes_fused(per_rank, cnt)
    achieved = abytes / (ms_fused * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(args.workload)
    # whole step: fused call + optimisers (24 B per updated element) + sparse TV
    # (16 B of link probes per cell of the run; 16 B data + 32 B read-modify-write per channel of
    # a non-empty cell, expected count = occupancy x cells)
    n_tv = max(int(SYN["tv_sparsity"] * sc["reso"] ** 3), 1)
    tv_bytes = 2 * n_tv * 16 + sc["occupancy"] * n_tv * 28 * 48
    opt_bytes = 24 * (touched[0] + 27 * touched[1])
    step_bytes = abytes + opt_bytes + tv_bytes
    line = {
        "metric": "train rays/sec (fused fwd+bwd step) at 512^3 SH2",
        "value": gq * args.steps / (ms * 1e-3),
        "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": sc["name"] + (" [rows in Morton order]" if args.zorder else ""), "reso": sc["reso"],
            "stored_voxels": sc["n_rows"],
            "occupancy": round(sc["occupancy"], 4), "rays_per_gpu": per_rank, "global_batch": gq,
            "step": "volume_render_fused + tv_grad_sparse(sigma,sh) + rmsprop(sigma,sh)" +
                    (" as one svox2_train_step call (compact gradient buffer)" if getattr(trainer, "_step_ws", None) is not None else "") +
                    (" in svox2_mg_step_* phases on the compact gradient buffer" if getattr(trainer, "_mg", None) is not None else ""),
            "l2": "inputs larger than L2 (grid+grads+rms >> 126 MB); 8 distinct ray batches cycled",
            "parallelism": (f"grid sharded into {world} x-slabs; " +
                            ("rays sharded for the traversal, sample records shipped to the owning slab with "
                             "NVLink stores into its queue (symmetric memory), per-ray partial colours and "
                             "halo planes exchanged with peer stores + device-side barriers"
                             if ship_mode(trainer) else
                             "every rank traverses the global batch; per-ray scalars and halo planes exchanged "
                             "with " + ("symmetric-memory peer stores over NVLink + device-side barriers"
                                        if getattr(trainer, "_peer", None) not in (None, False) else
                                        "NCCL all-gathers / P2P (no symmetric memory)"))) if slab else
                           (f"ray-sharded x{world}, replicated grid" +
                            (", masked packed grad all-reduce (NCCL)" if world > 1 else "")),
            "samples_per_ray": {"skip": cnt[0] / per_rank, "visited": cnt[1] / per_rank,
                                "active": cnt[2] / per_rank},
        },
        "e2e": {"value": gq * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": 3 * batches[0][0].size(0) * 3 * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": loss_val,
                "loss_readback": "every step, one step late (double-buffered pinned scalar, D2H on a side stream behind the step)",
                "h2d": "pinned host -> double-buffered device staging on a copy stream, one step ahead",
                "state": "parameters / RMS state / TV seeds of the start of the device-timed loop restored first: the same K steps"},
        "gpu_launches": trainer.kernels_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fused call: trace + shade + seed + scatter kernels",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes,
                     "ms_per_launch": ms_fused,
                     # dram__bytes_read+write of the kernels of one fused call (ncu --set full)
                     "traffic": traffic["fused_call_dram_bytes"] if traffic else None,
                     "traffic_source": traffic_src},
    }
    if world == 1:
        line["step_roofline"] = {
            "bound": "hbm", "what": "whole train step: fused call + sparse TV x2 + RMSProp x2",
            "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
            "algorithmic_bytes": {"fused": abytes, "optimisers": opt_bytes, "tv": tv_bytes},
            "rows_updated": {"density": touched[0], "sh": touched[1]},
            "traffic": traffic.get("step_dram_bytes") if traffic else None}
    if parity is not None:
        line["parity"] = parity
    line["render"] = {"value": world * 0.64 / (ms_img * 1e-3), "unit": "MPix/s",
                      "ms_per_image": ms_img, "what": "volume_render_image, 800x800, fx 1111.11, "
                      f"{N_CAMS} hemisphere cameras per GPU (images sharded over GPUs)"}
    if slab:
        halo = sum(h.numel() for h in (trainer.halo_left, trainer.halo_right) if h is not None)
        line["exchange"] = {"mode": "slab-ship" if ship_mode(trainer) else "slab", "slab_bounds": trainer.bounds,
                            "bytes_all_gathered_per_step": gq * (16 if ship_mode(trainer) else 20) * world,
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
    if parity is not None and not parity["ok"]:
        sys.exit("bench.py: multi-GPU parity check FAILED (see the `parity` object of the line above)")


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


def cpu_port_baseline(sc, grid, batch, budget_rays=BATCH, reps=2):
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


# ----------------------------------------------------------------------------------------------
# the other two GPU configurations of BASELINE.json: C4 (NDC full-image render), C5 (MSI train step)
# ----------------------------------------------------------------------------------------------
def _load_ref():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "svox2")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import svox2 as ref
    return ref


def build_c4(device, seed=11):
    """LLFF NDC grid 1408x1156x128 (llff.json): content slabs at three NDC depths, ~4 % stored."""
    reso = [1408, 1156, 128]
    gen = torch.Generator(device=device).manual_seed(seed)
    keep = torch.zeros(reso, dtype=torch.bool, device=device)
    for z0 in (20, 52, 90):
        blob = torch.rand(reso[0] // 16 + 1, reso[1] // 16 + 1, generator=gen, device=device) < 0.45
        keep[:, :, z0:z0 + 4] |= blob.repeat_interleave(16, 0).repeat_interleave(16, 1)[:reso[0], :reso[1], None]
    n_rows = int(keep.sum())
    links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
    links[~keep.view(-1)] = -1
    links = links.view(*reso).contiguous()
    density = torch.rand(n_rows, 1, generator=gen, device=device) * 40.0
    sh = torch.randn(n_rows, 27, generator=gen, device=device) * 0.1
    sh[:, ::9] += 0.5
    return dict(links=links, density=density, sh=sh, n_rows=n_rows, reso=reso,
                radius=[1.496031746031746, 1.6613756613756614, 1.0],
                name="C4: LLFF NDC grid 1408x1156x128 SH deg 2, full-image render at 1008x756")


def c4_grid(mod, sc, device):
    g = mod.SparseGrid(reso=[8, 8, 8], radius=sc["radius"], basis_dim=9, use_sphere_bound=False,
                       device=device)
    g.links = sc["links"].clone()
    g.density_data = torch.nn.Parameter(sc["density"].clone())
    g.sh_data = torch.nn.Parameter(sc["sh"].clone())
    g.capacity = sc["n_rows"]
    g.accelerate()
    g.opt.step_size, g.opt.sigma_thresh, g.opt.stop_thresh = 0.5, 1e-8, 1e-7
    g.opt.background_brightness, g.opt.last_sample_opaque = 0.5, False
    return g


C4_CAM = dict(fx=820.0, width=1008, height=756)


def c4_cameras(device, n=8):
    cams = []
    for i in range(n):      # a small forward-facing rig around the origin (LLFF poses after recentring)
        c2w = torch.eye(4)
        a = 2 * math.pi * i / n
        c2w[:3, 3] = torch.tensor([0.08 * math.cos(a), 0.06 * math.sin(a), 0.0])
        cams.append(c2w.to(device))
    return cams


def run_c4(args):
    import svox2_b200 as sv
    assert torch.cuda.is_available()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sc = build_c4(dev)
    grid = c4_grid(sv, sc, dev)
    W, H, fx = C4_CAM["width"], C4_CAM["height"], C4_CAM["fx"]
    ndc = (2 * fx / W, 2 * fx / H)
    cams = c4_cameras(dev)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_img = time_images(lambda c: grid.volume_render_image(sv.Camera(c, ndc_coeffs=ndc, **C4_CAM)), cams,
                         reps=max(3, args.steps // len(cams)), sampler=sampler)
    # end to end through the public call: camera pose from pinned host memory, image back to the host
    host_c2w = [c.cpu().pin_memory() for c in cams]
    out_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    n_e2e = max(len(cams), args.steps // 4)
    for i in range(n_e2e):
        c = host_c2w[i % len(cams)].to(dev, non_blocking=True)
        out_host.copy_(grid.volume_render_image(sv.Camera(c, ndc_coeffs=ndc, **C4_CAM)), non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1) / n_e2e
    if sampler:
        sampler.mark()
    clocks = sampler.stop() if sampler else None
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    from svox2_b200 import csrc as C
    rays = sv.Camera(cams[0], ndc_coeffs=ndc, **C4_CAM).gen_rays()
    C.volume_render_cuvol(grid._to_cpp(), rays._to_cpp(), grid.opt._to_cpp(), _counters=counters)
    cnt = counters.cpu().tolist()
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms_img, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_img, ms_e2e = tt.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    mpix = W * H / 1e6
    peak, peak_src = peaks()
    abytes = 36 * W * H + 4 * cnt[0] + 64 * cnt[1] + 864 * cnt[2]          # SURVEY 8d forward bytes
    line = {"metric": "render MPix/s (volume_render_image, LLFF NDC grid)", "value": world * mpix / (ms_img * 1e-3),
            "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_img, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": sc["name"], "reso": sc["reso"], "stored_voxels": sc["n_rows"],
                       "image": [W, H], "fx": fx, "ndc_coeffs": ndc, "step": "one 1008x756 image",
                       "l2": "links alone are 833 MB >> L2; 8 cameras cycled",
                       "parallelism": f"images sharded over {world} GPU(s), replicated grid",
                       "samples_per_ray": {"skip": cnt[0] / (W * H), "visited": cnt[1] / (W * H),
                                           "active": cnt[2] / (W * H)}},
            "e2e": {"value": world * mpix / (ms_e2e * 1e-3), "unit": "MPix/s", "h2d_bytes_per_step": 64,
                    "d2h_bytes_per_step": W * H * 12, "ms_per_step": ms_e2e},
            "gpu_launches": args.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "lattice_march_kernel<FWD, FROM_CAM>",
                         "achieved": abytes / (ms_img * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": abytes / (ms_img * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "ms_per_launch": ms_img, "traffic": None,
                         "note": "issue-bound, not bandwidth-bound (profiles/r1_ncu_image_render.txt)"}}
    if args.ref_cuda and world == 1:
        ref = _load_ref()
        if ref is None:
            line["ref_cuda"] = {"unavailable": "oracle/_ref not built"}
        else:
            rg = c4_grid(ref, sc, dev)
            ms_ref = time_images(lambda c: rg.volume_render_image(ref.Camera(c, ndc_coeffs=ndc, **C4_CAM)), cams)
            line["ref_cuda"] = {"value": mpix / (ms_ref * 1e-3), "unit": "MPix/s", "ms_per_image": ms_ref,
                                "what": "sxyu/svox2 CUDA extension, same grid / cameras / GPU"}
    print(json.dumps(line), flush=True)


def build_c5(device, reso=640, bg_reso=1024, bg_layers=64, seed=17):
    """tnt.json-sized scene: sparsified 640^3 foreground (~6 % stored), MSI 2048x1024x64 with sparse
    shells of density (the state after the first upsamples of a T&T run)."""
    gen0 = torch.Generator().manual_seed(99)
    nb = 40
    centers = (torch.rand(nb, 3, generator=gen0) - 0.5) * 1.3
    radii = torch.rand(nb, 3, generator=gen0) * 0.17 + 0.08
    inside = _inside_blobs(reso, centers, radii, 0.85, device)
    keep = _dilate(inside, 2)
    n_rows = int(keep.sum())
    links = torch.cumsum(keep.view(-1).to(torch.int32), 0, dtype=torch.int32) - 1
    links[~keep.view(-1)] = -1
    links = links.view(reso, reso, reso).contiguous()
    gen = torch.Generator(device=device).manual_seed(seed)
    density = (inside[keep].float() * 25.0).view(n_rows, 1).contiguous()
    del inside, keep
    sh = torch.randn(n_rows, 27, generator=gen, device=device) * 0.1
    sh[:, ::9] += 0.5
    nbg = 2 * bg_reso * bg_reso
    bg = torch.zeros(nbg, bg_layers, 4, device=device)
    bg[..., :3] = torch.randn(nbg, bg_layers, 3, generator=gen, device=device) * 0.4
    lay = torch.rand(nbg, bg_layers, generator=gen, device=device)
    tex = torch.rand(nbg, 1, generator=gen, device=device) < 0.2
    bg[..., 3] = torch.where((lay < 0.15) & tex, lay * 40.0, torch.zeros_like(lay))
    return dict(links=links, density=density, sh=sh, bg=bg, n_rows=n_rows, reso=reso, bg_reso=bg_reso,
                bg_layers=bg_layers,
                name=f"C5: T&T {reso}^3 foreground + {bg_layers}-layer MSI {2 * bg_reso}x{bg_reso} (tnt.json)")


def c5_grid(mod, sc, device):
    g = mod.SparseGrid(reso=[8, 8, 8], basis_dim=9, use_sphere_bound=False, device=device,
                       background_nlayers=sc["bg_layers"], background_reso=sc["bg_reso"])
    g.links = sc["links"].clone()
    g.density_data = torch.nn.Parameter(sc["density"].clone())
    g.sh_data = torch.nn.Parameter(sc["sh"].clone())
    g.capacity = sc["n_rows"]
    g.background_data = torch.nn.Parameter(sc["bg"].clone())
    g.accelerate()
    g.opt.step_size, g.opt.sigma_thresh, g.opt.stop_thresh = 0.5, 1e-8, 1e-7
    g.opt.background_brightness = 1.0
    return g


def c5_rays(n, seed, device):
    """Cameras INSIDE the grid at radius ~0.85 (the NSVF loader scales T&T cameras by 0.95,
    nsvf_dataset.py:46,167-170) looking at the centre region; most rays leave the foreground box and
    composite the MSI shells."""
    gen = torch.Generator().manual_seed(seed)
    v = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    origins = v * 0.85
    target = (torch.rand(n, 3, generator=gen) - 0.5) * 1.2
    dirs = torch.nn.functional.normalize(target - origins, dim=-1)
    return origins.to(device).contiguous(), dirs.to(device).contiguous()


TNT = dict(lr_sigma=30.0, lr_sh=1e-2, lambda_tv=5e-5, lambda_tv_sh=5e-3, tv_sparsity=0.01, rms_beta=0.95,
           beta_loss=1e-5, sparsity_loss=1e-11, lr_sigma_bg=3.0, lr_color_bg=0.1,
           lambda_tv_background_sigma=1e-3, lambda_tv_background_color=1e-3, tv_background_sparsity=0.01)


def run_c5(args):
    """tnt.json train step: fused render with MSI + beta/sparsity losses, sparse TV (sigma, SH, MSI),
    RMSProp (sigma, SH, MSI).  N > 1: rays sharded over a replicated grid (ShardedTrainer)."""
    import svox2_b200 as sv
    from svox2_b200.dist import ShardedTrainer
    assert torch.cuda.is_available()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sc = build_c5(dev)
    grid = c5_grid(sv, sc, dev)
    gq = BATCH * world
    o, d = c5_rays(POOL * gq, 5, dev)
    with torch.no_grad():
        gt = torch.cat([grid.volume_render(sv.Rays(o[i:i + 50000], d[i:i + 50000]))
                        for i in range(0, POOL * gq, 50000)])
        gt = (gt + 0.02 * torch.randn(gt.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(99))).clamp_(0, 1)
    sl = [slice(b * gq + rank * BATCH, b * gq + (rank + 1) * BATCH) for b in range(POOL)]
    batches = [(o[s].contiguous(), d[s].contiguous(), gt[s].contiguous()) for s in sl]
    trainer = ShardedTrainer(grid, world_size=world, rank=rank, **TNT)
    np.random.seed(1234)
    torch.manual_seed(1234)
    step = lambda bo, bd, bg: trainer.step(sv.Rays(bo, bd), bg, global_batch=gq)
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(*batches[i % POOL])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(*batches[i % POOL])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    n_bg = int(grid.sparse_background_indexer.sum())
    host = [tuple(x.cpu().pin_memory() for x in b) for b in batches]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        bo, bd, bg = (x.to(dev, non_blocking=True) for x in host[i % POOL])
        rgb = step(bo, bd, bg)
        loss_host.copy_(torch.mean((rgb - bg) ** 2), non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "train rays/sec (fused fwd+bwd step with MSI background), tnt.json", "value": gq / (ms * 1e-3),
            "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": sc["name"], "reso": sc["reso"], "stored_voxels": sc["n_rows"],
                       "msi_texels": int(sc["bg"].size(0)), "rays_per_gpu": BATCH, "global_batch": gq,
                       "step": "volume_render_fused(beta, sparsity) + tv_grad_sparse(sigma, sh) + msi_tv_grad_sparse "
                               "+ rmsprop(sigma, sh, background)",
                       "msi_shell_rows_touched_last_step": n_bg,
                       "l2": "inputs larger than L2; 8 distinct ray batches cycled",
                       "parallelism": f"ray-sharded x{world}, replicated grid" +
                                      (", masked packed grad all-reduce (NCCL)" if world > 1 else "")},
            "e2e": {"value": gq / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 3 * BATCH * 12,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "last_loss": float(loss_host)},
            "gpu_launches": 12 * args.steps, "clocks": clocks,
            "roofline": None}
    if args.ref_cuda and world == 1:
        ref = _load_ref()
        if ref is None:
            line["ref_cuda"] = {"unavailable": "oracle/_ref not built"}
        else:
            del trainer
            rg = c5_grid(ref, sc, dev)
            np.random.seed(1234)

            def rstep(bo, bd, bg):
                rgb = rg.volume_render_fused(ref.Rays(bo, bd), bg, beta_loss=TNT["beta_loss"],
                                             sparsity_loss=TNT["sparsity_loss"])
                rg.inplace_tv_grad(rg.density_data.grad, scaling=TNT["lambda_tv"], sparse_frac=0.01)
                rg.inplace_tv_color_grad(rg.sh_data.grad, scaling=TNT["lambda_tv_sh"], sparse_frac=0.01)
                rg.inplace_tv_background_grad(rg.background_data.grad, scaling=TNT["lambda_tv_background_color"],
                                              scaling_density=TNT["lambda_tv_background_sigma"],
                                              sparse_frac=0.01, contiguous=True)
                rg.optim_density_step(30.0, beta=0.95)
                rg.optim_sh_step(1e-2, beta=0.95)
                rg.optim_background_step(3.0, 0.1, beta=0.95)
                return rgb

            for i in range(min(args.warmup, 5)):
                rstep(*batches[i % POOL])
            torch.cuda.synchronize()
            n_ref = max(5, args.steps // 4)
            e0.record()
            for i in range(n_ref):
                rstep(*batches[i % POOL])
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / n_ref
            line["ref_cuda"] = {"value": BATCH / (ms_ref * 1e-3), "unit": "rays/s", "ms_per_step": ms_ref,
                                "what": "sxyu/svox2 CUDA extension, same grid / rays / step / GPU"}
    print(json.dumps(line), flush=True)


def run_reference(args):
    """Reference arm: the reference algorithm on the host cores.  The reference's own
    pure-PyTorch path cannot run the 512^3 workload (its autograd materialises a dense [N,27]
    gradient per gather: ~50 TB of traffic per batch), so the timed number is the C port of the
    same algorithm (oracle, all cores) on the SAME config as our arm: whole 5000-ray batches of the
    same pool, one whole step each (march + TV on the full 1 % cell run + RMSProp).  Only if K such
    steps would not finish within a few minutes on this host is the per-step sample shrunk (and
    said so in `cpu_baseline.sample`).  The reference's own Python path is timed beside it on its
    own CPU-runnable config (C1) and reported as an extra key."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    sc = build_scene(args.workload, dev)
    import oracle as og
    world = args.gpus
    per_rank = BATCH // world if args.scaling == "strong" else BATCH
    gq = per_rank * world
    origins, dirs = build_rays(POOL * gq, "cpu")
    links = og.accel_dist_prop(sc["links"].cpu().numpy())
    sample = min(args.cpu_rays, gq)
    budget_s = 150.0
    rng = np.random.default_rng(1)

    def make(sample):
        step = CpuPortStep(sc, links, sample)

        def one(i):
            s = (i % POOL) * gq
            o = np.ascontiguousarray(origins[s:s + sample].numpy())
            d = np.ascontiguousarray(dirs[s:s + sample].numpy())
            step(o, d, rng.uniform(0, 1, o.shape).astype(np.float32))
        return step, one

    step, one = make(sample)
    one(0)                                          # page-faults the 1.4 GB arrays
    t = time.perf_counter()
    one(1)
    t_step = time.perf_counter() - t
    shrunk = False
    if t_step * (args.steps + max(args.warmup - 2, 0)) > budget_s:
        step.close()
        sample = max(256, int(sample * budget_s / (t_step * (args.steps + max(args.warmup - 2, 0)))))
        step, one = make(sample)
        one(0)
        shrunk = True
    cores, n_tv = step.cores, step.n_tv
    for i in range(max(args.warmup - 2, 0)):
        one(i)
    t = time.perf_counter()
    for i in range(args.steps):
        one(i)
    dt = time.perf_counter() - t
    step.close()
    val = sample * args.steps / dt
    what = (f"{sample} rays per step" + (" (the whole batch)" if sample == gq else
            f" of each {gq}-ray batch" + (" (shrunk to keep the run within minutes)" if shrunk else "")))
    line = {
        "impl": "reference", "metric": "train rays/sec (fused fwd+bwd step) at 512^3 SH2",
        "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": sc["name"], "reso": sc["reso"], "stored_voxels": sc["n_rows"],
                   "occupancy": round(sc["occupancy"], 4), "rays_per_gpu": per_rank, "global_batch": gq,
                   "step": "volume_render_fused + tv_grad_sparse(sigma,sh) + rmsprop(sigma,sh)",
                   "rays_per_step": sample},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{what}: fused fwd+bwd march + sparse TV on {n_tv} cells x2 + "
                                   f"masked RMSProp x2 (oracle C port of the reference algorithm, "
                                   f"{cores} pthreads)"},
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
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS) + ["c4", "c5"],
                    help="c3 (default, the config the metric is quoted on), c2 / small (same step on smaller "
                         "grids), c4 (LLFF NDC full-image render), c5 (tnt.json MSI train step)")
    ap.add_argument("--dist", default="slab", choices=["slab", "rays"],
                    help="multi-GPU decomposition: grid sharded by x-slabs (default) or rays "
                         "sharded over a replicated grid with a packed gradient all-reduce")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 5000 rays per GPU per step; strong: the reference's global batch of "
                         "5000 rays (opt.py:79-81) whatever the GPU count")
    ap.add_argument("--zorder", action="store_true",
                    help="number the voxel rows along the Morton curve (resample(use_z_order=True)) instead of "
                         "row-major; a layout experiment, not the headline configuration")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true",
                    help="skip the N-GPU == 1-GPU parity block that precedes the timed loop when N > 1")
    ap.add_argument("--cpu-rays", type=int, default=BATCH,
                    help="CPU arm / cpu_baseline: rays per step (default: the full 5000-ray batch)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-ref-cuda", dest="ref_cuda", action="store_false")
    ap.add_argument("--no-ref-pytorch", dest="ref_pytorch", action="store_false")
    ap.add_argument("--ref-pytorch-rays", type=int, default=256)
    ap.add_argument("--profile-steps", type=int, default=2,
                    help="whole steps inside the --profile bracket (1 for the ncu --set full capture)")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the kernel-only loop + 2 full steps with cudaProfilerStart/Stop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
