"""Training driver of the hot path: the loop of the reference's opt/opt.py (opt.py:369-659) around
``SparseGrid`` -- exponential LR schedules with the reverse-cosine delay (util.py:78-111), a
device-resident shuffled ray pool (util.py:178-190 keeps it on the host), fused render + sparse TV
+ masked RMSProp per batch, the resolution schedule with weight-threshold resampling, PSNR eval.

Differences by design: the per-step MSE is accumulated on the device and read every
``print_every`` steps instead of ``.item()`` per step (opt.py:505-508), and the step itself is a
``ShardedTrainer`` / ``SlabTrainer`` call, so the same driver runs on 1..8 GPUs (one process per
GPU, ``torchrun``).  Data: ``load_nerf_synthetic`` reads a NeRF-synthetic (Blender) scene
directory; without one, ``synthetic_dataset`` renders posed images of a hidden analytic scene with
this package's own renderer (no dataset ships with the image).

    python -m svox2_b200.train --reso "[[64,64,64],[128,128,128]]" --n_iters 600
    python -m svox2_b200.train --config opt/configs/syn.json --data_dir <nerf_synthetic/lego>
"""
import argparse
import dataclasses
import json
import math
import time
from typing import List, Optional

import numpy as np
import torch

from .grid import Camera, Rays, SparseGrid


def expon_lr(lr_init: float, lr_final: float, delay_steps: int = 0, delay_mult: float = 1.0,
             max_steps: int = 1000000):
    """step -> learning rate: log-linear from lr_init to lr_final over max_steps, scaled by a
    reverse-cosine ease-in over the first delay_steps (behaviour of util.py:78-111)."""
    def at(step: int) -> float:
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        ease = 1.0
        if delay_steps > 0:
            ease = delay_mult + (1.0 - delay_mult) * math.sin(0.5 * math.pi * min(max(step / delay_steps, 0.0), 1.0))
        t = min(max(step / max_steps, 0.0), 1.0)
        return ease * math.exp(math.log(lr_init) * (1.0 - t) + math.log(lr_final) * t)
    return at


@dataclasses.dataclass
class TrainConfig:
    """opt.py's argument defaults + opt/configs/syn.json."""
    reso: str = "[[256, 256, 256], [512, 512, 512]]"
    upsamp_every: int = 3 * 12800
    init_iters: int = 0
    n_iters: int = 10 * 12800
    batch_size: int = 5000
    basis_dim: int = 9
    init_sigma: float = 0.1
    lr_sigma: float = 3e1
    lr_sigma_final: float = 5e-2
    lr_sigma_decay_steps: int = 250000
    lr_sigma_delay_steps: int = 15000
    lr_sigma_delay_mult: float = 1e-2
    lr_sh: float = 1e-2
    lr_sh_final: float = 5e-6
    lr_sh_decay_steps: int = 250000
    lr_sh_delay_steps: int = 0
    lr_sh_delay_mult: float = 1e-2
    rms_beta: float = 0.95
    lambda_tv: float = 1e-5
    lambda_tv_sh: float = 1e-3
    tv_sparsity: float = 0.01
    tv_early_only: int = 1
    tv_decay: float = 1.0
    lambda_beta: float = 0.0
    lambda_sparsity: float = 0.0
    thresh_type: str = "weight"
    weight_thresh: float = 0.0005 * 512
    density_thresh: float = 5.0
    max_grid_elements: int = 44_000_000
    upsample_density_add: float = 0.0
    step_size: float = 0.5
    sigma_thresh: float = 1e-8
    stop_thresh: float = 1e-7
    background_brightness: float = 1.0
    last_sample_opaque: bool = False
    near_clip: float = 0.0
    # MSI background (opt.py:108-124, opt/configs/tnt.json)
    background_nlayers: int = 0
    background_reso: int = 512
    init_sigma_bg: float = 0.1
    lr_sigma_bg: float = 3e0
    lr_color_bg: float = 1e-1
    lambda_tv_background_sigma: float = 1e-2
    lambda_tv_background_color: float = 1e-2
    tv_background_sparsity: float = 0.01
    background_density_thresh: float = 1.0 + 1e-9
    print_every: int = 20
    seed: int = 20200823

    @classmethod
    def from_json(cls, path: str, **overrides):
        with open(path) as f:
            d = json.load(f)
        d.update(overrides)
        return cls(**{k: v for k, v in d.items() if k in {f.name for f in dataclasses.fields(cls)}})

    def reso_list(self) -> List[List[int]]:
        r = json.loads(self.reso) if isinstance(self.reso, str) else self.reso
        return [[x, x, x] if isinstance(x, int) else list(x) for x in r]


@dataclasses.dataclass
class PosedImages:
    """A dataset as the driver needs it: cameras + [n,H,W,3] images on the device."""
    c2w: torch.Tensor          # [n,4,4]
    images: torch.Tensor       # [n,H,W,3]
    fx: float
    width: int
    height: int

    def camera(self, i: int) -> Camera:
        return Camera(self.c2w[i], fx=self.fx, width=self.width, height=self.height)


class RayPool:
    """All training rays and their colours resident in HBM; one device-side permutation per epoch
    (the reference shuffles on the host and uploads each batch, util.py:178-190 / opt.py:474-478)."""

    def __init__(self, data, seed: int = 0):
        if hasattr(data, "gen_rays"):    # datasets.PosedImageSet: float32 rays built on the device
            rays, rgb = data.gen_rays()  # (world or NDC), as the reference's loaders make them
            self.origins, self.dirs, self.rgb = rays.origins, rays.dirs, rgb
        else:
            origins, dirs = [], []
            for i in range(data.c2w.size(0)):
                r = data.camera(i).gen_rays()
                origins.append(r.origins)
                dirs.append(r.dirs)
            self.origins = torch.cat(origins).contiguous()
            self.dirs = torch.cat(dirs).contiguous()
            self.rgb = data.images.reshape(-1, 3).contiguous()
        self.gen = torch.Generator(device=self.rgb.device).manual_seed(seed)
        self.shuffle()

    def __len__(self):
        return self.rgb.size(0)

    def shuffle(self):
        perm = torch.randperm(len(self), device=self.rgb.device, generator=self.gen)
        self.origins, self.dirs, self.rgb = self.origins[perm], self.dirs[perm], self.rgb[perm]

    def batch(self, begin: int, end: int):
        return Rays(self.origins[begin:end], self.dirs[begin:end]), self.rgb[begin:end]


def hemisphere_cameras(n: int, radius: float, device, seed: int = 0) -> torch.Tensor:
    """n OpenCV-convention c2w matrices on the upper hemisphere, looking at the origin."""
    gen = torch.Generator().manual_seed(seed)
    z = 0.15 + 0.75 * torch.rand(n, generator=gen)
    phi = torch.rand(n, generator=gen) * 2 * math.pi
    rxy = torch.sqrt(1 - z * z)
    pos = torch.stack([rxy * torch.cos(phi), rxy * torch.sin(phi), z], -1).double() * radius
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    c2w = torch.eye(4, dtype=torch.float64).repeat(n, 1, 1)
    c2w[:, :3, :3] = torch.stack([right, down, fwd], dim=-1)
    c2w[:, :3, 3] = pos
    return c2w.float().to(device)


def synthetic_dataset(device, n_images: int = 24, width: int = 100, height: int = 100,
                      reso: int = 64, seed: int = 3) -> PosedImages:
    """Posed images of a hidden scene (coloured ellipsoids in a dense grid), rendered with this
    package's renderer -- stands in for NeRF-synthetic, which is not in the image."""
    gen = torch.Generator().manual_seed(seed)
    g = SparseGrid(reso=reso, basis_dim=1, device=device)
    ax = (torch.arange(reso, dtype=torch.float32) + 0.5) / reso * 2 - 1
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    sigma = torch.zeros(reso, reso, reso)
    col = torch.zeros(reso, reso, reso, 3)
    for _ in range(6):
        c = (torch.rand(3, generator=gen) - 0.5) * 1.0
        r = torch.rand(3, generator=gen) * 0.2 + 0.15
        inside = ((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2 < 1
        sigma[inside] = 60.0
        col[inside] = torch.rand(3, generator=gen)
    with torch.no_grad():
        g.density_data.data[:] = sigma.reshape(-1, 1).to(device)
        # colour = 0.5 + C0 * sh  (SH DC basis 0.2820948)
        g.sh_data.data[:] = ((col.reshape(-1, 3) - 0.5) / 0.28209479177387814).to(device)
    g.opt.background_brightness = 1.0
    g.opt.sigma_thresh, g.opt.stop_thresh = 1e-8, 1e-7
    c2w = hemisphere_cameras(n_images, 2.687, device, seed)
    fx = 1111.11 * width / 800.0
    with torch.no_grad():
        imgs = torch.stack([g.volume_render_image(Camera(c2w[i], fx=fx, width=width, height=height))
                            for i in range(n_images)])
    return PosedImages(c2w, imgs.clamp(0, 1), fx, width, height)


def load_nerf_synthetic(root: str, split: str, device, scene_scale: float = 2.0 / 3.0,
                        white_bkgd: bool = True, n_images: Optional[int] = None) -> PosedImages:
    """NeRF-synthetic (Blender) scene directory -> PosedImages (behaviour of
    opt/util/nerf_dataset.py:14-120): ``transforms_<split>.json`` + ``<split>/<name>.png``,
    OpenGL -> OpenCV camera convention (c2w @ diag(1,-1,-1,1)), translations scaled by
    ``scene_scale``, RGBA composited over white, focal = W / (2 tan(camera_angle_x / 2))."""
    import os
    from PIL import Image
    with open(os.path.join(root, f"transforms_{split}.json")) as f:
        meta = json.load(f)
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
    c2w, imgs = [], []
    for frame in meta["frames"][:n_images]:
        path = os.path.join(root, split, os.path.basename(frame["file_path"]) + ".png")
        im = torch.from_numpy(np.asarray(Image.open(path), dtype=np.uint8).copy()).float() / 255.0
        if im.size(-1) == 4:
            im = im[..., :3] * im[..., 3:] + (1.0 - im[..., 3:]) if white_bkgd else im[..., :3]
        m = torch.tensor(frame["transform_matrix"], dtype=torch.float32) @ flip
        m[:3, 3] *= scene_scale
        c2w.append(m)
        imgs.append(im)
    images = torch.stack(imgs).to(device)
    h, w = images.shape[1:3]
    fx = float(0.5 * w / math.tan(0.5 * meta["camera_angle_x"]))
    return PosedImages(torch.stack(c2w).to(device), images, fx, w, h)


def save_nerf_synthetic(data: PosedImages, root: str, split: str, scene_scale: float = 2.0 / 3.0):
    """Inverse of ``load_nerf_synthetic``: writes ``transforms_<split>.json`` + ``<split>/r_<i>.png``
    in the NeRF-synthetic (Blender) layout the reference's loader reads
    (opt/util/nerf_dataset.py:57-98): OpenGL camera convention, unscaled translations."""
    import os
    from PIL import Image
    os.makedirs(os.path.join(root, split), exist_ok=True)
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
    frames = []
    for i in range(data.c2w.size(0)):
        m = data.c2w[i].detach().cpu().float().clone()
        if m.size(0) == 3:
            m = torch.cat([m, torch.tensor([[0.0, 0.0, 0.0, 1.0]])])
        m[:3, 3] /= scene_scale
        m = m @ flip
        im = (data.images[i].detach().clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).cpu().numpy()
        Image.fromarray(im).save(os.path.join(root, split, f"r_{i}.png"))
        frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": m.tolist()})
    meta = {"camera_angle_x": 2.0 * math.atan(0.5 * data.width / data.fx), "frames": frames}
    with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
        json.dump(meta, f)


class Trainer:
    """One process of the training job.  ``dist_mode``: "slab" (grid sharded) or "rays"."""

    def __init__(self, cfg: TrainConfig, train: PosedImages, test: Optional[PosedImages] = None,
                 world_size: int = 1, rank: int = 0, dist_mode: str = "slab", log=print):
        self.cfg, self.train, self.test = cfg, train, test
        self.world_size, self.rank, self.dist_mode, self.log = world_size, rank, dist_mode, log
        dev = train.images.device
        self.reso_list = cfg.reso_list()
        self.reso_id = 0
        torch.manual_seed(cfg.seed)
        np.random.seed(cfg.seed)
        # scene bounds, sphere bound and MSI background as the dataset says (opt.py:280-304)
        nlayers = cfg.background_nlayers if getattr(train, "should_use_background", True) else 0
        self.grid = SparseGrid(reso=self.reso_list[0], basis_dim=cfg.basis_dim,
                               radius=list(getattr(train, "scene_radius", (1.0, 1.0, 1.0))),
                               center=list(getattr(train, "scene_center", (0.0, 0.0, 0.0))),
                               use_sphere_bound=bool(getattr(train, "use_sphere_bound", False)) and nlayers == 0,
                               background_nlayers=nlayers, background_reso=cfg.background_reso,
                               device=dev)
        with torch.no_grad():
            self.grid.sh_data.data[:] = 0.0
            self.grid.density_data.data[:] = cfg.init_sigma
            if self.grid.use_background:
                self.grid.background_data.data[..., -1] = cfg.init_sigma_bg
        o = self.grid.opt
        o.step_size, o.sigma_thresh, o.stop_thresh = cfg.step_size, cfg.sigma_thresh, cfg.stop_thresh
        o.background_brightness = cfg.background_brightness
        o.last_sample_opaque, o.near_clip = cfg.last_sample_opaque, cfg.near_clip
        self.pool = RayPool(train, seed=cfg.seed)      # same seed on every rank: same permutation
        self.lr_sigma = expon_lr(cfg.lr_sigma, cfg.lr_sigma_final, cfg.lr_sigma_delay_steps,
                                 cfg.lr_sigma_delay_mult, cfg.lr_sigma_decay_steps)
        self.lr_sh = expon_lr(cfg.lr_sh, cfg.lr_sh_final, cfg.lr_sh_delay_steps,
                              cfg.lr_sh_delay_mult, cfg.lr_sh_decay_steps)
        self.lambda_tv, self.lambda_tv_sh = cfg.lambda_tv, cfg.lambda_tv_sh
        self.step_id = 0
        self.last_upsamp = cfg.init_iters
        self._make_stepper()

    def _make_stepper(self):
        from .dist import ShardedTrainer
        from .slab import SlabTrainer
        c = self.cfg
        kw = dict(world_size=self.world_size, rank=self.rank, lr_sigma=c.lr_sigma, lr_sh=c.lr_sh,
                  lambda_tv=self.lambda_tv, lambda_tv_sh=self.lambda_tv_sh, tv_sparsity=c.tv_sparsity,
                  rms_beta=c.rms_beta, beta_loss=c.lambda_beta, sparsity_loss=c.lambda_sparsity)
        # MSI scenes use the ray-sharded step (the slab step has no background path)
        self.slab = self.world_size > 1 and self.dist_mode == "slab" and not self.grid.use_background
        if self.slab:
            self.stepper = SlabTrainer(self.grid, **kw)
        else:
            self.stepper = ShardedTrainer(
                self.grid, lr_sigma_bg=c.lr_sigma_bg, lr_color_bg=c.lr_color_bg,
                lambda_tv_background_sigma=c.lambda_tv_background_sigma,
                lambda_tv_background_color=c.lambda_tv_background_color,
                tv_background_sparsity=c.tv_background_sparsity, **kw)

    # ---- one epoch of batches -------------------------------------------------------------
    def train_epoch(self, max_steps: Optional[int] = None):
        c, W = self.cfg, self.world_size
        gq = c.batch_size * W                       # global batch: batch_size rays per GPU
        self.pool.shuffle()
        n_batches = len(self.pool) // gq
        if n_batches == 0:
            raise ValueError(
                f"the ray pool ({len(self.pool)} rays) is smaller than one global batch "
                f"({c.batch_size} rays x {W} GPUs): lower batch_size or use fewer GPUs")
        sq_err = torch.zeros((), device=self.pool.rgb.device)
        seen = 0
        for b in range(n_batches):
            if max_steps is not None and self.step_id >= max_steps:
                break
            lr_s, lr_c = self.lr_sigma(self.step_id), self.lr_sh(self.step_id)
            if self.slab:
                rays, gt = self.pool.batch(b * gq, (b + 1) * gq)
                rgb = self.stepper.step(rays, gt, lr_sigma=lr_s, lr_sh=lr_c)
            else:
                lo = b * gq + self.rank * c.batch_size
                rays, gt = self.pool.batch(lo, lo + c.batch_size)
                rgb = self.stepper.step(rays, gt, global_batch=gq, lr_sigma=lr_s, lr_sh=lr_c)
            fused_mse = getattr(self.stepper, "last_mse", None) \
                if getattr(self.stepper, "_step_ws", None) is not None else None
            sq_err += fused_mse if fused_mse is not None else ((rgb.detach() - gt) ** 2).mean()
            seen += 1
            self.step_id += 1
            if seen % c.print_every == 0 and self.slab:
                self.stepper.check_overflow()               # every rank, at the logging sync point
            if seen % c.print_every == 0 and self.rank == 0:
                mse = (sq_err / c.print_every).item()       # the only host sync of the loop
                sq_err.zero_()
                self.log(f"step {self.step_id} mse {mse:.6f} psnr {-10 * math.log10(max(mse, 1e-12)):.2f} "
                         f"lr_sigma {lr_s:.3g} lr_sh {lr_c:.3g}")
        self.maybe_upsample()

    def maybe_upsample(self):
        """opt.py:616-648: every upsamp_every steps move to the next resolution."""
        c = self.cfg
        if self.step_id - self.last_upsamp < c.upsamp_every:
            return
        self.last_upsamp = self.step_id
        if self.reso_id >= len(self.reso_list) - 1:
            return
        if self.slab:
            self.stepper.gather_params()            # resample reads the whole grid
        if c.tv_early_only > 0:
            self.lambda_tv = self.lambda_tv_sh = 0.0
        elif c.tv_decay != 1.0:
            self.lambda_tv *= c.tv_decay
            self.lambda_tv_sh *= c.tv_decay
        self.reso_id += 1
        reso = self.reso_list[self.reso_id]
        cams = [self.train.camera(i) for i in range(self.train.c2w.size(0))] \
            if c.thresh_type == "weight" else None
        self.grid.resample(reso=reso, sigma_thresh=c.density_thresh,
                           weight_thresh=c.weight_thresh / reso[2], dilate=2, cameras=cams,
                           max_elements=c.max_grid_elements)
        if self.grid.use_background and self.reso_id <= 1:      # opt.py:636-637
            self.grid.sparsify_background(c.background_density_thresh)
        if c.upsample_density_add:
            self.grid.density_data.data[:] += c.upsample_density_add
        if self.rank == 0:
            self.log(f"* upsampled to {reso}: {self.grid.density_data.size(0)} voxels")
        self._make_stepper()

    @torch.no_grad()
    def eval_psnr(self, data: Optional[PosedImages] = None) -> float:
        data = data or self.test or self.train
        if self.slab:
            self.stepper.gather_params()
        tot = 0.0
        for i in range(data.c2w.size(0)):
            img = self.grid.volume_render_image(data.camera(i)).clamp_(0, 1)
            tot += -10.0 * math.log10(max(((img - data.images[i]) ** 2).mean().item(), 1e-12))
        return tot / data.c2w.size(0)

    def eval_metrics(self, data=None, **kw):
        """PSNR / SSIM (/ LPIPS) per test camera: the loop of opt/render_imgs.py (evaluate.py)."""
        from .evaluate import evaluate
        if self.slab:
            self.stepper.gather_params()
        return evaluate(self.grid, data or self.test or self.train, **kw)

    def run(self):
        t0 = time.time()
        while self.step_id < self.cfg.n_iters:
            self.train_epoch(max_steps=self.cfg.n_iters)
        if self.rank == 0:
            self.log(f"done: {self.step_id} steps in {time.time() - t0:.1f} s")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=None, help="JSON with TrainConfig fields (e.g. opt/configs/syn.json)")
    ap.add_argument("--dist", default="slab", choices=["slab", "rays"])
    ap.add_argument("--save", default=None, help="write ckpt.npz here at the end")
    ap.add_argument("--data_dir", default=None,
                    help="scene directory: NeRF-synthetic (transforms_train.json), LLFF (poses_bounds.npy) "
                         "or NSVF / T&T (intrinsics.txt, pose/, rgb/), detected as opt/util/dataset.py does; "
                         "default: a synthetic posed-image set rendered on the fly")
    for f in dataclasses.fields(TrainConfig):
        kind = (lambda v: v.lower() in ("1", "true", "yes")) if isinstance(f.default, bool) else type(f.default)
        ap.add_argument(f"--{f.name}", type=kind, default=None)
    args = ap.parse_args()
    over = {f.name: getattr(args, f.name) for f in dataclasses.fields(TrainConfig)
            if getattr(args, f.name) is not None}
    cfg = TrainConfig.from_json(args.config, **over) if args.config else TrainConfig(**over)
    import os
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.data_dir:
        from .datasets import auto_dataset
        train = auto_dataset(args.data_dir, "train", dev)
        test = auto_dataset(args.data_dir, "test", dev)
    else:
        data = synthetic_dataset(dev, n_images=28)
        test = PosedImages(data.c2w[24:], data.images[24:], data.fx, data.width, data.height)
        train = PosedImages(data.c2w[:24], data.images[:24], data.fx, data.width, data.height)
    tr = Trainer(cfg, train, test, world_size=world, rank=rank, dist_mode=args.dist)
    tr.run()
    res = tr.eval_metrics(log=print if rank == 0 else None)
    if rank == 0:
        print(f"test PSNR {res['psnr']:.2f} dB  SSIM {res['ssim']:.4f}")
        if args.save:
            tr.grid.save(args.save)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
