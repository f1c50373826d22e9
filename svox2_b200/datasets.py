"""Posed-image datasets either side of the hot path: NeRF-synthetic (Blender),
LLFF forward-facing (NDC) and NSVF / Tanks&Temples, with ray generation ON THE GPU.

The reference builds every training ray on the host in float32 and keeps the pool there
(opt/util/dataset_base.py:36-71, llff_dataset.py:183-195, util.py:178-190); here the images and
cameras are uploaded once and ``generate_rays`` / ``PosedImageSet.gen_rays`` produce the
``[n*H*W, 3]`` origins / unit directions (world or NDC) directly in HBM with the same arithmetic:

    d_cam = normalise(((x + .5 - cx)/fx, (y + .5 - cy)/fy, 1));  d_world = R d_cam;  o = t
    NDC (utils.convert_to_ndc, utils.py:576-596) followed by a re-normalisation (llff_dataset.py:189)

File formats follow the reference's loaders (behaviour, not code):
  * NSVF / T&T  (opt/util/nsvf_dataset.py): ``intrinsics.txt``, ``pose/*.txt`` (c2w, OpenCV),
    ``rgb/*.png``; split by file prefix ``0_ / 1_ / 2_``; poses normalised by the camera
    similarity transform (util.py:319-373) times ``cam_scale_factor``.
  * LLFF (opt/util/llff_dataset.py + load_llff.py): ``poses_bounds.npy`` + ``images[_k]/``;
    [down, right, back] -> [right, up, back] axis fix, 1/(0.75 * min bound) rescale, recentering by
    the average pose, OpenGL -> OpenCV flip, every ``hold_every``-th image held out, scene radius
    padded by ``offset`` pixels, ``ndc_coeffs = (2 fx / W, 2 fy / H)``.
  * NeRF-synthetic: ``train.load_nerf_synthetic`` (re-exported here).
Image decoding uses PIL; area down-scaling uses OpenCV when present (as upstream), else PIL.
"""
import dataclasses
import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import utils
from .grid import Camera, Rays


@dataclasses.dataclass
class Intrin:
    """Pinhole intrinsics shared by all images of a set (util.py Intrin)."""
    fx: float
    fy: float
    cx: float
    cy: float

    def scale(self, s: float) -> "Intrin":
        return Intrin(self.fx * s, self.fy * s, self.cx * s, self.cy * s)

    def get(self, field: str, image_id: int = 0) -> float:
        return float(getattr(self, field))


def generate_rays(c2w: torch.Tensor, intrin: Intrin, height: int, width: int,
                  ndc_coeffs: Sequence[float] = (-1.0, -1.0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """All pixel rays of ``c2w`` [n, 3|4, 4] (OpenCV convention) -> origins, unit dirs [n, H*W, 3]
    on ``c2w.device``, float32 (dataset_base.py:36-63; with ``ndc_coeffs`` > 0 also
    llff_dataset.py:183-195).  The rotation is applied as three scaled column adds in a fixed order,
    so the result does not depend on the device's matmul kernels."""
    dev = c2w.device
    c2w = c2w.to(torch.float32)
    ys = (torch.arange(height, dtype=torch.float32, device=dev) + 0.5 - float(intrin.cy)) / float(intrin.fy)
    xs = (torch.arange(width, dtype=torch.float32, device=dev) + 0.5 - float(intrin.cx)) / float(intrin.fx)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    d_cam = torch.stack((xx, yy, torch.ones_like(xx)), dim=-1)
    d_cam = d_cam / torch.norm(d_cam, dim=-1, keepdim=True)
    dx, dy, dz = (d_cam[..., i].reshape(1, -1, 1) for i in range(3))
    R = c2w[:, None, :3, :3]                                          # [n,1,3,3]
    dirs = R[..., 0] * dx + R[..., 1] * dy + R[..., 2] * dz           # [n,HW,3]
    origins = c2w[:, None, :3, 3].expand(-1, height * width, -1).contiguous()
    if ndc_coeffs[0] > 0.0:
        origins, dirs = utils.convert_to_ndc(origins, dirs, ndc_coeffs)
        dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    return origins.contiguous(), dirs.contiguous()


@dataclasses.dataclass
class PosedImageSet:
    """One split of a dataset, resident on the device; duck-compatible with ``train.PosedImages``
    and with what opt.py / render_imgs.py read from the reference's dataset objects."""
    c2w: torch.Tensor                      # [n,4,4] OpenCV c2w
    gt: torch.Tensor                       # [n,H,W,3] float32 in [0,1]
    intrins: Intrin
    split: str = "train"
    ndc_coeffs: Tuple[float, float] = (-1.0, -1.0)
    scene_center: Sequence[float] = (0.0, 0.0, 0.0)
    scene_radius: Sequence[float] = (1.0, 1.0, 1.0)
    use_sphere_bound: bool = True
    should_use_background: bool = True
    scene_scale: float = 1.0
    render_c2w: Optional[torch.Tensor] = None
    names: Optional[List[str]] = None

    @property
    def images(self):
        return self.gt

    @property
    def n_images(self):
        return self.c2w.size(0)

    @property
    def height(self):
        return self.gt.size(1)

    @property
    def width(self):
        return self.gt.size(2)

    h = h_full = height
    w = w_full = width

    @property
    def fx(self):
        return self.intrins.fx

    def get_image_size(self, i: int):
        return self.height, self.width

    def camera(self, i: int, crop: float = 1.0) -> Camera:
        """Camera of image i; ``crop`` < 1 renders the central crop (render_imgs.py:176-187)."""
        W, H = self.width, self.height
        w = W if crop == 1.0 else int(W * crop)
        h = H if crop == 1.0 else int(H * crop)
        k = self.intrins
        return Camera(self.c2w[i], k.fx, k.fy, k.cx + (w - W) * 0.5, k.cy + (h - H) * 0.5, w, h,
                      ndc_coeffs=tuple(self.ndc_coeffs))

    def gen_rays(self, factor: int = 1):
        """(Rays, gt[n*h*w, 3]) of all images at 1/factor resolution (area-averaged colours), built
        on the device (dataset_base.py:36-71)."""
        h, w = self.height // factor, self.width // factor
        k = self.intrins.scale(1.0 / (self.height / h))
        origins, dirs = generate_rays(self.c2w, k, h, w, self.ndc_coeffs)
        gt = self.gt
        if factor != 1:
            gt = torch.nn.functional.interpolate(gt.permute(0, 3, 1, 2), size=(h, w),
                                                 mode="area").permute(0, 2, 3, 1)
        return Rays(origins.view(-1, 3), dirs.view(-1, 3)), gt.reshape(-1, 3).contiguous()


# ----------------------------------------------------------------------------------------
# image IO
# ----------------------------------------------------------------------------------------
def _read_image(path: str) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(path)).copy()


def _resize_area(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    try:
        import cv2
        return cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_AREA)
    except ImportError:  # pragma: no cover
        from PIL import Image
        return np.asarray(Image.fromarray(img).resize((new_w, new_h), Image.BOX))


def _to_rgb01(stack: np.ndarray, white_bkgd: bool, f64: bool) -> torch.Tensor:
    gt = torch.from_numpy(stack)
    gt = (gt.double() if f64 else gt.float()) / 255.0
    if gt.size(-1) == 4:
        gt = gt[..., :3] * gt[..., 3:] + (1.0 - gt[..., 3:]) if white_bkgd else gt[..., :3]
    return gt.float()


def _prefix_sort_key(name: str) -> str:
    return name[2:] if len(name) > 2 and name[1] == "_" else name


# ----------------------------------------------------------------------------------------
# NSVF / Tanks&Temples
# ----------------------------------------------------------------------------------------
def similarity_from_cameras(c2w: np.ndarray):
    """Similarity (4x4, scale) that makes z+ the mean camera-up axis, puts the median of the
    points where the optical axes pass closest to the origin at the origin, and scales the median
    camera distance to 1 (behaviour of util.py:319-373)."""
    t = c2w[:, :3, 3]
    R = c2w[:, :3, :3]
    cam_up = np.array([0.0, -1.0, 0.0])
    world_up = (R @ cam_up).mean(axis=0)
    world_up /= np.linalg.norm(world_up)
    cosang = float(cam_up @ world_up)
    axis = np.cross(world_up, cam_up)
    K = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
    if cosang > -1:
        R_align = np.eye(3) + K + (K @ K) * 1 / (1 + cosang)   # Rodrigues, rotating world_up onto cam_up
    else:
        R_align = np.diag([-1.0, 1.0, 1.0])
    R = R_align @ R
    fwd = R[:, :, 2]
    t = (R_align @ t[..., None])[..., 0]
    nearest = t + (fwd * -t).sum(-1)[:, None] * fwd
    translate = -np.median(nearest, axis=0)
    T = np.eye(4)
    T[:3, 3] = translate
    T[:3, :3] = R_align
    scale = 1.0 / np.median(np.linalg.norm(t + translate, axis=-1))
    return T, scale


def _find_dir(root: str, candidates: Sequence[str]) -> str:
    for c in candidates:
        if os.path.isdir(os.path.join(root, c)):
            return c
    raise FileNotFoundError(f"none of {list(candidates)} found in {root}")


def _load_pose_txt(path: str) -> np.ndarray:
    m = np.loadtxt(path).reshape(-1, 4)
    if len(m) == 3:
        m = np.concatenate([m, np.array([[0.0, 0.0, 0.0, 1.0]])], axis=0)
    return m


def load_nsvf(root: str, split: str, device="cuda", scene_scale: Optional[float] = None,
              scale: Optional[float] = 1.0, white_bkgd: bool = True, normalize_by_camera: bool = True,
              cam_scale_factor: float = 0.95, **_unused) -> PosedImageSet:
    """NSVF-format scene directory (also Tanks&Temples / BlendedMVS as packaged for NSVF)."""
    scene_scale = 1.0 if scene_scale is None else scene_scale
    scale = 1.0 if scale is None else scale
    img_dir = _find_dir(root, ["images", "image", "rgb"])
    pose_dir = _find_dir(root, ["poses", "pose"])
    every = sorted(os.listdir(os.path.join(root, img_dir)), key=_prefix_sort_key)
    if split in ("train", "test_train"):
        files = [f for f in every if f.startswith("0_")]
    elif split == "val":
        files = [f for f in every if f.startswith("1_")]
    elif split == "test":
        files = [f for f in every if f.startswith("2_")] or [f for f in every if f.startswith("1_")]
    else:
        files = every
    if not files:  # no split prefixes: every 16th image is held out
        files = [f for i, f in enumerate(every) if i % 16 != 0] if split == "train" else every[::16]
    if not files:
        raise FileNotFoundError(f"no matching images in {os.path.join(root, img_dir)}")

    resize = scale < 1
    if resize and abs(1.0 / scale - round(1.0 / scale)) < 1e-9:
        pre = f"{img_dir}_{round(1.0 / scale)}"
        if os.path.exists(os.path.join(root, pre)):
            img_dir, resize = pre, False
    imgs, poses, full = [], [], (0, 0)
    new_h = new_w = 0
    for f in files:
        im = _read_image(os.path.join(root, img_dir, f))
        poses.append(_load_pose_txt(os.path.join(root, pose_dir, os.path.splitext(f)[0] + ".txt")))
        full = im.shape[:2]
        new_h, new_w = round(full[0] * scale), round(full[1] * scale)
        if resize:
            im = _resize_area(im, new_w, new_h)
        imgs.append(im)
    c2w = torch.from_numpy(np.stack(poses))                      # float64
    if normalize_by_camera:
        all_pose_files = sorted(os.listdir(os.path.join(root, pose_dir)), key=_prefix_sort_key)
        all_poses = np.stack([np.loadtxt(os.path.join(root, pose_dir, f)).reshape(-1, 4)
                              for f in all_pose_files], axis=0)
        T, s = similarity_from_cameras(all_poses)
        c2w = torch.from_numpy(T) @ c2w
        scene_scale = cam_scale_factor * s
    c2w[:, :3, 3] *= scene_scale
    gt = _to_rgb01(np.stack(imgs), white_bkgd, f64=True)

    ipath = os.path.join(root, "intrinsics.txt")
    try:
        K = np.loadtxt(ipath)
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    except Exception:  # "f cx cy ..." on the first line (original NSVF synthetic scenes)
        with open(ipath) as fh:
            first = fh.readline().split()
        fx = fy = float(first[0])
        cx, cy = float(first[1]), float(first[2])
    if scale < 1.0:
        sw, sh = new_w / full[1], new_h / full[0]
        fx, cx, fy, cy = fx * sw, cx * sw, fy * sh, cy * sh
    return PosedImageSet(c2w.float().to(device), gt.to(device), Intrin(float(fx), float(fy), float(cx), float(cy)),
                         split=split, scene_scale=float(scene_scale), names=list(files))


# ----------------------------------------------------------------------------------------
# LLFF forward-facing
# ----------------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v)


def _average_pose(poses: np.ndarray) -> np.ndarray:
    """[n,3,4] -> 3x4: mean position, mean viewing axis, orthonormalised (load_llff.py:165-174)."""
    pos = poses[:, :3, 3].mean(0)
    z = _unit(poses[:, :3, 2].sum(0))
    x = _unit(np.cross(poses[:, :3, 1].sum(0), z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, pos], 1)


def llff_poses(poses_bounds: np.ndarray, bd_factor: float = 0.75):
    """``poses_bounds.npy`` [n,17] -> (c2w OpenCV [n,4,4] float32, (H, W, focal), bounds [n,2])
    after the LLFF axis fix, the bound rescale and the recentering (load_llff.py:318-345,
    llff_dataset.py:560-589)."""
    arr = np.asarray(poses_bounds)
    p = arr[:, :-2].reshape(-1, 3, 5)                                 # [n,3,5]
    bds = arr[:, -2:].astype(np.float32)
    hwf = p[0, :, 4].copy()
    m = np.concatenate([p[:, :, 1:2], -p[:, :, 0:1], p[:, :, 2:4]], axis=2).astype(np.float32)  # [n,3,4]
    sc = 1.0 if bd_factor is None else 1.0 / (bds.min() * bd_factor)
    m[:, :3, 3] *= sc
    bds = bds * sc
    bottom = np.array([[0.0, 0.0, 0.0, 1.0]])
    avg = np.concatenate([_average_pose(m), bottom], 0)
    m44 = np.concatenate([m, np.tile(bottom[None], (m.shape[0], 1, 1))], 1)
    rec = (np.linalg.inv(avg) @ m44).astype(np.float32)[:, :3, :4]
    rec = rec.astype(np.float32)
    # OpenGL -> OpenCV (nerf_pose_to_ours): negate y,z of the centre and the matching rotation entries
    c2w = np.tile(np.eye(4, dtype=np.float32)[None], (rec.shape[0], 1, 1))
    R = rec[:, :3, :3].copy()
    R[:, 1:, 0] *= -1
    R[:, 0, 1:] *= -1
    center = rec[:, :3, 3].copy()
    center[:, 1:] *= -1
    c2w[:, :3, :3] = R
    c2w[:, :3, 3] = center
    return c2w, hwf, bds


def load_llff(root: str, split: str, device="cuda", scale: Optional[float] = 1.0 / 4.0,
              hold_every: int = 8, offset: int = 250, **_unused) -> PosedImageSet:
    """LLFF scene directory (``poses_bounds.npy`` + ``images/``) -> NDC dataset."""
    scale = 1.0 / 4.0 if scale is None else scale
    arr = np.load(os.path.join(root, "poses_bounds.npy"))
    img_dir, prescaled = "images", False
    if scale != 1 and abs(1.0 / scale - round(1.0 / scale)) < 1e-9:
        cand = f"images_{round(1.0 / scale)}"
        if os.path.isdir(os.path.join(root, cand)):
            img_dir, prescaled = cand, True
    exts = (".png", ".jpg", ".jpeg", ".exr")
    files = sorted([f for f in os.listdir(os.path.join(root, img_dir))
                    if not f.startswith(".") and f.lower().endswith(exts)], key=_prefix_sort_key)
    c2w, (H, W, focal), _ = llff_poses(arr)
    if os.path.isfile(os.path.join(root, "hwf_cxcy.npy")):
        raise NotImplementedError("hwf_cxcy.npy (per-scene principal point) is not supported")
    if len(files) != c2w.shape[0]:
        raise ValueError(f"{len(files)} images but {c2w.shape[0]} poses in {root}")
    new_w, new_h = round(int(W) * scale), round(int(H) * scale)
    sw, sh = new_w / int(W), new_h / int(H)
    k = Intrin(float(focal) * sw, float(focal) * sh, float(W) / 2.0 * sw, float(H) / 2.0 * sh)
    want_train = split.endswith("train")
    keep = [i for i in range(len(files)) if (i % hold_every > 0) == want_train]
    imgs = []
    for i in keep:
        im = _read_image(os.path.join(root, img_dir, files[i]))
        if scale != 1 and not prescaled:
            im = _resize_area(im, round(im.shape[1] * scale), round(im.shape[0] * scale))
        imgs.append(im)
    gt = _to_rgb01(np.stack(imgs), white_bkgd=True, f64=False)
    if gt.shape[1] != new_h or gt.shape[2] != new_w:
        raise ValueError("image size does not match the intrinsics in poses_bounds.npy")
    radius = [1 + 2 * offset / gt.size(2), 1 + 2 * offset / gt.size(1), 1.0]
    return PosedImageSet(torch.from_numpy(c2w[keep]).to(device), gt.to(device), k, split=split,
                         ndc_coeffs=(2 * k.fx / new_w, 2 * k.fy / new_h), scene_radius=radius,
                         use_sphere_bound=False, should_use_background=False,
                         names=[files[i] for i in keep])


def load_nerf_synthetic(root: str, split: str, device="cuda", **kw) -> PosedImageSet:
    """NeRF-synthetic (Blender) scene -> PosedImageSet (``train.load_nerf_synthetic``)."""
    from .train import load_nerf_synthetic as _load
    d = _load(root, split, device, **kw)
    return PosedImageSet(d.c2w, d.images, Intrin(d.fx, d.fx, d.width * 0.5, d.height * 0.5), split=split,
                         should_use_background=False)


def auto_dataset(root: str, split: str, device="cuda", **kw) -> PosedImageSet:
    """Format detection as opt/util/dataset.py:10-28."""
    if os.path.isfile(os.path.join(root, "poses_bounds.npy")):
        return load_llff(root, split, device, **kw)
    if os.path.isfile(os.path.join(root, "transforms.json")) or \
            os.path.isfile(os.path.join(root, "transforms_train.json")):
        return load_nerf_synthetic(root, split, device, **{k: v for k, v in kw.items()
                                                           if k in ("scene_scale", "white_bkgd", "n_images")})
    return load_nsvf(root, split, device, **kw)
