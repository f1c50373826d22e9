"""Evaluation loop of the reference's opt/render_imgs.py (:150-241) on the B200 renderer: every
``n_eval``-th test camera is rendered with ``volume_render_image`` (the fused image kernel under
``no_grad``), clamped to [0,1] and scored with PSNR, SSIM and -- when a perceptual network is
supplied -- LPIPS.  Metrics stay on the device; one host read per image (as upstream's ``.item()``).

``compute_ssim`` restates opt/util/util.py:193-271 (itself modelled on tf.image.ssim): an 11-tap
Gaussian (sigma 1.5) applied separably with zero padding, variances clipped at 0, covariance
clipped to the Cauchy-Schwarz bound, k1 = 0.01, k2 = 0.03.

LPIPS needs pretrained VGG weights, which this image does not ship (no network): pass
``lpips_fn(gt_chw, pred_chw) -> scalar tensor`` if you have them (render_imgs.py:117-119 builds
``lpips.LPIPS(net='vgg')``); otherwise the field is None, like ``--no_lpips`` upstream.
"""
import math
import os
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


def compute_ssim(img0: torch.Tensor, img1: torch.Tensor, max_val: float = 1.0, filter_size: int = 11,
                 filter_sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03,
                 return_map: bool = False) -> torch.Tensor:
    """Mean SSIM of two images [..., H, W, C] (or the per-pixel map)."""
    shape = img0.shape
    H, W, C = shape[-3:]
    a = img0.reshape(-1, H, W, C).permute(0, 3, 1, 2)
    b = img1.reshape(-1, H, W, C).permute(0, 3, 1, 2)
    half = filter_size // 2
    offs = torch.arange(filter_size, device=a.device) - half + (2 * half - filter_size + 1) / 2
    g = torch.exp(-0.5 * (offs / filter_sigma) ** 2)
    g = g / g.sum()
    gv = g.view(1, 1, -1, 1).repeat(C, 1, 1, 1)
    gh = g.view(1, 1, 1, -1).repeat(C, 1, 1, 1)

    def blur(z):  # rows first, then columns (zero padding), one filter per channel
        z = F.conv2d(z, gh, padding=[0, half], groups=C)
        return F.conv2d(z, gv, padding=[half, 0], groups=C)

    mu_a, mu_b = blur(a), blur(b)
    aa, bb, ab = mu_a * mu_a, mu_b * mu_b, mu_a * mu_b
    var_a = torch.clamp(blur(a ** 2) - aa, min=0.0)
    var_b = torch.clamp(blur(b ** 2) - bb, min=0.0)
    cov = blur(a * b) - ab
    cov = torch.sign(cov) * torch.min(torch.sqrt(var_a * var_b), torch.abs(cov))
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    ssim_map = ((2 * ab + c1) * (2 * cov + c2)) / ((aa + bb + c1) * (var_a + var_b + c2))
    if return_map:
        return ssim_map
    return ssim_map.reshape(-1, C * H * W).mean(dim=-1).reshape(shape[:-3])


@torch.no_grad()
def evaluate(grid, dset, n_eval: int = 100000, crop: float = 1.0,
             lpips_fn: Optional[Callable] = None, out_dir: Optional[str] = None,
             save_images: bool = False, timing: bool = False, log=print) -> Dict:
    """render_imgs.py's metric loop.  ``dset``: a ``datasets.PosedImageSet`` (or anything with
    ``c2w``, ``gt``, ``n_images``, ``camera(i, crop)``).  Returns averages and per-image values;
    with ``out_dir`` also writes ``psnr.txt`` / ``ssim.txt`` (/ ``lpips.txt``) and, with
    ``save_images``, ``NNNN.png`` = ground truth | render side by side, as upstream."""
    n_images = dset.n_images
    interval = max(n_images // n_eval, 1)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    per_image = []
    for img_id in range(0, n_images, interval):
        cam = dset.camera(img_id, crop) if crop != 1.0 else dset.camera(img_id)
        im = grid.volume_render_image(cam, use_kernel=True)
        im.clamp_(0.0, 1.0)
        gt = dset.gt[img_id].to(im.device)
        if crop != 1.0:  # central crop of the ground truth to the rendered window
            H, W = gt.shape[:2]
            h, w = im.shape[:2]
            y0, x0 = (H - h) // 2, (W - w) // 2
            gt = gt[y0:y0 + h, x0:x0 + w]
        rec = {"id": img_id}
        mse = ((im - gt) ** 2).mean()
        if not timing:
            ssim = compute_ssim(gt, im)
            lp = lpips_fn(gt.permute(2, 0, 1).contiguous(), im.permute(2, 0, 1).contiguous()) \
                if lpips_fn is not None else None
            vals = torch.stack([mse, ssim] + ([lp.reshape(())] if lp is not None else [])).tolist()
            rec["ssim"] = vals[1]
            if lp is not None:
                rec["lpips"] = vals[2]
            mse_val = vals[0]
        else:
            mse_val = mse.item()
        rec["psnr"] = -10.0 * math.log10(max(mse_val, 1e-20))
        per_image.append(rec)
        if log is not None and not timing:
            log(" ".join([str(img_id), "PSNR", f"{rec['psnr']:.4f}", "SSIM", f"{rec['ssim']:.4f}"] +
                         (["LPIPS", f"{rec['lpips']:.4f}"] if "lpips" in rec else [])))
        if out_dir and save_images:
            from PIL import Image
            side = (torch.cat([gt, im], dim=1) * 255).to(torch.uint8).cpu().numpy()
            Image.fromarray(side).save(os.path.join(out_dir, f"{img_id:04d}.png"))
    n = max(len(per_image), 1)
    out = {"n_images": len(per_image), "per_image": per_image,
           "psnr": sum(r["psnr"] for r in per_image) / n,
           "ssim": (sum(r["ssim"] for r in per_image) / n) if not timing else None,
           "lpips": (sum(r["lpips"] for r in per_image) / n) if (per_image and "lpips" in per_image[0]) else None}
    if out_dir:
        for key in ("psnr", "ssim", "lpips"):
            if out[key] is not None:
                with open(os.path.join(out_dir, key + ".txt"), "w") as f:
                    f.write(str(out[key]))
    if log is not None:
        log("AVERAGES " + " ".join(f"{k.upper()}: {out[k]:.4f}" for k in ("psnr", "ssim", "lpips")
                                   if out[k] is not None))
    return out
