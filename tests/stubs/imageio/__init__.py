"""Minimal stand-in for `imageio` (absent from this image) so that the reference's UNMODIFIED
opt/opt.py, opt/render_imgs.py and opt/util/*_dataset.py can be executed by the drop-in tests.
PIL-backed: imread / imwrite, and a mimwrite that stores the frames as a numbered PNG sequence."""
import os

import numpy as np
from PIL import Image


def imread(path):
    return np.asarray(Image.open(path)).copy()


def imwrite(path, im, **kwargs):
    Image.fromarray(np.asarray(im)).save(path)


def mimwrite(path, frames, **kwargs):
    base = os.path.splitext(path)[0]
    for i, f in enumerate(frames):
        Image.fromarray(np.asarray(f)).save(f"{base}_{i:04d}.png")
