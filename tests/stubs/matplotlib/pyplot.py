import numpy as np


def get_cmap(name="viridis"):
    """A 3-stop approximation of viridis: callable on [0,1] arrays -> RGBA, like a Colormap."""
    stops = np.array([[0.267, 0.005, 0.329], [0.128, 0.567, 0.551], [0.993, 0.906, 0.144]])

    def cmap(x):
        x = np.clip(np.asarray(x, np.float64), 0.0, 1.0) * 2.0
        lo = np.clip(np.floor(x).astype(int), 0, 1)
        w = (x - lo)[..., None]
        rgb = stops[lo] * (1 - w) + stops[lo + 1] * w
        return np.concatenate([rgb, np.ones_like(rgb[..., :1])], -1)
    return cmap
