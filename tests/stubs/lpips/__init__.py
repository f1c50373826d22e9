"""Stand-in for `lpips` (absent from this image; its VGG weights need the network).  Importable so
that opt/render_imgs.py starts; constructing the metric raises -- run with --no_lpips."""


class LPIPS:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("lpips is not available in this image (no network for VGG weights)")
