"""Seeded synthetic scenes shared by the parity tests (numpy; no GPU needed to build them)."""
import numpy as np


def _dilate_np(mask, it=1):
    m = mask.copy()
    for _ in range(it):
        p = np.pad(m, 1, mode="edge")
        out = np.zeros_like(m)
        for dx in range(3):
            for dy in range(3):
                for dz in range(3):
                    out |= p[dx:dx + m.shape[0], dy:dy + m.shape[1], dz:dz + m.shape[2]]
        m = out
    return m


def make_scene(reso=32, basis_dim=9, seed=0, n_rays=256, sparse=True, radius=(1.0, 1.0, 1.0),
               center=(0.0, 0.0, 0.0), n_blobs=6, sigma_peak=30.0, bg=None, inside=False,
               accel=True):
    """Returns a dict of numpy arrays: links (with skip levels if accel), density [N,1],
    sh [N,3B], offset/scaling (the _offset/_scaling of SparseGridSpec), origins, dirs, rgb_gt,
    optional MSI background (bg=(reso, nlayers))."""
    import oracle as og
    rng = np.random.default_rng(seed)
    reso3 = (reso,) * 3 if isinstance(reso, int) else tuple(reso)
    radius = np.asarray(radius, np.float32)
    center = np.asarray(center, np.float32)
    ax = [(np.arange(r, dtype=np.float32) + 0.5) / r * 2 - 1 for r in reso3]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    field = np.zeros(reso3, np.float32)
    for _ in range(n_blobs):
        c = rng.uniform(-0.55, 0.55, 3)
        r = rng.uniform(0.12, 0.35, 3)
        d2 = ((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2
        field = np.maximum(field, sigma_peak * np.clip(1.0 - d2, 0, None))
    if sparse:
        keep = _dilate_np(field > 0, 2)
    else:
        keep = np.ones(reso3, bool)
        field = field + rng.uniform(0, 2, reso3).astype(np.float32) * (rng.random(reso3) < 0.5)
    links = np.full(reso3, -1, np.int32)
    links[keep] = np.arange(int(keep.sum()), dtype=np.int32)
    N = int(keep.sum())
    density = field[keep].reshape(N, 1).astype(np.float32)
    sh = (rng.standard_normal((N, 3 * basis_dim)) * 0.15).astype(np.float32)
    sh[:, ::basis_dim] += 0.4
    if accel:
        links = og.accel_dist_prop(links)
    gsz = np.asarray(reso3, np.float32)
    offset = (0.5 * (1.0 - center / radius) * gsz - 0.5).astype(np.float32)
    scaling = (0.5 / radius * gsz).astype(np.float32)

    v = rng.standard_normal((n_rays, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    dist = 0.6 if inside else 2.7
    origins = (v * dist * radius + center).astype(np.float32)
    targets = (rng.uniform(-0.5, 0.5, (n_rays, 3)) * radius + center).astype(np.float32)
    if inside:
        targets = (rng.standard_normal((n_rays, 3)) * radius * 3 + center).astype(np.float32)
    dirs = targets - origins
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    dirs = dirs.astype(np.float32)
    rgb_gt = rng.uniform(0, 1, (n_rays, 3)).astype(np.float32)
    out = dict(links=links, density=density, sh=sh, offset=offset, scaling=scaling,
               origins=origins, dirs=dirs, rgb_gt=rgb_gt, radius=radius, center=center,
               basis_dim=basis_dim)
    if bg is not None:
        R, L = bg
        bl = np.arange(2 * R * R, dtype=np.int32).reshape(2 * R, R)
        drop = rng.random((2 * R, R)) < 0.15
        bl[drop] = -1
        bl[~drop] = np.arange(int((~drop).sum()), dtype=np.int32)
        nbg = int((~drop).sum())
        bd = np.zeros((nbg, L, 4), np.float32)
        bd[..., :3] = rng.standard_normal((nbg, L, 3)) * 0.5
        bd[..., 3] = rng.uniform(0, 1.5, (nbg, L)) * (rng.random((nbg, L)) < 0.6)
        out["bg_links"] = bl
        out["bg_data"] = bd
    return out


def oracle_grid(sc):
    import oracle as og
    return og.Grid(sc["links"], sc["density"], sc["sh"], sc["offset"], sc["scaling"],
                   sc.get("bg_links"), sc.get("bg_data"))


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / denom)


def frac_bad(a, b, tol=1e-4):
    """Fraction of elements whose error exceeds tol * max|b| (robust to a few samples that
    land in a neighbouring cell because CUDA libm and C libm differ in the last ulp)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float((np.abs(a - b) > tol * denom).mean())
