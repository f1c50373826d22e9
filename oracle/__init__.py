"""CPU parity oracle for the svox2 hot path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrapper over ``libsvox2_oracle.so`` (``svox2_oracle.c``), a scalar C
restatement of the reference's CUDA algorithms (each C function cites the reference
file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may import this package; ``svox2_b200`` never does.

Parity pinning: ``tests/test_oracle_golden.py`` checks this oracle against vectors generated
by the reference's own pure-PyTorch renderer (``tests/golden/make_golden.py``), and
``tests/test_gpu_vs_reference.py`` checks it (and the product) against the reference CUDA
extension compiled for sm_100 (``oracle/_ref``, built by ``oracle/build_ref.sh``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsvox2_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "svox2_oracle.c")
    hdr = os.path.join(_HERE, "svox2_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsvox2_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class _Grid(C.Structure):
    _fields_ = [
        ("size", C.c_int32 * 3),
        ("links", C.c_void_p),
        ("density", C.c_void_p),
        ("sh", C.c_void_p),
        ("n_rows", C.c_int64),
        ("basis_dim", C.c_int32),
        ("offset", C.c_float * 3),
        ("scaling", C.c_float * 3),
        ("bg_links", C.c_void_p),
        ("bg_data", C.c_void_p),
        ("bg_reso", C.c_int32),
        ("bg_nlayers", C.c_int32),
        ("bg_rows", C.c_int64),
    ]


class _Opt(C.Structure):
    _fields_ = [
        ("background_brightness", C.c_float),
        ("step_size", C.c_float),
        ("sigma_thresh", C.c_float),
        ("stop_thresh", C.c_float),
        ("near_clip", C.c_float),
        ("use_spheric_clip", C.c_int32),
        ("last_sample_opaque", C.c_int32),
    ]


class _Cam(C.Structure):
    _fields_ = [
        ("c2w", C.c_float * 12),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32),
        ("ndc_coeffx", C.c_float), ("ndc_coeffy", C.c_float),
    ]


class _Counts(C.Structure):
    _fields_ = [("s_skip", C.c_int64), ("s_vis", C.c_int64), ("s_act", C.c_int64),
                ("bg_act", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.og_tv.restype = C.c_float
    return _lib


def set_num_threads(n: int):
    lib().og_set_num_threads(int(n))


def set_accum_float(on: bool):
    lib().og_set_accum_float(int(bool(on)))


def get_max_threads() -> int:
    return int(lib().og_get_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Grid:
    """Plain-numpy grid: links int32[X,Y,Z], density f32[N], sh f32[N,3B] (+ optional MSI)."""

    def __init__(self, links, density, sh, offset, scaling, bg_links=None, bg_data=None):
        self.links = np.ascontiguousarray(links, dtype=np.int32)
        self.density = _f32(density).reshape(-1)
        self.sh = _f32(sh)
        assert self.sh.ndim == 2 and self.sh.shape[1] % 3 == 0
        self.offset = _f32(offset).reshape(3)
        self.scaling = _f32(scaling).reshape(3)
        self.bg_links = None if bg_links is None else np.ascontiguousarray(bg_links, dtype=np.int32)
        self.bg_data = None if bg_data is None else _f32(bg_data)

    @property
    def basis_dim(self):
        return self.sh.shape[1] // 3

    def c(self):
        g = _Grid()
        g.size[:] = self.links.shape
        g.links = _ptr(self.links)
        g.density = _ptr(self.density)
        g.sh = _ptr(self.sh)
        g.n_rows = self.density.shape[0]
        g.basis_dim = self.basis_dim
        g.offset[:] = [float(v) for v in self.offset]
        g.scaling[:] = [float(v) for v in self.scaling]
        if self.bg_links is not None and self.bg_data is not None and self.bg_data.size > 0:
            g.bg_links = _ptr(self.bg_links)
            g.bg_data = _ptr(self.bg_data)
            g.bg_reso = self.bg_links.shape[1]
            g.bg_nlayers = self.bg_data.shape[1]
            g.bg_rows = self.bg_data.shape[0]
        return g


def make_opt(background_brightness=1.0, step_size=0.5, sigma_thresh=1e-10, stop_thresh=1e-7,
             near_clip=0.0, use_spheric_clip=False, last_sample_opaque=False):
    return _Opt(background_brightness, step_size, sigma_thresh, stop_thresh, near_clip,
                int(use_spheric_clip), int(last_sample_opaque))


def make_cam(c2w, fx, fy, cx, cy, width, height, ndc_coeffx=-1.0, ndc_coeffy=-1.0):
    cam = _Cam()
    m = _f32(c2w)[:3, :4].reshape(-1)
    cam.c2w[:] = [float(v) for v in m]
    cam.fx, cam.fy, cam.cx, cam.cy = fx, fy, cx, cy
    cam.width, cam.height = width, height
    cam.ndc_coeffx, cam.ndc_coeffy = ndc_coeffx, ndc_coeffy
    return cam


def _counts_dict(c):
    return dict(s_skip=c.s_skip, s_vis=c.s_vis, s_act=c.s_act, bg_act=c.bg_act)


def volume_render(grid, opt, origins, dirs, want_counts=False):
    origins, dirs = _f32(origins), _f32(dirs)
    Q = origins.shape[0]
    rgb = np.zeros((Q, 3), np.float32)
    logT = np.zeros((Q,), np.float32)
    cnt = _Counts()
    g = grid.c()
    lib().og_volume_render(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs), C.c_int64(Q),
                           _ptr(rgb), _ptr(logT), C.byref(cnt))
    if want_counts:
        return rgb, logT, _counts_dict(cnt)
    return rgb, logT


def volume_render_image(grid, cam, opt):
    out = np.zeros((cam.height, cam.width, 3), np.float32)
    g = grid.c()
    lib().og_volume_render_image(C.byref(g), C.byref(cam), C.byref(opt), _ptr(out))
    return out


def gen_rays(cam):
    n = cam.height * cam.width
    o = np.zeros((n, 3), np.float32)
    d = np.zeros((n, 3), np.float32)
    lib().og_gen_rays(C.byref(cam), _ptr(o), _ptr(d))
    return o, d


def _alloc_grads(grid, want_bg=True):
    gd = np.zeros_like(grid.density)
    gsh = np.zeros_like(grid.sh)
    mask = np.zeros(grid.density.shape[0], np.uint8)
    gbg = mbg = None
    if want_bg and grid.bg_data is not None and grid.bg_data.size > 0:
        gbg = np.zeros_like(grid.bg_data)
        mbg = np.zeros(grid.bg_data.shape[:2], np.uint8)
    return gd, gsh, gbg, mask, mbg


def volume_render_backward(grid, opt, origins, dirs, grad_out, color_cache):
    origins, dirs = _f32(origins), _f32(dirs)
    grad_out, color_cache = _f32(grad_out), _f32(color_cache)
    gd, gsh, gbg, mask, mbg = _alloc_grads(grid)
    g = grid.c()
    lib().og_volume_render_backward(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                    C.c_int64(origins.shape[0]), _ptr(grad_out),
                                    _ptr(color_cache), _ptr(gd), _ptr(gsh), _ptr(gbg),
                                    _ptr(mask), _ptr(mbg))
    return dict(grad_density=gd, grad_sh=gsh, grad_bg=gbg, mask=mask.astype(bool),
                mask_bg=None if mbg is None else mbg.astype(bool))


def volume_render_fused(grid, opt, origins, dirs, rgb_gt, beta_loss=0.0, sparsity_loss=0.0,
                        norm_rays=None):
    origins, dirs, rgb_gt = _f32(origins), _f32(dirs), _f32(rgb_gt)
    Q = origins.shape[0]
    gd, gsh, gbg, mask, mbg = _alloc_grads(grid)
    rgb = np.zeros((Q, 3), np.float32)
    cnt = _Counts()
    g = grid.c()
    lib().og_volume_render_fused(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                 C.c_int64(Q), C.c_int64(Q if norm_rays is None else norm_rays),
                                 _ptr(rgb_gt), C.c_float(beta_loss), C.c_float(sparsity_loss),
                                 _ptr(rgb), _ptr(gd), _ptr(gsh), _ptr(gbg), _ptr(mask),
                                 _ptr(mbg), C.byref(cnt))
    return dict(rgb=rgb, grad_density=gd, grad_sh=gsh, grad_bg=gbg, mask=mask.astype(bool),
                mask_bg=None if mbg is None else mbg.astype(bool), counts=_counts_dict(cnt))


def volume_render_nvol(grid, opt, origins, dirs):
    origins, dirs = _f32(origins), _f32(dirs)
    rgb = np.zeros((origins.shape[0], 3), np.float32)
    g = grid.c()
    lib().og_volume_render_nvol(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                C.c_int64(origins.shape[0]), _ptr(rgb))
    return rgb


def volume_render_nvol_backward(grid, opt, origins, dirs, grad_out, color_cache):
    origins, dirs = _f32(origins), _f32(dirs)
    grad_out, color_cache = _f32(grad_out), _f32(color_cache)
    gd, gsh, _, mask, _ = _alloc_grads(grid, want_bg=False)
    g = grid.c()
    lib().og_volume_render_nvol_backward(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                         C.c_int64(origins.shape[0]), _ptr(grad_out),
                                         _ptr(color_cache), _ptr(gd), _ptr(gsh), _ptr(mask))
    return dict(grad_density=gd, grad_sh=gsh, mask=mask.astype(bool))


def volume_render_nvol_fused(grid, opt, origins, dirs, rgb_gt, sparsity_loss=0.0, norm_rays=None):
    origins, dirs, rgb_gt = _f32(origins), _f32(dirs), _f32(rgb_gt)
    Q = origins.shape[0]
    gd, gsh, _, mask, _ = _alloc_grads(grid, want_bg=False)
    rgb = np.zeros((Q, 3), np.float32)
    g = grid.c()
    lib().og_volume_render_nvol_fused(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                      C.c_int64(Q), C.c_int64(Q if norm_rays is None else norm_rays),
                                      _ptr(rgb_gt), C.c_float(sparsity_loss), _ptr(rgb), _ptr(gd),
                                      _ptr(gsh), _ptr(mask))
    return dict(rgb=rgb, grad_density=gd, grad_sh=gsh, mask=mask.astype(bool))


def render_expected_term(grid, opt, origins, dirs):
    origins, dirs = _f32(origins), _f32(dirs)
    out = np.zeros(origins.shape[0], np.float32)
    g = grid.c()
    lib().og_render_expected_term(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                  C.c_int64(origins.shape[0]), _ptr(out))
    return out


def render_sigma_thresh(grid, opt, origins, dirs, sigma_thresh):
    origins, dirs = _f32(origins), _f32(dirs)
    out = np.zeros(origins.shape[0], np.float32)
    g = grid.c()
    lib().og_render_sigma_thresh(C.byref(g), C.byref(opt), _ptr(origins), _ptr(dirs),
                                 C.c_int64(origins.shape[0]), C.c_float(sigma_thresh), _ptr(out))
    return out


def accel_dist_prop(links):
    out = np.ascontiguousarray(links, dtype=np.int32).copy()
    X, Y, Z = out.shape
    lib().og_accel_dist_prop(_ptr(out), X, Y, Z)
    return out


def dilate(mask):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.zeros_like(m)
    X, Y, Z = m.shape
    lib().og_dilate(_ptr(m), X, Y, Z, _ptr(out))
    return out.astype(bool)


def grid_weight_render(data, cam, step_size, stop_thresh, last_sample_opaque, offset, scaling,
                       grid_weight=None):
    data = _f32(data)
    X, Y, Z = data.shape
    out = np.zeros_like(data) if grid_weight is None else _f32(grid_weight).copy()
    offset, scaling = _f32(offset), _f32(scaling)
    lib().og_grid_weight_render(_ptr(data), X, Y, Z, C.byref(cam), C.c_float(step_size),
                                C.c_float(stop_thresh), int(last_sample_opaque), _ptr(offset),
                                _ptr(scaling), _ptr(out))
    return out


def sample_grid(grid, points, want_colors=True):
    points = _f32(points)
    P = points.shape[0]
    od = np.zeros((P, 1), np.float32)
    osh = np.zeros((P, grid.sh.shape[1]), np.float32) if want_colors else None
    g = grid.c()
    lib().og_sample_grid(C.byref(g), _ptr(points), C.c_int64(P), _ptr(od), _ptr(osh))
    return od, osh


def sample_grid_backward(grid, points, grad_out_density, grad_out_sh=None):
    points = _f32(points)
    god = _f32(grad_out_density).reshape(-1)
    gosh = None if grad_out_sh is None else _f32(grad_out_sh)
    gd = np.zeros_like(grid.density)
    gsh = np.zeros_like(grid.sh) if gosh is not None else None
    g = grid.c()
    lib().og_sample_grid_backward(C.byref(g), _ptr(points), C.c_int64(points.shape[0]),
                                  _ptr(god), _ptr(gosh), _ptr(gd), _ptr(gsh))
    return gd, gsh


def tv(links, data, start_dim, end_dim, ignore_edge):
    links = np.ascontiguousarray(links, np.int32)
    data = _f32(data)
    X, Y, Z = links.shape
    return float(lib().og_tv(_ptr(links), X, Y, Z, _ptr(data), data.shape[1], start_dim,
                             end_dim, int(ignore_edge)))


def tv_grad(links, data, start_dim, end_dim, scale, ignore_edge, grad=None):
    links = np.ascontiguousarray(links, np.int32)
    data = _f32(data)
    X, Y, Z = links.shape
    out = np.zeros_like(data) if grad is None else _f32(grad).copy()
    lib().og_tv_grad(_ptr(links), X, Y, Z, _ptr(data), data.shape[1], start_dim, end_dim,
                     C.c_float(scale), int(ignore_edge), _ptr(out))
    return out


def tv_grad_sparse(links, data, rand_cells, start_dim, end_dim, scale, ignore_edge,
                   ignore_last_z, grad=None, mask=None):
    links = np.ascontiguousarray(links, np.int32)
    data = _f32(data)
    cells = np.ascontiguousarray(rand_cells, np.int32)
    X, Y, Z = links.shape
    out = np.zeros_like(data) if grad is None else _f32(grad).copy()
    m = np.zeros(data.shape[0], np.uint8) if mask is None else np.ascontiguousarray(mask, np.uint8).copy()
    lib().og_tv_grad_sparse(_ptr(links), X, Y, Z, _ptr(data), data.shape[1], _ptr(cells),
                            C.c_int64(cells.shape[0]), _ptr(m), start_dim, end_dim,
                            C.c_float(scale), int(ignore_edge), int(ignore_last_z), _ptr(out))
    return out, m.astype(bool)


def msi_tv_grad_sparse(bg_links, msi, rand_cells, scale, scale_last, grad=None, mask=None):
    bg_links = np.ascontiguousarray(bg_links, np.int32)
    msi = _f32(msi)
    cells = np.ascontiguousarray(rand_cells, np.int32)
    out = np.zeros_like(msi) if grad is None else _f32(grad).copy()
    m = np.zeros(msi.shape[:2], np.uint8) if mask is None else np.ascontiguousarray(mask, np.uint8).copy()
    lib().og_msi_tv_grad_sparse(_ptr(bg_links), bg_links.shape[1], _ptr(msi), msi.shape[1],
                                msi.shape[2], _ptr(cells), C.c_int64(cells.shape[0]), _ptr(m),
                                C.c_float(scale), C.c_float(scale_last), _ptr(out))
    return out, m.astype(bool)


def lumisphere_tv_grad_sparse(links, sh, rand_cells, basis_fn, basis_fn_u, scale, dir_factor,
                              grad=None, mask=None):
    links = np.ascontiguousarray(links, np.int32)
    sh = _f32(sh)
    cells = np.ascontiguousarray(rand_cells, np.int32)
    fn, fnu = _f32(basis_fn).reshape(-1), _f32(basis_fn_u).reshape(-1)
    out = np.zeros_like(sh) if grad is None else _f32(grad).copy()
    m = np.zeros(sh.shape[0], np.uint8) if mask is None else np.ascontiguousarray(mask, np.uint8).copy()
    X, Y, Z = links.shape
    lib().og_lumisphere_tv_grad_sparse(_ptr(links), X, Y, Z, _ptr(sh), sh.shape[1] // 3, _ptr(cells),
                                       C.c_int64(cells.shape[0]), _ptr(fn), _ptr(fnu), C.c_float(scale),
                                       C.c_float(dir_factor), _ptr(m), _ptr(out))
    return out, m.astype(bool)


def _indexer(indexer):
    if indexer is None:
        return 0, None, None, 0
    ind = np.asarray(indexer)
    if ind.dtype == np.bool_:
        m = np.ascontiguousarray(ind, np.uint8)
        return 1, m, None, 0
    idx = np.ascontiguousarray(ind, np.int64)
    return 2, None, idx, idx.shape[0]


def rmsprop_step(data, rms, grad, indexer, beta, lr, eps, minval, lr_last):
    """Returns updated copies (data, rms, grad)."""
    data, rms, grad = _f32(data).copy(), _f32(rms).copy(), _f32(grad).copy()
    kind, m, idx, n = _indexer(indexer)
    lib().og_rmsprop_step(_ptr(data), _ptr(rms), _ptr(grad), C.c_int64(data.shape[0]),
                          data.shape[1], kind, _ptr(m), _ptr(idx), C.c_int64(n), C.c_float(beta),
                          C.c_float(lr), C.c_float(eps), C.c_float(minval), C.c_float(lr_last))
    return data, rms, grad


def sgd_step(data, grad, indexer, lr, lr_last):
    data, grad = _f32(data).copy(), _f32(grad).copy()
    kind, m, idx, n = _indexer(indexer)
    lib().og_sgd_step(_ptr(data), _ptr(grad), C.c_int64(data.shape[0]), data.shape[1], kind,
                      _ptr(m), _ptr(idx), C.c_int64(n), C.c_float(lr), C.c_float(lr_last))
    return data, grad
