"""Drop-in replacement for the reference's pybind module ``svox2.csrc``.

Same attribute names, argument order and meaning as ``svox2/csrc/svox2.cpp:74-172`` of
sxyu/svox2: 5 spec classes with read/write attributes and the operator functions, looked up
by name at call time by the Python layer (``svox2.py:239,252,1125,1158-1161``).  Underneath,
each function adapts torch tensors to raw pointers and calls the C-ABI of
``libsvox2_b200.so`` (``include/svox2_b200.h``) on torch's CURRENT stream (the reference
uses the legacy default stream).  Argument violations and CUDA errors raise ``RuntimeError``
(the reference raises for the former and only prints the latter).

The CUDA sources of the library live next to this file (``*.cu``, ``common.cuh``).
"""
import ctypes as C

import torch

from .. import _lib
from .._lib import CameraT, GradsT, GridT, RenderOptsT, Svox2Error, check

BASIS_TYPE_SH = 1

__all__ = [
    "SparseGridSpec", "CameraSpec", "RaysSpec", "RenderOptions", "GridOutputGrads",
    "sample_grid", "sample_grid_backward", "volume_render_cuvol", "volume_render_cuvol_image",
    "volume_render_cuvol_backward", "volume_render_cuvol_fused", "volume_render_expected_term",
    "volume_render_sigma_thresh", "volume_render_nvol", "volume_render_nvol_backward",
    "volume_render_nvol_fused", "volume_render_svox1", "volume_render_svox1_backward",
    "volume_render_svox1_fused", "dilate", "accel_dist_prop", "grid_weight_render", "tv",
    "tv_grad", "tv_grad_sparse", "tv_grad_sparse_run", "msi_tv_grad_sparse", "lumisphere_tv_grad_sparse",
    "rmsprop_step", "sgd_step",
]


# ----------------------------------------------------------------------------------------
# spec classes (data_spec.hpp:18-135): default-constructible, plain read/write attributes;
# an unset tensor attribute (None) means "absent" (an undefined torch::Tensor upstream).
# ----------------------------------------------------------------------------------------
class SparseGridSpec:
    def __init__(self):
        self.density_data = None
        self.sh_data = None
        self.links = None
        self._offset = None
        self._scaling = None
        self.basis_dim = 0
        self.basis_type = BASIS_TYPE_SH
        self.basis_data = None
        self.background_links = None
        self.background_data = None


class CameraSpec:
    def __init__(self):
        self.c2w = None
        self.fx = self.fy = self.cx = self.cy = 0.0
        self.width = self.height = 0
        self.ndc_coeffx = self.ndc_coeffy = -1.0


class RaysSpec:
    def __init__(self):
        self.origins = None
        self.dirs = None


class RenderOptions:
    def __init__(self):
        self.background_brightness = 1.0
        self.step_size = 0.5
        self.sigma_thresh = 1e-10
        self.stop_thresh = 1e-7
        self.near_clip = 0.0
        self.use_spheric_clip = False
        self.last_sample_opaque = False


class GridOutputGrads:
    def __init__(self):
        self.grad_density_out = None
        self.grad_sh_out = None
        self.grad_basis_out = None
        self.grad_background_out = None
        self.mask_out = None
        self.mask_background_out = None


# ----------------------------------------------------------------------------------------
# checks (util.hpp CHECK_INPUT / data_spec.hpp check()) and adapters
# ----------------------------------------------------------------------------------------
def _fail(msg):
    raise RuntimeError(msg)


def _defined(t):
    return t is not None and isinstance(t, torch.Tensor)


def _check_cuda(t, name, dtype=None, ndim=None):
    if not _defined(t):
        _fail(f"{name} must be a tensor")
    if not t.is_cuda:
        _fail(f"{name} must be a CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        _fail(f"{name} must have dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        _fail(f"{name} must have {ndim} dimensions, got {t.dim()}")


def _check_input(t, name, dtype=None, ndim=None):
    _check_cuda(t, name, dtype, ndim)
    if not t.is_contiguous():
        _fail(f"{name} must be contiguous")


def _check_rows(t, name):
    """2-D float32 CUDA tensor whose rows are contiguous (row stride may be padded)."""
    _check_cuda(t, name, torch.float32, 2)
    if t.size(0) > 1 and t.size(1) > 1 and t.stride(1) != 1:
        _fail(f"{name} must be contiguous")
    if t.size(0) > 1 and t.stride(0) < t.size(1):
        _fail(f"{name} must be contiguous")


def _row_stride(t):
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _host3(t, name):
    if not _defined(t):
        _fail(f"{name} must be a tensor")
    if t.is_cuda:
        _fail(f"{name} must be a CPU tensor")
    v = t.detach().to(torch.float32).reshape(-1).tolist()
    if len(v) != 3:
        _fail(f"{name} must have 3 elements")
    return v


def _grid_struct(spec):
    """SparseGridSpec.check() (data_spec.hpp:32-51) + PackedSparseGridSpec."""
    _check_rows(spec.density_data, "density_data")
    _check_rows(spec.sh_data, "sh_data")
    _check_input(spec.links, "links", torch.int32, 3)
    if int(spec.basis_type) != BASIS_TYPE_SH:
        _fail("svox2_b200 supports basis_type == BASIS_TYPE_SH only (learned bases are out of scope)")
    g = GridT()
    g.density_data = spec.density_data.data_ptr()
    g.sh_data = spec.sh_data.data_ptr()
    g.links = spec.links.data_ptr()
    g.size[:] = list(spec.links.shape)
    g.basis_dim = int(spec.basis_dim)
    if spec.sh_data.size(1) != 3 * g.basis_dim:
        _fail(f"sh_data must have 3*basis_dim = {3 * g.basis_dim} channels, got {spec.sh_data.size(1)}")
    if spec.density_data.size(0) != spec.sh_data.size(0):
        _fail("density_data and sh_data must have the same number of rows")
    g.n_rows = spec.sh_data.size(0)
    g.sh_stride = _row_stride(spec.sh_data)
    g.density_stride = _row_stride(spec.density_data)
    g.offset[:] = _host3(spec._offset, "_offset")
    g.scaling[:] = _host3(spec._scaling, "_scaling")
    bl, bd = spec.background_links, spec.background_data
    if _defined(bl) and _defined(bd) and bl.numel() > 0 and bd.numel() > 0:
        _check_input(bl, "background_links", torch.int32, 2)
        _check_input(bd, "background_data", torch.float32, 3)
        if bd.size(2) != 4:
            _fail("background_data must have 4 channels")
        g.background_links = bl.data_ptr()
        g.background_data = bd.data_ptr()
        g.background_reso = bl.size(1)
        g.background_nlayers = bd.size(1)
        g.background_rows = bd.size(0)
    return g


def _opt_struct(opt):
    return RenderOptsT(float(opt.background_brightness), float(opt.step_size),
                       float(opt.sigma_thresh), float(opt.stop_thresh), float(opt.near_clip),
                       int(bool(opt.use_spheric_clip)), int(bool(opt.last_sample_opaque)))


def _rays(rays):
    """RaysSpec.check() (data_spec.hpp:107-112)."""
    _check_input(rays.origins, "origins", torch.float32, 2)
    _check_input(rays.dirs, "dirs", torch.float32, 2)
    if rays.origins.size(1) != 3 or rays.dirs.shape != rays.origins.shape:
        _fail("origins and dirs must both be [Q, 3]")
    return rays.origins, rays.dirs


def _cam_struct(cam):
    """CameraSpec.check() (data_spec.hpp:96-101); c2w stays on the device."""
    _check_input(cam.c2w, "c2w", torch.float32, 2)
    if cam.c2w.size(1) != 4 or cam.c2w.size(0) < 3:
        _fail("c2w must be [3|4, 4]")
    return CameraT(cam.c2w.data_ptr(), float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy),
                   int(cam.width), int(cam.height), float(cam.ndc_coeffx), float(cam.ndc_coeffy))


def _grads_struct(grads, gspec):
    """GridOutputGrads.check() (data_spec.hpp:69-80)."""
    gr = GradsT()
    keep = []
    if _defined(grads.grad_density_out):
        _check_rows(grads.grad_density_out, "grad_density_out")
        if grads.grad_density_out.shape != gspec.density_data.shape:
            _fail("grad_density_out must have the shape of density_data")
        gr.grad_density_out = grads.grad_density_out.data_ptr()
        gr.grad_density_stride = _row_stride(grads.grad_density_out)
    if _defined(grads.grad_sh_out):
        _check_rows(grads.grad_sh_out, "grad_sh_out")
        if grads.grad_sh_out.shape != gspec.sh_data.shape:
            _fail("grad_sh_out must have the shape of sh_data")
        gr.grad_sh_out = grads.grad_sh_out.data_ptr()
        gr.grad_sh_stride = _row_stride(grads.grad_sh_out)
    if _defined(grads.grad_background_out) and grads.grad_background_out.numel() > 0:
        _check_input(grads.grad_background_out, "grad_background_out", torch.float32)
        gr.grad_background_out = grads.grad_background_out.data_ptr()
    if _defined(grads.mask_out) and grads.mask_out.numel() > 0:
        _check_input(grads.mask_out, "mask_out", torch.bool)
        gr.mask_out = grads.mask_out.data_ptr()
    if _defined(grads.mask_background_out) and grads.mask_background_out.numel() > 0:
        _check_input(grads.mask_background_out, "mask_background_out", torch.bool)
        gr.mask_background_out = grads.mask_background_out.data_ptr()
    return gr, keep


def _workspace(n_rays, device):
    nbytes = _lib.lib().svox2_render_workspace_bytes(n_rays)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws, nbytes


# ----------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------
def sample_grid(grid, points, want_colors):
    """svox2.cpp:11-13 -> svox2_sample_grid."""
    g = _grid_struct(grid)
    _check_input(points, "points", torch.float32, 2)
    P = points.size(0)
    dev = points.device
    out_density = torch.empty((P, grid.density_data.size(1)), dtype=torch.float32, device=dev)
    out_sh = torch.empty((P if want_colors else 0, grid.sh_data.size(1)), dtype=torch.float32,
                         device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().svox2_sample_grid(C.byref(g), points.data_ptr(), P, out_density.data_ptr(),
                                           out_sh.data_ptr() if want_colors else None, _stream()))
    return out_density, out_sh


def sample_grid_backward(grid, points, grad_out_density, grad_out_sh, grad_density_out,
                         grad_sh_out, want_colors):
    """svox2.cpp:14-17 -> svox2_sample_grid_backward."""
    g = _grid_struct(grid)
    _check_input(points, "points", torch.float32, 2)
    _check_input(grad_out_density, "grad_out_density", torch.float32, 2)
    _check_input(grad_out_sh, "grad_out_sh", torch.float32, 2)
    _check_rows(grad_density_out, "grad_density_out")
    _check_rows(grad_sh_out, "grad_sh_out")
    with torch.cuda.device(points.device):
        check(_lib.lib().svox2_sample_grid_backward(
            C.byref(g), points.data_ptr(), points.size(0), grad_out_density.data_ptr(),
            grad_out_sh.data_ptr() if want_colors else None, grad_density_out.data_ptr(),
            _row_stride(grad_density_out), grad_sh_out.data_ptr() if want_colors else None,
            _row_stride(grad_sh_out), _stream()))


def volume_render_cuvol(grid, rays, opt, _counters=None):
    """svox2.cpp:18 -> svox2_volume_render_cuvol.  Returns [Q,3]."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    Q = origins.size(0)
    out = torch.empty_like(origins)
    o = _opt_struct(opt)
    with torch.cuda.device(origins.device):
        ws, nbytes = _workspace(Q, origins.device)
        check(_lib.lib().svox2_volume_render_cuvol(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), Q, C.byref(o), out.data_ptr(),
            ws.data_ptr(), nbytes, _ptr(_counters), _stream()))
    return out


def volume_render_cuvol_image(grid, cam, opt, _rows=None, _out=None):
    """svox2.cpp:19-20 -> svox2_volume_render_cuvol_image.  Returns [H,W,3].

    ``_rows=(begin, end)`` (extension) renders only those image rows into ``_out``."""
    g = _grid_struct(grid)
    c = _cam_struct(cam)
    o = _opt_struct(opt)
    dev = grid.sh_data.device
    H, W = int(cam.height), int(cam.width)
    out = _out if _out is not None else torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    r0, r1 = (0, H) if _rows is None else _rows
    with torch.cuda.device(dev):
        ws, nbytes = _workspace((r1 - r0) * W, dev)
        check(_lib.lib().svox2_volume_render_cuvol_image(
            C.byref(g), C.byref(c), C.byref(o), r0, r1, out.data_ptr(), ws.data_ptr(), nbytes,
            _stream()))
    return out


def volume_render_cuvol_backward(grid, rays, opt, grad_out, color_cache, grads):
    """svox2.cpp:21-23 -> svox2_volume_render_cuvol_backward."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    _check_input(grad_out, "grad_out", torch.float32, 2)
    _check_input(color_cache, "color_cache", torch.float32, 2)
    gr, _ = _grads_struct(grads, grid)
    o = _opt_struct(opt)
    Q = origins.size(0)
    with torch.cuda.device(origins.device):
        ws, nbytes = _workspace(Q, origins.device)
        check(_lib.lib().svox2_volume_render_cuvol_backward(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), Q, C.byref(o), grad_out.data_ptr(),
            color_cache.data_ptr(), C.byref(gr), ws.data_ptr(), nbytes, _stream()))


def volume_render_cuvol_fused(grid, rays, opt, rgb_gt, beta_loss, sparsity_loss, rgb_out, grads,
                              _norm_rays=None, _counters=None):
    """svox2.cpp:24-26 -> svox2_volume_render_cuvol_fused.

    ``_norm_rays`` (extension): global batch size when the batch is sharded over GPUs."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    _check_input(rgb_gt, "rgb_gt", torch.float32, 2)
    _check_input(rgb_out, "rgb_out", torch.float32, 2)
    gr, _ = _grads_struct(grads, grid)
    o = _opt_struct(opt)
    Q = origins.size(0)
    with torch.cuda.device(origins.device):
        # scratch for the sample queue: up to 512 active samples per ray on average (never more
        # than the lattice steps along the grid diagonal), capped at 1 GiB; overflow only costs speed
        slots = 0
        if Q > 0:
            diag = sum(float(s) ** 2 for s in grid.links.shape) ** 0.5
            slots = int(diag / max(float(opt.step_size), 1e-3)) + 2
            slots = max(64, min(slots, 512, (1 << 30) // (96 * Q)))
        nbytes = _lib.lib().svox2_fused_workspace_bytes(Q, slots)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=origins.device)
        check(_lib.lib().svox2_volume_render_cuvol_fused(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), Q, Q if _norm_rays is None else _norm_rays,
            C.byref(o), rgb_gt.data_ptr(), float(beta_loss), float(sparsity_loss), rgb_out.data_ptr(),
            C.byref(gr), ws.data_ptr(), nbytes, _ptr(_counters), _stream()))


def volume_render_expected_term(grid, rays, opt):
    """svox2.cpp:27-28 -> svox2_volume_render_expected_term.  Returns [Q]."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    o = _opt_struct(opt)
    out = torch.empty((origins.size(0),), dtype=torch.float32, device=origins.device)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_volume_render_expected_term(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o),
            out.data_ptr(), _stream()))
    return out


def volume_render_sigma_thresh(grid, rays, opt, sigma_thresh):
    """svox2.cpp:29-30 -> svox2_volume_render_sigma_thresh.  Returns [Q]."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    o = _opt_struct(opt)
    out = torch.empty((origins.size(0),), dtype=torch.float32, device=origins.device)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_volume_render_sigma_thresh(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o),
            float(sigma_thresh), out.data_ptr(), _stream()))
    return out


def _out_of_scope(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"{name}: the svox1 ablation backend is outside the B200 hot path (SURVEY.md 2.1); "
            "use backend='cuvol' or 'nvol'.")
    fn.__name__ = name
    return fn


def volume_render_nvol(grid, rays, opt):
    """svox2.cpp:80 -> svox2_volume_render_nvol (render_lerp_kernel_nvol.cu:406-426).  Returns [Q,3]."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    out = torch.empty_like(origins)
    o = _opt_struct(opt)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_volume_render_nvol(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o),
            out.data_ptr(), _stream()))
    return out


def volume_render_nvol_backward(grid, rays, opt, grad_out, color_cache, grads):
    """svox2.cpp:81-82 -> svox2_volume_render_nvol_backward (nvol.cu:428-459)."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    _check_input(grad_out, "grad_out", torch.float32, 2)
    _check_input(color_cache, "color_cache", torch.float32, 2)
    gr, _ = _grads_struct(grads, grid)
    o = _opt_struct(opt)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_volume_render_nvol_backward(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o),
            grad_out.data_ptr(), color_cache.data_ptr(), C.byref(gr), _stream()))


def volume_render_nvol_fused(grid, rays, opt, rgb_gt, beta_loss, sparsity_loss, rgb_out, grads,
                             _norm_rays=None, _counters=None):
    """svox2.cpp:83-85 -> svox2_volume_render_nvol_fused (nvol.cu:461-502); ``beta_loss`` is
    ignored exactly as upstream ("not supported")."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    _check_input(rgb_gt, "rgb_gt", torch.float32, 2)
    _check_input(rgb_out, "rgb_out", torch.float32, 2)
    gr, _ = _grads_struct(grads, grid)
    o = _opt_struct(opt)
    Q = origins.size(0)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_volume_render_nvol_fused(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), Q, Q if _norm_rays is None else _norm_rays,
            C.byref(o), rgb_gt.data_ptr(), float(sparsity_loss), rgb_out.data_ptr(), C.byref(gr),
            _stream()))


# Name-visible in the reference module (svox2.cpp:86-92); the svox1 ablation is not built here.
volume_render_svox1 = _out_of_scope("volume_render_svox1")
volume_render_svox1_backward = _out_of_scope("volume_render_svox1_backward")
volume_render_svox1_fused = _out_of_scope("volume_render_svox1_fused")


def lumisphere_tv_grad_sparse(grid, rand_cells, sphfunc_val, sphfunc_val_u, scale, ndc_coeffx,
                              ndc_coeffy, dir_factor, grads):
    """svox2.cpp:50-51 -> svox2_lumisphere_tv_grad_sparse (loss_kernel.cu:643-678).  Like upstream
    only the first ``basis_dim`` floats of ``sphfunc_val_u`` are read (it may be [cells, basis_dim])."""
    _grid_struct(grid)  # SparseGridSpec.check()
    _check_input(rand_cells, "rand_cells", torch.int32)
    _check_input(sphfunc_val, "basis_fn", torch.float32, 1)
    _check_input(sphfunc_val_u, "basis_fn_u", torch.float32)
    B = int(grid.basis_dim)
    if sphfunc_val.numel() < B or sphfunc_val_u.numel() < B:
        _fail("basis_fn / basis_fn_u must hold basis_dim values")
    if not _defined(grads.grad_sh_out):
        _fail("grad_sh_out is required")
    _check_rows(grads.grad_sh_out, "grad_sh_out")
    if grads.grad_sh_out.shape != grid.sh_data.shape:
        _fail("grad_sh_out must have the shape of sh_data")
    mask = grads.mask_out
    mask_ptr = None
    if _defined(mask) and mask.dim() > 0 and mask.numel() > 0:
        _check_input(mask, "mask_out", torch.bool)
        mask_ptr = mask.data_ptr()
    X, Y, Z = grid.links.shape
    with torch.cuda.device(grid.sh_data.device):
        check(_lib.lib().svox2_lumisphere_tv_grad_sparse(
            grid.links.data_ptr(), X, Y, Z, grid.sh_data.data_ptr(), _row_stride(grid.sh_data), B,
            rand_cells.data_ptr(), rand_cells.size(0), sphfunc_val.data_ptr(), sphfunc_val_u.data_ptr(),
            float(scale), float(ndc_coeffx), float(ndc_coeffy), float(dir_factor), mask_ptr,
            grads.grad_sh_out.data_ptr(), _row_stride(grads.grad_sh_out), _stream()))


def dilate(grid):
    """svox2.cpp:60 -> svox2_dilate.  bool[X,Y,Z] -> new bool[X,Y,Z]."""
    _check_input(grid, "grid", None, 3)
    if grid.is_floating_point():
        _fail("grid must not be floating point")
    if grid.dtype != torch.bool:
        _fail("grid must be a bool tensor")
    out = torch.empty_like(grid)
    X, Y, Z = grid.shape
    with torch.cuda.device(grid.device):
        check(_lib.lib().svox2_dilate(grid.data_ptr(), X, Y, Z, out.data_ptr(), _stream()))
    return out


def accel_dist_prop(grid):
    """svox2.cpp:61 -> svox2_accel_dist_prop.  In place on the links tensor."""
    _check_input(grid, "grid", torch.int32, 3)
    X, Y, Z = grid.shape
    with torch.cuda.device(grid.device):
        nbytes = _lib.lib().svox2_accel_workspace_bytes(X, Y, Z)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=grid.device)
        check(_lib.lib().svox2_accel_dist_prop(grid.data_ptr(), X, Y, Z, ws.data_ptr(), nbytes,
                                               _stream()))


def grid_weight_render(data, cam, step_size, stop_thresh, last_sample_opaque, offset, scaling,
                       grid_weight_out):
    """svox2.cpp:62-65 -> svox2_grid_weight_render."""
    _check_input(data, "data", torch.float32, 3)
    _check_input(offset, "offset", torch.float32)
    _check_input(scaling, "scaling", torch.float32)
    _check_input(grid_weight_out, "grid_weight_out", torch.float32, 3)
    c = _cam_struct(cam)
    X, Y, Z = data.shape
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_grid_weight_render(
            data.data_ptr(), X, Y, Z, C.byref(c), float(step_size), float(stop_thresh),
            int(bool(last_sample_opaque)), offset.data_ptr(), scaling.data_ptr(),
            grid_weight_out.data_ptr(), _stream()))


def _tv_common(links, data):
    _check_input(links, "links", torch.int32, 3)
    _check_rows(data, "data")
    return links.shape


def tv(links, data, start_dim, end_dim, use_logalpha, logalpha_delta, ignore_edge, ndc_coeffx,
       ndc_coeffy):
    """svox2.cpp:32-35 -> svox2_tv.  Returns a 0-dim tensor."""
    X, Y, Z = _tv_common(links, data)
    out = torch.empty((), dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_tv(links.data_ptr(), X, Y, Z, data.data_ptr(), _row_stride(data),
                                  int(start_dim), int(end_dim), int(bool(ignore_edge)),
                                  float(ndc_coeffx), float(ndc_coeffy), out.data_ptr(), _stream()))
    return out


def tv_grad(links, data, start_dim, end_dim, scale, use_logalpha, logalpha_delta, ignore_edge,
            ndc_coeffx, ndc_coeffy, grad_data):
    """svox2.cpp:36-40 -> svox2_tv_grad."""
    X, Y, Z = _tv_common(links, data)
    _check_rows(grad_data, "grad_data")
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_tv_grad(links.data_ptr(), X, Y, Z, data.data_ptr(), _row_stride(data),
                                       int(start_dim), int(end_dim), float(scale),
                                       int(bool(ignore_edge)), float(ndc_coeffx), float(ndc_coeffy),
                                       grad_data.data_ptr(), _row_stride(grad_data), _stream()))


def tv_grad_sparse(links, data, rand_cells, mask_out, start_dim, end_dim, scale, use_logalpha,
                   logalpha_delta, ignore_edge, ignore_last_z, ndc_coeffx, ndc_coeffy, grad_data,
                   _scale_cells=0):
    """svox2.cpp:41-46 -> svox2_tv_grad_sparse.  mask_out with dim()==0 or no elements: no mask."""
    X, Y, Z = _tv_common(links, data)
    _check_rows(grad_data, "grad_data")
    _check_input(rand_cells, "rand_cells", torch.int32)
    _check_input(mask_out, "mask_out")
    mask_ptr = mask_out.data_ptr() if (mask_out.dim() > 0 and mask_out.numel() > 0) else None
    if mask_ptr is not None and mask_out.dtype != torch.bool:
        _fail("mask_out must be a bool tensor")
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_tv_grad_sparse(
            links.data_ptr(), X, Y, Z, data.data_ptr(), _row_stride(data), rand_cells.data_ptr(),
            rand_cells.size(0), int(_scale_cells), mask_ptr, int(start_dim), int(end_dim),
            float(scale), int(bool(ignore_edge)), int(bool(ignore_last_z)), float(ndc_coeffx),
            float(ndc_coeffy), grad_data.data_ptr(), _row_stride(grad_data), _stream()))


def tv_grad_sparse_run(links, data, cell_start, n_cells, mask_out, start_dim, end_dim, scale,
                       ignore_edge, ignore_last_z, grad_data, _scale_cells=0):
    """Extension: tv_grad_sparse over the contiguous wrapped cell run [cell_start, cell_start+n_cells)
    without an index tensor (what inplace_tv_grad(contiguous=True) needs every step)."""
    X, Y, Z = _tv_common(links, data)
    _check_rows(grad_data, "grad_data")
    _check_input(mask_out, "mask_out")
    mask_ptr = mask_out.data_ptr() if (mask_out.dim() > 0 and mask_out.numel() > 0) else None
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_tv_grad_sparse_run(
            links.data_ptr(), X, Y, Z, data.data_ptr(), _row_stride(data), int(cell_start),
            int(n_cells), int(_scale_cells), mask_ptr, int(start_dim), int(end_dim), float(scale),
            int(bool(ignore_edge)), int(bool(ignore_last_z)), grad_data.data_ptr(),
            _row_stride(grad_data), _stream()))


def msi_tv_grad_sparse(links, msi, rand_cells, mask_out, scale, scale_last, grad_msi,
                       _scale_cells=0):
    """svox2.cpp:47-49 -> svox2_msi_tv_grad_sparse."""
    _check_input(links, "links", torch.int32, 2)
    _check_input(msi, "msi", torch.float32, 3)
    _check_input(grad_msi, "grad_msi", torch.float32, 3)
    _check_input(rand_cells, "rand_cells", torch.int32)
    _check_input(mask_out, "mask_out")
    mask_ptr = mask_out.data_ptr() if mask_out.numel() > 0 else None
    with torch.cuda.device(msi.device):
        check(_lib.lib().svox2_msi_tv_grad_sparse(
            links.data_ptr(), links.size(1), msi.data_ptr(), msi.size(1), msi.size(2),
            rand_cells.data_ptr(), rand_cells.size(0), int(_scale_cells), mask_ptr, float(scale),
            float(scale_last), grad_msi.data_ptr(), _stream()))


def _indexer_args(indexer, rows):
    """Indexer dispatch of optim_kernel.cu:175-215: 0-dim => dense, empty => skip,
    bool => mask, otherwise int64 indices."""
    _check_cuda(indexer, "indexer")
    if indexer.dim() == 0:
        return 0, None, None, 0
    if indexer.size(0) == 0:
        return None
    if not indexer.is_contiguous():
        _fail("indexer must be contiguous")
    if indexer.dtype == torch.bool:
        if indexer.numel() != rows:
            _fail("bool indexer must have one entry per row")
        return 1, indexer.data_ptr(), None, 0
    if indexer.dtype != torch.int64:
        _fail("indexer must be bool or int64")
    return 2, None, indexer.data_ptr(), indexer.size(0)


def rmsprop_step(data, rms, grad, indexer, beta, lr, epsilon, minval, lr_last):
    """svox2.cpp:52-55 -> svox2_rmsprop_step."""
    _check_rows(data, "data")
    _check_rows(rms, "rms")
    _check_rows(grad, "grad")
    if not (_row_stride(data) == _row_stride(rms) == _row_stride(grad)):
        _fail("data, rms and grad must share one layout")
    sel = _indexer_args(indexer, data.size(0))
    if sel is None:
        return
    kind, mask, index, n_index = sel
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_rmsprop_step(
            data.data_ptr(), rms.data_ptr(), grad.data_ptr(), data.size(0), data.size(1),
            _row_stride(data), kind, mask, index, n_index, float(beta), float(lr), float(epsilon),
            float(minval), float(lr_last), _stream()))


def sgd_step(data, grad, indexer, lr, lr_last):
    """svox2.cpp:56-58 -> svox2_sgd_step."""
    _check_rows(data, "data")
    _check_rows(grad, "grad")
    if _row_stride(data) != _row_stride(grad):
        _fail("data and grad must share one layout")
    sel = _indexer_args(indexer, data.size(0))
    if sel is None:
        return
    kind, mask, index, n_index = sel
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_sgd_step(
            data.data_ptr(), grad.data_ptr(), data.size(0), data.size(1), _row_stride(data), kind,
            mask, index, n_index, float(lr), float(lr_last), _stream()))


# ---- multi-GPU gradient exchange helpers (extension) -------------------------
def packed_row_floats(sh_channels: int) -> int:
    """Floats per packed exchange row: SH channels + 1 density slot, padded to a multiple of 4."""
    return (sh_channels + 1 + 3) // 4 * 4


def pack_grad_rows(grad_sh, grad_density, rows, packed, mask_a=None, mask_b=None):
    """packed[i] = [grad_sh[rows[i]], grad_density[rows[i]], mask_a[rows[i]], mask_b[rows[i]], 0...]
    (any pair of [N, C] / [N, 1] arrays in the padded layout, e.g. the parameters themselves)."""
    _check_rows(grad_sh, "grad_sh")
    _check_rows(grad_density, "grad_density")
    _check_rows(packed, "packed")
    _check_cuda(rows, "rows", torch.int64, 1)
    if not rows.is_contiguous() or packed.size(0) != rows.numel():
        _fail("rows must be contiguous and match packed.size(0)")
    with torch.cuda.device(grad_sh.device):
        check(_lib.lib().svox2_pack_grad_rows(
            grad_sh.data_ptr(), _row_stride(grad_sh), grad_sh.size(1), grad_density.data_ptr(),
            _row_stride(grad_density), rows.data_ptr(), rows.numel(), packed.data_ptr(),
            _row_stride(packed), _ptr(mask_a), _ptr(mask_b), _stream()))


def unpack_rows(packed, rows, dst_sh, dst_density, add, mask_a=None, mask_b=None):
    """dst[rows[i]] (+)= packed[i]; mask_a / mask_b |= the flags packed by pack_grad_rows."""
    _check_rows(packed, "packed")
    _check_rows(dst_sh, "dst_sh")
    _check_rows(dst_density, "dst_density")
    _check_cuda(rows, "rows", torch.int64, 1)
    if not rows.is_contiguous() or packed.size(0) != rows.numel():
        _fail("rows must be contiguous and match packed.size(0)")
    with torch.cuda.device(dst_sh.device):
        check(_lib.lib().svox2_unpack_rows(
            packed.data_ptr(), _row_stride(packed), rows.data_ptr(), rows.numel(), dst_sh.data_ptr(),
            _row_stride(dst_sh), dst_sh.size(1), dst_density.data_ptr(), _row_stride(dst_density),
            1 if add else 0, _ptr(mask_a), _ptr(mask_b), _stream()))


def slab_prefix(logT_all, dirs, rank, lt_in):
    """lt_in[q] = sum of logT_all[j, q] over the slabs ray q crosses before slab `rank`."""
    _check_input(logT_all, "logT_all", torch.float32, 2)
    _check_input(dirs, "dirs", torch.float32, 2)
    _check_input(lt_in, "lt_in", torch.float32, 1)
    with torch.cuda.device(dirs.device):
        check(_lib.lib().svox2_slab_prefix(logT_all.data_ptr(), dirs.data_ptr(), dirs.size(0),
                                           logT_all.size(0), int(rank), lt_in.data_ptr(), _stream()))


def slab_combine(part_all, logT_all, dirs, rgb_gt, rank, background_brightness, beta_loss, rgb, gout,
                 accum_start):
    """Total colour, dL/d(rgb) and the reference's accum at slab entry from the gathered partials."""
    _check_input(part_all, "part_all", torch.float32, 3)
    _check_input(logT_all, "logT_all", torch.float32, 2)
    for t_, n_ in ((dirs, "dirs"), (rgb_gt, "rgb_gt"), (rgb, "rgb"), (gout, "gout")):
        _check_input(t_, n_, torch.float32, 2)
    _check_input(accum_start, "accum_start", torch.float32, 1)
    if part_all.size(2) != 4:
        _fail("part_all must be [slabs, Q, 4]")
    with torch.cuda.device(dirs.device):
        check(_lib.lib().svox2_slab_combine(
            part_all.data_ptr(), logT_all.data_ptr(), dirs.data_ptr(), rgb_gt.data_ptr(), dirs.size(0),
            part_all.size(0), int(rank), float(background_brightness), float(beta_loss),
            rgb.data_ptr(), gout.data_ptr(), accum_start.data_ptr(), _stream()))


def rmsprop_step_packed(sh_data, sh_rms, density_data, density_rms, rows, packed, mask_sh,
                        mask_density, beta, lr_sh, lr_density, epsilon, minval_sh=-1e9,
                        minval_density=-1e9, grad_sh=None, grad_density=None):
    """Masked RMSProp on density + SH of the listed rows, gradients from the packed buffer;
    zeroes those rows of grad_sh / grad_density (the local gradients that were packed)."""
    if (grad_sh is None) != (grad_density is None):
        _fail("grad_sh and grad_density go together")
    if grad_sh is not None:
        _check_rows(grad_sh, "grad_sh")
        _check_rows(grad_density, "grad_density")
    for t, n in ((sh_data, "sh_data"), (sh_rms, "sh_rms"), (density_data, "density_data"),
                 (density_rms, "density_rms"), (packed, "packed")):
        _check_rows(t, n)
    _check_cuda(rows, "rows", torch.int64, 1)
    for m, n in ((mask_sh, "mask_sh"), (mask_density, "mask_density")):
        _check_cuda(m, n)
        if m.dtype not in (torch.bool, torch.uint8) or m.numel() != sh_data.size(0) or not m.is_contiguous():
            _fail(f"{n} must be a contiguous bool/uint8 tensor with one entry per row")
    if _row_stride(sh_data) != _row_stride(sh_rms) or _row_stride(density_data) != _row_stride(density_rms):
        _fail("data and rms must share one layout")
    with torch.cuda.device(sh_data.device):
        check(_lib.lib().svox2_rmsprop_step_packed(
            sh_data.data_ptr(), sh_rms.data_ptr(), _row_stride(sh_data), sh_data.size(1),
            density_data.data_ptr(), density_rms.data_ptr(), _row_stride(density_data),
            rows.data_ptr(), rows.numel(), packed.data_ptr(), _row_stride(packed),
            mask_sh.data_ptr(), mask_density.data_ptr(), _ptr(grad_sh),
            0 if grad_sh is None else _row_stride(grad_sh), _ptr(grad_density),
            0 if grad_density is None else _row_stride(grad_density), float(beta), float(lr_sh),
            float(lr_density), float(epsilon), float(minval_sh), float(minval_density), _stream()))


# ---- fused train step on a compact gradient buffer (csrc/step.cu) ---------------------------
class TrainStepWorkspace:
    """Persistent scratch of ``train_step`` for one (grid rows, basis, batch size): the sample-record
    queue plus the touched-row bit set, slot numbering and compact gradient rows.  Zero-initialised
    once; every step leaves it in the state the next one expects."""

    def __init__(self, n_rows: int, basis_dim: int, n_rays: int, slots_per_ray: int, device):
        self.key = (int(n_rows), int(basis_dim), int(n_rays), int(slots_per_ray))
        self.nbytes = _lib.lib().svox2_train_step_workspace_bytes(*self.key)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            check(_lib.lib().svox2_train_step_init(self.buf.data_ptr(), self.nbytes, *self.key, _stream()))

    def status(self, want_masks: bool = False):
        """(touched rows, overflowed last step, overflowed steps before) and optionally the two sparse
        masks of the LAST step as bool[N] (host sync)."""
        dev = self.buf.device
        st = torch.zeros(3, dtype=torch.int32, device=dev)
        md = msh = None
        if want_masks:
            md = torch.zeros(self.key[0], dtype=torch.bool, device=dev)
            msh = torch.zeros(self.key[0], dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().svox2_train_step_masks(self.buf.data_ptr(), self.nbytes, *self.key, _ptr(md),
                                                    _ptr(msh), st.data_ptr(), _stream()))
        n, ov, sticky = st.tolist()
        return (n, ov, max(sticky, 0)), md, msh


def train_step(grid, density_rms, sh_rms, rays, opt, rgb_gt, rgb_out, workspace: TrainStepWorkspace, *,
               tv_density=(0, 0, 0.0), tv_sh=(0, 0, 0.0), tv_scale_cells=0, beta_loss=0.0,
               sparsity_loss=0.0, norm_rays=None, lr_density=0.0, lr_sh=0.0, rms_beta=0.95,
               epsilon=1e-8, mse_out=None, _counters=None):
    """Extension: svox2_train_step -- fused render + sparse TV (density, SH; contiguous cell runs
    ``(start, count, lambda)``) + RMSProp (density, SH) in one call; density_data / sh_data (through
    ``grid``) and the RMS arrays are updated in place, ``rgb_out`` receives the colours and
    ``mse_out`` (optional 0-dim float32 CUDA tensor) the batch's mean squared error."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    _check_input(rgb_gt, "rgb_gt", torch.float32, 2)
    _check_input(rgb_out, "rgb_out", torch.float32, 2)
    _check_rows(density_rms, "density_rms")
    _check_rows(sh_rms, "sh_rms")
    if _row_stride(sh_rms) != _row_stride(grid.sh_data) or _row_stride(density_rms) != _row_stride(grid.density_data):
        _fail("data and rms must share one layout")
    Q = origins.size(0)
    if workspace.key != (grid.sh_data.size(0), int(grid.basis_dim), Q, workspace.key[3]):
        _fail("train_step workspace was created for another grid / batch size")
    o = _opt_struct(opt)
    p = _lib.StepParamsT(int(tv_density[0]), int(tv_density[1]), float(tv_density[2]), int(tv_sh[0]),
                         int(tv_sh[1]), float(tv_sh[2]), int(tv_scale_cells), float(beta_loss),
                         float(sparsity_loss), int(Q if norm_rays is None else norm_rays), float(lr_density),
                         float(lr_sh), float(rms_beta), float(epsilon))
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_train_step(
            C.byref(g), density_rms.data_ptr(), sh_rms.data_ptr(), origins.data_ptr(), dirs.data_ptr(), Q,
            C.byref(o), rgb_gt.data_ptr(), rgb_out.data_ptr(), C.byref(p), workspace.buf.data_ptr(),
            workspace.nbytes, workspace.key[3], _ptr(mse_out), _ptr(_counters), _stream()))


# ---- slab mode (grid sharded over GPUs by x-range of cells; svox2_b200/slab.py) --------------
def slab_workspace(n_rays: int, slots_per_ray: int, device):
    """Sample-record scratch shared by slab_trace / slab_shade / slab_backward."""
    nbytes = _lib.lib().svox2_fused_workspace_bytes(n_rays, max(int(slots_per_ray), 64))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    ws[:256].zero_()        # queue header: counter, overflow flag, sticky overflow count
    return ws


def slab_overflowed(workspace) -> bool:
    """True if the last slab_trace, or any earlier one on this workspace since it was created,
    ran out of record space (host sync)."""
    hdr = workspace[:256].view(torch.int32)
    return bool((hdr[32] != 0).item() or (hdr[48] != 0).item())


def slab_trace(grid, rays, opt, x_lo, x_hi, logT_local, workspace):
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    o = _opt_struct(opt)
    _check_input(logT_local, "logT_local", torch.float32, 1)
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_slab_trace(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o), int(x_lo),
            int(x_hi), logT_local.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()))


def slab_shade(grid, n_rays, opt, lt_in, part, workspace):
    """part [Q,4] (zeroed by the caller): .xyz += colour of this slab's samples, .w = stop flag."""
    g = _grid_struct(grid)
    o = _opt_struct(opt)
    _check_input(lt_in, "lt_in", torch.float32, 1)
    _check_input(part, "part", torch.float32, 2)
    if part.size(1) != 4:
        _fail("part must be [Q, 4]")
    with torch.cuda.device(lt_in.device):
        check(_lib.lib().svox2_slab_shade(
            C.byref(g), int(n_rays), C.byref(o), lt_in.data_ptr(), part.data_ptr(),
            workspace.data_ptr(), workspace.numel(), _stream()))


def slab_backward(grid, n_rays, opt, lt_in, gout, accum_start, sparsity_loss, grads, workspace):
    g = _grid_struct(grid)
    o = _opt_struct(opt)
    gr, _ = _grads_struct(grads, grid)
    _check_input(lt_in, "lt_in", torch.float32, 1)
    _check_input(gout, "gout", torch.float32, 2)
    _check_input(accum_start, "accum_start", torch.float32, 1)
    with torch.cuda.device(lt_in.device):
        check(_lib.lib().svox2_slab_backward(
            C.byref(g), int(n_rays), C.byref(o), lt_in.data_ptr(), gout.data_ptr(),
            accum_start.data_ptr(), float(sparsity_loss), C.byref(gr), workspace.data_ptr(),
            workspace.numel(), _stream()))


# ---- ship mode (default multi-GPU step): rays sharded for the traversal, records shipped to the
# ---- slab that owns the sample's cell (svox2_b200/slab.py, include/svox2_b200.h) -----------------
class ShipWorkspace:
    """Geometry of the per-rank symmetric record workspace: ``peer_ptrs[d]`` is rank d's buffer
    mapped into this process, ``local`` the uint8 tensor over this rank's own buffer."""

    def __init__(self, n_rays_global: int, n_slabs: int, rank: int, region_cap: int):
        self.n_rays, self.n_slabs, self.rank, self.region_cap = int(n_rays_global), int(n_slabs), int(rank), int(region_cap)
        self.nbytes = int(_lib.lib().svox2_ship_workspace_bytes(self.n_rays, self.n_slabs, self.region_cap))
        self.local, self.peer_ptrs = None, None

    def bind(self, local: torch.Tensor, peer_ptrs):
        if local.dtype != torch.uint8 or local.numel() < self.nbytes or len(peer_ptrs) != self.n_slabs:
            _fail("ShipWorkspace.bind: local must be a uint8 tensor of >= nbytes bytes, one pointer per slab")
        self.local = local
        self.peer_ptrs = [int(p) for p in peer_ptrs]
        self._arr = (C.c_uint64 * self.n_slabs)(*self.peer_ptrs)
        local[:512].zero_()            # header, counts, counters (the rest is written before it is read)

    def ray_logT(self) -> torch.Tensor:
        """[Qg] forward log-transmittance of every ray of the global batch (written by its tracer)."""
        ptr = _lib.lib().svox2_ship_ray_logT(self.local.data_ptr(), self.nbytes, self.n_rays, self.n_slabs,
                                             self.region_cap)
        if not ptr:
            _fail("ShipWorkspace: bad geometry")
        off = int(ptr) - self.local.data_ptr()
        return self.local[off:off + 4 * self.n_rays].view(torch.float32)

    def overflowed(self) -> bool:
        hdr = self.local[:256].view(torch.int32)
        return bool((hdr[32] != 0).item() or (hdr[48] != 0).item())


def ship_trace(grid, rays, opt, ray_begin, ray_count, slab_bounds, ws: ShipWorkspace, _counters=None):
    """Traverse rays [ray_begin, ray_begin + ray_count) of the global batch and ship their sample
    records to the owning slabs (NVLink stores into the peers' workspaces)."""
    g = _grid_struct(grid)
    origins, dirs = _rays(rays)
    o = _opt_struct(opt)
    if origins.size(0) != ws.n_rays or len(slab_bounds) != ws.n_slabs + 1:
        _fail("ship_trace: rays must hold the global batch the workspace was sized for")
    b = (C.c_int32 * len(slab_bounds))(*[int(x) for x in slab_bounds])
    with torch.cuda.device(origins.device):
        check(_lib.lib().svox2_ship_trace(
            C.byref(g), origins.data_ptr(), dirs.data_ptr(), ws.n_rays, int(ray_begin), int(ray_count),
            C.byref(o), C.addressof(b), ws.n_slabs, ws.rank, C.addressof(ws._arr), ws.nbytes, ws.region_cap,
            _ptr(_counters), _stream()))


def ship_shade(grid, part, ws: ShipWorkspace):
    """part [Qg,4] (zeroed by the caller): .xyz += colour of the records this rank received."""
    g = _grid_struct(grid)
    _check_input(part, "part", torch.float32, 2)
    if part.size(1) != 4 or part.size(0) != ws.n_rays:
        _fail("part must be [Qg, 4]")
    with torch.cuda.device(part.device):
        check(_lib.lib().svox2_ship_shade(C.byref(g), ws.n_rays, ws.n_slabs, ws.region_cap, part.data_ptr(),
                                          ws.local.data_ptr(), ws.nbytes, _stream()))


def ship_combine(part_all, ray_logT, dirs, rgb_gt, rank, background_brightness, beta_loss, rgb, gout,
                 accum_start):
    """slab_combine with the ray's total log-transmittance taken from its tracer."""
    _check_input(part_all, "part_all", torch.float32, 3)
    _check_input(ray_logT, "ray_logT", torch.float32, 1)
    for t_, n_ in ((dirs, "dirs"), (rgb_gt, "rgb_gt"), (rgb, "rgb"), (gout, "gout")):
        _check_input(t_, n_, torch.float32, 2)
    _check_input(accum_start, "accum_start", torch.float32, 1)
    if part_all.size(2) != 4:
        _fail("part_all must be [slabs, Q, 4]")
    with torch.cuda.device(dirs.device):
        check(_lib.lib().svox2_ship_combine(
            part_all.data_ptr(), ray_logT.data_ptr(), dirs.data_ptr(), rgb_gt.data_ptr(), dirs.size(0),
            part_all.size(0), int(rank), float(background_brightness), float(beta_loss),
            rgb.data_ptr(), gout.data_ptr(), accum_start.data_ptr(), _stream()))


def ship_backward(grid, opt, gout, accum_start, sparsity_loss, grads, ws: ShipWorkspace):
    g = _grid_struct(grid)
    o = _opt_struct(opt)
    gr, _ = _grads_struct(grads, grid)
    _check_input(gout, "gout", torch.float32, 2)
    _check_input(accum_start, "accum_start", torch.float32, 1)
    with torch.cuda.device(gout.device):
        check(_lib.lib().svox2_ship_backward(
            C.byref(g), ws.n_rays, C.byref(o), ws.n_slabs, ws.region_cap, gout.data_ptr(),
            accum_start.data_ptr(), float(sparsity_loss), C.byref(gr), ws.local.data_ptr(), ws.nbytes,
            _stream()))


# ---- multi-GPU step on the compact gradient buffer (step.cu; driven by svox2_b200/slab.py) --------
class MgStep:
    """One x-slab rank's view of the multi-GPU step: the train-step workspace (compact gradient
    slots, sized for the GLOBAL batch), optionally the ship workspace, and the per-step parameters.
    The methods are the phases of include/svox2_b200.h `svox2_mg_step_*`."""

    def __init__(self, step_ws: TrainStepWorkspace, n_slabs: int, rank: int, x_lo: int, x_hi: int,
                 ship: "ShipWorkspace" = None, slab_bounds=None):
        self.ws, self.ship = step_ws, ship
        self.p = _lib.MgStepT()
        self.p.n_slabs, self.p.rank, self.p.x_lo, self.p.x_hi = int(n_slabs), int(rank), int(x_lo), int(x_hi)
        self.p.ship = 1 if ship is not None else 0
        self._bounds = None
        if slab_bounds is not None:
            self._bounds = (C.c_int32 * len(slab_bounds))(*[int(b) for b in slab_bounds])

    def set_step(self, tv_density_runs, tv_density_total, lambda_tv, tv_sh_runs, tv_sh_total, lambda_tv_sh,
                 sparsity_loss, lr_density, lr_sh, rms_beta, epsilon=1e-8):
        p = self.p
        for k in range(2):
            d = tv_density_runs[k] if k < len(tv_density_runs) else (0, 0)
            h = tv_sh_runs[k] if k < len(tv_sh_runs) else (0, 0)
            p.tv_density_start[k], p.tv_density_count[k] = int(d[0]), int(d[1])
            p.tv_sh_start[k], p.tv_sh_count[k] = int(h[0]), int(h[1])
        p.tv_density_scale, p.tv_sh_scale = float(lambda_tv), float(lambda_tv_sh)
        p.tv_density_scale_cells, p.tv_sh_scale_cells = int(tv_density_total), int(tv_sh_total)
        p.sparsity_loss = float(sparsity_loss)
        p.lr_density, p.lr_sh, p.rms_beta, p.epsilon = float(lr_density), float(lr_sh), float(rms_beta), float(epsilon)

    def _ship_args(self):
        if self.ship is None:
            return None, 0, 0
        return self.ship.local.data_ptr(), self.ship.nbytes, self.ship.region_cap

    def begin(self, grid, rays, opt, boundary_rows, logT_local=None):
        g = _grid_struct(grid)
        origins, dirs = _rays(rays)
        o = _opt_struct(opt)
        _check_cuda(boundary_rows, "boundary_rows", torch.int64, 1)
        if origins.size(0) != self.ws.key[2]:
            _fail("the step workspace was sized for another global batch")
        peers = C.addressof(self.ship._arr) if self.ship is not None else None
        nb, cap = (self.ship.nbytes, self.ship.region_cap) if self.ship is not None else (0, 0)
        with torch.cuda.device(origins.device):
            check(_lib.lib().svox2_mg_step_begin(
                C.byref(g), origins.data_ptr(), dirs.data_ptr(), origins.size(0), C.byref(o), C.byref(self.p),
                boundary_rows.data_ptr(), boundary_rows.numel(), self.ws.buf.data_ptr(), self.ws.nbytes,
                self.ws.key[3], _ptr(logT_local), None if self._bounds is None else C.addressof(self._bounds),
                peers, nb, cap, _stream()))

    def shade(self, grid, opt, lt_in, part):
        g = _grid_struct(grid)
        o = _opt_struct(opt)
        _check_input(part, "part", torch.float32, 2)
        sp, nb, cap = self._ship_args()
        with torch.cuda.device(part.device):
            check(_lib.lib().svox2_mg_step_shade(
                C.byref(g), self.ws.key[2], C.byref(o), C.byref(self.p), _ptr(lt_in), part.data_ptr(),
                self.ws.buf.data_ptr(), self.ws.nbytes, self.ws.key[3], sp, nb, cap, _stream()))

    def backward(self, grid, opt, lt_in, gout, accum_start):
        g = _grid_struct(grid)
        o = _opt_struct(opt)
        _check_input(gout, "gout", torch.float32, 2)
        _check_input(accum_start, "accum_start", torch.float32, 1)
        sp, nb, cap = self._ship_args()
        with torch.cuda.device(gout.device):
            check(_lib.lib().svox2_mg_step_backward(
                C.byref(g), self.ws.key[2], C.byref(o), C.byref(self.p), _ptr(lt_in), gout.data_ptr(),
                accum_start.data_ptr(), self.ws.buf.data_ptr(), self.ws.nbytes, self.ws.key[3], sp, nb, cap,
                _stream()))

    def slot_pack(self, rows, packed, clear=False):
        """packed[i] = [slot row of rows[i] | density flag | SH flag | 0, 0] (zeros if untouched)."""
        _check_cuda(rows, "rows", torch.int64, 1)
        _check_rows(packed, "packed")
        if packed.size(0) != rows.numel():
            _fail("one packet per row")
        with torch.cuda.device(packed.device):
            check(_lib.lib().svox2_mg_slot_pack(rows.data_ptr(), rows.numel(), packed.data_ptr(), _row_stride(packed),
                                                1 if clear else 0, self.ws.buf.data_ptr(), self.ws.nbytes,
                                                *self.ws.key, _stream()))

    def slot_unpack_add(self, packed, rows):
        _check_cuda(rows, "rows", torch.int64, 1)
        _check_rows(packed, "packed")
        if packed.size(0) != rows.numel():
            _fail("one packet per row")
        with torch.cuda.device(packed.device):
            check(_lib.lib().svox2_mg_slot_unpack_add(packed.data_ptr(), _row_stride(packed), rows.data_ptr(),
                                                      rows.numel(), self.ws.buf.data_ptr(), self.ws.nbytes,
                                                      *self.ws.key, _stream()))

    def finish(self, grid, density_rms, sh_rms, send=None, peer_density=None):
        g = _grid_struct(grid)
        _check_rows(density_rms, "density_rms")
        _check_rows(sh_rms, "sh_rms")
        arr = None
        if send is not None:
            _check_cuda(send, "send", torch.uint8, 1)
            arr = (C.c_uint64 * len(peer_density))(*[int(x) for x in peer_density])
        with torch.cuda.device(sh_rms.device):
            check(_lib.lib().svox2_mg_step_finish(
                C.byref(g), density_rms.data_ptr(), sh_rms.data_ptr(), self.ws.key[2], C.byref(self.p),
                _ptr(send), None if arr is None else C.addressof(arr), self.ws.buf.data_ptr(), self.ws.nbytes,
                self.ws.key[3], _stream()))


def rmsprop_step_publish(data, rms, grad, mask, send, peer_ptrs, self_rank, skip_peer, beta, lr, epsilon,
                         minval=-1e9):
    """Masked RMSProp on a replicated single-channel array + publication of the updated rows
    (send: uint8 per row, 0 local only / 1 every peer / 2 every peer but ``skip_peer``)."""
    for t_, n_ in ((data, "data"), (rms, "rms"), (grad, "grad")):
        _check_rows(t_, n_)
    if data.size(1) != 1 or not (_row_stride(data) == _row_stride(rms) == _row_stride(grad)):
        _fail("data, rms and grad must be [N,1] arrays of one layout")
    _check_cuda(mask, "mask", torch.bool, 1)
    _check_cuda(send, "send", torch.uint8, 1)
    if mask.numel() != data.size(0) or send.numel() != data.size(0):
        _fail("mask / send must have one entry per row")
    arr = (C.c_uint64 * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(data.device):
        check(_lib.lib().svox2_rmsprop_step_publish(
            data.data_ptr(), rms.data_ptr(), grad.data_ptr(), data.size(0), _row_stride(data), mask.data_ptr(),
            send.data_ptr(), C.addressof(arr), len(peer_ptrs), int(self_rank), int(skip_peer), float(beta),
            float(lr), float(epsilon), float(minval), _stream()))


def peer_store(peer_ptrs, dst_offset_bytes, src):
    """src -> (base + dst_offset_bytes) of every peer buffer (list of device pointers mapped into
    this process, e.g. torch symmetric memory ``buffer_ptrs``); one launch, NVLink P2P stores."""
    _check_input(src, "src")
    arr = (C.c_uint64 * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(src.device):
        check(_lib.lib().svox2_peer_store(C.addressof(arr), len(peer_ptrs), int(dst_offset_bytes),
                                          src.data_ptr(), src.numel() * src.element_size(), _stream()))
