"""ctypes binding of ``libsvox2_b200.so`` (the C-ABI declared in ``include/svox2_b200.h``).

There is NO fallback: if the CUDA library has not been built, or a call fails, this raises.
Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C svox2_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVOX2_B200_LIB selects an A/B build variant (tools/build_variants.sh); default: the in-tree build
LIB_PATH = os.environ.get("SVOX2_B200_LIB") or os.path.join(_HERE, "csrc", "libsvox2_b200.so")


class Svox2Error(RuntimeError):
    """Non-zero status from the C-ABI (the reference raises RuntimeError via TORCH_CHECK)."""


class GridT(C.Structure):
    _fields_ = [
        ("density_data", C.c_void_p),
        ("sh_data", C.c_void_p),
        ("links", C.c_void_p),
        ("size", C.c_int32 * 3),
        ("basis_dim", C.c_int32),
        ("n_rows", C.c_int64),
        ("sh_stride", C.c_int64),
        ("density_stride", C.c_int64),
        ("offset", C.c_float * 3),
        ("scaling", C.c_float * 3),
        ("background_links", C.c_void_p),
        ("background_data", C.c_void_p),
        ("background_reso", C.c_int32),
        ("background_nlayers", C.c_int32),
        ("background_rows", C.c_int64),
    ]


class RenderOptsT(C.Structure):
    _fields_ = [
        ("background_brightness", C.c_float),
        ("step_size", C.c_float),
        ("sigma_thresh", C.c_float),
        ("stop_thresh", C.c_float),
        ("near_clip", C.c_float),
        ("use_spheric_clip", C.c_int32),
        ("last_sample_opaque", C.c_int32),
    ]


class CameraT(C.Structure):
    _fields_ = [
        ("c2w", C.c_void_p),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32),
        ("ndc_coeffx", C.c_float), ("ndc_coeffy", C.c_float),
    ]


class GradsT(C.Structure):
    _fields_ = [
        ("grad_density_out", C.c_void_p),
        ("grad_sh_out", C.c_void_p),
        ("grad_background_out", C.c_void_p),
        ("mask_out", C.c_void_p),
        ("mask_background_out", C.c_void_p),
        ("grad_sh_stride", C.c_int64),
        ("grad_density_stride", C.c_int64),
    ]


class StepParamsT(C.Structure):
    _fields_ = [
        ("tv_density_start", C.c_int64), ("tv_density_count", C.c_int64), ("tv_density_scale", C.c_float),
        ("tv_sh_start", C.c_int64), ("tv_sh_count", C.c_int64), ("tv_sh_scale", C.c_float),
        ("tv_scale_cells", C.c_int64), ("beta_loss", C.c_float), ("sparsity_loss", C.c_float),
        ("norm_rays", C.c_int64), ("lr_density", C.c_float), ("lr_sh", C.c_float),
        ("rms_beta", C.c_float), ("epsilon", C.c_float),
    ]


class MgStepT(C.Structure):
    _fields_ = [
        ("n_slabs", C.c_int32), ("rank", C.c_int32), ("x_lo", C.c_int32), ("x_hi", C.c_int32),
        ("ship", C.c_int32),
        ("tv_density_start", C.c_int64 * 2), ("tv_density_count", C.c_int64 * 2),
        ("tv_sh_start", C.c_int64 * 2), ("tv_sh_count", C.c_int64 * 2),
        ("tv_density_scale", C.c_float), ("tv_sh_scale", C.c_float),
        ("tv_density_scale_cells", C.c_int64), ("tv_sh_scale_cells", C.c_int64),
        ("sparsity_loss", C.c_float),
        ("lr_density", C.c_float), ("lr_sh", C.c_float), ("rms_beta", C.c_float), ("epsilon", C.c_float),
    ]


_P = C.c_void_p
_I32, _I64, _F, _SZ = C.c_int32, C.c_int64, C.c_float, C.c_size_t
_G, _O, _CAM, _GR = C.POINTER(GridT), C.POINTER(RenderOptsT), C.POINTER(CameraT), C.POINTER(GradsT)

# name -> (restype, argtypes); mirrors include/svox2_b200.h one to one
SIGNATURES = {
    "svox2_last_error": (C.c_char_p, []),
    "svox2_version": (C.c_char_p, []),
    "svox2_cuda_available": (C.c_int, []),
    "svox2_render_workspace_bytes": (_SZ, [_I64]),
    "svox2_fused_workspace_bytes": (_SZ, [_I64, _I64]),
    "svox2_volume_render_cuvol": (C.c_int, [_G, _P, _P, _I64, _O, _P, _P, _SZ, _P, _P]),
    "svox2_volume_render_cuvol_image": (C.c_int, [_G, _CAM, _O, _I32, _I32, _P, _P, _SZ, _P]),
    "svox2_volume_render_cuvol_backward": (C.c_int, [_G, _P, _P, _I64, _O, _P, _P, _GR, _P, _SZ, _P]),
    "svox2_volume_render_cuvol_fused": (C.c_int, [_G, _P, _P, _I64, _I64, _O, _P, _F, _F, _P, _GR, _P,
                                                  _SZ, _P, _P]),
    "svox2_volume_render_expected_term": (C.c_int, [_G, _P, _P, _I64, _O, _P, _P]),
    "svox2_volume_render_sigma_thresh": (C.c_int, [_G, _P, _P, _I64, _O, _F, _P, _P]),
    "svox2_volume_render_nvol": (C.c_int, [_G, _P, _P, _I64, _O, _P, _P]),
    "svox2_volume_render_nvol_backward": (C.c_int, [_G, _P, _P, _I64, _O, _P, _P, _GR, _P]),
    "svox2_volume_render_nvol_fused": (C.c_int, [_G, _P, _P, _I64, _I64, _O, _P, _F, _P, _GR, _P]),
    "svox2_train_step_workspace_bytes": (_SZ, [_I64, _I32, _I64, _I64]),
    "svox2_train_step_init": (C.c_int, [_P, _SZ, _I64, _I32, _I64, _I64, _P]),
    "svox2_train_step": (C.c_int, [_G, _P, _P, _P, _P, _I64, _O, _P, _P, C.POINTER(StepParamsT), _P, _SZ,
                                   _I64, _P, _P, _P]),
    "svox2_train_step_masks": (C.c_int, [_P, _SZ, _I64, _I32, _I64, _I64, _P, _P, _P, _P]),
    "svox2_sample_grid": (C.c_int, [_G, _P, _I64, _P, _P, _P]),
    "svox2_sample_grid_backward": (C.c_int, [_G, _P, _I64, _P, _P, _P, _I64, _P, _I64, _P]),
    "svox2_dilate": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "svox2_accel_workspace_bytes": (_SZ, [_I32, _I32, _I32]),
    "svox2_accel_dist_prop": (C.c_int, [_P, _I32, _I32, _I32, _P, _SZ, _P]),
    "svox2_grid_weight_render": (C.c_int, [_P, _I32, _I32, _I32, _CAM, _F, _F, _I32, _P, _P, _P, _P]),
    "svox2_tv": (C.c_int, [_P, _I32, _I32, _I32, _P, _I64, _I32, _I32, _I32, _F, _F, _P, _P]),
    "svox2_tv_grad": (C.c_int, [_P, _I32, _I32, _I32, _P, _I64, _I32, _I32, _F, _I32, _F, _F, _P, _I64, _P]),
    "svox2_tv_grad_sparse": (C.c_int, [_P, _I32, _I32, _I32, _P, _I64, _P, _I64, _I64, _P, _I32, _I32,
                                       _F, _I32, _I32, _F, _F, _P, _I64, _P]),
    "svox2_tv_grad_sparse_run": (C.c_int, [_P, _I32, _I32, _I32, _P, _I64, _I64, _I64, _I64, _P, _I32,
                                           _I32, _F, _I32, _I32, _P, _I64, _P]),
    "svox2_msi_tv_grad_sparse": (C.c_int, [_P, _I32, _P, _I32, _I32, _P, _I64, _I64, _P, _F, _F, _P, _P]),
    "svox2_lumisphere_tv_grad_sparse": (C.c_int, [_P, _I32, _I32, _I32, _P, _I64, _I32, _P, _I64, _P, _P,
                                                  _F, _F, _F, _F, _P, _P, _I64, _P]),
    "svox2_rmsprop_step": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, _I32, _P, _P, _I64, _F, _F, _F, _F,
                                     _F, _P]),
    "svox2_sgd_step": (C.c_int, [_P, _P, _I64, _I32, _I64, _I32, _P, _P, _I64, _F, _F, _P]),
    "svox2_slab_trace": (C.c_int, [_G, _P, _P, _I64, _O, _I32, _I32, _P, _P, _SZ, _P]),
    "svox2_slab_shade": (C.c_int, [_G, _I64, _O, _P, _P, _P, _SZ, _P]),
    "svox2_slab_backward": (C.c_int, [_G, _I64, _O, _P, _P, _P, _F, _GR, _P, _SZ, _P]),
    "svox2_ship_workspace_bytes": (_SZ, [_I64, _I32, _I64]),
    "svox2_ship_trace": (C.c_int, [_G, _P, _P, _I64, _I64, _I64, _O, _P, _I32, _I32, _P, _SZ, _I64, _P, _P]),
    "svox2_ship_shade": (C.c_int, [_G, _I64, _I32, _I64, _P, _P, _SZ, _P]),
    "svox2_ship_ray_logT": (_P, [_P, _SZ, _I64, _I32, _I64]),
    "svox2_ship_combine": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _P, _P, _P, _P]),
    "svox2_ship_backward": (C.c_int, [_G, _I64, _O, _I32, _I64, _P, _P, _F, _GR, _P, _SZ, _P]),
    "svox2_mg_step_begin": (C.c_int, [_G, _P, _P, _I64, _O, C.POINTER(MgStepT), _P, _I64, _P, _SZ, _I64, _P, _P, _P,
                                      _SZ, _I64, _P]),
    "svox2_mg_step_shade": (C.c_int, [_G, _I64, _O, C.POINTER(MgStepT), _P, _P, _P, _SZ, _I64, _P, _SZ, _I64, _P]),
    "svox2_mg_step_backward": (C.c_int, [_G, _I64, _O, C.POINTER(MgStepT), _P, _P, _P, _P, _SZ, _I64, _P, _SZ, _I64,
                                         _P]),
    "svox2_mg_slot_pack": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _SZ, _I64, _I32, _I64, _I64, _P]),
    "svox2_mg_slot_unpack_add": (C.c_int, [_P, _I64, _P, _I64, _P, _SZ, _I64, _I32, _I64, _I64, _P]),
    "svox2_mg_step_finish": (C.c_int, [_G, _P, _P, _I64, C.POINTER(MgStepT), _P, _P, _P, _SZ, _I64, _P]),
    "svox2_rmsprop_step_publish": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _P, _I32, _I32, _I32, _F, _F, _F, _F, _P]),
    "svox2_pack_grad_rows": (C.c_int, [_P, _I64, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P]),
    "svox2_unpack_rows": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, _P, _I64, _I32, _P, _P, _P]),
    "svox2_peer_store": (C.c_int, [_P, _I32, _I64, _P, _I64, _P]),
    "svox2_slab_prefix": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "svox2_slab_combine": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _P, _P, _P, _P]),
    "svox2_rmsprop_step_packed": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _I64, _P, _I64, _P, _I64, _P, _P,
                                            _P, _I64, _P, _I64, _F, _F, _F, _F, _F, _F, _P]),
}

_lib = None


def lib():
    """Load the CUDA library; raise loudly if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"svox2_b200: CUDA library not built ({LIB_PATH} missing). Run "
                "`make -C svox2_b200/csrc` or `__graft_entry__.build()`; there is no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().svox2_last_error().decode("utf-8", "replace")
        raise Svox2Error(msg or f"svox2_b200 call failed with status {status}")
