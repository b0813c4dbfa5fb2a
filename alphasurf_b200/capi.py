"""ctypes binding of the C ABI declared in ``include/asurf.h`` (library: ``alphasurf_b200/csrc/libasurf.so``).

This is plumbing only: torch supplies device memory (``tensor.data_ptr()``) and the current CUDA stream; every
computation happens in the hand-written sm_100a kernels behind the C ABI.  There is NO fallback: if the library
is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libasurf.so")
_LIB = None

ASURF_OK = 0
ASURF_E_INVALID = -1
ASURF_E_UNSUPPORTED = -2
ASURF_E_NOMEM = -3


class AsurfError(RuntimeError):
    pass


class GridT(C.Structure):
    _fields_ = [("links", C.c_void_p), ("size", C.c_int32 * 3), ("density", C.c_void_p), ("surface", C.c_void_p),
                ("sh", C.c_void_p), ("level_set", C.c_void_p), ("level_set_num", C.c_int32), ("basis_dim", C.c_int32),
                ("sh_dim", C.c_int32), ("capacity", C.c_int64), ("offset", C.c_float * 3), ("scaling", C.c_float * 3),
                ("fake_sample_std", C.c_float), ("truncated_vol_render_a", C.c_float), ("accel", C.c_void_p),
                ("work", C.c_void_p), ("background_links", C.c_void_p), ("background_data", C.c_void_p),
                ("background_reso", C.c_int32), ("background_nlayers", C.c_int32)]


class OptT(C.Structure):
    _fields_ = [("background_brightness", C.c_float), ("step_size", C.c_float), ("sigma_thresh", C.c_float),
                ("stop_thresh", C.c_float), ("near_clip", C.c_float), ("use_spheric_clip", C.c_int32),
                ("last_sample_opaque", C.c_int32), ("surf_fake_sample", C.c_int32),
                ("surf_fake_sample_min_vox_len", C.c_float), ("limited_fake_sample", C.c_int32),
                ("no_surf_grad_from_sh", C.c_int32), ("alpha_activation_type", C.c_int32),
                ("fake_sample_l_dist", C.c_int32), ("fake_sample_normalize_surf", C.c_int32),
                ("only_outward_intersect", C.c_int32), ("truncated_vol_render", C.c_int32),
                ("trunc_vol_weight_min", C.c_float)]


class RaysT(C.Structure):
    _fields_ = [("origins", C.c_void_p), ("dirs", C.c_void_p), ("n_rays", C.c_int64)]


class GradsT(C.Structure):
    _fields_ = [("grad_density", C.c_void_p), ("grad_surface", C.c_void_p), ("grad_sh", C.c_void_p),
                ("grad_fake_sample_std", C.c_void_p), ("mask", C.c_void_p), ("grad_background", C.c_void_p),
                ("mask_background", C.c_void_p)]


class FusedT(C.Structure):
    _fields_ = [("beta_loss", C.c_float), ("sparsity_loss", C.c_float), ("fused_surf_norm_reg_scale", C.c_float),
                ("lambda_l2", C.c_float), ("lambda_l1", C.c_float), ("lambda_l_dist", C.c_float),
                ("lambda_l_entropy", C.c_float), ("no_norm_weight_l_entropy", C.c_int32),
                ("lambda_l_dist_a", C.c_float), ("lambda_l_entropy_a", C.c_float), ("lambda_l_samp_dist", C.c_float),
                ("lambda_l_di", C.c_float), ("l_di_alpha_thresh", C.c_float), ("surf_sparse_alpha_thresh", C.c_float),
                ("lambda_inplace_surf_sparse", C.c_float), ("lambda_inwards_norm_loss", C.c_float),
                ("lambda_conv_mode_samp", C.c_float), ("l_dist_max_sample", C.c_int32), ("norm_rays", C.c_int64)]


def lib():
    """Load (building first if the sources are newer and nvcc is available) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        from . import build
        build.build_library()
    L = C.CDLL(LIB_PATH)
    L.asurf_last_error.restype = C.c_char_p
    L.asurf_accel_words.restype = C.c_int64
    L.asurf_launch_count.restype = C.c_uint64
    L.asurf_accel_words.argtypes = [C.POINTER(C.c_int32)]
    for name in EXPORTS:
        getattr(L, name)  # fail loudly on a stale library
    if os.environ.get("ASURF_WAVE") == "0":   # debugging knob: persistent shading kernels only
        L.asurf_debug_set_wave(C.c_int32(0))
    _LIB = L
    return L


# every symbol include/asurf.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = [
    "asurf_last_error", "asurf_abi_version", "asurf_accel_words", "asurf_accel_build", "asurf_work_build", "asurf_surf_trav_forward", "asurf_surf_trav_scalar",
    "asurf_surf_trav_backward", "asurf_surf_trav_fused", "asurf_debug_ray_bounds", "asurf_debug_trace", "asurf_debug_set_skip", "asurf_debug_set_wave", "asurf_debug_set_seg", "asurf_debug_counters", "asurf_debug_work_cache_copy", "asurf_debug_work_cache_valid", "asurf_debug_set_normal_tile", "asurf_debug_last_verdict",
    "asurf_cuvol_forward", "asurf_cuvol_stats", "asurf_cuvol_image", "asurf_cuvol_scalar", "asurf_cuvol_backward", "asurf_cuvol_fused",
    "asurf_accel_dist_prop", "asurf_sample_grid", "asurf_sample_grid_backward", "asurf_cubic_extract_iso_pts", "asurf_dilate", "asurf_grid_weight_render", "asurf_sparse_grid_weight_render",
    "asurf_sparse_grid_mask_render", "asurf_sparse_grid_visibility_render_surf", "asurf_rmsprop_step", "asurf_sgd_step", "asurf_rows_pack", "asurf_rows_unpack_add", "asurf_mask_pack", "asurf_mask_unpack_or", "asurf_msi_forward", "asurf_msi_forward_image", "asurf_msi_backward", "asurf_msi_tv_grad_sparse", "asurf_debug_bg_state", "asurf_tv", "asurf_tv_grad", "asurf_tv_grad_sparse",
    "asurf_surf_tv_grad_sparse", "asurf_surf_sign_change_grad_sparse", "asurf_surface_normal_grad", "asurf_lumisphere_tv_grad_sparse", "asurf_alpha_surf_sparsify_grad_sparse", "asurf_surface_normal_grad_sparse", "asurf_profile_enable", "asurf_profile_read", "asurf_profile_read_stages", "asurf_launch_count", "asurf_release",
]


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().asurf_last_error().decode("utf-8", "replace")
        if rc == ASURF_E_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise AsurfError("%s failed (code %d): %s" % (what or "asurf call", rc, msg))


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def size3(sz):
    return (C.c_int32 * 3)(int(sz[0]), int(sz[1]), int(sz[2]))


def make_opt(d) -> OptT:
    """dict or object with the RenderOptions fields -> OptT."""
    o = OptT()
    get = (lambda k: d[k]) if isinstance(d, dict) else (lambda k: getattr(d, k))
    for name, ctype in OptT._fields_:
        v = get(name)
        setattr(o, name, float(v) if ctype is C.c_float else int(v))
    return o


def make_fused(d: dict, norm_rays: int = 0) -> FusedT:
    f = FusedT()
    for name, ctype in FusedT._fields_:
        if name == "norm_rays":
            continue
        v = d.get(name, 0)
        setattr(f, name, float(v) if ctype is C.c_float else int(v))
    f.norm_rays = int(norm_rays)
    return f
