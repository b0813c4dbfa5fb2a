"""Drop-in replacement for the reference's ``svox2.csrc`` extension module (hot path only).

The reference's ``svox2`` Python package does ``import svox2.csrc as _C`` (/root/reference/svox2/utils.py:32-46)
and calls ``_C.<function>(specs..., scalars..., tensors...)`` positionally
(/root/reference/svox2/csrc/svox2.cpp:139-291).  This module offers the same names, argument order, spec classes
and error behaviour (``RuntimeError`` where the reference's ``TORCH_CHECK`` fires) on top of the C ABI in
``include/asurf.h``.  Installing it as ``svox2/csrc.py`` makes ``svox2.SparseGrid`` run on
the B200 kernels unchanged.  Torch is used for memory and streams only; nothing is computed in Python.

Deliberately absent (part of the contract): ``volume_render_surf_trav_image``.
"""
import ctypes as C

import weakref

import torch

from . import capi

BASIS_TYPE_SH = 1
SURFACE_TYPE_NONE = 100


# ---- spec classes: default-constructible, fields as svox2.cpp:209-290 ------------------------------------------
class SparseGridSpec:
    def __init__(self):
        self.density_data = None
        self.surface_data = None
        self.level_set_data = None
        self.sh_data = None
        self.links = None
        self._offset = None
        self._scaling = None
        self.basis_dim = 0
        self.basis_type = BASIS_TYPE_SH
        self.surface_type = SURFACE_TYPE_NONE
        self.basis_data = None
        self.background_links = None
        self.background_data = None
        self.fake_sample_std = 1.0
        self.truncated_vol_render_a = 1.0


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
        self.masks = None


class RayVoxIntersecSpec:
    def __init__(self):
        self.voxel_ls = None
        self.vox_start_i = None
        self.vox_num = None


class RenderOptions:
    def __init__(self):
        self.background_brightness = 1.0
        self.step_size = 0.5
        self.sigma_thresh = 1e-10
        self.stop_thresh = 1e-7
        self.near_clip = 0.0
        self.use_spheric_clip = False
        self.last_sample_opaque = False
        self.surf_fake_sample = False
        self.surf_fake_sample_min_vox_len = 0.0
        self.limited_fake_sample = False
        self.no_surf_grad_from_sh = False
        self.alpha_activation_type = 0
        self.fake_sample_l_dist = True
        self.fake_sample_normalize_surf = False
        self.only_outward_intersect = False
        self.truncated_vol_render = False
        self.trunc_vol_weight_min = 0.0


class GridOutputGrads:
    def __init__(self):
        self.grad_density_out = None
        self.grad_sh_out = None
        self.grad_surface_out = None
        self.grad_fake_sample_std_out = None
        self.grad_basis_out = None
        self.grad_background_out = None
        self.mask_out = None
        self.mask_background_out = None


# ---- argument checks (include/util.hpp:3-12, include/data_spec.hpp:58-81) ----------------------------------------
def _check_input(t, name):
    if t is None or not torch.is_tensor(t):
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _check_cpu_input(t, name):
    if t is None or not torch.is_tensor(t) or t.is_cuda:
        raise RuntimeError("%s must be a CPU tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _check_f32(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)


def _defined(t):
    return t is not None and torch.is_tensor(t)


def _check_grid(grid: SparseGridSpec):
    _check_input(grid.density_data, "density_data")
    _check_input(grid.sh_data, "sh_data")
    _check_input(grid.links, "links")
    if grid.surface_type != SURFACE_TYPE_NONE:
        _check_input(grid.surface_data, "surface_data")
        _check_input(grid.level_set_data, "level_set_data")
    _check_cpu_input(grid._offset, "_offset")
    _check_cpu_input(grid._scaling, "_scaling")
    if grid.density_data.dim() != 2 or grid.sh_data.dim() != 2 or grid.links.dim() != 3:
        raise RuntimeError("density_data / sh_data must be 2-D and links 3-D")
    if grid.links.dtype != torch.int32:
        raise RuntimeError("links must be int32")
    _check_f32(grid.density_data, "density_data")
    _check_f32(grid.sh_data, "sh_data")
    if _defined(grid.background_links) and grid.background_links.numel() > 0:     # MSI background (data_spec.hpp:47-48)
        _check_input(grid.background_links, "background_links")
        _check_input(grid.background_data, "background_data")
        if grid.background_links.dim() != 2 or grid.background_links.dtype != torch.int32:
            raise RuntimeError("background_links must be a 2-D int32 tensor")
        if grid.background_data.dim() != 3 or grid.background_data.shape[2] != 4 or grid.background_data.dtype != torch.float32:
            raise RuntimeError("background_data must be a float32 (n, nlayers, 4) tensor")
    if grid.basis_type != BASIS_TYPE_SH:
        raise NotImplementedError("only the SH basis is on the B200 hot path")


def _check_rays(rays: RaysSpec):
    _check_input(rays.origins, "origins")
    _check_input(rays.dirs, "dirs")
    if rays.masks is not None:
        _check_input(rays.masks, "masks")
    _check_f32(rays.origins, "origins")
    _check_f32(rays.dirs, "dirs")


def _check_grads(g: GridOutputGrads):
    for n in ("grad_density_out", "grad_sh_out", "grad_surface_out", "grad_fake_sample_std_out"):
        t = getattr(g, n)
        if _defined(t):
            _check_input(t, n)
            _check_f32(t, n)
    if _defined(g.mask_out) and g.mask_out.numel() > 0:
        _check_input(g.mask_out, "mask_out")


# ---- occupancy pyramid cache: one per (links storage, version) ------------------------------------------------------
_ACCEL = {}


def accel_for(links: torch.Tensor) -> torch.Tensor:
    """Bitmap pyramid of ``links`` (built by asurf_accel_build); rebuilt when links is modified in place or when another
    tensor object turns up at the same address (the caching allocator hands freed blocks out again: a pruned grid's new
    ``links`` may well land where the old one was, with the same shape and a fresh version counter)."""
    key = (links.data_ptr(), tuple(links.shape), links.device.index)
    ver = links._version
    hit = _ACCEL.get(key)
    if hit is not None and hit[0]() is links and hit[1] == ver:
        return hit[2]
    L = capi.lib()
    words = L.asurf_accel_words(capi.size3(links.shape))
    acc = torch.empty((words,), dtype=torch.int64, device=links.device)
    capi.check(L.asurf_accel_build(capi.ptr(links), capi.size3(links.shape), capi.ptr(acc),
                                   capi.current_stream(links.device)), "asurf_accel_build")
    if len(_ACCEL) > 8:
        _ACCEL.clear()
    _ACCEL[key] = (weakref.ref(links), ver, acc)
    return acc


def _grid_t(grid: SparseGridSpec, need_surface=True, need_accel=True):
    g = capi.GridT()
    g.links = grid.links.data_ptr()
    g.size[:] = [int(s) for s in grid.links.shape]
    g.density = grid.density_data.data_ptr()
    has_surf = grid.surface_type != SURFACE_TYPE_NONE and _defined(grid.surface_data)
    g.surface = grid.surface_data.data_ptr() if has_surf else None
    g.level_set = grid.level_set_data.data_ptr() if has_surf else None
    g.level_set_num = int(grid.level_set_data.shape[0]) if has_surf else 0
    g.sh = grid.sh_data.data_ptr()
    g.basis_dim = int(grid.basis_dim)
    g.sh_dim = int(grid.sh_data.shape[1])
    g.capacity = int(grid.density_data.shape[0])
    g.offset[:] = [float(v) for v in grid._offset.tolist()]
    g.scaling[:] = [float(v) for v in grid._scaling.tolist()]
    g.fake_sample_std = float(grid.fake_sample_std)
    g.truncated_vol_render_a = float(grid.truncated_vol_render_a)
    if _defined(grid.background_links) and grid.background_links.numel() > 0:
        g.background_links = grid.background_links.data_ptr()
        g.background_data = grid.background_data.data_ptr()
        g.background_reso = int(grid.background_links.shape[1])
        g.background_nlayers = int(grid.background_data.shape[1])
    acc = accel_for(grid.links) if need_accel else None
    g.accel = acc.data_ptr() if need_accel else None
    return g, acc


def _rays_t(rays: RaysSpec):
    r = capi.RaysT()
    r.origins = rays.origins.data_ptr()
    r.dirs = rays.dirs.data_ptr()
    r.n_rays = int(rays.origins.shape[0])
    return r


def _grads_t(g: GridOutputGrads):
    o = capi.GradsT()
    o.grad_density = g.grad_density_out.data_ptr() if _defined(g.grad_density_out) else None
    o.grad_surface = g.grad_surface_out.data_ptr() if _defined(g.grad_surface_out) else None
    o.grad_sh = g.grad_sh_out.data_ptr() if _defined(g.grad_sh_out) else None
    o.grad_fake_sample_std = (g.grad_fake_sample_std_out.data_ptr()
                              if _defined(g.grad_fake_sample_std_out) and g.grad_fake_sample_std_out.numel() > 0 else None)
    o.mask = g.mask_out.data_ptr() if _defined(g.mask_out) and g.mask_out.numel() > 0 else None
    o.grad_background = (g.grad_background_out.data_ptr()
                         if _defined(g.grad_background_out) and g.grad_background_out.numel() > 0 else None)
    o.mask_background = (g.mask_background_out.data_ptr()
                         if _defined(g.mask_background_out) and g.mask_background_out.numel() > 0 else None)
    return o


# ---- alpha-Surf renderer (render_lerp_kernel_surf_trav.cu:3596-3942) -------------------------------------------------
def volume_render_surf_trav(grid, rays, opt):
    _check_grid(grid)
    _check_rays(rays)
    out = torch.empty_like(rays.origins)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_surf_trav_forward(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                      capi.ptr(out), None, capi.current_stream()),
                   "volume_render_surf_trav")
    return out


def volume_render_surf_trav_backward(grid, rays, opt, grad_out, color_cache, grads):
    _check_grid(grid)
    _check_rays(rays)
    _check_grads(grads)
    _check_input(grad_out, "grad_out")
    _check_input(color_cache, "color_cache")
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_surf_trav_backward(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                       capi.ptr(grad_out), capi.ptr(color_cache),
                                                       C.byref(_grads_t(grads)), capi.current_stream()),
                   "volume_render_surf_trav_backward")


# ---- scalar renders for evaluation (render_lerp_kernel_surf_trav.cu:3944-4050+; svox2.py:3690-3830) ------------------
def _surf_trav_scalar(name, grid, rays, opt, mode, param, width=1, max_sample=0):
    _check_grid(grid)
    _check_rays(rays)
    Q = rays.origins.shape[0]
    out = torch.empty((Q,) if width == 1 else (Q, width), dtype=rays.origins.dtype, device=rays.origins.device)
    out2 = torch.empty_like(out) if mode == 5 else None
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_surf_trav_scalar(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                     C.c_int32(mode), C.c_float(param), C.c_int32(max_sample), capi.ptr(out),
                                                     capi.ptr(out2) if out2 is not None else None,
                                                     capi.current_stream()), name)
    return (out, out2) if mode == 5 else out


def volume_render_expected_term_surf_trav(grid, rays, opt):
    """expected ray termination depth, (Q,) (:3944-3963)"""
    return _surf_trav_scalar("volume_render_expected_term_surf_trav", grid, rays, opt, 0, 0.0)


def volume_render_mode_term_surf_trav(grid, rays, opt, weight_thresh):
    """depth of the sample with the largest weight, 0 if the accumulated weight <= weight_thresh, (Q,) (:3965-3985)"""
    return _surf_trav_scalar("volume_render_mode_term_surf_trav", grid, rays, opt, 1, float(weight_thresh))


def volume_render_sigma_thresh_surf_trav(grid, rays, opt, sigma_thresh):
    """depth of the first sample whose alpha exceeds sigma_thresh, (Q,) (:3987-4007)"""
    return _surf_trav_scalar("volume_render_sigma_thresh_surf_trav", grid, rays, opt, 2, float(sigma_thresh))


def volume_render_alpha_surf_trav(grid, rays, opt, thresh):
    """alpha of the first sample whose alpha exceeds thresh, (Q,) (:4009-4029)"""
    return _surf_trav_scalar("volume_render_alpha_surf_trav", grid, rays, opt, 3, float(thresh))


def render_normal_surf_trav(grid, rays, opt):
    """un-normalised surface gradient at the first sample with alpha > 0, (Q, 3) (:4031-4050)"""
    return _surf_trav_scalar("render_normal_surf_trav", grid, rays, opt, 4, 0.0, width=3)


def extract_pts_surf_trav(grid, rays, opt, max_sample, alpha_thresh):
    """(depths, alphas) of the first max_sample samples with alpha > alpha_thresh per ray, zero-padded, (Q, max_sample)
    (:4052-4081; svox2.py:3938)"""
    if int(max_sample) <= 0:
        z = torch.zeros((rays.origins.shape[0], 0), dtype=rays.origins.dtype, device=rays.origins.device)
        return z, z.clone()
    return _surf_trav_scalar("extract_pts_surf_trav", grid, rays, opt, 5, float(alpha_thresh), width=int(max_sample),
                             max_sample=int(max_sample))


# Global batch size used to normalise the fused losses; None = this call's ray count (the reference behaviour).
# The ray-sharded data-parallel wrapper (alphasurf_b200.dist) sets it to the global Q.
_NORM_RAYS = None


def set_loss_norm_rays(n):
    global _NORM_RAYS
    _NORM_RAYS = None if n is None else int(n)


def volume_render_surf_trav_fused(grid, rays, opt, rgb_gt, beta_loss, sparsity_loss, fused_surf_norm_reg_scale,
                                  fused_surf_norm_reg_con_check, fused_surf_norm_reg_ignore_empty, lambda_l2, lambda_l1,
                                  lambda_l_dist, lambda_l_entropy, no_norm_weight_l_entropy, lambda_l_dist_a,
                                  lambda_l_entropy_a, lambda_l_samp_dist, lambda_l_di, l_di_alpha_thresh,
                                  surf_sparse_alpha_thresh, lambda_inplace_surf_sparse, lambda_inwards_norm_loss,
                                  lambda_conv_mode_samp, l_dist_max_sample, rgb_out, grads):
    _check_input(rgb_gt, "rgb_gt")
    _check_input(rgb_out, "rgb_out")
    _check_grid(grid)
    _check_rays(rays)
    _check_grads(grads)
    f = capi.make_fused(dict(
        beta_loss=beta_loss, sparsity_loss=sparsity_loss, fused_surf_norm_reg_scale=fused_surf_norm_reg_scale,
        lambda_l2=lambda_l2, lambda_l1=lambda_l1, lambda_l_dist=lambda_l_dist, lambda_l_entropy=lambda_l_entropy,
        no_norm_weight_l_entropy=no_norm_weight_l_entropy, lambda_l_dist_a=lambda_l_dist_a,
        lambda_l_entropy_a=lambda_l_entropy_a, lambda_l_samp_dist=lambda_l_samp_dist, lambda_l_di=lambda_l_di,
        l_di_alpha_thresh=l_di_alpha_thresh, surf_sparse_alpha_thresh=surf_sparse_alpha_thresh,
        lambda_inplace_surf_sparse=lambda_inplace_surf_sparse, lambda_inwards_norm_loss=lambda_inwards_norm_loss,
        lambda_conv_mode_samp=lambda_conv_mode_samp, l_dist_max_sample=l_dist_max_sample), _NORM_RAYS or 0)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_surf_trav_fused(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                    capi.ptr(rgb_gt), C.byref(f), capi.ptr(rgb_out),
                                                    C.byref(_grads_t(grads)), None, capi.current_stream()),
                   "volume_render_surf_trav_fused")


# ---- Plenoxels renderer (render_lerp_kernel_cuvol.cu:1120-1354) -----------------------------------------------------------
def volume_render_cuvol(grid, rays, opt):
    _check_grid(grid)
    _check_rays(rays)
    out = torch.empty_like(rays.origins)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_forward(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                  capi.ptr(out), None, capi.current_stream()), "volume_render_cuvol")
    return out


# depth renders of the cuvol backend (render_lerp_kernel_cuvol.cu:1356-1442; svox2.py:3732-3775)
def _cuvol_scalar(name, grid, rays, opt, mode, param, max_sample=0):
    _check_grid(grid)
    _check_rays(rays)
    Q = rays.origins.shape[0]
    kw = dict(dtype=rays.origins.dtype, device=rays.origins.device)
    out = torch.empty((Q, max_sample) if mode == 2 else (Q,), **kw)
    out2 = torch.empty((Q, max_sample), **kw) if mode == 2 else None
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_scalar(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                 C.c_int32(mode), C.c_float(param), C.c_int32(max_sample), capi.ptr(out),
                                                 capi.ptr(out2) if out2 is not None else None, capi.current_stream()), name)
    return (out, out2) if mode == 2 else out


def volume_render_expected_term(grid, rays, opt, weight_thresh):
    """expected termination depth; 0 where the accumulated weight <= weight_thresh, (Q,)"""
    return _cuvol_scalar("volume_render_expected_term", grid, rays, opt, 0, float(weight_thresh))


def volume_render_mode_term(grid, rays, opt, weight_thresh):
    """depth of the heaviest sample; 0 where the accumulated weight <= weight_thresh, (Q,)"""
    return _cuvol_scalar("volume_render_mode_term", grid, rays, opt, 1, float(weight_thresh))


def volume_render_med_term(grid, rays, opt, max_sample):
    """(depths, sigmas) of the first max_sample samples per ray, zero-padded, each (Q, max_sample)"""
    return _cuvol_scalar("volume_render_med_term", grid, rays, opt, 2, 0.0, int(max_sample))


def volume_render_sigma_thresh(grid, rays, opt, sigma_thresh):
    """depth of the first sample whose sigma exceeds sigma_thresh, (Q,)"""
    return _cuvol_scalar("volume_render_sigma_thresh", grid, rays, opt, 3, float(sigma_thresh))


def volume_render_cuvol_image(grid, cam, opt):
    _check_grid(grid)
    _check_input(cam.c2w, "c2w")
    if cam.ndc_coeffx > 0.0:
        raise NotImplementedError("NDC cameras are outside the B200 hot path")
    out = torch.empty((int(cam.height), int(cam.width), 3), dtype=grid.sh_data.dtype, device=grid.sh_data.device)
    c2w = (C.c_float * 12)(*[float(v) for v in cam.c2w[:3, :4].reshape(-1).tolist()])
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_image(C.byref(g), c2w, C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx),
                                                C.c_float(cam.cy), C.c_int32(int(cam.width)), C.c_int32(int(cam.height)),
                                                C.byref(capi.make_opt(opt)), capi.ptr(out), capi.current_stream()),
                   "volume_render_cuvol_image")
    return out


def volume_render_cuvol_backward(grid, rays, opt, grad_out, color_cache, grads):
    _check_grid(grid)
    _check_rays(rays)
    _check_grads(grads)
    _check_input(grad_out, "grad_out")
    _check_input(color_cache, "color_cache")
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_backward(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                   capi.ptr(grad_out), capi.ptr(color_cache), C.byref(_grads_t(grads)),
                                                   capi.current_stream()), "volume_render_cuvol_backward")


def volume_render_cuvol_fused(grid, rays, opt, rgb_gt, beta_loss, sparsity_loss, rgb_out, grads):
    _check_input(rgb_gt, "rgb_gt")
    _check_input(rgb_out, "rgb_out")
    _check_grid(grid)
    _check_rays(rays)
    _check_grads(grads)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_fused(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                capi.ptr(rgb_gt), C.c_float(beta_loss), C.c_float(sparsity_loss),
                                                C.c_int64(_NORM_RAYS or 0), capi.ptr(rgb_out), C.byref(_grads_t(grads)),
                                                capi.current_stream()), "volume_render_cuvol_fused")


# ---- grid maintenance (misc_kernel.cu:1022-1058) ----------------------------------------------------------------------------
def accel_dist_prop(links):
    """In place: empty vertices of ``links`` receive the negative skip codes the cuvol marcher reads."""
    _check_input(links, "grid")
    if links.is_floating_point() or links.dim() != 3:
        raise RuntimeError("accel_dist_prop expects the 3-D integer links tensor")
    if links.dtype != torch.int32:
        raise RuntimeError("links must be int32")
    with torch.cuda.device(links.device):
        capi.check(capi.lib().asurf_accel_dist_prop(capi.ptr(links), capi.size3(links.shape), capi.current_stream()),
                   "accel_dist_prop")
    links.add_(0)   # torch bumps tensor._version only for its own ops: force it so that cached pyramids are rebuilt


def dilate(grid):
    """3x3x3 OR of a 3-D bool tensor (misc_kernel.cu:1005-1020); returns a new tensor."""
    _check_input(grid, "grid")
    if grid.is_floating_point() or grid.dim() != 3:
        raise RuntimeError("dilate expects a 3-D non-floating tensor")
    if grid.dtype != torch.bool:
        raise RuntimeError("dilate expects a bool tensor")     # the reference's accessor is bool as well
    out = torch.empty_like(grid)
    with torch.cuda.device(grid.device):
        capi.check(capi.lib().asurf_dilate(capi.ptr(grid), capi.size3(grid.shape), capi.ptr(out), capi.current_stream()),
                   "dilate")
    return out


def _f3(t):
    return (C.c_float * 3)(*[float(v) for v in t.reshape(-1).tolist()[:3]])


def _cam_args(cam):
    _check_input(cam.c2w, "c2w")
    if cam.ndc_coeffx > 0.0:
        raise NotImplementedError("NDC cameras are outside the B200 hot path")
    c2w = (C.c_float * 12)(*[float(v) for v in cam.c2w[:3, :4].reshape(-1).tolist()])
    return (c2w, C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.c_int32(int(cam.width)),
            C.c_int32(int(cam.height)))


def grid_weight_render(data, cam, step_size, stop_thresh, last_sample_opaque, offset, scaling, grid_weight_out):
    """max rendering weight per vertex of a dense (X,Y,Z) sigma volume over the camera's pixels (misc_kernel.cu:1084-1111)"""
    _check_input(data, "data")
    _check_input(offset, "offset")
    _check_input(scaling, "scaling")
    _check_input(grid_weight_out, "grid_weight_out")
    if data.dim() != 3 or tuple(grid_weight_out.shape) != tuple(data.shape):
        raise RuntimeError("grid_weight_render expects (X,Y,Z) data and an output of the same shape")
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_grid_weight_render(capi.ptr(data), capi.size3(data.shape), _f3(offset), _f3(scaling),
                                                       *_cam_args(cam), C.c_float(step_size), C.c_float(stop_thresh),
                                                       C.c_int32(1 if last_sample_opaque else 0), capi.ptr(grid_weight_out),
                                                       capi.current_stream()), "grid_weight_render")


def sparse_grid_weight_render(grid, cam, step_size, stop_thresh, offset, scaling, grid_weight_out):
    """max transmittance reaching each vertex, marched through the sparse grid (misc_kernel.cu:1113-1138)"""
    _check_grid(grid)
    _check_input(offset, "offset")
    _check_input(scaling, "scaling")
    _check_input(grid_weight_out, "grid_weight_out")
    if tuple(grid_weight_out.shape) != tuple(grid.links.shape):
        raise RuntimeError("sparse_grid_weight_render expects an output shaped like links")
    with torch.cuda.device(grid.links.device):
        capi.check(capi.lib().asurf_sparse_grid_weight_render(capi.ptr(grid.links), capi.ptr(grid.density_data),
                                                              capi.size3(grid.links.shape), _f3(offset), _f3(scaling),
                                                              *_cam_args(cam), C.c_float(step_size), C.c_float(stop_thresh),
                                                              capi.ptr(grid_weight_out), capi.current_stream()),
                   "sparse_grid_weight_render")


def sparse_grid_mask_render(grid, rays, near_clip, grid_mask):
    """rows of the voxels the rays pass through are set to 1 in the float tensor grid_mask (N,) (misc_kernel.cu:1158-1175)"""
    _check_grid(grid)
    _check_rays(rays)
    _check_input(grid_mask, "grid_mask")
    with torch.cuda.device(grid.links.device):
        capi.check(capi.lib().asurf_sparse_grid_mask_render(capi.ptr(grid.links), capi.size3(grid.links.shape),
                                                            _f3(grid._offset), _f3(grid._scaling), capi.ptr(rays.origins),
                                                            capi.ptr(rays.dirs), C.c_int64(rays.origins.shape[0]),
                                                            C.c_float(near_clip), capi.ptr(grid_mask), capi.current_stream()),
                   "sparse_grid_mask_render")


def sparse_grid_visbility_render_surf(grid, cam, visibility_out):
    """visibility_out (N,) += number of pixels whose ray reaches the vertex's voxel before its first surface intersection
    (misc_kernel.cu:1140-1156; the name is the reference's spelling)"""
    _check_grid(grid)
    _check_input(visibility_out, "visibility_out")
    if not _defined(grid.surface_data) or grid.surface_type == SURFACE_TYPE_NONE:
        raise RuntimeError("sparse_grid_visbility_render_surf needs a grid with surface data")
    with torch.cuda.device(grid.links.device):
        capi.check(capi.lib().asurf_sparse_grid_visibility_render_surf(
            capi.ptr(grid.links), capi.ptr(grid.surface_data), capi.ptr(grid.level_set_data),
            C.c_int32(int(grid.level_set_data.shape[0])), capi.size3(grid.links.shape), _f3(grid._offset), _f3(grid._scaling),
            *_cam_args(cam), capi.ptr(visibility_out), capi.current_stream()), "sparse_grid_visbility_render_surf")


# ---- point queries (svox2_kernel.cu:384-582) --------------------------------------------------------------------------------
def _check_points(points):
    _check_input(points, "points")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must be (P, 3)")


def _sample(grid, data, missing, points, name):
    out = torch.empty((points.shape[0], data.shape[1]), dtype=points.dtype, device=points.device)
    capi.check(capi.lib().asurf_sample_grid(capi.ptr(grid.links), capi.size3(grid.links.shape), _f3(grid._offset),
                                            _f3(grid._scaling), capi.ptr(data), C.c_int32(int(data.shape[1])),
                                            C.c_float(missing), capi.ptr(points), C.c_int64(points.shape[0]), capi.ptr(out),
                                            capi.current_stream()), name)
    return out


def sample_grid(grid, points, want_colors):
    """(density (P,1), sh (P,D) -- (0,D) unless want_colors) at world-space points; empty corners read 0"""
    _check_grid(grid)
    _check_points(points)
    with torch.cuda.device(points.device):
        dens = _sample(grid, grid.density_data, 0.0, points, "sample_grid")
        sh = (_sample(grid, grid.sh_data, 0.0, points, "sample_grid") if want_colors
              else torch.empty((0, grid.sh_data.shape[1]), dtype=points.dtype, device=points.device))
    return dens, sh


def sample_grid_sh_surf(grid, points, want_colors, want_surfaces, default_surf):
    """(sh (P,D), surface (P,1)); empty corners read 0 for SH and default_surf for the surface"""
    _check_grid(grid)
    _check_points(points)
    empty = lambda t: torch.empty((0, t.shape[1]), dtype=points.dtype, device=points.device)
    with torch.cuda.device(points.device):
        sh = _sample(grid, grid.sh_data, 0.0, points, "sample_grid_sh_surf") if want_colors else empty(grid.sh_data)
        surf = (_sample(grid, grid.surface_data, float(default_surf), points, "sample_grid_sh_surf") if want_surfaces
                else empty(grid.surface_data))
    return sh, surf


def sample_grid_raw_alpha(grid, points, empty_raw):
    """raw opacity (P,1); empty corners read empty_raw"""
    _check_grid(grid)
    _check_points(points)
    with torch.cuda.device(points.device):
        return _sample(grid, grid.density_data, float(empty_raw), points, "sample_grid_raw_alpha")


def sample_grid_backward(grid, points, grad_out_density, grad_out_sh, grad_density_out, grad_sh_out, want_colors):
    _check_grid(grid)
    _check_points(points)
    for t, n in ((grad_out_density, "grad_out_density"), (grad_out_sh, "grad_out_sh"), (grad_density_out, "grad_density_out"),
                 (grad_sh_out, "grad_sh_out")):
        _check_input(t, n)
    if grad_out_density.dim() != 2 or grad_out_sh.dim() != 2:
        raise RuntimeError("sample_grid_backward expects 2-D output gradients")
    args = (capi.ptr(grid.links), capi.size3(grid.links.shape), _f3(grid._offset), _f3(grid._scaling), capi.ptr(points),
            C.c_int64(points.shape[0]))
    with torch.cuda.device(points.device):
        capi.check(capi.lib().asurf_sample_grid_backward(*args, capi.ptr(grad_out_density),
                                                         C.c_int32(int(grad_density_out.shape[1])), capi.ptr(grad_density_out),
                                                         capi.current_stream()), "sample_grid_backward")
        if want_colors:
            capi.check(capi.lib().asurf_sample_grid_backward(*args, capi.ptr(grad_out_sh), C.c_int32(int(grad_sh_out.shape[1])),
                                                             capi.ptr(grad_sh_out), capi.current_stream()),
                       "sample_grid_backward")


def cubic_extract_iso_pts(links, level_data, mask_data, cell_ids, n_sample, density_thresh):
    """zero crossings of the level function along 3 n_sample^2 lattice lines per cell, (n_cells, 3 n_sample^2, 3) in grid
    coordinates, zeros where there is none (svox2_kernel.cu:542-582; svox2.py:4537)"""
    for t, n in ((level_data, "level_data"), (mask_data, "mask_data"), (links, "links"), (cell_ids, "cell_ids")):
        _check_input(t, n)
    if links.dtype != torch.int32 or cell_ids.dtype != torch.int32:
        raise RuntimeError("links and cell_ids must be int32")
    n_sample = int(n_sample)
    out = torch.zeros((cell_ids.shape[0], 3 * n_sample * n_sample, 3), dtype=level_data.dtype, device=level_data.device)
    with torch.cuda.device(level_data.device):
        capi.check(capi.lib().asurf_cubic_extract_iso_pts(capi.ptr(links), capi.size3(links.shape), capi.ptr(level_data),
                                                          capi.ptr(mask_data), capi.ptr(cell_ids), C.c_int64(cell_ids.shape[0]),
                                                          C.c_int32(n_sample), C.c_float(density_thresh), capi.ptr(out),
                                                          capi.current_stream()), "cubic_extract_iso_pts")
    return out


# ---- optimizer steps (optim_kernel.cu:154-267) -------------------------------------------------------------------------
def _indexer(indexer):
    """-> (kind, pointer, n): 0 all rows (0-dim tensor), bool mask, int64 row list; n == 0 means skip."""
    _check_input(indexer, "indexer")
    if indexer.dim() == 0:
        return 0, None, -1
    if indexer.shape[0] == 0:
        return 1, None, 0
    if indexer.dtype == torch.bool:
        return 1, capi.ptr(indexer), int(indexer.shape[0])
    if indexer.dtype != torch.int64:
        raise RuntimeError("indexer must be a bool mask or an int64 index list")
    return 2, capi.ptr(indexer), int(indexer.shape[0])


def rmsprop_step(data, rms, grad, indexer, beta, lr, epsilon, minval, lr_last):
    _check_input(data, "data")
    _check_input(rms, "rms")
    _check_input(grad, "grad")
    kind, p, n = _indexer(indexer)
    if n == 0:
        return
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_rmsprop_step(
            capi.ptr(data), capi.ptr(rms), capi.ptr(grad), C.c_int64(data.shape[0]), C.c_int32(data.shape[1]),
            C.c_int32(kind), p, C.c_int64(max(n, 0)), C.c_float(beta), C.c_float(lr), C.c_float(epsilon),
            C.c_float(minval), C.c_float(lr_last), capi.current_stream()), "rmsprop_step")


def sgd_step(data, grad, indexer, lr, lr_last):
    _check_input(data, "data")
    _check_input(grad, "grad")
    kind, p, n = _indexer(indexer)
    if n == 0:
        return
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_sgd_step(
            capi.ptr(data), capi.ptr(grad), C.c_int64(data.shape[0]), C.c_int32(data.shape[1]), C.c_int32(kind), p,
            C.c_int64(max(n, 0)), C.c_float(lr), C.c_float(lr_last), capi.current_stream()), "sgd_step")


# ---- grid-side regularisers (loss_kernel.cu:1214-1622) ------------------------------------------------------------------
def _mask_ptr(mask_out):
    """The reference passes mask_out.data_ptr() when dim() > 0 (:1368); an empty tensor means "no mask"."""
    _check_input(mask_out, "mask_out")
    return capi.ptr(mask_out) if (mask_out.dim() > 0 and mask_out.numel() > 0) else None


def _check_loss_common(links, data, grad_data=None):
    _check_input(data, "data")
    _check_input(links, "links")
    if grad_data is not None:
        _check_input(grad_data, "grad_data")
        if not grad_data.is_floating_point() or grad_data.dim() != 2:
            raise RuntimeError("grad_data must be a 2-D floating point tensor")
    if not data.is_floating_point() or links.is_floating_point() or data.dim() != 2 or links.dim() != 3:
        raise RuntimeError("data must be a 2-D floating point tensor and links a 3-D integer tensor")
    _check_f32(data, "data")
    if links.dtype != torch.int32:
        raise RuntimeError("links must be int32")


def _check_cells(rand_cells):
    _check_input(rand_cells, "rand_cells")
    if rand_cells.dtype != torch.int32:
        raise RuntimeError("rand_cells must be int32")


def tv(links, data, start_dim, end_dim, use_logalpha, logalpha_delta, ignore_edge, ndc_coeffx, ndc_coeffy):
    _check_loss_common(links, data)
    out = torch.zeros((), dtype=data.dtype, device=data.device)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_tv(capi.ptr(links), capi.size3(links.shape), capi.ptr(data), C.c_int32(data.shape[1]),
                                       C.c_int32(start_dim), C.c_int32(end_dim), C.c_int32(bool(ignore_edge)),
                                       capi.ptr(out), capi.current_stream()), "tv")
    return out


def tv_grad(links, data, start_dim, end_dim, scale, use_logalpha, logalpha_delta, ignore_edge, ndc_coeffx, ndc_coeffy,
            grad_data):
    _check_loss_common(links, data, grad_data)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_tv_grad(capi.ptr(links), capi.size3(links.shape), capi.ptr(data),
                                            C.c_int32(data.shape[1]), C.c_int32(start_dim), C.c_int32(end_dim),
                                            C.c_float(scale), C.c_int32(bool(ignore_edge)), capi.ptr(grad_data),
                                            capi.current_stream()), "tv_grad")


def tv_grad_sparse(links, data, rand_cells, mask_out, start_dim, end_dim, scale, use_logalpha, logalpha_delta,
                   ignore_edge, ignore_last_z, ndc_coeffx, ndc_coeffy, grad_data):
    _check_loss_common(links, data, grad_data)
    _check_cells(rand_cells)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_tv_grad_sparse(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(data), C.c_int32(data.shape[1]), capi.ptr(rand_cells),
            C.c_int64(rand_cells.shape[0]), _mask_ptr(mask_out), C.c_int32(start_dim), C.c_int32(end_dim), C.c_float(scale),
            C.c_int32(bool(ignore_edge)), C.c_int32(bool(ignore_last_z)), capi.ptr(grad_data), capi.current_stream()),
            "tv_grad_sparse")


def surf_tv_grad_sparse(links, data, density_data, rand_cells, mask_out, start_dim, end_dim, scale, ignore_edge,
                        edge_value, ignore_last_z, ndc_coeffx, ndc_coeffy, alpha_dependency, grad_data):
    _check_loss_common(links, data, grad_data)
    _check_input(density_data, "density_data")
    _check_cells(rand_cells)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_surf_tv_grad_sparse(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(data), C.c_int32(data.shape[1]), capi.ptr(density_data),
            C.c_int32(density_data.shape[1]), capi.ptr(rand_cells), C.c_int64(rand_cells.shape[0]), _mask_ptr(mask_out),
            C.c_int32(start_dim), C.c_int32(end_dim), C.c_float(scale), C.c_int32(bool(ignore_edge)), C.c_float(edge_value),
            C.c_int32(bool(ignore_last_z)), C.c_int32(bool(alpha_dependency)), capi.ptr(grad_data),
            capi.ptr(accel_for(links)), capi.current_stream()), "surf_tv_grad_sparse")


def surf_sign_change_grad_sparse(links, data, rand_cells, mask_out, start_dim, end_dim, scale, grad_data):
    """constant-gradient penalty where a stored vertex and a +x / +y / +z neighbour differ in sign (loss_kernel.cu:895-977)"""
    _check_loss_common(links, data, grad_data)
    _check_cells(rand_cells)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_surf_sign_change_grad_sparse(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(data), C.c_int32(data.shape[1]), capi.ptr(rand_cells),
            C.c_int64(rand_cells.shape[0]), _mask_ptr(mask_out), C.c_int32(start_dim), C.c_int32(end_dim), C.c_float(scale),
            capi.ptr(grad_data), capi.current_stream()), "surf_sign_change_grad_sparse")


def surface_normal_grad(links, data, lv_set, start_dim, end_dim, scale, ndc_coeffx, ndc_coeffy, grad_data):
    """dense normal-consistency loss over the whole lattice (loss_kernel.cu:1289-1325; the ndc coefficients are unused there too)"""
    _check_loss_common(links, data, grad_data)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_surface_normal_grad(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(data), C.c_int32(data.shape[1]), C.c_float(lv_set),
            C.c_int32(start_dim), C.c_int32(end_dim), C.c_float(scale), capi.ptr(grad_data), capi.current_stream()),
            "surface_normal_grad")


def lumisphere_tv_grad_sparse(grid, rand_cells, basis_fn, basis_fn_u, scale, ndc_coeffx, ndc_coeffy, dir_factor, grads):
    """TV of the radiance seen from one direction (loss_kernel.cu:1661-1697); basis_fn must be 1-D (:1674), basis_fn_u is read
    as its first basis_dim values like the reference kernel does (:1127)"""
    _check_input(grid.sh_data, "grid.sh_data")
    _check_input(grid.links, "grid.links")
    _check_cells(rand_cells)
    _check_input(basis_fn, "basis_fn")
    _check_input(basis_fn_u, "basis_fn_u")
    if basis_fn.dim() != 1:
        raise RuntimeError("basis_fn must be 1-D (loss_kernel.cu:1674)")
    if grads.grad_sh_out is None:
        raise RuntimeError("grads.grad_sh_out is required")
    _check_input(grads.grad_sh_out, "grads.grad_sh_out")
    basis_dim = int(grid.basis_dim)
    if basis_fn.numel() < basis_dim or basis_fn_u.numel() < basis_dim:
        raise RuntimeError("basis_fn / basis_fn_u hold fewer than basis_dim values")
    with torch.cuda.device(grid.sh_data.device):
        capi.check(capi.lib().asurf_lumisphere_tv_grad_sparse(
            capi.ptr(grid.links), capi.size3(grid.links.shape), capi.ptr(grid.sh_data), C.c_int32(grid.sh_data.shape[1]),
            C.c_int32(basis_dim), capi.ptr(rand_cells), C.c_int64(rand_cells.shape[0]), capi.ptr(basis_fn.float()),
            capi.ptr(basis_fn_u.float()), C.c_float(scale), C.c_float(dir_factor),
            None if grads.mask_out is None else _mask_ptr(grads.mask_out),
            capi.ptr(grads.grad_sh_out), capi.current_stream()), "lumisphere_tv_grad_sparse")


def alpha_surf_sparsify_grad_sparse(links, alpha_data, surf_data, rand_cells, mask_out, scale_alpha, scale_surf,
                                    surf_sparse_decrease, surf_sparse_thresh, alpha_bound, surf_bound, grad_alpha, grad_surf):
    _check_loss_common(links, alpha_data, grad_alpha)
    _check_input(surf_data, "surf_data")
    _check_input(grad_surf, "grad_surf")
    _check_cells(rand_cells)
    with torch.cuda.device(alpha_data.device):
        capi.check(capi.lib().asurf_alpha_surf_sparsify_grad_sparse(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(alpha_data), C.c_int32(alpha_data.shape[1]),
            capi.ptr(surf_data), C.c_int32(surf_data.shape[1]), capi.ptr(rand_cells), C.c_int64(rand_cells.shape[0]),
            _mask_ptr(mask_out), C.c_float(scale_alpha), C.c_float(scale_surf), C.c_int32(bool(surf_sparse_decrease)),
            C.c_float(surf_sparse_thresh), C.c_float(alpha_bound), C.c_float(surf_bound), capi.ptr(grad_alpha),
            capi.ptr(grad_surf), capi.current_stream()), "alpha_surf_sparsify_grad_sparse")


def surface_normal_grad_sparse(links, data, rand_cells, mask_out, lv_set, start_dim, end_dim, scale, eikonal_scale,
                               ndc_coeffx, ndc_coeffy, con_check, ignore_empty, use_l1, grad_data):
    _check_loss_common(links, data, grad_data)
    _check_cells(rand_cells)
    with torch.cuda.device(data.device):
        capi.check(capi.lib().asurf_surface_normal_grad_sparse(
            capi.ptr(links), capi.size3(links.shape), capi.ptr(data), capi.ptr(rand_cells), C.c_int64(rand_cells.shape[0]),
            _mask_ptr(mask_out), C.c_float(lv_set), C.c_int32(start_dim), C.c_int32(end_dim), C.c_float(scale),
            C.c_int32(bool(con_check)), C.c_int32(bool(ignore_empty)), C.c_int32(bool(use_l1)), capi.ptr(grad_data),
            capi.ptr(accel_for(links)), capi.current_stream()), "surface_normal_grad_sparse")


def msi_tv_grad_sparse(links, msi, rand_cells, mask_out, scale, scale_last, grad_msi):
    """TV over (texel, layer) cells of the MSI background (loss_kernel.cu:979-1064, :1624-1659)"""
    for t, n in ((links, "links"), (msi, "msi"), (grad_msi, "grad_msi"), (mask_out, "mask_out")):
        _check_input(t, n)
    _check_cells(rand_cells)
    if msi.dim() != 3 or links.dim() != 2 or not msi.is_floating_point():
        raise RuntimeError("msi must be a (n, nlayers, channels) floating point tensor and links 2-D")
    with torch.cuda.device(msi.device):
        capi.check(capi.lib().asurf_msi_tv_grad_sparse(
            capi.ptr(links), C.c_int32(links.shape[0]), C.c_int32(links.shape[1]), capi.ptr(msi), C.c_int32(msi.shape[1]),
            C.c_int32(msi.shape[2]), capi.ptr(rand_cells), C.c_int64(rand_cells.shape[0]),
            capi.ptr(mask_out) if mask_out.numel() > 0 else None, C.c_float(scale), C.c_float(scale_last), capi.ptr(grad_msi),
            capi.current_stream()), "msi_tv_grad_sparse")


def debug_bg_state(n_rays, device="cuda"):
    """(log_transmit (Q,), accum (Q,)) the last foreground pass over a grid with a background left for the MSI pass"""
    lt = torch.empty((n_rays,), dtype=torch.float32, device=device)
    acc = torch.empty((n_rays,), dtype=torch.float32, device=device)
    capi.check(capi.lib().asurf_debug_bg_state(capi.ptr(lt), capi.ptr(acc), C.c_int64(n_rays)), "debug_bg_state")
    return lt, acc


# ---- test hooks -----------------------------------------------------------------------------------------------------------
def debug_ray_bounds(grid, rays, opt):
    """(Q,9) grid-space rays as the kernels see them: origin3, dir3, tmin, tmax, world_step."""
    _check_grid(grid)
    _check_rays(rays)
    xf = torch.zeros((rays.origins.shape[0], 9), dtype=torch.float32, device=rays.origins.device)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_debug_ray_bounds(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                     capi.ptr(xf), capi.current_stream()), "debug_ray_bounds")
    return xf


def debug_trace(grid, rays, opt, max_hits=64):
    """Composited-sample trace of the forward march: (count (Q,), cell (Q,H), kind (Q,H), t (Q,H))."""
    _check_grid(grid)
    _check_rays(rays)
    Q, dev = rays.origins.shape[0], rays.origins.device
    cnt = torch.zeros((Q,), dtype=torch.int32, device=dev)
    cell = torch.full((Q, max_hits), -1, dtype=torch.int32, device=dev)
    kind = torch.full((Q, max_hits), -1, dtype=torch.int32, device=dev)
    t = torch.zeros((Q, max_hits), dtype=torch.float32, device=dev)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_debug_trace(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                C.c_int32(max_hits), capi.ptr(cnt), capi.ptr(cell), capi.ptr(kind),
                                                capi.ptr(t), capi.current_stream()), "debug_trace")
    return cnt, cell, kind, t


def render_stats(grid, rays, opt):
    """Counters of SURVEY.md 8(d) for one forward march: dict(n_steps, n_linked, n_active, n_samples)."""
    _check_grid(grid)
    _check_rays(rays)
    st = torch.zeros((6,), dtype=torch.int64, device=rays.origins.device)
    out = torch.empty_like(rays.origins)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid)
        capi.check(capi.lib().asurf_surf_trav_forward(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)),
                                                      capi.ptr(out), capi.ptr(st), capi.current_stream()),
                   "render_stats")
    v = st.tolist()
    return dict(n_steps=v[0], n_skips=v[1], n_linked=v[2], n_active=v[3], n_samples=v[4])


def cuvol_render_stats(grid, rays, opt):
    """Counters of SURVEY.md 8(d) for one cuvol forward march: dict(n_steps = sample positions, n_skips, n_linked = gathered
    samples, n_samples = samples with sigma > sigma_thresh)."""
    _check_grid(grid)
    _check_rays(rays)
    st = torch.zeros((6,), dtype=torch.int64, device=rays.origins.device)
    with torch.cuda.device(grid.sh_data.device):
        g, _keep = _grid_t(grid, need_accel=False)
        capi.check(capi.lib().asurf_cuvol_stats(C.byref(g), C.byref(_rays_t(rays)), C.byref(capi.make_opt(opt)), capi.ptr(st),
                                                capi.current_stream()), "cuvol_render_stats")
    v = st.tolist()
    return dict(n_steps=v[0], n_skips=v[1], n_linked=v[2], n_active=v[3], n_samples=v[4])


# ---- entry points of svox2.cpp that are outside the hot path ----------------------------------
# They exist so that `hasattr(_C, name)` probes of the reference behave (svox2/utils.py:36 requires `sample_grid`), and
# raise instead of silently doing nothing: there is no fallback implementation in this package.
def _not_on_hot_path(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("svox2.csrc.%s is outside the B200 hot path of alphasurf_b200 (SURVEY.md 8f)" % name)
    fn.__name__ = name
    return fn


for _name in ("volume_render_surface", "volume_render_surface_backward", "volume_render_surface_fused",
              "volume_render_nvol", "volume_render_nvol_backward", "volume_render_nvol_fused", "volume_render_svox1",
              "volume_render_svox1_backward", "volume_render_svox1_fused", "test_cubic_root_grad"):
    if _name not in globals():
        globals()[_name] = _not_on_hot_path(_name)
del _name
