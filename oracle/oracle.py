"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/*.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module.  The product package ``alphasurf_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_cubic_solve.restype = C.c_int
    return _LIB


class OGrid(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("links", C.c_void_p), ("density", C.c_void_p), ("surface", C.c_void_p),
                ("sh", C.c_void_p), ("level_set", C.c_void_p), ("level_set_num", C.c_int32),
                ("basis_dim", C.c_int32), ("sh_dim", C.c_int32), ("offset", C.c_float * 3),
                ("scaling", C.c_float * 3), ("fake_sample_std", C.c_float), ("truncated_vol_render_a", C.c_float)]


class OOpt(C.Structure):
    _fields_ = [("background_brightness", C.c_float), ("step_size", C.c_float), ("sigma_thresh", C.c_float),
                ("stop_thresh", C.c_float), ("near_clip", C.c_float), ("use_spheric_clip", C.c_int32),
                ("last_sample_opaque", C.c_int32), ("surf_fake_sample", C.c_int32),
                ("surf_fake_sample_min_vox_len", C.c_float), ("limited_fake_sample", C.c_int32),
                ("no_surf_grad_from_sh", C.c_int32), ("alpha_activation_type", C.c_int32),
                ("fake_sample_l_dist", C.c_int32), ("fake_sample_normalize_surf", C.c_int32),
                ("only_outward_intersect", C.c_int32), ("truncated_vol_render", C.c_int32),
                ("trunc_vol_weight_min", C.c_float)]


class OFused(C.Structure):
    _fields_ = [("beta_loss", C.c_float), ("sparsity_loss", C.c_float), ("lambda_l2", C.c_float),
                ("lambda_l1", C.c_float), ("lambda_l_dist", C.c_float), ("lambda_l_entropy", C.c_float),
                ("no_norm_weight_l_entropy", C.c_int32), ("lambda_l_dist_a", C.c_float),
                ("lambda_l_entropy_a", C.c_float), ("lambda_l_samp_dist", C.c_float), ("lambda_l_di", C.c_float),
                ("l_di_alpha_thresh", C.c_float), ("surf_sparse_alpha_thresh", C.c_float),
                ("lambda_inplace_surf_sparse", C.c_float), ("lambda_inwards_norm_loss", C.c_float),
                ("lambda_conv_mode_samp", C.c_float), ("l_dist_max_sample", C.c_int32)]


class OGrads(C.Structure):
    _fields_ = [("grad_density", C.c_void_p), ("grad_surface", C.c_void_p), ("grad_sh", C.c_void_p),
                ("grad_fake_sample_std", C.c_void_p), ("mask", C.c_void_p)]


class OTrace(C.Structure):
    _fields_ = [("max_hits", C.c_int32), ("hit_count", C.c_void_p), ("hit_cell", C.c_void_p),
                ("hit_kind", C.c_void_p), ("hit_t", C.c_void_p), ("counters", C.c_void_p)]


def _np(t, dtype):
    """torch tensor / ndarray -> contiguous ndarray of dtype (copy only if needed)."""
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Grid:
    """Holds numpy copies of a SynthGrid-like object (attributes links/density/surface/sh/level_set/offset/scaling)."""

    def __init__(self, g):
        self.links = _np(g.links, np.int32)
        self.density = _np(g.density, np.float32)
        self.surface = _np(g.surface, np.float32)
        self.sh = _np(g.sh, np.float32)
        self.level_set = _np(g.level_set, np.float32)
        self.offset = _np(g.offset, np.float32)
        self.scaling = _np(g.scaling, np.float32)
        self.basis_dim = int(g.basis_dim)
        self.fake_sample_std = float(g.fake_sample_std)
        self.truncated_vol_render_a = float(g.truncated_vol_render_a)
        s = OGrid()
        s.size[:] = list(self.links.shape)
        s.links = _ptr(self.links)
        s.density = _ptr(self.density)
        s.surface = _ptr(self.surface)
        s.sh = _ptr(self.sh)
        s.level_set = _ptr(self.level_set)
        s.level_set_num = 0 if self.level_set is None else int(self.level_set.shape[0])
        s.basis_dim = self.basis_dim
        s.sh_dim = int(self.sh.shape[1])
        s.offset[:] = self.offset.tolist()
        s.scaling[:] = self.scaling.tolist()
        s.fake_sample_std = self.fake_sample_std
        s.truncated_vol_render_a = self.truncated_vol_render_a
        self.c = s

    @property
    def N(self):
        return self.density.shape[0]


def make_opt(d: dict) -> OOpt:
    o = OOpt()
    for name, ctype in OOpt._fields_:
        v = d[name]
        setattr(o, name, float(v) if ctype is C.c_float else int(v))
    return o


def make_fused(d: dict) -> OFused:
    f = OFused()
    for name, ctype in OFused._fields_:
        v = d.get(name, 0)
        setattr(f, name, float(v) if ctype is C.c_float else int(v))
    return f


class Grads:
    def __init__(self, grid: Grid, with_mask=True, with_std=True):
        self.density = np.zeros_like(grid.density)
        self.surface = None if grid.surface is None else np.zeros_like(grid.surface)
        self.sh = np.zeros_like(grid.sh)
        self.fake_sample_std = np.zeros((1,), np.float32) if with_std else None
        self.mask = np.zeros((grid.N,), np.uint8) if with_mask else None
        c = OGrads()
        c.grad_density = _ptr(self.density)
        c.grad_surface = _ptr(self.surface)
        c.grad_sh = _ptr(self.sh)
        c.grad_fake_sample_std = _ptr(self.fake_sample_std)
        c.mask = _ptr(self.mask)
        self.c = c


class Trace:
    def __init__(self, Q, max_hits=64):
        self.hit_count = np.zeros((Q,), np.int32)
        self.hit_cell = np.full((Q, max_hits), -1, np.int32)
        self.hit_kind = np.full((Q, max_hits), -1, np.int32)
        self.hit_t = np.zeros((Q, max_hits), np.float32)
        self.counters = np.zeros((Q, 4), np.int64)
        c = OTrace()
        c.max_hits = max_hits
        c.hit_count = _ptr(self.hit_count)
        c.hit_cell = _ptr(self.hit_cell)
        c.hit_kind = _ptr(self.hit_kind)
        c.hit_t = _ptr(self.hit_t)
        c.counters = _ptr(self.counters)
        self.c = c


def surf_trav_forward(grid: Grid, opt: dict, origins, dirs, xf=None, trace: Trace = None):
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    xf = _np(xf, np.float32)
    Q = o.shape[0]
    out = np.zeros((Q, 3), np.float32)
    lib().oracle_surf_trav_forward(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf),
                                   C.c_int64(Q), _ptr(out), C.c_int(0), None, None, None,
                                   C.byref(trace.c) if trace else None)
    return out


SCALAR_MODES = dict(expected_term=0, mode_term=1, thresh_depth=2, thresh_alpha=3, normal=4, extract_pts=5)


def surf_trav_scalar(grid: Grid, opt: dict, origins, dirs, mode, param=0.0, xf=None, max_sample=0):
    """depth / alpha / normal / point renders (render_lerp_kernel_surf_trav.cu:564-1708); mode: key of SCALAR_MODES;
    extract_pts returns (depths, alphas), each (Q, max_sample)"""
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    xf = _np(xf, np.float32)
    Q = o.shape[0]
    m = SCALAR_MODES[mode]
    out = np.zeros((Q, 3) if m == 4 else ((Q, max_sample) if m == 5 else (Q,)), np.float32)
    out2 = np.zeros_like(out) if m == 5 else None
    lib().oracle_surf_trav_scalar(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(Q),
                                  C.c_int(m), C.c_float(param), C.c_int(max_sample), _ptr(out), _ptr(out2))
    return (out, out2) if m == 5 else out


def surf_trav_backward(grid: Grid, opt: dict, origins, dirs, grad_out, color_cache, xf=None, grads: Grads = None):
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    xf = _np(xf, np.float32)
    go, cc = _np(grad_out, np.float32), _np(color_cache, np.float32)
    grads = grads or Grads(grid, with_mask=False)
    lib().oracle_surf_trav_backward(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf),
                                    C.c_int64(o.shape[0]), _ptr(go), _ptr(cc), C.byref(grads.c))
    return grads


def surf_trav_fused(grid: Grid, opt: dict, origins, dirs, rgb_gt, fused: dict, xf=None, grads: Grads = None,
                    trace: Trace = None, q_norm=None):
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    xf = _np(xf, np.float32)
    gt = _np(rgb_gt, np.float32)
    Q = o.shape[0]
    out = np.zeros((Q, 3), np.float32)
    grads = grads or Grads(grid)
    lib().oracle_surf_trav_fused(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(Q),
                                 C.c_int64(Q if q_norm is None else q_norm), _ptr(gt), C.byref(make_fused(fused)),
                                 _ptr(out), C.byref(grads.c), C.byref(trace.c) if trace else None)
    return out, grads


def cubic_solve(fs):
    fs = np.ascontiguousarray(fs, np.float64)
    st = np.zeros(3, np.float64)
    typ = lib().oracle_cubic_solve(_ptr(fs), _ptr(st))
    return typ, st


def cubic_root_grad(typ, st_id, fs):
    fs = np.ascontiguousarray(fs, np.float64)
    g = np.ones(4, np.float32)
    lib().oracle_cubic_root_grad(C.c_int(typ), C.c_int(st_id), _ptr(fs), _ptr(g))
    return g


# ---- optimizer steps: numpy restatement of optim_kernel.cu:15-25 (rmsprop_once) and :97-103 (sgd_once) ----
def _rows(n_rows, indexer):
    """indexer as the reference dispatches it (optim_kernel.cu:175-215): None = all rows, bool mask, int64 list."""
    if indexer is None:
        return np.arange(n_rows)
    indexer = np.asarray(indexer)
    if indexer.dtype == np.bool_:
        return np.nonzero(indexer)[0]
    return indexer.astype(np.int64)


def rmsprop_step(data, rms, grad, indexer, beta, lr, eps, minval, lr_last):
    """In place on float32 arrays (N, C).  fmaf is emulated in float64 (exact product, one final rounding)."""
    if lr_last < 0:
        lr_last = lr
    rows = _rows(data.shape[0], indexer)
    f32 = np.float32
    g = grad[rows]
    g2 = (g * g).astype(f32)
    r = rms[rows]
    lerp = (np.float64(f32(beta)) * (r - g2).astype(f32).astype(np.float64) + g2.astype(np.float64)).astype(f32)
    r_new = np.where(r == 0, g2, lerp).astype(f32)
    lrs = np.full((data.shape[1],), f32(lr), f32)
    lrs[-1] = f32(lr_last)
    step = ((lrs[None, :] * g).astype(f32) / (np.sqrt(r_new).astype(f32) + f32(eps)).astype(f32)).astype(f32)
    data[rows] = np.maximum((data[rows] - step).astype(f32), f32(minval))
    rms[rows] = r_new
    grad[rows] = 0


def sgd_step(data, grad, indexer, lr, lr_last):
    if lr_last < 0:
        lr_last = lr
    rows = _rows(data.shape[0], indexer)
    lrs = np.full((data.shape[1],), np.float32(lr), np.float32)
    lrs[-1] = np.float32(lr_last)
    d = data[rows].astype(np.float64) - lrs[None, :].astype(np.float64) * grad[rows].astype(np.float64)
    data[rows] = d.astype(np.float32)
    grad[rows] = 0


# ---- grid regularisers (oracle_loss.c) --------------------------------------------------------------------------------
def _sz(links):
    return (C.c_int32 * 3)(*[int(s) for s in links.shape])


def tv(links, data, start_dim, end_dim, ignore_edge):
    links, data = _np(links, np.int32), _np(data, np.float32)
    lib().oracle_tv.restype = C.c_float
    return float(lib().oracle_tv(_ptr(links), _sz(links), _ptr(data), C.c_int(data.shape[1]), C.c_int(start_dim),
                                 C.c_int(end_dim), C.c_int(int(ignore_edge))))


def tv_grad(links, data, start_dim, end_dim, scale, ignore_edge, grad):
    links, data = _np(links, np.int32), _np(data, np.float32)
    lib().oracle_tv_grad(_ptr(links), _sz(links), _ptr(data), C.c_int(data.shape[1]), C.c_int(start_dim), C.c_int(end_dim),
                         C.c_float(scale), C.c_int(int(ignore_edge)), _ptr(grad))


def tv_grad_sparse(links, data, density, cells, mask, start_dim, end_dim, scale, ignore_edge, edge_value, ignore_last_z,
                   alpha_dependency, surf, grad):
    links, data, cells = _np(links, np.int32), _np(data, np.float32), _np(cells, np.int32)
    density = _np(density, np.float32)
    lib().oracle_tv_grad_sparse(_ptr(links), _sz(links), _ptr(data), C.c_int(data.shape[1]), _ptr(density),
                                C.c_int(0 if density is None else density.shape[1]), _ptr(cells),
                                C.c_int64(cells.shape[0]), _ptr(mask), C.c_int(start_dim), C.c_int(end_dim),
                                C.c_float(scale), C.c_int(int(ignore_edge)), C.c_float(edge_value),
                                C.c_int(int(ignore_last_z)), C.c_int(int(alpha_dependency)), C.c_int(int(surf)), _ptr(grad))


def alpha_surf_sparsify(links, alpha, surf, cells, mask, scale_alpha, scale_surf, surf_decrease, surf_thresh, alpha_bound,
                        surf_bound, grad_alpha, grad_surf):
    links, alpha, surf, cells = _np(links, np.int32), _np(alpha, np.float32), _np(surf, np.float32), _np(cells, np.int32)
    lib().oracle_alpha_surf_sparsify(_ptr(links), _ptr(alpha), C.c_int(alpha.shape[1]), _ptr(surf), C.c_int(surf.shape[1]),
                                     _ptr(cells), C.c_int64(cells.shape[0]), _ptr(mask), C.c_float(scale_alpha),
                                     C.c_float(scale_surf), C.c_int(int(surf_decrease)), C.c_float(surf_thresh),
                                     C.c_float(alpha_bound), C.c_float(surf_bound), _ptr(grad_alpha), _ptr(grad_surf))


def surface_normal_grad_sparse(links, surf, cells, mask, lv_set, start_dim, end_dim, scale, con_check, ignore_empty, use_l1,
                               grad):
    links, surf, cells = _np(links, np.int32), _np(surf, np.float32), _np(cells, np.int32)
    lib().oracle_surface_normal_grad_sparse(_ptr(links), _sz(links), _ptr(surf), _ptr(cells), C.c_int64(cells.shape[0]),
                                            _ptr(mask), C.c_float(lv_set), C.c_int(start_dim), C.c_int(end_dim),
                                            C.c_float(scale), C.c_int(int(con_check)), C.c_int(int(ignore_empty)),
                                            C.c_int(int(use_l1)), _ptr(grad))


# ---- Plenoxels cuvol renderer (oracle_cuvol.c) ---------------------------------------------------------------------------
def cuvol_forward(grid: Grid, opt: dict, origins, dirs, xf=None, want_log_transmit=False):
    o, d, xf = _np(origins, np.float32), _np(dirs, np.float32), _np(xf, np.float32)
    Q = o.shape[0]
    out = np.zeros((Q, 3), np.float32)
    lt = np.zeros((Q,), np.float32)
    lib().oracle_cuvol_forward(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(Q), _ptr(out),
                               _ptr(lt))
    return (out, lt) if want_log_transmit else out


CUVOL_SCALAR_MODES = dict(expected_term=0, mode_term=1, med_term=2, sigma_thresh=3)


def cuvol_scalar(grid: Grid, opt: dict, origins, dirs, mode, param=0.0, max_sample=0, xf=None):
    """depth renders of the cuvol backend (render_lerp_kernel_cuvol.cu:127-369); med_term returns (depths, sigmas)"""
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    xf = _np(xf, np.float32)
    Q = o.shape[0]
    m = CUVOL_SCALAR_MODES[mode]
    shape = (Q, max_sample) if m == 2 else (Q,)
    out, out2 = np.zeros(shape, np.float32), np.zeros(shape if m == 2 else (1,), np.float32)
    lib().oracle_cuvol_scalar(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(Q), C.c_int(m),
                              C.c_float(param), C.c_int(max_sample), _ptr(out), _ptr(out2))
    return (out, out2) if m == 2 else out


def cuvol_backward(grid: Grid, opt: dict, origins, dirs, grad_out, color_cache, xf=None, grads: Grads = None):
    o, d, xf = _np(origins, np.float32), _np(dirs, np.float32), _np(xf, np.float32)
    go, cc = _np(grad_out, np.float32), _np(color_cache, np.float32)
    grads = grads or Grads(grid, with_std=False)
    lib().oracle_cuvol_backward(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(o.shape[0]),
                                _ptr(go), _ptr(cc), C.c_int(0), C.c_float(0.0), None, C.c_float(0.0), C.c_float(0.0),
                                C.byref(grads.c))
    return grads


def cuvol_fused(grid: Grid, opt: dict, origins, dirs, rgb_gt, beta_loss=0.0, sparsity_loss=0.0, xf=None, grads: Grads = None,
                q_norm=None):
    """volume_render_cuvol_fused (render_lerp_kernel_cuvol.cu:1272-1354): forward, then backward of the MSE."""
    o, d, xf = _np(origins, np.float32), _np(dirs, np.float32), _np(xf, np.float32)
    gt = _np(rgb_gt, np.float32)
    Q = o.shape[0]
    qn = Q if q_norm is None else int(q_norm)
    out, lt = cuvol_forward(grid, opt, o, d, xf=xf, want_log_transmit=True)
    grads = grads or Grads(grid, with_std=False)
    norm = np.float32(2.0) / np.float32(3 * qn)
    lib().oracle_cuvol_backward(C.byref(grid.c), C.byref(make_opt(opt)), _ptr(o), _ptr(d), _ptr(xf), C.c_int64(Q), _ptr(gt),
                                _ptr(out), C.c_int(1), C.c_float(norm), _ptr(lt) if beta_loss > 0 else None,
                                C.c_float(np.float32(beta_loss) / np.float32(qn)), C.c_float(sparsity_loss), C.byref(grads.c))
    return out, grads


# ---- accel_dist_prop (misc_kernel.cu:113-182, :1022-1058), numpy restatement --------------------------------------------------
def accel_dist_prop(links):
    """Returns a copy of ``links`` (X,Y,Z) int32 with every negative entry replaced by -(1 + number of empty octree levels
    above the vertex).  Level sizes halve with ceil until an axis reaches 1 (misc_kernel.cu:1035-1041)."""
    links = np.array(_np(links, np.int32), copy=True)
    X, Y, Z = links.shape
    occ = links >= 0
    levels = []
    cur, (sx, sy, sz) = occ, (X, Y, Z)
    while sx > 1 and sy > 1 and sz > 1:
        nx, ny, nz = (sx + 1) // 2, (sy + 1) // 2, (sz + 1) // 2
        pad = np.zeros((2 * nx, 2 * ny, 2 * nz), bool)
        pad[:sx, :sy, :sz] = cur
        cur = pad.reshape(nx, 2, ny, 2, nz, 2).any(axis=(1, 3, 5))
        levels.append(cur)
        sx, sy, sz = nx, ny, nz
    xs, ys, zs = np.nonzero(~occ)
    result = np.full(xs.shape, -1, np.int32)
    alive = np.ones(xs.shape, bool)
    x, y, z = xs.copy(), ys.copy(), zs.copy()
    for lv in levels:
        x >>= 1
        y >>= 1
        z >>= 1
        hit = lv[x, y, z]
        alive &= ~hit
        result[alive] -= 1
    links[xs, ys, zs] = result
    return links


# ---- grid maintenance renders (oracle_gridtools.c, oracle_surf_trav.c) ---------------------------------------------------
def cam_rays(c2w, fx, fy, cx, cy, width, height):
    """cam2world_ray for every pixel in raster order -> (origins, dirs), each (H*W, 3)"""
    c = np.ascontiguousarray(np.asarray(c2w, np.float32)[:3, :4].reshape(-1))
    o = np.zeros((height * width, 3), np.float32)
    d = np.zeros((height * width, 3), np.float32)
    lib().oracle_cam_rays(_ptr(c), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(width), C.c_int(height),
                          _ptr(o), _ptr(d))
    return o, d


def dilate(grid_bool):
    g = np.ascontiguousarray(_np(grid_bool, np.uint8))
    out = np.zeros_like(g)
    lib().oracle_dilate(_ptr(g), _ptr(np.asarray(g.shape, np.int32)), _ptr(out))
    return out.astype(bool)


def weight_render(data, links, size, offset, scaling, origins, dirs, step_size, stop_thresh, last_sample_opaque, out):
    """links None: dense (X,Y,Z) volume (grid_weight_render); else sparse (sparse_grid_weight_render).  out (X,Y,Z) max-updated."""
    dat = _np(data, np.float32)
    lk = None if links is None else _np(links, np.int32)
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    lib().oracle_weight_render(C.c_int(0 if lk is None else 1), _ptr(dat), _ptr(lk), _ptr(np.asarray(size, np.int32)),
                               _ptr(_np(offset, np.float32)), _ptr(_np(scaling, np.float32)), _ptr(o), _ptr(d), None,
                               C.c_int64(o.shape[0]), C.c_float(step_size), C.c_float(stop_thresh),
                               C.c_int(1 if last_sample_opaque else 0), _ptr(out))
    return out


def mask_render(links, offset, scaling, origins, dirs, near_clip, out):
    lk = _np(links, np.int32)
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    lib().oracle_mask_render(_ptr(lk), _ptr(np.asarray(lk.shape, np.int32)), _ptr(_np(offset, np.float32)),
                             _ptr(_np(scaling, np.float32)), _ptr(o), _ptr(d), None, C.c_int64(o.shape[0]), C.c_float(near_clip),
                             _ptr(out))
    return out


def visibility_surf(grid: Grid, origins, dirs, out):
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    lib().oracle_visibility_surf(C.byref(grid.c), _ptr(o), _ptr(d), None, C.c_int64(o.shape[0]), _ptr(out))
    return out


# ---- point queries (svox2_kernel.cu:11-376) -------------------------------------------------------------------------------
def sample_grid(links, offset, scaling, data, missing, points):
    lk, dat, pts = _np(links, np.int32), _np(data, np.float32), _np(points, np.float32)
    out = np.zeros((pts.shape[0], dat.shape[1]), np.float32)
    lib().oracle_sample_grid(_ptr(lk), _ptr(np.asarray(lk.shape, np.int32)), _ptr(_np(offset, np.float32)),
                             _ptr(_np(scaling, np.float32)), _ptr(dat), C.c_int(dat.shape[1]), C.c_float(missing), _ptr(pts),
                             C.c_int64(pts.shape[0]), _ptr(out))
    return out


def sample_grid_backward(links, offset, scaling, points, grad_out, grad_data):
    lk, pts, go = _np(links, np.int32), _np(points, np.float32), _np(grad_out, np.float32)
    lib().oracle_sample_grid_backward(_ptr(lk), _ptr(np.asarray(lk.shape, np.int32)), _ptr(_np(offset, np.float32)),
                                      _ptr(_np(scaling, np.float32)), _ptr(pts), C.c_int64(pts.shape[0]), _ptr(go),
                                      C.c_int(go.shape[1]), _ptr(grad_data))
    return grad_data


def cubic_extract_iso_pts(links, level, maskv, cell_ids, n_sample, density_thresh):
    lk, lv, mv, ids = _np(links, np.int32), _np(level, np.float32), _np(maskv, np.float32), _np(cell_ids, np.int32)
    out = np.zeros((ids.shape[0], 3 * n_sample * n_sample, 3), np.float32)
    lib().oracle_cubic_extract_iso_pts(_ptr(lk), _ptr(np.asarray(lk.shape, np.int32)), _ptr(lv), _ptr(mv), _ptr(ids),
                                       C.c_int64(ids.shape[0]), C.c_int(n_sample), C.c_float(density_thresh), _ptr(out))
    return out


# ---- MSI background (oracle_msi.c) ---------------------------------------------------------------------------------------------
def _f3(t):
    a = np.ascontiguousarray(np.asarray(t, dtype=np.float32).reshape(-1)[:3])
    return a


def msi_forward(bg_links, bg_data, size, offset, scaling, opt: dict, origins, dirs, log_transmit, rgb):
    """rgb (Q,3) float32 numpy, updated in place: += background colours (render_background_kernel)"""
    links, data = _np(bg_links, np.int32), _np(bg_data, np.float32)
    o, d, lt = _np(origins, np.float32), _np(dirs, np.float32), _np(log_transmit, np.float32)
    sz = (C.c_int32 * 3)(*[int(v) for v in size])
    off, sc = _f3(offset), _f3(scaling)
    oo = make_opt(opt)
    lib().oracle_msi_forward(_ptr(links), _ptr(data), C.c_int(links.shape[1]), C.c_int(data.shape[1]), sz, _ptr(off), _ptr(sc),
                             C.byref(oo), _ptr(o), _ptr(d), C.c_int64(o.shape[0]), _ptr(lt), _ptr(rgb))


def msi_backward(bg_links, bg_data, size, offset, scaling, opt: dict, origins, dirs, grad_in, color_cache, grad_is_rgb,
                 log_transmit, accum, sparsity_loss, grad_bg, mask_bg):
    links, data = _np(bg_links, np.int32), _np(bg_data, np.float32)
    o, d = _np(origins, np.float32), _np(dirs, np.float32)
    gi, cc = _np(grad_in, np.float32), _np(color_cache, np.float32)
    lt, acc = _np(log_transmit, np.float32), _np(accum, np.float32)
    sz = (C.c_int32 * 3)(*[int(v) for v in size])
    off, sc = _f3(offset), _f3(scaling)
    oo = make_opt(opt)
    lib().oracle_msi_backward(_ptr(links), _ptr(data), C.c_int(links.shape[1]), C.c_int(data.shape[1]), sz, _ptr(off), _ptr(sc),
                              C.byref(oo), _ptr(o), _ptr(d), C.c_int64(o.shape[0]), _ptr(gi), _ptr(cc), C.c_int(int(grad_is_rgb)),
                              _ptr(lt), _ptr(acc), C.c_float(sparsity_loss), _ptr(grad_bg), _ptr(mask_bg))


def msi_tv_grad_sparse(bg_links, msi, cells, mask, scale, scale_last, grad):
    links, data, cells = _np(bg_links, np.int32), _np(msi, np.float32), _np(cells, np.int32)
    lib().oracle_msi_tv_grad_sparse(_ptr(links), C.c_int(links.shape[0]), C.c_int(links.shape[1]), _ptr(data),
                                    C.c_int(data.shape[1]), C.c_int(data.shape[2]), _ptr(cells), C.c_int64(cells.shape[0]),
                                    _ptr(mask), C.c_float(scale), C.c_float(scale_last), _ptr(grad))


def surf_sign_change_grad_sparse(links, data, cells, mask, start_dim, end_dim, scale, grad):
    links, data, cells = _np(links, np.int32), _np(data, np.float32), _np(cells, np.int32)
    lib().oracle_surf_sign_change_grad_sparse(_ptr(links), _sz(links), _ptr(data), C.c_int(data.shape[1]), _ptr(cells),
                                              C.c_int64(cells.shape[0]), _ptr(mask), C.c_int(start_dim), C.c_int(end_dim),
                                              C.c_float(scale), _ptr(grad))


def surface_normal_grad(links, data, lv_set, start_dim, end_dim, scale, grad):
    """dense normal-consistency loss (loss_kernel.cu:245-396, :1289-1325); grad (N, n_cols) float32 numpy, accumulated"""
    links, data = _np(links, np.int32), _np(data, np.float32)
    lib().oracle_surface_normal_grad(_ptr(links), _sz(links), _ptr(data), C.c_int(data.shape[1]), C.c_float(lv_set),
                                     C.c_int(start_dim), C.c_int(end_dim), C.c_float(scale), _ptr(grad))


def lumisphere_tv_grad_sparse(links, sh, basis_dim, cells, basis_fn, basis_fn_u, scale, dir_factor, mask, grad):
    """loss_kernel.cu:1067-1177, :1661-1697; grad (N, sh_dim) float32 numpy accumulated, mask uint8 (N,) or None"""
    links, sh, cells = _np(links, np.int32), _np(sh, np.float32), _np(cells, np.int32)
    sv, su = _np(basis_fn, np.float32), _np(basis_fn_u, np.float32)
    lib().oracle_lumisphere_tv_grad_sparse(_ptr(links), _sz(links), _ptr(sh), C.c_int(sh.shape[1]), C.c_int(basis_dim), _ptr(cells),
                                           C.c_int64(cells.shape[0]), _ptr(sv), _ptr(su), C.c_float(scale), C.c_float(dir_factor),
                                           _ptr(mask), _ptr(grad))
