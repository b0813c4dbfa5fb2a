"""Shared test plumbing: SynthGrid -> spec objects of our module / the compiled reference module / the CPU oracle."""
import glob
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from alphasurf_b200 import synth  # noqa: E402

SURFACE_TYPE_SDF = 102
BASIS_TYPE_SH = 1


def fill_grid_spec(mod, sg):
    g = mod.SparseGridSpec()
    g.density_data = sg.density
    g.surface_type = SURFACE_TYPE_SDF
    g.surface_data = sg.surface
    g.level_set_data = sg.level_set
    g.sh_data = sg.sh
    g.links = sg.links
    g._offset = sg.offset
    g._scaling = sg.scaling
    g.basis_dim = sg.basis_dim
    g.basis_type = BASIS_TYPE_SH
    g.fake_sample_std = float(sg.fake_sample_std)
    g.truncated_vol_render_a = float(sg.truncated_vol_render_a)
    return g


def fill_rays_spec(mod, origins, dirs):
    r = mod.RaysSpec()
    r.origins = origins
    r.dirs = dirs
    r.masks = torch.ones((origins.shape[0],), dtype=torch.bool, device=origins.device)
    return r


def fill_opt(mod, d):
    o = mod.RenderOptions()
    for k, v in d.items():
        if k == "backend":
            continue
        setattr(o, k, v)
    return o


class GradSet:
    def __init__(self, sg, device, with_std=True):
        self.density = torch.zeros_like(sg.density, device=device)
        self.surface = torch.zeros_like(sg.surface, device=device)
        self.sh = torch.zeros_like(sg.sh, device=device)
        self.std = torch.zeros((1, 1), dtype=torch.float32, device=device) if with_std else None
        self.mask = torch.zeros((sg.capacity,), dtype=torch.bool, device=device)

    def spec(self, mod):
        g = mod.GridOutputGrads()
        g.grad_density_out = self.density
        g.grad_surface_out = self.surface
        g.grad_sh_out = self.sh
        if self.std is not None:
            g.grad_fake_sample_std_out = self.std
        g.mask_out = self.mask
        return g


FUSED_ORDER = ["beta_loss", "sparsity_loss", "fused_surf_norm_reg_scale", "fused_surf_norm_reg_con_check",
               "fused_surf_norm_reg_ignore_empty", "lambda_l2", "lambda_l1", "lambda_l_dist", "lambda_l_entropy",
               "no_norm_weight_l_entropy", "lambda_l_dist_a", "lambda_l_entropy_a", "lambda_l_samp_dist", "lambda_l_di",
               "l_di_alpha_thresh", "surf_sparse_alpha_thresh", "lambda_inplace_surf_sparse", "lambda_inwards_norm_loss",
               "lambda_conv_mode_samp", "l_dist_max_sample"]


def fused_positional(fd):
    return [fd[k] for k in FUSED_ORDER]


def load_reference_cuda(required=None):
    """The UNMODIFIED reference extension compiled by oracle/build_ref_cuda.sh.  On a machine with a GPU the comparator is
    REQUIRED: a missing build fails the calling test instead of silently reducing it to the oracle comparison
    (``required=False``: return None instead, for bench.py's optional extra)."""
    cands = glob.glob(os.path.join(ROOT, "oracle", "_ref", "svox2_ref_csrc*.so"))
    if not cands:
        if required is None:
            required = torch.cuda.is_available()
        if required:
            raise RuntimeError("oracle/_ref/svox2_ref_csrc*.so is missing: run __graft_entry__.build() where /root/reference "
                               "exists (oracle/build_ref_cuda.sh); the GPU parity tests compare against it")
        return None
    name = "svox2_ref_csrc"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def rel_err(a, b):
    """max |a-b| relative to the scale of b (atomics make element order differ, so scale by the tensor max)."""
    a, b = a.double(), b.double()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def ulp_perturbed(t, seed=0):
    """every element moved by ONE ulp, up or down at random"""
    g = torch.Generator(device=t.device).manual_seed(seed)
    up = torch.rand(t.shape, device=t.device, generator=g) < 0.5
    return torch.where(up, torch.nextafter(t, t + 1), torch.nextafter(t, t - 1))


def assert_close_conditioned(a, b, b_pert, tol, what, max_elems=2048):
    """a: ours, b: the reference, b_pert: the REFERENCE ITSELF on inputs moved by one ulp (ulp_perturbed).

    The root Jacobian of the cubic ray / level-set intersection (render_util.cuh:1206-1415) divides by the discriminant: at
    a near-double root a last-bit difference (fp64 libm, FMA contraction) is amplified ~1e5-fold, in the reference as in any
    other implementation.  Elements the reference cannot pin down to 0.1 tol under a one-ulp change of its own inputs are
    "ill-conditioned": there must be few of them and we must agree within tol + 3x the reference's own movement; every
    other element must agree within tol (relative to the tensor max, as rel_err)."""
    a, b, p = a.double().reshape(-1), b.double().reshape(-1), b_pert.double().reshape(-1)
    mx = b.abs().max().clamp_min(1e-30)
    own = (p - b).abs()
    cond = own > 0.1 * tol * mx
    err = (a - b).abs()
    n_cond = int(cond.sum())
    assert n_cond <= max_elems, "%s: %d ill-conditioned elements" % (what, n_cond)
    if bool((~cond).any()):
        e = float(err[~cond].max() / mx)
        assert e < tol, "%s: %.3e on the well-conditioned elements" % (what, e)
    if n_cond:
        e, o = float(err[cond].max() / mx), float(own[cond].max() / mx)
        assert e < tol + 3 * o, "%s: %.3e on the %d ill-conditioned elements (reference moves %.3e by itself)" % (what, e, n_cond, o)
    return n_cond
