"""GPU: the drop-in claim, proven through the reference's OWN Python.

`svox2/svox2.py` of the reference (UNMODIFIED, staged beside the comparator in oracle/_ref/pyref by
oracle/build_ref_cuda.sh) is imported twice -- once with `svox2.csrc` = OUR compiled extension module
(alphasurf_b200/csrc/host/svox2_shim.cpp over the C ABI, what INTEGRATION.md installs), once with the reference's own
extension -- and drives the same alpha-Surf training iterations and evaluation renders exactly as opt/opt.py does
(:806-830 volume_render_fused, :950-1060 inplace_*_grad, :1120-1145 optim_*_step, :443-563 volume_render_image /
volume_render_depth_image).  Nothing of this repository sits between the reference's Python and the extension.
"""
import os
import sys
import types
import warnings

import numpy as np
import pytest
import torch

from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
PYREF = os.path.join(H.ROOT, "oracle", "_ref", "pyref")
_MODS = ("svox2", "svox2.csrc", "svox2.svox2", "svox2.utils", "svox2.defs", "svox2.version")


def _import_reference_python(csrc):
    """a fresh import of the reference package with `svox2.csrc` = csrc (svox2/utils.py:32-46 picks it up)"""
    if not os.path.isdir(os.path.join(PYREF, "svox2")):
        raise RuntimeError("oracle/_ref/pyref/svox2 is missing: run oracle/build_ref_cuda.sh where /root/reference exists")
    for k in _MODS:
        sys.modules.pop(k, None)
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))    # module-level import of an absent package (svox2.py:16)
    sys.modules["svox2.csrc"] = csrc
    sys.path.insert(0, PYREF)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import svox2
            from svox2 import utils
        assert utils._get_c_extension() is csrc
        return svox2
    finally:
        sys.path.remove(PYREF)


def _forget_reference_python():
    for k in _MODS:
        sys.modules.pop(k, None)


def _make_grid(svox2, sg):
    """SparseGrid of the reference carrying the tensors of a synthetic grid (what load() does with a checkpoint,
    svox2.py:4749-4838)."""
    R = sg.links.shape[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        grid = svox2.SparseGrid(reso=R, center=[0.0, 0.0, 0.0], radius=[1.0, 1.0, 1.0], basis_dim=sg.basis_dim,
                                use_z_order=True, device="cuda", background_nlayers=0, basis_type=svox2.BASIS_TYPE_SH,
                                surface_type=svox2.SURFACE_TYPE_SDF, use_sphere_bound=False, trainable_fake_sample_std=False,
                                surface_init=None, use_octree=False)
    grid.links = sg.links.clone()
    grid.capacity = sg.capacity
    grid.density_data = torch.nn.Parameter(sg.density.clone())
    grid.sh_data = torch.nn.Parameter(sg.sh.clone())
    grid.surface_data = torch.nn.Parameter(sg.surface.clone())
    grid.level_set_data = sg.level_set.clone()
    grid.truncated_vol_render_a = sg.truncated_vol_render_a
    for k, v in synth.alphasurf_render_options().items():
        if hasattr(grid.opt, k):
            setattr(grid.opt, k, v)
    grid.opt.backend = "surf_trav"
    return grid


def _train_and_render(csrc, sg, n_iters, Q):
    """opt.py's iteration, through SparseGrid only"""
    svox2 = _import_reference_python(csrc)
    try:
        grid = _make_grid(svox2, sg)
        np.random.seed(1234)            # the cell windows are drawn with np.random (svox2.py:6344, :6368)
        fused = synth.alphasurf_fused_args()
        rgbs = []
        for it in range(n_iters):
            o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=500 + it)
            rays = svox2.Rays(o, d)
            out = grid.volume_render_fused(rays, gt, beta_loss=0.0, sparsity_loss=0.0, lambda_l2=fused["lambda_l2"],
                                           lambda_l1=fused["lambda_l1"], lambda_l_dist=0.0,
                                           lambda_l_entropy=fused["lambda_l_entropy"], no_norm_weight_l_entropy=False,
                                           lambda_l_dist_a=0.0, lambda_l_entropy_a=0.0, lambda_l_samp_dist=0.0, lambda_l_di=0.0,
                                           l_di_alpha_thresh=0.0, surf_sparse_alpha_thresh=0.15, lambda_inplace_surf_sparse=0.0,
                                           lambda_inwards_norm_loss=0.0, lambda_conv_mode_samp=fused["lambda_conv_mode_samp"],
                                           l_dist_max_sample=64, randomize=False, no_surface=False)
            rgbs.append(out["rgb"].clone())
            grid.inplace_tv_grad(grid.density_data.grad, scaling=1e-5, sparse_frac=0.01, ndc_coeffs=(-1.0, -1.0), contiguous=True)
            grid.inplace_tv_surface_grad(grid.surface_data.grad, scaling=1e-3, sparse_frac=1.0, ndc_coeffs=(-1.0, -1.0),
                                         contiguous=True, ignore_edge=True, edge_value=-1.0, alpha_dependency=False)
            grid.inplace_surface_normal_grad(grid.surface_data.grad, scaling=1e-6, sparse_frac=1.0, ndc_coeffs=(-1.0, -1.0),
                                             contiguous=True, connectivity_check=False, ignore_empty=False, use_l1=True)
            grid.inplace_alpha_surf_sparsify_grad(grid.density_data.grad, grid.surface_data.grad, scaling_alpha=1e-9,
                                                  scaling_surf=0.0, sparse_frac=0.1, surf_sparse_decrease=True,
                                                  surf_sparse_thresh=0.15, alpha_sparsify_bound=0.0, surf_sparsify_bound=-0.1,
                                                  only_trained_cells=False, trained_cells_mask=None, contiguous=True)
            grads = {k: getattr(grid, k + "_data").grad.clone() for k in ("density", "surface", "sh")}
            masks = (grid.sparse_grad_indexer.clone(), grid.sparse_sh_grad_indexer.clone())
            grid.optim_density_step(1e-2, beta=0.95, optim="rmsprop")
            grid.optim_surface_step(1e-5, beta=0.95, optim="rmsprop")
            grid.optim_sh_step(1e-3, beta=0.95, optim="rmsprop")
        # evaluation as opt.py does it: 5000-ray chunks through volume_render / volume_render_depth (svox2.py:3671-3683)
        c2w = torch.eye(4, device="cuda")
        c2w[:3, 3] = torch.tensor([0.0, 0.0, -2.6], device="cuda")
        cam = svox2.Camera(c2w, 180.0, 180.0, 64.0, 48.0, 128, 96, ndc_coeffs=(-1.0, -1.0))
        with torch.no_grad():
            img = grid.volume_render_image(cam, use_kernel=True)
            depth = grid.volume_render_depth_image(cam)
        torch.cuda.synchronize()
        params = {k: getattr(grid, k + "_data").data.clone() for k in ("density", "surface", "sh")}
        return dict(rgbs=rgbs, grads=grads, masks=masks, params=params, img=img.clone(), depth=depth.clone())
    finally:
        _forget_reference_python()


@pytest.mark.parametrize("variant,reso", [("G", 128), ("G*", 64)])
def test_reference_python_runs_unchanged_on_our_extension(variant, reso):
    from alphasurf_b200 import build_shim
    ours = build_shim.load()
    ref = H.load_reference_cuda()
    sg = synth.make_shell_grid(reso, basis_dim=9, variant=variant).to("cuda")
    Q = 8192
    a = _train_and_render(ours, sg, 1, Q)
    b = _train_and_render(ref, sg, 1, Q)
    # the reference on surface scalars moved by one ulp: the conditioning yardstick (helpers.assert_close_conditioned)
    sg_p = synth.SynthGrid(sg.links, sg.density, H.ulp_perturbed(sg.surface), sg.sh, sg.level_set, sg.offset, sg.scaling,
                           sg.basis_dim, sg.fake_sample_std, sg.truncated_vol_render_a, dict(sg.meta))
    p = _train_and_render(ref, sg_p, 1, Q)
    # one iteration: identical inputs on both sides
    assert H.rel_err(a["rgbs"][0], b["rgbs"][0]) < 1e-4
    assert torch.equal(a["masks"][0], b["masks"][0]) and torch.equal(a["masks"][1], b["masks"][1])
    assert int(a["masks"][1].sum()) > 0 and int(a["masks"][0].sum()) > sg.capacity // 2
    for k in ("density", "surface", "sh"):
        H.assert_close_conditioned(a["grads"][k], b["grads"][k], p["grads"][k], 1e-4, k)
    for k in ("img", "depth"):
        assert a[k].shape == b[k].shape
        assert H.rel_err(a[k], b[k]) < 1e-3, k       # parameters after one RMSprop step differ by atomic-order noise
    assert float((a["img"] - 1.0).abs().max()) > 1e-2
    # three iterations on our side alone: the optimizer moved the grid, everything stays finite
    c = _train_and_render(ours, sg, 3, Q)
    for k in ("density", "surface", "sh"):
        assert bool(torch.isfinite(c["params"][k]).all())
    assert float((c["params"]["density"] - sg.density).abs().max()) > 0
    assert float((c["params"]["sh"] - sg.sh).abs().max()) > 0
