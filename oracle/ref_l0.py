"""TEST INFRASTRUCTURE ONLY -- runs the reference's own pure-PyTorch "gradcheck" renderer (L0) on CPU.

Imports /root/reference/svox2 UNMODIFIED (nothing is copied into the repository's history).  On the GPU box, which has no
/root/reference, the package is found where oracle/build_ref_cuda.sh staged it (oracle/_ref/pyref, git-ignored, travels with
the snapshot like the compiled comparator).  Used by oracle/gen_golden.py to write tests/golden/*.npz, by the CPU tests
that are skipped when the package is absent, and by bench.py's cpu_baseline (config C1, kind "reference").

Shims needed to import and run it on CPU:
  * ``mcubes`` is not installed -> stub module (module-level import at svox2/svox2.py:16);
  * ``svox2.csrc`` is absent -> the reference sets ``_C = None`` and warns (svox2/utils.py:32-46);
  * ``torch.tensor(..., device='cuda')`` is hard-coded at svox2/svox2.py:2482 -> redirected to CPU;
  * ``use_octree=False`` (kaolin), ``surf_fake_sample=True`` (else UnboundLocalError at svox2/svox2.py:2550).
"""
import contextlib
import os
import sys
import types
import warnings

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")   # oracle/build_ref_cuda.sh
REF_ROOT = os.environ.get("ASURF_REFERENCE", "/root/reference" if os.path.isdir("/root/reference/svox2") else _STAGED)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "svox2"))


def import_reference():
    if "mcubes" not in sys.modules:
        sys.modules["mcubes"] = types.ModuleType("mcubes")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import svox2  # noqa
    return svox2


@contextlib.contextmanager
def cpu_tensor_redirect():
    orig = torch.tensor

    def patched(*a, **k):
        if k.get("device", None) is not None and "cuda" in str(k["device"]):
            k["device"] = "cpu"
        return orig(*a, **k)

    torch.tensor = patched
    try:
        yield
    finally:
        torch.tensor = orig


def build_reference_grid(sg, opt: dict):
    """SparseGrid carrying the tensors of a SynthGrid (cubic grids only)."""
    svox2 = import_reference()
    R = sg.links.shape[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        grid = svox2.SparseGrid(reso=R, center=[0.0, 0.0, 0.0], radius=[1.0, 1.0, 1.0], basis_dim=sg.basis_dim,
                                use_z_order=False, device="cpu", background_nlayers=0,
                                basis_type=svox2.BASIS_TYPE_SH, surface_type=svox2.SURFACE_TYPE_SDF,
                                use_sphere_bound=False, trainable_fake_sample_std=True, surface_init=None,
                                use_octree=False)
    grid.links = sg.links.clone().cpu()
    grid.capacity = sg.capacity
    grid.density_data = torch.nn.Parameter(sg.density.clone().cpu())
    grid.sh_data = torch.nn.Parameter(sg.sh.clone().cpu())
    grid.surface_data = torch.nn.Parameter(sg.surface.clone().cpu())
    grid.level_set_data = sg.level_set.clone().cpu()
    grid.fake_sample_std = torch.nn.Parameter(torch.tensor([[float(sg.fake_sample_std)]], dtype=torch.float32))
    grid.truncated_vol_render_a = sg.truncated_vol_render_a
    for k, v in opt.items():
        if hasattr(grid.opt, k):
            setattr(grid.opt, k, v)
    return svox2, grid


def render_l0(sg, opt: dict, origins, dirs, run_backward=True, lambda_l_entropy=0.0, lambda_conv_mode_samp=0.0,
              sparsity_loss=0.0):
    """Forward (+ backward of mean|rgb|, the loss hard-wired at svox2/svox2.py:2817-2828) through
    SparseGrid.volume_render(use_kernel=False) -> _surface_render_gradcheck_lerp (svox2/svox2.py:1596-2857)."""
    svox2, grid = build_reference_grid(sg, opt)
    rays = svox2.Rays(origins.clone().cpu(), dirs.clone().cpu())
    with cpu_tensor_redirect(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = grid.volume_render(rays, use_kernel=False, allow_outside=False, run_backward=run_backward,
                                 lambda_l_dist=0, lambda_l_entropy=lambda_l_entropy, sparsity_loss=sparsity_loss,
                                 lambda_conv_mode_samp=lambda_conv_mode_samp)
    res = {"rgb": out["rgb"].detach().float()}
    if run_backward:
        z = lambda p: (p.grad.detach().float() if p.grad is not None else torch.zeros_like(p.data))
        res.update(grad_density=z(grid.density_data), grad_sh=z(grid.sh_data), grad_surface=z(grid.surface_data),
                   grad_fake_sample_std=z(grid.fake_sample_std))
    return res


def render_l0_cuvol(sg, opt: dict, origins, dirs, rgb_gt=None):
    """Plenoxels flavour: SparseGrid.volume_render(use_kernel=False) -> _volume_render_gradcheck_lerp
    (svox2/svox2.py:1215-1441) on a grid WITHOUT a surface; backward of mean((rgb - rgb_gt)^2), the loss the fused CUDA
    call differentiates (render_lerp_kernel_cuvol.cu:873-880)."""
    svox2 = import_reference()
    R = sg.links.shape[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        grid = svox2.SparseGrid(reso=R, center=[0.0, 0.0, 0.0], radius=[1.0, 1.0, 1.0], basis_dim=sg.basis_dim,
                                use_z_order=False, device="cpu", background_nlayers=0, basis_type=svox2.BASIS_TYPE_SH,
                                surface_type=svox2.SURFACE_TYPE_NONE, use_sphere_bound=False, use_octree=False)
    grid.links = sg.links.clone().cpu()
    grid.capacity = sg.capacity
    grid.density_data = torch.nn.Parameter(sg.density.clone().cpu())
    grid.sh_data = torch.nn.Parameter(sg.sh.clone().cpu())
    for k, v in opt.items():
        if hasattr(grid.opt, k):
            setattr(grid.opt, k, v)
    grid.opt.backend = "cuvol"
    rays = svox2.Rays(origins.clone().cpu(), dirs.clone().cpu())
    with cpu_tensor_redirect(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb = grid.volume_render(rays, use_kernel=False)["rgb"]
    res = {"rgb": rgb.detach().float()}
    if rgb_gt is not None:
        loss = ((rgb - rgb_gt.cpu()) ** 2).mean()
        loss.backward()
        res.update(grad_density=grid.density_data.grad.detach().float(), grad_sh=grid.sh_data.grad.detach().float())
    return res
