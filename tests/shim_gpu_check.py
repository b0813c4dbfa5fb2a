"""Body of tests/test_zz_shim_gpu.py, run in a process of its own (python -m tests.shim_gpu_check): the compiled svox2.csrc
shim against the ctypes mirror on a GPU.  Exit code 0 = equal."""
import sys

import torch

from alphasurf_b200 import build_shim
from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H


def check_fused_render(shim):
    opts, fused = synth.alphasurf_render_options(), synth.alphasurf_fused_args()
    sg = synth.make_shell_grid(48, basis_dim=9, variant="G*").to("cuda")
    o, d, gt = synth.make_camera_rays(1024, device="cuda", seed=31)
    res = []
    for mod in (ours, shim):
        G = H.GradSet(sg, "cuda")
        rgb = torch.zeros_like(o)
        mod.volume_render_surf_trav_fused(H.fill_grid_spec(mod, sg), H.fill_rays_spec(mod, o, d), H.fill_opt(mod, opts), gt,
                                          *H.fused_positional(fused), rgb, G.spec(mod))
        fwd = mod.volume_render_surf_trav(H.fill_grid_spec(mod, sg), H.fill_rays_spec(mod, o, d), H.fill_opt(mod, opts))
        depth = mod.volume_render_expected_term_surf_trav(H.fill_grid_spec(mod, sg), H.fill_rays_spec(mod, o, d),
                                                          H.fill_opt(mod, opts))
        torch.cuda.synchronize()
        res.append((rgb, fwd, depth, G))
    (rgb_a, fwd_a, dep_a, Ga), (rgb_b, fwd_b, dep_b, Gb) = res
    assert torch.equal(rgb_a, rgb_b) and torch.equal(fwd_a, fwd_b) and torch.equal(dep_a, dep_b)
    assert torch.equal(Ga.mask, Gb.mask)
    for k in ("density", "surface", "sh"):
        assert H.rel_err(getattr(Ga, k), getattr(Gb, k)) < 1e-5      # atomic order only


def check_losses_optimizer_queries(shim):
    sg = synth.make_shell_grid(32, basis_dim=4, variant="G").to("cuda")
    cells = torch.nonzero(sg.links.reshape(-1) >= 0).flatten().to(torch.int32)
    pts = (torch.rand((2000, 3), generator=torch.Generator().manual_seed(1)) * 2.2 - 1.1).cuda()
    out = []
    for mod in (ours, shim):
        g_surf = torch.zeros_like(sg.surface)
        mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
        mod.surf_tv_grad_sparse(sg.links, sg.surface, sg.density, cells, mask, 0, 1, 1e-3, True, -1.0, False, -1.0, -1.0, False,
                                g_surf)
        mod.surface_normal_grad_sparse(sg.links, sg.surface, cells, mask, 0.0, 0, 1, 1e-2, 0.0, -1.0, -1.0, False, False, True,
                                       g_surf)
        data, rms, grad = sg.density.clone(), torch.zeros_like(sg.density), torch.full_like(sg.density, 0.01)
        mod.rmsprop_step(data, rms, grad, mask, 0.95, 1e-2, 1e-8, -1e9, 1e-2)
        dens, sh = mod.sample_grid(H.fill_grid_spec(mod, sg), pts, True)
        dil = mod.dilate(sg.links >= 0)
        torch.cuda.synchronize()
        out.append((g_surf, mask, data, rms, grad, dens, sh, dil))
    a, b = out
    assert H.rel_err(a[0], b[0]) < 1e-5
    for x, y in zip(a[1:], b[1:]):
        assert torch.equal(x, y)
    for exc, call in ((RuntimeError, lambda: shim.sample_grid(H.fill_grid_spec(shim, sg), pts.cpu(), True)),
                      (NotImplementedError, lambda: shim.volume_render_nvol(None))):
        try:
            call()
        except exc:
            continue
        raise AssertionError("expected %s" % exc.__name__)


if __name__ == "__main__":
    assert torch.cuda.is_available()
    m = build_shim.load()
    check_fused_render(m)
    check_losses_optimizer_queries(m)
    print("shim == ctypes mirror")
    sys.exit(0)
