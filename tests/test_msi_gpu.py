"""GPU parity of the multi-sphere-image background (render_lerp_kernel_surf_trav.cu:2914-3137, :3370-3455,
render_util.cuh:207-283, loss_kernel.cu:979-1064): our kernels through the svox2.csrc-compatible API against

  (a) the CPU oracle (oracle/oracle_msi.c) fed with the per-ray state our foreground pass left (log-transmittance, accum), and
  (b) the UNMODIFIED reference CUDA kernels (oracle/_ref) for the whole calls: surf_trav and cuvol, forward / backward / fused /
      image, and msi_tv_grad_sparse.

Tolerance: 1e-4 relative of the tensor max (fast intrinsics, atomic order), masks bit-equal."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H
from tests.test_cuvol_gpu import plenoxels_options

pytestmark = pytest.mark.gpu
TOL = 1e-4
SURFACE_TYPE_NONE = 100


def make_background(reso=32, nlayers=16, seed=3, device="cuda"):
    """(links (2 reso, reso) int32 with a hole of empty texels, data (n, nlayers, 4) float32: rgb ~ N(0, 1), sigma ~ N(0.3, 1)
    so that some layers are transparent (sigma <= 0 is skipped, :2966))"""
    g = torch.Generator().manual_seed(seed)
    keep = torch.rand((2 * reso, reso), generator=g) > 0.15
    links = torch.full((2 * reso, reso), -1, dtype=torch.int32)
    links[keep] = torch.arange(int(keep.sum()), dtype=torch.int32)
    n = int(keep.sum())
    data = torch.randn((n, nlayers, 4), generator=g)
    data[..., 3] = data[..., 3] * 1.0 + 0.3
    return links.to(device).contiguous(), data.to(device).contiguous()


def grid_spec(mod, sg, bg, surface=True, links=None):
    g = H.fill_grid_spec(mod, sg) if surface else mod.SparseGridSpec()
    if not surface:
        g.density_data, g.sh_data = sg.density, sg.sh
        g._offset, g._scaling = sg.offset, sg.scaling
        g.basis_dim, g.basis_type, g.surface_type = sg.basis_dim, H.BASIS_TYPE_SH, SURFACE_TYPE_NONE
    g.links = sg.links if links is None else links
    g.background_links, g.background_data = bg
    return g


class Grads:
    def __init__(self, sg, bg, surface=True):
        self.density = torch.zeros_like(sg.density)
        self.sh = torch.zeros_like(sg.sh)
        self.surface = torch.zeros_like(sg.surface) if surface else None
        self.mask = torch.zeros((sg.capacity,), dtype=torch.bool, device=sg.density.device)
        self.bg = torch.zeros_like(bg[1])
        self.mask_bg = torch.zeros(bg[1].shape[:2], dtype=torch.bool, device=bg[1].device)

    def spec(self, mod):
        g = mod.GridOutputGrads()
        g.grad_density_out, g.grad_sh_out, g.mask_out = self.density, self.sh, self.mask
        if self.surface is not None:
            g.grad_surface_out = self.surface
        g.grad_background_out, g.mask_background_out = self.bg, self.mask_bg
        return g


def _check(G, Gr, keys):
    assert torch.equal(G.mask, Gr.mask) and torch.equal(G.mask_bg, Gr.mask_bg)
    assert int(G.mask_bg.sum()) > 0
    for k in keys:
        assert H.rel_err(getattr(G, k), getattr(Gr, k)) < TOL, k


@pytest.mark.parametrize("variant,reso,Q", [("G", 64, 4096), ("G*", 48, 2048)])
def test_surf_trav_with_background_vs_reference_cuda(variant, reso, Q):
    ref = H.load_reference_cuda()
    sg = synth.make_shell_grid(reso, basis_dim=9, variant=variant).to("cuda")
    bg = make_background()
    opts, fused = synth.alphasurf_render_options(), synth.alphasurf_fused_args()
    fused["beta_loss"] = 0.0
    o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=31)
    res = {}
    for name, mod in (("ours", ours), ("ref", ref)):
        grid, rays, opt = grid_spec(mod, sg, bg), H.fill_rays_spec(mod, o, d), H.fill_opt(mod, opts)
        out = mod.volume_render_surf_trav(grid, rays, opt)
        G = Grads(sg, bg)
        rgb = torch.zeros_like(o)
        mod.volume_render_surf_trav_fused(grid, rays, opt, gt, *H.fused_positional(fused), rgb, G.spec(mod))
        G2 = Grads(sg, bg)
        gout = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
        mod.volume_render_surf_trav_backward(grid, rays, opt, gout, out, G2.spec(mod))
        torch.cuda.synchronize()
        res[name] = (out, rgb, G, G2)
    a, b = res["ours"], res["ref"]
    assert H.rel_err(a[0], b[0]) < TOL and H.rel_err(a[1], b[1]) < TOL
    assert torch.equal(a[0], a[1])                       # fused forward == plain forward
    assert float((a[0] - 1.0).abs().max()) > 1e-2        # the background is not the constant brightness
    _check(a[2], b[2], ("bg", "sh", "density"))
    _check(a[3], b[3], ("bg", "sh", "density"))
    assert H.rel_err(a[2].surface, b[2].surface) < 5 * TOL      # (near-double roots: see helpers.assert_close_conditioned)
    assert H.rel_err(a[3].surface, b[3].surface) < 5 * TOL


@pytest.mark.parametrize("skip_codes", [False, True])
def test_cuvol_with_background_vs_reference_cuda(skip_codes):
    ref = H.load_reference_cuda()
    sg = synth.make_shell_grid(64, basis_dim=9, variant="G", sigma_density=True).to("cuda")
    sg.density.mul_(0.05)                                # thin foreground: light reaches the background
    bg = make_background(seed=5)
    links = sg.links.clone()
    if skip_codes:
        ref.accel_dist_prop(links)
    opts = plenoxels_options()
    o, d, gt = synth.make_camera_rays(2048, device="cuda", seed=32)
    res = {}
    for name, mod in (("ours", ours), ("ref", ref)):
        grid, rays, opt = grid_spec(mod, sg, bg, surface=False, links=links), H.fill_rays_spec(mod, o, d), H.fill_opt(mod, opts)
        out = mod.volume_render_cuvol(grid, rays, opt)
        G = Grads(sg, bg, surface=False)
        rgb = torch.zeros_like(o)
        mod.volume_render_cuvol_fused(grid, rays, opt, gt, 1e-2, 1e-3, rgb, G.spec(mod))
        G2 = Grads(sg, bg, surface=False)
        gout = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        mod.volume_render_cuvol_backward(grid, rays, opt, gout, out, G2.spec(mod))
        cam = mod.CameraSpec()
        c2w = torch.eye(4, device="cuda")
        c2w[:3, 3] = torch.tensor([0.1, -0.2, -2.4], device="cuda")
        cam.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height = c2w, 90.0, 90.0, 40.0, 30.0, 80, 60
        cam.ndc_coeffx = cam.ndc_coeffy = -1.0
        img = mod.volume_render_cuvol_image(grid, cam, opt)
        torch.cuda.synchronize()
        res[name] = (out, rgb, G, G2, img)
    a, b = res["ours"], res["ref"]
    assert H.rel_err(a[0], b[0]) < TOL and H.rel_err(a[1], b[1]) < TOL and H.rel_err(a[4], b[4]) < TOL
    _check(a[2], b[2], ("bg", "sh", "density"))
    _check(a[3], b[3], ("bg", "sh", "density"))


def test_msi_passes_vs_cpu_oracle():
    """The background kernels alone against oracle/oracle_msi.c, from the per-ray state our surf_trav pass left."""
    from oracle import oracle
    sg = synth.make_shell_grid(48, basis_dim=4, variant="G").to("cuda")
    bg = make_background(reso=24, nlayers=12, seed=9)
    opts = synth.alphasurf_render_options()
    o, d, gt = synth.make_camera_rays(1024, device="cuda", seed=33)
    grid, rays, opt = grid_spec(ours, sg, bg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts)
    # forward: colours of the foreground alone (same grid without a background, no brightness) + oracle background
    out = ours.volume_render_surf_trav(grid, rays, opt)
    lt, _ = ours.debug_bg_state(o.shape[0])
    opts0 = dict(opts)
    opts0["background_brightness"] = 0.0
    fg = ours.volume_render_surf_trav(H.fill_grid_spec(ours, sg), rays, H.fill_opt(ours, opts0))
    # fg ends on exp(lt) * 0; with a background the foreground pass adds nothing for the brightness either
    rgb_o = fg.cpu().numpy().copy()
    size = list(sg.links.shape)
    oracle.msi_forward(bg[0], bg[1], size, sg.offset, sg.scaling, opts, o, d, lt, rgb_o)
    assert H.rel_err(out.cpu(), torch.from_numpy(rgb_o)) < TOL
    assert float(lt.min()) < -1e-3 and float((lt == 0).float().mean()) > 0.05      # rays that hit and rays that miss
    # backward (stand-alone form)
    G = Grads(sg, bg)
    gout = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    ours.volume_render_surf_trav_backward(grid, rays, opt, gout, out, G.spec(ours))
    lt_b, acc = ours.debug_bg_state(o.shape[0])
    init = (out * gout).sum(1)
    acc = torch.where(torch.isfinite(acc), acc, init)        # rays the foreground backward never visited start from the initial sum
    g_o = np.zeros(tuple(bg[1].shape), np.float32)
    m_o = np.zeros(tuple(bg[1].shape[:2]), np.uint8)
    oracle.msi_backward(bg[0], bg[1], size, sg.offset, sg.scaling, opts, o, d, gout, out, False, lt_b, acc, 0.0, g_o, m_o)
    assert H.rel_err(G.bg.cpu(), torch.from_numpy(g_o)) < TOL
    assert np.array_equal(G.mask_bg.cpu().numpy().astype(np.uint8), m_o) and m_o.sum() > 0


@pytest.mark.parametrize("with_mask", [True, False])
def test_msi_tv_grad_sparse(with_mask):
    from oracle import oracle
    ref = H.load_reference_cuda()
    links, data = make_background(reso=20, nlayers=10, seed=4)
    n_cells = links.numel() * data.shape[1]
    cells = torch.randint(0, n_cells, (n_cells // 3,), generator=torch.Generator().manual_seed(1)).int().cuda()
    mshape = tuple(data.shape[:2]) if with_mask else (0, 0)
    grad, mask = torch.zeros_like(data), torch.zeros(mshape, dtype=torch.bool, device="cuda")
    ours.msi_tv_grad_sparse(links, data, cells, mask, 0.7, 0.3, grad)
    g_o = np.zeros(tuple(data.shape), np.float32)
    m_o = np.zeros(tuple(data.shape[:2]), np.uint8)
    oracle.msi_tv_grad_sparse(links, data, cells, m_o if with_mask else None, 0.7, 0.3, g_o)
    assert H.rel_err(grad.cpu(), torch.from_numpy(g_o)) < 2e-5
    if with_mask:
        assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o) and m_o.sum() > 0
        g_r, m_r = torch.zeros_like(data), torch.zeros_like(mask)
        ref.msi_tv_grad_sparse(links, data, cells, m_r, 0.7, 0.3, g_r)
        assert H.rel_err(grad, g_r) < 2e-5 and torch.equal(mask, m_r)
