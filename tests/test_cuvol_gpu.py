"""GPU parity of the Plenoxels "cuvol" renderer: ours (svox2.csrc-compatible module -> C ABI -> sm_100a kernels) against
(a) the CPU oracle (oracle/oracle_cuvol.c, pinned on the reference's PyTorch renderer by tests/golden/l0_cuvol_*.npz) and
(b) the UNMODIFIED reference CUDA kernels, with and without the negative skip codes that accel_dist_prop writes into links.

Tolerances: sample selection is exact by construction (same t sequence); colours and gradients <= 1e-4 relative of the
tensor max (fast intrinsics, atomic order)."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4
SURFACE_TYPE_NONE = 100


def plenoxels_options(**kw):
    """opt/util/config_util.py:81-92 defaults for the cuvol backend (configs/syn.yaml)."""
    o = synth.alphasurf_render_options()
    o.update(backend="cuvol", sigma_thresh=1e-8, stop_thresh=1e-7, step_size=0.5)
    o.update(kw)
    return o


def _grid_spec(mod, sg, links=None):
    """SparseGrid._to_cpp for a grid without a surface (svox2.py:6234-6272): surface tensors are left unset."""
    g = mod.SparseGridSpec()
    g.density_data = sg.density
    g.sh_data = sg.sh
    g.links = sg.links if links is None else links
    g._offset = sg.offset
    g._scaling = sg.scaling
    g.basis_dim = sg.basis_dim
    g.basis_type = H.BASIS_TYPE_SH
    g.surface_type = SURFACE_TYPE_NONE
    return g


class Grads:
    def __init__(self, sg):
        self.density = torch.zeros_like(sg.density)
        self.sh = torch.zeros_like(sg.sh)
        self.mask = torch.zeros((sg.capacity,), dtype=torch.bool, device=sg.density.device)

    def spec(self, mod):
        g = mod.GridOutputGrads()
        g.grad_density_out = self.density
        g.grad_sh_out = self.sh
        g.mask_out = self.mask
        return g


def _xf(sg, o, d, opts):
    """grid-space rays exactly as the GPU computes them (rnorm3df), handed to the oracle"""
    return ours.debug_ray_bounds(H.fill_grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts))


def _with_skip_codes(sg):
    """links with accel_dist_prop's negative codes (from the reference kernel when available, else hand-made 4^3 codes)."""
    ref = H.load_reference_cuda()
    links = sg.links.clone()
    if ref is not None:
        ref.accel_dist_prop(links)
        return links
    occ = links >= 0
    pad = torch.nn.functional.pad(occ.float()[None, None], (1, 1, 1, 1, 1, 1))
    blk = torch.nn.functional.max_pool3d(pad, kernel_size=6, stride=4)[0, 0]
    code = torch.where(blk == 0, torch.tensor(-3, device=links.device), torch.tensor(-1, device=links.device)).to(torch.int32)
    full = code.repeat_interleave(4, 0).repeat_interleave(4, 1).repeat_interleave(4, 2)
    return torch.where(occ, links, full[:links.shape[0], :links.shape[1], :links.shape[2]]).contiguous()


CASES = [("sh2-64", 64, 9, 2048, {}), ("sh1-48", 48, 4, 1024, {}),
         ("sh2-64-opaque-sparsity", 64, 9, 1024, dict(last_sample_opaque=True))]


@pytest.mark.parametrize("name,reso,bd,Q,okw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("skip_codes", [False, True])
def test_cuvol_vs_oracle_and_reference(name, reso, bd, Q, okw, skip_codes):
    from oracle import oracle
    opts = plenoxels_options(**okw)
    sg = synth.make_shell_grid(reso, basis_dim=bd, variant="G", sigma_density=True).to("cuda")
    o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=synth.SEED + 5)
    links = _with_skip_codes(sg) if skip_codes else sg.links
    if skip_codes:
        assert int((links < -1).sum()) > 0
    beta, sparsity = (0.0, 0.0) if not okw else (1e-2, 1e-3)
    grid, rays, opt = _grid_spec(ours, sg, links), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts)

    G = Grads(sg)
    rgb = torch.zeros_like(o)
    ours.volume_render_cuvol_fused(grid, rays, opt, gt, beta, sparsity, rgb, G.spec(ours))
    rgb_f = ours.volume_render_cuvol(grid, rays, opt)
    torch.cuda.synchronize()
    assert torch.equal(rgb, rgb_f)
    assert float((rgb - 1.0).abs().max()) > 1e-2 and int(G.mask.sum()) > 0

    # (a) CPU oracle on the GPU's grid-space rays
    sg_cpu = sg.to("cpu")
    og = oracle.Grid(synth.SynthGrid(links.cpu(), sg_cpu.density, None, sg_cpu.sh, None, sg.offset, sg.scaling, bd))
    xf = _xf(sg, o, d, opts).cpu()
    rgb_o, g_o = oracle.cuvol_fused(og, opts, o.cpu(), d.cpu(), gt.cpu(), beta_loss=beta, sparsity_loss=sparsity, xf=xf)
    assert H.rel_err(rgb.cpu(), torch.from_numpy(rgb_o)) < TOL
    assert H.rel_err(G.sh.cpu(), torch.from_numpy(g_o.sh)) < TOL
    assert H.rel_err(G.density.cpu(), torch.from_numpy(g_o.density)) < TOL
    assert np.array_equal(G.mask.cpu().numpy().astype(np.uint8), g_o.mask)

    # (b) the reference kernels
    ref = H.load_reference_cuda()
    if ref is None:
        return
    Gr = Grads(sg)
    rgb_r = torch.zeros_like(o)
    ref.volume_render_cuvol_fused(_grid_spec(ref, sg, links), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts), gt, beta,
                                  sparsity, rgb_r, Gr.spec(ref))
    torch.cuda.synchronize()
    assert torch.equal(G.mask, Gr.mask)
    assert H.rel_err(rgb, rgb_r) < TOL
    assert H.rel_err(G.sh, Gr.sh) < TOL
    assert H.rel_err(G.density, Gr.density) < TOL
    # non-fused backward entry point
    gout = torch.randn(rgb.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    G2, G2r = Grads(sg), Grads(sg)
    ours.volume_render_cuvol_backward(grid, rays, opt, gout, rgb_r, G2.spec(ours))
    ref.volume_render_cuvol_backward(_grid_spec(ref, sg, links), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts), gout,
                                     rgb_r, G2r.spec(ref))
    torch.cuda.synchronize()
    assert torch.equal(G2.mask, G2r.mask)
    assert H.rel_err(G2.sh, G2r.sh) < TOL
    assert H.rel_err(G2.density, G2r.density) < TOL


def test_cuvol_image_equals_ray_render():
    """volume_render_cuvol_image generates the camera rays in the kernel (cam2world_ray); it must agree with rendering the
    same rays passed explicitly, and with the reference's image kernel."""
    opts = plenoxels_options()
    sg = synth.make_shell_grid(64, basis_dim=9, variant="G", sigma_density=True).to("cuda")
    W = Hh = 96
    cam = ours.CameraSpec()
    c2w = torch.eye(4)
    v = torch.tensor([0.6, -0.5, 0.62]); v = v / v.norm()
    fwd = -v
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0])); right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, v * 2.6
    cam.c2w = c2w[:3, :4].contiguous().cuda()
    cam.fx = cam.fy = 110.0
    cam.cx, cam.cy = W * 0.5, Hh * 0.5
    cam.width, cam.height = W, Hh
    img = ours.volume_render_cuvol_image(_grid_spec(ours, sg), cam, H.fill_opt(ours, opts))
    yy, xx = torch.meshgrid(torch.arange(Hh, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    x = (xx + 0.5 - cam.cx) / cam.fx
    y = (yy + 0.5 - cam.cy) / cam.fy
    z = torch.sqrt(x * x + y * y + 1.0)
    dcam = torch.stack([x / z, y / z, 1.0 / z], -1).reshape(-1, 3)
    dirs = (dcam @ c2w[:3, :3].T).contiguous().cuda()
    orig = c2w[:3, 3].expand_as(dirs).contiguous().cuda()
    out = ours.volume_render_cuvol(_grid_spec(ours, sg), H.fill_rays_spec(ours, orig, dirs), H.fill_opt(ours, opts))
    torch.cuda.synchronize()
    assert img.shape == (Hh, W, 3)
    assert H.rel_err(img.reshape(-1, 3), out) < 1e-4
    assert float((img - 1.0).abs().max()) > 1e-2
    ref = H.load_reference_cuda()
    if ref is not None and hasattr(ref, "volume_render_cuvol_image"):
        rc = ref.CameraSpec()
        rc.c2w, rc.fx, rc.fy, rc.cx, rc.cy, rc.width, rc.height = cam.c2w, cam.fx, cam.fy, cam.cx, cam.cy, W, Hh
        rc.ndc_coeffx = rc.ndc_coeffy = -1.0
        img_r = ref.volume_render_cuvol_image(_grid_spec(ref, sg), rc, H.fill_opt(ref, opts))
        assert H.rel_err(img, img_r) < 1e-4


def test_cuvol_full_size_properties():
    """C2-sized property checks (256^3, SH deg 2, 5000 rays): rendering twice is bit-identical, an empty grid renders the
    background, and a fully opaque first sample saturates the transmittance (colours independent of what lies behind)."""
    opts = plenoxels_options()
    sg = synth.make_shell_grid(256, basis_dim=9, variant="G", sigma_density=True).to("cuda")
    o, d, gt = synth.make_camera_rays(5000, device="cuda")
    grid, rays, opt = _grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts)
    a = ours.volume_render_cuvol(grid, rays, opt)
    b = ours.volume_render_cuvol(grid, rays, opt)
    assert torch.equal(a, b)
    hit = (a - 1.0).abs().max(dim=1).values > 1e-3
    assert 0.2 < float(hit.float().mean()) < 0.5
    empty = synth.SynthGrid(torch.full_like(sg.links, -1), sg.density, None, sg.sh, None, sg.offset, sg.scaling, 9)
    c = ours.volume_render_cuvol(_grid_spec(ours, empty), rays, opt)
    assert float((c - 1.0).abs().max()) == 0.0
    G = Grads(sg)
    rgb = torch.zeros_like(o)
    ours.volume_render_cuvol_fused(grid, rays, opt, gt, 0.0, 0.0, rgb, G.spec(ours))
    torch.cuda.synchronize()
    # gradients only on rows that a hit ray can touch, and the SH DC gradient has the sign of (rgb - gt) on average
    assert int(G.mask.sum()) > 0 and float(G.sh[~G.mask].abs().max()) == 0.0
    assert bool(torch.isfinite(G.sh).all()) and bool(torch.isfinite(G.density).all())


@pytest.mark.parametrize("skip_codes", [False, True])
def test_cuvol_depth_renders(skip_codes):
    """volume_render_expected_term / _mode_term / _med_term / _sigma_thresh (the depth maps opt.py logs for the cuvol
    backend): ours vs the CPU oracle on the same grid-space rays and vs the reference kernels.  The sample positions are
    exact; a ray may flip only where a weight / sigma sits within rounding of the threshold."""
    from oracle import oracle
    opts = plenoxels_options()
    bd, Q = 9, 4096
    sg = synth.make_shell_grid(64, basis_dim=bd, variant="G", sigma_density=True).to("cuda")
    o, d, _ = synth.make_camera_rays(Q, device="cuda", seed=synth.SEED + 9)
    links = _with_skip_codes(sg) if skip_codes else sg.links
    grid, rays, opt = _grid_spec(ours, sg, links), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts)
    sg_cpu = sg.to("cpu")
    og = oracle.Grid(synth.SynthGrid(links.cpu(), sg_cpu.density, None, sg_cpu.sh, None, sg.offset, sg.scaling, bd))
    xf = _xf(sg, o, d, opts).cpu()
    ref = H.load_reference_cuda()

    def close(a, b, what, flips=0):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert a.shape == b.shape
        bad = np.abs(a - b) > 1e-5 * np.maximum(1.0, np.abs(b))
        if bad.ndim == 2:
            bad = bad.any(1)
        assert int(bad.sum()) <= flips, (what, int(bad.sum()))
        assert np.count_nonzero(b) > 0, what

    calls = [("expected_term", "volume_render_expected_term", 0.0), ("expected_term", "volume_render_expected_term", 0.9),
             ("mode_term", "volume_render_mode_term", 0.5), ("sigma_thresh", "volume_render_sigma_thresh", 18.0)]
    for mode, fn, param in calls:
        got = getattr(ours, fn)(grid, rays, opt, param)
        torch.cuda.synchronize()
        want = oracle.cuvol_scalar(og, opts, o.cpu(), d.cpu(), mode, param, xf=xf)
        close(got.cpu().numpy(), want, (fn, param, "oracle"), flips=4)
        if ref is not None:
            want_r = getattr(ref, fn)(_grid_spec(ref, sg, links), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts), param)
            close(got.cpu().numpy(), want_r.cpu().numpy(), (fn, param, "reference"), flips=4)
    dep, sig = ours.volume_render_med_term(grid, rays, opt, 12)
    torch.cuda.synchronize()
    dep_o, sig_o = oracle.cuvol_scalar(og, opts, o.cpu(), d.cpu(), "med_term", 0.0, 12, xf=xf)
    close(dep.cpu().numpy(), dep_o, "med depths vs oracle")
    close(sig.cpu().numpy(), sig_o, "med sigmas vs oracle")
    if ref is not None:
        dep_r, sig_r = ref.volume_render_med_term(_grid_spec(ref, sg, links), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts), 12)
        close(dep.cpu().numpy(), dep_r.cpu().numpy(), "med depths vs reference")
        close(sig.cpu().numpy(), sig_r.cpu().numpy(), "med sigmas vs reference")
    # empty batch
    e = torch.zeros((0, 3), device="cuda")
    assert ours.volume_render_expected_term(grid, H.fill_rays_spec(ours, e, e), opt, 0.0).shape == (0,)
