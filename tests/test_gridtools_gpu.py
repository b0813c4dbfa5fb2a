"""GPU parity of the grid-maintenance renders (dilate, grid_weight_render, sparse_grid_weight_render,
sparse_grid_mask_render, sparse_grid_visbility_render_surf): ours through the svox2.csrc-compatible module against the
UNMODIFIED reference kernels (masks and counts exact -- max / count reductions do not depend on the order of the atomics --
weights to 2e-5 of the maximum) and against the
CPU oracle (the host derives the unit ray direction with 1/sqrt instead of rnorm3df, so a vanishing share of the
grazing samples may land in a neighbouring voxel: bounded mismatch counts instead of equality)."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _camera(mod, W=96, Hh=80, f=100.0, v=(0.6, -0.5, 0.62), dist=2.6):
    cam = mod.CameraSpec()
    c2w = torch.eye(4)
    v = torch.tensor(v)
    v = v / v.norm()
    fwd = -v
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, v * dist
    cam.c2w = c2w[:3, :4].contiguous().cuda()
    cam.fx = cam.fy = f
    cam.cx, cam.cy = W * 0.5, Hh * 0.5
    cam.width, cam.height = W, Hh
    cam.ndc_coeffx = cam.ndc_coeffy = -1.0
    return cam


def _oracle_rays(cam):
    from oracle import oracle
    return oracle.cam_rays(cam.c2w.cpu().numpy(), cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height)


def _few_differ(a, b, frac, what):
    a, b = np.asarray(a), np.asarray(b)
    bad = int((np.abs(a.astype(np.float64) - b.astype(np.float64)) > 1e-5 * np.maximum(1.0, np.abs(b))).sum())
    assert bad <= max(2, int(frac * max(np.count_nonzero(b), 1))), (what, bad, np.count_nonzero(b))
    assert np.count_nonzero(b) > 0, what


def test_dilate():
    from oracle import oracle
    g = torch.rand((37, 20, 45), generator=torch.Generator().manual_seed(0)) < 0.03
    out = ours.dilate(g.cuda())
    assert out.dtype == torch.bool and out.shape == g.shape
    assert np.array_equal(out.cpu().numpy(), oracle.dilate(g.numpy()))
    want = torch.nn.functional.max_pool3d(g[None, None].float(), 3, 1, 1)[0, 0] > 0
    assert torch.equal(out.cpu(), want)
    ref = H.load_reference_cuda()
    if ref is not None:
        assert torch.equal(out, ref.dilate(g.cuda()))
    with pytest.raises(RuntimeError):
        ours.dilate(g.cuda().float())


@pytest.mark.parametrize("last_sample_opaque", [False, True])
def test_grid_weight_render_dense(last_sample_opaque):
    from oracle import oracle
    sg = synth.make_shell_grid(48, basis_dim=1, variant="G", sigma_density=True).to("cuda")
    dense = torch.zeros(tuple(sg.links.shape), device="cuda")
    m = sg.links >= 0
    dense[m] = sg.density[sg.links[m].long(), 0]
    cam = _camera(ours)
    off, scl = sg.offset.cuda(), sg.scaling.cuda()
    out = torch.zeros_like(dense)
    ours.grid_weight_render(dense, cam, 0.5, 1e-7, last_sample_opaque, off, scl, out)
    torch.cuda.synchronize()
    assert float(out.max()) > 0.1 and float(out.min()) >= 0.0
    ref = H.load_reference_cuda()
    if ref is not None:
        out_r = torch.zeros_like(dense)
        ref.grid_weight_render(dense, _camera(ref), 0.5, 1e-7, last_sample_opaque, off, scl, out_r)
        torch.cuda.synchronize()
        assert H.rel_err(out, out_r) < 2e-5     # nvcc contracts log_light += -world_step * sigma differently per kernel
        assert torch.equal(out > 0, out_r > 0)
    o, d = _oracle_rays(cam)
    out_o = np.zeros(tuple(dense.shape), np.float32)
    oracle.weight_render(dense.cpu(), None, dense.shape, sg.offset, sg.scaling, o, d, 0.5, 1e-7, last_sample_opaque, out_o)
    _few_differ(out.cpu().numpy(), out_o, 2e-3, "grid_weight_render vs oracle")


def test_sparse_grid_weight_render():
    from oracle import oracle
    sg = synth.make_shell_grid(48, basis_dim=4, variant="G", sigma_density=True).to("cuda")
    sg.density.mul_(0.05)          # thin medium: the transmittance stays well above the stop threshold through the shell
    cam = _camera(ours)
    off, scl = sg.offset.cuda(), sg.scaling.cuda()
    out = torch.zeros(tuple(sg.links.shape), device="cuda")
    ours.sparse_grid_weight_render(H.fill_grid_spec(ours, sg), cam, 0.5, 1e-7, off, scl, out)
    torch.cuda.synchronize()
    assert 0.0 < float(out[out > 0].min()) < 0.9 and float(out.max()) == 1.0
    ref = H.load_reference_cuda()
    if ref is not None:
        out_r = torch.zeros_like(out)
        ref.sparse_grid_weight_render(H.fill_grid_spec(ref, sg), _camera(ref), 0.5, 1e-7, off, scl, out_r)
        torch.cuda.synchronize()
        assert H.rel_err(out, out_r) < 2e-5     # nvcc contracts log_light += -world_step * sigma differently per kernel
        assert torch.equal(out > 0, out_r > 0)
    o, d = _oracle_rays(cam)
    out_o = np.zeros(tuple(out.shape), np.float32)
    oracle.weight_render(sg.density.cpu(), sg.links.cpu(), out.shape, sg.offset, sg.scaling, o, d, 0.5, 1e-7, False, out_o)
    _few_differ(out.cpu().numpy(), out_o, 2e-3, "sparse_grid_weight_render vs oracle")


@pytest.mark.parametrize("near_clip", [0.0, 30.0])
def test_sparse_grid_mask_render(near_clip):
    from oracle import oracle
    sg = synth.make_shell_grid(48, basis_dim=4, variant="G").to("cuda")
    o, d, _ = synth.make_camera_rays(3000, device="cuda", seed=21)
    mask = torch.zeros((sg.capacity,), device="cuda")
    ours.sparse_grid_mask_render(H.fill_grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), near_clip, mask)
    torch.cuda.synchronize()
    n = int((mask > 0).sum())
    assert 0 < n < sg.capacity and set(mask.unique().tolist()) <= {0.0, 1.0}
    ref = H.load_reference_cuda()
    if ref is not None:
        mask_r = torch.zeros_like(mask)
        ref.sparse_grid_mask_render(H.fill_grid_spec(ref, sg), H.fill_rays_spec(ref, o, d), near_clip, mask_r)
        torch.cuda.synchronize()
        assert torch.equal(mask, mask_r)
    mask_o = np.zeros((sg.capacity,), np.float32)
    oracle.mask_render(sg.links.cpu(), sg.offset, sg.scaling, o.cpu(), d.cpu(), near_clip, mask_o)
    assert int((mask.cpu().numpy() != mask_o).sum()) <= max(2, n // 500)


@pytest.mark.parametrize("variant", ["G", "G*"])
def test_sparse_grid_visibility_render_surf(variant):
    from oracle import oracle
    sg = synth.make_shell_grid(48, basis_dim=4, variant=variant).to("cuda")
    cam = _camera(ours)
    vis = torch.zeros((sg.capacity,), device="cuda")
    ours.sparse_grid_visbility_render_surf(H.fill_grid_spec(ours, sg), cam, vis)
    torch.cuda.synchronize()
    assert float(vis.max()) >= 1.0 and int((vis == 0).sum()) > 0       # the far side of the shell is occluded
    ref = H.load_reference_cuda()
    if ref is not None:
        # the reference marches on with an uninitialised tmax once a ray has stepped off the voxel range
        # (misc_kernel.cu:600-609) and then reads links out of bounds: compare on a narrow camera whose rays all end on the
        # surface (the level-set sphere fills the view), where that code is never reached
        vis_n, vis_r = torch.zeros_like(vis), torch.zeros_like(vis)
        ours.sparse_grid_visbility_render_surf(H.fill_grid_spec(ours, sg), _camera(ours, f=320.0), vis_n)
        ref.sparse_grid_visbility_render_surf(H.fill_grid_spec(ref, sg), _camera(ref, f=320.0), vis_r)
        torch.cuda.synchronize()
        assert float(vis_r.max()) >= 1.0
        assert torch.equal(vis_n, vis_r)
    o, d = _oracle_rays(cam)
    vis_o = np.zeros((sg.capacity,), np.float32)
    oracle.visibility_surf(oracle.Grid(sg.to("cpu")), o, d, vis_o)
    assert int((vis.cpu().numpy() != vis_o).sum()) <= max(4, sg.capacity // 200)
    # accumulates: a second pass doubles every count
    ours.sparse_grid_visbility_render_surf(H.fill_grid_spec(ours, sg), cam, vis)
    torch.cuda.synchronize()
    assert int((vis.cpu().numpy() != 2 * vis_o).sum()) <= max(8, sg.capacity // 100)
