"""CPU: the oracle's scalar renders (oracle/oracle_surf_trav.c, trace_ray_scalar) on a grid with a known answer -- a sphere
of radius r0 voxels stored as its signed distance, seen from outside: the first-hit depth is the analytic ray / sphere
distance (to the trilinear interpolation error of the field), the normal is radial, every mode agrees on the hit."""
import numpy as np
import torch

from alphasurf_b200 import synth
from oracle import oracle


def _sphere_grid(R=32, r0=9.6):
    sg = synth.make_shell_grid(R, basis_dim=4, variant="G", shell_mid=r0 / R, shell_half=0.12)
    lin = torch.nonzero(sg.links.reshape(-1) >= 0).flatten()
    rows = sg.links.reshape(-1)[lin].long()
    x, y, z = lin // (R * R), (lin // R) % R, lin % R
    rad = torch.sqrt((x - R / 2.0) ** 2 + (y - R / 2.0) ** 2 + (z - R / 2.0) ** 2)
    sg.surface[rows, 0] = (rad - r0).float()
    sg.density[:] = 2.0     # alpha = sigmoid-like activation of a large value: well above any threshold used here
    return sg


def test_sphere_depth_normal():
    R, r0 = 32, 9.6
    sg = _sphere_grid(R, r0)
    opts = synth.alphasurf_render_options()
    g = torch.Generator().manual_seed(3)
    Q = 64
    o = torch.randn((Q, 3), generator=g)
    o = 2.5 * o / o.norm(dim=1, keepdim=True)
    aim = 0.2 * torch.randn((Q, 3), generator=g) * (r0 / (R / 2))
    d = aim - o
    d = d / d.norm(dim=1, keepdim=True)
    og = oracle.Grid(sg)
    depth = oracle.surf_trav_scalar(og, opts, o, d, "thresh_depth", 0.1)
    alpha = oracle.surf_trav_scalar(og, opts, o, d, "thresh_alpha", 0.1)
    normal = oracle.surf_trav_scalar(og, opts, o, d, "normal")
    expd = oracle.surf_trav_scalar(og, opts, o, d, "expected_term")
    mode = oracle.surf_trav_scalar(og, opts, o, d, "mode_term", 0.1)
    # analytic: sphere centre in world coordinates = (R/2 - offset) / scaling, radius r0 / scaling
    c = ((R / 2.0 - sg.offset) / sg.scaling).numpy()
    rw = r0 / float(sg.scaling[0])
    oc = o.numpy() - c
    b = (oc * d.numpy()).sum(1)
    disc = b * b - ((oc * oc).sum(1) - rw * rw)
    assert (disc > 0).all()
    want = -b - np.sqrt(disc)
    assert (depth > 0).all() and (alpha > 0.1).all()
    vox = 1.0 / float(sg.scaling[0])
    assert np.abs(depth - want).max() < 0.75 * vox, np.abs(depth - want).max() / vox
    hit = o.numpy() + depth[:, None] * d.numpy() - c
    hit /= np.linalg.norm(hit, axis=1, keepdims=True)
    nrm = normal / np.linalg.norm(normal, axis=1, keepdims=True)
    assert ((hit * nrm).sum(1) > 0.98).all()
    # the first sample carries the largest weight; the expected depth lies between the two crossings of the sphere
    np.testing.assert_allclose(mode, depth, rtol=1e-6)
    assert ((expd >= depth * (1 - 1e-6)) & (expd <= (-b + np.sqrt(disc)) + vox)).all()


def test_miss_and_thresholds():
    sg = _sphere_grid()
    opts = synth.alphasurf_render_options()
    og = oracle.Grid(sg)
    o = torch.tensor([[0.0, 0.0, 3.0], [0.0, 0.0, 3.0]])
    d = torch.tensor([[0.0, 0.0, 1.0], [0.02, 0.01, -1.0]])
    d = d / d.norm(dim=1, keepdim=True)
    for m in ("expected_term", "mode_term", "thresh_depth", "thresh_alpha", "normal"):
        out = oracle.surf_trav_scalar(og, opts, o, d, m, 0.0)
        assert np.abs(out[0]).max() == 0.0 and np.abs(out[1]).max() > 0.0, m
    # thresholds above every alpha: nothing is reported
    assert oracle.surf_trav_scalar(og, opts, o, d, "thresh_depth", 1.5).max() == 0.0
    assert oracle.surf_trav_scalar(og, opts, o, d, "mode_term", 1.5).max() == 0.0


def test_cuvol_depth_oracle_homogeneous_box():
    """oracle_cuvol_scalar on a fully stored grid of constant sigma (a homogeneous box): the thresholded depth is the
    distance to the box (to one step), the expected termination depth
    lies about one mean free path behind the entry and equals the re-compositing of the med-term samples, which are equally
    spaced by the world step."""
    R = 24
    links = torch.arange(R ** 3, dtype=torch.int32).reshape(R, R, R)
    sigma = 3.0
    dens = torch.full((R ** 3, 1), sigma)
    sh = torch.zeros((R ** 3, 3))
    gsz = torch.tensor([R, R, R], dtype=torch.float32)
    sg = synth.SynthGrid(links, dens, None, sh, None, 0.5 * gsz, 0.5 * gsz, 1)
    opts = synth.alphasurf_render_options()
    opts.update(backend="cuvol", sigma_thresh=1e-8, stop_thresh=1e-7, step_size=0.5)
    og = oracle.Grid(sg)
    # rays along -z from above the box, spread over its top face
    g = torch.Generator().manual_seed(5)
    xy = (torch.rand((32, 2), generator=g) - 0.5) * 1.2
    o = torch.cat([xy, torch.full((32, 1), 3.0)], dim=1)
    d = torch.tensor([[0.0, 0.0, -1.0]]).repeat(32, 1)
    # world <-> grid: p_grid = p * R/2 + R/2, the sampled volume spans grid [-0.5, R-0.5] -> world z in [-1 - 1/R, 1 - 1/R]
    top = 1.0 - 1.0 / R
    bottom = -1.0 - 1.0 / R
    d0 = 3.0 - top
    step_w = 0.5 / (R / 2.0)                      # step_size voxels in world units
    depth = oracle.cuvol_scalar(og, opts, o, d, "sigma_thresh", 1.0)
    assert np.abs(depth - d0).max() <= step_w * 1.01
    dep, sig = oracle.cuvol_scalar(og, opts, o, d, "med_term", 0.0, 6)
    np.testing.assert_allclose(np.diff(dep, axis=1), step_w, rtol=1e-4)
    assert np.abs(sig[:, 1:] - sigma).max() < 1e-5          # interior samples see the constant density
    e = oracle.cuvol_scalar(og, opts, o, d, "expected_term", 0.0)
    # the expected depth re-composited from the per-sample depths / sigmas of med_term: weight_k = T_k (1 - exp(-ws sigma_k))
    dep_all, sig_all = oracle.cuvol_scalar(og, opts, o, d, "med_term", 0.0, 80)
    att = np.exp(-step_w * sig_all.astype(np.float64))
    trans = np.cumprod(np.concatenate([np.ones((32, 1)), att[:, :-1]], axis=1), axis=1)
    want = (trans * (1.0 - att) * dep_all).sum(1)
    np.testing.assert_allclose(e, want, rtol=2e-5)
    assert ((e > d0) & (e < d0 + 3.0 / sigma)).all()           # about one mean free path (1/sigma) behind the entry
    m = oracle.cuvol_scalar(og, opts, o, d, "mode_term", 0.0)
    assert np.abs(m - dep[:, 1]).max() <= step_w * 1.01     # heaviest sample: the first interior one
