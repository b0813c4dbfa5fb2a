"""CPU: the oracle (oracle/*.c) against golden vectors generated from the reference's own pure-PyTorch renderer
(/root/reference/svox2/svox2.py:1596-2857, through oracle/gen_golden.py).  The reference ships no golden vectors for this
path; these fixtures are what pins the oracle.  Tolerance 1e-4 relative (north_star)."""
import glob
import os
import types

import numpy as np
import torch
import pytest

from alphasurf_b200 import synth
from oracle import oracle

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l0_sh*.npz")))


def _grid(z):
    return types.SimpleNamespace(links=z["links"], density=z["density"], surface=z["surface"], sh=z["sh"],
                                 level_set=z["level_set"], offset=z["offset"], scaling=z["scaling"],
                                 basis_dim=int(z["basis_dim"]), fake_sample_std=float(z["fake_sample_std"]),
                                 truncated_vol_render_a=float(z["truncated_vol_render_a"]))


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


def test_fixtures_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_l0(path):
    z = np.load(path)
    og = oracle.Grid(_grid(z))
    Q = z["origins"].shape[0]
    # the L0 backward is hard-wired to mean|rgb| (svox2.py:2817-2828) == fused call with rgb_gt = 0, l1 = 1, l2 = 0
    rgb, g = oracle.surf_trav_fused(og, synth.parity_render_options(), z["origins"], z["dirs"],
                                    np.zeros((Q, 3), np.float32), dict(lambda_l2=0.0, lambda_l1=1.0, l_dist_max_sample=64))
    assert (np.abs(z["rgb"] - 1.0).max(axis=1) > 1e-3).sum() > Q // 4, "fixture must contain rays that hit the shell"
    assert _rel(rgb, z["rgb"]) < 1e-4
    assert _rel(g.sh, z["grad_sh"]) < 1e-4
    assert _rel(g.density, z["grad_density"]) < 1e-4
    assert _rel(g.surface, z["grad_surface"]) < 1e-4
    assert _rel(g.fake_sample_std, z["grad_fake_sample_std"].reshape(-1)) < 1e-4


def test_forward_entry_equals_fused_forward():
    z = np.load(GOLD[0])
    og = oracle.Grid(_grid(z))
    rgb = oracle.surf_trav_forward(og, synth.parity_render_options(), z["origins"], z["dirs"])
    assert _rel(rgb, z["rgb"]) < 1e-4


def test_cubic_solver_known_roots():
    # (t-0.2)(t-0.5)(t-0.9) = t^3 - 1.6 t^2 + 0.73 t - 0.09 : three roots, ascending
    typ, st = oracle.cubic_solve([-0.09, 0.73, -1.6, 1.0])
    assert typ == 205
    np.testing.assert_allclose(st, [0.2, 0.5, 0.9], atol=1e-12)
    # (t-0.3)(t^2+1): one real root
    typ, st = oracle.cubic_solve([-0.3, 1.0, -0.3, 1.0])
    assert typ == 206 and abs(st[0] - 0.3) < 1e-12 and st[1] == -1 and st[2] == -1
    # quadratic (f3 ~ 0): roots 0.25, 0.75 ; linear ; none
    typ, st = oracle.cubic_solve([0.1875, -1.0, 1.0, 0.0])
    assert typ == 203
    np.testing.assert_allclose(st[:2], [0.25, 0.75], atol=1e-12)
    typ, st = oracle.cubic_solve([-0.5, 2.0, 0.0, 0.0])
    assert typ == 201 and abs(st[0] - 0.25) < 1e-15
    typ, st = oracle.cubic_solve([1.0, 0.0, 0.0, 0.0])
    assert typ == 200


def test_cubic_root_grad_matches_finite_differences():
    rng = np.random.default_rng(0)
    for fs in ([-0.09, 0.73, -1.6, 1.0], [-0.3, 1.0, -0.3, 1.0], [0.1875, -1.0, 1.0, 0.0], [-0.5, 2.0, 0.0, 0.0]):
        fs = np.array(fs, np.float64)
        typ, st = oracle.cubic_solve(fs)
        for rid in range(3):
            if st[rid] < 0:
                continue
            g = oracle.cubic_root_grad(typ, rid, fs)
            for k in range(4):
                if (typ in (201,) and k >= 2) or (typ in (202, 203) and k == 3):
                    continue
                h = 1e-7
                fp, fm = fs.copy(), fs.copy()
                fp[k] += h
                fm[k] -= h
                tp, sp = oracle.cubic_solve(fp)
                tm, sm = oracle.cubic_solve(fm)
                if tp != typ or tm != typ:
                    continue
                fd = (sp[rid] - sm[rid]) / (2 * h)
                assert abs(g[k] - fd) < 1e-4 * max(1.0, abs(fd)), (fs, rid, k, g[k], fd)


# ---- Plenoxels cuvol renderer ---------------------------------------------------------------------------------------------
GOLD_CUVOL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l0_cuvol_*.npz")))


def test_cuvol_fixtures_present():
    assert len(GOLD_CUVOL) >= 2


@pytest.mark.parametrize("path", GOLD_CUVOL, ids=[os.path.basename(p) for p in GOLD_CUVOL])
def test_cuvol_oracle_matches_reference_l0(path):
    """oracle/oracle_cuvol.c vs the reference's _volume_render_gradcheck_lerp (svox2/svox2.py:1215-1441): colours and the
    gradients of mean((rgb - gt)^2), the loss volume_render_cuvol_fused differentiates."""
    z = np.load(path)
    g = types.SimpleNamespace(links=z["links"], density=z["density"], surface=None, sh=z["sh"], level_set=None,
                              offset=z["offset"], scaling=z["scaling"], basis_dim=int(z["basis_dim"]), fake_sample_std=1.0,
                              truncated_vol_render_a=1.0)
    og = oracle.Grid(g)
    opts = synth.alphasurf_render_options()
    opts.update(sigma_thresh=0.0, stop_thresh=0.0)
    rgb, gr = oracle.cuvol_fused(og, opts, z["origins"], z["dirs"], z["rgb_gt"])
    Q = z["origins"].shape[0]
    assert (np.abs(z["rgb"] - 1.0).max(axis=1) > 1e-3).sum() > Q // 4, "fixture must contain rays that hit the shell"
    assert _rel(rgb, z["rgb"]) < 1e-4
    assert _rel(gr.sh, z["grad_sh"]) < 1e-4
    assert _rel(gr.density, z["grad_density"]) < 1e-4
    assert np.array_equal(oracle.cuvol_forward(og, opts, z["origins"], z["dirs"]), rgb)


def test_cuvol_skip_codes_do_not_change_the_image():
    """Negative links <= -2 announce empty 2^(k-1) blocks (accel_dist_prop); skipping them must leave colours unchanged up to
    the rounding of t (t += ceil(skip/step)*step instead of repeated t += step)."""
    sg = synth.make_shell_grid(32, basis_dim=4, variant="G", sigma_density=True, z_order=False)
    links = sg.links.clone()
    occ = (links >= 0)
    # block-level codes: an aligned 4^3 block with no stored vertex in its 6^3 neighbourhood gets code -3 (side 4)
    R = 32
    pad = torch.nn.functional.pad(occ.float()[None, None], (1, 1, 1, 1, 1, 1))[0, 0]
    blk = torch.nn.functional.max_pool3d(pad[None, None], kernel_size=6, stride=4)[0, 0]   # (8,8,8) over [4b-1, 4b+4]
    empty = (blk == 0)
    code = torch.where(empty, torch.tensor(-3), torch.tensor(-1)).to(torch.int32)
    code_full = code.repeat_interleave(4, 0).repeat_interleave(4, 1).repeat_interleave(4, 2)
    links2 = torch.where(occ, links, code_full)
    o, d, gt = synth.make_camera_rays(256, cam_radius=2.2)
    opts = synth.alphasurf_render_options()
    opts.update(sigma_thresh=1e-8, stop_thresh=1e-7)
    mk = lambda l: oracle.Grid(types.SimpleNamespace(links=l, density=sg.density, surface=None, sh=sg.sh, level_set=None,
                                                     offset=sg.offset, scaling=sg.scaling, basis_dim=4, fake_sample_std=1.0,
                                                     truncated_vol_render_a=1.0))
    a = oracle.cuvol_forward(mk(links), opts, o, d)
    b = oracle.cuvol_forward(mk(links2), opts, o, d)
    assert int((links2 < -1).sum()) > 1000
    assert _rel(b, a) < 1e-4 and (np.abs(a - 1.0).max(axis=1) > 1e-3).sum() > 50
