"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the reference's own pure-PyTorch renderer (L0).

Run in the development container (needs /root/reference):  python -m oracle.gen_golden
Each fixture holds the complete inputs (grid tensors, rays, options) and the reference outputs
(rgb, d(mean|rgb|... see below)/d(density, sh, surface)) so that the CPU tests can replay them without the reference.

The L0 renderer's backward is hard-wired to the loss  L = mean|rgb - 0| ... precisely: svox2.py:2817-2828 runs
``s = torch.abs(rgb - torch.zeros).mean(); s.backward()`` (lambda_l2 = 0, lambda_l1 = 1), which the CUDA semantics
express as the fused call with rgb_gt = 0, lambda_l1 = 1, lambda_l2 = 0.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from alphasurf_b200 import synth  # noqa: E402
from oracle import ref_l0  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, reso, basis_dim, n_rays, variant, seed
    ("l0_sh1_r24", 24, 4, 96, "G*", 1),
    ("l0_sh2_r20", 20, 9, 64, "G*", 2),
    ("l0_sh1_r32_G", 32, 4, 96, "G", 3),
]


def main():
    assert ref_l0.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for name, reso, bd, nr, variant, seed in CASES:
        torch.manual_seed(seed)
        sg = synth.make_shell_grid(reso, basis_dim=bd, variant=variant, seed=synth.SEED + seed, z_order=False)
        o, d, _ = synth.make_camera_rays(nr, seed=synth.SEED + 10 * seed, cam_radius=2.2)
        opts = synth.parity_render_options()
        res = ref_l0.render_l0(sg, opts, o, d, run_backward=True)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            links=sg.links.numpy(), density=sg.density.numpy(), surface=sg.surface.numpy(), sh=sg.sh.numpy(),
            level_set=sg.level_set.numpy(), offset=sg.offset.numpy(), scaling=sg.scaling.numpy(),
            basis_dim=np.int32(bd), fake_sample_std=np.float32(sg.fake_sample_std),
            truncated_vol_render_a=np.float32(sg.truncated_vol_render_a), origins=o.numpy(), dirs=d.numpy(),
            rgb=res["rgb"].numpy(), grad_density=res["grad_density"].numpy(), grad_sh=res["grad_sh"].numpy(),
            grad_surface=res["grad_surface"].numpy(), grad_fake_sample_std=res["grad_fake_sample_std"].numpy(),
            opts=np.array(repr(opts)))
        print(name, "rgb mean", float(res["rgb"].mean()), "N", sg.capacity)


CUVOL_CASES = [
    # name, reso, basis_dim, n_rays, seed
    ("l0_cuvol_sh1_r24", 24, 4, 96, 11),
    ("l0_cuvol_sh2_r20", 20, 9, 64, 12),
]


def cuvol_options():
    """Plenoxels options the L0 renderer honours (it has no sigma / stop thresholds and no skipping)."""
    o = synth.alphasurf_render_options()
    o.update(backend="cuvol", sigma_thresh=0.0, stop_thresh=0.0)
    return o


def main_cuvol():
    assert ref_l0.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for name, reso, bd, nr, seed in CUVOL_CASES:
        sg = synth.make_shell_grid(reso, basis_dim=bd, variant="G", seed=synth.SEED + seed, z_order=False, sigma_density=True)
        o, d, gt = synth.make_camera_rays(nr, seed=synth.SEED + 10 * seed, cam_radius=2.2)
        opts = cuvol_options()
        res = ref_l0.render_l0_cuvol(sg, opts, o, d, gt)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            links=sg.links.numpy(), density=sg.density.numpy(), sh=sg.sh.numpy(), offset=sg.offset.numpy(),
            scaling=sg.scaling.numpy(), basis_dim=np.int32(bd), origins=o.numpy(), dirs=d.numpy(), rgb_gt=gt.numpy(),
            rgb=res["rgb"].numpy(), grad_density=res["grad_density"].numpy(), grad_sh=res["grad_sh"].numpy(),
            opts=np.array(repr(opts)))
        print(name, "rgb mean", float(res["rgb"].mean()), "N", sg.capacity)


if __name__ == "__main__":
    if "--cuvol-only" not in sys.argv:
        main()
    main_cuvol()
