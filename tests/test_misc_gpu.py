"""accel_dist_prop (misc_kernel.cu:1022-1058): ours vs the numpy oracle and vs the UNMODIFIED reference kernel -- integer
work, so bit-exact -- on cubic, non-cubic and odd-sized grids, and idempotence at full size."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _links(shape, seed, fill=0.1):
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand(shape, generator=g) < fill
    # clustered occupancy: keep only blobs
    if min(shape) >= 5:
        blob = torch.nn.functional.avg_pool3d(occ.float()[None, None], 5, stride=1, padding=2)[0, 0] > 0.16
    else:
        blob = occ
    links = torch.full(shape, -1, dtype=torch.int32)
    links[blob] = torch.arange(int(blob.sum()), dtype=torch.int32)
    return links


@pytest.mark.parametrize("shape", [(32, 32, 32), (33, 47, 20), (64, 64, 64), (5, 9, 17), (1, 8, 8)])
def test_accel_dist_prop_bit_exact(shape):
    from oracle import oracle
    links = _links(shape, sum(shape))
    want = oracle.accel_dist_prop(links)
    a = links.clone().cuda()
    ours.accel_dist_prop(a)
    assert np.array_equal(a.cpu().numpy(), want)
    ref = H.load_reference_cuda()
    if ref is not None:
        b = links.clone().cuda()
        ref.accel_dist_prop(b)
        assert torch.equal(a, b)
    # applying it again changes nothing (codes are recomputed from the stored vertices only)
    c = a.clone()
    ours.accel_dist_prop(c)
    assert torch.equal(a, c)


def test_accel_dist_prop_full_size_and_cuvol_skips():
    """512^3 shell grid: codes are in [-10, -1], stored vertices untouched, and cuvol renders the same image with the codes
    (block skipping) as without (sample every step)."""
    sg = synth.make_shell_grid(512, basis_dim=1, variant="G", sigma_density=True).to("cuda")
    links = sg.links.clone()
    ours.accel_dist_prop(links)
    assert torch.equal(links[sg.links >= 0], sg.links[sg.links >= 0])
    neg = links[sg.links < 0]
    assert int(neg.max()) == -1 and int(neg.min()) >= -10 and int((neg < -5).sum()) > 0
    from tests.test_cuvol_gpu import _grid_spec, plenoxels_options
    opts = plenoxels_options()
    o, d, _ = synth.make_camera_rays(4096, device="cuda")
    a = ours.volume_render_cuvol(_grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts))
    b = ours.volume_render_cuvol(_grid_spec(ours, sg, links), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts))
    assert H.rel_err(b, a) < 1e-4 and float((a - 1.0).abs().max()) > 1e-2
