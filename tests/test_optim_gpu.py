"""RMSprop / SGD grid steps: ours vs the numpy oracle and (bit-exact) vs the reference CUDA build, in the three indexer
modes the reference dispatches (optim_kernel.cu:175-215): 0-dim tensor = all rows, bool mask, int64 row list."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _make(N, C, seed, touched=0.2):
    g = torch.Generator().manual_seed(seed)
    data = torch.randn((N, C), generator=g)
    rms = torch.rand((N, C), generator=g) * 0.1
    rms[torch.rand((N, C), generator=g) < 0.3] = 0.0          # first-touch branch (rms == 0)
    grad = torch.randn((N, C), generator=g) * 0.01
    mask = torch.rand((N,), generator=g) < touched
    return data, rms, grad, mask


def _indexers(mask, dev):
    return {"all": torch.empty((), device=dev), "mask": mask.to(dev), "index": torch.nonzero(mask).flatten().to(dev),
            "empty": torch.empty((0,), dtype=torch.bool, device=dev)}


@pytest.mark.parametrize("N,C", [(1000, 1), (1003, 1), (4097, 27), (333, 12)])
@pytest.mark.parametrize("mode", ["all", "mask", "index", "empty"])
def test_rmsprop(N, C, mode):
    from oracle import oracle
    dev = "cuda"
    data, rms, grad, mask = _make(N, C, 7 * N + C)
    args = dict(beta=0.95, lr=1e-2, eps=1e-8, minval=-0.5, lr_last=3e-3)
    d, r, g = data.to(dev), rms.to(dev), grad.to(dev)
    ours.rmsprop_step(d, r, g, _indexers(mask, dev)[mode], args["beta"], args["lr"], args["eps"], args["minval"], args["lr_last"])
    dn, rn, gn = data.numpy().copy(), rms.numpy().copy(), grad.numpy().copy()
    if mode != "empty":
        idx = {"all": None, "mask": mask.numpy(), "index": np.nonzero(mask.numpy())[0]}[mode]
        oracle.rmsprop_step(dn, rn, gn, idx, **args)
    np.testing.assert_allclose(d.cpu().numpy(), dn, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(r.cpu().numpy(), rn, rtol=2e-6, atol=1e-12)
    assert np.array_equal(g.cpu().numpy(), gn)
    ref = H.load_reference_cuda()
    if ref is not None:
        d2, r2, g2 = data.to(dev), rms.to(dev), grad.to(dev)
        ref.rmsprop_step(d2, r2, g2, _indexers(mask, dev)[mode], args["beta"], args["lr"], args["eps"], args["minval"], args["lr_last"])
        assert torch.equal(d, d2) and torch.equal(r, r2) and torch.equal(g, g2), "not bit-exact with the reference kernel"


@pytest.mark.parametrize("N,C", [(2049, 27), (1003, 1)])
@pytest.mark.parametrize("mode", ["all", "mask", "index", "empty"])
def test_sgd(mode, N, C):
    from oracle import oracle
    dev = "cuda"
    data, _, grad, mask = _make(N, C, 11)
    d, g = data.to(dev), grad.to(dev)
    ours.sgd_step(d, g, _indexers(mask, dev)[mode], 0.1, 0.02)
    dn, gn = data.numpy().copy(), grad.numpy().copy()
    if mode != "empty":
        idx = {"all": None, "mask": mask.numpy(), "index": np.nonzero(mask.numpy())[0]}[mode]
        oracle.sgd_step(dn, gn, idx, 0.1, 0.02)
    np.testing.assert_allclose(d.cpu().numpy(), dn, rtol=1e-6, atol=1e-7)
    assert np.array_equal(g.cpu().numpy(), gn)
    ref = H.load_reference_cuda()
    if ref is not None:
        d2, g2 = data.to(dev), grad.to(dev)
        ref.sgd_step(d2, g2, _indexers(mask, dev)[mode], 0.1, 0.02)
        assert torch.equal(d, d2) and torch.equal(g, g2)


def test_rmsprop_large_rows_no_overflow():
    """64-bit indexing: N*C beyond 2^31 elements is the reference's documented overflow (cuda_util.cuh:14); here a
    smaller stand-in that still crosses the 2^31 BYTE offset boundary."""
    dev = "cuda"
    N, C = 20_000_000, 27          # 540M elements, 2.16 GB per tensor
    data = torch.zeros((N, C), device=dev)
    rms = torch.zeros((N, C), device=dev)
    grad = torch.zeros((N, C), device=dev)
    mask = torch.zeros((N,), dtype=torch.bool, device=dev)
    rows = torch.tensor([0, 12345, N - 1], device=dev)
    mask[rows] = True
    grad[rows] = 1.0
    ours.rmsprop_step(data, rms, grad, mask, 0.9, 0.5, 1e-8, -1e9, 0.5)
    torch.cuda.synchronize()
    assert float(grad.abs().sum()) == 0.0
    exp = torch.zeros((N,), device=dev)
    exp[rows] = -0.5
    assert torch.allclose(data[:, 0], exp) and torch.allclose(data[:, C - 1], exp)
    assert int((data != 0).sum()) == 3 * C
