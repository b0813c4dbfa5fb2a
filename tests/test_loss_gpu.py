"""Grid-side regularisers (loss_kernel.cu): our CUDA kernels through the svox2.csrc-compatible API vs the CPU oracle
(oracle/oracle_loss.c) and vs the UNMODIFIED reference kernels (oracle/_ref) on the same inputs.

Cell lists follow the reference's callers (svox2/svox2.py:4950-5163, :5690-5724): int32 flat cell ids, either a
contiguous run or a random subset; the mask is the bool "sparse grad indexer" of the grid.
"""
import numpy as np
import pytest
import torch

from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 2e-5   # fp32 atomics in a different order; values are otherwise the same formulas


def _grid(reso, bd=4, variant="G*", z_order=None):
    return synth.make_shell_grid(reso, basis_dim=bd, variant=variant, z_order=z_order)


def _cells(sg, frac, seed, contiguous):
    n = sg.links.numel()
    k = max(int(n * frac), 1)
    g = torch.Generator().manual_seed(seed)
    if contiguous:
        start = int(torch.randint(0, n, (1,), generator=g))
        c = (torch.arange(start, start + k) % n)
    else:
        c = torch.randint(0, n, (k,), generator=g)
    return c.to(torch.int32)


def _close(a, b, what):
    e = H.rel_err(a.cpu(), torch.as_tensor(b))
    assert e < TOL, (what, e)


@pytest.mark.parametrize("reso,ignore_edge", [(24, False), (24, True), (33, True)])
def test_tv_and_tv_grad_dense(reso, ignore_edge):
    from oracle import oracle
    sg = _grid(reso)
    links, sh = sg.links.cuda(), sg.sh.cuda()
    D = sh.shape[1]
    tv = ours.tv(links, sh, 1, D, False, 2.0, ignore_edge, -1.0, -1.0)
    want = oracle.tv(sg.links, sg.sh, 1, D, ignore_edge)
    assert abs(float(tv) - want) < 1e-5 * abs(want)
    grad = torch.zeros_like(sh)
    ours.tv_grad(links, sh, 1, D, 0.37, False, 2.0, ignore_edge, -1.0, -1.0, grad)
    g_o = np.zeros(tuple(sg.sh.shape), np.float32)
    oracle.tv_grad(sg.links, sg.sh, 1, D, 0.37, ignore_edge, g_o)
    _close(grad, g_o, "tv_grad")
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r = torch.zeros_like(sh)
        ref.tv_grad(links, sh, 1, D, 0.37, False, 2.0, ignore_edge, -1.0, -1.0, g_r)
        _close(grad, g_r.cpu(), "tv_grad vs reference CUDA")
        # The VALUE of the reference's tv() is not compared: its kernel lets threads return before the cub::BlockReduce
        # (ignore_edge at loss_kernel.cu:89, and CUDA_GET_THREAD_ID for the tail of the last block), so exited threads leave
        # stale partials in the reduction -- seen on B200: 2-3 % off from run to run, and NaN.  The value is checked
        # against the oracle above.


@pytest.mark.parametrize("what", ["density", "sh"])
@pytest.mark.parametrize("contiguous", [True, False])
@pytest.mark.parametrize("ignore_edge,ignore_last_z", [(False, False), (True, False), (False, True)])
def test_tv_grad_sparse(what, contiguous, ignore_edge, ignore_last_z):
    from oracle import oracle
    sg = _grid(28)
    data_c = sg.density if what == "density" else sg.sh
    s, e = (0, 1) if what == "density" else (1, data_c.shape[1])
    cells_c = _cells(sg, 0.3, 5, contiguous)
    links, data, cells = sg.links.cuda(), data_c.cuda(), cells_c.cuda()
    grad = torch.zeros_like(data)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.tv_grad_sparse(links, data, cells, mask, s, e, 0.81, False, 2.0, ignore_edge, ignore_last_z, -1.0, -1.0, grad)
    g_o = np.zeros(tuple(data_c.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.tv_grad_sparse(sg.links, data_c, None, cells_c, m_o, s, e, 0.81, ignore_edge, 0.0, ignore_last_z, False, False, g_o)
    _close(grad, g_o, "tv_grad_sparse")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r, m_r = torch.zeros_like(data), torch.zeros_like(mask)
        ref.tv_grad_sparse(links, data, cells, m_r, s, e, 0.81, False, 2.0, ignore_edge, ignore_last_z, -1.0, -1.0, g_r)
        _close(grad, g_r.cpu(), "tv_grad_sparse vs reference CUDA")
        assert torch.equal(mask, m_r)


def test_tv_grad_sparse_without_mask():
    """An empty mask tensor means "no mask" (loss_kernel.cu:1368)."""
    sg = _grid(20)
    links, data, cells = sg.links.cuda(), sg.density.cuda(), _cells(sg, 0.5, 1, False).cuda()
    g1, g2 = torch.zeros_like(data), torch.zeros_like(data)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.tv_grad_sparse(links, data, cells, mask, 0, 1, 1.0, False, 2.0, False, False, -1.0, -1.0, g1)
    ours.tv_grad_sparse(links, data, cells, torch.empty((0,), dtype=torch.bool, device="cuda"), 0, 1, 1.0, False, 2.0,
                        False, False, -1.0, -1.0, g2)
    assert H.rel_err(g1, g2) < TOL
    assert int(mask.sum()) > 0


@pytest.mark.parametrize("alpha_dependency", [False, True])
@pytest.mark.parametrize("ignore_edge,edge_value", [(True, -1.0), (False, -1.0), (False, 0.5)])
def test_surf_tv_grad_sparse(alpha_dependency, ignore_edge, edge_value):
    from oracle import oracle
    sg = _grid(28, variant="G")
    dens_c = sg.density * 0.2   # so that some max-alpha values fall under the 0.1 up-weighting threshold
    cells_c = _cells(sg, 1.0, 3, True)
    links, surf, dens, cells = sg.links.cuda(), sg.surface.cuda(), dens_c.cuda(), cells_c.cuda()
    grad = torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.surf_tv_grad_sparse(links, surf, dens, cells, mask, 0, 1, 1e-3, ignore_edge, edge_value, False, -1.0, -1.0,
                             alpha_dependency, grad)
    g_o = np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.tv_grad_sparse(sg.links, sg.surface, dens_c, cells_c, m_o, 0, 1, 1e-3, ignore_edge, edge_value, False,
                          alpha_dependency, True, g_o)
    _close(grad, g_o, "surf_tv_grad_sparse")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r, m_r = torch.zeros_like(surf), torch.zeros_like(mask)
        ref.surf_tv_grad_sparse(links, surf, dens, cells, m_r, 0, 1, 1e-3, ignore_edge, edge_value, False, -1.0, -1.0,
                                alpha_dependency, g_r)
        _close(grad, g_r.cpu(), "surf_tv_grad_sparse vs reference CUDA")
        assert torch.equal(mask, m_r)


@pytest.mark.parametrize("surf_decrease", [False, True])
def test_alpha_surf_sparsify(surf_decrease):
    from oracle import oracle
    sg = _grid(28, variant="G")
    dens_c = sg.density - 0.45   # mix of positive / negative raw alpha around the bounds
    cells_c = _cells(sg, 0.4, 9, False)
    links, surf, dens, cells = sg.links.cuda(), sg.surface.cuda(), dens_c.cuda(), cells_c.cuda()
    ga, gs = torch.zeros_like(dens), torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    args = (1e-3, 2e-3, surf_decrease, 0.15, 0.0, -0.1)
    ours.alpha_surf_sparsify_grad_sparse(links, dens, surf, cells, mask, *args, ga, gs)
    ga_o, gs_o = np.zeros(tuple(dens_c.shape), np.float32), np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.alpha_surf_sparsify(sg.links, dens_c, sg.surface, cells_c, m_o, *args, ga_o, gs_o)
    _close(ga, ga_o, "sparsify grad_alpha")
    _close(gs, gs_o, "sparsify grad_surf")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    ref = H.load_reference_cuda()
    if ref is not None:
        ga_r, gs_r, m_r = torch.zeros_like(dens), torch.zeros_like(surf), torch.zeros_like(mask)
        ref.alpha_surf_sparsify_grad_sparse(links, dens, surf, cells, m_r, *args, ga_r, gs_r)
        _close(ga, ga_r.cpu(), "sparsify grad_alpha vs reference CUDA")
        _close(gs, gs_r.cpu(), "sparsify grad_surf vs reference CUDA")
        assert torch.equal(mask, m_r)


@pytest.mark.parametrize("con_check,ignore_empty,use_l1", [(False, False, True), (True, False, False), (True, True, True),
                                                          (False, True, False)])
def test_surface_normal_grad_sparse(con_check, ignore_empty, use_l1):
    from oracle import oracle
    sg = _grid(28, variant="G*")
    cells_c = _cells(sg, 1.0, 2, True)
    links, surf, cells = sg.links.cuda(), sg.surface.cuda(), cells_c.cuda()
    lv = float(sg.level_set[0])
    grad = torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.surface_normal_grad_sparse(links, surf, cells, mask, lv, 0, 1, 1e-2, 0.0, -1.0, -1.0, con_check, ignore_empty,
                                    use_l1, grad)
    g_o = np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.surface_normal_grad_sparse(sg.links, sg.surface, cells_c, m_o, lv, 0, 1, 1e-2, con_check, ignore_empty, use_l1, g_o)
    _close(grad, g_o, "surface_normal_grad_sparse")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r, m_r = torch.zeros_like(surf), torch.zeros_like(mask)
        ref.surface_normal_grad_sparse(links, surf, cells, m_r, lv, 0, 1, 1e-2, 0.0, -1.0, -1.0, con_check, ignore_empty,
                                       use_l1, g_r)
        _close(grad, g_r.cpu(), "surface_normal_grad_sparse vs reference CUDA")
        assert torch.equal(mask, m_r)


@pytest.mark.parametrize("reso", [32, 28])        # power-of-two sizes decode the flat cell id by shifts
def test_surface_normal_exact_zero_contributions(reso):
    """Where neighbouring cells have IDENTICAL normals the loss direction is exactly zero and the reference's
    `val != 0` test leaves the rows unmarked: a field that is linear in half of the grid and noisy in the other half
    exercises both the all-non-zero fast path and the exact per-corner path of the mask."""
    from oracle import oracle
    sg = _grid(reso, variant="G")
    lin = torch.nonzero(sg.links.reshape(-1) >= 0).flatten()
    rows = sg.links.reshape(-1)[lin].long()
    x, y, z = lin // (reso * reso), (lin // reso) % reso, lin % reso
    flat = (0.25 * x + 0.5 * y - 0.125 * z).float()          # exactly representable: all cell normals identical
    surf_c = sg.surface.clone()
    half = x < reso // 2
    surf_c[rows[half], 0] = flat[half]
    cells_c = _cells(sg, 1.0, 2, True)
    links, surf, cells = sg.links.cuda(), surf_c.cuda(), cells_c.cuda()
    for use_l1 in (True, False):
        grad = torch.zeros_like(surf)
        mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
        ours.surface_normal_grad_sparse(links, surf, cells, mask, 0.0, 0, 1, 1e-2, 0.0, -1.0, -1.0, False, False, use_l1, grad)
        g_o = np.zeros(tuple(surf_c.shape), np.float32)
        m_o = np.zeros((sg.capacity,), np.uint8)
        oracle.surface_normal_grad_sparse(sg.links, surf_c, cells_c, m_o, 0.0, 0, 1, 1e-2, False, False, use_l1, g_o)
        _close(grad, g_o, "normal loss with exact zeros")
        assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
        assert 0 < int(m_o.sum()) < int((sg.links >= 0).sum())      # part of the rows stays unmarked
        ref = H.load_reference_cuda()
        if ref is not None:
            g_r, m_r = torch.zeros_like(surf), torch.zeros_like(mask)
            ref.surface_normal_grad_sparse(links, surf, cells, m_r, 0.0, 0, 1, 1e-2, 0.0, -1.0, -1.0, False, False, use_l1, g_r)
            _close(grad, g_r.cpu(), "normal loss with exact zeros vs reference CUDA")
            assert torch.equal(mask, m_r)


def test_loss_kernels_at_full_size_properties():
    """512^3-sized property checks: the sparse TV gradient over ALL cells equals the dense TV gradient (same formula,
    scale/nl vs scale/n_cells normalisation accounted for), and gradients of a constant field vanish."""
    R = 256
    sg = synth.make_shell_grid(R, basis_dim=1, variant="G").to("cuda")
    n = sg.links.numel()
    # cells of the dense kernel: x,y,z < R-1
    ar = torch.arange(R - 1, device="cuda", dtype=torch.int32)
    cells = ((ar[:, None, None] * R + ar[None, :, None]) * R + ar[None, None, :]).reshape(-1).contiguous()
    g_d, g_s = torch.zeros_like(sg.density), torch.zeros_like(sg.density)
    ours.tv_grad(sg.links, sg.density, 0, 1, 1.0, False, 2.0, False, -1.0, -1.0, g_d)
    ours.tv_grad_sparse(sg.links, sg.density, cells, torch.empty((0,), dtype=torch.bool, device="cuda"), 0, 1, 1.0, False,
                        2.0, False, False, -1.0, -1.0, g_s)
    assert H.rel_err(g_s, g_d) < 1e-4
    const = torch.full_like(sg.density, 0.7)
    g_c = torch.zeros_like(const)
    stored = torch.where(sg.links.view(-1) >= 0)[0].int()    # cells whose own vertex exists: missing neighbours copy it
    ours.tv_grad_sparse(sg.links, const, stored, torch.empty((0,), dtype=torch.bool, device="cuda"), 0, 1, 1.0, False, 2.0,
                        True, False, -1.0, -1.0, g_c)
    assert float(g_c.abs().max()) == 0.0
    assert n == R ** 3


def _tile_path(enabled):
    from alphasurf_b200 import capi
    capi.lib().asurf_debug_set_normal_tile(1 if enabled else 0)


def _window(sg, frac, seed):
    """the list svox2.py:6354-6372 builds: all stored vertices in ascending order, or a contiguous window of that list"""
    ne = torch.where(sg.links.view(-1) >= 0)[0].int()
    if frac >= 1.0:
        return ne
    n = int(ne.shape[0] * frac)
    start = int(torch.randint(0, ne.shape[0] - n + 1, (1,), generator=torch.Generator().manual_seed(seed)))
    return ne[start:start + n].contiguous()


# shell_half = 1.0: every vertex stored (the far faces of the grid are reached: the reference's link-0 reads, :761-763);
# shell_half = 0.25: the shell touches the faces but not the corners
TILE_GRIDS = [(40, "G*", None, 0.05), (64, "G", None, 0.05), (33, "G*", False, 0.05), (24, "G*", False, 1.0), (36, "G", False, 0.25)]


@pytest.mark.parametrize("reso,variant,z_order,shell_half", TILE_GRIDS)
@pytest.mark.parametrize("con_check,ignore_empty,use_l1", [(False, False, True), (True, True, False)])
@pytest.mark.parametrize("frac", [1.0, 0.37])
def test_surface_normal_window_lists_take_the_tile_path(reso, variant, z_order, shell_half, con_check, ignore_empty, use_l1,
                                                        frac):
    """norm_surface_sparsity = 1 hands the kernel every stored vertex in ascending order, a smaller fraction a contiguous
    window of that list (svox2.py:6354-6372); the library recognises both on the device and runs the tiled kernel.  Same
    result as the oracle / the reference kernel, and as our list kernel (tile path switched off)."""
    from oracle import oracle
    sg = synth.make_shell_grid(reso, basis_dim=4, variant=variant, z_order=z_order, shell_half=shell_half)
    cells_c = _window(sg, frac, 11)
    links, surf, cells = sg.links.cuda(), sg.surface.cuda(), cells_c.cuda()
    lv = float(sg.level_set[0])
    args = (lv, 0, 1, 1e-2, 0.0, -1.0, -1.0, con_check, ignore_empty, use_l1)
    grad = torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.surface_normal_grad_sparse(links, surf, cells, mask, *args, grad)
    g_o = np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.surface_normal_grad_sparse(sg.links, sg.surface, cells_c, m_o, lv, 0, 1, 1e-2, con_check, ignore_empty, use_l1, g_o)
    _close(grad, g_o, "normal loss, tile path vs oracle")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    g_p, m_p = torch.zeros_like(surf), torch.zeros_like(mask)
    _tile_path(False)
    try:
        ours.surface_normal_grad_sparse(links, surf, cells, m_p, *args, g_p)
    finally:
        _tile_path(True)
    assert H.rel_err(grad, g_p) < TOL and torch.equal(mask, m_p)
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r, m_r = torch.zeros_like(surf), torch.zeros_like(mask)
        ref.surface_normal_grad_sparse(links, surf, cells, m_r, *args, g_r)
        _close(grad, g_r.cpu(), "normal loss, tile path vs reference CUDA")
        assert torch.equal(mask, m_r)


@pytest.mark.parametrize("reso,variant,z_order,shell_half", TILE_GRIDS)
@pytest.mark.parametrize("ignore_edge,edge_value,ignore_last_z", [(True, -1.0, False), (False, -1.0, False), (False, 0.5, True)])
@pytest.mark.parametrize("frac", [1.0, 0.37])
def test_surf_tv_window_lists_take_the_tile_path(reso, variant, z_order, shell_half, ignore_edge, edge_value, ignore_last_z, frac):
    """tv_surface_sparsity = 1 / a window of the stored vertices: tiled surface TV vs the oracle, the reference kernel and
    our list kernel; includes grids whose stored vertices reach the far faces (out-of-range neighbours read link 0 there)."""
    from oracle import oracle
    sg = synth.make_shell_grid(reso, basis_dim=4, variant=variant, z_order=z_order, shell_half=shell_half)
    cells_c = _window(sg, frac, 12)
    links, surf, dens, cells = sg.links.cuda(), sg.surface.cuda(), sg.density.cuda(), cells_c.cuda()
    args = (0, 1, 1e-3, ignore_edge, edge_value, ignore_last_z, -1.0, -1.0, False)
    grad = torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.surf_tv_grad_sparse(links, surf, dens, cells, mask, *args, grad)
    g_o = np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.tv_grad_sparse(sg.links, sg.surface, sg.density, cells_c, m_o, 0, 1, 1e-3, ignore_edge, edge_value, ignore_last_z,
                          False, True, g_o)
    _close(grad, g_o, "surface TV, tile path vs oracle")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o)
    g_p, m_p = torch.zeros_like(surf), torch.zeros_like(mask)
    _tile_path(False)
    try:
        ours.surf_tv_grad_sparse(links, surf, dens, cells, m_p, *args, g_p)
    finally:
        _tile_path(True)
    assert H.rel_err(grad, g_p) < TOL and torch.equal(mask, m_p)
    ref = H.load_reference_cuda()
    if ref is not None:
        g_r, m_r = torch.zeros_like(surf), torch.zeros_like(mask)
        ref.surf_tv_grad_sparse(links, surf, dens, cells, m_r, *args, g_r)
        _close(grad, g_r.cpu(), "surface TV, tile path vs reference CUDA")
        assert torch.equal(mask, m_r)


def test_lists_that_are_not_windows_fall_back_to_the_list_kernels():
    """A window with one entry removed / duplicated / unordered is not the set of stored vertices between its ends: the
    device-side check must send it to the list kernels (same result as the oracle)."""
    from oracle import oracle
    sg = synth.make_shell_grid(40, basis_dim=4, variant="G*")
    ne = _window(sg, 1.0, 0)
    n = ne.shape[0]
    variants = {
        "hole": torch.cat([ne[:n // 2], ne[n // 2 + 1:]]),
        "duplicate": torch.cat([ne[:n // 2], ne[n // 2 - 1:]]),
        "swapped": torch.cat([ne[:7], ne[8:9], ne[7:8], ne[9:]]),
    }
    links, surf = sg.links.cuda(), sg.surface.cuda()
    for name, cells_c in variants.items():
        cells_c = cells_c.contiguous()
        grad = torch.zeros_like(surf)
        mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
        ours.surface_normal_grad_sparse(links, surf, cells_c.cuda(), mask, 0.0, 0, 1, 1e-2, 0.0, -1.0, -1.0, False, False, True, grad)
        g_o = np.zeros(tuple(sg.surface.shape), np.float32)
        m_o = np.zeros((sg.capacity,), np.uint8)
        oracle.surface_normal_grad_sparse(sg.links, sg.surface, cells_c, m_o, 0.0, 0, 1, 1e-2, False, False, True, g_o)
        _close(grad, g_o, "normal loss, %s list" % name)
        assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o), name
        grad = torch.zeros_like(surf)
        mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
        ours.surf_tv_grad_sparse(links, surf, sg.density.cuda(), cells_c.cuda(), mask, 0, 1, 1e-3, True, -1.0, False, -1.0, -1.0,
                                 False, grad)
        g_o = np.zeros(tuple(sg.surface.shape), np.float32)
        m_o = np.zeros((sg.capacity,), np.uint8)
        oracle.tv_grad_sparse(sg.links, sg.surface, sg.density, cells_c, m_o, 0, 1, 1e-3, True, -1.0, False, False, True, g_o)
        _close(grad, g_o, "surface TV, %s list" % name)
        assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o), name


@pytest.mark.parametrize("contiguous", [True, False])
def test_surf_sign_change_grad_sparse(contiguous):
    """Against the oracle only: the reference kernel's loop counter is uninitialised (loss_kernel.cu:944, undefined
    behaviour), both restatements implement the documented intent."""
    from oracle import oracle
    sg = _grid(32, variant="G*")
    cells_c = _cells(sg, 0.6, 4, contiguous)
    links, surf, cells = sg.links.cuda(), sg.surface.cuda(), cells_c.cuda()
    grad = torch.zeros_like(surf)
    mask = torch.zeros((sg.capacity,), dtype=torch.bool, device="cuda")
    ours.surf_sign_change_grad_sparse(links, surf, cells, mask, 0, 1, 0.3, grad)
    g_o = np.zeros(tuple(sg.surface.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.surf_sign_change_grad_sparse(sg.links, sg.surface, cells_c, m_o, 0, 1, 0.3, g_o)
    _close(grad, g_o, "surf_sign_change_grad_sparse")
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), m_o) and 0 < m_o.sum() < sg.capacity


@pytest.mark.parametrize("reso,cols,lv", [(24, (0, 1), 0.0), (20, (1, 3), 0.1)])
def test_surface_normal_grad_dense(reso, cols, lv):
    """dense variant over the whole lattice (loss_kernel.cu:1289-1325): vs the oracle and the reference kernel, on a
    single-column tensor (the surface) and on two columns of a wider one"""
    from oracle import oracle
    ref = H.load_reference_cuda()
    sg = _grid(reso, variant="G*")
    data_c = sg.surface if cols == (0, 1) else (sg.sh[:, :4] * 0.5 + sg.surface).contiguous()
    links, data = sg.links.cuda(), data_c.cuda()
    grad = torch.zeros_like(data)
    ours.surface_normal_grad(links, data, lv, cols[0], cols[1], 0.7, -1.0, -1.0, grad)
    g_o = np.zeros(tuple(data_c.shape), np.float32)
    oracle.surface_normal_grad(sg.links, data_c, lv, cols[0], cols[1], 0.7, g_o)
    _close(grad, g_o, "surface_normal_grad vs oracle")
    assert np.abs(g_o).max() > 0
    g_r = torch.zeros_like(data)
    ref.surface_normal_grad(links, data, lv, cols[0], cols[1], 0.7, -1.0, -1.0, g_r)
    _close(grad, g_r.cpu().numpy(), "surface_normal_grad vs reference kernel")
    untouched = [c for c in range(data.shape[1]) if not cols[0] <= c < cols[1]]
    assert float(grad[:, untouched].abs().max()) == 0.0 if untouched else True


@pytest.mark.parametrize("bd,dir_factor,with_mask", [(9, 1.0, True), (4, 0.0, True), (9, 0.5, False)])
def test_lumisphere_tv_grad_sparse(bd, dir_factor, with_mask):
    """loss_kernel.cu:1661-1697 through the GridSpec / GridOutputGrads objects, vs the oracle and the reference kernel.
    Cell ids are decoded on the (size - 1) lattice (:1092-1096); row 0's cell is skipped (:1110)."""
    from oracle import oracle
    ref = H.load_reference_cuda()
    reso = 24
    sg = _grid(reso, bd=bd, variant="G*").to("cuda")
    n_cells = (reso - 1) ** 3
    cells = torch.randint(0, n_cells, (n_cells // 2,), generator=torch.Generator().manual_seed(6)).int().cuda()
    g = torch.Generator().manual_seed(8)
    sv, su = torch.randn(bd, generator=g).cuda(), torch.randn(bd, generator=g).cuda()
    res = {}
    for name, mod in (("ours", ours), ("ref", ref)):
        grid = H.fill_grid_spec(mod, sg)
        G = H.GradSet(sg, "cuda")
        holder = mod.GridOutputGrads()
        holder.grad_sh_out = G.sh
        if with_mask:
            holder.mask_out = G.mask
        mod.lumisphere_tv_grad_sparse(grid, cells, sv, su, 0.4, -1.0, -1.0, dir_factor, holder)
        torch.cuda.synchronize()
        res[name] = G
    g_o = np.zeros(tuple(sg.sh.shape), np.float32)
    m_o = np.zeros((sg.capacity,), np.uint8)
    oracle.lumisphere_tv_grad_sparse(sg.links, sg.sh, bd, cells, sv, su, 0.4, dir_factor, m_o if with_mask else None, g_o)
    _close(res["ours"].sh, g_o, "lumisphere vs oracle")
    _close(res["ours"].sh, res["ref"].sh.cpu().numpy(), "lumisphere vs reference kernel")
    assert np.abs(g_o).max() > 0
    if with_mask:
        assert np.array_equal(res["ours"].mask.cpu().numpy().astype(np.uint8), m_o) and m_o.sum() > 0
        assert torch.equal(res["ours"].mask, res["ref"].mask)
