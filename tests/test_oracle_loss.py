"""CPU: the loss oracle (oracle/oracle_loss.c) against an independent vectorised restatement of the same formulas
(loss_kernel.cu:119-184, :664-807) and against its own size-independent properties.  The reference ships neither
vectors nor a CPU implementation of these kernels; tests/test_loss_gpu.py pins the oracle on the UNMODIFIED reference
kernels on the GPU box."""
import numpy as np
import torch

from alphasurf_b200 import synth
from oracle import oracle


def _tv_grad_torch(links, data, start, end, scale, ignore_edge):
    """Dense TV gradient with autograd-free tensor algebra (float64 accumulate)."""
    X, Y, Z = links.shape
    sc = [s / 256.0 for s in (X, Y, Z)]
    nl = (X - 1) * (Y - 1) * (Z - 1)
    l000 = links[:-1, :-1, :-1].long()
    nbr = [links[1:, :-1, :-1].long(), links[:-1, 1:, :-1].long(), links[:-1, :-1, 1:].long()]
    grad = torch.zeros(data.shape, dtype=torch.float64)
    for idx in range(start, end):
        col = data[:, idx]
        v000 = torch.where(l000 >= 0, col[l000.clamp_min(0)], torch.zeros(()))
        skip = (l000 == 0) if ignore_edge else torch.zeros_like(l000, dtype=torch.bool)
        d = []
        for l in nbr:
            v = torch.where(l >= 0, col[l.clamp_min(0)], v000 if ignore_edge else torch.zeros(()))
            d.append(v - v000)
        idelta = (scale / np.float32(nl)) * torch.rsqrt(1e-9 + d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        idelta = torch.where(skip, torch.zeros(()), idelta)
        tot = torch.zeros_like(idelta)
        for a in range(3):
            da = d[a] * sc[a]
            tot = tot + da
            ok = (nbr[a] >= 0) & (da != 0)
            grad[:, idx].index_add_(0, nbr[a][ok], (da * idelta)[ok].double())
        ok = l000 >= 0
        grad[:, idx].index_add_(0, l000[ok], (-(tot) * idelta)[ok].double())
    return grad


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_tv_grad_dense_matches_tensor_restatement():
    sg = synth.make_shell_grid(20, basis_dim=4, variant="G*")
    for ignore_edge in (False, True):
        g = np.zeros(tuple(sg.sh.shape), np.float32)
        oracle.tv_grad(sg.links, sg.sh, 1, 5, 0.5, ignore_edge, g)
        want = _tv_grad_torch(sg.links, sg.sh, 1, 5, 0.5, ignore_edge)
        assert _rel(g, want.numpy()) < 1e-5
        assert np.abs(g[:, 0]).max() == 0 and np.abs(g[:, 5:]).max() == 0   # only channels [1, 5) are touched


def test_sparse_over_all_cells_equals_dense():
    sg = synth.make_shell_grid(18, basis_dim=1, variant="G")
    R = 18
    ar = np.arange(R - 1)
    cells = ((ar[:, None, None] * R + ar[None, :, None]) * R + ar[None, None, :]).reshape(-1).astype(np.int32)
    gd, gs = np.zeros(tuple(sg.density.shape), np.float32), np.zeros(tuple(sg.density.shape), np.float32)
    oracle.tv_grad(sg.links, sg.density, 0, 1, 1.0, False, gd)
    mask = np.zeros((sg.capacity,), np.uint8)
    oracle.tv_grad_sparse(sg.links, sg.density, None, cells, mask, 0, 1, 1.0, False, 0.0, False, False, False, gs)
    assert _rel(gs, gd) < 1e-6
    assert mask.sum() > 0 and set(np.nonzero(mask)[0]) >= set(np.nonzero(gs[:, 0])[0])


def test_tv_value_and_constant_field():
    # fully linked 8^3 grid holding a constant: every difference is 0 -> tv = sqrt(1e-5) over the cells whose link != 0
    R = 8
    links = torch.arange(R ** 3, dtype=torch.int32).reshape(R, R, R)
    const = np.full((R ** 3, 1), 0.25, np.float32)
    nl = (R - 1) ** 3
    assert abs(oracle.tv(links, const, 0, 1, False) - np.sqrt(np.float32(1e-5))) < 1e-7
    assert abs(oracle.tv(links, const, 0, 1, True) - np.sqrt(np.float32(1e-5)) * (nl - 1) / nl) < 1e-7   # link 0 is skipped (sic)
    g = np.zeros_like(const)
    cells = np.arange(R ** 3, dtype=np.int32)
    oracle.tv_grad_sparse(links, const, None, cells, None, 0, 1, 1.0, True, 0.0, False, False, False, g)
    assert np.abs(g).max() == 0
    # a linear ramp along z: |dz| = slope everywhere, gradient cancels in the interior
    ramp = (np.arange(R ** 3) % R).astype(np.float32).reshape(-1, 1) * 0.5
    tv = oracle.tv(links, ramp, 0, 1, False)
    assert abs(tv - np.sqrt(np.float32(1e-5) + (0.5 * R / 256.0) ** 2)) < 1e-6


def test_sparsify_matches_tensor_restatement():
    sg = synth.make_shell_grid(16, basis_dim=1, variant="G")
    gen = torch.Generator().manual_seed(3)
    cells = torch.randint(0, sg.links.numel(), (5000,), generator=gen).to(torch.int32)
    alpha = (sg.density - 0.45).numpy()
    surf = sg.surface.numpy()
    ga, gs = np.zeros_like(alpha), np.zeros_like(surf)
    mask = np.zeros((sg.capacity,), np.uint8)
    oracle.alpha_surf_sparsify(sg.links, alpha, surf, cells, mask, 1e-3, 2e-3, False, 0.15, 0.0, -0.1, ga, gs)
    l = sg.links.reshape(-1)[cells.long()].numpy()
    l = l[l >= 0]
    want_a, want_s = np.zeros(alpha.shape[0], np.float64), np.zeros(alpha.shape[0], np.float64)
    a = alpha[l, 0]
    safe = 1.0 / np.maximum(a, np.float32(1e-8))
    np.add.at(want_a, l[a > 0.0], 1e-3 * safe[a > 0.0])
    sel = (surf[l, 0] < -0.1) & (a < 0.15)
    np.add.at(want_s, l[sel], -2e-3 * safe[sel])
    assert _rel(ga[:, 0], want_a) < 1e-5
    assert _rel(gs[:, 0], want_s) < 1e-5 or (np.abs(want_s).max() == 0 and np.abs(gs).max() == 0)
    assert np.array_equal(np.nonzero(mask)[0], np.unique(l))


def test_accel_dist_prop_oracle_properties():
    """numpy restatement of accel_dist_prop: stored vertices untouched; an empty vertex next to a stored one gets -1 when
    they share the 2^3 parent; a vertex whose whole 2^k block is empty gets at most -(k + 1); idempotent."""
    links = torch.full((16, 16, 16), -1, dtype=torch.int32)
    links[0, 0, 0] = 0
    links[9, 9, 9] = 1
    out = oracle.accel_dist_prop(links)
    assert out[0, 0, 0] == 0 and out[9, 9, 9] == 1
    assert out[1, 1, 1] == -1            # shares the 2^3 parent of (0,0,0)
    assert out[2, 2, 2] == -2            # parent 2^3 empty, 4^3 block holds (0,0,0)
    assert out[4, 4, 4] == -3            # 2^3 and 4^3 empty, 8^3 block [0,8) holds (0,0,0)
    assert out[15, 0, 0] == -4           # first occupied ancestor is the 16^3 root
    assert out[8, 8, 8] == -1 and out[10, 10, 10] == -2
    assert np.array_equal(oracle.accel_dist_prop(out), out)
    sg = synth.make_shell_grid(24, basis_dim=1, variant="G")
    o2 = oracle.accel_dist_prop(sg.links)
    assert np.array_equal(o2[sg.links.numpy() >= 0], sg.links.numpy()[sg.links.numpy() >= 0])
    assert o2[sg.links.numpy() < 0].max() == -1


def test_dense_normal_loss_is_the_gradient_of_its_energy():
    """oracle_surface_normal_grad (loss_kernel.cu:245-396) against autograd of the energy its expressions differentiate:
    sum over cells and used +x/+y/+z pairs of |n0/|n0| - n1/|n1||^2 / norm_count, times scale / n_lattice_cells."""
    R, lv = 20, 0.0
    sg = synth.make_shell_grid(R, basis_dim=1, variant="G*")
    links = sg.links.long()
    s = sg.surface[:, 0].double().clone().requires_grad_(True)
    ok8 = torch.ones((R - 1,) * 3, dtype=torch.bool)
    corner = {}
    for k in range(8):
        dx, dy, dz = k >> 2, (k >> 1) & 1, k & 1
        l = links[dx:R - 1 + dx, dy:R - 1 + dy, dz:R - 1 + dz]
        ok8 &= l >= 0
        corner[k] = s[l.clamp_min(0)]
    c = corner
    n = torch.stack([((c[4] + c[5] + c[6] + c[7]) - (c[0] + c[1] + c[2] + c[3])) / 4,
                     ((c[2] + c[3] + c[6] + c[7]) - (c[0] + c[1] + c[4] + c[5])) / 4,
                     ((c[1] + c[3] + c[5] + c[7]) - (c[0] + c[2] + c[4] + c[6])) / 4], -1)
    nh = n / torch.sqrt(1e-9 + (n * n).sum(-1, keepdim=True))

    def connected(vals):
        v = torch.stack(vals, -1).detach()
        return ~((v <= lv).all(-1) | (v >= lv).all(-1))
    faces = [[c[4], c[5], c[6], c[7]], [c[2], c[3], c[6], c[7]], [c[1], c[3], c[5], c[7]]]
    use, energy = [], []
    for a in range(3):
        sl_hi = [slice(None)] * 3
        sl_lo = [slice(None)] * 3
        sl_hi[a], sl_lo[a] = slice(1, None), slice(0, -1)
        u = torch.zeros_like(ok8)
        u[tuple(sl_lo)] = ok8[tuple(sl_lo)] & ok8[tuple(sl_hi)] & connected(faces[a])[tuple(sl_lo)]
        e = torch.zeros(ok8.shape, dtype=torch.float64)
        e[tuple(sl_lo)] = ((nh[tuple(sl_lo)] - nh[tuple(sl_hi)]) ** 2).sum(-1)
        use.append(u)
        energy.append(e)
    cnt = (use[0].long() + use[1].long() + use[2].long()).clamp_min(1)
    total = sum((energy[a] * use[a] / cnt).sum() for a in range(3)) * (0.7 / (R - 1) ** 3)
    total.backward()
    g = np.zeros(tuple(sg.surface.shape), np.float32)
    oracle.surface_normal_grad(sg.links, sg.surface, lv, 0, 1, 0.7, g)
    assert np.abs(g).max() > 0 and _rel(g[:, 0], s.grad.numpy()) < 2e-5


def test_lumisphere_tv_is_the_gradient_of_its_energy():
    """oracle_lumisphere_tv_grad_sparse (loss_kernel.cu:1067-1177) against autograd of
    scale / n * sum_cells sum_colour sqrt(1e-9 + dx^2 + dy^2 + dz^2 + du^2) over a list of DISTINCT cells."""
    R, bd = 20, 4
    sg = synth.make_shell_grid(R, basis_dim=bd, variant="G*")
    links = sg.links.long()
    g = torch.Generator().manual_seed(2)
    sv, su = torch.randn(bd, generator=g), torch.randn(bd, generator=g)
    n_lat = (R - 1) ** 3
    cells = torch.randperm(n_lat, generator=g)[: n_lat // 2].int()
    z, xy = cells.long() % (R - 1), cells.long() // (R - 1)
    y, x = xy % (R - 1), xy // (R - 1)
    sh = sg.sh.double().clone().requires_grad_(True)
    l0 = links[x, y, z]
    keep = l0 != 0
    v000 = torch.where((l0 >= 0)[:, None], sh[l0.clamp_min(0)], torch.zeros((), dtype=torch.float64))

    def nb(l):
        return torch.where((l >= 0)[:, None], sh[l.clamp_min(0)], v000)
    proj = lambda v, b: (v.view(-1, 3, bd) * b.double()).sum(-1)
    a0 = proj(v000, sv)
    sc = [R / 256.0] * 3
    dx = (proj(nb(links[x + 1, y, z]), sv) - a0) * sc[0]
    dy = (proj(nb(links[x, y + 1, z]), sv) - a0) * sc[1]
    dz = (proj(nb(links[x, y, z + 1]), sv) - a0) * sc[2]
    du = (proj(v000, su) - a0) * 0.8
    e = torch.sqrt(1e-9 + dx * dx + dy * dy + dz * dz + du * du)
    (e[keep].sum() * (0.4 / cells.shape[0])).backward()
    got = np.zeros(tuple(sg.sh.shape), np.float32)
    mask = np.zeros((sg.capacity,), np.uint8)
    oracle.lumisphere_tv_grad_sparse(sg.links, sg.sh, bd, cells, sv, su, 0.4, 0.8, mask, got)
    assert np.abs(got).max() > 0 and _rel(got, sh.grad.numpy()) < 2e-5
    assert mask.sum() > 0 and not mask[np.abs(got).max(1) == 0].all()


def test_msi_tv_matches_autograd_of_its_surrogate():
    """oracle_msi_tv_grad_sparse (loss_kernel.cu:979-1064).  The kernel's update is idelta * axis_scale * d_axis per
    neighbour (the axis scale is applied AFTER the norm), i.e. the gradient of the surrogate
    sum_cells idelta.detach() * sum_axis axis_scale * d_axis.detach() * d_axis -- which autograd differentiates here through an
    independent gather (wrap-around in both texel axes, missing texels read 0, the layer past the last one reads v00, or 0 for
    sigma).  Distinct cells, so that every (texel, layer) is updated by a deterministic set of cells."""
    R, L = 10, 6
    g = torch.Generator().manual_seed(3)
    keep = torch.rand((2 * R, R), generator=g) > 0.2
    links = torch.full((2 * R, R), -1, dtype=torch.int32)
    links[keep] = torch.arange(int(keep.sum()), dtype=torch.int32)
    data = torch.randn((int(keep.sum()), L, 4), generator=g)
    n_cells = links.numel() * L
    cells = torch.randperm(n_cells, generator=g)[: n_cells // 2].int()
    scale, scale_last = 0.7, 0.3
    msi = data.double().clone().requires_grad_(True)
    z = cells.long() % L
    t = cells.long() // L
    y, x = t % R, t // R
    nx, ny = (x + 1) % (2 * R), (y + 1) % R
    l00, l01, l10 = links[x, y].long(), links[x, ny].long(), links[nx, y].long()

    def at(l, zz):
        return torch.where((l >= 0)[:, None], msi[l.clamp_min(0), zz], torch.zeros((), dtype=torch.float64))
    v00, v01, v10 = at(l00, z), at(l01, z), at(l10, z)
    has_next = (l00 >= 0) & (z + 1 < L)
    last = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64)
    v_nx = torch.where(has_next[:, None], msi[l00.clamp_min(0), (z + 1).clamp_max(L - 1)], v00 * (1.0 - last))
    dx, dy, dz = v10 - v00, v01 - v00, v_nx - v00
    sc = torch.tensor([scale, scale, scale, scale_last], dtype=torch.float64) / (n_cells // 2)
    idelta = (sc * torch.rsqrt(1e-9 + dx * dx + dy * dy + dz * dz)).detach()
    ax = (2 * R / 256.0, R / 256.0, L / 256.0)
    (idelta * (ax[0] * dx.detach() * dx + ax[1] * dy.detach() * dy + ax[2] * dz.detach() * dz)).sum().backward()
    got = np.zeros(tuple(data.shape), np.float32)
    mask = np.zeros(tuple(data.shape[:2]), np.uint8)
    oracle.msi_tv_grad_sparse(links, data, cells, mask, scale, scale_last, got)
    assert np.abs(got).max() > 0 and _rel(got, msi.grad.numpy()) < 2e-5
    assert mask.sum() > 0 and (np.abs(got).max(-1)[mask == 0] == 0).all()
