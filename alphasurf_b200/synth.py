"""Seeded synthetic SparseGrid contents and ray batches.

There is no network for datasets or checkpoints, so every test and benchmark runs on the grids and
rays generated here.  The layouts are exactly what the reference's ``SparseGrid`` hands to ``svox2.csrc``
(``/root/reference/svox2/svox2.py:580-990`` for the tensors, ``:6234-6272`` for ``_to_cpp``):

* ``links``   int32 (X, Y, Z): row index into the data tensors, or a negative value for an empty vertex;
* ``density`` float32 (N, 1): raw opacity (surf_trav) or sigma (cuvol);
* ``surface`` float32 (N, 1): level-set scalar; ``level_set`` float32 (L,);
* ``sh``      float32 (N, 3*basis_dim), channel-major (``svox2.py:1333``);
* ``_offset`` / ``_scaling`` float32 CPU (3,): world -> grid transform, already multiplied by the grid size.
"""
from dataclasses import dataclass, field
from typing import Optional

import math
import torch

SEED = 20200823  # same constant as /root/reference/opt/opt.py:88


def _expand_bits(v: torch.Tensor) -> torch.Tensor:
    # 10-bit -> 30-bit interleave (z-order curve), cf. /root/reference/svox2/utils.py:49-66
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton_code_3(x, y, z):
    return (_expand_bits(x) << 2) + (_expand_bits(y) << 1) + _expand_bits(z)


def _is_pow2(x: int) -> bool:
    return x > 0 and (x & (x - 1)) == 0


@dataclass
class SynthGrid:
    links: torch.Tensor
    density: torch.Tensor
    surface: Optional[torch.Tensor]
    sh: torch.Tensor
    level_set: Optional[torch.Tensor]
    offset: torch.Tensor          # CPU (3,)
    scaling: torch.Tensor         # CPU (3,)
    basis_dim: int
    fake_sample_std: float = 1.0
    truncated_vol_render_a: float = 5.0
    meta: dict = field(default_factory=dict)

    @property
    def capacity(self) -> int:
        return self.density.shape[0]

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return SynthGrid(mv(self.links), mv(self.density), mv(self.surface), mv(self.sh), mv(self.level_set),
                         self.offset, self.scaling, self.basis_dim, self.fake_sample_std,
                         self.truncated_vol_render_a, dict(self.meta))


def make_shell_grid(reso: int, basis_dim: int = 9, device="cpu", variant: str = "G", seed: int = SEED,
                    z_order: Optional[bool] = None, sigma_density: bool = False,
                    shell_mid: float = 0.30, shell_half: float = 0.05) -> SynthGrid:
    """G(R) / G*(R) of SURVEY.md 8(d).

    Occupancy: vertex (x,y,z) is stored iff | |p-c| - shell_mid*R | <= shell_half*R, c = (R/2,)*3.
    variant "G":  surface = 0.1*(|p-c| - shell_mid*R) + U(-0.02, 0.02)  (about two crossings per ray)
    variant "G*": concentric level sets + U(-.5, .5) noise, as the reference gradcheck test
                  (/root/reference/test/test_render_gradcheck_surface.py:73-77): a crossing in most voxels.
    density ~ N(0.5, 0.1) (raw alpha), or N(20, 5) clamped >= 0 when ``sigma_density`` (cuvol).
    sh: DC = 0.5, other coefficients ~ N(0, 0.1).
    """
    R = int(reso)
    if z_order is None:
        z_order = _is_pow2(R)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    ar = torch.arange(R, device=device, dtype=torch.int64)
    X, Y, Z = torch.meshgrid(ar, ar, ar, indexing="ij")
    c = R / 2.0
    rad = torch.sqrt((X.float() - c) ** 2 + (Y.float() - c) ** 2 + (Z.float() - c) ** 2)
    occ = (rad - shell_mid * R).abs() <= shell_half * R
    if z_order:
        order = morton_code_3(X, Y, Z)
    else:
        order = (X * R + Y) * R + Z
    del X, Y, Z
    occ_flat = occ.reshape(-1)
    order_occ = order.reshape(-1)[occ_flat]
    rank = torch.argsort(torch.argsort(order_occ))
    links = torch.full((R * R * R,), -1, dtype=torch.int32, device=device)
    links[occ_flat] = rank.to(torch.int32)
    links = links.reshape(R, R, R).contiguous()
    N = int(order_occ.numel())
    rad_occ = rad.reshape(-1)[occ_flat]
    rad_rows = torch.empty(N, device=device, dtype=torch.float32)
    rad_rows[rank] = rad_occ
    del rad, occ, order

    D = 3 * basis_dim
    sh = (torch.randn((N, D), generator=gen, dtype=torch.float32) * 0.1).to(device)
    sh[:, 0::basis_dim] = 0.5
    if sigma_density:
        density = (torch.randn((N, 1), generator=gen, dtype=torch.float32) * 5.0 + 20.0).clamp_min(0.0).to(device)
    else:
        density = (torch.randn((N, 1), generator=gen, dtype=torch.float32) * 0.1 + 0.5).to(device)
    u = torch.rand((N, 1), generator=gen, dtype=torch.float32).to(device)
    if variant == "G":
        surface = 0.1 * (rad_rows[:, None] - shell_mid * R) + (u * 0.04 - 0.02)
    elif variant in ("G*", "Gstar"):
        # signed distance to the nearest of the concentric spheres of radius 0.5, 2.5, 4.5, ...
        # (surface_init='sphere', /root/reference/svox2/svox2.py:779-792) plus U(-.5,.5) noise
        d = rad_rows[:, None] - 0.5
        surface = d - 2.0 * torch.round(d / 2.0) + (u - 0.5)
    else:
        raise ValueError(variant)
    level_set = torch.zeros((1,), dtype=torch.float32, device=device)
    gsz = torch.tensor([R, R, R], dtype=torch.float32)
    offset = 0.5 * gsz     # radius 1, center 0: _offset = 0.5, _scaling = 0.5 (svox2.py:612-616) times gsz
    scaling = 0.5 * gsz
    return SynthGrid(links, density.contiguous(), surface.contiguous(), sh.contiguous(), level_set, offset, scaling,
                     basis_dim, 1.0, 5.0, {"reso": R, "variant": variant, "z_order": bool(z_order), "N": N})


def make_camera_rays(Q: int, device="cpu", seed: int = SEED, n_cams: int = 100, width: int = 800, height: int = 800,
                     fx: float = 1111.11, cam_radius: float = 2.6875):
    """Blender-shaped pinhole cameras on a sphere looking at the origin; Q (camera, pixel) pairs drawn with a
    seeded generator (the statistical equivalent of the first Q entries of a permutation of all pixels).
    Returns origins (Q,3), unit dirs (Q,3), rgb_gt (Q,3) ~ U(0,1); float32."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed + 1)
    v = torch.randn((n_cams, 3), generator=gen, dtype=torch.float64)
    v = v / v.norm(dim=1, keepdim=True)
    centers = v * cam_radius
    fwd = -v                                     # look at the origin (OpenCV: +z forward)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(fwd)
    right = torch.cross(fwd, up, dim=1)
    right = right / right.norm(dim=1, keepdim=True).clamp_min(1e-9)
    down = torch.cross(fwd, right, dim=1)
    rot = torch.stack([right, down, fwd], dim=2)  # columns = camera axes in world space
    cam = torch.randint(0, n_cams, (Q,), generator=gen)
    px = torch.randint(0, width, (Q,), generator=gen).double()
    py = torch.randint(0, height, (Q,), generator=gen).double()
    d_cam = torch.stack([(px + 0.5 - width * 0.5) / fx, (py + 0.5 - height * 0.5) / fx, torch.ones_like(px)], dim=1)
    dirs = torch.einsum("qij,qj->qi", rot[cam], d_cam)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    origins = centers[cam]
    rgb_gt = torch.rand((Q, 3), generator=gen, dtype=torch.float32)
    return (origins.float().contiguous().to(device), dirs.float().contiguous().to(device), rgb_gt.to(device))


def make_image_rays(device="cpu", seed: int = SEED, cam_index: int = 0, width: int = 800, height: int = 800,
                    fx: float = 1111.11, cam_radius: float = 2.6875):
    """One full image in raster order (config C5)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed + 2 + cam_index)
    v = torch.randn((3,), generator=gen, dtype=torch.float64)
    v = v / v.norm()
    fwd = -v
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    rot = torch.stack([right, down, fwd], dim=1)
    yy, xx = torch.meshgrid(torch.arange(height, dtype=torch.float64), torch.arange(width, dtype=torch.float64), indexing="ij")
    d_cam = torch.stack([(xx + 0.5 - width * 0.5) / fx, (yy + 0.5 - height * 0.5) / fx, torch.ones_like(xx)], dim=-1).reshape(-1, 3)
    dirs = d_cam @ rot.T
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    origins = (v * cam_radius).expand_as(dirs)
    return origins.float().contiguous().to(device), dirs.float().contiguous().to(device)


def alphasurf_render_options():
    """RenderOptions of /root/reference/opt/configs/surface_cuda_syn.yaml."""
    return dict(backend="surf_trav", background_brightness=1.0, step_size=0.5, sigma_thresh=-10000.0,
                stop_thresh=-10000.0, near_clip=0.0, use_spheric_clip=False, last_sample_opaque=False,
                surf_fake_sample=False, surf_fake_sample_min_vox_len=0.1, limited_fake_sample=False,
                no_surf_grad_from_sh=False, alpha_activation_type=1, fake_sample_l_dist=True,
                fake_sample_normalize_surf=True, only_outward_intersect=True, truncated_vol_render=True,
                trunc_vol_weight_min=1e-10)


def parity_render_options():
    """Option set of /root/reference/test/test_render_gradcheck_surface.py:44-62 (the only one the
    reference's pure-PyTorch renderer supports)."""
    o = alphasurf_render_options()
    o.update(sigma_thresh=-20.0, stop_thresh=0.0, surf_fake_sample=True, surf_fake_sample_min_vox_len=0.0,
             limited_fake_sample=True, only_outward_intersect=False, trunc_vol_weight_min=0.0)
    return o


def alphasurf_fused_args():
    """Fused-loss scalars of surface_cuda_syn.yaml as passed by /root/reference/opt/opt.py:808-830."""
    return dict(beta_loss=0.0, sparsity_loss=0.0, fused_surf_norm_reg_scale=0.0, fused_surf_norm_reg_con_check=True,
                fused_surf_norm_reg_ignore_empty=False, lambda_l2=1.0, lambda_l1=0.0, lambda_l_dist=0.0,
                lambda_l_entropy=1e-4, no_norm_weight_l_entropy=False, lambda_l_dist_a=0.0, lambda_l_entropy_a=0.0,
                lambda_l_samp_dist=0.0, lambda_l_di=0.0, l_di_alpha_thresh=0.0, surf_sparse_alpha_thresh=0.0,
                lambda_inplace_surf_sparse=0.0, lambda_inwards_norm_loss=0.0, lambda_conv_mode_samp=1e-6,
                l_dist_max_sample=64)
