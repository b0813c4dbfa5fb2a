"""One alpha-Surf training iteration (config C3 of SURVEY.md 8d) as ``opt/opt.py`` drives it through ``svox2.csrc``.

Host-side mirror of /root/reference/opt/opt.py:800-1125 + the ``SparseGrid`` helpers it calls
(/root/reference/svox2/svox2.py:3480-3640 ``volume_render_fused``, :4950-5163 and :5690-5724 ``inplace_*_grad``,
:5972-6100 ``optim_*_step``, :6314-6376 indexers / cell lists).  Every line below is a call into the
``svox2.csrc``-compatible module; no arithmetic happens in Python.  bench.py times this step; the GPU tests compare it
with the same sequence run on the UNMODIFIED reference kernels.
"""
import numpy as np
import torch

from . import synth

RMS_BETA, RMS_EPS = 0.95, 1e-8   # opt/util/config_util.py:473 (rms_beta), svox2.py:5972 (epsilon)
SURFACE_TYPE_SDF = 102
BASIS_TYPE_SH = 1
FUSED_ORDER = ["beta_loss", "sparsity_loss", "fused_surf_norm_reg_scale", "fused_surf_norm_reg_con_check",
               "fused_surf_norm_reg_ignore_empty", "lambda_l2", "lambda_l1", "lambda_l_dist", "lambda_l_entropy",
               "no_norm_weight_l_entropy", "lambda_l_dist_a", "lambda_l_entropy_a", "lambda_l_samp_dist", "lambda_l_di",
               "l_di_alpha_thresh", "surf_sparse_alpha_thresh", "lambda_inplace_surf_sparse", "lambda_inwards_norm_loss",
               "lambda_conv_mode_samp", "l_dist_max_sample"]


# ---- SparseGrid._to_cpp / Rays._to_cpp / RenderOptions._to_cpp (svox2.py:6234-6272, :118-128, :57-90) ---------------
def grid_to_cpp(C, sg):
    g = C.SparseGridSpec()
    g.density_data = sg.density
    g.surface_type = SURFACE_TYPE_SDF
    g.surface_data = sg.surface
    g.level_set_data = sg.level_set
    g.sh_data = sg.sh
    g.links = sg.links
    g._offset = sg.offset
    g._scaling = sg.scaling
    g.basis_dim = sg.basis_dim
    g.basis_type = BASIS_TYPE_SH
    g.fake_sample_std = float(sg.fake_sample_std)
    g.truncated_vol_render_a = float(sg.truncated_vol_render_a)
    return g


_RAY_MASKS = {}


def rays_to_cpp(C, origins, dirs):
    r = C.RaysSpec()
    r.origins = origins
    r.dirs = dirs
    key = (origins.shape[0], origins.device)
    m = _RAY_MASKS.get(key)             # Rays.masks defaults to all-true (svox2.py:104-108); one tensor per batch size
    if m is None:
        if len(_RAY_MASKS) > 16:
            _RAY_MASKS.clear()
        m = _RAY_MASKS[key] = torch.ones((origins.shape[0],), dtype=torch.bool, device=origins.device)
    r.masks = m
    return r


def opt_to_cpp(C, d):
    o = C.RenderOptions()
    for k, v in d.items():
        if k != "backend":
            setattr(o, k, v)
    return o


def fused_positional(fd):
    return [fd[k] for k in FUSED_ORDER]


def c3_hyper():
    """Regulariser / optimizer settings of opt/configs/surface_cuda_syn.yaml (+ config_util defaults)."""
    return dict(lr_density=1e-2, lr_surface=1e-5, lr_sh=1e-3,
                lambda_tv_alpha=1e-5, tv_sparsity=0.01,
                lambda_tv_surface=1e-3, tv_surface_sparsity=1.0, surf_tv_ignore_edge=True, surf_tv_edge_value=-1.0,
                surf_tv_alpha_dependency=False,
                lambda_normal_loss=1e-6, norm_surface_sparsity=1.0, norm_con_check=False, norm_ignore_empty=False,
                lambda_sparsify_alpha=1e-9, lambda_sparsify_surf=0.0, alpha_surf_sparsify_sparsity=0.1,
                sparsify_surf_decrease=True, sparsify_surf_thresh=0.15, alpha_sparsify_bound=0.0,
                surf_sparsify_bound=-0.1)


class TrainStep:
    """Grid state + the per-iteration call sequence.  ``C`` is the csrc-compatible module (ours or the reference's)."""

    def __init__(self, C, sg, render_opts=None, fused=None, hyper=None, seed=synth.SEED):
        self.C, self.sg = C, sg
        dev = sg.density.device
        self.dev = dev
        self.opts = render_opts or synth.alphasurf_render_options()
        self.fused = fused or synth.alphasurf_fused_args()
        self.hp = hyper or c3_hyper()
        self.grad = {k: torch.zeros_like(getattr(sg, k)) for k in ("density", "surface", "sh")}
        self.rms = {k: torch.zeros_like(getattr(sg, k)) for k in ("density", "surface", "sh")}
        self.mask = torch.zeros((sg.capacity,), dtype=torch.bool, device=dev)       # sparse_grad_indexer
        self.mask_sh = torch.zeros((sg.capacity,), dtype=torch.bool, device=dev)    # sparse_sh_grad_indexer
        self.rng = np.random.RandomState(seed & 0x7fffffff)
        self.grid_size = sg.links.numel()
        # _get_rand_cells_non_empty (svox2.py:6354-6376): links never change during training -> computed once
        self.non_empty = torch.where(sg.links.view(-1) >= 0)[0].int().contiguous()
        self.grid_spec, self.opt_spec = grid_to_cpp(C, sg), opt_to_cpp(C, self.opts)
        gs = C.GridOutputGrads()
        gs.grad_density_out, gs.grad_surface_out, gs.grad_sh_out = self.grad["density"], self.grad["surface"], self.grad["sh"]
        gs.mask_out = self.mask
        self.grad_spec = gs
        self.fpos = fused_positional(self.fused)
        self.no_mask = torch.empty((0,), dtype=torch.bool, device=dev)

    # ---- cell lists ----------------------------------------------------------------------------------------------
    def rand_cells(self, frac):           # svox2.py:6335-6352, contiguous=True
        if frac >= 1.0:
            return None
        n = max(int(frac * self.grid_size), 1)
        start = int(self.rng.randint(0, self.grid_size))
        arr = torch.arange(start, start + n, dtype=torch.int32, device=self.dev)
        if start > self.grid_size - n:
            arr[self.grid_size - n - start:] -= self.grid_size
        return arr

    def rand_cells_non_empty(self, frac):  # svox2.py:6354-6373, contiguous=True
        if frac >= 1.0:
            return self.non_empty
        ne = self.non_empty.shape[0]
        n = int(ne * frac)
        start = int(self.rng.randint(0, ne - n + 1))
        return self.non_empty[start:start + n]

    # ---- the iteration -----------------------------------------------------------------------------------------------
    def render(self, origins, dirs, rgb_gt, rgb_out):
        """volume_render_fused (svox2.py:3480-3640): fresh bool indexer, fused kernel, sh indexer = its copy."""
        self.mask.zero_()
        self.C.volume_render_surf_trav_fused(self.grid_spec, rays_to_cpp(self.C, origins, dirs), self.opt_spec, rgb_gt,
                                             *self.fpos, rgb_out, self.grad_spec)
        self.mask_sh.copy_(self.mask)

    def step(self, origins, dirs, rgb_gt, rgb_out, exchange=None):
        """``exchange``: alphasurf_b200.dist.GradExchange for the ray-sharded multi-GPU step: the sparse all-reduce of the
        render gradients overlaps the regularisers, which are themselves sharded over the ranks by cell."""
        if exchange is not None:
            exchange.step(self, origins, dirs, rgb_gt, rgb_out)
            return
        self.render(origins, dirs, rgb_gt, rgb_out)
        self.regularisers()
        self.optimizer()

    @staticmethod
    def _shard(cells, rank, world):
        """this rank's contiguous share of a cell list (multi-GPU: the regularisers are data-parallel over cells)"""
        if world == 1:
            return cells
        n = cells.shape[0]
        return cells[(n * rank) // world:(n * (rank + 1)) // world]

    def density_terms_replicable(self):
        """True when the regularisers that write the density gradient (density TV over a 1 % window, opacity sparsity over a
        10 % window: ~35 us together) can run in full on every rank of a multi-GPU step instead of being sharded -- the
        density gradient of the regularisers then needs no exchange.  Not so when the sparsity term also pushes the
        surface (lambda_sparsify_surf != 0): its surface part would be summed once per rank by the surface exchange."""
        return self.hp["lambda_sparsify_surf"] == 0

    def regularisers(self, rank=0, world=1, grad=None, mask=None, replicate_density_terms=False):
        """With world > 1 every rank takes 1/world of each cell list.  The kernels normalise by the length of the list they
        are given (n_r on rank r), so the scale handed to them is lambda * n_r / n: every cell then weighs lambda / n, the
        single-process normalisation, whether or not the list divides evenly.  GradExchange sums the shards (dense
        all-reduce of the density / surface gradients, OR of the masks).  ``grad`` / ``mask``: accumulate into these
        buffers instead of the step's own (the overlapped multi-GPU step keeps the regulariser gradients apart until both
        exchanges are done)."""
        C, sg, hp = self.C, self.sg, self.hp
        g = self.grad if grad is None else grad
        mask_t = self.mask if mask is None else mask

        def sh_(cells, lam, replicate=False):
            """-> (this rank's share, its scale)"""
            if replicate or world == 1:
                return cells, lam
            part = self._shard(cells, rank, world)
            return part, lam * part.shape[0] / cells.shape[0]

        rep = bool(replicate_density_terms) and self.density_terms_replicable()
        if hp["lambda_tv_alpha"] > 0:      # inplace_tv_grad, opt.py:952-957
            cells, lam = sh_(self.rand_cells(hp["tv_sparsity"]), hp["lambda_tv_alpha"], rep)
            C.tv_grad_sparse(sg.links, sg.density, cells, mask_t, 0, 1, lam, False, 2.0, False,
                             bool(self.opts["last_sample_opaque"]), -1.0, -1.0, g["density"])
        if hp["lambda_tv_surface"] > 0:    # inplace_tv_surface_grad, opt.py:959-968
            cells, lam = sh_(self.rand_cells_non_empty(hp["tv_surface_sparsity"]), hp["lambda_tv_surface"])
            C.surf_tv_grad_sparse(sg.links, sg.surface, sg.density, cells, mask_t, 0, 1, lam,
                                  hp["surf_tv_ignore_edge"], hp["surf_tv_edge_value"], bool(self.opts["last_sample_opaque"]),
                                  -1.0, -1.0, hp["surf_tv_alpha_dependency"], g["surface"])
        if hp["lambda_normal_loss"] > 0:   # inplace_surface_normal_grad, opt.py:970-981
            cells, lam = sh_(self.rand_cells_non_empty(hp["norm_surface_sparsity"]), hp["lambda_normal_loss"])
            C.surface_normal_grad_sparse(sg.links, sg.surface, cells, mask_t, 0.0, 0, 1, lam, 0.0,
                                         -1.0, -1.0, hp["norm_con_check"], hp["norm_ignore_empty"], True, g["surface"])
        if hp["lambda_sparsify_alpha"] > 0 or hp["lambda_sparsify_surf"] > 0:   # opt.py:1046-1060
            # (the sparsity loss is NOT normalised by the list length, loss_kernel.cu:1555-1558: its scale stays)
            cells = self.rand_cells_non_empty(hp["alpha_surf_sparsify_sparsity"])
            if not rep:
                cells = self._shard(cells, rank, world)
            C.alpha_surf_sparsify_grad_sparse(sg.links, sg.density, sg.surface, cells, mask_t, hp["lambda_sparsify_alpha"],
                                              hp["lambda_sparsify_surf"], hp["sparsify_surf_decrease"],
                                              hp["sparsify_surf_thresh"], hp["alpha_sparsify_bound"],
                                              hp["surf_sparsify_bound"], g["density"], g["surface"])

    def optimizer(self):
        """optim_density_step / optim_surface_step / optim_sh_step with the bool indexers (svox2.py:5972-6100).  The
        reference converts a sparse mask to an index list on the host (a count_nonzero().item() sync per tensor,
        :6314-6333); the masked kernels here make that conversion unnecessary, the update is the same."""
        C, sg, hp = self.C, self.sg, self.hp
        C.rmsprop_step(sg.density, self.rms["density"], self.grad["density"], self.mask, RMS_BETA, hp["lr_density"], RMS_EPS,
                       -1e9, hp["lr_density"])
        C.rmsprop_step(sg.surface, self.rms["surface"], self.grad["surface"], self.mask, RMS_BETA, hp["lr_surface"], RMS_EPS,
                       -1e9, hp["lr_surface"])
        C.rmsprop_step(sg.sh, self.rms["sh"], self.grad["sh"], self.mask_sh, RMS_BETA, hp["lr_sh"], RMS_EPS, -1e9,
                       hp["lr_sh"])


SURFACE_TYPE_NONE = 100


def plenoxels_render_options():
    """RenderOptions of the cuvol backend: opt/util/config_util.py:81-92 defaults as configs/syn.yaml leaves them."""
    o = synth.alphasurf_render_options()
    o.update(backend="cuvol", sigma_thresh=1e-8, stop_thresh=1e-7, step_size=0.5)
    return o


def c2_hyper():
    """Plenoxels regulariser / optimizer settings (config_util.py:564-640 defaults + configs/syn.yaml)."""
    return dict(lr_sigma=3e1, lr_sh=1e-2, lambda_tv=1e-5, tv_sparsity=0.01, lambda_tv_sh=1e-3, tv_sh_sparsity=0.01)


class CuvolStep:
    """One Plenoxels iteration (config C2) as opt/opt.py drives it for a grid without a surface: volume_render_fused with the
    cuvol backend (svox2.py:3477-3638), inplace_tv_grad on sigma and inplace_tv_color_grad on the SH coefficients over a
    contiguous 1 % window of cells (:4947-4985, :5768-5812), RMSprop on sigma and SH (:5972-6009, :6110-6150)."""

    def __init__(self, C, sg, render_opts=None, hyper=None, seed=synth.SEED):
        self.C, self.sg = C, sg
        dev = sg.density.device
        self.dev = dev
        self.opts = render_opts or plenoxels_render_options()
        self.hp = hyper or c2_hyper()
        self.grad = {k: torch.zeros_like(getattr(sg, k)) for k in ("density", "sh")}
        self.rms = {k: torch.zeros_like(getattr(sg, k)) for k in ("density", "sh")}
        self.mask = torch.zeros((sg.capacity,), dtype=torch.bool, device=dev)
        self.mask_sh = torch.zeros((sg.capacity,), dtype=torch.bool, device=dev)
        self.rng = np.random.RandomState(seed & 0x7fffffff)
        self.grid_size = sg.links.numel()
        g = C.SparseGridSpec()
        g.density_data, g.sh_data, g.links = sg.density, sg.sh, sg.links
        g._offset, g._scaling = sg.offset, sg.scaling
        g.basis_dim, g.basis_type, g.surface_type = sg.basis_dim, BASIS_TYPE_SH, SURFACE_TYPE_NONE
        self.grid_spec, self.opt_spec = g, opt_to_cpp(C, self.opts)
        gs = C.GridOutputGrads()
        gs.grad_density_out, gs.grad_sh_out, gs.mask_out = self.grad["density"], self.grad["sh"], self.mask
        self.grad_spec = gs

    rand_cells = TrainStep.rand_cells

    def render(self, origins, dirs, rgb_gt, rgb_out):
        self.mask.zero_()
        self.C.volume_render_cuvol_fused(self.grid_spec, rays_to_cpp(self.C, origins, dirs), self.opt_spec, rgb_gt, 0.0, 0.0,
                                         rgb_out, self.grad_spec)
        self.mask_sh.copy_(self.mask)

    def regularisers(self):
        C, sg, hp = self.C, self.sg, self.hp
        C.tv_grad_sparse(sg.links, sg.density, self.rand_cells(hp["tv_sparsity"]), self.mask, 0, 1, hp["lambda_tv"], False, 2.0,
                         False, bool(self.opts["last_sample_opaque"]), -1.0, -1.0, self.grad["density"])
        C.tv_grad_sparse(sg.links, sg.sh, self.rand_cells(hp["tv_sh_sparsity"]), self.mask_sh, 0, sg.sh.shape[1],
                         hp["lambda_tv_sh"], False, 2.0, True, False, -1.0, -1.0, self.grad["sh"])

    def optimizer(self):
        C, sg, hp = self.C, self.sg, self.hp
        C.rmsprop_step(sg.density, self.rms["density"], self.grad["density"], self.mask, RMS_BETA, hp["lr_sigma"], RMS_EPS, -1e9,
                       hp["lr_sigma"])
        C.rmsprop_step(sg.sh, self.rms["sh"], self.grad["sh"], self.mask_sh, RMS_BETA, hp["lr_sh"], RMS_EPS, -1e9, hp["lr_sh"])

    def step(self, origins, dirs, rgb_gt, rgb_out):
        self.render(origins, dirs, rgb_gt, rgb_out)
        self.regularisers()
        self.optimizer()
