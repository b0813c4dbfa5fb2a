"""Reader / writer of the reference's on-disk grid format.

The reference keeps checkpoint I/O in Python: ``SparseGrid.save`` / ``SparseGrid.load`` (/root/reference/svox2/svox2.py:4693-4727,
:4750-4838) are ``np.savez`` / ``np.load`` of the tensors the render path receives, and ``opt/opt.py:317-346`` starts an
alpha-Surf run from a pretrained Plenoxels file of this format.  This module reads and writes the same archive for the grid
container this package's step drivers use (``synth.SynthGrid`` + an optional MSI background), so that a trainer built on
``alphasurf_b200.step`` can start from, and hand back, files the reference's own tools open.  No arithmetic beyond the
reference's: ``_offset = 0.5 (1 - center / radius)``, ``_scaling = 0.5 / radius`` (svox2.py:644-645), SH stored as float16.
Field by field (reference line):

  radius, center (3,) float32            :4700-4701        links (X,Y,Z) int32                 :4702
  density_data (N,1) float32              :4703             sh_data (N,D) float16                :4704 (read back as float32 :4797-4798)
  step_id                                 :4705             surface_data (N,1) if surface_type   :4707-4708
  level_set_data, fake_sample_std         :4709-4712        background_links / background_data   :4722-4724
  basis_type, surface_type                :4725-4726        legacy "data" = [density | sh]       :4757-4761
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .synth import SynthGrid

BASIS_TYPE_SH = 1            # data_spec.hpp: BASIS_TYPE_SH (svox2/defs.py)
SURFACE_TYPE_NONE = 100      # svox2/defs.py:7-9
SURFACE_TYPE_SDF = 102


@dataclass
class Checkpoint:
    grid: SynthGrid
    radius: torch.Tensor                 # (3,) float32
    center: torch.Tensor                 # (3,) float32
    surface_type: int = SURFACE_TYPE_NONE
    basis_type: int = BASIS_TYPE_SH
    step_id: int = 0
    background_links: Optional[torch.Tensor] = None     # (2R, R) int32
    background_data: Optional[torch.Tensor] = None      # (n, layers, 4) float32


def _np(t):
    return t.detach().cpu().numpy()


def save(path, ck: Checkpoint, compress: bool = False):
    """Write ``ck`` the way SparseGrid.save does (svox2.py:4693-4727): same keys, dtypes and the float16 SH."""
    g = ck.grid
    data = {"radius": _np(ck.radius).astype(np.float32), "center": _np(ck.center).astype(np.float32),
            "links": _np(g.links).astype(np.int32), "density_data": _np(g.density).astype(np.float32),
            "sh_data": _np(g.sh).astype(np.float16), "step_id": ck.step_id}
    if ck.surface_type != SURFACE_TYPE_NONE:
        if g.surface is None:
            raise ValueError("surface_type %d needs surface data" % ck.surface_type)
        data["surface_data"] = _np(g.surface).astype(np.float32)
    if g.level_set is not None:
        data["level_set_data"] = _np(g.level_set).astype(np.float32)
    if ck.background_data is not None:
        data["background_links"] = _np(ck.background_links).astype(np.int32)
        data["background_data"] = _np(ck.background_data).astype(np.float32)
    data["basis_type"] = ck.basis_type
    data["surface_type"] = ck.surface_type
    (np.savez_compressed if compress else np.savez)(path, **data)


def load(path, device="cpu") -> Checkpoint:
    """Read a file written by SparseGrid.save (or by ``save`` above), following SparseGrid.load (svox2.py:4750-4838):
    the legacy single-array layout, missing radius / center defaults, float16 -> float32."""
    z = np.load(path, allow_pickle=True)
    surface = None
    surface_type = SURFACE_TYPE_NONE
    if "data" in z.files:                       # compatibility layout (:4757-4761)
        all_data = z["data"]
        sh, density = all_data[..., 1:], all_data[..., :1]
    else:
        sh, density = z["sh_data"], z["density_data"]
        surface_type = int(z["surface_type"].item()) if "surface_type" in z.files else SURFACE_TYPE_NONE
        if surface_type != SURFACE_TYPE_NONE:
            surface = z["surface_data"].astype(np.float32)
    basis_type = int(z["basis_type"].item()) if "basis_type" in z.files else BASIS_TYPE_SH
    if basis_type != BASIS_TYPE_SH or "basis_data" in z.files:
        raise NotImplementedError("only the SH basis is on the B200 hot path (learned bases: SURVEY.md 8, out of scope)")
    links = z["links"]
    radius = np.asarray(z["radius"] if "radius" in z.files else [1.0, 1.0, 1.0], np.float32).reshape(-1)
    center = np.asarray(z["center"] if "center" in z.files else [0.0, 0.0, 0.0], np.float32).reshape(-1)
    if radius.size == 1:
        radius = np.repeat(radius, 3)
    radius_t, center_t = torch.from_numpy(radius.copy()), torch.from_numpy(center.copy())
    level_set = torch.from_numpy(z["level_set_data"].astype(np.float32)) if "level_set_data" in z.files else None
    fake_std = float(np.asarray(z["fake_sample_std"]).reshape(-1)[0]) if "fake_sample_std" in z.files else 1.0
    dev = torch.device(device)
    mv = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    grid = SynthGrid(links=mv(links.astype(np.int32)), density=mv(density.astype(np.float32)), surface=mv(surface),
                     sh=mv(sh.astype(np.float32)), level_set=None if level_set is None else level_set.to(dev),
                     offset=0.5 * (1.0 - center_t / radius_t), scaling=0.5 / radius_t, basis_dim=int(sh.shape[1]) // 3,
                     fake_sample_std=fake_std, meta={"source": str(path)})
    ck = Checkpoint(grid=grid, radius=radius_t, center=center_t, surface_type=surface_type, basis_type=basis_type,
                    step_id=int(z["step_id"].item()) if "step_id" in z.files else 0)
    if "background_data" in z.files:
        ck.background_links = mv(z["background_links"].astype(np.int32))
        ck.background_data = mv(z["background_data"].astype(np.float32))
    return ck
