"""CPU: the on-disk grid format (/root/reference/svox2/svox2.py:4693-4838).  Round trip through our writer /
reader, the legacy layout, and -- where the reference package is importable (dev container, or the copy staged under
oracle/_ref/pyref) -- both directions against the reference's own SparseGrid.save / SparseGrid.load."""
import numpy as np
import pytest
import torch

from alphasurf_b200 import checkpoint as ckpt
from alphasurf_b200 import synth
from oracle import ref_l0


def _ck(reso=12, bg=False):
    sg = synth.make_shell_grid(reso, basis_dim=4, variant="G*")
    sg.level_set = torch.tensor([0.0, 0.25])
    ck = ckpt.Checkpoint(grid=sg, radius=torch.tensor([1.0, 1.5, 2.0]), center=torch.tensor([0.1, -0.2, 0.3]),
                         surface_type=ckpt.SURFACE_TYPE_SDF, step_id=7)
    if bg:
        g = torch.Generator().manual_seed(0)
        ck.background_links = torch.arange(2 * 8 * 8, dtype=torch.int32).reshape(16, 8)
        ck.background_data = torch.randn((128, 4, 4), generator=g)
    return ck


@pytest.mark.parametrize("compress,bg", [(False, False), (True, True)])
def test_round_trip(tmp_path, compress, bg):
    ck = _ck(bg=bg)
    p = str(tmp_path / "g.npz")
    ckpt.save(p, ck, compress=compress)
    z = np.load(p)
    assert z["sh_data"].dtype == np.float16 and z["links"].dtype == np.int32      # the reference's storage types
    back = ckpt.load(p)
    g, h = ck.grid, back.grid
    assert torch.equal(g.links, h.links) and torch.equal(g.density, h.density) and torch.equal(g.surface, h.surface)
    assert torch.equal(g.sh.half().float(), h.sh) and torch.equal(g.level_set, h.level_set)
    assert back.step_id == 7 and back.surface_type == ckpt.SURFACE_TYPE_SDF and h.basis_dim == 4
    assert torch.allclose(h.offset, 0.5 * (1 - ck.center / ck.radius)) and torch.allclose(h.scaling, 0.5 / ck.radius)
    if bg:
        assert torch.equal(back.background_links, ck.background_links) and torch.equal(back.background_data, ck.background_data)
    else:
        assert back.background_data is None


def test_legacy_layout_and_defaults(tmp_path):
    sg = synth.make_shell_grid(10, basis_dim=4, variant="G")
    p = str(tmp_path / "old.npz")
    np.savez(p, data=np.concatenate([sg.density.numpy(), sg.sh.numpy()], 1), links=sg.links.numpy())
    back = ckpt.load(p)
    assert back.surface_type == ckpt.SURFACE_TYPE_NONE and back.grid.surface is None
    assert torch.equal(back.grid.density, sg.density) and torch.equal(back.grid.sh, sg.sh)
    assert torch.equal(back.radius, torch.ones(3)) and torch.equal(back.grid.scaling, torch.full((3,), 0.5))


def test_learned_basis_is_refused(tmp_path):
    sg = synth.make_shell_grid(8, basis_dim=4)
    p = str(tmp_path / "b.npz")
    np.savez(p, sh_data=sg.sh.numpy(), density_data=sg.density.numpy(), links=sg.links.numpy(), basis_type=4,
             surface_type=ckpt.SURFACE_TYPE_NONE)
    with pytest.raises(NotImplementedError):
        ckpt.load(p)


@pytest.mark.skipif(not ref_l0.available(), reason="reference svox2 package not present")
def test_against_the_reference_reader_and_writer(tmp_path):
    svox2 = ref_l0.import_reference()
    ck = _ck(reso=12)
    p = str(tmp_path / "ours.npz")
    ckpt.save(p, ck)
    # the reference opens our file
    ref = svox2.SparseGrid.load(p, device="cpu")
    g = ck.grid
    assert torch.equal(ref.links, g.links) and torch.equal(ref.density_data.data, g.density)
    assert torch.equal(ref.surface_data.data, g.surface) and torch.equal(ref.sh_data.data, g.sh.half().float())
    assert torch.equal(ref.level_set_data, g.level_set) and ref.step_id == 7
    assert ref.surface_type == svox2.defs.SURFACE_TYPE_SDF == ckpt.SURFACE_TYPE_SDF and ckpt.BASIS_TYPE_SH == svox2.defs.BASIS_TYPE_SH
    assert torch.allclose(ref._offset, 0.5 * (1 - ck.center / ck.radius)) and torch.allclose(ref._scaling, 0.5 / ck.radius)
    # and we open the reference's
    q = str(tmp_path / "ref.npz")
    ref.save(q, step_id=11)
    back = ckpt.load(q)
    h = back.grid
    assert torch.equal(h.links, g.links) and torch.equal(h.density, g.density) and torch.equal(h.surface, g.surface)
    assert torch.equal(h.sh, g.sh.half().float()) and back.step_id == 11 and back.surface_type == ref.surface_type
    assert torch.allclose(h.offset, ref._offset.float()) and torch.allclose(h.scaling, ref._scaling.float())
