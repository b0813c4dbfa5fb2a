"""The whole C3 iteration (alphasurf_b200.step.TrainStep: fused render -> TV / normal / sparsity regularisers -> RMSprop) on
our module vs the same call sequence on the UNMODIFIED reference kernels (oracle/_ref), same seeded grid, rays and cells."""
import pytest
import torch

from alphasurf_b200 import step as S
from alphasurf_b200 import svox2_csrc as ours
from alphasurf_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _clone_grid(sg):
    return synth.SynthGrid(sg.links, sg.density.clone(), sg.surface.clone(), sg.sh.clone(), sg.level_set, sg.offset,
                           sg.scaling, sg.basis_dim, sg.fake_sample_std, sg.truncated_vol_render_a, dict(sg.meta))


@pytest.mark.parametrize("variant,reso", [("G*", 40), ("G", 64)])
def test_train_step_matches_reference_kernels(variant, reso):
    ref = H.load_reference_cuda()
    if ref is None:
        pytest.skip("reference CUDA build (oracle/_ref) not present")
    sg = synth.make_shell_grid(reso, basis_dim=9, variant=variant).to("cuda")
    a, b = S.TrainStep(ours, _clone_grid(sg), seed=5), S.TrainStep(ref, _clone_grid(sg), seed=5)
    Q = 4096
    out_a, out_b = torch.zeros((Q, 3), device="cuda"), torch.zeros((Q, 3), device="cuda")
    for it in range(3):
        o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=77 + it)
        for ts, out in ((a, out_a), (b, out_b)):
            ts.render(o, d, gt, out)
            ts.regularisers()
        torch.cuda.synchronize()
        assert H.rel_err(out_a, out_b) < 1e-4
        assert torch.equal(a.mask, b.mask) and torch.equal(a.mask_sh, b.mask_sh)
        for k in ("density", "surface", "sh"):
            assert H.rel_err(a.grad[k], b.grad[k]) < 2e-4, (it, k)
        # same gradients into both optimizers: the update itself is bit-exact (tests/test_optim_gpu.py), and feeding
        # both sides one gradient keeps RMSprop's g/|g| from amplifying atomic-order noise on near-zero entries
        for k in ("density", "surface", "sh"):
            b.grad[k].copy_(a.grad[k])
        a.optimizer()
        b.optimizer()
        torch.cuda.synchronize()
        for k in ("density", "surface", "sh"):
            assert torch.equal(getattr(a.sg, k), getattr(b.sg, k)), (it, k)
            assert float(a.grad[k].abs().max()) == 0.0 or k == "sh"    # grads of the stepped rows are zeroed
    assert float((a.sg.density - sg.density).abs().max()) > 0      # the step did change the grid


def test_train_step_full_size_smoke():
    """256^3 / 16k rays: the step runs, touches the expected share of rows, and leaves finite parameters."""
    sg = synth.make_shell_grid(256, basis_dim=9, variant="G").to("cuda")
    ts = S.TrainStep(ours, sg)
    Q = 16384
    out = torch.zeros((Q, 3), device="cuda")
    o, d, gt = synth.make_camera_rays(Q, device="cuda")
    ts.step(o, d, gt, out)
    torch.cuda.synchronize()
    assert 0 < int(ts.mask_sh.sum()) < sg.capacity // 4          # the render touches a small share of the rows
    assert int(ts.mask.sum()) > sg.capacity // 2                 # surface TV / normal loss run over every stored cell
    for k in ("density", "surface", "sh"):
        assert bool(torch.isfinite(getattr(sg, k)).all())


def test_incremental_work_pyramid_equals_full_rebuild():
    """Across training steps the library updates its cached work pyramid only around the vertices whose level-set side or
    density gate changed; after every render call it must equal a pyramid built from scratch on the same data, bit for bit.
    Large learning rates make many vertices change side per step."""
    import ctypes as C
    from alphasurf_b200 import capi
    L = capi.lib()
    for variant, reso in (("G", 96), ("G*", 64)):
        sg = synth.make_shell_grid(reso, basis_dim=9, variant=variant).to("cuda")
        hp = S.c3_hyper()
        hp.update(lr_surface=3e-2, lr_density=0.2)
        ts = S.TrainStep(ours, sg, hyper=hp)
        Q = 8192
        out = torch.zeros((Q, 3), device="cuda")
        n3 = L.asurf_accel_words(capi.size3(sg.links.shape))
        cached = torch.zeros((n3,), dtype=torch.int64, device="cuda")
        full = torch.zeros((n3,), dtype=torch.int64, device="cuda")
        n_changed_total = 0
        prev = None
        for it in range(5):
            o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=300 + it)
            ts.render(o, d, gt, out)
            assert L.asurf_debug_work_cache_valid() == 1
            # size of the three pyramid levels
            from alphasurf_b200.svox2_csrc import accel_for, _grid_t
            g, keep = _grid_t(ts.grid_spec)
            b0 = (max(reso - 1, 1) + 3) // 4                              # AccelLayout (common.cuh): 4^3, 16^3, 64^3 cells
            b1 = (b0 + 3) // 4
            lay_words = b0 ** 3 + b1 ** 3 + ((b1 + 3) // 4) ** 3
            capi.check(L.asurf_debug_work_cache_copy(capi.ptr(cached), C.c_int64(lay_words), capi.current_stream()), "copy")
            capi.check(L.asurf_work_build(C.byref(g), C.byref(capi.make_opt(ts.opt_spec)), capi.ptr(full),
                                          capi.current_stream()), "work_build")
            torch.cuda.synchronize()
            assert torch.equal(cached[:lay_words], full[:lay_words]), (variant, it)
            if prev is not None:
                n_changed_total += int((prev != full[:lay_words]).sum())
            prev = full[:lay_words].clone()
            ts.regularisers()
            ts.optimizer()
        assert n_changed_total > 0, "the steps were supposed to move the level set through some voxels"
