"""CPU: host-side mirror of the reference interface (alphasurf_b200/svox2_csrc.py): names, spec fields and the error
behaviour of the reference's TORCH_CHECKs (include/util.hpp:3-12, include/data_spec.hpp:58-81)."""
import inspect

import pytest
import torch

from alphasurf_b200 import svox2_csrc as C
from alphasurf_b200 import synth
from tests import helpers as H


def test_surface_of_the_module():
    for name in ("volume_render_surf_trav", "volume_render_surf_trav_backward", "volume_render_surf_trav_fused",
                 "rmsprop_step", "sgd_step", "SparseGridSpec", "RaysSpec", "RenderOptions", "GridOutputGrads",
                 "CameraSpec", "RayVoxIntersecSpec"):
        assert hasattr(C, name), name
    # part of the contract (svox2.py:3660 probes for it): the image variant must NOT exist
    assert not hasattr(C, "volume_render_surf_trav_image")
    # positional signature of the fused call, svox2.cpp:85-111
    assert len(inspect.signature(C.volume_render_surf_trav_fused).parameters) == 26
    assert len(inspect.signature(C.rmsprop_step).parameters) == 9
    assert len(inspect.signature(C.sgd_step).parameters) == 5


def test_spec_fields_match_pybind_definitions():
    assert set(vars(C.SparseGridSpec())) == {
        "density_data", "surface_data", "level_set_data", "sh_data", "links", "_offset", "_scaling", "basis_dim",
        "basis_type", "surface_type", "basis_data", "background_links", "background_data", "fake_sample_std",
        "truncated_vol_render_a"}
    assert set(vars(C.RenderOptions())) == {
        "background_brightness", "step_size", "sigma_thresh", "stop_thresh", "near_clip", "use_spheric_clip",
        "last_sample_opaque", "surf_fake_sample", "surf_fake_sample_min_vox_len", "limited_fake_sample",
        "no_surf_grad_from_sh", "alpha_activation_type", "fake_sample_l_dist", "fake_sample_normalize_surf",
        "only_outward_intersect", "truncated_vol_render", "trunc_vol_weight_min"}
    assert set(vars(C.GridOutputGrads())) == {
        "grad_density_out", "grad_sh_out", "grad_surface_out", "grad_fake_sample_std_out", "grad_basis_out",
        "grad_background_out", "mask_out", "mask_background_out"}
    assert set(vars(C.RaysSpec())) == {"origins", "dirs", "masks"}


def test_cpu_tensors_are_rejected_like_torch_check():
    sg = synth.make_shell_grid(16, basis_dim=4)
    o, d, gt = synth.make_camera_rays(8)
    grid, rays, opt = H.fill_grid_spec(C, sg), H.fill_rays_spec(C, o, d), H.fill_opt(C, synth.alphasurf_render_options())
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        C.volume_render_surf_trav(grid, rays, opt)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        C.rmsprop_step(sg.density, sg.density.clone(), sg.density.clone(), torch.empty(()), 0.9, 0.1, 1e-8, -1e9, 0.1)


def test_render_options_round_trip_into_the_abi_struct():
    from alphasurf_b200 import capi
    o = capi.make_opt(synth.alphasurf_render_options())
    assert o.only_outward_intersect == 1 and o.truncated_vol_render == 1 and abs(o.trunc_vol_weight_min - 1e-10) < 1e-16
    f = capi.make_fused(synth.alphasurf_fused_args(), norm_rays=123)
    assert f.l_dist_max_sample == 64 and f.norm_rays == 123 and abs(f.lambda_conv_mode_samp - 1e-6) < 1e-12


def test_accel_cache_is_keyed_by_tensor_identity(monkeypatch):
    """The cached occupancy pyramid of `links` must not survive a NEW tensor at the same address (caching allocator reuse
    after pruning), nor an in-place edit; it must survive repeated calls with the same tensor."""
    import types
    builds = []
    fake = types.SimpleNamespace(asurf_accel_words=lambda sz: 8,
                                 asurf_accel_build=lambda *a: builds.append(1) or 0)
    monkeypatch.setattr(C.capi, "lib", lambda: fake)
    monkeypatch.setattr(C.capi, "current_stream", lambda dev=None: None)
    monkeypatch.setattr(C, "_ACCEL", {})
    storage = torch.zeros((4, 4, 4), dtype=torch.int32)
    a = C.accel_for(storage)
    assert C.accel_for(storage) is a and len(builds) == 1          # same tensor, same version: cached
    storage.add_(1)
    b = C.accel_for(storage)
    assert b is not a and len(builds) == 2                         # in-place edit: rebuilt
    alias = storage.view(4, 4, 4)                                  # another tensor object on the same memory, same version
    assert alias.data_ptr() == storage.data_ptr() and alias._version == storage._version
    C.accel_for(alias)
    assert len(builds) == 3                                        # not trusted: rebuilt


def test_round_two_entries_check_their_inputs_like_the_reference():
    """surface_normal_grad (loss_kernel.cu:1297-1306), lumisphere_tv_grad_sparse (:1671-1676), msi_tv_grad_sparse,
    surf_sign_change_grad_sparse: CHECK_INPUT -> RuntimeError on CPU tensors; basis_fn must be 1-D; positional arity of
    svox2.cpp:117-135."""
    sg = synth.make_shell_grid(8, basis_dim=4)
    assert len(inspect.signature(C.surface_normal_grad).parameters) == 9
    assert len(inspect.signature(C.lumisphere_tv_grad_sparse).parameters) == 9
    assert len(inspect.signature(C.msi_tv_grad_sparse).parameters) == 7
    assert len(inspect.signature(C.surf_sign_change_grad_sparse).parameters) == 8
    with pytest.raises(RuntimeError):
        C.surface_normal_grad(sg.links, sg.surface, 0.0, 0, 1, 1.0, -1.0, -1.0, torch.zeros_like(sg.surface))
    cells = torch.zeros((4,), dtype=torch.int32)
    with pytest.raises(RuntimeError):
        C.surf_sign_change_grad_sparse(sg.links, sg.surface, cells, torch.zeros((0,), dtype=torch.bool), 0, 1, 1.0,
                                       torch.zeros_like(sg.surface))
    grid = H.fill_grid_spec(C, sg)
    holder = C.GridOutputGrads()
    holder.grad_sh_out = torch.zeros_like(sg.sh)
    with pytest.raises(RuntimeError):
        C.lumisphere_tv_grad_sparse(grid, cells, torch.zeros(4), torch.zeros(4), 1.0, -1.0, -1.0, 1.0, holder)
    for name in ("volume_render_surface_fused", "volume_render_nvol", "volume_render_svox1", "test_cubic_root_grad"):
        with pytest.raises(NotImplementedError):            # the out-of-scope backends exist and refuse, never fall back
            getattr(C, name)()
