import sys, json, torch
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as ours, synth
from tests import helpers as H
from tests.test_cuvol_gpu import _grid_spec, Grads, plenoxels_options, _with_skip_codes
ref = H.load_reference_cuda()
opts = plenoxels_options()
out = {}
for reso, Q in ((256, 5000), (256, 65536), (512, 65536)):
    sg = synth.make_shell_grid(reso, basis_dim=9, variant="G", sigma_density=True).to("cuda")
    links = _with_skip_codes(sg)
    batches = [synth.make_camera_rays(Q, device="cuda", seed=100 + i) for i in range(4)]
    res = {}
    for name, mod in (("ours", ours), ("ref", ref)):
        if mod is None: continue
        G = Grads(sg); rgb = torch.zeros((Q, 3), device="cuda")
        grid, opt, gs = _grid_spec(mod, sg, links), H.fill_opt(mod, opts), G.spec(mod)
        def call(i):
            o, d, gt = batches[i % 4]
            mod.volume_render_cuvol_fused(grid, H.fill_rays_spec(mod, o, d), opt, gt, 0.0, 0.0, rgb, gs)
        for i in range(3): call(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n): call(i)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / n
    out["%d^3 Q=%d" % (reso, Q)] = {k: {"ms": v, "rays_per_s": Q / v * 1e3} for k, v in res.items()}
print(json.dumps(out, indent=1))
