import sys, torch
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as ours, synth
from tests import helpers as H
from tests.test_surf_trav_gpu import _full_fused, _setup
ref = H.load_reference_cuda()
opts = synth.alphasurf_render_options(); fused = _full_fused(synth.alphasurf_fused_args())
sg, o, d, gt = _setup(64, 9, 2048, "G*", opts)
out_r = ref.volume_render_surf_trav(H.fill_grid_spec(ref, sg), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts))
worst = (0, None)
for seed in range(30):
    gout = torch.randn(out_r.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed))
    G2, G2r = H.GradSet(sg, "cuda"), H.GradSet(sg, "cuda")
    ours.volume_render_surf_trav_backward(H.fill_grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts), gout, out_r, G2.spec(ours))
    ref.volume_render_surf_trav_backward(H.fill_grid_spec(ref, sg), H.fill_rays_spec(ref, o, d), H.fill_opt(ref, opts), gout, out_r, G2r.spec(ref))
    torch.cuda.synchronize()
    a, b = G2.surface.double().view(-1), G2r.surface.double().view(-1)
    mx = b.abs().max()
    err = (a - b).abs()
    e = float(err.max() / mx)
    i = int(err.argmax())
    nbig = int((err > 1e-5 * mx).sum())
    print("seed %d rel %.3e  at %d ours %.6e ref %.6e  max|ref| %.4e  n(err>1e-5 max) %d  sh %.2e dens %.2e" % (
        seed, e, i, float(a[i]), float(b[i]), float(mx), nbig, H.rel_err(G2.sh, G2r.sh), H.rel_err(G2.density, G2r.density)))
