"""scratch: one small pass over every hot-path kernel family, to be run under compute-sanitizer (memcheck / racecheck):
fused surf_trav render (wavefront + persistent paths), tiled surface TV + normal loss, list-kernel regularisers, RMSprop,
MSI background, cuvol fused."""
import os, sys
os.environ.setdefault("ASURF_DEBUG_HOOKS", "1")
sys.path.insert(0, ".")
import torch
from alphasurf_b200 import svox2_csrc as C, synth, step as S, capi
from tests import helpers as H
from tests.test_msi_gpu import make_background, grid_spec, Grads

R, Q = 48, 512
for variant in ("G", "G*"):
    sg = synth.make_shell_grid(R, basis_dim=9, variant=variant).to("cuda")
    ts = S.TrainStep(C, sg)
    out = torch.zeros((Q, 3), device="cuda")
    for it in range(2):
        o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=7 + it)
        ts.step(o, d, gt, out)
    capi.lib().asurf_debug_set_wave(0)          # persistent shading kernels
    ts.step(o, d, gt, out)
    capi.lib().asurf_debug_set_wave(1)
    capi.lib().asurf_debug_set_normal_tile(0)   # list kernels of the regularisers
    ts.regularisers()
    capi.lib().asurf_debug_set_normal_tile(1)
    torch.cuda.synchronize()
    print(variant, "train steps ok", float(out.mean()))
sg = synth.make_shell_grid(R, basis_dim=9, variant="G").to("cuda")
bg = make_background(reso=16, nlayers=8)
o, d, gt = synth.make_camera_rays(Q, device="cuda", seed=3)
grid, rays, opt = grid_spec(C, sg, bg), H.fill_rays_spec(C, o, d), H.fill_opt(C, synth.alphasurf_render_options())
G = Grads(sg, bg)
rgb = torch.zeros_like(o)
C.volume_render_surf_trav_fused(grid, rays, opt, gt, *H.fused_positional(synth.alphasurf_fused_args()), rgb, G.spec(C))
torch.cuda.synchronize()
print("msi fused ok", float(rgb.mean()))
sgc = synth.make_shell_grid(R, basis_dim=9, variant="G", sigma_density=True).to("cuda")
C.accel_dist_prop(sgc.links)
cs = S.CuvolStep(C, sgc)
cs.step(o, d, gt, rgb)
torch.cuda.synchronize()
print("cuvol step ok", float(rgb.mean()))
