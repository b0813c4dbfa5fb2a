"""scratch: queue counters + timing of the fused call on G*(512) (sample-dominated), before/after the capacity feedback."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from alphasurf_b200 import svox2_csrc as ours, synth, capi
from tests import helpers as H
L = capi.lib()
variant = sys.argv[1] if len(sys.argv) > 1 else "G*"
sg = synth.make_shell_grid(512, basis_dim=9, variant=variant).to("cuda")
opts, fused = synth.alphasurf_render_options(), synth.alphasurf_fused_args()
o, d, gt = synth.make_camera_rays(65536, device="cuda")
G = H.GradSet(sg, "cuda", with_std=False)
rgb = torch.zeros_like(o)
grid, rays, opt = H.fill_grid_spec(ours, sg), H.fill_rays_spec(ours, o, d), H.fill_opt(ours, opts)
for it in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ours.volume_render_surf_trav_fused(grid, rays, opt, gt, *H.fused_positional(fused), rgb, G.spec(ours))
    b.record(); torch.cuda.synchronize()
    c = (ctypes.c_uint64 * 8)()
    capi.check(L.asurf_debug_counters(c), "counters")
    print(variant, "call", it, "ms %.3f" % a.elapsed_time(b), "long rays", c[2], "short rays", c[3], "items", c[4], "samples", c[5], "fine items", c[6], flush=True)
