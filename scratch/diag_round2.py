"""scratch: (1) tile path on/off timings + verdict at 512^3; (2) where the C3 surface gradient differs from the reference."""
import ctypes, os, sys, torch
os.environ.setdefault("ASURF_DEBUG_HOOKS", "1")
sys.path.insert(0, ".")
from alphasurf_b200 import svox2_csrc as ours, synth, step as S, capi
from tests import helpers as H
L = capi.lib()
sg = synth.make_shell_grid(512, basis_dim=9, variant="G").to("cuda")
ts = S.TrainStep(ours, sg)
C, hp, g = ours, ts.hp, ts.grad
cells = ts.rand_cells_non_empty(1.0)
def normal():
    C.surface_normal_grad_sparse(sg.links, sg.surface, cells, ts.mask, 0.0, 0, 1, hp["lambda_normal_loss"], 0.0,
                                 -1.0, -1.0, hp["norm_con_check"], hp["norm_ignore_empty"], True, g["surface"])
def surftv():
    C.surf_tv_grad_sparse(sg.links, sg.surface, sg.density, cells, ts.mask, 0, 1, hp["lambda_tv_surface"],
                          hp["surf_tv_ignore_edge"], hp["surf_tv_edge_value"], False, -1.0, -1.0, hp["surf_tv_alpha_dependency"], g["surface"])
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for mode in (1, 0, 1):
    L.asurf_debug_set_normal_tile(mode)
    tn = t(normal)
    v = (ctypes.c_int32 * 4)()
    if mode: capi.check(L.asurf_debug_last_verdict(v), "verdict")
    print("tile", mode, "normal ms %.4f" % tn, "surf tv ms %.4f" % t(surftv), "all %.4f" % t(ts.regularisers), "verdict", list(v), "n_cells", cells.shape[0], flush=True)
L.asurf_debug_set_normal_tile(1)

ref = H.load_reference_cuda()
opts, fused = synth.alphasurf_render_options(), synth.alphasurf_fused_args()
o, d, gt = synth.make_camera_rays(65536, device="cuda")
def run(mod, sgx):
    G = H.GradSet(sgx, "cuda", with_std=False)
    rgb = torch.zeros_like(o)
    mod.volume_render_surf_trav_fused(H.fill_grid_spec(mod, sgx), H.fill_rays_spec(mod, o, d), H.fill_opt(mod, opts), gt,
                                      *H.fused_positional(fused), rgb, G.spec(mod))
    torch.cuda.synchronize()
    return rgb, G
rgb_a, Ga = run(ours, sg)
rgb_r, Gr = run(ref, sg)
a, b = Ga.surface.double().view(-1), Gr.surface.double().view(-1)
mx = b.abs().max(); err = (a - b).abs()
print("surface rel %.3e max|ref| %.4e rows err>1e-4max: %d  >1e-5max: %d  touched %d" % (float(err.max() / mx), float(mx),
      int((err > 1e-4 * mx).sum()), int((err > 1e-5 * mx).sum()), int(Ga.mask.sum())))
top = torch.topk(err, 24).indices
inv = torch.full((sg.capacity,), -1, dtype=torch.int64, device="cuda")
flat = torch.where(sg.links.view(-1) >= 0)[0]
inv[sg.links.view(-1)[flat].long()] = flat
for i in top.tolist():
    f = int(inv[i]); x, y, z = f // (512 * 512), (f // 512) % 512, f % 512
    print("row %d vertex (%d,%d,%d) ours %.6e ref %.6e err/max %.2e" % (i, x, y, z, float(a[i]), float(b[i]), float(err[i] / mx)))
# the reference's own sensitivity: surface values moved by one ulp at random
gen = torch.Generator(device="cuda").manual_seed(0)
up = torch.rand(sg.surface.shape, device="cuda", generator=gen) < 0.5
sg2 = synth.SynthGrid(sg.links, sg.density, torch.where(up, torch.nextafter(sg.surface, sg.surface + 1), torch.nextafter(sg.surface, sg.surface - 1)),
                      sg.sh, sg.level_set, sg.offset, sg.scaling, sg.basis_dim, sg.fake_sample_std, sg.truncated_vol_render_a, dict(sg.meta))
rgb_p, Gp = run(ref, sg2)
for k in ("surface", "density", "sh"):
    print(k, "ours vs ref %.3e" % H.rel_err(getattr(Ga, k), getattr(Gr, k)), " ref(1-ulp perturbed surface) vs ref %.3e" % H.rel_err(getattr(Gp, k), getattr(Gr, k)))
print("rgb ours vs ref %.3e, ref perturbed vs ref %.3e, masks equal %s / %s" % (H.rel_err(rgb_a, rgb_r), H.rel_err(rgb_p, rgb_r), bool(torch.equal(Ga.mask, Gr.mask)), bool(torch.equal(Gp.mask, Gr.mask))))
pe = (Gp.surface.double().view(-1) - b).abs()
print("rows where the perturbed reference moves > 1e-4 max: %d; overlap with our top rows: %d" % (int((pe > 1e-4 * mx).sum()), int((pe[top] > 1e-5 * mx).sum())))
