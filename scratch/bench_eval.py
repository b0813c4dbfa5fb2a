"""C5 of SURVEY.md 8d: full-image evaluation renders (800x800, 640^3 grid) -- colour, depth, normal -- ours vs the
UNMODIFIED reference kernels on the same GPU, in 5000-ray chunks (what svox2.py does) and as one call."""
import sys, json, torch
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as ours, synth
from tests import helpers as H
if "--shim" in sys.argv:          # the compiled svox2.csrc module instead of the ctypes layer (lower per-call host cost)
    sys.argv.remove("--shim")
    from alphasurf_b200 import build_shim
    ours = build_shim.load()
ref = H.load_reference_cuda()
opts = synth.alphasurf_render_options()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 640
sg = synth.make_shell_grid(R, basis_dim=9, variant="G").to("cuda")
o, d = synth.make_image_rays(device="cuda")[:2]
Q = o.shape[0]
out = {"grid": "%d^3" % R, "rays": Q}

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

for name, mod in (("ours", ours), ("reference_cuda", ref)):
    if mod is None: continue
    grid, opt = H.fill_grid_spec(mod, sg), H.fill_opt(mod, opts)
    def chunks(f, extra=()):
        return torch.cat([f(grid, H.fill_rays_spec(mod, o[i:i + 5000].contiguous(), d[i:i + 5000].contiguous()), opt, *extra)
                          for i in range(0, Q, 5000)])
    whole = H.fill_rays_spec(mod, o, d)
    res = {}
    for label, f, extra in (("colour", mod.volume_render_surf_trav, ()), ("depth_expected", mod.volume_render_expected_term_surf_trav, ()),
                            ("depth_mode", mod.volume_render_mode_term_surf_trav, (0.1,)), ("normal", mod.render_normal_surf_trav, ())):
        ms_c, img_c = timed(lambda: chunks(f, extra))
        ms_w, img_w = timed(lambda: f(grid, whole, opt, *extra))
        assert torch.equal(img_c, img_w)
        res[label] = {"ms_5000_ray_chunks": ms_c, "ms_one_call": ms_w, "rays_per_s_one_call": Q / ms_w * 1e3, "img": img_w}
    out[name] = res
if ref is not None:
    for k in out["ours"]:
        a, b = out["ours"][k].pop("img"), out["reference_cuda"][k].pop("img")
        out["ours"][k]["max_rel_err_vs_reference"] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        out["ours"][k]["speedup_one_call"] = out["reference_cuda"][k]["ms_one_call"] / out["ours"][k]["ms_one_call"]
else:
    for k in out["ours"]:
        out["ours"][k].pop("img")
print("EVAL " + json.dumps(out))
