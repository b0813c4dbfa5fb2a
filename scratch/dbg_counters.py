import sys, ctypes as C, torch
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as ours, synth, capi, step as S
sg = synth.make_shell_grid(512, basis_dim=9, variant="G").to("cuda")
ts = S.TrainStep(ours, sg)
o, d, gt = synth.make_camera_rays(65536, device="cuda")
out = torch.zeros_like(o)
for seg in (1, 0):
    capi.lib().asurf_debug_set_seg(seg)
    ts.render(o, d, gt, out)
    torch.cuda.synchronize()
    c = (C.c_uint64 * 8)()
    capi.check(capi.lib().asurf_debug_counters(c), "ctr")
    print("seg", seg, list(c))
