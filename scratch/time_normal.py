"""A/B timing of the regulariser kernels at 512^3 (scratch)."""
import sys, torch
sys.path.insert(0, ".")
from alphasurf_b200 import svox2_csrc as ours, synth, step as S, capi
sg = synth.make_shell_grid(512, basis_dim=9, variant="G").to("cuda")
ts = S.TrainStep(ours, sg)
C, hp, g = ours, ts.hp, ts.grad
cells = ts.rand_cells_non_empty(hp["norm_surface_sparsity"])
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
for mode in (0, 0, 0, 2, 0, 0, 2):
    capi.lib().asurf_debug_set_normal_tile(mode)
    print("mode", mode, "normal ms", round(t(normal), 4), "surf tv ms", round(t(surftv), 4), "all", round(t(ts.regularisers), 4), flush=True)
capi.lib().asurf_debug_set_normal_tile(0)
