"""torchrun --nproc-per-node N scratch/bench_eval_dist.py : C5 (800x800 image, 640^3 grid) with the pixels sharded over N GPUs."""
import os, sys, json, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as C, synth, dist as adist
from tests import helpers as H
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
sg = synth.make_shell_grid(640, basis_dim=9, variant="G").to(dev)
o, d = synth.make_image_rays(device=dev)[:2]
grid, opt = H.fill_grid_spec(C, sg), H.fill_opt(C, synth.alphasurf_render_options())
fns = {"colour": lambda a, b: C.volume_render_surf_trav(grid, H.fill_rays_spec(C, a, b), opt),
       "depth_expected": lambda a, b: C.volume_render_expected_term_surf_trav(grid, H.fill_rays_spec(C, a, b), opt),
       "normal": lambda a, b: C.render_normal_surf_trav(grid, H.fill_rays_spec(C, a, b), opt)}
out = {"world": world, "rays": int(o.shape[0])}
for name, fn in fns.items():
    single = fn(o, d) if rank == 0 else None
    for _ in range(2):
        img = adist.render_sharded(fn, o, d)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        img = adist.render_sharded(fn, o, d)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"ms": float(t.item()), "rays_per_s": o.shape[0] / float(t.item()) * 1e3}
    if rank == 0:
        out[name]["equals_single_gpu_render"] = bool(torch.equal(img, single))
if rank == 0:
    print("EVALDIST " + json.dumps(out))
dist.destroy_process_group()
