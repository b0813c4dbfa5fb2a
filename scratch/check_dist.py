"""torchrun --nproc-per-node N scratch/check_dist.py : N-rank ray-sharded render + exchange == 1-rank render of the whole batch."""
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as C, synth, step as S, dist as adist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
R, Qg = 256, 32768
sg = synth.make_shell_grid(R, basis_dim=9, variant="G").to(dev)
o, d, gt = synth.make_camera_rays(Qg, device=dev, seed=123)
Ql = Qg // world
sl = slice(rank * Ql, (rank + 1) * Ql)
ts = S.TrainStep(C, sg)
C.set_loss_norm_rays(Qg)
rgb = torch.zeros((Ql, 3), device=dev)
ex = adist.GradExchange(ts)
ts.optimizer = lambda: None     # compare the gradients the optimizer would see
for _ in range(2):              # twice: the second pass exercises the stream hand-over between iterations
    for k in ts.grad:
        ts.grad[k].zero_()
    ex.step(ts, o[sl].contiguous(), d[sl].contiguous(), gt[sl].contiguous(), rgb)
n = ex.last_rows
torch.cuda.synchronize()
res = {}
if rank == 0:
    C.set_loss_norm_rays(None)
    ts1 = S.TrainStep(C, sg)
    rgb1 = torch.zeros((Qg, 3), device=dev)
    for _ in range(2):          # same number of passes: the random cell windows of the regularisers advance per pass
        for k in ts1.grad:
            ts1.grad[k].zero_()
        ts1.render(o, d, gt, rgb1)
        ts1.regularisers()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    res = {"world": world, "rows_exchanged": n, "mask_equal": bool(torch.equal(ts.mask, ts1.mask)),
           "rgb_equal_slice": bool(torch.equal(rgb, rgb1[sl])),
           "rel_sh": rel(ts.grad["sh"], ts1.grad["sh"]), "rel_density": rel(ts.grad["density"], ts1.grad["density"]),
           "rel_surface": rel(ts.grad["surface"], ts1.grad["surface"])}
    print(json.dumps(res))
    assert res["mask_equal"] and res["rgb_equal_slice"] and max(res["rel_sh"], res["rel_density"], res["rel_surface"]) < 1e-4
dist.destroy_process_group()
