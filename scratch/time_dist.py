"""torchrun --nproc-per-node N scratch/time_dist.py : timing of the multi-GPU step variants and of the bare collectives."""
import os, sys, json, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
from alphasurf_b200 import svox2_csrc as C, synth, step as S, dist as adist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
Q = 65536
sg = synth.make_shell_grid(512, basis_dim=9, variant="G").to(dev)
ts = S.TrainStep(C, sg)
C.set_loss_norm_rays(Q * world)
batches = [synth.make_camera_rays(Q, device=dev, seed=1000 * rank + b) for b in range(4)]
rgb = torch.zeros((Q, 3), device=dev)
def timeit(fn, n=10, w=3):
    for i in range(w): fn(i)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3
out = {}
N = sg.capacity
m8 = torch.zeros((N,), dtype=torch.uint8, device=dev)
f2 = torch.zeros((2, N, 1), device=dev)
f1 = torch.zeros((N, 1), device=dev)
rows = torch.zeros((65536 * world, 29), device=dev)
out["allreduce_mask_u8_max"] = timeit(lambda i: dist.all_reduce(m8, op=dist.ReduceOp.MAX))
out["allreduce_2N_f32"] = timeit(lambda i: dist.all_reduce(f2))
out["allreduce_N_f32"] = timeit(lambda i: dist.all_reduce(f1))
out["allreduce_rows"] = timeit(lambda i: dist.all_reduce(rows))
ex_old = adist.GradExchange(ts)
def old(i):
    o, d, gt = batches[i % 4]
    ts.render(o, d, gt, rgb); ex_old.begin(ts); ts.regularisers(ex_old.rank, ex_old.world); ex_old.end(ts); ts.optimizer()
out["old_sharded_inline"] = timeit(old)
ex_ns = adist.GradExchange(ts, shard_regularisers=False)
def old_ns(i):
    o, d, gt = batches[i % 4]
    ts.render(o, d, gt, rgb); ex_ns.begin(ts); ts.regularisers(); ex_ns.end(ts); ts.optimizer()
out["old_replicated_inline"] = timeit(old_ns)
ex_l = adist.GradExchange(ts)
out["lanes_sharded"] = timeit(lambda i: ex_l.step(ts, *batches[i % 4], rgb))
ex_l2 = adist.GradExchange(ts, shard_regularisers=False)
out["lanes_replicated"] = timeit(lambda i: ex_l2.step(ts, *batches[i % 4], rgb))
if rank == 0:
    print("TIMES " + json.dumps(out))
dist.destroy_process_group()
