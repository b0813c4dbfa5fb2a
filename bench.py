#!/usr/bin/env python
"""Benchmark of the alpha-Surf hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one training iteration of config C3 on one batch of synthetic rays per GPU:
fused surf_trav render forward+backward (L2 + entropy + conv-mode losses) -> density TV, surface TV, surface-normal and
opacity-sparsity regularisers -> RMSprop steps on density / surface / SH for the touched voxels, on a synthetic 512^3 SH-degree-2 shell grid.
Metric: rays/s, whole job (all ranks).  ``--impl reference`` times the CPU oracle (a port of the reference CUDA
semantics; the reference has no compiled CPU implementation) on a bounded ray sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rays/sec fwd+bwd (512^3 SH2 surface render)"
LR = dict(density=1e-2, surface=1e-5, sh=1e-3)
RMS_BETA, RMS_EPS = 0.95, 1e-8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period; the timed region of
    the default run lasts ~50 ms, too short for `nvidia-smi -lms`), with the nvidia-smi one-shot query as a fallback."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20))

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = False
        self._thread = None
        self._nvml = None

    def _loop(self):
        n = self._nvml
        try:
            h = n.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            while not self._stop:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception:
            pass

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            import threading
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            time.sleep(0.02)
        except Exception:
            self._nvml = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml thread, 5 ms period"}
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
        if not self.samples:   # fallback: one nvidia-smi query right after the timed region
            try:
                q = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in q.strip().split(",")[:2]]
                self.samples, self.max_mhz = [a], b
                out["source"] = "nvidia-smi query after the timed region"
            except Exception:
                return out
        sm = sorted(self.samples)
        out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(sm))
        return out


def algorithmic_bytes(st, Q, D, M=64):
    """compulsory bytes of the reference algorithm for one fused call, from the march counters."""
    Nv, Nl, Na, S = st["n_steps"], st["n_linked"], st["n_active"], st["n_samples"]
    cache = 12 * min(S, M * Q)
    fwd = 16 * Nv + 16 * Nl + 16 * Na + S * 32 * (1 + D) + (24 + 12) * Q + cache
    bwd = 16 * Nv + 16 * Nl + 16 * Na + S * 32 * (1 + D) + S * 2 * 32 * D + S * 2 * 32 * 2 + 8 * S + (24 + 24) * Q + cache
    return fwd, bwd


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from alphasurf_b200 import synth
    from oracle import oracle
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    sg = synth.make_shell_grid(args.reso, basis_dim=9, variant="G", device="cpu")
    og = oracle.Grid(sg)
    opts, fused = synth.alphasurf_render_options(), synth.alphasurf_fused_args()
    sample = args.ref_rays
    o, d, gt = synth.make_camera_rays(sample, device="cpu")
    grads = oracle.Grads(og)

    def step():
        oracle.surf_trav_fused(og, opts, o, d, gt, fused, grads=grads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_note="CPU step = fused render fwd+bwd of a %d-ray sample" % sample),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays x %d steps of the same 512^3 workload, fused render fwd+bwd only "
                                   "(oracle/*.c, OpenMP over rays)" % (sample, args.steps)},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_note=None):
    c = {"workload": "C3: alpha-Surf surf_trav fused render fwd+bwd + TV/normal/sparsity regularisers + RMSprop(density,"
                     "surface,sh) step, synthetic "
                     "%d^3 shell grid G(R) SH deg 2 (D=27), %d rays/step/GPU, options of surface_cuda_syn.yaml"
                     % (args.reso, args.rays),
         "grid": "%d^3" % args.reso, "rays_per_step_per_gpu": args.rays, "sh_dim": 27,
         "l2_policy": "inputs larger than L2 (grid data 2.3 GB); a different ray batch every step",
         "parallelism": "ray-sharded dp%d, grid replicated" % args.gpus}
    if sample_note:
        c["note"] = sample_note
    return c


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from alphasurf_b200 import capi, synth
    from alphasurf_b200 import step as S
    from alphasurf_b200 import svox2_csrc as C
    import ctypes

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    Q, D = args.rays, 27
    sg = synth.make_shell_grid(args.reso, basis_dim=9, variant="G", device="cpu").to(dev)
    ts = S.TrainStep(C, sg)
    NB = args.batches
    dev_batches, host_batches = [], []
    for b in range(NB):
        o, d, gt = synth.make_camera_rays(Q, device="cpu", seed=synth.SEED + 1000 * rank + b)
        host_batches.append(tuple(t.pin_memory() for t in (o, d, gt)))
        dev_batches.append(tuple(t.to(dev) for t in (o, d, gt)))
    rgb_out = torch.zeros((Q, 3), dtype=torch.float32, device=dev)
    rgb_host = torch.zeros((Q, 3), dtype=torch.float32).pin_memory()
    stage = tuple(torch.empty((Q, 3), dtype=torch.float32, device=dev) for _ in range(3))
    exchange = None
    if world > 1:
        from alphasurf_b200 import dist as adist
        C.set_loss_norm_rays(Q * world)
        exchange = adist.GradExchange(ts)
    L = capi.lib()
    phase_ev = []   # per step: 4 events (start, after render [+ exchange], after regularisers, after optimizer)

    def device_step(o, d, gt, record=False):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if exchange is not None:
            # render + sparse row exchange on this stream; cell-sharded regularisers + their dense exchange on a side stream
            # and a second communicator, joined before the optimizer (alphasurf_b200/dist.py::GradExchange.step)
            exchange.step(ts, o, d, gt, rgb_out, events=evs)
            if record:
                phase_ev.append(evs)
            return
        if record:
            evs[0].record()
        ts.render(o, d, gt, rgb_out)
        if record:
            evs[1].record()
        ts.regularisers()
        if record:
            evs[2].record()
        ts.optimizer()
        if record:
            evs[3].record()
            phase_ev.append(evs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # march counters of batch 0 (outside the timed region) -> algorithmic bytes
    st = C.render_stats(ts.grid_spec, S.rays_to_cpp(C, dev_batches[0][0], dev_batches[0][1]), ts.opt_spec)
    fwd_bytes, bwd_bytes = algorithmic_bytes(st, Q, D)

    # ---------------- device-resident timing ----------------
    for i in range(args.warmup):
        device_step(*dev_batches[i % NB])
    capi.check(L.asurf_profile_enable(ctypes.c_int32(args.steps)), "profile_enable")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # BEFORE the barrier: its start-up sleep must not delay rank 0 against the other ranks
    barrier()
    L.asurf_launch_count(ctypes.c_int32(1))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        device_step(*dev_batches[(args.warmup + i) % NB], record=True)
    ev1.record()
    barrier()
    n_launch = int(L.asurf_launch_count(ctypes.c_int32(0)))
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ncalls, stages = ctypes.c_int32(0), (ctypes.c_float * 4)()
    capi.check(L.asurf_profile_read_stages(ctypes.byref(ncalls), stages), "profile_read_stages")
    capi.check(L.asurf_profile_enable(ctypes.c_int32(0)), "profile_disable")
    stage_ms = [v / max(ncalls.value, 1) for v in stages]
    fwd_ms = stage_ms[0] + stage_ms[1] + stage_ms[2]    # work pyramid build + pre-march + forward shading
    bwd_ms = stage_ms[3]
    ms_step = ms_total / args.steps
    value = Q * world * args.steps / (ms_total * 1e-3)
    phases = {"render_ms": 0.0, "regularisers_ms": 0.0, "optimizer_ms": 0.0}
    for evs in phase_ev:
        phases["render_ms"] += evs[0].elapsed_time(evs[1]) / len(phase_ev)
        phases["regularisers_ms"] += evs[1].elapsed_time(evs[2]) / len(phase_ev)
        phases["optimizer_ms"] += evs[2].elapsed_time(evs[3]) / len(phase_ev)

    # ---------------- end to end: host buffers in, colours out, through the svox2.csrc-compatible API ----------------
    def e2e_step(b):
        ho, hd, hgt = host_batches[b]
        stage[0].copy_(ho, non_blocking=True)
        stage[1].copy_(hd, non_blocking=True)
        stage[2].copy_(hgt, non_blocking=True)
        device_step(stage[0], stage[1], stage[2])
        rgb_host.copy_(rgb_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(((rgb_host - hgt) ** 2).mean())  # the mse opt.py logs every step (opt/opt.py:832-860)

    for i in range(max(args.warmup, 3)):
        e2e_step(i % NB)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step((args.warmup + i) % NB)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = Q * world * args.steps / e2e_s

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # the unit the north_star names: the fused surface-render call, forward + backward passes together
    dom = ("fused", fwd_ms + bwd_ms, fwd_bytes + bwd_bytes)
    achieved = dom[2] / (dom[1] * 1e-3) / 1e9 if dom[1] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("fused")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": 3 * Q * 3 * 4, "d2h_bytes_per_step": Q * 3 * 4,
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": n_launch,
        "roofline": {"bound": "hbm", "kernel": "volume_render_surf_trav_fused: work pyramid update + pre-march + wavefront "
                                               "shading kernels, forward and backward (one call)",
                     "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom[2], "kernel_ms": dom[1],
                     "kernels": {"forward_ms": fwd_ms, "backward_ms": bwd_ms,
                                 "stages_ms": {"work_pyramid": stage_ms[0], "premarch": stage_ms[1],
                                               "forward_shading": stage_ms[2], "backward": stage_ms[3]},
                                 "forward_bytes": fwd_bytes,
                                 "backward_bytes": bwd_bytes,
                                 "fused_frac": (fwd_bytes + bwd_bytes) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak
                                 if fwd_ms + bwd_ms > 0 else 0.0},
                     "counters_per_ray": {k: st[k] / Q for k in ("n_steps", "n_linked", "n_active", "n_samples")}},
        "step_phases": phases,
        "step_phases_note": ("sequential: fused render | regularisers | optimizer" if world == 1 else
                             "two lanes (dist.GradExchange.step): render_ms = render + mask OR + row pack on the main stream while the "
                             "cell-sharded regularisers and their dense all-reduce run on a side stream; regularisers_ms = sparse row "
                             "all-reduce wait + join of the lanes; optimizer_ms = RMSprop steps"),
        "render_only_rays_per_s": Q * world / ((fwd_ms + bwd_ms) * 1e-3) if fwd_ms + bwd_ms > 0 else None,
    }
    if world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_baseline(args, sg, ts.opts, ts.fused, host_batches[0])
        ref = reference_cuda_same_gpu(args, sg, dev_batches)
        if ref is not None:
            line["reference_cuda_same_gpu"] = ref
    print(json.dumps(line), flush=True)


def cpu_baseline(args, sg, opts, fused, batch0):
    """The oracle port on the host cores of this box: bounded sample of the same workload (render fwd+bwd)."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    og = oracle.Grid(sg.to("cpu"))
    n = args.ref_rays
    o, d, gt = (t[:n] for t in batch0)
    grads = oracle.Grads(og)
    oracle.surf_trav_fused(og, opts, o, d, gt, fused, grads=grads)
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.surf_trav_fused(og, opts, o, d, gt, fused, grads=grads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > args.cpu_seconds or reps >= 200:
            break
    return {"value": n * reps / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "first %d rays of batch 0, %d repeats (%.1f s), fused render fwd+bwd only, oracle/*.c with OpenMP"
                      % (n, reps, dt)}


def reference_cuda_same_gpu(args, sg, dev_batches):
    """Extra (not part of the contract): the UNMODIFIED reference CUDA kernels (oracle/_ref) on the same GPU, driven
    through the same TrainStep sequence (its own copy of the grid)."""
    import torch
    from alphasurf_b200 import step as S
    from tests import helpers as H
    try:
        ref = H.load_reference_cuda(required=False)
    except Exception as e:  # noqa
        return {"unavailable": repr(e)[:200]}
    if ref is None:
        return None
    from alphasurf_b200 import synth
    sg2 = synth.SynthGrid(sg.links, sg.density.clone(), sg.surface.clone(), sg.sh.clone(), sg.level_set, sg.offset,
                          sg.scaling, sg.basis_dim, sg.fake_sample_std, sg.truncated_vol_render_a, dict(sg.meta))
    ts = S.TrainStep(ref, sg2)
    Q = dev_batches[0][0].shape[0]
    out = torch.zeros((Q, 3), dtype=torch.float32, device=sg.density.device)
    n, ev = 5, [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(2):
        ts.step(*dev_batches[i % len(dev_batches)], out)
    torch.cuda.synchronize()
    t_render = t_rest = 0.0
    for i in range(n):
        o, d, gt = dev_batches[(2 + i) % len(dev_batches)]
        ev[0].record()
        ts.render(o, d, gt, out)
        ev[1].record()
        ts.regularisers()
        ts.optimizer()
        ev[2].record()
        torch.cuda.synchronize()
        t_render += ev[0].elapsed_time(ev[1]) / n
        t_rest += ev[1].elapsed_time(ev[2]) / n
    return {"what": "reference svox2.csrc kernels, same C3 step sequence (fused render, regularisers, RMSprop)",
            "render_ms": t_render, "regularisers_optimizer_ms": t_rest, "ms_per_step": t_render + t_rest,
            "rays_per_s": Q / ((t_render + t_rest) * 1e-3), "render_only_rays_per_s": Q / (t_render * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reso", type=int, default=512)
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--ref-rays", type=int, default=16384)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline and the reference-CUDA comparison (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # stdout carries the one JSON line only: NCCL announces its version there when the first communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
