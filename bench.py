#!/usr/bin/env python
"""Benchmark of the alpha-Surf hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c3|c3-gstar|c2] [--scaling weak|strong]

One "step" = one training iteration on one batch of synthetic rays:
  c3        (default, the headline) config C3 of SURVEY.md 8d: fused surf_trav render forward+backward (L2 + entropy + conv-mode
            losses) -> density TV, surface TV, surface-normal and opacity-sparsity regularisers -> RMSprop on density /
            surface / SH for the touched voxels, synthetic 512^3 SH-degree-2 shell grid G(512), 65 536 rays
  c3-gstar  the same step on the stress grid G*(512) (a level-set crossing in almost every voxel: sample-dominated)
  c2        config C2: Plenoxels cuvol fused render + sigma / SH TV + RMSprop, synthetic 256^3 SH-degree-2 grid, 5000 rays
Metric: rays/s, whole job (all ranks).  --scaling weak: the batch above PER GPU; strong: the batch split over the ranks (C4).
``--impl reference`` times the CPU restatement of the same step (oracle/*.c, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rays/sec fwd+bwd (512^3 SH2 surface render)"
WORKLOADS = {
    "c3": dict(reso=512, rays=65536, variant="G", kind="surf_trav",
               name="C3: alpha-Surf surf_trav fused render fwd+bwd + TV/normal/sparsity regularisers + RMSprop(density,surface,sh) "
                    "step, synthetic %d^3 shell grid G(R) SH deg 2 (D=27), %d rays/step/GPU, options of surface_cuda_syn.yaml"),
    "c3-gstar": dict(reso=512, rays=65536, variant="G*", kind="surf_trav",
                     name="C3 on the stress grid G*(R) (SURVEY 8d 'report both'; level sets of test_render_gradcheck_surface.py:73-77, "
                          "a crossing in almost every voxel): same step, synthetic %d^3 shell grid SH deg 2, %d rays/step/GPU"),
    "c2": dict(reso=256, rays=5000, variant="G", kind="cuvol",
               name="C2: Plenoxels cuvol fused render fwd+bwd + sigma/SH TV + RMSprop(sigma,sh) step, synthetic %d^3 shell grid "
                    "SH deg 2 (D=27), sigma ~ N(20,5), %d rays/step/GPU, options of configs/syn.yaml"),
}


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


# ---- algorithmic bytes ----------------------------------------------------------------------------------------
def surf_trav_bytes(st, Q, D, M=64):
    """Compulsory bytes of the REFERENCE algorithm for one fused surf_trav call, from the march counters."""
    Nv, Nl, Na, S = st["n_steps"], st["n_linked"], st["n_active"], st["n_samples"]
    cache = 12 * min(S, M * Q)
    fwd = 16 * Nv + 16 * Nl + 16 * Na + S * 32 * (1 + D) + (24 + 12) * Q + cache
    bwd = 16 * Nv + 16 * Nl + 16 * Na + S * 32 * (1 + D) + S * 2 * 32 * D + S * 2 * 32 * 2 + 8 * S + (24 + 24) * Q + cache
    return fwd, bwd


def surf_trav_own_bytes(st, Q, D, n_rows, M=64):
    """Compulsory bytes of THIS implementation's algorithm for the same call: the per-voxel link reads of the reference are
    replaced by one bit of the work pyramid (L2 resident: not DRAM traffic); what remains is the row-order class scan that
    validates the pyramid (8 B per stored row), the 8-link / surface / density reads of the voxels that need work, and the
    per-sample gathers and read-modify-writes, which are the same as the reference's."""
    Na, S = st["n_active"], st["n_samples"]
    cache = 12 * min(S, M * Q)
    fwd = 8 * n_rows + Na * (32 + 32 + 32) + S * 32 * (1 + D) + (24 + 12) * Q + cache
    bwd = S * 32 * (1 + D) + S * 2 * 32 * D + S * 2 * 32 * 2 + 8 * S + (24 + 24) * Q + cache
    return fwd + bwd


def cuvol_bytes(st, Q, D):
    """cuvol fused call: per sample position one skip link; per gathered sample 4 new links + 4 new densities; per
    contributing sample 8 x D SH forward, and the same again + SH / density RMW + mask in the backward."""
    P, G, S = st["n_steps"], st["n_linked"], st["n_samples"]
    fwd = 4 * P + 32 * G + S * 32 * D + (24 + 12) * Q
    bwd = 4 * P + 32 * G + S * 32 * D + S * (2 * 32 * D + 2 * 32 + 8) + (24 + 24) * Q
    return fwd, bwd


def load_traffic(workload):
    """DRAM bytes per call measured by `ncu --set full` (profiles/traffic.json, written from the committed ncu summaries)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(workload) or {}
    except Exception:
        return {}


def workload_config(args, world, sample_note=None):
    w = WORKLOADS[args.workload]
    rays_rank = args.rays if args.scaling == "weak" else args.rays // world
    c = {"workload": w["name"] % (args.reso, rays_rank), "grid": "%d^3" % args.reso, "rays_per_step_per_gpu": rays_rank,
         "rays_per_step_global": rays_rank * world, "sh_dim": 27, "scaling": args.scaling,
         "l2_policy": "inputs larger than L2 (grid data %s); a different ray batch every step" %
                      ("2.3 GB" if args.reso >= 512 else "0.3 GB"),
         "parallelism": "ray-sharded dp%d, grid replicated%s" % (
             world, "" if world == 1 else (", regularisers cell-sharded" if args.shard_regularisers else ", regularisers redundant"))}
    if sample_note:
        c["note"] = sample_note
    return c


# ---- the CPU arm --------------------------------------------------------------------------------------------------------------
def omp_threads():
    import ctypes
    try:
        return int(ctypes.CDLL("libgomp.so.1").omp_get_max_threads())
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


class CpuStep:
    """The same C3 / C2 iteration on the host cores, on a bounded sample: n rays of the batch through the oracle's renderer
    (oracle/*.c, OpenMP over rays) and the fraction f = n / batch of every regulariser list and of the optimizer rows, so that
    n / time estimates the throughput of the whole step.  Test infrastructure driven from the benchmark's baseline legs only."""

    def __init__(self, args, sg_cpu):
        import numpy as np
        from alphasurf_b200 import step as S, synth
        from oracle import oracle
        self.np, self.oracle, self.S = np, oracle, S
        self.kind = WORKLOADS[args.workload]["kind"]
        self.n = min(args.ref_rays, args.rays)
        self.f = self.n / float(args.rays)
        self.sg = sg_cpu
        self.og = oracle.Grid(sg_cpu)
        self.grads = oracle.Grads(self.og)
        self.n_vert = sg_cpu.links.numel()
        self.non_empty = np.nonzero(sg_cpu.links.numpy().reshape(-1) >= 0)[0].astype(np.int32)
        self.rng = np.random.RandomState(7)
        if self.kind == "surf_trav":
            self.opts, self.fused, self.hp = synth.alphasurf_render_options(), synth.alphasurf_fused_args(), S.c3_hyper()
        else:
            self.opts, self.hp = S.plenoxels_render_options(), S.c2_hyper()
        N = sg_cpu.capacity
        self.rms = {k: np.zeros(tuple(getattr(sg_cpu, k).shape), np.float32) for k in ("density", "sh")}
        if self.kind == "surf_trav":
            self.rms["surface"] = np.zeros((N, 1), np.float32)
        self.mask = np.zeros((N,), np.uint8)

    def _window(self, frac_of_list, n_list):
        n = max(int(n_list * frac_of_list * self.f), 1)
        start = int(self.rng.randint(0, n_list - n + 1))
        return start, n

    def step(self, o, d, gt):
        np, oracle, hp, sg, g = self.np, self.oracle, self.hp, self.sg, self.grads
        o, d, gt = o[:self.n], d[:self.n], gt[:self.n]
        self.mask[:] = 0
        if self.kind == "surf_trav":
            oracle.surf_trav_fused(self.og, self.opts, o, d, gt, self.fused, grads=g)
            s, n = self._window(hp["tv_sparsity"], self.n_vert)
            oracle.tv_grad_sparse(sg.links, sg.density, None, np.arange(s, s + n, dtype=np.int32), self.mask, 0, 1,
                                  hp["lambda_tv_alpha"], False, 0.0, False, False, False, g.density)
            s, n = self._window(1.0, self.non_empty.shape[0])
            cells = self.non_empty[s:s + n]
            oracle.tv_grad_sparse(sg.links, sg.surface, sg.density, cells, self.mask, 0, 1, hp["lambda_tv_surface"], True, -1.0,
                                  False, False, True, g.surface)
            oracle.surface_normal_grad_sparse(sg.links, sg.surface, cells, self.mask, 0.0, 0, 1, hp["lambda_normal_loss"], False,
                                              False, True, g.surface)
            s, n = self._window(hp["alpha_surf_sparsify_sparsity"], self.non_empty.shape[0])
            oracle.alpha_surf_sparsify(sg.links, sg.density, sg.surface, self.non_empty[s:s + n], self.mask,
                                       hp["lambda_sparsify_alpha"], hp["lambda_sparsify_surf"], True, 0.15, 0.0, -0.1, g.density,
                                       g.surface)
            tensors = (("density", hp["lr_density"]), ("surface", hp["lr_surface"]), ("sh", hp["lr_sh"]))
        else:
            oracle.cuvol_fused(self.og, self.opts, o, d, gt, grads=g)
            s, n = self._window(hp["tv_sparsity"], self.n_vert)
            oracle.tv_grad_sparse(sg.links, sg.density, None, np.arange(s, s + n, dtype=np.int32), self.mask, 0, 1, hp["lambda_tv"],
                                  False, 0.0, False, False, False, g.density)
            s, n = self._window(hp["tv_sh_sparsity"], self.n_vert)
            oracle.tv_grad_sparse(sg.links, sg.sh, None, np.arange(s, s + n, dtype=np.int32), self.mask, 0, sg.sh.shape[1],
                                  hp["lambda_tv_sh"], True, 0.0, False, False, False, g.sh)
            tensors = (("density", hp["lr_sigma"]), ("sh", hp["lr_sh"]))
        # the rows of the sample (already the fraction f of the step's): SH moves where the render touched, the rest where
        # the render or a regulariser did (sparse_sh_grad_indexer / sparse_grad_indexer, svox2.py:3637)
        rows_sh = np.nonzero(g.mask)[0].astype(np.int64)
        rows = np.nonzero(np.maximum(self.mask, g.mask))[0].astype(np.int64)
        for k, lr in tensors:
            oracle.rmsprop_step(getattr(sg, k).numpy(), self.rms[k], getattr(g, k), rows_sh if k == "sh" else rows,
                                self.S.RMS_BETA, lr, self.S.RMS_EPS, -1e9, lr)
        g.mask[:] = 0

    def note(self):
        return ("%d rays of the batch + the fraction %.3g of every regulariser list and optimizer row set "
                "(oracle/*.c, OpenMP; RMSprop in numpy)" % (self.n, self.f))


def make_grid(args, device="cpu"):
    from alphasurf_b200 import synth
    w = WORKLOADS[args.workload]
    return synth.make_shell_grid(args.reso, basis_dim=9, variant=w["variant"], device=device,
                                 sigma_density=(w["kind"] == "cuvol"))


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the same step on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from alphasurf_b200 import synth
    sg = make_grid(args)
    cpu = CpuStep(args, sg)
    cores = omp_threads()
    o, d, gt = synth.make_camera_rays(cpu.n, device="cpu")
    for _ in range(args.warmup):
        cpu.step(o, d, gt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.step(o, d, gt)
    dt = time.perf_counter() - t0
    val = cpu.n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1), sample_note="CPU step = " + cpu.note()),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%s x %d steps" % (cpu.note(), args.steps)},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, sg, batch0):
    """The same step on the host cores of this box (bounded sample, see CpuStep)."""
    cpu = CpuStep(args, sg.to("cpu"))
    o, d, gt = batch0
    cpu.step(o, d, gt)
    reps, t0 = 0, time.perf_counter()
    while True:
        cpu.step(o, d, gt)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > args.cpu_seconds or reps >= 200:
            break
    return {"value": cpu.n * reps / dt, "unit": "rays/s", "cores": omp_threads(), "kind": "port",
            "sample": "%s, %d repeats (%.1f s)" % (cpu.note(), reps, dt)}


def cpu_reference_l0(args):
    """Config C1 of BASELINE.md 2a: the reference's OWN pure-PyTorch gradcheck renderer (svox2/svox2.py:1596-2857, unmodified,
    staged in oracle/_ref/pyref) on the host cores: 128^3, SH deg 1, forward + backward, bounded to --l0-rays rays."""
    try:
        import torch
        from alphasurf_b200 import synth
        from oracle import ref_l0
        if not ref_l0.available():
            return {"unavailable": "oracle/_ref/pyref (the reference's Python package) is not staged"}
        sg = synth.make_shell_grid(128, basis_dim=4, variant="G*")
        n = args.l0_rays
        o, d, _ = synth.make_camera_rays(n, device="cpu")
        opts = synth.parity_render_options()
        t0 = time.perf_counter()
        ref_l0.render_l0(sg, opts, o, d, run_backward=False)
        t_f = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref_l0.render_l0(sg, opts, o, d, run_backward=True)
        t_fb = time.perf_counter() - t0
        return {"kind": "reference", "what": "SparseGrid.volume_render(use_kernel=False) -> _surface_render_gradcheck_lerp, "
                "C1: 128^3 G* grid, SH deg 1 (D=12), parity option set of test_render_gradcheck_surface.py:44-62",
                "rays": n, "forward_rays_per_s": n / t_f, "forward_backward_rays_per_s": n / t_fb,
                "torch_threads": torch.get_num_threads(), "host_cores": os.cpu_count()}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


# ---- the GPU arm --------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes
    import torch
    import torch.distributed as dist
    from alphasurf_b200 import capi, synth
    from alphasurf_b200 import step as S
    from alphasurf_b200 import build_shim

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    C = build_shim.load()      # the compiled svox2.csrc replacement (csrc/host/svox2_shim.cpp over the C ABI): the drop-in itself
    w = WORKLOADS[args.workload]
    Q = args.rays if args.scaling == "weak" else args.rays // world
    D = 27
    sg = make_grid(args).to(dev)
    if w["kind"] == "cuvol":
        C.accel_dist_prop(sg.links)          # the skip codes the cuvol marcher reads (SparseGrid.accelerate())
        ts = S.CuvolStep(C, sg)
    else:
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
        assert w["kind"] == "surf_trav", "the multi-GPU step is implemented for the alpha-Surf workloads"
        from alphasurf_b200 import dist as adist
        C.set_loss_norm_rays(Q * world)
        exchange = adist.GradExchange(ts, shard_regularisers=(args.shard_regularisers != 0))
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
    rays0 = S.rays_to_cpp(C, dev_batches[0][0], dev_batches[0][1])
    if w["kind"] == "cuvol":
        st = C.cuvol_render_stats(ts.grid_spec, rays0, ts.opt_spec)
        fwd_bytes, bwd_bytes = cuvol_bytes(st, Q, D)
        own_bytes = fwd_bytes + bwd_bytes
    else:
        st = C.render_stats(ts.grid_spec, rays0, ts.opt_spec)
        fwd_bytes, bwd_bytes = surf_trav_bytes(st, Q, D)
        own_bytes = surf_trav_own_bytes(st, Q, D, sg.capacity)

    # ---------------- device-resident timing ----------------
    for i in range(args.warmup):
        device_step(*dev_batches[i % NB])
    capi.check(L.asurf_profile_enable(ctypes.c_int32(args.steps)), "profile_enable")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # BEFORE the barrier: its start-up sleep must not delay rank 0 against the other ranks
    barrier()
    L.asurf_launch_count(ctypes.c_int32(1))
    if exchange is not None:
        exchange.host_ms, exchange.host_steps = {}, 0      # host-time sections of the timed steps only
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        device_step(*dev_batches[(args.warmup + i) % NB], record=True)
    ev1.record()
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host0) / args.steps   # CPU time to ENQUEUE a step (no synchronisation inside)
    host_sections = None
    if exchange is not None:
        host_sections = {k: v / max(exchange.host_steps, 1) for k, v in exchange.host_ms.items()}
    barrier()
    n_launch = int(L.asurf_launch_count(ctypes.c_int32(0)))
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ncalls, stages = ctypes.c_int32(0), (ctypes.c_float * 4)()
    capi.check(L.asurf_profile_read_stages(ctypes.byref(ncalls), stages), "profile_read_stages")
    capi.check(L.asurf_profile_enable(ctypes.c_int32(0)), "profile_disable")
    stage_ms = [v / max(ncalls.value, 1) for v in stages]
    ms_step = ms_total / args.steps
    value = Q * world * args.steps / (ms_total * 1e-3)
    phases = {"render_ms": 0.0, "regularisers_ms": 0.0, "optimizer_ms": 0.0}
    for evs in phase_ev:
        phases["render_ms"] += evs[0].elapsed_time(evs[1]) / len(phase_ev)
        phases["regularisers_ms"] += evs[1].elapsed_time(evs[2]) / len(phase_ev)
        phases["optimizer_ms"] += evs[2].elapsed_time(evs[3]) / len(phase_ev)
    if w["kind"] == "cuvol":      # no in-library stage events for the cuvol call: the render phase IS the fused call
        fwd_ms, bwd_ms = None, None
        kernel_ms = phases["render_ms"]
    else:
        fwd_ms = stage_ms[0] + stage_ms[1] + stage_ms[2]    # work pyramid update + pre-march + forward shading
        bwd_ms = stage_ms[3]
        kernel_ms = fwd_ms + bwd_ms

    # ---------------- per-call table of the step (single GPU: the calls run back to back on one stream) ----------------
    call_table = None
    if world == 1:
        call_table = time_calls(torch, ts, w["kind"], dev_batches, rgb_out, args)

    # ---------------- end to end: host buffers in, colours out, through the compiled svox2.csrc module ----------------
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

    parity = breakdown = None
    if world > 1 and not args.no_extras:      # collective: every rank takes part, rank 0 reports
        breakdown = exchange.collective_breakdown(ts)
        breakdown["host_enqueue_ms_by_section"] = host_sections
        parity = dist_parity(torch, dist, C, S, synth, sg, rank, world, Q, dev)

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # the ncu captures are of the workload's own size: another grid size / ray count has no measured traffic
    traffic = load_traffic(args.workload) if (Q == w["rays"] and args.reso == w["reso"]) else {}
    alg = fwd_bytes + bwd_bytes
    achieved = alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    tr_fused = traffic.get("fused")
    roof = {"bound": "hbm",
            "kernel": ("volume_render_surf_trav_fused: work pyramid update + pre-march + wavefront shading kernels, forward and "
                       "backward (one call)") if w["kind"] == "surf_trav" else
                      "volume_render_cuvol_fused: forward kernel + backward kernel (one call)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": tr_fused,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_ms,
            "frac_note": "achieved = SURVEY 8d algorithmic bytes of the REFERENCE algorithm / measured time: a speed score "
                         "against the reference's compulsory traffic, NOT DRAM utilisation; dram_frac is the physical figure",
            "dram_frac": (tr_fused / (kernel_ms * 1e-3) / 1e9 / peak) if (tr_fused and kernel_ms > 0) else None,
            "own_algorithmic_bytes": own_bytes,
            "own_frac": own_bytes / (kernel_ms * 1e-3) / 1e9 / peak if kernel_ms > 0 else None,
            "counters_per_ray": {k: st[k] / Q for k in ("n_steps", "n_linked", "n_active", "n_samples")}}
    if w["kind"] == "surf_trav":
        roof["kernels"] = {"forward_ms": fwd_ms, "backward_ms": bwd_ms,
                           "stages_ms": {"work_pyramid": stage_ms[0], "premarch": stage_ms[1], "forward_shading": stage_ms[2],
                                         "backward": stage_ms[3]},
                           "forward_bytes": fwd_bytes, "backward_bytes": bwd_bytes}
    if call_table is not None:
        roof["step_calls"] = annotate_calls(call_table, sg, Q, D, st, alg, own_bytes, traffic, peak)
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": 3 * Q * 3 * 4, "d2h_bytes_per_step": Q * 3 * 4,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "alphasurf_b200/csrc/svox2_csrc_shim*.so (compiled pybind11/torch module replacing svox2.csrc)"},
        "gpu_launches": n_launch,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "host_enqueue_note": "CPU time rank 0 spends issuing one step.  On one GPU nothing in the step waits for the device, so a "
                             "value close to ms_per_step would mean a launch-bound step; the multi-GPU step waits once per step "
                             "for the previous step's row count (GradExchange._touched_rows), so there the value tracks the "
                             "device time by construction",
        "roofline": roof,
        "step_phases": phases,
        "step_phases_note": ("sequential: fused render | regularisers | optimizer" if world == 1 else
                             "two lanes (dist.GradExchange.step): render_ms = render + mask OR + row pack + sparse row all-reduce + "
                             "unpack on the main stream while the cell-sharded surface regularisers and their dense all-reduce run "
                             "on a side stream; regularisers_ms = wait for that lane + join; optimizer_ms = RMSprop steps "
                             "(ASURF_MERGED_EXCHANGE=1: one gradient and one mask all-reduce instead, then render_ms holds the whole "
                             "exchange)"),
        "render_only_rays_per_s": Q * world / (kernel_ms * 1e-3) if kernel_ms > 0 else None,
    }
    if parity is not None:
        line["parity"] = parity
    if breakdown is not None:
        line["exchange"] = breakdown
    if world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_baseline(args, sg, host_batches[0])
        line["cpu_reference_l0"] = cpu_reference_l0(args)
        ref = reference_cuda_same_gpu(args, sg, dev_batches, w["kind"])
        if ref is not None:
            line["reference_cuda_same_gpu"] = ref
    print(json.dumps(line), flush=True)


def time_calls(torch, ts, kind, dev_batches, rgb_out, args):
    """CUDA-event time of every svox2.csrc call of the step, averaged over a few steps run after the timed region."""
    C, sg, hp = ts.C, ts.sg, ts.hp
    if kind == "surf_trav":
        ne = ts.non_empty
        calls = [
            ("volume_render_surf_trav_fused", lambda b: ts.render(b[0], b[1], b[2], rgb_out)),
            ("tv_grad_sparse(density, 1% window)", lambda b: C.tv_grad_sparse(
                sg.links, sg.density, ts.rand_cells(hp["tv_sparsity"]), ts.mask, 0, 1, hp["lambda_tv_alpha"], False, 2.0, False,
                False, -1.0, -1.0, ts.grad["density"])),
            ("surf_tv_grad_sparse(all stored cells)", lambda b: C.surf_tv_grad_sparse(
                sg.links, sg.surface, sg.density, ne, ts.mask, 0, 1, hp["lambda_tv_surface"], True, -1.0, False, -1.0, -1.0, False,
                ts.grad["surface"])),
            ("surface_normal_grad_sparse(all stored cells)", lambda b: C.surface_normal_grad_sparse(
                sg.links, sg.surface, ne, ts.mask, 0.0, 0, 1, hp["lambda_normal_loss"], 0.0, -1.0, -1.0, False, False, True,
                ts.grad["surface"])),
            ("alpha_surf_sparsify_grad_sparse(10% window)", lambda b: C.alpha_surf_sparsify_grad_sparse(
                sg.links, sg.density, sg.surface, ts.rand_cells_non_empty(hp["alpha_surf_sparsify_sparsity"]), ts.mask,
                hp["lambda_sparsify_alpha"], hp["lambda_sparsify_surf"], True, 0.15, 0.0, -0.1, ts.grad["density"],
                ts.grad["surface"])),
            ("rmsprop_step(density)", lambda b: C.rmsprop_step(sg.density, ts.rms["density"], ts.grad["density"], ts.mask, 0.95,
                                                               hp["lr_density"], 1e-8, -1e9, hp["lr_density"])),
            ("rmsprop_step(surface)", lambda b: C.rmsprop_step(sg.surface, ts.rms["surface"], ts.grad["surface"], ts.mask, 0.95,
                                                               hp["lr_surface"], 1e-8, -1e9, hp["lr_surface"])),
            ("rmsprop_step(sh)", lambda b: C.rmsprop_step(sg.sh, ts.rms["sh"], ts.grad["sh"], ts.mask_sh, 0.95, hp["lr_sh"], 1e-8,
                                                          -1e9, hp["lr_sh"])),
        ]
    else:
        calls = [
            ("volume_render_cuvol_fused", lambda b: ts.render(b[0], b[1], b[2], rgb_out)),
            ("tv_grad_sparse(sigma, 1% window) + tv_grad_sparse(sh, 1% window)", lambda b: ts.regularisers()),
            ("rmsprop_step(sigma) + rmsprop_step(sh)", lambda b: ts.optimizer()),
        ]
    n = 7
    samples = [[] for _ in calls]
    for it in range(n + 1):
        b = dev_batches[it % len(dev_batches)]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for k, (_, fn) in enumerate(calls):
            fn(b)
            evs[k + 1].record()
        torch.cuda.synchronize()
        if it > 0:
            for k in range(len(calls)):
                samples[k].append(evs[k].elapsed_time(evs[k + 1]))
    # median over the repeats: an event interval also contains any wait of the GPU for the host between two calls
    return [(name, sorted(s)[len(s) // 2], s) for (name, _), s in zip(calls, samples)]


def annotate_calls(table, sg, Q, D, st, alg_fused, own_fused, traffic, peak):
    """Per call: time, share of the step, algorithmic bytes (SURVEY 8d formulas) and, where an ncu capture exists
    (profiles/traffic.json), measured DRAM bytes -> both fractions of the HBM peak.  Sorted by time: dominant call first."""
    N = sg.capacity
    n_vert = sg.links.numel()
    total = sum(t[1] for t in table)
    out = []
    touched = st["n_samples"] * 8          # upper bound of the rows the render touches
    for name, ms, raw in table:
        alg, key, note = None, None, None
        if name.startswith("volume_render"):
            alg, key = alg_fused, "fused"
            note = "reference-algorithm bytes (8d); own-algorithm bytes: %d" % own_fused
        elif name.startswith("surf_tv_grad_sparse"):
            alg, key = N * 68, "surf_tv"                       # 4 links + 4 values + 4 RMW + 4 mask bytes per cell (8d)
        elif name.startswith("surface_normal_grad_sparse"):
            alg, key = N * 17, "normal_loss"
            note = "one pass over the stored vertices: link 4 B + value 4 B + gradient RMW 8 B + mask 1 B; issue-bound, not HBM-bound"
        elif name.startswith("tv_grad_sparse(density"):
            alg = int(0.01 * n_vert) * 68
        elif name.startswith("alpha_surf_sparsify"):
            alg = int(0.1 * N) * (4 + 4 + 4 + 4 + 8 + 1)
        elif name.startswith("rmsprop_step(sh)"):
            alg, key = N * 1 + min(touched, N) * D * 24, "rmsprop_sh"
        elif name.startswith("rmsprop_step("):
            alg, key = N * (24 + 1), "rmsprop_col1"
        e = {"call": name, "ms": ms, "share": ms / total if total > 0 else None}
        if max(raw) > 2.0 * max(min(raw), 1e-6):
            e["ms_samples"] = [round(v, 4) for v in raw]     # an unsteady interval: shown in full
        if alg is not None and ms > 0:
            e["algorithmic_bytes"] = alg
            e["achieved_GBps"] = alg / (ms * 1e-3) / 1e9
            e["frac"] = e["achieved_GBps"] / peak
        tr = traffic.get(key) if key else None
        if tr and ms > 0:
            e["traffic"] = tr
            e["dram_frac"] = tr / (ms * 1e-3) / 1e9 / peak
        if note:
            e["note"] = note
        out.append(e)
    out.sort(key=lambda e: -e["ms"])
    return out


def dist_parity(torch, dist, C, S, synth, sg, rank, world, Q, dev):
    """Outside the timed region: one N-rank step (every rank its own Q rays) against the 1-rank step on the concatenated
    batch, from identical grid copies -- touched masks must be equal, gradients and rendered colours within 1e-4."""
    from alphasurf_b200 import dist as adist

    def clone(s):
        return synth.SynthGrid(s.links, s.density.clone(), s.surface.clone(), s.sh.clone(), s.level_set, s.offset, s.scaling,
                               s.basis_dim, s.fake_sample_std, s.truncated_vol_render_a, dict(s.meta))
    try:
        base = synth.make_shell_grid(sg.links.shape[0], basis_dim=9, variant=sg.meta.get("variant", "G")).to(dev)
        rays = [synth.make_camera_rays(Q, device=dev, seed=4242 + r) for r in range(world)]
        # N-rank step (gradients after both exchanges, before the optimizer)
        a = S.TrainStep(C, clone(base), seed=11)
        C.set_loss_norm_rays(Q * world)
        ex = adist.GradExchange(a)
        out_a = torch.zeros((Q, 3), device=dev)
        ex.step(a, *rays[rank], out_a, skip_optimizer=True)
        torch.cuda.synchronize()
        # 1-rank step on the concatenated batch (every rank computes it redundantly)
        b = S.TrainStep(C, clone(base), seed=11)
        C.set_loss_norm_rays(None)
        o = torch.cat([r[0] for r in rays])
        d = torch.cat([r[1] for r in rays])
        gt = torch.cat([r[2] for r in rays])
        out_b = torch.zeros((Q * world, 3), device=dev)
        b.render(o, d, gt, out_b)
        b.regularisers()
        torch.cuda.synchronize()
        C.set_loss_norm_rays(Q * world)

        def rel(x, y):
            return float((x.double() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-30))
        res = {"what": "%d-rank step vs 1-rank step on the concatenated %d-ray batch, same grid" % (world, Q * world),
               "mask_equal": bool(torch.equal(a.mask, b.mask)), "mask_sh_equal": bool(torch.equal(a.mask_sh, b.mask_sh)),
               "rgb_rel_err": rel(out_a, out_b[rank * Q:(rank + 1) * Q]),
               "grad_rel_err": {k: rel(a.grad[k], b.grad[k]) for k in ("density", "surface", "sh")}}
        ok = torch.tensor([1 if (res["mask_equal"] and res["mask_sh_equal"] and res["rgb_rel_err"] < 1e-4 and
                                 max(res["grad_rel_err"].values()) < 1e-4) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["all_ranks_ok"] = bool(int(ok.item()))
        return res
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:400]}


def reference_cuda_same_gpu(args, sg, dev_batches, kind):
    """Extra (not part of the contract): the UNMODIFIED reference CUDA kernels (oracle/_ref) on the same GPU, driven
    through the same step sequence (its own copy of the grid)."""
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
    sg2 = synth.SynthGrid(sg.links, sg.density.clone(), None if sg.surface is None else sg.surface.clone(), sg.sh.clone(),
                          sg.level_set, sg.offset, sg.scaling, sg.basis_dim, sg.fake_sample_std, sg.truncated_vol_render_a,
                          dict(sg.meta))
    ts = S.CuvolStep(ref, sg2) if kind == "cuvol" else S.TrainStep(ref, sg2)
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
    return {"what": "reference svox2.csrc kernels, same step sequence (fused render, regularisers, RMSprop)",
            "render_ms": t_render, "regularisers_optimizer_ms": t_rest, "ms_per_step": t_render + t_rest,
            "rays_per_s": Q / ((t_render + t_rest) * 1e-3), "render_only_rays_per_s": Q / (t_render * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default=None, help="alias: --variant Gstar == --workload c3-gstar")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--reso", type=int, default=None)
    ap.add_argument("--rays", type=int, default=None)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--ref-rays", type=int, default=8192)
    ap.add_argument("--l0-rays", type=int, default=256)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--shard-regularisers", type=int, default=1,
                    help="multi-GPU: 1 = every rank takes 1/N of the regulariser cell lists and the shards are all-reduced (dense), "
                         "0 = every rank runs the whole lists redundantly (no regulariser exchange)")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline, the L0 baseline, the reference-CUDA comparison and "
                                                             "the N-rank parity block (profiling runs)")
    args = ap.parse_args()
    if args.variant and args.variant.lower().replace("*", "star") in ("gstar", "g-star"):
        args.workload = "c3-gstar"
    w = WORKLOADS[args.workload]
    args.reso = args.reso or w["reso"]
    args.rays = args.rays or w["rays"]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # The CPU legs (reference arm; cpu_baseline at N = 1) use every host thread.  torch.distributed.run exports
    # OMP_NUM_THREADS=1 to its workers: set it explicitly, BEFORE torch / libgomp load and read it.
    if args.impl == "reference" or world == 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
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
