"""Ray-sharded data parallelism for the render hot path.

The grid (links, density, surface, SH, RMS state) is replicated on every GPU; each rank renders its slice of the global
ray batch into its local gradient buffers with the fused losses normalised by the GLOBAL ray count
(``svox2_csrc.set_loss_norm_rays``); the gradients and touched-row masks are then summed / OR-ed across ranks, and the
grid regularisers + RMSprop steps run redundantly (same cell lists on every rank: same seed), so no parameter
broadcast is needed.  The reference has no multi-GPU path; the contract is "equal to the single-GPU result on the
concatenated batch" up to fp32 summation order.

The exchange is the one collective step of the path.  A batch touches ~1% of the 15 M voxel rows, so the dense
gradients (1.8 GB at 512^3 x 29 floats) are never sent: the masks are OR-ed first (one byte per row), which gives every
rank the same list of touched rows, and only those rows travel -- packed into one (n_rows, 2 + D) bucket, summed with a
single NCCL all-reduce over NVLink / NVSwitch, and scattered back.
"""
import os
import time

import torch
import torch.distributed as dist


class GradExchange:
    """Sum density / surface / SH gradients of the touched rows and OR the touched masks over all ranks.

    Two-phase use (what TrainStep / bench.py do): ``begin`` right after the fused render -- OR the masks, pack the touched
    rows into one bucket, clear them in the gradient tensors and start the all-reduce asynchronously -- then the grid
    regularisers run (they add identical values on every rank into the cleared rows, so they must not travel), and ``end``
    before the optimizer waits for the collective and adds the reduced render gradients back.  The NCCL transfer overlaps
    the regulariser kernels.  ``run`` does both phases back to back.
    """

    def __init__(self, ts=None, group=None, dense_threshold=0.25, shard_regularisers=True, sync_free=True):
        self.group = group
        # sync_free: after the first step the list of touched rows is taken with a fixed capacity (1.5 x the largest count
        # seen so far) instead of a data-dependent size, so the host never waits for the device inside the step; the true
        # count is read back one step later and an overflow raises (it cannot be repaired after the optimizer has run).
        self.sync_free = sync_free
        self.cap = None
        self._count_probe = None     # (pinned host count, event, capacity it was taken with)
        self._mask_bufs = {}
        self._probe_bufs = None
        self.host_ms = {}            # host time per section of step() (enqueue cost, no synchronisation), summed
        self.host_steps = 0
        self.bitpack_masks = False
        # step() with ONE gradient and ONE mask all-reduce (_step_merged).  Measured (C3 step): 1.780 ms against 1.718 ms for
        # the two lanes at 2 GPUs, 1.991 against 2.002 ms at 8 GPUs -- no gain, so it is off unless asked for.
        self.merged_exchange = os.environ.get("ASURF_MERGED_EXCHANGE", "0") != "0"
        self._merged = None
        self._masks2 = None
        # order of the two lanes of step(): "render_first" / "regs_first" -- both lanes start together in that enqueue order
        # (their kernels share the SMs); "regs_kernels_first" -- the regulariser kernels run first and alone, the render starts
        # when they are done and overlaps their collectives.  Measured at 2 GPUs (C3 step): 1.717 / 1.699 / 1.789 ms.
        self.lane_order = os.environ.get("ASURF_LANE_ORDER", "render_first")
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_regularisers = shard_regularisers   # end() also sums the cell-sharded regulariser gradients
        self.dense_threshold = dense_threshold   # above this touched fraction the dense all-reduce is cheaper
        self.bucket = None
        self.last_rows = 0
        self._pending = None
        self._hold = None
        self._reg = None
        self._group_b = None
        self._side = None
        if ts is not None and hasattr(ts, "sg"):
            self._lane_b(ts)     # buffers, side stream and second communicator up front (collective: every rank is here)

    def _bucket(self, n, width, like):
        if self.bucket is None or self.bucket.shape[0] < n or self.bucket.shape[1] != width or self.bucket.device != like.device:
            cap = max(int(n * 1.5), 1024)
            self.bucket = torch.empty((cap, width), dtype=like.dtype, device=like.device)
        return self.bucket[:n]

    def mask_or(self, mask, group=None, bitpack=None):
        """mask (N,) bool <- OR over the ranks: a MAX all-reduce of the N bytes, or (bitpack) asurf_mask_pack -> all-gather
        of N / 8 bytes per rank -> asurf_mask_unpack_or.  Measured on B200 / NVSwitch (profiles/r2_scale_*): the 15 MB
        byte-wise all-reduce takes 0.064 ms at 2 ranks and 0.090 ms at 8, the bit-packed form 0.086 ms at 2 ranks (two extra
        launches and the all-gather's latency outweigh the 8x smaller message), so the byte-wise form is the default."""
        group = self.group if group is None else group
        bitpack = self.bitpack_masks if bitpack is None else bitpack
        if not mask.is_cuda or not bitpack:
            dist.all_reduce(mask.view(torch.uint8), op=dist.ReduceOp.MAX, group=group)
            return
        from . import capi
        import ctypes as C
        N = mask.shape[0]
        nw = (N + 31) // 32
        key = (N, mask.device, torch.cuda.current_stream(mask.device).cuda_stream)
        buf = self._mask_bufs.get(key)
        if buf is None:
            buf = (torch.empty((nw,), dtype=torch.int32, device=mask.device),
                   torch.empty((self.world * nw,), dtype=torch.int32, device=mask.device))
            self._mask_bufs[key] = buf
        words, gathered = buf
        st = capi.current_stream(mask.device)
        capi.check(capi.lib().asurf_mask_pack(capi.ptr(mask), C.c_int64(N), capi.ptr(words), st), "mask_pack")
        dist.all_gather_into_tensor(gathered, words, group=group)
        capi.check(capi.lib().asurf_mask_unpack_or(capi.ptr(gathered), C.c_int32(self.world), C.c_int64(N), capi.ptr(mask), st),
                   "mask_unpack_or")

    def begin(self, ts):
        """``ts``: object with ``grad`` = {density (N,1), surface (N,1), sh (N,D)}, ``mask`` and ``mask_sh`` (N,) bool."""
        g = ts.grad
        self.mask_or(ts.mask)
        ts.mask_sh.copy_(ts.mask)
        N = ts.mask.shape[0]
        rows, n = self._touched_rows(ts.mask)             # identical on every rank
        self.last_rows = n
        if n == 0:
            self._pending = None
            return 0
        if n > self.dense_threshold * N:
            for k in ("density", "surface", "sh"):
                dist.all_reduce(g[k], op=dist.ReduceOp.SUM, group=self.group)
            self._pending = None
            if self.shard_regularisers and self.world > 1:   # keep the summed render part out of end()'s dense sum
                self._hold = (g["density"].clone(), g["surface"].clone())
                g["density"].zero_()
                g["surface"].zero_()
            return n
        D = g["sh"].shape[1]
        buf = self._bucket(n, 2 + D, g["sh"])
        self._pack(g, rows, n, buf)
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending = (work, rows, buf)
        return n

    def _touched_rows(self, mask):
        """-> (rows int64 (n,), n).  First call (and host tensors): torch.nonzero, which waits for the count.  Afterwards, on
        CUDA with sync_free: a list of fixed capacity padded with -1 (the pack / unpack kernels skip negative rows)."""
        if not (self.sync_free and mask.is_cuda) or self.cap is None:
            rows = torch.nonzero(mask).flatten()
            n = int(rows.shape[0])
            if self.sync_free and mask.is_cuda:
                self.cap = min(int(mask.shape[0]), int(1.5 * n) + 1024)
            return rows, n
        if self._count_probe is not None:                 # the count of the PREVIOUS step: long since on the host
            host, ev, cap_used = self._count_probe
            ev.synchronize()
            cnt = int(host[0])
            if cnt > cap_used:
                raise RuntimeError("GradExchange: %d touched rows exceeded the list capacity %d of the previous step; "
                                   "its gradient exchange was incomplete (use sync_free=False)" % (cnt, cap_used))
            if cnt > 0.8 * self.cap:
                self.cap = min(int(mask.shape[0]), int(1.5 * cnt) + 1024)
        cap = self.cap
        rows = torch.nonzero_static(mask, size=cap, fill_value=-1).flatten()
        if self._probe_bufs is None:      # two pinned words + two events, reused alternately (allocating pinned memory is slow)
            self._probe_bufs = [(torch.zeros((1,), dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(2)]
            self._probe_i = 0
        host, ev = self._probe_bufs[self._probe_i]
        self._probe_i ^= 1
        host.copy_(mask.sum().reshape(1), non_blocking=True)
        ev.record()
        self._count_probe = (host, ev, cap)
        return rows, cap

    def collective_breakdown(self, ts, iters=5):
        """Each collective / kernel of one exchange timed ALONE (CUDA events on this stream, synchronous NCCL calls), outside
        any timed region: names the limiter of the multi-GPU step.  Sizes are those of the last step."""
        g, dev = ts.grad, ts.mask.device
        N, D = ts.mask.shape[0], g["sh"].shape[1]
        n = max(int(self.last_rows), 1)
        out = {"rows_in_bucket": n, "world": self.world}
        mask_u8 = torch.zeros((N,), dtype=torch.uint8, device=dev)
        bucket = torch.zeros((n, 2 + D), dtype=g["sh"].dtype, device=dev)
        dense = torch.zeros((2, N, 1), dtype=g["sh"].dtype, device=dev)

        def t(fn):
            fn()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / iters
        out["mask_or_allreduce_bytewise_ms"] = t(lambda: dist.all_reduce(mask_u8, op=dist.ReduceOp.MAX, group=self.group))
        mb = torch.zeros((N,), dtype=torch.bool, device=dev)
        out["mask_or_bitpacked_allgather_ms"] = t(lambda: self.mask_or(mb, bitpack=True))
        out["mask_or_bytes_per_rank"] = (N + 31) // 32 * 4
        out["sparse_bucket_allreduce_ms"] = t(lambda: dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group))
        out["sparse_bucket_bytes"] = n * (2 + D) * 4
        rep = hasattr(ts, "density_terms_replicable") and ts.density_terms_replicable()
        dense_part = dense[1] if rep else dense
        out["dense_regulariser_allreduce_ms"] = t(lambda: dist.all_reduce(dense_part, op=dist.ReduceOp.SUM, group=self.group))
        out["dense_regulariser_bytes"] = dense_part.numel() * 4
        out["nonzero_static_ms"] = t(lambda: torch.nonzero_static(ts.mask, size=n, fill_value=-1))
        return out

    def end(self, ts):
        g = ts.grad
        if self.shard_regularisers and self.world > 1:
            # the regularisers ran on this rank's share of the cells: sum the shards (they touch every stored row, so
            # this exchange is dense) and OR the masks they set
            self.mask_or(ts.mask)
            dist.all_reduce(g["density"], op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(g["surface"], op=dist.ReduceOp.SUM, group=self.group)
            if self._hold is not None:
                g["density"].add_(self._hold[0])
                g["surface"].add_(self._hold[1])
                self._hold = None
        if self._pending is None:
            return
        work, rows, buf = self._pending
        self._pending = None
        work.wait()
        self._unpack(g, rows, buf)

    @staticmethod
    def _pack(g, rows, n, buf):
        """touched rows of the three gradients -> bucket (n, 2 + D), rows cleared; negative (padding) rows give zero rows"""
        D = g["sh"].shape[1]
        if buf.is_cuda:    # one kernel (asurf_rows_pack)
            from . import capi
            import ctypes as C
            capi.check(capi.lib().asurf_rows_pack(capi.ptr(rows), C.c_int64(n), capi.ptr(g["density"]), capi.ptr(g["surface"]),
                                                  capi.ptr(g["sh"]), C.c_int32(D), capi.ptr(buf), C.c_int32(1),
                                                  capi.current_stream(buf.device)), "rows_pack")
        else:              # host tensors (gloo tests of the protocol)
            buf[:, 0] = g["density"].view(-1)[rows]
            buf[:, 1] = g["surface"].view(-1)[rows]
            buf[:, 2:] = g["sh"][rows]
            g["density"].view(-1).index_fill_(0, rows, 0.0)
            g["surface"].view(-1).index_fill_(0, rows, 0.0)
            g["sh"].index_fill_(0, rows, 0.0)

    @staticmethod
    def _unpack(g, rows, buf):
        if buf.is_cuda:
            from . import capi
            import ctypes as C
            capi.check(capi.lib().asurf_rows_unpack_add(capi.ptr(rows), C.c_int64(rows.shape[0]), capi.ptr(g["density"]),
                                                        capi.ptr(g["surface"]), capi.ptr(g["sh"]), C.c_int32(g["sh"].shape[1]),
                                                        capi.ptr(buf), capi.current_stream(buf.device)), "rows_unpack_add")
        else:
            g["density"].view(-1).index_add_(0, rows, buf[:, 0])
            g["surface"].view(-1).index_add_(0, rows, buf[:, 1])
            g["sh"].index_add_(0, rows, buf[:, 2:])

    def run(self, ts):
        n = self.begin(ts)
        self.end(ts)
        return n

    # ---- the whole multi-GPU iteration ---------------------------------------------------------------------------------
    def _lane_b(self, ts):
        """buffers and communicator of the regulariser lane (created on first use; every rank does so at the same point)"""
        if self._reg is None:
            N = ts.grad["density"].shape[0]
            dev = ts.grad["density"].device
            buf = torch.zeros((2, N, 1), dtype=ts.grad["density"].dtype, device=dev)    # one bucket: density, surface
            self._reg = dict(buf=buf, grad={"density": buf[0], "surface": buf[1]},
                             mask=torch.zeros((N,), dtype=torch.bool, device=dev))
            # a communicator of its own, so that the dense exchange of this lane and the sparse one of the render lane
            # are not serialised on one NCCL stream
            self._group_b = dist.new_group(ranks=list(range(self.world))) if self.group is None else self.group
            self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
            if self._side is not None:
                # build the cached occupancy pyramid of `links` now, on the main stream: afterwards both lanes only read it
                ts.C.accel_for(ts.sg.links)
                torch.cuda.synchronize(dev)
        return self._reg

    def step(self, ts, origins, dirs, rgb_gt, rgb_out, events=None, skip_optimizer=False):
        """One training iteration on this rank's rays.  The regularisers depend on the parameters only, not on the render:
        they run cell-sharded on a side stream into buffers of their own and their dense all-reduce (2 x N floats + N mask
        bytes) proceeds on a second communicator WHILE the main stream renders, ORs the touched masks and exchanges the
        touched rows (begin / end above with shard_regularisers off).  Both lanes join before the optimizer.
        ``events``: optional 4 CUDA events recorded at start / after render + begin / after the join / after the optimizer."""
        if (self.merged_exchange and self.shard_regularisers and self.world > 1 and hasattr(ts, "density_terms_replicable")
                and ts.density_terms_replicable()):
            return self._step_merged(ts, origins, dirs, rgb_gt, rgb_out, events, skip_optimizer)
        reg = self._lane_b(ts)
        cuda = self._side is not None
        t_h = [time.perf_counter()]

        def lap(name):
            now = time.perf_counter()
            self.host_ms[name] = self.host_ms.get(name, 0.0) + 1e3 * (now - t_h[0])
            t_h[0] = now
        if events:
            events[0].record()
        import contextlib
        ev_start = None
        if cuda:
            ev_start = torch.cuda.Event()
            ev_start.record()                     # the previous optimizer step has updated the parameters

        def side_lane():
            if cuda:
                self._side.wait_event(ev_start)
                ctx = torch.cuda.stream(self._side)
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                reg["buf"].zero_()
                reg["mask"].zero_()
                if self.shard_regularisers:
                    # the expensive terms (surface TV, normal loss: both write the SURFACE gradient) are sharded over the
                    # ranks by cell; the cheap ones that write the density gradient run in full on every rank when that is
                    # exact (TrainStep.density_terms_replicable), so that only the surface half of the buffer travels
                    rep = hasattr(ts, "density_terms_replicable") and ts.density_terms_replicable()
                    if rep:
                        ts.regularisers(self.rank, self.world, grad=reg["grad"], mask=reg["mask"], replicate_density_terms=True)
                    else:
                        ts.regularisers(self.rank, self.world, grad=reg["grad"], mask=reg["mask"])
                    if cuda:
                        ev_kernels.record()       # the regulariser kernels are done: what follows on this lane is communication
                    self.mask_or(reg["mask"], group=self._group_b)
                    dist.all_reduce(reg["buf"][1] if rep else reg["buf"], op=dist.ReduceOp.SUM, group=self._group_b)
                else:
                    ts.regularisers(grad=reg["grad"], mask=reg["mask"])     # every rank the whole lists: nothing to exchange

        # Enqueue order (self.lane_order), the same on every rank.  Both lanes are compute-bound in their kernels and wire-bound in
        # their collectives, and concurrent kernels only share the SMs; what can be hidden is communication under computation.
        render_first = self.lane_order == "render_first"
        ev_kernels = torch.cuda.Event() if cuda else None
        if not render_first:
            side_lane()
            if cuda and self.lane_order == "regs_kernels_first" and self.shard_regularisers:
                torch.cuda.current_stream().wait_event(ev_kernels)
            lap("regulariser lane: kernels + mask OR + dense all-reduce")
        shard = self.shard_regularisers
        self.shard_regularisers = False           # begin / end: the sparse exchange of the render gradients alone
        try:
            ts.render(origins, dirs, rgb_gt, rgb_out)
            lap("render")
            self.begin(ts)
            lap("begin: mask OR, row list, pack, sparse all-reduce start")
            if events:
                events[1].record()
            if render_first:
                self.shard_regularisers = shard
                side_lane()
                self.shard_regularisers = False
                lap("regulariser lane: kernels + mask OR + dense all-reduce")
            self.end(ts)
            lap("end: wait + unpack")
        finally:
            self.shard_regularisers = shard
        if cuda:
            torch.cuda.current_stream().wait_stream(self._side)
        ts.grad["density"].add_(reg["grad"]["density"])
        ts.grad["surface"].add_(reg["grad"]["surface"])
        ts.mask.logical_or_(reg["mask"])
        lap("join")
        if events:
            events[2].record()
        if not skip_optimizer:
            ts.optimizer()
        lap("optimizer")
        self.host_steps += 1
        if events:
            events[3].record()

    def _step_merged(self, ts, origins, dirs, rgb_gt, rgb_out, events=None, skip_optimizer=False):
        """The iteration with ONE gradient all-reduce and ONE mask all-reduce (optional: self.merged_exchange).  At 8 GPUs the
        four collectives of step() (two mask ORs, the sparse bucket, the dense regulariser gradient) add up to the time the
        step spends outside its kernels, so this variant sends them as two: the cell-sharded surface regularisers still run on
        the side stream beside the render, but write their gradient into the HEAD of the exchange buffer, the touched rows of
        the render gradients are packed behind it, and the whole buffer travels once; the render mask and the regulariser
        mask travel as one 2N-byte MAX all-reduce.  (The density regularisers run in full on every rank:
        TrainStep.density_terms_replicable.)  Same results as step(); not faster on NVSwitch (see merged_exchange above)."""
        reg = self._lane_b(ts)
        cuda = self._side is not None
        g = ts.grad
        N, D = ts.mask.shape[0], g["sh"].shape[1]
        dev, dt = g["sh"].device, g["sh"].dtype
        max_rows = int(self.dense_threshold * N) + 1          # beyond this the exchange is dense anyway
        if self._merged is None:
            self._merged = torch.zeros((N + max_rows * (2 + D),), dtype=dt, device=dev)
            self._masks2 = torch.zeros((2 * N,), dtype=torch.bool, device=dev)
        head = self._merged[:N].view(N, 1)                    # surface gradient of this rank's regulariser shard
        mask_reg = self._masks2[N:]
        t_h = [time.perf_counter()]

        def lap(name):
            now = time.perf_counter()
            self.host_ms[name] = self.host_ms.get(name, 0.0) + 1e3 * (now - t_h[0])
            t_h[0] = now
        if events:
            events[0].record()
        import contextlib
        ev_start = ev_side = None
        if cuda:
            ev_start = torch.cuda.Event()
            ev_start.record()                     # the previous optimizer step has updated the parameters
            self._side.wait_event(ev_start)
        with (torch.cuda.stream(self._side) if cuda else contextlib.nullcontext()):
            head.zero_()
            reg["grad"]["density"].zero_()
            mask_reg.zero_()
            ts.regularisers(self.rank, self.world, grad={"density": reg["grad"]["density"], "surface": head}, mask=mask_reg,
                            replicate_density_terms=True)
            if cuda:
                ev_side = torch.cuda.Event()
                ev_side.record()
        lap("regulariser lane: kernels")
        ts.render(origins, dirs, rgb_gt, rgb_out)
        lap("render")
        if cuda:
            torch.cuda.current_stream().wait_event(ev_side)
        self._masks2[:N].copy_(ts.mask)
        dist.all_reduce(self._masks2.view(torch.uint8), op=dist.ReduceOp.MAX, group=self.group)
        ts.mask_sh.copy_(self._masks2[:N])
        rows, n = self._touched_rows(self._masks2[:N])        # identical on every rank
        self.last_rows = n
        if n > max_rows:                                      # dense fallback: the three gradients whole, the head alone
            for k in ("density", "surface", "sh"):
                dist.all_reduce(g[k], op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self._merged[:N], op=dist.ReduceOp.SUM, group=self.group)
        else:
            bucket = self._merged[N:N + n * (2 + D)].view(n, 2 + D)
            if n:
                self._pack(g, rows, n, bucket)
            dist.all_reduce(self._merged[:N + n * (2 + D)], op=dist.ReduceOp.SUM, group=self.group)
            if n:
                self._unpack(g, rows, bucket)
        lap("exchange: masks, row list, pack, all-reduce, unpack")
        if events:
            events[1].record()
        g["density"].add_(reg["grad"]["density"])
        g["surface"].add_(head)
        torch.logical_or(self._masks2[:N], mask_reg, out=ts.mask)
        lap("join")
        if events:
            events[2].record()
        if not skip_optimizer:
            ts.optimizer()
        lap("optimizer")
        self.host_steps += 1
        if events:
            events[3].record()


def allreduce_grads(G, group=None):
    """Dense variant (kept for small grids / tests): G has .density .surface .sh .mask."""
    dist.all_reduce(G.mask.view(torch.uint8), op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(G.density, op=dist.ReduceOp.SUM, group=group)
    if G.surface is not None:
        dist.all_reduce(G.surface, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(G.sh, op=dist.ReduceOp.SUM, group=group)


def render_sharded(render_fn, origins, dirs, group=None, gather_to=None):
    """Evaluation render of a ray set split over the ranks (a full image, pixels sharded).

    ``render_fn(origins, dirs) -> (n, C)`` is any per-ray render of the csrc-compatible module bound to its grid and
    options (``volume_render_surf_trav``, a depth / normal render, ...); every rank passes the SAME full ``origins`` /
    ``dirs`` (Q, 3), renders the contiguous slice ``[rank * ceil(Q / world), ...)`` and the slices are concatenated in ray
    order -- on every rank (``gather_to=None``, all-gather) or on rank ``gather_to`` only (the others get ``None``).
    Rays are independent, so the result equals the single-process render bit for bit."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    Q = origins.shape[0]
    per = (Q + world - 1) // world
    lo, hi = min(rank * per, Q), min((rank + 1) * per, Q)
    part = render_fn(origins[lo:hi].contiguous(), dirs[lo:hi].contiguous())
    width = tuple(part.shape[1:])
    if part.shape[0] < per:        # equal-sized slices for the collective; the padding is cut off below
        pad = torch.zeros((per - part.shape[0],) + width, dtype=part.dtype, device=part.device)
        part = torch.cat([part, pad])
    part = part.contiguous()
    if gather_to is None:
        out = torch.empty((world * per,) + width, dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(out, part, group=group)
        return out[:Q]
    bufs = [torch.empty_like(part) for _ in range(world)] if rank == gather_to else None
    dist.gather(part, bufs, dst=gather_to, group=group)
    return torch.cat(bufs)[:Q] if rank == gather_to else None
