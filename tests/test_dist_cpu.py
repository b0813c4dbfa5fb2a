"""CPU, world_size 2, gloo: the gradient / mask exchange of the ray-sharded data-parallel path (alphasurf_b200.dist).
Two processes hold different gradients on overlapping touched rows; after the exchange both must hold the sum on the union
of the rows (what a single process rendering the concatenated batch would have accumulated) and identical masks."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphasurf_b200 import dist as adist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_state(rank, N, D, touched_frac):
    g = torch.Generator().manual_seed(100 + rank)
    mask = torch.rand((N,), generator=g) < touched_frac
    grad = {"density": torch.randn((N, 1), generator=g), "surface": torch.randn((N, 1), generator=g),
            "sh": torch.randn((N, D), generator=g)}
    for k in grad:
        grad[k][~mask] = 0.0     # a rank only has gradient on rows it touched
    return types.SimpleNamespace(grad=grad, mask=mask.clone(), mask_sh=mask.clone())


def _worker(rank, world, port, N, D, frac, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts = _local_state(rank, N, D, frac)
        ex = adist.GradExchange(shard_regularisers=False)   # protocol of the sparse render-gradient exchange alone
        n = ex.run(ts)
        # expected: sum over ranks, union of masks
        states = [_local_state(r, N, D, frac) for r in range(world)]
        want_mask = torch.stack([s.mask for s in states]).any(0)
        ok = torch.equal(ts.mask, want_mask) and torch.equal(ts.mask_sh, want_mask) and n == int(want_mask.sum())
        for k in ("density", "surface", "sh"):
            want = sum(s.grad[k] for s in states)
            ok = ok and torch.allclose(ts.grad[k], want, rtol=0, atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(N, D, frac):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, N, D, frac, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sparse_exchange_world2():
    _run(5000, 12, 0.02)


def test_dense_fallback_world2():
    _run(2000, 27, 0.6)


def test_empty_exchange_world2():
    _run(1000, 3, 0.0)


def _worker_sharded(rank, world, port, N, D, frac, ret):
    """full protocol: sparse exchange of the render gradients around cell-sharded regularisers"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def reg(r):   # what rank r's share of the regularisers adds: rows [r*N/world, (r+1)*N/world)
            g = torch.Generator().manual_seed(500 + r)
            lo, hi = (N * r) // world, (N * (r + 1)) // world
            dd, ds = torch.zeros((N, 1)), torch.zeros((N, 1))
            dd[lo:hi] = torch.randn((hi - lo, 1), generator=g)
            ds[lo:hi] = torch.randn((hi - lo, 1), generator=g)
            m = torch.zeros((N,), dtype=torch.bool)
            m[lo:hi] = True
            return dd, ds, m
        ts = _local_state(rank, N, D, frac)
        ex = adist.GradExchange()
        ex.begin(ts)
        dd, ds, m = reg(rank)
        ts.grad["density"] += dd
        ts.grad["surface"] += ds
        ts.mask |= m
        ex.end(ts)
        states = [_local_state(r, N, D, frac) for r in range(world)]
        regs = [reg(r) for r in range(world)]
        ok = torch.equal(ts.mask, torch.ones((N,), dtype=torch.bool))
        ok = ok and torch.equal(ts.mask_sh, torch.stack([s.mask for s in states]).any(0))
        ok = ok and torch.allclose(ts.grad["density"], sum(s.grad["density"] for s in states) + sum(x[0] for x in regs), atol=1e-6)
        ok = ok and torch.allclose(ts.grad["surface"], sum(s.grad["surface"] for s in states) + sum(x[1] for x in regs), atol=1e-6)
        ok = ok and torch.allclose(ts.grad["sh"], sum(s.grad["sh"] for s in states), atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run_sharded(N, D, frac):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker_sharded, args=(r, world, port, N, D, frac, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sharded_regularisers_sparse_world2():
    _run_sharded(4000, 12, 0.03)


def test_sharded_regularisers_dense_world2():
    _run_sharded(1500, 27, 0.7)


class _FakeStep:
    """stands in for TrainStep in the protocol test of GradExchange.step: render / regularisers / optimizer are tensor ops"""

    def __init__(self, rank, N, D, frac):
        self.rank, self.N, self.D, self.frac = rank, N, D, frac
        self.grad = {"density": torch.zeros((N, 1)), "surface": torch.zeros((N, 1)), "sh": torch.zeros((N, D))}
        self.mask = torch.zeros((N,), dtype=torch.bool)
        self.mask_sh = torch.zeros((N,), dtype=torch.bool)
        self.seen = None

    def render(self, o, d, gt, out):
        st = _local_state(self.rank, self.N, self.D, self.frac)
        self.mask.zero_()
        for k in self.grad:
            self.grad[k] += st.grad[k]
        self.mask |= st.mask
        self.mask_sh.copy_(self.mask)

    @staticmethod
    def reg_part(r, world, N):
        g = torch.Generator().manual_seed(900 + r)
        lo, hi = (N * r) // world, (N * (r + 1)) // world
        dd, ds = torch.zeros((N, 1)), torch.zeros((N, 1))
        dd[lo:hi] = torch.randn((hi - lo, 1), generator=g)
        ds[lo:hi] = torch.randn((hi - lo, 1), generator=g)
        m = torch.zeros((N,), dtype=torch.bool)
        m[lo:hi] = True
        return dd, ds, m

    def regularisers(self, rank, world, grad=None, mask=None):
        dd, ds, m = self.reg_part(rank, world, self.N)
        grad["density"] += dd
        grad["surface"] += ds
        mask |= m

    def optimizer(self):
        self.seen = {k: v.clone() for k, v in self.grad.items()}, self.mask.clone(), self.mask_sh.clone()
        for v in self.grad.values():
            v.zero_()


def _worker_step(rank, world, port, N, D, frac, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts = _FakeStep(rank, N, D, frac)
        ex = adist.GradExchange()
        ok = True
        for _ in range(2):       # two iterations: buffers of the regulariser lane are reused
            ex.step(ts, None, None, None, None)
            grads, mask, mask_sh = ts.seen
            states = [_local_state(r, N, D, frac) for r in range(world)]
            regs = [_FakeStep.reg_part(r, world, N) for r in range(world)]
            ok = ok and torch.equal(mask, torch.ones((N,), dtype=torch.bool))
            ok = ok and torch.equal(mask_sh, torch.stack([s.mask for s in states]).any(0))
            ok = ok and torch.allclose(grads["density"], sum(s.grad["density"] for s in states) + sum(x[0] for x in regs), atol=1e-6)
            ok = ok and torch.allclose(grads["surface"], sum(s.grad["surface"] for s in states) + sum(x[1] for x in regs), atol=1e-6)
            ok = ok and torch.allclose(grads["sh"], sum(s.grad["sh"] for s in states), atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_overlapped_step_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker_step, args=(r, world, port, 3000, 12, 0.03, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


class _FakeStepRep(_FakeStep):
    """density regularisers replicated (the same full gradient on every rank), surface regularisers sharded: the shape of
    TrainStep that takes GradExchange's merged exchange (one gradient all-reduce, one mask all-reduce)"""

    def density_terms_replicable(self):
        return True

    @staticmethod
    def dens_full(N):
        return torch.randn((N, 1), generator=torch.Generator().manual_seed(77))

    def regularisers(self, rank, world, grad=None, mask=None, replicate_density_terms=False):
        assert replicate_density_terms
        _, ds, m = self.reg_part(rank, world, self.N)
        grad["density"] += self.dens_full(self.N)
        grad["surface"] += ds
        mask |= m


def _worker_step_merged(rank, world, port, N, D, frac, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts = _FakeStepRep(rank, N, D, frac)
        ex = adist.GradExchange()
        ex.merged_exchange = True
        ok = True
        for _ in range(2):
            ex.step(ts, None, None, None, None)
            grads, mask, mask_sh = ts.seen
            states = [_local_state(r, N, D, frac) for r in range(world)]
            regs = [_FakeStep.reg_part(r, world, N) for r in range(world)]
            ok = ok and ex._merged is not None                      # the merged path ran
            ok = ok and torch.equal(mask, torch.ones((N,), dtype=torch.bool))
            ok = ok and torch.equal(mask_sh, torch.stack([s.mask for s in states]).any(0))
            ok = ok and torch.allclose(grads["density"], sum(s.grad["density"] for s in states) + _FakeStepRep.dens_full(N), atol=1e-6)
            ok = ok and torch.allclose(grads["surface"], sum(s.grad["surface"] for s in states) + sum(x[1] for x in regs), atol=1e-6)
            ok = ok and torch.allclose(grads["sh"], sum(s.grad["sh"] for s in states), atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frac", [0.03, 0.6])     # sparse bucket behind the regulariser head / dense fallback
def test_merged_exchange_step_world2(frac):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker_step_merged, args=(r, world, port, 3000, 12, frac, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _worker_render(rank, world, port, Q, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        o, d = torch.randn((Q, 3), generator=g), torch.randn((Q, 3), generator=g)
        fn = lambda a, b: torch.cat([a * 2.0 + b, (a * b).sum(1, keepdim=True)], dim=1)     # any per-ray function
        want = fn(o, d)
        full = adist.render_sharded(fn, o, d)
        on0 = adist.render_sharded(fn, o, d, gather_to=0)
        depth = adist.render_sharded(lambda a, b: (a * b).sum(1), o, d)                      # 1-D results (depth maps)
        ok = torch.equal(full, want) and torch.equal(depth, (o * d).sum(1))
        ok = ok and ((on0 is None) if rank != 0 else torch.equal(on0, want))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_render_sharded_world2():
    for Q in (1001, 4):          # ragged split, and fewer rays than a full slice on the last rank
        world = 2
        mgr = mp.Manager()
        ret = mgr.dict()
        port = _free_port()
        procs = [mp.get_context("spawn").Process(target=_worker_render, args=(r, world, port, Q, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert all(ret.get(r) for r in range(world)), (Q, dict(ret))
