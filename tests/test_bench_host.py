"""CPU: the host-side bookkeeping of bench.py -- byte formulas of SURVEY.md 8d, the per-call table, the traffic file -- so
that a slip in the measurement code shows here and not on the GPU box at round end.  No CUDA, no oracle."""
import argparse
import json
import os
import types

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_surf_trav_byte_formulas():
    """B_f / B_b of SURVEY.md 8d on one hand-computed ray population"""
    Q, D = 100, 27
    st = dict(n_steps=10.0 * Q, n_linked=4.0 * Q, n_active=3.0 * Q, n_samples=2.0 * Q)      # counters are batch totals
    fwd, bwd = bench.surf_trav_bytes(st, Q, D)
    per_f = 16 * 10 + 16 * 4 + 16 * 3 + 2 * 32 * (1 + D) + 36 + 12 * 2
    per_b = 16 * 10 + 16 * 4 + 16 * 3 + 2 * 32 * (1 + D) + 2 * (2 * 32 * D + 2 * 32 * 2 + 8) + 48 + 12 * 2
    assert fwd == pytest.approx(per_f * Q) and bwd == pytest.approx(per_b * Q)
    own = bench.surf_trav_own_bytes(st, Q, D, n_rows=1000)
    assert 0 < own < fwd + bwd + 8 * 1000 + 1       # never more than the reference algorithm's bytes plus the class scan


def test_traffic_file_covers_the_bench_workloads():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for w in bench.WORKLOADS:
        assert bench.load_traffic(w).get("fused", 0) > 0, w
        assert "source" in t[w] and os.path.exists(os.path.join(ROOT, t[w]["source"].split(":")[0])), w
    assert bench.load_traffic("no-such-workload") == {}


def test_workload_config_names_the_workload():
    for name, w in bench.WORKLOADS.items():
        args = argparse.Namespace(workload=name, reso=w["reso"], rays=w["rays"], scaling="weak", shard_regularisers=1)
        c = bench.workload_config(args, 1)
        assert c["rays_per_step_per_gpu"] == w["rays"] and str(w["reso"]) in c["grid"] and "model" not in c
        c8 = bench.workload_config(argparse.Namespace(**{**vars(args), "scaling": "strong"}), 8)
        assert c8["rays_per_step_global"] == w["rays"] // 8 * 8 and c8["scaling"] == "strong"


def test_annotate_calls_table():
    sg = types.SimpleNamespace(capacity=1000, links=types.SimpleNamespace(numel=lambda: 8000))
    st = dict(n_steps=10.0, n_linked=4.0, n_active=3.0, n_samples=0.5)
    table = [("volume_render_surf_trav_fused", 0.4, [0.4] * 7), ("surf_tv_grad_sparse(all stored cells)", 0.2, [0.2] * 7),
             ("surface_normal_grad_sparse(all stored cells)", 0.6, [0.6, 0.6, 5.0, 0.6, 0.6, 0.6, 0.6]),
             ("rmsprop_step(sh)", 0.1, [0.1] * 7), ("rmsprop_step(density)", 0.05, [0.05] * 7)]
    out = bench.annotate_calls(table, sg, 64, 27, st, alg_fused=10_000_000, own_fused=4_000_000,
                               traffic={"fused": 2_000_000, "normal_loss": 1_000_000}, peak=6542.7)
    assert [e["call"] for e in out][0].startswith("surface_normal")            # dominant call first
    assert sum(e["share"] for e in out) == pytest.approx(1.0)
    fused = next(e for e in out if e["call"].startswith("volume_render"))
    assert fused["frac"] == pytest.approx(10_000_000 / 0.4e-3 / 1e9 / 6542.7)
    assert fused["dram_frac"] == pytest.approx(2_000_000 / 0.4e-3 / 1e9 / 6542.7)
    normal = out[0]
    assert "ms_samples" in normal and "dram_frac" in normal                    # the outlier sample is shown, not hidden
    assert "traffic" not in next(e for e in out if e["call"].startswith("surf_tv"))
