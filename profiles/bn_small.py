"""Fused BatchNorm kernels at strong-scaling sizes (128 graphs per GPU: 28 792 tuples, 2 990 nodes,
128 graphs): per-kernel time in CUDA-graph replay against the CTAs-per-SM knob (key 2).

    python profiles/bn_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygho_b200  # noqa: E402,F401
from pygho_b200 import _lib, ops  # noqa: E402,F401

O = torch.ops.pygho_b200
dev = torch.device("cuda", 0)


def graph_time(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


for rows, C in ((128, 128), (2990, 128), (28792, 128), (28792, 384), (230147, 128), (230147, 384)):
    gen = torch.Generator(device=dev).manual_seed(0)
    ns = 4
    ys = [torch.randn((rows, C), device=dev, generator=gen) for _ in range(ns)]
    dzs = [torch.randn((rows, C), device=dev, generator=gen) for _ in range(ns)]
    gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    for bps in (0, 2, 8, 16):
        _lib.call("pgh_set_tuning", 2, bps)
        m0, r0 = O.bn_stats(ys[0], 1e-5, 0.1, rm, rv)
        sums0, _, _ = O.bn_act_bwd_reduce(dzs[0], ys[0], m0, r0, gam, bet, 1, None, None, None)
        t_stats = graph_time(lambda i: O.bn_stats(ys[i % ns], 1e-5, 0.1, rm, rv))
        t_fwd = graph_time(lambda i: O.bn_act_fwd(ys[i % ns], m0, r0, gam, bet, 1))
        t_red = graph_time(lambda i: O.bn_act_bwd_reduce(dzs[i % ns], ys[i % ns], m0, r0, gam, bet, 1, None, None, None))
        t_app = graph_time(lambda i: O.bn_act_bwd_apply(dzs[i % ns], ys[i % ns], m0, r0, gam, bet, sums0, None, 1, None, True, None))
        mb = rows * C * 4 / 1e6
        print(f"rows={rows:6d} C={C:3d} ({mb:6.1f} MB) ctas/sm={bps or 4:2d}: stats {t_stats:6.1f}  fwd {t_fwd:6.1f}  "
              f"bwd_reduce {t_red:6.1f}  bwd_apply {t_app:6.1f} us", flush=True)
    _lib.call("pgh_set_tuning", 2, 0)
