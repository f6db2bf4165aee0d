"""Kernel-variant sweep for the fused gather-multiply-segmented-reduce family at cfg5 sizes
(B=1024 ZINC-shaped graphs, d=128).  Every case is captured into a CUDA graph (NSET rotating
operand sets x REP calls) so the numbers are device time per launch without Python /
dispatcher overhead; results of every variant are checked bit-exact against variant 2
(same reduction order by construction).

    python profiles/bench_gmr.py [--variants -1,2,10,11] [--epw 32,64,128]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygho_b200 import SparseTensor, _lib  # noqa: E402
from pygho_b200 import plans as P  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="-1,2,10,11,12,13,14,15")
ap.add_argument("--epw", default="64")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--cases", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
d = 128
hb = make_batch(args.batch, seed=0)
ei, tid = torch.from_numpy(hb.edge_index).to(dev), torch.from_numpy(hb.tupleid).to(dev)
N, nA, nX = hb.num_nodes, ei.shape[1], tid.shape[1]
gen = torch.Generator(device=dev).manual_seed(0)
NSET, REP = 4, 3
Xs = [torch.randn(nX, d, device=dev, generator=gen) for _ in range(NSET)]
As = [torch.randn(nA, d, device=dev, generator=gen) for _ in range(NSET)]
xs = [torch.randn(N, d, device=dev, generator=gen) for _ in range(NSET)]
ops = torch.ops.pygho_b200


def gtime(fn):
    """us per call of fn(i), measured by replaying a captured graph of NSET*REP calls."""
    for i in range(NSET):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(REP):
            for i in range(NSET):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (NSET * REP))
    return best


cases = []   # (name, fn(i) -> tensor, algorithmic bytes)
keys = {"X___X___1___A___0": (tid, 1, ei, 0), "X___A___1___X___0": (ei, 1, tid, 0),
        "X___X___1___X___0": (tid, 1, tid, 0)}
for key, (i1, d1, i2, d2) in keys.items():
    acd, _ = P.filtered_plan(tid, i1, d1, i2, d2, k2_sorted=True)
    T = acd.shape[1]
    n1, n2 = i1.shape[1], i2.shape[1]
    v1 = Xs if n1 == nX else As
    v2 = Xs if n2 == nX else As
    plan = P.plan_from_acd(acd, nX, n1, n2).prefetch()
    ga, gc, gd = plan.group("a"), plan.group("c"), plan.group("d")
    short = key.replace("___", "")
    cases.append((f"fwd sum {short}", lambda i, v1=v1, v2=v2, ga=ga: ops.seg_gmr(v1[i], ga.first, None, v2[i], ga.second, ga.rowptr, nX, 0),
                  4 * d * (n1 + n2 + nX) + 4 * (2 * T + nX + 1)))
    cases.append((f"fwd max {short}", lambda i, v1=v1, v2=v2, ga=ga: ops.seg_gmr(v1[i], ga.first, None, v2[i], ga.second, ga.rowptr, nX, 2),
                  4 * d * (n1 + n2 + nX) + 4 * (2 * T + nX + 1)))
    cases.append((f"bwd dA  {short}", lambda i, v2=v2, gc=gc, n1=n1: ops.seg_gmr(Xs[i], gc.first, None, v2[i], gc.second, gc.rowptr, n1, 0),
                  4 * d * (nX + n2 + n1) + 4 * (2 * T + n1 + 1)))
    cases.append((f"bwd dB  {short}", lambda i, v1=v1, gd=gd, n2=n2: ops.seg_gmr(Xs[i], gd.first, None, v1[i], gd.second, gd.rowptr, n2, 0),
                  4 * d * (nX + n1 + n2) + 4 * (2 * T + n2 + 1)))

X0 = SparseTensor(tid, Xs[0], (N, N, d), True)
A0 = SparseTensor(ei, As[0], (N, N, d), True)
from pygho_b200.backend.Spmm import _spmm_plan  # noqa: E402
sp = _spmm_plan(A0, 1).group("a")
cases.append(("spmm A x sum", lambda i: ops.seg_gmr(As[i], sp.first, None, xs[i], sp.second, sp.rowptr, N, 0),
              4 * (d * nA + 2 * d * N) + 4 * (nA + N + 1)))
for dims, keyrow in (([1], 0), ([0], 1)):
    pg = X0._key_plan((keyrow,)).group("a")
    for aggr, code in (("sum", 0), ("mean", 1), ("max", 2)):
        cases.append((f"pool {aggr} dims={dims}", lambda i, pg=pg, code=code: ops.seg_gmr(Xs[i], pg.first, None, None, None, pg.rowptr, N, code),
                      4 * d * (nX + N) + 4 * (N + 1) + (4 * nX if keyrow else 0)))
for dim in (0, 1):
    ug = X0._key_plan((dim,)).transposed().group("a")
    cases.append((f"unpool dim={dim}", lambda i, ug=ug: ops.seg_gmr(xs[i], ug.first, None, None, None, ug.rowptr, nX, 0),
                  4 * d * (N + nX) + 4 * nX))

if args.cases:
    want = args.cases.split(",")
    cases = [c for c in cases if any(w in c[0] for w in want)]

variants = [int(v) for v in args.variants.split(",")]
epws = [int(v) for v in args.epw.split(",")]
lib.pgh_set_tuning(0, 2)
ref = {name: fn(0).clone() for name, fn, _ in cases}
print(f"peak {PEAK} GB/s; B={args.batch} N={N} nA={nA} nX={nX}", flush=True)
hdr = "case".ljust(34) + "".join(f"v{v}/e{e}".rjust(12) for v in variants for e in (epws if v >= 10 else epws[:1]))
print(hdr, flush=True)
for name, fn, nbytes in cases:
    line = name.ljust(34)
    for v in variants:
        for e in (epws if v >= 10 else epws[:1]):
            lib.pgh_set_tuning(0, v)
            lib.pgh_set_tuning(1, e)
            got = fn(0)
            ok = torch.equal(got, ref[name])
            us = gtime(fn)
            line += f"{us:7.1f}{'' if ok else '!'}{100 * nbytes / us / 1e3 / PEAK:4.0f}%".rjust(12)
    print(line, flush=True)
lib.pgh_set_tuning(0, -1)
