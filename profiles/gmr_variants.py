"""Time seg_gmr variants (argv: variant ids) on the B=1024 SSWL key: forward, dA, dB launches.
    python profiles/gmr_variants.py 30 34 35"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygho_b200 import _lib  # noqa: E402
from pygho_b200 import plans as P  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

ops = torch.ops.pygho_b200
dev = torch.device("cuda", 0)
d = 128
hb = make_batch(1024, seed=0)
ei, tid = torch.from_numpy(hb.edge_index).to(dev), torch.from_numpy(hb.tupleid).to(dev)
nX, nA = tid.shape[1], ei.shape[1]
acd, _ = P.filtered_plan(tid, tid, 1, ei, 0, k2_sorted=True)
plan = P.plan_from_acd(acd, nX, nX, nA)
ga, gc, gd = plan.group("a"), plan.group("c"), plan.group("d")
gen = torch.Generator(device=dev).manual_seed(0)
NSET = 4
Xs = [torch.randn(nX, d, device=dev, generator=gen) for _ in range(NSET)]
As = [torch.randn(nA, d, device=dev, generator=gen) for _ in range(NSET)]
alg = 4 * d * (2 * nX + nA) + 4 * (2 * plan.T + nX + 1)


def timeit(fn, iters=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


ref = None
for v in [int(a) for a in sys.argv[1:]] or [30]:
    _lib.call("pgh_set_tuning", 0, v)
    out = ops.seg_gmr(Xs[0], ga.first, None, As[0], ga.second, ga.rowptr, nX, 0)
    if ref is None:
        ref = out
    same = bool(torch.equal(out, ref))
    f = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], ga.first, None, As[i % NSET], ga.second, ga.rowptr, nX, 0))
    a = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gc.first, None, As[i % NSET], gc.second, gc.rowptr, nX, 0))
    b = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gd.first, None, Xs[(i + 1) % NSET], gd.second, gd.rowptr, nA, 0))
    print(f"variant {v}: fwd {f:6.1f} us ({alg / f / 1e3:5.0f} GB/s)  dA {a:6.1f} us  dB {b:6.1f} us  "
          f"bitwise equal to first variant: {same}", flush=True)
_lib.call("pgh_set_tuning", 0, -1)
