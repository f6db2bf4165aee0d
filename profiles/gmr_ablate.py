"""Where does the time of the spspmm forward launch (key X___X___1___A___0, B=1024) go?
Ablations with the shipped kernel (results of the no-store mode are WRONG; timing only):

    full            the launch bench.py's roofline object times
    no stores       tuning key 6 = 1: rows are reduced but never written
    single operand  only the X gather (no A rows, no multiply)
    A only          only the (L2-resident) A gather
    sequential X    same plan shape, X gathered in row order (no reuse, perfect streaming)
    rows/warp       the full launch with other work splits (PYGHO_B200 tuning is per process,
                    so the split is varied through the plan: not available -> variants only)
    variants        stream kernel with 8 / 4 / 2 entries in flight, ring variants
"""
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
g = plan.group("a")
T = plan.T
gen = torch.Generator(device=dev).manual_seed(0)
NSET = 4
Xs = [torch.randn(nX, d, device=dev, generator=gen) for _ in range(NSET)]
As = [torch.randn(nA, d, device=dev, generator=gen) for _ in range(NSET)]
alg = 4 * d * (2 * nX + nA) + 4 * (2 * T + nX + 1)


def tune(k, v):
    _lib.call("pgh_set_tuning", k, v)


def timeit(fn, iters=40):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def show(name, us, nbytes=alg):
    print(f"{name:52s} {us:7.1f} us  {nbytes / us / 1e3:6.0f} GB/s on {nbytes / 1e6:.0f} MB", flush=True)


full = lambda i: ops.seg_gmr(Xs[i % NSET], g.first, None, As[i % NSET], g.second, g.rowptr, nX, 0)  # noqa: E731
show("full (default variant)", timeit(full))
tune(6, 1)
show("no stores", timeit(full), 4 * d * (nX + nA) + 4 * (2 * T + nX + 1))
tune(6, 0)
tune(0, 2)     # keep the register-staged stream kernel for the single-operand cases
show("single operand: X gather only (stream kernel)",
     timeit(lambda i: ops.seg_gmr(Xs[i % NSET], g.first, None, None, None, g.rowptr, nX, 0)),
     4 * d * 2 * nX + 4 * (T + nX + 1))
show("single operand: A gather only (stream kernel)",
     timeit(lambda i: ops.seg_gmr(As[i % NSET], g.second, None, None, None, g.rowptr, nX, 0)),
     4 * d * (nA + nX) + 4 * (T + nX + 1))
seq_c = (torch.arange(T, device=dev, dtype=torch.int64) * nX // T).to(torch.int32)
show("sequential X rows + real A gather",
     timeit(lambda i: ops.seg_gmr(Xs[i % NSET], seq_c, None, As[i % NSET], g.second, g.rowptr, nX, 0)))
seq_d = (torch.arange(T, device=dev, dtype=torch.int64) * nA // T).to(torch.int32)
show("sequential X rows + sequential A rows",
     timeit(lambda i: ops.seg_gmr(Xs[i % NSET], seq_c, None, As[i % NSET], seq_d, g.rowptr, nX, 0)))
tune(6, 1)
show("sequential X + sequential A, no stores",
     timeit(lambda i: ops.seg_gmr(Xs[i % NSET], seq_c, None, As[i % NSET], seq_d, g.rowptr, nX, 0)),
     4 * d * (nX + nA) + 4 * (2 * T + nX + 1))
tune(6, 0)
for v, name in ((1, "stream 8 in flight, 1 CTA/SM bound"), (2, "stream 4 in flight, 4 CTAs/SM"),
                (3, "stream 2 in flight, 5 CTAs/SM"), (0, "stream 4 in flight, no occupancy bound"),
                (10, "ring 16 stages x4, 4 warps"), (11, "ring 8 stages x4, 8 warps"),
                (13, "ring 8 stages x2, 4 warps"), (17, "ring 16 stages x2, 8 warps")):
    tune(0, v)
    show(f"variant {v}: {name}", timeit(full))
for v, name in ((30, "lean 4 in flight, 4 CTAs/SM"), (31, "lean 8 in flight, 2 CTAs/SM"),
                (32, "lean 2 in flight, 5 CTAs/SM"), (33, "lean 4 in flight, 3 CTAs/SM")):
    tune(0, v)
    show(f"variant {v}: {name}", timeit(full))
    show(f"variant {v}: single operand X gather",
         timeit(lambda i: ops.seg_gmr(Xs[i % NSET], g.first, None, None, None, g.rowptr, nX, 0)),
         4 * d * 2 * nX + 4 * (T + nX + 1))
gc_, gd_ = plan.group("c"), plan.group("d")
for v in (-1, 30, 31):
    tune(0, v)
    show(f"variant {v}: bwd dA (by c, {nX} rows)",
         timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gc_.first, None, As[i % NSET], gc_.second, gc_.rowptr, nX, 0)))
    show(f"variant {v}: bwd dB (by d, {nA} rows)",
         timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gd_.first, None, Xs[(i + 1) % NSET], gd_.second, gd_.rowptr, nA, 0)))
for ent in (32,):
    tune(7, ent)
    for v, name in ((20, "bulk G4 NS4 W4 (64 KB, 3 CTAs/SM)"), (21, "bulk G8 NS3 W4 (96 KB, 2 CTAs/SM)"),
                    (22, "bulk G4 NS4 W8 (128 KB, 1 CTA/SM)"), (23, "bulk G4 NS6 W4 (96 KB, 2 CTAs/SM)"),
                    (24, "bulk G8 NS3 W2 (48 KB, 4 CTAs/SM)"), (25, "bulk G4 NS2 W4 (32 KB, 7 CTAs/SM)")):
        tune(0, v)
        show(f"variant {v}: {name}, {ent} entries/warp", timeit(full))
tune(7, 0)
tune(0, 20)
tune(6, 0)
show("bulk 20: single operand X gather only",
     timeit(lambda i: ops.seg_gmr(Xs[i % NSET], g.first, None, None, None, g.rowptr, nX, 0)),
     4 * d * 2 * nX + 4 * (T + nX + 1))
pool_rp = torch.arange(0, nX + 1, 10, device=dev, dtype=torch.int32)
npool = pool_rp.numel() - 1
for v in (13, 20, 30, 31):
    tune(0, v)
    show(f"variant {v}: pooling-like (identity index, 10 rows per segment)",
         timeit(lambda i: ops.seg_gmr(Xs[i % NSET], None, None, None, None, pool_rp, npool, 0)),
         4 * d * (nX + npool) + 4 * (npool + 1))
tune(0, -1)
# a plain copy of the same byte volume for reference (what 'peak' means on this box)
src = torch.empty(alg // 8, dtype=torch.float32, device=dev)
dst = torch.empty_like(src)
show("torch copy_ of the same byte volume", timeit(lambda i: dst.copy_(src)), alg // 8 * 8)
