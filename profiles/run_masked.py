"""Masked (dense) path at cfg3 (b = 128, n <= 40, d = 128): masked pooling forward / backward,
masked fill and mamamm, CUDA-graph replays over 6 rotating input sets (630 MB >> L2).

    python profiles/run_masked.py [--mamamm]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygho_b200 import ops  # noqa: E402,F401

ap = argparse.ArgumentParser()
ap.add_argument("--mamamm", action="store_true")
ap.add_argument("--b", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda", 0)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
b, n, d = args.b, 40, 128
rng = np.random.default_rng(0)
sizes = torch.from_numpy(np.clip(np.rint(rng.normal(23.2, 4.5, b)), 9, n).astype(np.int64))
sizes[0] = n
ar = torch.arange(n)
mask = ((ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])).to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
NM = 6
Ms = [torch.randn(b, n, n, d, device=dev, generator=gen) * mask.unsqueeze(-1) for _ in range(NM)]
P = torch.ops.pygho_b200


def timeit(fn, iters=12):
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
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * iters)


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:58s} {us:7.1f} us  {gbs:6.0f} GB/s  {gbs / PEAK:5.2f} of peak  ({nbytes / 1e6:.1f} MB)", flush=True)


s2 = float((sizes.double() ** 2).sum())
for red, nm in ((2, "dim 2"), (1, "dim 1"), (3, "dims 1+2")):
    n_out = b * (1 if red == 3 else n)
    for aggr, code in (("sum", 0), ("mean", 1), ("max", 2)):
        t = timeit(lambda i: P.masked_pool(Ms[i % NM], mask, red, code))
        valid = 4 * d * (s2 + n_out) + b * n * n
        report(f"masked pool {aggr} {nm} ({b},{n},{n},{d})", t, valid)
        out, _ = P.masked_pool(Ms[0], mask, red, code)
        gs = [torch.randn_like(out) for _ in range(3)]
        outs = [P.masked_pool(Ms[i], mask, red, code)[0] for i in range(NM)]
        t = timeit(lambda i: P.masked_pool_bwd(Ms[i % NM], mask, outs[i % NM], gs[i % 3], red, code))
        # sum / mean write the whole gradient (pads = 0) and read g; max / min also read the data
        bw = 4 * d * (b * n * n + n_out) + b * n * n + (4 * d * (s2 + n_out) if code == 2 else 0)
        report(f"masked pool bwd {aggr} {nm}", t, bw)
rows = b * n * n
t = timeit(lambda i: P.masked_fill_rows(Ms[i % NM], mask, 0.0))
report("masked fill (pads -> 0)", t, 4 * d * (s2 + rows) + rows)

if args.mamamm:
    ext = torch.stack((sizes, sizes, sizes), 1).to(torch.int32).to(dev)
    valid = 4 * d * float(2 * s2 + b * n * n) + b * n * n
    lpt = torch.argsort(sizes, descending=True, stable=True).to(torch.int32).to(dev)
    for algo, order, name in ((2, None, "tcgen05 tf32 pipeline"), (4, None, "fp32 smem ring, input order"),
                              (4, lpt, "fp32 smem ring, largest first")):
        t = timeit(lambda i: P.mamamm(Ms[i % NM], False, Ms[(i + 1) % NM], False, mask, ext, algo, order))
        report(f"mamamm algo {algo} ({name})", t, valid)
