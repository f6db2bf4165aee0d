"""mamamm algo 4: kernel time against the number of graphs (fixed cost vs per-unit cost), CUDA-graph
replay, for the whole kernel and for its ablations (pgh_set_tuning key 7)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygho_b200 import _lib, ops  # noqa: E402,F401

n, d = 40, 128
dev = torch.device("cuda", 0)
mm = torch.ops.pygho_b200.mamamm


def graph_time(A, B, mask, e, algo, dbg=0, reps=12, order=None):
    _lib.load().pgh_set_tuning(7, dbg)
    for i in range(3):
        mm(A[i], False, B[i], False, mask, e, algo, order)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            mm(A[i % 3], False, B[i % 3], False, mask, e, algo, order)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    _lib.load().pgh_set_tuning(7, 0)
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


for b in (1, 8, 32, 64, 128, 256, 512):
    rng = np.random.default_rng(0)
    sizes = torch.from_numpy(np.clip(np.rint(rng.normal(23.2, 4.5, b)), 9, n).astype(np.int64))
    ar = torch.arange(n)
    mask = ((ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])).to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    A = [torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1) for _ in range(3)]
    B = [torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1) for _ in range(3)]
    ext = torch.stack((sizes, sizes, sizes), 1).to(torch.int32).to(dev)
    valid = 4 * d * float((2 * sizes.double() ** 2).sum() + b * n * n) + b * n * n
    row = [f"b={b:4d} ({valid / 1e6:6.1f} MB)"]
    lpt = torch.argsort(sizes, descending=True, stable=True).to(torch.int32).to(dev)
    row.append(f"algo 2 {graph_time(A, B, mask, ext, 2):6.1f}")
    row.append(f"algo 4 {graph_time(A, B, mask, ext, 4):6.1f}")
    for dbg, name in ((0, "LPT"), (4, "no epi"), (1, "no FMA"), (13, "loads"), (14, "FMAs"), (15, "skel")):
        row.append(f"{name} {graph_time(A, B, mask, ext, 4, dbg, order=lpt):6.1f}")
    print("  ".join(row), flush=True)
    del A, B
