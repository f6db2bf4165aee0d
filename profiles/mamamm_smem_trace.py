"""Timeline of the mamamm algo-4 kernel (csrc/mamamm_smem.cu): clock64 stamps of CTA 0's producer
thread and first consumer warp (pgh_debug_trace hook), plus CUDA-graph-replay timing of the kernel
(no host overhead between launches)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygho_b200 import _lib, ops  # noqa: E402,F401

b, n, d = 128, 40, 128
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
sizes = torch.from_numpy(np.clip(np.rint(rng.normal(23.2, 4.5, b)), 9, n).astype(np.int64))
sizes[0] = n
ar = torch.arange(n)
mask = ((ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])).to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
sets = [(torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1),
         torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1)) for _ in range(3)]
ext = torch.stack((sizes, sizes, sizes), 1).to(torch.int32).to(dev)
mm = torch.ops.pygho_b200.mamamm


def graph_time(algo, e, dbg=0, reps=12):
    _lib.load().pgh_set_tuning(7, dbg)
    for i in range(3):
        mm(sets[i][0], False, sets[i][1], False, mask, e, algo)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            mm(sets[i % 3][0], False, sets[i % 3][1], False, mask, e, algo)
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


for algo in (() if os.environ.get("TRACE_ONLY") else (2, 4)):
    print(f"graph replay: algo {algo} ext {graph_time(algo, ext):7.1f} us   full {graph_time(algo, None):7.1f} us")
for dbg, name in () if os.environ.get("TRACE_ONLY") else ((1, "no FMAs"), (4, "no stores"), (13, "loads only"), (14, "FMAs only"), (15, "queue + barriers only"),
                  (16, "four channels per lane only"), (32, "no mask loads"), (64, "mask loads, no tile stores"),
                  (33, "no FMAs, no mask loads"), (65, "no FMAs, mask loads, no tile stores")):
    print(f"graph replay: algo 4 ext ablation {dbg:2d} ({name}): {graph_time(4, ext, dbg):7.1f} us")

K = 4096
buf = torch.zeros(K, dtype=torch.int64, device=dev)
_lib.load().pgh_set_tuning(7, int(os.environ.get("TRACE_DBG", "0")))
_lib.call("pgh_debug_trace", buf.data_ptr(), buf.numel())
mm(sets[0][0], False, sets[0][1], False, mask, ext, 4,
   torch.argsort(sizes, descending=True, stable=True).to(torch.int32).to(dev) if os.environ.get('TRACE_LPT') else None)
torch.cuda.synchronize()
_lib.call("pgh_debug_trace", None, 0)
_lib.load().pgh_set_tuning(7, 0)
t = buf.cpu().numpy()
half = (K - 2) // 2
ev = []
for region in (0, 1):
    cnt = int(t[region])
    ev += t[2 + region * half: 2 + region * half + 2 * cnt].reshape(-1, 2).tolist()
t0 = min(e[1] for e in ev)
names = {1: "P issue", 2: "C start", 3: "C done ", 4: "C unit ", 5: "C epi  ", 6: "P begin", 7: "P unit issued", 8: "P next unit known",
         9: "P end marker sent", 10: "P pad fill drained"}
print(f"{len(ev)} events of CTA 0 (cycles since first stamp; 1 us = ~1900 cycles), TRACE_DBG={os.environ.get('TRACE_DBG', '0')}")
for tag, clk in sorted(ev, key=lambda r: r[1]):
    role, item, slab, ch = tag >> 56, (tag >> 16) & 0xFFFFFF, (tag >> 8) & 0xFF, tag & 0xFF
    print(f"{clk - t0:9d}  {names.get(role, role)}  item {item:3d} (n={int(sizes[item]):2d}) slab {slab} chunk {ch}")
