"""Timeline of the pipelined mamamm kernel (algo 2): %globaltimer stamps of the copy,
transposer, MMA and epilogue roles of CTA 0 for its first work items (pgh_debug_trace hook)."""
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
A = torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1)
B = torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1)
ext = torch.stack((sizes, sizes, sizes), 1).to(torch.int32).to(dev)
for _ in range(3):
    torch.ops.pygho_b200.mamamm(A, False, B, False, mask, ext, 2)
K = 32
buf = torch.zeros(4 * K * 4, dtype=torch.int64, device=dev)
_lib.call("pgh_debug_trace", buf.data_ptr(), buf.numel())
torch.ops.pygho_b200.mamamm(A, False, B, False, mask, ext, 2)
torch.cuda.synchronize()
_lib.call("pgh_debug_trace", None, 0)
t = buf.cpu().numpy().reshape(4, K, 4)
t0 = t[t > 0].min()
names = [("C", ()), ("T", ()), ("M", ()), ("E", ())]
items = [it for it in range(0, b * d // 8, 148)]
for q in range(16):
    g = items[q] // 16 if q < len(items) else -1
    line = f"item {q:2d} (n={int(sizes[g]) if g >= 0 else -1:2d}) "
    for r, (nm, evs) in enumerate(names):
        line += f"| {nm} " + " ".join(f"{(t[r, q, e] - t0) / 1e3:6.2f}" if t[r, q, e] else "   -  " for e in range(4)) + " "
    print(line)
print("us since first stamp.  C(opy warp) = zeros written / ring space granted / copies issued; "
      "T(ransposers) = mask loads issued / operands landed / tiles free / tiles written; "
      "M(MA warp) = top / tiles full / accumulator free / all rounds issued; "
      "E(pilogue) = top / mask row read / accumulator full / item stored")
