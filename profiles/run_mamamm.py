"""mamamm (2-FWL contraction) microbench at cfg3: b=128 graphs, n<=40, d=128.
Times algo 0 (CUDA-core), algo 1 (tcgen05, if available) and the reference's own GPU path
(permute + torch.matmul, backend/Mamamm.py:40-63) with CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygho_b200  # noqa: E402,F401
from pygho_b200 import ops  # noqa: E402,F401

b, n, d = int(os.environ.get("B", "128")), int(os.environ.get("N", "40")), int(os.environ.get("D", "128"))
ITERS = int(os.environ.get("ITERS", "10"))
ALGOS = [int(a) for a in os.environ.get("ALGOS", "0,1").split(",")]
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
sizes = torch.from_numpy(np.clip(np.rint(rng.normal(23.2, 4.5, b)), 9, n).astype(np.int64))
sizes[0] = n
ar = torch.arange(n)
mask = ((ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])).to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
sets = [(torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1),
         torch.randn((b, n, n, d), device=dev, generator=gen) * mask.unsqueeze(-1)) for _ in range(3)]
alg_bytes = 4 * d * b * 3 * n * n + b * n * n
useful_flops = 2 * d * float((sizes.double() ** 3).sum())
padded_flops = 2 * d * b * n ** 3
torch.backends.cuda.matmul.allow_tf32 = True


def timeit(fn):
    for i in range(3):
        fn(*sets[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(*sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / ITERS


def ref_path(A, B):
    # what the reference does on the GPU: move batch/dense dims, matmul, move back
    tA, tB = A.permute(3, 0, 1, 2), B.permute(3, 0, 1, 2)
    return torch.matmul(tA, tB).permute(1, 2, 3, 0) * mask.unsqueeze(-1)


want = torch.einsum("bijd,bjkd->bikd", sets[0][0].double(), sets[0][1].double()) * mask.unsqueeze(-1)
ext = torch.stack((sizes, sizes, sizes), 1).to(torch.int32).to(dev)
valid_bytes = 4 * d * float((2 * sizes.double() ** 2).sum() + b * n * n) + b * n * n
for algo in ALGOS:
    for tag, e in (("full", None), ("ext ", ext)):
        try:
            out = torch.ops.pygho_b200.mamamm(sets[0][0], False, sets[0][1], False, mask, e, algo)
        except Exception as ex:  # noqa: BLE001
            print(f"algo {algo}: unavailable ({str(ex)[:80]})")
            continue
        err = float((out.double() - want).abs().max() / want.abs().max())
        us = timeit(lambda A, B: torch.ops.pygho_b200.mamamm(A, False, B, False, mask, e, algo))
        nbytes = alg_bytes if e is None else valid_bytes
        print(f"algo {algo} {tag}: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s ({nbytes / 1e6:.0f} MB moved by design)  "
              f"{padded_flops / us / 1e6:6.2f} TFLOP/s padded, {useful_flops / us / 1e6:6.2f} useful  rel.err {err:.2e}")
if os.environ.get("ABLATE") and 4 in ALGOS:
    from pygho_b200 import _lib
    names = {0: "all", 1: "no FMAs", 2: "no loads", 3: "no loads, no FMAs", 4: "no stores", 8: "no pad fill",
             12: "no stores, no pad fill", 13: "loads only", 14: "FMAs only", 15: "queue + barriers only",
             16: "four channels per lane only", 30: "FMAs only, four channels per lane"}
    for dbg, name in names.items():
        _lib.load().pgh_set_tuning(7, dbg)
        for tag, e in (("full", None), ("ext ", ext)):
            us = timeit(lambda A, B: torch.ops.pygho_b200.mamamm(A, False, B, False, mask, e, 4))
            print(f"algo 4 {tag} ablation {dbg:2d} ({name}): {us:8.1f} us")
    for hint in (0, 1000000):
        for dbg in (0, 14, 15):
            _lib.load().pgh_set_tuning(7, dbg | (hint << 8))
            us = timeit(lambda A, B: torch.ops.pygho_b200.mamamm(A, False, B, False, mask, ext, 4))
            print(f"algo 4 ext  wait hint {hint:7d} ns, ablation {dbg:2d}: {us:8.1f} us")
    _lib.load().pgh_set_tuning(7, 0)
# the gradient contractions' layouts (g @ B^T, A^T @ g) on the shipped kernels
for algo in ALGOS:
    if algo not in (2, 4):
        continue
    for ta, tb in ((False, True), (True, False), (True, True)):
        us = timeit(lambda A, B: torch.ops.pygho_b200.mamamm(A, ta, B, tb, mask, ext, algo))
        print(f"algo {algo} ext  trans_a={int(ta)} trans_b={int(tb)}: {us:8.1f} us  {valid_bytes / us / 1e3:7.1f} GB/s")
us = timeit(ref_path)
out = ref_path(*sets[0])
err = float((out.double() - want).abs().max() / want.abs().max())
print(f"torch permute+matmul (reference GPU path, tf32): {us:8.1f} us  {alg_bytes / us / 1e3:7.1f} GB/s  rel.err {err:.2e}")
