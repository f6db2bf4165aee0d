"""linear_stats_kernel (TMA + tcgen05 GEMM with BatchNorm-statistics epilogue) and the ticketed
BatchNorm kernels alone at the SSWL+ sizes (230 147 tuples), for ncu and for timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygho_b200.ops  # noqa: E402,F401

M = int(os.environ.get("ROWS", "230147"))
K = int(os.environ.get("K", "384"))
N = int(os.environ.get("N", "128"))
ITERS = int(os.environ.get("ITERS", "12"))
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
ops = torch.ops.pygho_b200
gen = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn((M, K), device=dev, generator=gen) for _ in range(3)]
w = torch.randn((N, K), device=dev, generator=gen) / K ** 0.5
b = torch.randn((N,), device=dev, generator=gen)
ys = [torch.randn((M, N), device=dev, generator=gen) for _ in range(3)]
dz = torch.randn((M, N), device=dev, generator=gen)
gam, bet = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)


def timeit(name, fn, nbytes):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / ITERS
    print(f"{name}: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s  ({nbytes / 1e6:.0f} MB algorithmic)")


timeit(f"linear_stats {M}x{K}->{N}", lambda i: ops.linear_stats(xs[i % 3], w, b, 1e-5, 0.1, None, None, None, None, False),
       4 * (M * K + N * K + M * N))
timeit(f"cuBLAS linear {M}x{K}->{N}", lambda i: torch.nn.functional.linear(xs[i % 3], w, b), 4 * (M * K + N * K + M * N))
timeit(f"bn_stats {M}x{N}", lambda i: ops.bn_stats(ys[i % 3], 1e-5, 0.1, None, None, None, None), 4 * M * N)
mean, rstd = ops.bn_stats(ys[0], 1e-5, 0.1, None, None, None, None)
timeit(f"bn_act_fwd {M}x{N}", lambda i: ops.bn_act_fwd(ys[i % 3], mean, rstd, gam, bet, 1, None, None), 8 * M * N)
timeit(f"bn_act_bwd_reduce {M}x{N}", lambda i: ops.bn_act_bwd_reduce(dz, ys[i % 3], mean, rstd, gam, bet, 1, None, None, None),
       8 * M * N)
sums, _, _ = ops.bn_act_bwd_reduce(dz, ys[0], mean, rstd, gam, bet, 1, None, None, None)
timeit(f"bn_act_bwd_apply {M}x{N}", lambda i: ops.bn_act_bwd_apply(dz, ys[i % 3], mean, rstd, gam, bet, sums, None, 1, None, True, None),
       12 * M * N)
