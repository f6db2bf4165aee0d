"""Op-level roofline table: every kernel family of the hot path at BASELINE sizes
(cfg5: B=1024 ZINC-shaped graphs, d=128; cfg3: b=128, n<=40, d=128), timed with CUDA events
on rotating inputs, against (a) the algorithmic bytes of SURVEY.md 8d / the measured HBM peak
and (b) the reference's own GPU path (the same ATen call chain the reference executes:
index_select + mul + scatter_reduce_(include_self=False), permute + matmul, ...).

    python profiles/run_ops.py [--md profiles/r1_op_rooflines.md]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygho_b200 import MaskedTensor, SparseTensor  # noqa: E402
from pygho_b200 import plans as P  # noqa: E402
from pygho_b200.backend import (filterind, spmm, spspmm, spspmm_ind,  # noqa: E402
                                torch_scatter_reduce)
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--md", default="")
ap.add_argument("--batch", type=int, default=1024)
args = ap.parse_args()
dev = torch.device("cuda", 0)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
d = 128
hb = make_batch(args.batch, seed=0)
ei, tid = torch.from_numpy(hb.edge_index).to(dev), torch.from_numpy(hb.tupleid).to(dev)
bvec = torch.from_numpy(hb.batch).to(dev)
N, nA, nX, B = hb.num_nodes, ei.shape[1], tid.shape[1], hb.num_graphs
gen = torch.Generator(device=dev).manual_seed(0)
NSET = 4
Xs = [torch.randn(nX, d, device=dev, generator=gen) for _ in range(NSET)]
As = [torch.randn(nA, d, device=dev, generator=gen) for _ in range(NSET)]
xs = [torch.randn(N, d, device=dev, generator=gen) for _ in range(NSET)]
rows = []


def timeit(fn, iters=12, graph=True):
    """us per call.  Kernel-level cases are replayed from a captured CUDA graph (device time,
    no Python / dispatcher overhead); cases that synchronise or allocate outside the pool
    (plan builders, reference chains with host syncs) fall back to eager event timing."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(iters):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
            return best
        except Exception:
            torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def report(name, ours_us, ref_us, nbytes, note=""):
    gbs = nbytes / ours_us / 1e3
    rows.append((name, ours_us, gbs, gbs / PEAK, ref_us, (ref_us / ours_us) if ref_us else None,
                 nbytes / 1e6, note))
    print(f"{name:46s} {ours_us:9.1f} us {gbs:8.0f} GB/s {100 * gbs / PEAK:5.1f}%  "
          f"ref-gpu {ref_us if ref_us else float('nan'):9.1f} us  {note}", flush=True)


def ref_scatter(src, ind, n, aggr):
    out = src.new_zeros((n,) + src.shape[1:])
    idx = ind.reshape((-1,) + (1,) * (src.ndim - 1)).expand_as(src)
    return out.scatter_reduce_(0, idx, src, {"sum": "sum", "mean": "mean", "max": "amax"}[aggr],
                               include_self=False)


# ------------------------------------------------------------------ plans
def two_step(i1, d1, i2, d2):
    tar, bcd = spspmm_ind(i1, d1, i2, d2)
    return filterind(tid, tar, bcd)


keys = {"X___X___1___A___0": (tid, 1, ei, 0), "X___A___1___X___0": (ei, 1, tid, 0),
        "X___X___1___X___0": (tid, 1, tid, 0)}
acds = {}
for key, (i1, d1, i2, d2) in keys.items():
    acd, _ = P.filtered_plan(tid, i1, d1, i2, d2, k2_sorted=True)
    acds[key] = acd
    T = acd.shape[1]
    t_fused = timeit(lambda i: P.filtered_plan(tid, i1, d1, i2, d2, k2_sorted=True), 5, graph=False)
    t_two = timeit(lambda i: two_step(i1, d1, i2, d2), 5, graph=False)
    plan_bytes = 8 * (i1.numel() + i2.numel() + tid.numel()) + 24 * T
    report(f"plan {key} fused (T={T})", t_fused, None, plan_bytes, "int64 in + acd out")
    report(f"plan {key} spspmm_ind+filterind", t_two, None, plan_bytes, "reference-API path, 4 syncs")
    t_csr = timeit(lambda i: P.plan_from_acd(acd.clone(), nX, i1.shape[1], i2.shape[1]).prefetch(), 5, graph=False)
    report(f"plan {key} CSR regroup (a,c,d)", t_csr, None, 24 * T + 3 * 12 * T, "3 stable sorts")

# ------------------------------------------------------------------ spspmm
for key, (i1, d1, i2, d2) in keys.items():
    acd = acds[key]
    T = acd.shape[1]
    n1, n2 = i1.shape[1], i2.shape[1]
    v1 = Xs if n1 == nX else As
    v2 = Xs if n2 == nX else As
    plan = P.plan_from_acd(acd, nX, n1, n2).prefetch()
    ops = torch.ops.pygho_b200
    ga, gc, gd = plan.group("a"), plan.group("c"), plan.group("d")
    fwd_bytes = 4 * d * (n1 + n2 + nX) + 4 * (2 * T + nX + 1)
    for aggr, code in (("sum", 0), ("max", 2)):
        t = timeit(lambda i: ops.seg_gmr(v1[i % NSET], ga.first, None, v2[i % NSET], ga.second, ga.rowptr, nX, code))
        tr = timeit(lambda i: ref_scatter(v1[i % NSET][acd[1]] * v2[i % NSET][acd[2]], acd[0], nX, aggr), 5)
        report(f"spspmm fwd {aggr} {key}", t, tr, fwd_bytes)
    t = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gc.first, None, v2[i % NSET], gc.second, gc.rowptr, n1, 0))
    report(f"spspmm bwd dA {key}", t, None, 4 * d * (nX + n2 + n1) + 4 * (2 * T + n1 + 1))
    t = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], gd.first, None, v1[i % NSET], gd.second, gd.rowptr, n2, 0))
    report(f"spspmm bwd dB {key}", t, None, 4 * d * (nX + n1 + n2) + 4 * (2 * T + n2 + 1))

# ------------------------------------------------------------------ spmm / pooling / unpooling
# (kernel-level: torch.ops.pygho_b200.seg_gmr on the cached plan groups, like the autograd
#  functions call it; the Python API adds ~25 us of host time per call on top)
A_sp = [SparseTensor(ei, As[i], (N, N, d), True) for i in range(NSET)]
X_sp = [SparseTensor(tid, Xs[i], (N, N, d), True) for i in range(NSET)]
ops = torch.ops.pygho_b200
from pygho_b200.backend.Spmm import _spmm_plan  # noqa: E402
sp = _spmm_plan(A_sp[0], 1).group("a")
t = timeit(lambda i: ops.seg_gmr(As[i % NSET], sp.first, None, xs[i % NSET], sp.second, sp.rowptr, N, 0))
tr = timeit(lambda i: ref_scatter(As[i % NSET] * xs[i % NSET][ei[1]], ei[0], N, "sum"), 5)
report("spmm A x (sum)", t, tr, 4 * (d * nA + 2 * d * N) + 4 * (nA + N + 1))
t_api = timeit(lambda i: spmm(A_sp[i % NSET], 1, xs[i % NSET], "sum"), graph=False)
report("spmm A x (sum) through the Python API", t_api, tr, 4 * (d * nA + 2 * d * N) + 4 * (nA + N + 1),
       "host-overhead bound")
for dims, keyrow, note in (([1], 0, "sorted key"), ([0], 1, "unsorted key -> perm")):
    pg = X_sp[0]._key_plan((keyrow,)).group("a")
    for aggr, code in (("sum", 0), ("mean", 1), ("max", 2)):
        t = timeit(lambda i: ops.seg_gmr(Xs[i % NSET], pg.first, None, None, None, pg.rowptr, N, code))
        tr = timeit(lambda i: ref_scatter(Xs[i % NSET], tid[keyrow], N, aggr), 5)
        report(f"sparse pool {aggr} dims={dims}", t, tr, 4 * d * (nX + N) + 4 * (N + 1) + (4 * nX if keyrow else 0), note)
for dim in (0, 1):
    ug = X_sp[0]._key_plan((dim,)).transposed().group("a")
    t = timeit(lambda i: ops.seg_gmr(xs[i % NSET], ug.first, None, None, None, ug.rowptr, nX, 0))
    tr = timeit(lambda i: xs[i % NSET][tid[dim]], 5)
    report(f"unpooling from dense dim={dim}", t, tr, 4 * d * (N + nX) + 4 * nX)
rg = P.plan_from_key(bvec, B, False).group("a")
t = timeit(lambda i: ops.seg_gmr(xs[i % NSET], rg.first, None, None, None, rg.rowptr, B, 0))
tr = timeit(lambda i: ref_scatter(xs[i % NSET], bvec, B, "sum"), 5)
report("graph read-out (N -> B) sum", t, tr, 4 * d * (N + B) + 4 * (B + 1))

# 3-D tuples: pooling to sparse (coalesce path of the reference)
hb3 = make_batch(64, seed=1, tuples="i2")
tid3 = torch.from_numpy(hb3.tupleid).to(dev)
n3, N3 = tid3.shape[1], hb3.num_nodes
X3 = [SparseTensor(tid3, torch.randn(n3, d, device=dev, generator=gen), (N3, N3, N3, d), True) for _ in range(2)]
t = timeit(lambda i: X3[i % 2].sum([2], return_sparse=True), 6, graph=False)


def ref_pool_sparse(i):
    key = tid3[0] * N3 + tid3[1]
    uk, inv = torch.unique(key, return_inverse=True)
    return ref_scatter(X3[i % 2].values, inv, uk.shape[0], "sum")


tr = timeit(ref_pool_sparse, 5, graph=False)
report(f"3-D pool to sparse (nnz={n3}, B=64 i2 tuples)", t, tr, 4 * d * (n3 + n3 // 20) + 4 * n3,
       "plan cached vs unique per call")

# ------------------------------------------------------------------ fused BatchNorm + SiLU
for C in (384, 128):
    ys = [torch.randn(nX, C, device=dev, generator=gen) for _ in range(3)]
    dz = torch.randn(nX, C, device=dev, generator=gen)
    gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    ops = torch.ops.pygho_b200

    def ours_fwd(i):
        m, r = ops.bn_stats(ys[i % 3], 1e-5, 0.1, None, None)
        return ops.bn_act_fwd(ys[i % 3], m, r, gam, bet, 1), m, r

    def ref_fwd(i):
        return torch.nn.functional.silu(torch.nn.functional.batch_norm(ys[i % 3], None, None, gam, bet, True, 0.1, 1e-5))

    t, tr = timeit(ours_fwd), timeit(ref_fwd, 5)
    report(f"BN(train)+SiLU fwd ({nX}x{C})", t, tr, 4 * C * nX * 3, "3 passes by design")
    _, m, r = ours_fwd(0)
    def _bn_bwd(i):
        sums, dg, db = ops.bn_act_bwd_reduce(dz, ys[i % 3], m, r, gam, bet, 1, None, None, None)
        return ops.bn_act_bwd_apply(dz, ys[i % 3], m, r, gam, bet, sums, None, 1, None, True, None)
    t = timeit(_bn_bwd)
    yr = [y.clone().requires_grad_(True) for y in ys]

    def ref_bwd(i):
        z = torch.nn.functional.silu(torch.nn.functional.batch_norm(yr[i % 3], None, None, gam, bet, True, 0.1, 1e-5))
        z.backward(dz)

    tr = timeit(ref_bwd, 5, graph=False) - timeit(ref_fwd, 5, graph=False)
    report(f"BN(train)+SiLU bwd ({nX}x{C})", t, tr, 4 * C * nX * 5, "5 passes by design; ref = fwd+bwd minus fwd")
    del ys, dz, yr

# ------------------------------------------ Linear + BatchNorm statistics in one pass (TMA + tcgen05)
torch.backends.cuda.matmul.allow_tf32 = True
for K in (384, 128):
    xl = [torch.randn(nX, K, device=dev, generator=gen) for _ in range(3)]
    wl = torch.randn(128, K, device=dev, generator=gen) / K ** 0.5
    bl = torch.randn(128, device=dev, generator=gen)
    ops = torch.ops.pygho_b200
    t = timeit(lambda i: ops.linear_stats(xl[i % 3], wl, bl, 1e-5, 0.1, None, None, None, None, False))

    def ref_lin(i):
        y = torch.nn.functional.linear(xl[i % 3], wl, bl)
        return y, y.mean(0), y.var(0, unbiased=False)

    def cublas_plus_stats(i):
        y = torch.nn.functional.linear(xl[i % 3], wl, bl)
        return y, ops.bn_stats(y, 1e-5, 0.1, None, None, None, None)

    tr = timeit(ref_lin, 5)
    t2 = timeit(cublas_plus_stats)
    report(f"Linear+BN stats ({nX}x{K}->128), one pass", t, tr, 4 * (nX * K + 128 * K + nX * 128),
           f"TF32; cuBLAS TF32 + bn_stats kernel: {t2:.1f} us")
    del xl

# ------------------------------------------------------------------ masked path (cfg3)
b, n = 128, 40
rng = np.random.default_rng(0)
sizes = torch.from_numpy(np.clip(np.rint(rng.normal(23.2, 4.5, b)), 9, n).astype(np.int64))
sizes[0] = n
ar = torch.arange(n)
mask = ((ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])).to(dev)
NM = 6   # 6 x 105 MB: well beyond the 126 MB L2 even with a non-LRU replacement policy
Ms = [torch.randn(b, n, n, d, device=dev, generator=gen) * mask.unsqueeze(-1) for _ in range(NM)]
MT = [MaskedTensor(m, mask, 0.0, True) for m in Ms]
torch.backends.cuda.matmul.allow_tf32 = True
full_bytes = 4 * d * b * 3 * n * n + b * n * n
valid_bytes = 4 * d * float((2 * sizes.double() ** 2).sum() + b * n * n) + b * n * n
from pygho_b200.backend import mamamm  # noqa: E402
for algo in (4, 2, 1, 0):
    os.environ["PYGHO_B200_MAMAMM_ALGO"] = str(algo)
    t = timeit(lambda i: mamamm(MT[i % NM], 2, MT[(i + 1) % NM], 1, mask))
    tr = timeit(lambda i: torch.matmul(Ms[i % NM].permute(3, 0, 1, 2), Ms[(i + 1) % NM].permute(3, 0, 1, 2)).permute(1, 2, 3, 0) * mask.unsqueeze(-1), 5)
    flops = 2 * d * float((sizes.double() ** 3).sum())
    report(f"mamamm algo {algo} ({ {4: 'exact fp32, TMA-fed smem ring, largest graph first', 2: 'tcgen05 tf32, persistent pipeline', 1: 'tcgen05 tf32, CTA per item', 0: 'fp32 simt'}[algo] }) b={b} n={n}", t, tr, valid_bytes,
           f"{full_bytes / 1e6:.0f} MB if pads were read; useful {flops / t / 1e6:.1f} TFLOP/s")
for aggr in ("sum", "max"):
    code = {"sum": 0, "max": 2}[aggr]
    v4 = [m.reshape(b * n, n, 1, d) for m in Ms]
    m4 = mask.reshape(b * n, n, 1)
    t = timeit(lambda i: torch.ops.pygho_b200.masked_pool(v4[i % NM], m4, 1, code))

    def ref_pool(i):
        x = Ms[i % NM]
        if aggr == "sum":
            return x.sum(2)
        r = x.masked_fill(~mask.unsqueeze(-1), float("-inf")).amax(2)
        return r.masked_fill(torch.isinf(r), 0)

    tr = timeit(ref_pool, 5)
    # masked-out positions are never read (the kernel tests the mask first): count the bytes
    # of valid positions only, like mamamm above; the SURVEY 8d figure with pads is in the note
    pool_valid = 4 * d * float((sizes.double() ** 2).sum() + b * n) + b * n * n
    report(f"masked pool {aggr} dim 2 ({b},{n},{n},{d})", t, tr, pool_valid,
           f"{(4 * d * (b * n * n + b * n) + b * n * n) / 1e6:.0f} MB if pads were read")

# ------------------------------------------------------------------ batch preparation (hodata.cu)
import time  # noqa: E402

from pygho_b200.hodata.MaData import to_dense_adj, to_dense_x  # noqa: E402
from pygho_b200.hodata.MaTupleSampler import spdsampler  # noqa: E402
from pygho_b200.hodata.SpTupleSampler import KhopSampler, graph_distances  # noqa: E402
from pygho_b200.hodata.synthetic import khop_tuples  # noqa: E402

nptr_d = torch.from_numpy(hb.node_ptr).to(dev)
sq = int((np.diff(hb.node_ptr) ** 2).sum())
t0 = time.perf_counter()
for g in range(64):                                   # host sampler of the generator, 64 graphs
    e0_, e1_ = int(hb.edge_ptr[g]), int(hb.edge_ptr[g + 1])
    khop_tuples(int(hb.node_ptr[g + 1] - hb.node_ptr[g]), hb.edge_index[:, e0_:e1_] - hb.node_ptr[g], 3)
cpu_us = (time.perf_counter() - t0) * 1e6 * (B / 64)
t = timeit(lambda i: graph_distances(ei, hb.node_ptr, 3), 8, graph=False)
report(f"hop-distance matrices, B={B} (bit-set BFS, 1 launch + torch glue)", t, None,
       16 * nA + 8 * 3 * (B + 1) + sq + 4 * N, "int64 edges in, u8 distances + counts out")
t = timeit(lambda i: KhopSampler(ei, hb.node_ptr, 3), 8, graph=False)
report(f"KhopSampler hop 3, B={B} -> {nX} tuples (2 launches + scan, 1 size read-back)", t, None,
       16 * nA + 2 * sq + 4 * N + 8 * (N + 1) + 24 * nX,
       f"host numpy sampler (1 core): {cpu_us / 1e3:.0f} ms per batch = {cpu_us / t:.0f}x")
t = timeit(lambda i: spdsampler(ei, hb.node_ptr, 5), 8, graph=False)
nmax_ = int(np.diff(hb.node_ptr).max())
report(f"spdsampler hop 5 -> ({B},{nmax_},{nmax_}) padded", t, None,
       16 * nA + 2 * sq + 9 * B * nmax_ * nmax_, "int64 features + mask out")
xcol = torch.from_numpy(hb.x).to(dev).unsqueeze(-1)
t = timeit(lambda i: to_dense_x(xcol, nptr_d, nmax_, B), 8, graph=False)
report(f"to_dense_x ({N},1) -> ({B},{nmax_},1)", t, None, 8 * N + 9 * B * nmax_, "host-overhead bound")
ea_d = torch.from_numpy(hb.edge_attr).to(dev)
eb_d = bvec[ei[0]]
t = timeit(lambda i: to_dense_adj(ei, eb_d, ea_d, nmax_, B, node_ptr=nptr_d), 8, graph=False)
report(f"to_dense_adj {nA} edges -> ({B},{nmax_},{nmax_})", t, None,
       32 * nA + 9 * B * nmax_ * nmax_, "fill pass + scatter pass")


if args.md:
    with open(args.md, "w") as f:
        f.write(f"# Op-level rooflines (B200, measured HBM peak {PEAK:.0f} GB/s)\n\n"
                f"`python profiles/run_ops.py` -- cfg5 batch: B={B} graphs, N={N}, nnzA={nA}, nnzX={nX}, d={d}; "
                f"cfg3: b={b}, n={n}. CUDA events around replays of a captured CUDA graph (device time; plan builders and host-syncing reference chains are timed eagerly), inputs rotated over {NSET} sets. "
                "`ref-gpu` = the reference's own ATen call chain executed on the same B200.\n\n"
                "| op | ours (us) | GB/s (algorithmic) | frac of peak | ref-gpu (us) | speed-up | alg. MB | note |\n"
                "|---|---:|---:|---:|---:|---:|---:|---|\n")
        for name, us, gbs, frac, ref, sp, mb, note in rows:
            f.write(f"| {name} | {us:.1f} | {gbs:.0f} | {100 * frac:.1f}% | "
                    f"{'' if not ref else f'{ref:.1f}'} | {'' if not sp else f'{sp:.1f}x'} | {mb:.1f} | {note} |\n")
