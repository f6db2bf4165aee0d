"""Staged (shared-memory first-operand rows) vs streaming segmented-reduce kernels on the keys
with many entries per row: ZINC 2-FWL key (B=1024), sr25 X.A (B=64), I2 key on sr25 (B=64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygho_b200.backend as B  # noqa: E402
from pygho_b200 import ops as OPS  # noqa: E402
from pygho_b200 import plans as P  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

dev = torch.device("cuda", 0)
ops = torch.ops.pygho_b200
CASES = [("zinc", "khop", "X___X___1___X___0", 1024), ("sr25", "khop", "X___X___1___A___0", 64),
         ("sr25", "i2", "X___X___2___A___0", 64), ("zinc", "khop", "X___X___1___A___0", 1024)]
ONLY = os.environ.get("CASE")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for ci, (shape, tuples, key, graphs) in enumerate(CASES):
    if ONLY is not None and int(ONLY) != ci:
        continue
    hb = make_batch(graphs, seed=0, tuples=tuples, shape=shape)
    ei, tid = torch.from_numpy(hb.edge_index).to(dev), torch.from_numpy(hb.tupleid).to(dev)
    _o0, o1, d1, o2, d2 = key.split("___")
    pick = lambda op: ei if op == "A" else tid  # noqa: E731
    acd = B.filterind(tid, *B.spspmm_ind(pick(o1), int(d1), pick(o2), int(d2)))
    n_out, n1, n2 = tid.shape[1], pick(o1).shape[1], pick(o2).shape[1]
    plan = P.plan_from_acd(acd, n_out, n1, n2)
    gen = torch.Generator(device=dev).manual_seed(0)
    sets = [(torch.randn((n1, 128), device=dev, generator=gen), torch.randn((n2, 128), device=dev, generator=gen),
             torch.randn((n_out, 128), device=dev, generator=gen)) for _ in range(3)]
    alg = 4 * 128 * (n1 + n2 + n_out) + 4 * (2 * plan.T + n_out + 1)
    print(f"== {shape} {tuples} {key} B={graphs}: rows {n_out}, T {plan.T} ({plan.T / n_out:.1f}/row), algorithmic {alg / 1e6:.0f} MB")
    i = [0]
    for which, first, second, rows in (("a", 0, 1, n_out), ("c", 2, 1, n1), ("d", 2, 0, n2)):
        t = plan.tiles(which)
        g = plan.group(which)

        def lean():
            s = sets[i[0] % 3]; i[0] += 1
            return ops.seg_gmr(s[first], g.first, None, s[second], g.second, g.rowptr, rows, 0)
        us0 = timeit(lean)
        msg = f"   grouping {which}: streaming {us0:7.1f} us ({alg / us0 / 1e3:6.0f} GB/s)"
        if t is not None:
            def staged():
                s = sets[i[0] % 3]; i[0] += 1
                return ops.seg_gmr_staged(s[first], g.first, None, s[second], g.second, g.rowptr, rows, 0,
                                          t[0], t[1], plan.STAGE_ROWS_PER_TILE, plan.STAGE_MAX_ROWS)
            us1 = timeit(staged)
            same = torch.equal(lean.__call__() if False else ops.seg_gmr(sets[0][first], g.first, None, sets[0][second], g.second, g.rowptr, rows, 0),
                               ops.seg_gmr_staged(sets[0][first], g.first, None, sets[0][second], g.second, g.rowptr, rows, 0,
                                                  t[0], t[1], plan.STAGE_ROWS_PER_TILE, plan.STAGE_MAX_ROWS))
            fit = float((t[1] <= plan.STAGE_MAX_ROWS).float().mean())
            msg += f" | staged {us1:7.1f} us ({alg / us1 / 1e3:6.0f} GB/s, {us0 / us1:.2f}x, tiles staged {fit:.2f}, mean range {float(t[1].float().mean()):.1f} rows, identical {same})"
        else:
            msg += " | staged: not selected (low reuse or ranges too long)"
        print(msg)
