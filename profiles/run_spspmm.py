"""The fused spspmm forward/backward kernels alone on a B=1024 batch (d=128), rotating
operand sets larger than L2; for `ncu --set full -k regex:seg_gmr`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygho_b200 import plans as P  # noqa: E402
from pygho_b200.hodata.device import sp_datadict  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

B = int(os.environ.get("BATCH", "1024"))
KEY = os.environ.get("KEY", "X___X___1___A___0")
ITERS = int(os.environ.get("ITERS", "12"))
dev = torch.device("cuda", 0)
dd = sp_datadict(make_batch(B, seed=0), dev, [KEY])
op1, op2 = KEY.split("___")[1], KEY.split("___")[3]
n1, n2 = dd[op1].nnz, dd[op2].nnz
nX = dd["X"].nnz
plan = P.plan_from_acd(dd[KEY + "___acd"], nX, n1, n2)
ga, gc, gd = plan.group("a"), plan.group("c"), plan.group("d")
ops = torch.ops.pygho_b200
gen = torch.Generator(device=dev).manual_seed(0)
sets = [(torch.randn((n1, 128), device=dev, generator=gen), torch.randn((n2, 128), device=dev, generator=gen),
         torch.randn((nX, 128), device=dev, generator=gen)) for _ in range(4)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda a, b, g: ops.seg_gmr(a, ga.first, None, b, ga.second, ga.rowptr, nX, 0)),
                 ("bwd_a", lambda a, b, g: ops.seg_gmr(g, gc.first, None, b, gc.second, gc.rowptr, n1, 0)),
                 ("bwd_b", lambda a, b, g: ops.seg_gmr(g, gd.first, None, a, gd.second, gd.rowptr, n2, 0))):
    for i in range(3):
        fn(*sets[i % 4])
    torch.cuda.synchronize()
    e0.record()
    for i in range(ITERS):
        fn(*sets[i % 4])
    e1.record()
    torch.cuda.synchronize()
    print(f"{KEY} {name}: {e0.elapsed_time(e1) * 1e3 / ITERS:.1f} us/launch  rows={nX} T={plan.T} nA={n1} nB={n2}")
