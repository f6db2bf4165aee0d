"""cProfile of the host side of eager SSWL+ training steps (B=1024): where do the ~14 ms of
Python / dispatcher time per step go?   python profiles/host_profile.py > gpurun_out/host_profile.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples.zinc_models import SpModel  # noqa: E402
from pygho_b200.dist import FlatGradBucket  # noqa: E402
from pygho_b200.hodata.device import prefetch_plans, sp_datadict  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402
from pygho_b200.honn.SpOperator import parse_precomputekey  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=6, hiddim=128).to(dev)
keys = parse_precomputekey(model)
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)
dd = sp_datadict(make_batch(1024, seed=0), dev, keys)
prefetch_plans(dd, keys, embeddings={"x": 32, "A": 16, "X": 16})


def step():
    bucket.zero()
    loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
host = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
print(f"host enqueue time per step: {host:.2f} ms")
# forward / backward / optimizer split
for name, fn in (("forward+loss", lambda: torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))),):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = [fn() for _ in range(N)]
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e3:.2f} ms host")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for o in outs:
        o.backward()
    print(f"backward: {(time.perf_counter() - t0) / N * 1e3:.2f} ms host")
    torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    opt.step()
print(f"optimizer: {(time.perf_counter() - t0) / N * 1e3:.2f} ms host")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
