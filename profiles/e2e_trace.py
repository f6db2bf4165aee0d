"""Per-step trace of the end-to-end loop of bench.py (pinned host datadict -> H2D -> CSR
regroup -> train step -> loss D2H): host time of every phase, allocator activity and the
device time of the step, to find where e2e loses time against the resident-batch loop.

    python profiles/e2e_trace.py [--steps 30] [--mode prefetch|sequential|pipelined]
"""
import argparse
import os
import sys
import time

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF",
                      "expandable_segments:True,roundup_power2_divisions:8")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from examples.zinc_models import SpModel  # noqa: E402
from pygho_b200.dist import FlatGradBucket  # noqa: E402
from pygho_b200.hodata.device import (DevicePrefetcher, attach_host_plans,  # noqa: E402
                                      prefetch_plans, sp_datadict)
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402
from pygho_b200.honn.SpOperator import parse_precomputekey  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--mode", default="prefetch")
ap.add_argument("--sync", action="store_true", help="full device sync at the end of every step")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=6, hiddim=128).to(dev)
keys = parse_precomputekey(model)
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
hbs = [make_batch(1024, seed=i) for i in range(3)]
pinned = {}
for hb in hbs:
    dd = sp_datadict(hb, dev, keys, pinned)
    attach_host_plans(hb, dd, keys)
for hb in hbs:
    for k, v in hb.plans.items():
        pinned[id(v)] = torch.from_numpy(v).pin_memory()


def train_step(dd):
    bucket.zero()
    loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
    loss.backward()
    opt.step()
    return loss


for hb in hbs:
    train_step(sp_datadict(hb, dev, keys, pinned))
torch.cuda.synchronize()
feeder = DevicePrefetcher(hbs, dev, keys, pinned) if args.mode != "sequential" else None
now = time.perf_counter
stat = lambda: torch.cuda.memory_stats(dev)  # noqa: E731
lines = []
for i in range(args.steps):
    s0 = stat()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = now()
    if feeder is None:
        dd = sp_datadict(hbs[i % 3], dev, keys, pinned)
        prefetch_plans(dd, keys)
    else:
        dd = feeder.get()
    t1 = now()
    e0.record()
    loss = train_step(dd)
    e1.record()
    t2 = now()
    if feeder is not None:
        feeder.advance()
    t3 = now()
    val = float(loss.item())
    t4 = now()
    if args.sync:
        torch.cuda.synchronize()
    t5 = now()
    s1 = stat()
    lines.append(f"step {i:2d} get {1e3*(t1-t0):6.2f} launch {1e3*(t2-t1):6.2f} advance {1e3*(t3-t2):6.2f} "
                 f"item {1e3*(t4-t3):6.2f} tail {1e3*(t5-t4):6.2f} total {1e3*(t5-t0):6.2f} | dev step "
                 f"{e0.elapsed_time(e1):6.2f} | segs +{s1['num_device_alloc']-s0['num_device_alloc']} "
                 f"-{s1['num_device_free']-s0['num_device_free']} reserved {s1['reserved_bytes.all.current']/2**30:.2f} GB "
                 f"allocs {s1['allocation.all.allocated']-s0['allocation.all.allocated']}")
print("\n".join(lines))
