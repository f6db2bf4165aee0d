import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from examples.zinc_models import SpModel
from pygho_b200.dist import FlatGradBucket
from pygho_b200.hodata.device import sp_datadict, attach_host_plans
from pygho_b200.hodata.synthetic import make_batch
from pygho_b200.honn.SpOperator import parse_precomputekey
from pygho_b200 import plans as P
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=6, hiddim=128).to(dev)
keys = parse_precomputekey(model)
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
hb = make_batch(1024, seed=0)
pinned = {}
dd = sp_datadict(hb, dev, keys, pinned)
attach_host_plans(hb, dd, keys)
for k, v in hb.plans.items():
    pinned[id(v)] = torch.from_numpy(v).pin_memory()
def step(dd):
    bucket.zero()
    loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
    loss.backward(); opt.step(); return loss
def T(): torch.cuda.synchronize(); return time.perf_counter()
for _ in range(3): step(dd)
for rep in range(3):
    t0 = T(); dd2 = sp_datadict(hb, dev, keys, pinned); t1 = T()
    for k in keys:
        a = dd2[k + "___acd"]
        n1 = dd2["X"].nnz if k.split("___")[1] == "X" else dd2["A"].nnz
        n2 = dd2["X"].nnz if k.split("___")[3] == "X" else dd2["A"].nnz
        P.plan_from_acd(a, dd2["X"].nnz, n1, n2).prefetch()
    t2 = T(); l = step(dd2); t3 = T(); l.item(); t4 = T()
    # host-only launch time of a step (no sync inside)
    h0 = time.perf_counter(); l = step(dd2); h1 = time.perf_counter(); torch.cuda.synchronize(); h2 = time.perf_counter()
    print(f"h2d+wrap {1e3*(t1-t0):.2f} ms  plans {1e3*(t2-t1):.2f} ms  step(sync) {1e3*(t3-t2):.2f} ms  item {1e3*(t4-t3):.2f}  | host launch time of a step {1e3*(h1-h0):.2f} ms, total {1e3*(h2-h0):.2f}")

import gc
hbs = [hb, make_batch(1024, seed=1), make_batch(1024, seed=2)]
for h in hbs[1:]:
    d_ = sp_datadict(h, dev, keys, pinned); attach_host_plans(h, d_, keys)
    for k, v in h.plans.items():
        pinned[id(v)] = torch.from_numpy(v).pin_memory()
    step(d_)
del d_
for mode in ("gc on", "gc off"):
    if mode == "gc off":
        gc.collect(); gc.disable()
    times = []
    st0 = torch.cuda.memory_stats()
    for i in range(12):
        t0 = T()
        dd2 = sp_datadict(hbs[i % 3], dev, keys, pinned)
        l = step(dd2); l.item()
        times.append(1e3 * (T() - t0))
    st1 = torch.cuda.memory_stats()
    print(mode, " ".join(f"{t:.1f}" for t in times), "| cudaMalloc segs", st1["num_device_alloc"] - st0["num_device_alloc"],
          "frees", st1["num_device_free"] - st0["num_device_free"], "retries", st1["num_alloc_retries"] - st0["num_alloc_retries"],
          "reserved GB", round(st1["reserved_bytes.all.current"] / 2**30, 1))
